"""Host-side mirror of the reference's model interface for the hot path.

`AttentionFontRenderer` here is a drop-in for the class at model.py:129-204: same constructor
argument, same attribute `max_length` (read by helpers.py:52,58), same sub-module and parameter
names in the same construction order -- hence the same state_dict keys/shapes (font_renderer.pth
round-trips) and, under the same torch.manual_seed, bit-identical initial weights. What differs
is where the arithmetic runs: forward / backward / optimizer call the sm_100a kernels of
libafr_sm100.so through the C ABI (include/afr_sm100.h). PyTorch only owns the memory.

There is no CPU path: a non-CUDA input or a missing library raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch
import torch.nn as nn

from . import _lib

# module constants of the reference (model.py:64-66,79-81)
SHEET_HEIGHT = 80
SHEET_WIDTH = 240
MAX_CHARS_PER_SHEET = 100
EMBEDDING_DIM = 32
DROPOUT_RATE = 0.2
NUM_ATTENTION_HEADS = 4
FC1_WIDTH = 64
VOCAB = 128


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _Context:
    """Owns one afr_ctx and the tensors whose raw pointers are bound into it."""

    def __init__(self, model: "AttentionFontRenderer", device: torch.device, max_batch: int,
                 training: bool):
        self.lib = _lib.load()
        self.device = device
        self.max_batch = max_batch
        self.training = training
        cfg = _lib.AfrConfig(
            device=device.index if device.index is not None else torch.cuda.current_device(),
            vocab=model.embedding.num_embeddings, max_length=model.max_length,
            embed_dim=model.embedding_dim, num_heads=model.num_heads, hidden=model.fc1_width,
            sheet_h=model.sheet_height, sheet_w=model.sheet_width, max_batch=max_batch,
            training=1 if training else 0)
        handle = C.c_void_p()
        _lib.check(self.lib.afr_create(C.byref(cfg), C.byref(handle)))
        self.handle = handle
        self.param_ptrs = None
        self.grad_ptrs = None
        self.state_ptrs = None
        self.shadow_version = None
        self.keepalive = {}

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle.value:
            self.lib.afr_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        _lib.check(rc, self.handle)

    @staticmethod
    def _pack(tensors) -> _lib.AfrTensors:
        t = _lib.AfrTensors()
        for name, tensor in zip(_lib.TENSOR_FIELDS, tensors):
            setattr(t, name, tensor.data_ptr())
        return t

    def bind_params(self, params):
        ptrs = tuple(p.data_ptr() for p in params)
        if ptrs != self.param_ptrs:
            self.check(self.lib.afr_bind_params(self.handle, C.byref(self._pack(params))))
            self.param_ptrs = ptrs
            self.shadow_version = None

    def bind_grads(self, grads):
        ptrs = tuple(g.data_ptr() for g in grads)
        if ptrs != self.grad_ptrs:
            self.check(self.lib.afr_bind_grads(self.handle, C.byref(self._pack(grads))))
            self.grad_ptrs = ptrs
            self.keepalive["grads"] = list(grads)

    def bind_adam_state(self, exp_avg, exp_avg_sq):
        ptrs = tuple(t.data_ptr() for t in list(exp_avg) + list(exp_avg_sq))
        if ptrs != self.state_ptrs:
            self.check(self.lib.afr_bind_adam_state(self.handle, C.byref(self._pack(exp_avg)),
                                                    C.byref(self._pack(exp_avg_sq))))
            self.state_ptrs = ptrs
            self.keepalive["adam"] = (list(exp_avg), list(exp_avg_sq))

    def workspace(self, which: int):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        self.check(self.lib.afr_workspace_ptr(self.handle, which, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    def workspace_tensor(self, which: int, shape, dtype) -> torch.Tensor:
        """Copy of a private workspace as a torch tensor (tests / diagnostics)."""
        out = torch.empty(shape, dtype=dtype, device=self.device)
        self.check(self.lib.afr_workspace_copy(self.handle, which, out.data_ptr(),
                                               out.numel() * out.element_size(),
                                               _stream_ptr(self.device)))
        return out

    def launch_count(self) -> int:
        return int(self.lib.afr_launch_count(self.handle))


class _GenericTrainFn(torch.autograd.Function):
    """Keeps forward() differentiable for an arbitrary PyTorch loss (unfused-loss path)."""

    @staticmethod
    def forward(fctx, model, tokens, drop, *params):
        ctx = model._context(tokens.shape[0], training=True)
        B, S = tokens.shape[0], min(tokens.shape[1], model.max_length)
        out = torch.empty((B, model.sheet_height, model.sheet_width), dtype=torch.float32,
                          device=tokens.device)
        ctx.check(ctx.lib.afr_forward_train(ctx.handle, tokens.data_ptr(), tokens.stride(0), B, S,
                                            C.byref(drop), out.data_ptr(),
                                            _stream_ptr(tokens.device)))
        fctx.model = model
        fctx.tokens = tokens      # the library re-reads the tokens in backward
        fctx.drop = drop
        return out

    @staticmethod
    def backward(fctx, dsheet):
        model = fctx.model
        ctx = model._ctx
        dsheet = dsheet.contiguous().float()
        scratch = model._scratch_grads()
        ctx.bind_grads(scratch)
        ctx.check(ctx.lib.afr_backward(ctx.handle, dsheet.data_ptr(), _stream_ptr(dsheet.device)))
        grads = tuple(g.clone() for g in scratch)
        model._rebind_param_grads()
        return (None, None, None) + grads


class AttentionFontRenderer(nn.Module):
    """Reference: class AttentionFontRenderer, model.py:129-204."""

    def __init__(self, max_length: int = MAX_CHARS_PER_SHEET, sheet_height: int = SHEET_HEIGHT,
                 sheet_width: int = SHEET_WIDTH, vocab: int = VOCAB, embedding_dim: int = EMBEDDING_DIM,
                 num_heads: int = NUM_ATTENTION_HEADS, fc1_width: int = FC1_WIDTH, n_fonts: int = 0):
        """The reference's constructor takes max_length only (model.py:130); its other sizes are
        module constants (model.py:64-66,79-81,148). They are arguments here so that the scaled
        workloads of BASELINE.json can be built: sheet size, vocabulary (config 5), and the
        width of the net -- embedding_dim / num_heads / fc1_width other than 32 / 4 / 64 select
        the GEMM-based front-end (csrc/afr_wide.cu; config 4: 128 / 8 / 128). n_fonts > 0 adds
        multi-font conditioning (config 3): a thirteenth parameter `font_embedding.weight`
        [n_fonts, embedding_dim] whose row is added to every token embedding of a sample before
        the embedding dropout; it is constructed LAST, so the first twelve tensors keep the
        reference's initial values under the same seed."""
        super().__init__()
        self.max_length = max_length
        self.sheet_height = sheet_height
        self.sheet_width = sheet_width
        self.embedding_dim = embedding_dim
        self.num_heads = num_heads
        self.fc1_width = fc1_width
        # Same sub-modules, same order as model.py:136-152 => same RNG draws, same state_dict.
        self.embedding = nn.Embedding(vocab, self.embedding_dim)
        self.embedding_dropout = nn.Dropout(DROPOUT_RATE)
        self.positional_encoding = nn.Parameter(torch.zeros(max_length, self.embedding_dim))
        nn.init.normal_(self.positional_encoding, mean=0, std=0.02)
        self.attention = nn.MultiheadAttention(embed_dim=self.embedding_dim,
                                               num_heads=num_heads, dropout=DROPOUT_RATE)
        self.layer_norm = nn.LayerNorm(self.embedding_dim)
        self.fc1 = nn.Linear(self.embedding_dim, fc1_width)
        self.dropout1 = nn.Dropout(DROPOUT_RATE + 0.05)
        self.fc_output = nn.Linear(fc1_width * max_length, sheet_height * sheet_width)
        self.n_fonts = n_fonts
        if n_fonts > 0:
            self.font_embedding = nn.Embedding(n_fonts, self.embedding_dim)
        self._font_ids: Optional[torch.Tensor] = None
        # dropout generator of the fused kernels: keyed by torch's seed, advanced every train step
        self.dropout_seed = int(torch.initial_seed()) & 0xFFFFFFFFFFFFFFFF
        self.dropout_step = 0
        self._ctx: Optional[_Context] = None
        self._scratch: Optional[list] = None
        self._side: Optional[torch.cuda.Stream] = None
        self._shadow: Optional[list] = None     # caller-owned bf16 copies of fc_output.weight (DP)

    # ------------------------------------------------------------------ plumbing
    def side_stream(self) -> torch.cuda.Stream:
        """Second CUDA stream of the training step: the HBM-bound AdamW sweep over
        fc_output.weight runs here, concurrently with the GEMMs / front-end backward that the
        compute (current) stream runs (training.backward_and_step)."""
        dev = self.fc_output.weight.device
        if self._side is None or self._side.device != dev:
            self._side = torch.cuda.Stream(device=dev)
        return self._side

    def _ordered_params(self):
        sd = dict(self.named_parameters())
        return [sd[k] for k in _lib.STATE_DICT_KEYS]

    def _all_params(self):
        """The reference's 12 tensors, then font_embedding.weight if the model has one."""
        ps = self._ordered_params()
        if self.n_fonts > 0:
            ps.append(self.font_embedding.weight)
        return ps

    def set_fonts(self, font_ids: Optional[torch.Tensor]):
        """Font of every sample of the NEXT forward call (int tensor [B]); None = no conditioning."""
        if font_ids is None:
            self._font_ids = None
        else:
            if self.n_fonts <= 0:
                raise ValueError("this model has no font_embedding (construct it with n_fonts > 0)")
            dev = self.fc_output.weight.device
            self._font_ids = font_ids.to(device=dev, dtype=torch.int32).contiguous()

    def _bind_fonts(self, c: "_Context", batch: int):
        """(Re)bind the font table / its gradient and Adam state when they exist, and hand the
        library the font ids of the coming forward."""
        if self.n_fonts <= 0:
            return
        w = self.font_embedding.weight
        st = getattr(self, "_font_state", None)
        has_grad = w.grad is not None
        ptrs = (w.data_ptr(), w.grad.data_ptr() if has_grad else 0,
                st[0].data_ptr() if (st and has_grad) else 0, st[1].data_ptr() if (st and has_grad) else 0)
        if getattr(c, "font_bound", None) != ptrs:
            c.check(c.lib.afr_bind_font_embedding(c.handle, self.n_fonts, ptrs[0], ptrs[1] or None,
                                                  ptrs[2] or None, ptrs[3] or None))
            c.font_bound = ptrs
        ids = self._font_ids
        if ids is not None and ids.numel() != batch:
            raise ValueError(f"font ids for {ids.numel()} samples, batch of {batch}")
        c.check(c.lib.afr_set_font_ids(c.handle, ids.data_ptr() if ids is not None else None))
        c.keepalive["font_ids"] = ids

    def _context(self, batch: int, training: bool) -> _Context:
        params = self._ordered_params()
        dev = params[0].device
        if dev.type != "cuda":
            raise RuntimeError("ai_font_renderer_b200 runs on a B200 only: move the model to CUDA "
                               "(there is no CPU fallback)")
        for p in params:
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("parameters must be contiguous fp32 (checkpoint contract)")
        c = self._ctx
        if (c is None or c.device != dev or c.max_batch < batch or (training and not c.training)):
            cap = 256
            while cap < batch:
                cap *= 2
            if c is not None:
                cap = max(cap, c.max_batch)
                training = training or c.training
                c.close()
            c = _Context(self, dev, cap, training)
            self._ctx = c
            if getattr(self, "_sm_limit", 0):
                c.check(c.lib.afr_set_sm_limit(c.handle, self._sm_limit))
            if getattr(self, "_smem_reserve", 0):
                c.check(c.lib.afr_set_smem_reserve(c.handle, self._smem_reserve))
            if self._shadow is not None and training and self._shadow[0].device == dev:
                c.check(c.lib.afr_bind_shadow(c.handle, self._shadow[0].data_ptr(), self._shadow[1].data_ptr()))
                c.shadow_bound = tuple(t.data_ptr() for t in self._shadow)
        c.bind_params(params)
        w = self.fc_output.weight
        if c.shadow_version != w._version:      # torch wrote the master weights in place
            c.check(c.lib.afr_sync_shadow(c.handle, _stream_ptr(dev)))
            c.shadow_version = w._version
        return c

    def own_shadow_copies(self, batch: int = 1, copies=None):
        """Data parallel with a row-sharded optimizer: the two bf16 copies of fc_output.weight the
        GEMMs read become torch tensors (so NCCL can all-gather updated rows into the inactive
        one). Returns [copy0, copy1]; `shadow_index()` tells which one the next forward reads."""
        c = self._context(batch, training=True)
        w = self.fc_output.weight
        if copies is not None:
            self._shadow = list(copies)          # e.g. symmetric-memory allocations (PeerLink)
        if self._shadow is None or self._shadow[0].device != w.device:
            self._shadow = [torch.empty(w.shape, dtype=torch.bfloat16, device=w.device) for _ in range(2)]
        if getattr(c, "shadow_bound", None) != tuple(t.data_ptr() for t in self._shadow):
            c.check(c.lib.afr_bind_shadow(c.handle, self._shadow[0].data_ptr(), self._shadow[1].data_ptr()))
            c.shadow_bound = tuple(t.data_ptr() for t in self._shadow)
            c.shadow_version = w._version       # rebuilt from the master at the next forward
        return self._shadow

    def defer_join(self, side: torch.cuda.Stream, commit: bool = True):
        """End of a step whose fc_output.weight update is still running on `side` (data parallel:
        the gather / AdamW / broadcast kernel; single GPU: the background AdamW sweep). Nothing
        needs those weights before the next fc_output GEMM, so the join (stream wait + activating
        the written copy when `commit`) is postponed until then; the next step's front-end kernel
        runs under it."""
        self._pending = (side, commit)

    def join_pending(self):
        pending = getattr(self, "_pending", None)
        if pending is not None:
            side, commit = pending
            torch.cuda.current_stream(side.device).wait_stream(side)
            if commit:
                self.shadow_commit()
            self._pending = None

    def set_smem_reserve(self, nbytes: int):
        """Shared memory per SM the wgrad / dgrad GEMMs and the front-end backward leave to the
        background AdamW sweep (afr_set_smem_reserve); 0 = none."""
        self._smem_reserve = nbytes
        if self._ctx is not None:
            self._ctx.check(self._ctx.lib.afr_set_smem_reserve(self._ctx.handle, nbytes))

    def set_sm_limit(self, sms: int):
        """Data parallel: leave `#SMs - sms` SMs to the collective's CTAs (0 = all SMs)."""
        self._sm_limit = sms
        if self._ctx is not None:
            self._ctx.check(self._ctx.lib.afr_set_sm_limit(self._ctx.handle, sms))

    def shadow_index(self) -> int:
        return int(self._ctx.lib.afr_shadow_index(self._ctx.handle))

    def shadow_commit(self):
        self._ctx.check(self._ctx.lib.afr_shadow_commit(self._ctx.handle))

    def _param_grads(self):
        """p.grad for the 12 tensors in state_dict order, allocated on first use. The ten small
        ones plus fc_output.bias are views into one flat buffer (`small_grad_flat`) so a
        data-parallel run reduces them with a single collective."""
        params = self._ordered_params()
        if any(p.grad is None for p in self._all_params()):
            small = [p for k, p in zip(_lib.STATE_DICT_KEYS, params) if k != "fc_output.weight"]
            if self.n_fonts > 0:
                small.append(self.font_embedding.weight)
            flat = torch.zeros(sum(p.numel() for p in small), dtype=torch.float32, device=params[0].device)
            off = 0
            for p in small:
                p.grad = flat[off:off + p.numel()].view_as(p)
                off += p.numel()
            self.small_grad_flat = flat
            w = self.fc_output.weight
            if w.grad is None:
                w.grad = torch.zeros_like(w)
        return [p.grad for p in params]

    def _rebind_param_grads(self):
        if self._ctx is not None and all(p.grad is not None for p in self._ordered_params()):
            self._ctx.bind_grads(self._param_grads())

    def _scratch_grads(self):
        if self._scratch is None or self._scratch[0].device != self.fc_output.weight.device:
            self._scratch = [torch.zeros_like(p) for p in self._ordered_params()]
        return self._scratch

    def _check_tokens(self, x: torch.Tensor) -> torch.Tensor:
        if x.dim() != 2:
            raise ValueError("expected tokens of shape [batch, seq_len]")
        if not x.is_cuda:
            raise RuntimeError("ai_font_renderer_b200 runs on a B200 only: tokens must be a CUDA "
                               "tensor (there is no CPU fallback)")
        if x.dtype != torch.int64:
            x = x.long()
        if x.stride(1) != 1:
            x = x.contiguous()
        return x

    def make_dropout(self, batch: int, seq: int, masks: Optional[Dict[str, torch.Tensor]] = None,
                     sample_offset: int = 0, enabled: bool = True) -> _lib.AfrDropout:
        d = _lib.AfrDropout()
        d.p_embed, d.p_attn, d.p_fc1 = (self.embedding_dropout.p, self.attention.dropout,
                                        self.dropout1.p)
        if not enabled:
            d.mode = 0
        elif masks is not None:
            d.mode = 2
            keep = {}
            for key, shape in (("embed", (batch, seq, self.embedding_dim)),
                               ("attn", (batch, self.num_heads, seq, seq)),
                               ("fc1", (batch, seq, self.fc1_width))):
                m = masks[key].to(device=self.fc_output.weight.device, dtype=torch.uint8).contiguous()
                if tuple(m.shape) != shape:
                    raise ValueError(f"mask '{key}' must have shape {shape}, got {tuple(m.shape)}")
                keep[key] = m
            d.mask_embed, d.mask_attn, d.mask_fc1 = (keep["embed"].data_ptr(),
                                                     keep["attn"].data_ptr(), keep["fc1"].data_ptr())
            d._keep = keep  # keep the tensors alive as long as the struct
        else:
            d.mode = 1
            d.seed = self.dropout_seed
            d.step = self.dropout_step
            d.sample_offset = sample_offset
        return d

    # ------------------------------------------------------------------ forward paths
    def _eval_forward(self, x: torch.Tensor, kind: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        self.join_pending()
        x = self._check_tokens(x)
        B, S = x.shape[0], min(x.shape[1], self.max_length)
        c = self._context(B, training=False)
        self._bind_fonts(c, B)
        if out is not None:
            want = torch.uint8 if kind == _lib.OUT_SHEET_U8 else torch.float32
            if (out.dtype != want or not out.is_contiguous() or out.device != x.device
                    or out.numel() != B * self.sheet_height * self.sheet_width):
                raise ValueError("out must be a contiguous tensor of B*H*W elements of the output dtype on the "
                                 "tokens' device")
        elif kind == _lib.OUT_SHEET_U8:
            out = torch.empty((B, self.sheet_height, self.sheet_width), dtype=torch.uint8, device=x.device)
        elif kind == _lib.OUT_SHEET_F32:
            out = torch.empty((B, self.sheet_height, self.sheet_width), dtype=torch.float32, device=x.device)
        else:
            out = torch.empty((B, self.sheet_height * self.sheet_width), dtype=torch.float32, device=x.device)
        c.check(c.lib.afr_forward_eval(c.handle, x.data_ptr(), x.stride(0), B, S, out.data_ptr(), kind,
                                       _stream_ptr(x.device)))
        return out

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """model.py:158-204. Eval mode (or no_grad) -> fused render; train mode -> differentiable."""
        if self.training:
            self.join_pending()
            x = self._check_tokens(x)
            S = min(x.shape[1], self.max_length)
            drop = self.make_dropout(x.shape[0], S)
            self.dropout_step += 1
            if torch.is_grad_enabled():
                return _GenericTrainFn.apply(self, x, drop, *self._ordered_params())
            c = self._context(x.shape[0], training=True)
            out = torch.empty((x.shape[0], self.sheet_height, self.sheet_width), dtype=torch.float32,
                              device=x.device)
            c.check(c.lib.afr_forward_train(c.handle, x.data_ptr(), x.stride(0), x.shape[0], S,
                                            C.byref(drop), out.data_ptr(), _stream_ptr(x.device)))
            return out
        return self._eval_forward(x, _lib.OUT_SHEET_F32)

    @torch.no_grad()
    def render_u8(self, x: torch.Tensor, out: Optional[torch.Tensor] = None,
                  font_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Eval forward + helpers.py:33 quantisation fused in the GEMM epilogue -> uint8 [B,H,W]
        (written into `out` if given: render.RenderPipeline reuses two device buffers)."""
        if font_ids is not None or self.n_fonts > 0:
            self.set_fonts(font_ids)
        return self._eval_forward(x, _lib.OUT_SHEET_U8, out=out)

    @torch.no_grad()
    def logits(self, x: torch.Tensor, font_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fc_output before the clamp (model.py:196), eval mode -> fp32 [B, H*W]."""
        if font_ids is not None or self.n_fonts > 0:
            self.set_fonts(font_ids)
        return self._eval_forward(x, _lib.OUT_LOGITS_F32)

    # ------------------------------------------------------------------ fused training step
    def fused_forward_loss(self, tokens: torch.Tensor, targets: torch.Tensor,
                           loss_count: Optional[float] = None,
                           masks: Optional[Dict[str, torch.Tensor]] = None, dropout: bool = True,
                           sample_offset: int = 0, loss_out: Optional[torch.Tensor] = None, marks=None,
                           font_ids: Optional[torch.Tensor] = None):
        """model.py:299 + 304-306 in one pass; leaves d(loss)/d(logits) inside the library.
        targets: uint8 [B,H,W] (k/255 grey levels) or fp32 [B,H,W]. Returns the device loss scalar
        (sum of squared errors / loss_count; loss_count defaults to B*H*W = mse_loss's mean)."""
        tokens = self._check_tokens(tokens)
        if font_ids is not None or self.n_fonts > 0:
            self.set_fonts(font_ids)
        B, S = tokens.shape[0], min(tokens.shape[1], self.max_length)
        if targets.shape[0] != B or targets.numel() != B * self.sheet_height * self.sheet_width:
            raise ValueError("targets must be [B, H, W]")
        if targets.dtype == torch.uint8:
            kind = _lib.TARGET_U8
        elif targets.dtype == torch.float32:
            kind = _lib.TARGET_F32
        else:
            raise ValueError("targets must be uint8 or float32")
        if not targets.is_cuda or not targets.is_contiguous():
            targets = targets.to(tokens.device).contiguous()
        c = self._context(B, training=True)
        c.bind_grads(self._param_grads())
        self._bind_fonts(c, B)
        drop = self.make_dropout(B, S, masks=masks, sample_offset=sample_offset,
                                 enabled=dropout and self.training)
        if loss_out is None:
            loss_out = torch.empty((), dtype=torch.float32, device=tokens.device)
        count = float(loss_count) if loss_count is not None else float(B * self.sheet_height * self.sheet_width)
        st = _stream_ptr(tokens.device)
        c.check(c.lib.afr_train_frontend(c.handle, tokens.data_ptr(), tokens.stride(0), B, S,
                                         C.byref(drop), st))
        if marks is not None:
            marks("frontend")        # bench.py: CUDA event between the front-end and the GEMM
        self.join_pending()          # fc_output weights of the previous data-parallel step
        c.check(c.lib.afr_train_loss(c.handle, targets.data_ptr(), kind, count, loss_out.data_ptr(), st))
        self._live = (tokens, targets, drop)   # the library reads tokens / masks again in backward
        if drop.mode == 1:
            self.dropout_step += 1
        return loss_out

    def fused_backward(self, row_buckets=None, on_bucket=None, wgrad_fn=None, marks=None):
        """loss.backward() (model.py:309): overwrites every p.grad. row_buckets: list of
        (row_begin, row_end) over fc_output's rows; on_bucket(i, begin, end) is called after the
        launches of each bucket so a data-parallel caller can start its all-reduce.
        wgrad_fn(begin, end), if given, replaces the wgrad launch of a bucket (the optimizer's
        wgrad_step_rows: gradient and AdamW step of those rows in one kernel)."""
        c = self._ctx
        st = _stream_ptr(self.fc_output.weight.device)
        P = self.sheet_height * self.sheet_width
        buckets = row_buckets or [(0, P)]
        for i, (lo, hi) in enumerate(buckets):
            if wgrad_fn is not None:
                wgrad_fn(lo, hi)
            else:
                c.check(c.lib.afr_train_wgrad(c.handle, lo, hi, st))
            if on_bucket is not None:
                on_bucket(i, lo, hi)
        if marks is None:
            c.check(c.lib.afr_train_dgrad(c.handle, st))
        else:                        # bench.py: CUDA events around the dgrad GEMM
            marks("dgrad_gemm_begin")
            c.check(c.lib.afr_train_dgrad_gemm(c.handle, st))
            marks("dgrad_gemm_end")
            c.check(c.lib.afr_train_frontend_backward(c.handle, st))

    def set_coresident(self, on: bool):
        """Single GPU: the next wgrad+AdamW GEMM and dgrad GEMM launch with half-an-SM footprints
        (afr_set_coresident) so that, enqueued on two streams, they share every SM."""
        c = self._ctx
        c.check(c.lib.afr_set_coresident(c.handle, 1 if on else 0))

    def wgrad_rows(self, row_begin: int, row_end: int):
        """d(loss)/d(fc_output.weight / .bias) for pixel rows [row_begin, row_end) (afr_train_wgrad)."""
        c = self._ctx
        c.check(c.lib.afr_train_wgrad(c.handle, row_begin, row_end, _stream_ptr(self.fc_output.weight.device)))

    def wgrad_rows_to(self, row_begin: int, row_end: int, dst: torch.Tensor, bf16: bool = False,
                      with_bias: bool = True):
        """afr_train_wgrad_to(_bf16): the gradient rows go to `dst` ([H*W, 64*max_length], fp32 or
        bf16 -- e.g. a peer-mapped buffer) instead of fc_output.weight.grad."""
        c = self._ctx
        K = self.fc_output.weight.shape[1]
        ptr = dst.data_ptr() + row_begin * K * dst.element_size()
        fn = c.lib.afr_train_wgrad_to_bf16 if bf16 else c.lib.afr_train_wgrad_to
        c.check(fn(c.handle, row_begin, row_end, ptr, 1 if with_bias else 0,
                   _stream_ptr(self.fc_output.weight.device)))

    def dgrad_gemm(self):
        """d(features) = d(logits) W on the current stream (first half of afr_train_dgrad)."""
        c = self._ctx
        c.check(c.lib.afr_train_dgrad_gemm(c.handle, _stream_ptr(self.fc_output.weight.device)))

    def frontend_backward(self):
        """fc1 / LayerNorm / attention / embedding backward into the ten small gradients."""
        c = self._ctx
        c.check(c.lib.afr_train_frontend_backward(c.handle, _stream_ptr(self.fc_output.weight.device)))

    def fused_train_step(self, tokens, targets, **kw) -> torch.Tensor:
        """optimizer.zero_grad(); loss = mse(model(x), t); loss.backward()  (model.py:292-309)."""
        loss = self.fused_forward_loss(tokens, targets, **kw)
        self.fused_backward()
        return loss

    def check_tokens_in_range(self):
        """Synchronises; raises IndexError like nn.Embedding would (model.py:167)."""
        if self._ctx is not None:
            c = self._ctx
            c.check(c.lib.afr_check_tokens(c.handle, _stream_ptr(c.device)))

    def kernel_launches(self) -> int:
        return self._ctx.launch_count() if self._ctx is not None else 0

    def state_dict(self, *args, **kwargs):
        """The optimizer step of fc_output.weight may still be running on the side stream
        (background sweep / data-parallel gather): join it before the weights are handed out."""
        self.join_pending()
        return super().state_dict(*args, **kwargs)

    # nn.Module hooks that can move / replace parameter storage
    def _apply(self, fn, recurse=True):
        self.join_pending()
        out = super()._apply(fn, recurse)
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None
        self._scratch = None
        self._shadow = None
        return out
