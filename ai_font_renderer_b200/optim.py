"""Fused AdamW for the renderer: one HBM pass per tensor (p, g, m, v read; p, m, v + the bf16
shadow of fc_output.weight written) in libafr_sm100.so instead of torch's multi-pass foreach
implementation. Reference: optim.AdamW(model.parameters(), lr, weight_decay, betas=(0.9, 0.99))
at model.py:273 and optimizer.step() at model.py:310.

It is a torch.optim.Optimizer, so ReduceLROnPlateau (model.py:276-278,337) drives its
param_groups[0]['lr'] unchanged, and its state uses torch's own keys (step / exp_avg /
exp_avg_sq).
"""
from __future__ import annotations

import torch

from .renderer import AttentionFontRenderer, _stream_ptr


class FusedAdamW(torch.optim.Optimizer):
    def __init__(self, model: AttentionFontRenderer, lr: float = 1e-3, betas=(0.9, 0.999),
                 eps: float = 1e-8, weight_decay: float = 1e-2, fuse_wgrad: bool = True,
                 overlap_dgrad: bool = False, background: bool = False, bg_chunks: int = 1,
                 bg_ctas: int = 0, bg_stages: int = 0, bg_after_dgrad: bool = True):
        if not isinstance(model, AttentionFontRenderer):
            raise TypeError("FusedAdamW is bound to an ai_font_renderer_b200.AttentionFontRenderer")
        if lr < 0 or eps < 0 or weight_decay < 0 or not (0 <= betas[0] < 1) or not (0 <= betas[1] < 1):
            raise ValueError("invalid AdamW hyper-parameter")
        self.model = model
        # single GPU: training.backward_and_step folds the step of fc_output.weight into the
        # wgrad GEMM (wgrad_step_rows); False keeps the two-kernel form (gradient materialised)
        self.fuse_wgrad = fuse_wgrad
        # ... and, optionally, runs the dgrad GEMM on a second stream under it with half-an-SM
        # footprints (afr_set_coresident). Measured on B200: not a win -- two operand stages and
        # two 128-column accumulators per kernel cost the AdamW GEMM more (0.72 -> 0.90 ms) than the
        # overlap can return (dgrad is 0.22 ms) -- so it is off by default.
        self.overlap_dgrad = overlap_dgrad
        # single GPU: the step of fc_output.weight as a BACKGROUND sweep (afr_adamw_rows_bg) on a
        # second stream, chunk k starting as soon as wgrad chunk k is done: the HBM-bound sweep
        # then runs under the rest of backward (dgrad GEMM, front-end backward) and the next
        # step's front-end forward instead of holding the GPU for itself. Takes precedence over
        # fuse_wgrad.
        self.background = background
        self.bg_chunks, self.bg_ctas, self.bg_stages = bg_chunks, bg_ctas, bg_stages
        self.bg_after_dgrad = bg_after_dgrad
        super().__init__(model._all_params(), dict(lr=lr, betas=betas, eps=eps,
                                                   weight_decay=weight_decay))

    def _ensure_state(self):
        params = self.model._ordered_params()
        for p in self.model._all_params():
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0)
                st["exp_avg"] = torch.zeros_like(p)
                st["exp_avg_sq"] = torch.zeros_like(p)
        return params

    def _bind(self):
        params = self._ensure_state()
        model = self.model
        ctx = model._context(1, training=True)
        ctx.bind_grads(model._param_grads())
        ctx.bind_adam_state([self.state[p]["exp_avg"] for p in params],
                            [self.state[p]["exp_avg_sq"] for p in params])
        if model.n_fonts > 0:       # thirteenth tensor: font_embedding.weight (config 3)
            fw = model.font_embedding.weight
            model._font_state = (self.state[fw]["exp_avg"], self.state[fw]["exp_avg_sq"])
            model._bind_fonts(ctx, model._font_ids.numel() if model._font_ids is not None else 0)
        return ctx, params

    def _hyper(self):
        g = self.param_groups[0]
        return (float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
                float(g["weight_decay"]))

    def _next_step(self, params) -> int:
        return int(self.state[params[0]]["step"].item()) + 1

    def _advance(self, params):
        for p in self.model._all_params():
            self.state[p]["step"] += 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        ctx, params = self._bind()
        t = self._next_step(params)
        ctx.check(ctx.lib.afr_adamw_step(ctx.handle, *self._hyper(), t, _stream_ptr(ctx.device)))
        self._advance(params)
        return loss

    # Row-bucketed form used by the data-parallel trainer: bucket b's all-reduce overlaps the
    # AdamW sweep of bucket b-1. Call begin_step(), any number of step_rows(), step_small(), end_step().
    @torch.no_grad()
    def begin_step(self):
        self._bucket = self._bind()
        return self._next_step(self._bucket[1])

    @torch.no_grad()
    def step_rows(self, t: int, row_begin: int, row_end: int):
        ctx, _ = self._bucket
        ctx.check(ctx.lib.afr_adamw_rows(ctx.handle, *self._hyper(), t, row_begin, row_end,
                                         _stream_ptr(ctx.device)))

    @torch.no_grad()
    def step_rows_bg(self, t: int, row_begin: int, row_end: int, ctas: int = 0, stages: int = 0,
                     grad_ptr: int = 0):
        """step_rows as the small-footprint background kernel (afr_adamw_rows_bg) on the current
        stream: meant for a side stream, where it shares the SMs with the compute kernels."""
        ctx, _ = self._bucket
        ctx.check(ctx.lib.afr_adamw_rows_bg(ctx.handle, *self._hyper(), t, row_begin, row_end,
                                            grad_ptr or None, ctas, stages, _stream_ptr(ctx.device)))

    @torch.no_grad()
    def wgrad_step_rows(self, t: int, row_begin: int, row_end: int):
        """Backward of fc_output w.r.t. weight / bias rows [row_begin, row_end) and the AdamW step
        of those weight rows in ONE kernel (afr_train_wgrad_adamw): the gradient stays in tensor
        memory, fc_output.weight.grad is not written. Needs a preceding fused_forward_loss."""
        ctx, _ = self._bucket
        ctx.check(ctx.lib.afr_train_wgrad_adamw(ctx.handle, *self._hyper(), t, row_begin, row_end,
                                                _stream_ptr(ctx.device)))

    @torch.no_grad()
    def bias_grad_rows(self, row_begin: int, row_end: int):
        """fc_output.bias.grad for the rows wgrad_step_rows handled (afr_train_bgrad)."""
        ctx, _ = self._bucket
        ctx.check(ctx.lib.afr_train_bgrad(ctx.handle, row_begin, row_end, _stream_ptr(ctx.device)))

    @torch.no_grad()
    def step_rows_gather(self, t: int, row_begin: int, row_end: int, peer_grads, peer_shadows,
                         world: int, ctas: int):
        """Row-sharded data parallel (training.PeerLink): AdamW on the owned rows with the
        gradient summed straight out of the peers' dW buffers and the bf16 result stored into
        every peer's inactive shadow copy, over NVLink, in one kernel on `ctas` SMs."""
        ctx, _ = self._bucket
        ctx.check(ctx.lib.afr_adamw_rows_gather(ctx.handle, *self._hyper(), t, row_begin, row_end,
                                                peer_grads, peer_shadows, world, ctas,
                                                _stream_ptr(ctx.device)))

    @torch.no_grad()
    def step_rows_gather_nvls(self, t: int, row_begin: int, row_end: int, grad_mc: int, shadow_mc: int,
                              ctas: int):
        """step_rows_gather through NVSwitch multicast addresses (in-switch gradient sum, one
        multicast store of the bf16 rows)."""
        ctx, _ = self._bucket
        ctx.check(ctx.lib.afr_adamw_rows_gather_nvls(ctx.handle, *self._hyper(), t, row_begin, row_end,
                                                     grad_mc, shadow_mc, ctas, _stream_ptr(ctx.device)))

    @torch.no_grad()
    def step_rows_gather_any(self, t: int, row_begin: int, row_end: int, link, shadow_index: int, world: int):
        """Dispatch over the four forms of the gather / AdamW / broadcast kernel a PeerLink selects:
        peer loads or NVSwitch multicast, fp32 or bf16 gradient buffers."""
        ctx, _ = self._bucket
        st = _stream_ptr(ctx.device)
        args = (ctx.handle, *self._hyper(), t, row_begin, row_end)
        lib = ctx.lib
        if link.nvls:
            fn = lib.afr_adamw_rows_gather_nvls_bf16 if link.grad_bf16 else lib.afr_adamw_rows_gather_nvls
            ctx.check(fn(*args, link.grad_mc, link.shadow_mc[shadow_index], link.ctas, st))
        else:
            fn = lib.afr_adamw_rows_gather_bf16 if link.grad_bf16 else lib.afr_adamw_rows_gather
            ctx.check(fn(*args, link.grad_ptrs, link.shadow_ptrs[shadow_index], world, link.ctas, st))

    @torch.no_grad()
    def step_small(self, t: int):
        ctx, _ = self._bucket
        ctx.check(ctx.lib.afr_adamw_small(ctx.handle, *self._hyper(), t, _stream_ptr(ctx.device)))

    @torch.no_grad()
    def end_step(self):
        self._advance(self._bucket[1])
        self._bucket = None

    def zero_grad(self, set_to_none: bool = False):
        """The fused backward overwrites every gradient, so the trainer never needs this; it is
        kept with in-place semantics (the gradient buffers stay bound to the library)."""
        for p in self.model._all_params():
            if p.grad is not None:
                p.grad.zero_()
