"""Training driver around the fused step (reference: train_attention_model, model.py:209-384).

The reference's per-batch body `zero_grad -> model(x) -> mse_loss -> backward -> step`
(model.py:291-311) becomes `fused_forward_loss -> fused_backward -> FusedAdamW`, the dataset
lives on the device (uint8 sheets, 2.9 GB for 150k samples) and batches are gathered there, but
everything that decides WHAT is computed is kept: the 80/20 random_split with seed 42, the shared
loader generator (so batch composition and order equal the reference's), unweighted epoch means,
ReduceLROnPlateau, early stopping with its shallow-copy "best state", the files written.

Data parallel (one process per GPU, torch.distributed): every rank walks the same global batch
order and takes a contiguous slice of each batch; the loss is normalised by the GLOBAL batch so
per-rank gradients just add. The optimizer over fc_output.weight is sharded by rows; its gradient
rows are summed and its updated bf16 rows broadcast inside ONE kernel per rank over NVLink peer
memory / NVSwitch multicast (PeerLink), overlapped with the rest of backward on a side stream; the
33 k small parameters use one NCCL all-reduce. NCCL reduce-scatter / all-gather is the fallback.
"""
from __future__ import annotations

import datetime
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.utils.data as tud

from .data import targets_as_u8
from .optim import FusedAdamW
from .render import render_strings


@dataclass
class TrainConfig:
    """Hyper-parameters; defaults are the reference's module constants (model.py:64-93)."""
    output_dir: str = ""
    num_epochs: int = 10000
    learning_rate: float = 0.001
    early_stopping_patience: int = 70
    validation_split: float = 0.2
    weight_decay: float = 0.0005
    embedding_dim: int = 32
    dropout_rate: float = 0.2
    num_attention_heads: int = 4
    scheduler_patience: int = 20
    scheduler_factor: float = 0.7
    min_learning_rate: float = 1e-6
    seed: int = 42
    sheet_height: int = 80
    sheet_width: int = 240
    max_chars_per_sheet: int = 100
    num_samples: int = 150000
    betas: Tuple[float, float] = (0.9, 0.99)
    test_strings: Sequence[str] = field(default_factory=list)
    test_font_ids: Optional[Sequence[int]] = None   # multi-font models: font of every test string
    # models with a font_embedding table (AttentionFontRenderer(n_fonts=N)): font of every SAMPLE
    sample_font_ids: Optional[torch.Tensor] = None
    render_every: int = 5
    grad_buckets: int = 8            # row buckets of fc_output.weight.grad (data parallel overlap)
    adam_buckets: int = 1            # single GPU: row buckets of the wgrad GEMM / AdamW sweep
    peer_memory: bool = True         # data parallel: NVLink peer-memory / NVLS optimizer step (PeerLink); False = NCCL
    max_steps: Optional[int] = None  # stop after this many optimizer steps (tests / smoke)
    # The reference keeps `model.state_dict().copy()` as its "best state" (model.py:344): a shallow
    # copy that aliases the live parameters, so both load_state_dict calls (model.py:365,370) are
    # no-ops and the saved weights are the LAST epoch's. False reproduces that (default: drop-in);
    # True keeps a real clone of the best epoch's weights and restores it.
    restore_best_weights: bool = False
    # single GPU: optimizer.step() of fc_output.weight as the background sweep on a second stream
    # (FusedAdamW(background=True)); False = inside the wgrad GEMM's epilogue
    background_adamw: bool = True
    # data parallel over NVSwitch: ship the fc_output.weight gradient rows as bf16 (half the NVLink
    # egress; summed in fp32 inside the switch; weights within 1e-3 of the fp32 exchange)
    dp_grad_bf16: bool = True
    quiet: bool = False


class _Indices(tud.Dataset):
    """Index-only stand-in for the TensorDataset: the DataLoader / random_split machinery of torch
    decides the order (model.py:239-266), the gather happens on the device."""

    def __init__(self, n: int):
        self.n = n

    def __len__(self):
        return self.n

    def __getitem__(self, i):
        return int(i)

    def __getitems__(self, idxs):
        return [int(i) for i in idxs]


def _world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of a global batch of n samples owned by `rank`."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def row_buckets(P: int, n: int) -> List[Tuple[int, int]]:
    """Split fc_output's P rows into <= n buckets whose bounds are multiples of 128."""
    n = max(1, n)
    tiles = (P + 127) // 128
    per = (tiles + n - 1) // n
    out, lo = [], 0
    while lo < P:
        hi = min(P, lo + per * 128)
        out.append((lo, hi))
        lo = hi
    return out


import ctypes as _C

_SMALL_GROUP = None
# ring depth of the background AdamW sweep: stages x 8 KB (+ the 1 KB the hardware reserves per CTA)
# of every SM's shared memory are left to it; 8 stages fit beside the front-end kernels' 2 x 76 KB
# (forward), 131 / 2 x 76 / 151 KB (backward head / attention / tail). With the sweep starting before the dgrad GEMM
# (bg_after_dgrad = False) 4 stages are the better trade: the GEMM's operand ring pays for them.
BG_DEFAULT_STAGES = 8
BG_DEFAULT_STAGES_BESIDE_DGRAD = 4


class PeerLink:
    """NVLink peer memory for the data-parallel step (torch symmetric memory): fc_output's gradient
    buffer and the two bf16 shadow copies of its weight are allocated symmetrically on every rank
    and mapped into every other rank's address space. With them the row-sharded optimizer needs no
    collective call for fc_output: one kernel per rank (afr_adamw_rows_gather) reads the owned rows
    of every rank's dW over NVLink, applies AdamW and stores the bf16 result into every rank's
    inactive shadow copy; the signal-pad barriers of the symmetric allocation order it against the
    wgrad GEMMs before and the next forward after. Construct on all ranks at the same time."""

    @staticmethod
    def default_ctas(world: int, grad_bf16: bool = False) -> int:
        """SMs given to the gather/AdamW/broadcast kernel (measured on 2/4/8 B200, bench.py
        --comm-ctas sweeps): the owned shard -- and with it the kernel's local HBM work -- shrinks
        with the number of ranks while its NVLink volume stays, so fewer SMs saturate it. With bf16
        gradient rows the link volume halves and the kernel hides under the rest of backward on
        half as many SMs (8 GPUs, profiles/r02: 24 CTAs 1.40 ms/step, 16: 1.29, 12: 1.27; with the final
        kernels 12: 1.204, 16 and 20 no better, 8: 1.28)."""
        if grad_bf16:
            return 32 if world <= 2 else (20 if world <= 4 else 12)
        return 48 if world <= 2 else (32 if world <= 4 else 24)

    def __init__(self, model, ctas: int = 0, group=None, inline: bool = False, nvls: bool = False,
                 grad_bf16: bool = False):
        """inline: run the gather/AdamW/broadcast kernel on the compute stream on ALL SMs right
        after the wgrad GEMM instead of on a side stream on `ctas` SMs next to the rest of
        backward: no SM is withheld from the GEMMs / front-end kernels, and the kernel itself is
        several times faster with the whole GPU's memory-level parallelism."""
        import torch.distributed._symmetric_memory as symm_mem
        group = group or dist.group.WORLD
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        self.inline = inline
        self.nvls = nvls
        # grad_bf16: the wgrad GEMM writes its gradient rows as bf16 into the peer-visible buffer
        # (half the NVLink egress per step: 215 instead of 430 MB per rank at 8 GPUs); the sum over
        # ranks is still accumulated in fp32 (in the gather kernel's registers / inside the switch)
        self.grad_bf16 = grad_bf16
        sms_all = torch.cuda.get_device_properties(model.fc_output.weight.device).multi_processor_count
        ctas = sms_all if inline else (ctas or self.default_ctas(self.world, grad_bf16))
        w = model.fc_output.weight
        dev = w.device
        self.ctas = ctas
        self.wgrad = symm_mem.empty(tuple(w.shape), dtype=torch.bfloat16 if grad_bf16 else torch.float32,
                                    device=dev)
        self.wgrad.zero_()
        self.h_grad = symm_mem.rendezvous(self.wgrad, group)
        self.shadow = [symm_mem.empty(tuple(w.shape), dtype=torch.bfloat16, device=dev) for _ in range(2)]
        self.h_shadow = [symm_mem.rendezvous(t, group) for t in self.shadow]
        if not grad_bf16:
            w.grad = self.wgrad                   # the wgrad GEMM writes the peer-visible buffer
        model._param_grads()
        model.own_shadow_copies(copies=self.shadow)
        arr = _C.c_void_p * self.world
        self.grad_ptrs = arr(*[int(p) for p in self.h_grad.buffer_ptrs])
        self.shadow_ptrs = [arr(*[int(p) for p in h.buffer_ptrs]) for h in self.h_shadow]
        if nvls:
            # NVSwitch multicast objects over the same allocations (NVLS): in-switch reduction of
            # the gradient rows, one multicast store for the bf16 rows
            self.grad_mc = int(self.h_grad.multicast_ptr or 0)
            self.shadow_mc = [int(h.multicast_ptr or 0) for h in self.h_shadow]
            if not self.grad_mc or not all(self.shadow_mc):
                raise RuntimeError("NVLS multicast is not available for the symmetric allocations on this system")
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        model.set_sm_limit(0 if inline else sms - ctas)   # side-stream mode: the gather kernel's SMs stay free
        model._peer_link = self

    @staticmethod
    def detach(model):
        """Undo a (partly) constructed link: plain gradient buffer and library-owned shadow copies
        again, all SMs back to the compute kernels."""
        model._peer_link = None
        w = model.fc_output.weight
        w.grad = torch.zeros_like(w)
        model._param_grads()
        model._rebind_param_grads()
        model._shadow = None
        model.own_shadow_copies()
        model.set_sm_limit(0)

    @staticmethod
    def nvls_available() -> bool:
        """True when symmetric allocations get an NVSwitch multicast mapping (collective: every rank
        must call it at the same point)."""
        try:
            import torch.distributed._symmetric_memory as symm_mem
            dev = torch.device("cuda", torch.cuda.current_device())
            probe = symm_mem.empty((1024,), dtype=torch.float32, device=dev)
            handle = symm_mem.rendezvous(probe, dist.group.WORLD)
            return bool(getattr(handle, "multicast_ptr", 0))
        except Exception:
            return False

    @staticmethod
    def available() -> bool:
        try:
            import torch.distributed._symmetric_memory as symm_mem  # noqa: F401
            return True
        except Exception:
            return False


def _small_group():
    """Second NCCL communicator (all ranks) for the 0.13 MB all-reduce of the small gradients."""
    global _SMALL_GROUP
    if _SMALL_GROUP is None:
        _SMALL_GROUP = dist.new_group()
    return _SMALL_GROUP


def owned_rows(P: int, rank: int, world: int) -> Tuple[int, int]:
    """Rows of fc_output.weight whose optimizer state (and fp32 master) rank `rank` owns."""
    if P % world != 0:
        raise ValueError(f"fc_output has {P} rows: not divisible by world size {world}")
    per = P // world
    return rank * per, (rank + 1) * per


def backward_and_step(model, optimizer, buckets, world: int, has_samples: bool = True, marks=None,
                      rank: Optional[int] = None):
    """loss.backward(); optimizer.step() (model.py:309-310) for one rank of a data-parallel job.

    world == 1, optimizer.fuse_wgrad (default): the wgrad GEMM with the AdamW step of
    fc_output.weight in its epilogue (afr_train_wgrad_adamw; fc_output.weight.grad is not
    materialised) with, if optimizer.overlap_dgrad, the dgrad GEMM co-resident on every SM from a
    second stream; then the front-end backward and the small-tensor AdamW.
    world == 1, fuse_wgrad off: wgrad, the AdamW sweep over fc_output.weight (right behind the
    gradient that is still partly in L2), dgrad, the front-end backward, the small-tensor AdamW.
    (Measured on B200, tools/overlap_probe.py: running the HBM-bound sweep on a second stream
    under the GEMMs / front-end backward buys nothing -- those kernels take the SM's whole
    shared-memory carveout, the sweep then runs with ~no L1 and both slow down by what the
    overlap saves.) The sweep writes the inactive copy of the bf16 shadow weights, so the dgrad
    GEMM behind it still reads the weights the forward used.

    world > 1: the optimizer over fc_output.weight (99.97 % of the parameters) is sharded by rows:
      compute stream : wgrad | dgrad, front-end backward | all-reduce + AdamW of the 33 k small
                       parameters | join
      side stream    : reduce-scatter of dW (each rank receives the sum of ITS rows, in place) ->
                       AdamW sweep over the owned rows only (1/world of the 3.7 GB sweep) ->
                       all-gather of the updated bf16 rows into the inactive shadow copy
    so the wire carries 491 MB fp32 + 246 MB bf16 per step instead of the 2 x 491 MB of an
    all-reduce, under dgrad and the front-end backward, and the HBM-bound sweep shrinks with the
    number of ranks. Each rank's fp32 master / Adam moments are current for its own rows only
    (Trainer.gather_master gathers them for checkpoints).
    `marks(label)`, if given, is called at phase boundaries with the stream to record on current
    (bench.py records CUDA events there): 'wgrad', 'dgrad', 'tail' on the compute stream,
    'adamw_begin' / 'adamw_end' around every sweep launch."""
    mark = marks or (lambda label: None)
    dev = model.fc_output.weight.device
    main = torch.cuda.current_stream(dev)
    model._param_grads()
    wgrad = model.fc_output.weight.grad
    t_step = optimizer.begin_step()
    P = wgrad.shape[0]

    if world == 1 and getattr(optimizer, "background", False):
        # compute stream : wgrad | dgrad GEMM | front-end backward | small AdamW | next front-end fwd
        # side stream    :                      AdamW sweep over fc_output.weight (background kernel:
        #                          one 128-thread CTA per SM beside whatever the compute stream runs)
        # The join is deferred to the next fc_output GEMM, so the next front-end forward runs under
        # the tail of the sweep too. The sweep writes the inactive bf16 copy: dgrad reads the old
        # one. bg_chunks > 1 splits wgrad / sweep into row chunks (the sweep of chunk k starts
        # under wgrad chunk k + 1); measured slower: the wgrad GEMM and the sweep are both
        # HBM-heavy and gain nothing from sharing the GPU (tools/overlap_probe2.py).
        model.join_pending()
        after_dgrad = getattr(optimizer, "bg_after_dgrad", True) and max(1, optimizer.bg_chunks) == 1
        stages = optimizer.bg_stages or (BG_DEFAULT_STAGES if after_dgrad else BG_DEFAULT_STAGES_BESIDE_DGRAD)
        sms = torch.cuda.get_device_properties(dev).multi_processor_count
        ring_ctas_per_sm = max(1, -(-int(optimizer.bg_ctas or 0) // sms))
        reserve = ring_ctas_per_sm * (stages * 8192 + 1024)
        model.set_smem_reserve(reserve)
        side = model.side_stream()
        chunks = row_buckets(P, max(1, optimizer.bg_chunks))
        last = len(chunks) - 1
        def dgrad():
            if marks is None:
                model.dgrad_gemm()
            else:
                mark("dgrad_gemm_begin")
                model.dgrad_gemm()
                mark("dgrad_gemm_end")

        # bg_after_dgrad (default): the sweep starts only after the dgrad GEMM (which then keeps
        # its full operand ring and HBM to itself: 0.17 instead of 0.30 ms) and runs beside the
        # front-end backward and the NEXT step's front-end forward. Off: it starts right after
        # wgrad, beside dgrad (1.42 instead of 1.385 ms per step at B = 1024).
        if after_dgrad:
            model.set_smem_reserve(0)
        # one chunk: fc_output.bias.grad (two small kernels, 16 us) leaves the compute stream too; it
        # only needs d(logits), and only the small-tensor AdamW at the end of the step needs it
        ev_bias = None
        if last == 0 and after_dgrad:       # (afr_train_wgrad_to honours the shared-memory reserve, which is 0 here)
            ev_fwd = torch.cuda.Event()
            ev_fwd.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev_fwd)
                optimizer.bias_grad_rows(0, P)
                ev_bias = torch.cuda.Event()
                ev_bias.record(side)
        for i, (r0, r1) in enumerate(chunks):
            if ev_bias is not None:
                model.wgrad_rows_to(r0, r1, wgrad, with_bias=False)
            else:
                model.wgrad_rows(r0, r1)
            if i == last:
                mark("wgrad")
            if after_dgrad:
                dgrad()
                model.set_smem_reserve(reserve)
            ev = torch.cuda.Event()
            ev.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev)
                if i == 0:
                    mark("adamw_begin")
                optimizer.step_rows_bg(t_step, r0, r1, optimizer.bg_ctas, stages)
                if i == last:
                    mark("adamw_end")
        if not after_dgrad:
            dgrad()
        model.frontend_backward()
        mark("dgrad")
        if ev_bias is not None:
            main.wait_event(ev_bias)
        optimizer.step_small(t_step)
        model.defer_join(side, commit=False)   # afr_adamw_rows_bg activates the copy it completed
        optimizer.end_step()
        mark("tail")
        return

    if getattr(model, "_smem_reserve", 0):
        model.set_smem_reserve(0)           # full footprints again after a background-sweep step

    if world == 1 and getattr(optimizer, "fuse_wgrad", False):
        # one kernel per bucket: wgrad GEMM whose epilogue applies AdamW to the fc_output.weight
        # tiles (the gradient never reaches HBM); 'adamw_*' marks bracket that kernel
        last = len(buckets) - 1
        overlap = getattr(optimizer, "overlap_dgrad", False) and len(buckets) == 1
        model.set_coresident(overlap)
        if overlap:
            # two streams, one CTA of each kernel on every SM: the HBM-bound wgrad+AdamW GEMM
            # (tensor pipe ~17 % busy) on the compute stream, the tensor-bound dgrad GEMM under it
            # on the side stream; both only read d(logits) and disjoint weight copies. The
            # front-end backward needs whole SMs and d(features): it joins both.
            side = model.side_stream()
            side.wait_stream(main)
            r0, r1 = buckets[0]
            mark("adamw_begin")
            optimizer.wgrad_step_rows(t_step, r0, r1)
            mark("adamw_end")
            with torch.cuda.stream(side):
                mark("dgrad_gemm_begin")
                model.dgrad_gemm()
                mark("dgrad_gemm_end")
            optimizer.bias_grad_rows(r0, r1)
            mark("wgrad")
            main.wait_stream(side)
            model.frontend_backward()
            mark("dgrad")
            optimizer.step_small(t_step)
            optimizer.end_step()
            mark("tail")
            return

        # the two small bias-gradient kernels (column sums of d(logits), 17 us) need no shared
        # memory: on the side stream they slot in next to the AdamW GEMM's CTAs instead of
        # running after it
        side = model.side_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side):
            for r0, r1 in buckets:
                optimizer.bias_grad_rows(r0, r1)

        def fused_bucket(r0, r1):
            mark("adamw_begin")
            optimizer.wgrad_step_rows(t_step, r0, r1)
            mark("adamw_end")

        model.fused_backward(buckets, lambda i, r0, r1: mark("wgrad") if i == last else None,
                             wgrad_fn=fused_bucket, marks=marks)
        mark("dgrad")
        main.wait_stream(side)
        optimizer.step_small(t_step)
        optimizer.end_step()
        mark("tail")
        return

    if world == 1:
        last = len(buckets) - 1

        def after_bucket(i, r0, r1):
            if i == last:
                mark("wgrad")
            mark("adamw_begin")
            optimizer.step_rows(t_step, r0, r1)
            mark("adamw_end")

        model.fused_backward(buckets, after_bucket, marks=marks)
        mark("dgrad")
        optimizer.step_small(t_step)
        optimizer.end_step()
        mark("tail")
        return

    model.join_pending()
    side = model.side_stream()
    rank = dist.get_rank() if rank is None else rank
    lo, hi = owned_rows(P, rank, world)
    link = getattr(model, "_peer_link", None)
    shadows = model.own_shadow_copies()
    nxt_index = 1 - model.shadow_index()
    nxt = shadows[nxt_index]

    def gather():
        """the one kernel of the exchange step: sum of the owned gradient rows over all ranks (peer
        loads / in-switch reduction), AdamW, bf16 rows to every rank (peer / multicast stores)"""
        optimizer.step_rows_gather_any(t_step, lo, hi, link, nxt_index, world)

    def after_wgrad(i, r0, r1):
        mark("wgrad")
        if link is not None and link.inline:
            # compute stream, all SMs: barrier -> gather-sum + AdamW + bf16 broadcast -> barrier
            link.h_grad.barrier(channel=0)
            mark("adamw_begin")
            gather()
            mark("adamw_end")
            link.h_grad.barrier(channel=1)
            return
        side.wait_stream(main)
        if link is not None:
            # NVLink peer memory, no collective call: barrier (every rank's dW complete) -> one
            # kernel: gather-sum of the owned gradient rows, AdamW, bf16 rows to every rank ->
            # barrier (rows delivered everywhere, nobody still reads this rank's dW)
            with torch.cuda.stream(side):
                link.h_grad.barrier(channel=0)
                mark("adamw_begin")
                gather()
                mark("adamw_end")
                link.h_grad.barrier(channel=1)
            return
        # NCCL: sum over ranks of dW[lo:hi] lands in place; ordered behind the wgrad GEMM on `main`
        rs = dist.reduce_scatter_tensor(wgrad[lo:hi], wgrad, op=dist.ReduceOp.SUM, async_op=True)
        with torch.cuda.stream(side):
            rs.wait()                                   # stream-level wait on NCCL, no host sync
            mark("adamw_begin")
            optimizer.step_rows(t_step, lo, hi)          # writes nxt[lo:hi]
            mark("adamw_end")
            ag = dist.all_gather_into_tensor(nxt, nxt[lo:hi], async_op=True)
            ag.wait()

    wgrad_fn = None
    if link is not None and link.grad_bf16:
        wgrad_fn = lambda r0, r1: model.wgrad_rows_to(r0, r1, link.wgrad, bf16=True)   # noqa: E731
    if has_samples:
        model.fused_backward([(0, P)], after_wgrad, wgrad_fn=wgrad_fn, marks=marks)
    else:
        after_wgrad(0, 0, P)
    mark("dgrad")
    # its own communicator: on the default one it would queue behind the all-gather above, which
    # waits for the sharded AdamW sweep, which waits for the reduce-scatter
    dist.all_reduce(model.small_grad_flat, op=dist.ReduceOp.SUM, group=_small_group())
    optimizer.step_small(t_step)
    if link is not None and link.inline:
        model.shadow_commit()   # everything is on the compute stream: the next forward reads the new copy
    else:
        model.defer_join(side)  # joined right before the next fc_output GEMM (renderer.join_pending)
    optimizer.end_step()
    mark("tail")


class Trainer:
    def __init__(self, model, tokens: torch.Tensor, targets: torch.Tensor, batch_size: int,
                 cfg: TrainConfig, device: torch.device):
        self.model, self.cfg, self.device, self.batch_size = model, cfg, device, batch_size
        self.rank, self.world = _world()
        u8 = targets_as_u8(targets)
        self.targets = (u8 if u8 is not None else targets.float()).to(device).contiguous()
        self.tokens = tokens.long().to(device).contiguous()
        self.fonts = None
        if cfg.sample_font_ids is not None:
            self.fonts = torch.as_tensor(cfg.sample_font_ids, dtype=torch.int32).to(device).contiguous()
        n = tokens.shape[0]
        # model.py:232-242
        val_size = int(cfg.validation_split * n)
        train_size = n - val_size
        self.train_size, self.val_size = train_size, val_size
        index_ds = _Indices(n)
        train_ds, val_ds = tud.random_split(index_ds, [train_size, val_size],
                                            generator=torch.Generator().manual_seed(cfg.seed))
        # model.py:245-266: ONE generator shared by both loaders (its stream fixes the batch order)
        g = torch.Generator()
        g.manual_seed(cfg.seed)
        self.train_loader = tud.DataLoader(train_ds, batch_size=batch_size, shuffle=True, generator=g,
                                           num_workers=0)
        self.val_loader = tud.DataLoader(val_ds, batch_size=batch_size, shuffle=False, generator=g,
                                         num_workers=0)
        self.optimizer = FusedAdamW(model, lr=cfg.learning_rate, weight_decay=cfg.weight_decay,
                                    betas=cfg.betas, background=cfg.background_adamw)  # model.py:273
        self.scheduler = torch.optim.lr_scheduler.ReduceLROnPlateau(
            self.optimizer, mode="min", factor=cfg.scheduler_factor,
            patience=cfg.scheduler_patience, min_lr=cfg.min_learning_rate)            # model.py:276-278
        self.P = cfg.sheet_height * cfg.sheet_width
        if self.world > 1:
            if self.P % self.world != 0:
                raise ValueError(f"data parallel: fc_output has {self.P} rows, not divisible by the world "
                                 f"size {self.world} (the optimizer is sharded by rows)")
            if device.type == "cuda":
                self._check_replicas()
        if (self.world > 1 and cfg.peer_memory and self.world <= 8 and device.type == "cuda" and self.P % self.world == 0
                and getattr(model, "_peer_link", None) is None and PeerLink.available()):
            # collective on every rank: symmetric allocations for dW and the bf16 weight copies; the
            # NCCL reduce-scatter / all-gather form stays as the fallback
            self._setup_peer_link()
        self.buckets = row_buckets(self.P, cfg.grad_buckets if self.world > 1 else cfg.adam_buckets)
        self.steps_done = 0

    def _setup_peer_link(self):
        """PeerLink construction is collective (symmetric-memory rendezvous): every rank must end
        up on the same path. Ranks agree on NVLS availability and on success with all-reduces; if
        any rank failed, all of them drop the link (NCCL reduce-scatter / all-gather path)."""
        model, cfg = self.model, self.cfg
        flag = torch.tensor([1 if PeerLink.nvls_available() else 0], device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        nvls = bool(flag.item())
        err = None
        try:
            # NVSwitch multicast: the gradient rows cross the links as bf16 (fp32 accumulation in the
            # switch); TrainConfig.dp_grad_bf16 = False keeps the fp32 exchange
            PeerLink(model, nvls=nvls, grad_bf16=nvls and cfg.dp_grad_bf16)
        except Exception as exc:      # no peer access / symmetric memory on this system
            err = exc
        ok = torch.tensor([0 if err is not None else 1], device=self.device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if not bool(ok.item()):
            if getattr(model, "_peer_link", None) is not None:
                PeerLink.detach(model)
            if self.rank == 0 and not cfg.quiet:
                print(f"PeerLink unavailable ({err or 'failed on another rank'}); "
                      "using NCCL reduce-scatter / all-gather")

    def _check_replicas(self):
        """Data parallel needs identical replicas: same initial parameters and the same dropout
        key on every rank (a caller that seeds ranks differently would silently diverge). Rank 0's
        values are broadcast; a mismatch is reported once."""
        model = self.model
        flat = torch.cat([p.detach().reshape(-1)[:4096].float() for p in model._ordered_params()])
        mine = torch.cat([flat, torch.tensor([float(model.dropout_seed % (1 << 24)),
                                              float(model.dropout_step)], device=flat.device)])
        ref = mine.clone()
        dist.broadcast(ref, 0)
        same = torch.tensor([1 if torch.equal(ref, mine) else 0], device=self.device)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        if not bool(same.item()):
            if self.rank == 0 and not self.cfg.quiet:
                print("data parallel: replicas differ at start-up; broadcasting rank 0's parameters "
                      "and dropout key")
            for p in model._ordered_params():
                dist.broadcast(p.data, 0)
            key = torch.tensor([model.dropout_seed & 0x7FFFFFFFFFFFFFFF, model.dropout_step],
                               dtype=torch.int64, device=self.device)
            dist.broadcast(key, 0)
            model.dropout_seed, model.dropout_step = int(key[0].item()), int(key[1].item())
            if model._ctx is not None:
                model._ctx.shadow_version = None      # the bf16 copy is rebuilt from the new master

    # ------------------------------------------------------------------ one optimizer step
    def train_batch(self, idx: torch.Tensor, loss_slot: torch.Tensor):
        """model.py:292-311 for one global batch given by dataset indices `idx` (host int64)."""
        model = self.model
        gB = idx.numel()
        lo, hi = shard_bounds(gB, self.rank, self.world)
        local = idx[lo:hi].to(self.device, non_blocking=True)
        x = self.tokens.index_select(0, local)
        t = self.targets.index_select(0, local)
        count = float(gB) * self.P
        fonts = self.fonts.index_select(0, local) if self.fonts is not None else None
        if hi > lo:
            model.fused_forward_loss(x, t, loss_count=count, sample_offset=lo, loss_out=loss_slot,
                                     font_ids=fonts)
        else:   # this rank has no sample of a ragged last batch: contribute zeros
            # the previous step's gather kernel (side stream) may still be reading this rank's
            # peer-mapped dW buffer: join it BEFORE the buffer is zeroed on the compute stream
            model.join_pending()
            loss_slot.zero_()
            for p in model._ordered_params():
                if p.grad is not None:
                    p.grad.zero_()
            if model.training:
                model.dropout_step += 1     # keep the dropout stream in step with the other ranks
        backward_and_step(model, self.optimizer, self.buckets, self.world, has_samples=hi > lo)
        self.steps_done += 1

    @torch.no_grad()
    def eval_batch(self, idx: torch.Tensor, loss_slot: torch.Tensor):
        """model.py:318-330: eval forward + mse_loss for one global validation batch."""
        model = self.model
        gB = idx.numel()
        lo, hi = shard_bounds(gB, self.rank, self.world)
        if hi <= lo:
            loss_slot.zero_()
            return
        local = idx[lo:hi].to(self.device, non_blocking=True)
        x = self.tokens.index_select(0, local)
        t = self.targets.index_select(0, local)
        # eval forward + MSE in the fused GEMM epilogue (model.eval() => dropout off)
        fonts = self.fonts.index_select(0, local) if self.fonts is not None else None
        model.fused_forward_loss(x, t, loss_count=float(gB) * self.P, dropout=False,
                                 loss_out=loss_slot, font_ids=fonts)

    def _epoch_losses(self, slots: torch.Tensor, n: int) -> float:
        """Sum of the per-batch mean losses, added on the host in batch order in double precision
        like `total += loss.item()` (model.py:311,330)."""
        v = slots[:n].clone()
        if self.world > 1:
            dist.all_reduce(v, op=dist.ReduceOp.SUM)
        total = 0.0
        for x in v.cpu().tolist():
            total += x
        return total

    # ------------------------------------------------------------------ the run
    def fit(self):
        cfg, model = self.cfg, self.model
        say = (lambda *a: None) if (cfg.quiet or self.rank != 0) else print
        if self.rank == 0 and cfg.output_dir:
            self._write_config()
        say(f"Dataset split: {self.train_size} training samples, {self.val_size} validation samples")
        best_val_loss = float("inf")
        patience_counter = 0
        best_model_state = None
        n_train, n_val = len(self.train_loader), len(self.val_loader)
        slots = torch.zeros(max(n_train, n_val, 1), dtype=torch.float32, device=self.device)
        history, lr_trace = [], []
        epoch = -1
        for epoch in range(cfg.num_epochs):
            model.train()
            for i, idx in enumerate(self.train_loader):
                self.train_batch(idx, slots[i])
                if cfg.max_steps is not None and self.steps_done >= cfg.max_steps:
                    n_train_done = i + 1
                    break
            else:
                n_train_done = n_train
            model.check_tokens_in_range()
            total_train_loss = self._epoch_losses(slots, n_train_done)
            model.eval()
            for i, idx in enumerate(self.val_loader):
                self.eval_batch(idx, slots[i])
            total_val_loss = self._epoch_losses(slots, n_val)
            avg_train_loss = total_train_loss / n_train_done            # model.py:333-334
            avg_val_loss = total_val_loss / max(n_val, 1)
            history.append((avg_train_loss, avg_val_loss))
            self.scheduler.step(avg_val_loss)                           # model.py:337
            is_best = avg_val_loss < best_val_loss                      # model.py:340-346
            if is_best:
                best_val_loss = avg_val_loss
                patience_counter = 0
                if cfg.restore_best_weights:
                    self.gather_master()          # row-sharded optimizer: complete fp32 master first
                    best_model_state = {k: v.detach().clone() for k, v in model.state_dict().items()}
                else:
                    best_model_state = model.state_dict().copy()   # shallow, as in the reference
            else:
                patience_counter += 1
            lr_now = self.optimizer.param_groups[0]["lr"]
            lr_trace.append(lr_now)
            if epoch % cfg.render_every == 0:                           # model.py:349-360
                status = (f"Epoch {epoch}, Train Loss: {avg_train_loss:.6f}, "
                          f"Val Loss: {avg_val_loss:.6f}, LR: {lr_now:.6f}")
                if is_best:
                    status += " (New Best)"
                say(status)
                if self.rank == 0 and cfg.output_dir and cfg.test_strings:
                    render_strings(model, cfg.test_strings, output_dir=f"{cfg.output_dir}/epoch_{epoch}",
                                   sheet_height=cfg.sheet_height, sheet_width=cfg.sheet_width,
                                   device=self.device, font_ids=cfg.test_font_ids)
            elif is_best:
                say(f"Epoch {epoch}, New best validation loss: {avg_val_loss:.6f}")
            if patience_counter >= cfg.early_stopping_patience:         # model.py:362-366
                say(f"Early stopping at epoch {epoch}, Best Val Loss: {best_val_loss:.6f}")
                self._load_best(best_model_state)
                break
            if cfg.max_steps is not None and self.steps_done >= cfg.max_steps:
                break
        if best_model_state is not None and patience_counter < cfg.early_stopping_patience:
            self._load_best(best_model_state)                           # model.py:369-371
            say(f"Training completed, Best Val Loss: {best_val_loss:.6f}")
        self.gather_master()
        model.join_pending()
        if self.rank == 0 and cfg.output_dir:
            final_epoch = epoch + 1 if patience_counter < cfg.early_stopping_patience else epoch
            self._write_results(final_epoch, best_val_loss, patience_counter)
        self.history, self.lr_trace = history, lr_trace
        self.best_val_loss, self.early_stopped = best_val_loss, patience_counter >= cfg.early_stopping_patience
        return model

    def _load_best(self, best_model_state):
        model = self.model
        model.join_pending()
        model.load_state_dict(best_model_state)
        if self.cfg.restore_best_weights and model._ctx is not None:
            model._ctx.shadow_version = None      # weights really changed: rebuild the bf16 copy

    @torch.no_grad()
    def gather_master(self):
        """Row-sharded optimizer (world > 1): every rank holds current fp32 values of
        fc_output.weight and its Adam moments only for the rows it owns; all-gather them so
        state_dict() / optimizer.state_dict() are complete on every rank (checkpoint contract)."""
        if self.world == 1:
            return
        self.model.join_pending()
        w = self.model.fc_output.weight
        lo, hi = owned_rows(w.shape[0], self.rank, self.world)
        tensors = [w.data]
        st = self.optimizer.state.get(w, {})
        tensors += [st[k] for k in ("exp_avg", "exp_avg_sq") if k in st]
        for t in tensors:
            dist.all_gather_into_tensor(t, t[lo:hi])
        torch.cuda.synchronize(self.device)

    def _write_config(self):
        cfg = self.cfg
        os.makedirs(cfg.output_dir, exist_ok=True)
        rows = [("num_epochs", cfg.num_epochs), ("learning_rate", cfg.learning_rate),
                ("batch_size", self.batch_size),
                ("early_stopping_patience", cfg.early_stopping_patience),
                ("validation_split", cfg.validation_split), ("weight_decay", cfg.weight_decay),
                ("embedding_dim", cfg.embedding_dim), ("dropout_rate", cfg.dropout_rate),
                ("num_attention_heads", cfg.num_attention_heads),
                ("max_length", self.model.max_length),
                ("max_chars_per_sheet", cfg.max_chars_per_sheet), ("num_samples", cfg.num_samples),
                ("data_size", self.tokens.shape[0]), ("random_seed", cfg.seed),
                ("sheet_height", cfg.sheet_height), ("sheet_width", cfg.sheet_width)]
        with open(f"{cfg.output_dir}/config.txt", "w") as f:
            f.write("# Training configuration\n")
            for k, v in rows:
                f.write(f"{k} = {v}\n")

    def _write_results(self, final_epoch, best_val_loss, patience_counter):
        cfg = self.cfg
        stopped = patience_counter >= cfg.early_stopping_patience
        with open(f"{cfg.output_dir}/training_results.txt", "w") as f:
            f.write("# Training Results\n")
            f.write(f"final_epoch = {final_epoch}\n")
            f.write(f"best_validation_loss = {best_val_loss:.6f}\n")
            f.write(f"final_learning_rate = {self.optimizer.param_groups[0]['lr']:.6f}\n")
            f.write(f"early_stopped = {stopped}\n")
            f.write(f"training_duration_epochs = {final_epoch}\n")
            f.write(f"training_completed = {datetime.datetime.now().strftime('%Y-%m-%d %H:%M:%S')}\n")


def train_attention_model(model, dataset, batch_size, cfg: Optional[TrainConfig] = None,
                          device: Optional[torch.device] = None):
    """Same call as model.py:209: `dataset` is the TensorDataset(int64 [N,L], fp32 [N,H,W]) that
    helpers.load_string_dataset returns (or any (tokens, targets) pair)."""
    cfg = cfg or TrainConfig()
    if isinstance(dataset, tud.TensorDataset):
        tokens, targets = dataset.tensors
    else:
        tokens, targets = dataset
    device = device or next(model.parameters()).device
    trainer = Trainer(model, tokens, targets, batch_size, cfg, device)
    trainer.fit()
    model._trainer = trainer
    return model
