#!/usr/bin/env python
"""bench.py -- train glyphs/s of the B200 hot path (BASELINE.json metric) on synthetic sheets.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-fuse] [--no-render]

N > 1 is launched by the driver as `python -m torch.distributed.run --nproc-per-node N bench.py
--gpus N ...` (one rank per GPU, NCCL). Prints ONE JSON line on rank 0.

Workload (config[1] of BASELINE.json, the configuration the metric is quoted on): the FiraCode
model at the reference's shapes (100 chars -> 80x240 sheet, 122.9 M parameters, weights from the
reference's seed-42 initialisation), the reference's GPU batch of 1024 glyphs per GPU
(model.py:409), one step = forward + clamp/MSE loss + backward + AdamW on one batch. Weak scaling:
the per-GPU batch stays 1024, gradients are all-reduced over NCCL.

  value   whole-job glyphs/s with the batches already resident in HBM (CUDA events, max over ranks)
  e2e     the same step driven through the public API with HOST buffers: every step copies its
          tokens + uint8 sheets from pinned host memory and reads the loss back
  roofline  the dominant HBM-bound kernel (single GPU: the AdamW sweep over fc_output.weight -- by
            default the background ring kernel that shares the SMs with the rest of backward, or with
            --step-mode fused the wgrad GEMM with the AdamW epilogue; N > 1: the row-sharded AdamW
            kernel) timed live with CUDA events on the stream it runs on
  gpu_eager_baseline  the reference's step as plain eager PyTorch on the same B200 (fp32 as the
            reference configures it, and under autocast(bf16)), the "existing Blackwell path" bar
  dp_parity (N > 1) the data-parallel step against the single-GPU step on the same global batch,
            checked before the timed region; a mismatch fails the run
  render    batched inference render glyphs/s (uint8 sheets), device-resident and end to end
  cpu_baseline  the oracle port of the reference's CPU path timed on this box's host cores
`--impl reference` times only that CPU path (the reference arm of the contract).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

BATCH_PER_GPU = 1024        # model.py:409
CPU_BATCH = 256             # model.py:411 (the reference's CPU batch)
SEED = 42
METRIC = "train_glyphs_per_sec"
UNIT = "glyphs/s"
K_FEAT, P_PIX, N_PARAMS_W = 6400, 19200, 19200 * 6400
GEMM_FLOP_PER_GLYPH = 3 * 2 * K_FEAT * P_PIX          # fwd + dgrad + wgrad of fc_output
ADAMW_BYTES_PER_PARAM = 30                            # p,g,m,v read; p,m,v write; bf16 shadow write
ADAMW_TRAFFIC_NCU = 3629e6                            # dram read+write per full sweep, profiles/r01_ncu_summary.md
FUSED_BYTES_PER_PARAM = 26                            # wgrad GEMM + AdamW epilogue: p,m,v read; p,m,v + bf16 copy write
FUSED_TRAFFIC_NCU = None                              # filled from profiles/ once captured (see load_fused_traffic)
RENDER_BATCH = 4096


_STDOUT_FD = None


def claim_stdout():
    """stdout carries exactly ONE JSON line: libraries that print there (NCCL's version banner,
    torchrun's OMP notice) are sent to stderr for the whole run; emit_line() writes the result to
    the real stdout."""
    global _STDOUT_FD
    if _STDOUT_FD is None:
        sys.stdout.flush()
        _STDOUT_FD = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _STDOUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_STDOUT_FD, data)


def load_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return dict(hbm=float(d["hbm_gbs"]), tf_burst=float(d["bf16_tflops"]),
                    tf_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])),
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, power, reasons = [], [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm),
                "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_train_glyphs_per_sec(steps: int, warmup: int):
    """The reference's CPU path (model.py:291-311 at its CPU batch of 256), as restated by the
    oracle port, with all host threads torch will use."""
    from oracle import afr_oracle as orc
    cfg = orc.OracleConfig()
    state = orc.init_state(cfg, seed=SEED)
    strings = orc.dataset_strings(CPU_BATCH)
    tokens = orc.encode_strings(strings, cfg.max_length)
    targets = orc.targets_to_f32(orc.synthetic_targets_u8(strings, cfg, seed=1234))
    opt = orc.AdamWState()
    gen = torch.Generator().manual_seed(SEED)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        masks = {"embed": torch.rand((CPU_BATCH, 100, 32), generator=gen) >= cfg.p_embed,
                 "attn": torch.rand((CPU_BATCH, 4, 100, 100), generator=gen) >= cfg.p_attn,
                 "fc1": torch.rand((CPU_BATCH, 100, 64), generator=gen) >= cfg.p_fc1}
        _, grads, _ = orc.loss_and_grads(state, tokens, targets, cfg, masks)
        orc.adamw_step(state, grads, opt)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    total = sum(times)
    return CPU_BATCH * len(times) / total, total / len(times)


def run_reference_arm(args, rank):
    if rank != 0:
        return
    # torchrun exports OMP_NUM_THREADS=1 for every rank: the CPU arm uses all host cores anyway
    try:
        torch.set_num_threads(max(1, len(os.sched_getaffinity(0))))
    except (AttributeError, OSError):
        torch.set_num_threads(max(1, os.cpu_count() or 1))
    steps = max(1, args.steps)
    value, sec = cpu_train_glyphs_per_sec(steps, max(1, args.warmup))
    cores = torch.get_num_threads()
    sample = f"{steps} train steps of the oracle port at the reference CPU batch {CPU_BATCH} (model.py:411)"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, args.warmup), "ms_per_step": sec * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "config[0] reference default on CPU: the reference's training step "
                               "(model.py:291-311: forward, mse_loss, backward, AdamW) at its CPU batch of "
                               f"{CPU_BATCH} (model.py:411), same FiraCode model (100 chars -> 80x240, 122.9 M "
                               "parameters), train-mode dropout, fp32, oracle port of the reference",
                   "implementation": "oracle/afr_oracle.py (CPU restatement pinned to the reference's outputs; "
                                     "measured faster than the unmodified reference in the build container)",
                   "batch": CPU_BATCH, "device": "host CPU", "threads": cores,
                   "max_length": 100, "sheet": "80x240", "params": 122912896, "parallelism": "single process"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host_cpus": os.cpu_count()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


def load_traffic(name, default=None):
    """DRAM bytes (read + write) of one launch of a kernel from a committed `ncu --set full` capture
    (profiles/<name>), or `default` before it exists."""
    try:
        with open(os.path.join(REPO, "profiles", name)) as f:
            return float(json.load(f)["dram_bytes_read_plus_write"])
    except (OSError, KeyError, ValueError):
        return default


def load_fused_traffic():
    return load_traffic("r01_fused_traffic.json", FUSED_TRAFFIC_NCU)


def gpu_eager_baseline(device, steps=8, warmup=3):
    """The reference's training step as plain eager PyTorch on the SAME B200 (the "existing
    Blackwell library path": ATen / cuBLAS kernels, torch.optim.AdamW), batch 1024, train-mode
    dropout, timed with CUDA events: once in fp32 exactly as the reference configures the device
    (model.py:86-93 sets no TF32 / autocast), once under torch.autocast(bfloat16). The module
    arithmetic is the oracle's restatement of model.py:158-204 (the reference itself cannot travel
    to the GPU box); the optimizer is torch's own AdamW with the reference's hyper-parameters."""
    from oracle import afr_oracle as orc
    import torch.nn.functional as F
    cfg = orc.OracleConfig()
    out = {}
    strings = orc.dataset_strings(BATCH_PER_GPU)
    tokens = orc.encode_strings(strings, cfg.max_length).to(device)
    targets = orc.targets_to_f32(orc.synthetic_targets_u8(strings, cfg, seed=1234)).to(device)
    gen = torch.Generator(device=device).manual_seed(SEED)
    prev_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    try:
        for name, autocast in (("fp32", False), ("autocast_bf16", True)):
            params = {k: torch.nn.Parameter(v.to(device)) for k, v in orc.init_state(cfg, seed=SEED).items()}
            opt = torch.optim.AdamW(params.values(), lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99))
            ev = []
            for i in range(warmup + steps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                masks = {"embed": torch.rand((BATCH_PER_GPU, 100, 32), generator=gen, device=device) >= cfg.p_embed,
                         "attn": torch.rand((BATCH_PER_GPU, 4, 100, 100), generator=gen, device=device) >= cfg.p_attn,
                         "fc1": torch.rand((BATCH_PER_GPU, 100, 64), generator=gen, device=device) >= cfg.p_fc1}
                opt.zero_grad(set_to_none=True)
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                    y = orc.forward(params, tokens, cfg, masks)
                    loss = F.mse_loss(y.float(), targets.view(y.shape))
                loss.backward()
                opt.step()
                e1.record()
                if i >= warmup:
                    ev.append((e0, e1))
            torch.cuda.synchronize()
            ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
            out[name] = {"ms_per_step": ms, "value": BATCH_PER_GPU / (ms / 1e3), "unit": UNIT,
                         "final_loss": float(loss.detach())}
            del params, opt
            torch.cuda.empty_cache()
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = prev_tf32
    out["what"] = ("reference step (model.py:291-311) as eager PyTorch on this GPU: oracle restatement of the module, "
                   f"torch.optim.AdamW, batch {BATCH_PER_GPU}, train-mode dropout, data resident, {steps} steps "
                   f"after {warmup} warm-up, CUDA events")
    return out


def measure_render(model, device, rank, world, dist, bmp_set=False):
    """Batched inference render (helpers.py:46-74 without the file writes): glyph strings sharded
    over the ranks, no collective; one pass = RENDER_BATCH strings -> uint8 sheets. Reported
    device-resident (tokens in HBM, sheets left in HBM) and end to end through
    render.RenderPipeline (int64 tokens from pinned host memory, uint8 sheets copied back to pinned
    host memory on a copy stream under the next batch's render).
    bmp_set: BASELINE config 5 -- a model whose embedding covers the 65,536 code points of the
    Unicode BMP renders sample i = [code point i, 0, 0, ...]; the code points are sharded over the
    ranks (65536 / world each)."""
    from ai_font_renderer_b200.render import RenderPipeline
    if bmp_set:
        from ai_font_renderer_b200.renderer import AttentionFontRenderer
        torch.manual_seed(SEED)
        model = AttentionFontRenderer(vocab=65536).to(device)
        per_rank = 65536 // world
        codes = torch.arange(rank * per_rank, (rank + 1) * per_rank)
        tok_h = torch.zeros((per_rank, 100), dtype=torch.int64)
        tok_h[:, 0] = codes
        n_pass, n_rot = max(1, per_rank // RENDER_BATCH), max(1, per_rank // RENDER_BATCH)
        batch = min(RENDER_BATCH, per_rank)
    else:
        n_pass, n_rot, batch = 20, 4, RENDER_BATCH
        g = torch.Generator().manual_seed(99 + rank)
        n = batch * n_rot
        lengths = torch.randint(10, 101, (n, 1), generator=g)
        letters = torch.randint(65, 91, (n, 100), generator=g)
        letters[torch.rand((n, 100), generator=g) < 0.148] = 32
        tok_h = torch.where(torch.arange(100).unsqueeze(0) < lengths, letters, torch.zeros_like(letters)).long()
        if world > 1:
            model.join_pending()
            model.set_sm_limit(0)          # rendering has no collective: all SMs
    tok_h = tok_h.pin_memory()
    tok_d = tok_h.to(device)
    was_training = model.training
    model.eval()
    pipe = RenderPipeline(model, device, batch)
    out_h = torch.empty((batch * n_pass, 80, 240), dtype=torch.uint8).pin_memory()
    idx = torch.cat([torch.arange((i % n_rot) * batch, (i % n_rot + 1) * batch) for i in range(n_pass)])
    tok_seq = tok_h[idx].pin_memory()      # the n_pass batches back to back

    def resident():
        for i in range(n_pass):
            s = (i % n_rot) * batch
            model.render_u8(tok_d[s:s + batch], out=pipe.dev[i % 2][:batch])

    def e2e():
        pipe.render_to_host(tok_seq, out=out_h)

    res = {}
    for name, fn in (("resident", resident), ("e2e", e2e)):
        fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        res[name] = ms
    # one string per call, the way the reference's render_strings drives the model (helpers.py:62-64)
    single_ms = None
    if not bmp_set:
        one = tok_d[:1]
        for _ in range(5):
            model.render_u8(one)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            model.render_u8(one)
        e1.record()
        torch.cuda.synchronize()
        single_ms = e0.elapsed_time(e1) / 50
    model.check_tokens_in_range()
    model.train(was_training)
    glyphs = batch * n_pass * world
    flop = 2.0 * K_FEAT * P_PIX
    return {"metric": "render_glyphs_per_sec", "unit": UNIT, "batch_per_gpu": batch,
            "passes": n_pass, "glyphs": glyphs, "value": glyphs / (res["resident"] / 1e3),
            "ms_per_pass": res["resident"] / n_pass,
            "single_string_ms": single_ms,
            "tflops_per_gpu": glyphs / world * flop / (res["resident"] / 1e3) / 1e12,
            "e2e": {"value": glyphs / (res["e2e"] / 1e3), "unit": UNIT,
                    "h2d_bytes_per_pass": batch * 100 * 8, "d2h_bytes_per_pass": batch * P_PIX,
                    "ms_per_pass": res["e2e"] / n_pass,
                    "how": "render.RenderPipeline: D2H of batch i on a copy stream under the render of batch i+1"},
            "workload": ("config[4]: 65,536 BMP code points, embedding [65536,32], sharded over the ranks"
                         if bmp_set else "config[1] model, synthetic strings of 10-100 chars"),
            "output": "uint8 sheets (helpers.py:33 quantisation fused in the GEMM epilogue)"}


def workload_config(n_gpus, step_mode="two-kernel", dp_mode=None):
    return {"workload": "config[1] FiraCode model 100 chars -> 80x240, fused fwd/bwd/AdamW, "
                        "1024 glyphs per GPU per step (model.py:409)",
            "optimizer": {"fused": "AdamW of fc_output.weight inside the wgrad GEMM epilogue (gradient not materialised)",
                          "background": "AdamW of fc_output.weight as a background sweep (128-thread cp.async ring "
                                        "CTAs on a second stream, co-resident with the front-end backward and the "
                                        "next front-end forward; --bg-after-dgrad 0: also with the dgrad GEMM)",
                          "two-kernel": "AdamW sweep kernel over fc_output.weight"}[step_mode],
            "step_mode": step_mode,
            "batch_per_gpu": BATCH_PER_GPU, "global_batch": BATCH_PER_GPU * n_gpus,
            "max_length": 100, "sheet": "80x240", "params": 122912896,
            "parallelism": (f"dp{n_gpus} ({dp_mode}): batch sharded; AdamW of fc_output.weight sharded by rows, "
                            + ("gradient rows summed inside the NVSwitch (multimem.ld_reduce), bf16 rows multicast"
                               if (dp_mode or "").startswith("nvls") else
                               "NCCL reduce-scatter / all-gather" if dp_mode == "nccl" else
                               "gradient rows read from / bf16 rows stored to NVLink peer memory")
                            + ("; gradient rows cross the links as bf16, summed in fp32" if (dp_mode or "").endswith("-bf16") else "")
                            ) if n_gpus > 1 else "single",
            "l2": "no flush: one step streams 3.9 GB (fp32 master, grads, Adam moments, bf16 shadow) "
                  "through the 126 MB L2 and rotates over 8 resident batches"}



# ------------------------------------------------------------------------------------ config 4
CFG4 = dict(max_length=64, sheet_height=64, sheet_width=64, vocab=128, embedding_dim=128, num_heads=8,
            fc1_width=128)
CFG4_BATCH = 8192


def run_config4(args, rank, local_rank, world):
    """BASELINE.json configs[3]: scaled synthetic workload -- 64 x 64 glyph bitmaps, 64-char strings,
    widened net (embed 128, 8 heads, fc1 128: fc_output 8192 -> 4096), 8192 glyphs per GPU per step
    (65,536 global at 8 GPUs, weak scaling). An extension: the reference has no such configuration
    (restated-oracle parity, tests/test_wide_gpu.py). Same step as config 2: forward + clamp/MSE +
    backward + AdamW, batch resident / from pinned host memory for e2e."""
    import torch.distributed as dist
    from ai_font_renderer_b200.data import HostBatchFeeder
    from ai_font_renderer_b200.optim import FusedAdamW
    from ai_font_renderer_b200.renderer import AttentionFontRenderer
    from ai_font_renderer_b200.training import backward_and_step, row_buckets
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    args.warmup = max(3, args.warmup)
    B = args.batch if args.batch != BATCH_PER_GPU else CFG4_BATCH
    gB = B * world
    L, H_, W_ = CFG4["max_length"], CFG4["sheet_height"], CFG4["sheet_width"]
    P, K = H_ * W_, L * CFG4["fc1_width"]
    torch.manual_seed(SEED)
    model = AttentionFontRenderer(**CFG4).to(device).train()
    step_mode = ("background" if args.step_mode == "auto" else args.step_mode) if world == 1 else "two-kernel"
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), fuse_wgrad=step_mode == "fused",
                     background=step_mode == "background")
    dp_mode = None
    if world > 1:
        from ai_font_renderer_b200.training import PeerLink
        dp_mode = args.dp_mode
        if dp_mode == "auto":
            dp_mode = "nvls-side-bf16" if PeerLink.nvls_available() else "peer-side"
        if dp_mode != "nccl":
            base = dp_mode[:-5] if dp_mode.endswith("-bf16") else dp_mode
            PeerLink(model, ctas=args.comm_ctas or 12, inline=base in ("peer", "nvls"), nvls=base.startswith("nvls"),
                     grad_bf16=dp_mode.endswith("-bf16"))
    n_rot = 4
    g = torch.Generator().manual_seed(1234 + rank)
    n = B * n_rot
    lengths = torch.randint(10, L + 1, (n, 1), generator=g)
    letters = torch.randint(65, 91, (n, L), generator=g)
    letters[torch.rand((n, L), generator=g) < 0.148] = 32
    tok_h = torch.where(torch.arange(L).unsqueeze(0) < lengths, letters, torch.zeros_like(letters)).long()
    # U-shaped grey levels, 80 % white
    tgt_h = torch.where(torch.rand((n, H_, W_), generator=g) < 0.8, torch.full((n, H_, W_), 255, dtype=torch.uint8),
                        (torch.randint(0, 4, (n, H_, W_), generator=g) * 64).to(torch.uint8))
    tok_h, tgt_h = tok_h.pin_memory(), tgt_h.pin_memory()
    tok_d, tgt_d = tok_h.to(device), tgt_h.to(device)
    buckets = row_buckets(P, 1)
    count = float(gB) * P
    loss_buf = torch.zeros(args.steps + args.warmup + 8, dtype=torch.float32, device=device)

    def step_resident(i):
        s = (i % n_rot) * B
        model.fused_forward_loss(tok_d[s:s + B], tgt_d[s:s + B], loss_count=count, sample_offset=rank * B,
                                 loss_out=loss_buf[i])
        backward_and_step(model, opt, buckets, world)

    feeder = HostBatchFeeder(tok_h, tgt_h, B, device)
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    loss_ready, loss_copied = torch.cuda.Event(), torch.cuda.Event()

    def step_e2e(i):
        x_dev, t_dev = feeder.get(i)
        loss = model.fused_forward_loss(x_dev, t_dev, loss_count=count, sample_offset=rank * B)
        loss_ready.record()
        with torch.cuda.stream(feeder.copy_stream):
            feeder.copy_stream.wait_event(loss_ready)
            loss_host.copy_(loss.view(1), non_blocking=True)
            loss_copied.record()
        backward_and_step(model, opt, buckets, world)
        feeder.done(i)
        loss_copied.synchronize()

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.kernel_launches()
        e0.record()
        for i in range(k):
            fn(i)
        model.join_pending()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = model.kernel_launches() - l0
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        return ms, launches

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, launches = timed(step_resident, args.steps)
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    model.check_tokens_in_range()
    # the three fc_output GEMMs alone (tensor-bound part of the step), CUDA events around each
    ctx = model._ctx
    st = torch.cuda.current_stream(device).cuda_stream
    model.fused_forward_loss(tok_d[:B], tgt_d[:B], loss_count=count)
    gemm_ms = {}
    for name, call in (("wgrad", lambda: ctx.lib.afr_train_wgrad(ctx.handle, 0, P, st)),
                       ("dgrad", lambda: ctx.lib.afr_train_dgrad_gemm(ctx.handle, st))):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ctx.check(call())
        e1.record()
        torch.cuda.synchronize()
        gemm_ms[name] = e0.elapsed_time(e1) / 5
    final_loss = float(loss_buf[args.steps - 1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    peaks = load_peaks()
    value = gB * args.steps / (ms_res / 1e3)
    flop_gemm = 2.0 * B * K * P
    front_flop = B * (2.0 * L * 128 * (384 + 128 + 128) * 3 * 3 + 2.0 * 8 * L * L * 16 * 2 * 3.5)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "config[3] scaled synthetic: 64x64 glyph sheets, 64-char strings, widened net "
                               "(embed 128, 8 heads, fc1 128; fc_output 8192 -> 4096), 8192 glyphs per GPU per step "
                               "(65,536 global at 8 GPUs); extension of the reference, restated-oracle parity",
                   "batch_per_gpu": B, "global_batch": gB, "max_length": L, "sheet": f"{H_}x{W_}",
                   "params": sum(p.numel() for p in model.parameters()), "step_mode": step_mode,
                   "parallelism": f"dp{world} ({dp_mode})" if world > 1 else "single",
                   "l2": "no flush: one step streams > 5 GB of activations through the 126 MB L2 and rotates over "
                         "4 resident batches"},
        "e2e": {"value": gB * args.steps / (ms_e2e / 1e3), "unit": UNIT,
                "h2d_bytes_per_step": int(feeder.h2d_bytes_per_batch), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks,
        "roofline": {"kernel": "gemm_bf16_tcgen05_kernel (fc_output wgrad GEMM, 4096 x 8192 x 8192)",
                     "bound": "tensor", "achieved": flop_gemm / (gemm_ms["wgrad"] / 1e3) / 1e12,
                     "peak": peaks["tf_burst"], "unit": "TFLOP/s",
                     "frac": flop_gemm / (gemm_ms["wgrad"] / 1e3) / 1e12 / peaks["tf_burst"], "traffic": None,
                     "ms_per_launch": gemm_ms["wgrad"], "dgrad_ms": gemm_ms["dgrad"],
                     "dgrad_tflops": flop_gemm / (gemm_ms["dgrad"] / 1e3) / 1e12,
                     "peak_source": peaks["source"],
                     "note": "the step is dominated by the fp32 SIMT attention kernels of the wide front-end "
                             "(profiles/r02_config4_launches.csv), not by this GEMM"},
        "train_tflops_whole_step": value * (3 * 2.0 * K * P + front_flop / B) / 1e12 / world,
        "final_loss": final_loss,
    }
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0

# ------------------------------------------------------------------------------------ GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batch", type=int, default=BATCH_PER_GPU)
    ap.add_argument("--dp-mode", default="auto",
                    choices=["auto", "peer", "peer-side", "nvls", "nvls-side", "nccl", "peer-bf16", "peer-side-bf16",
                             "nvls-bf16", "nvls-side-bf16"],
                    help="N > 1: 'peer' = row-sharded AdamW reading/writing NVLink peer memory in one "
                         "kernel on all SMs of the compute stream, 'peer-side' = the same kernel on a side "
                         "stream on --comm-ctas SMs under the rest of backward, 'nccl' = reduce-scatter / "
                         "all-gather, 'auto' = peer at 2 GPUs, peer-side above")
    ap.add_argument("--comm-ctas", type=int, default=0,
                    help="N > 1: SMs left to the communication kernel (0 = PeerLink.default_ctas / 32 for "
                         "NCCL); the persistent kernels use the rest")
    ap.add_argument("--step-mode", default="auto", choices=["auto", "background", "fused", "two-kernel"],
                    help="single GPU: how optimizer.step() of fc_output.weight runs. background = sweep kernel "
                         "on a second stream sharing the SMs with the rest of backward (default), fused = in the "
                         "wgrad GEMM's epilogue, two-kernel = wgrad then a stand-alone sweep")
    ap.add_argument("--bg-stages", type=int, default=0, help="background sweep: ring stages of 8 KB (0 = default)")
    ap.add_argument("--bg-chunks", type=int, default=0, help="background sweep: row chunks of wgrad / sweep (0 = default)")
    ap.add_argument("--bg-ctas", type=int, default=0, help="background sweep: CTAs (0 = one per SM)")
    ap.add_argument("--bg-after-dgrad", type=int, default=-1,
                    help="background sweep: 1 = start it after the dgrad GEMM, 0 = right after wgrad (-1 = default)")
    ap.add_argument("--no-fuse", action="store_true",
                    help="single GPU: same as --step-mode two-kernel")
    ap.add_argument("--config", type=int, default=2, choices=[2, 4],
                    help="BASELINE.json workload: 2 = configs[1], the reference's model on one B200 (the metric's "
                         "configuration, default); 4 = configs[3], the scaled synthetic workload (64x64 sheets, "
                         "64-char strings, widened net 128/8/128, 8192 glyphs per GPU = 64k global at 8 GPUs)")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the eager-PyTorch-on-B200 baseline")
    ap.add_argument("--no-dp-parity", action="store_true", help="N > 1: skip the parity check before timing")
    ap.add_argument("--overlap-dgrad", action="store_true",
                    help="single GPU: run the dgrad GEMM co-resident under the wgrad+AdamW GEMM (measured slower)")
    ap.add_argument("--no-render", action="store_true", help="skip the batched-render measurement")
    ap.add_argument("--adam-buckets", type=int, default=1,
                    help="single GPU: row buckets of the wgrad GEMM / AdamW sweep over fc_output.weight")
    args = ap.parse_args()
    claim_stdout()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return 0
    if args.config == 4:
        return run_config4(args, rank, local_rank, world)

    import torch.distributed as dist
    from ai_font_renderer_b200.data import HostBatchFeeder, fast_synthetic_batch
    from ai_font_renderer_b200.optim import FusedAdamW
    from ai_font_renderer_b200.renderer import AttentionFontRenderer
    from ai_font_renderer_b200.training import backward_and_step, row_buckets

    if not torch.cuda.is_available():
        emit_line({"error": "no CUDA device: the B200 path has no CPU fallback"})
        return 2
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # keep stdout to the one JSON line
        if args.dp_mode == "nccl":
            os.environ.setdefault("NCCL_MAX_CTAS", str(args.comm_ctas or 32))
        dist.init_process_group("nccl", device_id=device)
    args.warmup = max(3, args.warmup)
    B = args.batch
    gB = B * world

    # model: the reference's construction under its seed (model.py:87-90,402)
    torch.manual_seed(SEED)
    model = AttentionFontRenderer().to(device).train()
    if args.no_fuse:
        args.step_mode = "two-kernel"
    step_mode = ("background" if args.step_mode == "auto" else args.step_mode) if world == 1 else "two-kernel"
    fused = step_mode == "fused"
    bg_kw = {}
    if step_mode == "background":
        bg_kw = dict(background=True)
        if args.bg_stages:
            bg_kw["bg_stages"] = args.bg_stages
        if args.bg_chunks:
            bg_kw["bg_chunks"] = args.bg_chunks
        if args.bg_ctas:
            bg_kw["bg_ctas"] = args.bg_ctas
        if args.bg_after_dgrad >= 0:
            bg_kw["bg_after_dgrad"] = bool(args.bg_after_dgrad)
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), fuse_wgrad=fused,
                     overlap_dgrad=fused and args.overlap_dgrad, **bg_kw)
    dp_parity = None
    if world > 1:
        from ai_font_renderer_b200.training import PeerLink
        if args.dp_mode == "auto":
            # measured on B200: the gather / AdamW / broadcast kernel is bound by
            # the NVLink egress of the gradient rows (0.43 GB per rank and step at 8 GPUs), so it hides
            # on a side stream under the rest of backward; with NVSwitch multicast (NVLS) its inbound
            # traffic and SM time shrink (2 GPUs 1.65 vs 1.82 ms, 8 GPUs 1.51 vs 1.55 ms)
            # round 2: with NVLS the gradient rows cross the links as bf16 (fp32 accumulation inside the
            # switch): 8 GPUs 1.27 vs 1.52 ms (profiles/r02)
            args.dp_mode = "nvls-side-bf16" if PeerLink.nvls_available() else ("peer" if world == 2 else "peer-side")
        dp_base = args.dp_mode[:-5] if args.dp_mode.endswith("-bf16") else args.dp_mode
        if dp_base in ("peer", "peer-side", "nvls", "nvls-side"):
            try:
                PeerLink(model, ctas=args.comm_ctas, inline=dp_base in ("peer", "nvls"),
                         nvls=dp_base.startswith("nvls"), grad_bf16=args.dp_mode.endswith("-bf16"))
            except Exception as exc:     # no symmetric memory / peer access on this system
                sys.stderr.write(f"[bench] PeerLink unavailable ({exc}); falling back to --dp-mode nccl\n")
                args.dp_mode = "nccl"
                model.set_sm_limit(torch.cuda.get_device_properties(device).multi_processor_count - 32)
        elif args.dp_mode == "nccl":
            sms = torch.cuda.get_device_properties(device).multi_processor_count
            model.set_sm_limit(sms - int(os.environ.get("NCCL_MAX_CTAS", "32")))
        if not args.no_dp_parity:
            # correctness of the shipped data-parallel mode at THIS world size, before anything is
            # timed: two steps of the sharded step against the single-GPU step on the same global
            # batch (tools/dp_check.py). A mismatch fails the run.
            sys.path.insert(0, os.path.join(REPO, "tools"))
            import dp_check
            ok, dp_parity = dp_check.compare(rank, world, device, args.dp_mode, steps=2, per_rank=96,
                                             ctas=args.comm_ctas or PeerLink.default_ctas(
                                                 world, args.dp_mode.endswith("-bf16")))
            if not ok:
                if rank == 0:
                    emit_line({"error": "data-parallel parity check failed", "dp_parity": dp_parity})
                dist.destroy_process_group()
                return 3
    n_rot = 8
    tok_h, tgt_h = fast_synthetic_batch(B * n_rot, seed=1234 + rank)
    tok_h, tgt_h = tok_h.pin_memory(), tgt_h.pin_memory()
    tok_d, tgt_d = tok_h.to(device), tgt_h.to(device)
    buckets = row_buckets(P_PIX, 8 if world > 1 else args.adam_buckets)
    count = float(gB) * P_PIX
    loss_buf = torch.zeros(args.steps + args.warmup + 8, dtype=torch.float32, device=device)

    class Marks:
        """CUDA events on whatever stream is current when a phase boundary is reached."""
        def __init__(self):
            self.ev = {}

        def __call__(self, label):
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.ev.setdefault(label, []).append(e)

    step_marks = [Marks() for _ in range(args.steps)]

    def step_resident(i, marks=None):
        s = (i % n_rot) * B
        x, t = tok_d[s:s + B], tgt_d[s:s + B]
        if marks is not None:
            marks("start")
        model.fused_forward_loss(x, t, loss_count=count, sample_offset=rank * B, loss_out=loss_buf[i],
                                 marks=marks)
        if marks is not None:
            marks("forward")
        backward_and_step(model, opt, buckets, world, marks=marks)

    # end to end: every step's tokens + uint8 sheets come from pinned HOST memory through the
    # package's HostBatchFeeder (copy stream, one batch ahead) and the step's loss is read back.
    feeder = HostBatchFeeder(tok_h, tgt_h, B, device)
    loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()
    loss_ready, loss_copied = torch.cuda.Event(), torch.cuda.Event()

    def step_e2e(i):
        x_dev, t_dev = feeder.get(i)                             # H2D of batch i (and i+1 started)
        loss = model.fused_forward_loss(x_dev, t_dev, loss_count=count, sample_offset=rank * B)
        # D2H read of THIS step's loss, every step: copied on the feeder's copy stream as soon as
        # the forward has produced it, waited for on the host after the rest of the step has been
        # enqueued -- the host blocks on the value (model.py:311 `.item()`), the GPU never idles
        loss_ready.record()
        with torch.cuda.stream(feeder.copy_stream):
            feeder.copy_stream.wait_event(loss_ready)
            loss_host.copy_(loss.view(1), non_blocking=True)
            loss_copied.record()
        backward_and_step(model, opt, buckets, world)
        feeder.done(i)
        loss_copied.synchronize()
        return float(loss_host[0])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k, with_events):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        launches0 = model.kernel_launches()
        e0.record()
        for i in range(k):
            fn(i, step_marks[i]) if with_events else fn(i)
        model.join_pending()      # a deferred optimizer sweep (side stream) belongs to the timed steps
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        launches = model.kernel_launches() - launches0
        if world > 1:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
            dist.barrier()
        return ms, launches

    for i in range(args.warmup):
        step_resident(i)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_res, launches = timed(step_resident, args.steps, True)
    for i in range(2):
        step_e2e(i)
    ms_e2e, _ = timed(step_e2e, args.steps, False)
    clocks = sampler.stop() if rank == 0 else None
    model.check_tokens_in_range()
    final_loss = float(loss_buf[args.steps - 1])

    # per-phase device times (this rank), averaged over the timed steps. Compute stream:
    # forward | wgrad | dgrad + front-end backward | tail (small AdamW + join with the side stream).
    # Side stream: every AdamW sweep launch, begin -> end.
    order = ("start", "forward", "wgrad", "dgrad", "tail")
    phase_ms = {}
    for a, b_ in zip(order, order[1:]):
        phase_ms[b_] = sum(m.ev[a][0].elapsed_time(m.ev[b_][0]) for m in step_marks) / args.steps
    # single GPU: the AdamW sweep sits between the 'wgrad' and 'dgrad' marks of the one stream
    adam_launch_ms = [x.elapsed_time(y) for m in step_marks
                      for x, y in zip(m.ev["adamw_begin"], m.ev["adamw_end"])]
    adamw_ms_per_launch = sum(adam_launch_ms) / len(adam_launch_ms)
    adamw_ms_per_step = sum(adam_launch_ms) / args.steps
    launches_per_step = max(1, len(adam_launch_ms) // args.steps)
    phase_ms["adamw_side_stream"] = adamw_ms_per_step
    # the two tensor-bound GEMMs by themselves (forward incl. its 5 us loss-finalize kernel)
    fwd_gemm_ms = sum(m.ev["frontend"][0].elapsed_time(m.ev["forward"][0]) for m in step_marks) / args.steps
    dgrad_gemm_ms = sum(m.ev["dgrad_gemm_begin"][0].elapsed_time(m.ev["dgrad_gemm_end"][0])
                        for m in step_marks) / args.steps
    phase_ms["frontend_forward"] = phase_ms["forward"] - fwd_gemm_ms
    phase_ms["forward_gemm"] = fwd_gemm_ms
    phase_ms["dgrad_gemm"] = dgrad_gemm_ms

    from ai_font_renderer_b200 import training as _tr
    bg_stages_eff = getattr(opt, "bg_stages", 0) or (
        _tr.BG_DEFAULT_STAGES if getattr(opt, "bg_after_dgrad", True) else _tr.BG_DEFAULT_STAGES_BESIDE_DGRAD)
    # the same sweep alone on the device (nothing else running), for reference
    iso = []
    if world == 1:
        model.join_pending()
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_step = opt.begin_step()
            torch.cuda.synchronize()
            e0.record()
            if fused:
                opt.wgrad_step_rows(t_step, 0, P_PIX)   # dZ / features of the last step are still there
            elif step_mode == "background":
                opt.step_rows_bg(t_step, 0, P_PIX, opt.bg_ctas, bg_stages_eff)
            else:
                opt.step_rows(t_step, 0, P_PIX)
            e1.record()
            opt.step_small(t_step)
            opt.end_step()
            torch.cuda.synchronize()
            iso.append(e0.elapsed_time(e1))
    adamw_iso_ms = sorted(iso)[len(iso) // 2] if iso else None      # median of 5

    render = None if args.no_render else measure_render(model, device, rank, world, dist)
    render_bmp = None if args.no_render else measure_render(model, device, rank, world, dist, bmp_set=True)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peaks = load_peaks()
    value = gB * args.steps / (ms_res / 1e3)
    e2e_value = gB * args.steps / (ms_e2e / 1e3)
    if fused:
        # wgrad GEMM with the AdamW epilogue: p, m, v read + written, bf16 copy written, plus the
        # two bf16 operands (dZ [B,P], features [B,K]) read once
        kernel_name = "gemm_bf16_tcgen05_kernel<kEpiAdamW> (wgrad GEMM, AdamW of fc_output.weight in its epilogue)"
        adamw_bytes = FUSED_BYTES_PER_PARAM * N_PARAMS_W + 2 * B * (P_PIX + K_FEAT)
        traffic = load_fused_traffic()
    elif world > 1 and args.dp_mode != "nccl":
        # row-sharded optimizer step of this rank: HBM bytes of the owned rows (p, m, v read +
        # written, local gradient read, local bf16 copy written); the kernel is bound by the
        # NVLink egress of the other ranks' gradient rows, reported next to it
        owned = N_PARAMS_W // world
        kernel_name = ("adamw_gather_nvls_kernel (in-switch gradient sum + AdamW + multicast bf16 rows)"
                       if args.dp_mode.startswith("nvls") else
                       "adamw_gather_kernel (peer gradient rows + AdamW + bf16 rows to peers)")
        adamw_bytes = (24 + (2 if args.dp_mode.endswith("-bf16") else 4) + 2) * owned
        traffic = None
    elif step_mode == "background":
        # the same 30 B/parameter as the plain sweep (p, g, m, v read; p, m, v + bf16 copy written),
        # timed begin -> end on ITS stream while dgrad / the front-end backward / the next
        # front-end forward run beside it on the same SMs
        kernel_name = ("adamw_ring_kernel (background AdamW sweep over fc_output.weight + bf16 shadow, "
                       "co-resident with the compute kernels of the step)")
        adamw_bytes = ADAMW_BYTES_PER_PARAM * N_PARAMS_W
        traffic = load_traffic("r02_ring_traffic.json")
    else:
        kernel_name = "adamw_kernel (fc_output.weight sweep + bf16 shadow)"
        adamw_bytes = ADAMW_BYTES_PER_PARAM * N_PARAMS_W
        traffic = ADAMW_TRAFFIC_NCU
    adamw_gbs = adamw_bytes / (adamw_ms_per_step / 1e3) / 1e9
    one_gemm_tf = lambda ms: 2.0 * B * K_FEAT * P_PIX / (ms / 1e3) / 1e12
    gemm_tf = B * GEMM_FLOP_PER_GLYPH / ((phase_ms["forward"] + phase_ms["wgrad"] + phase_ms["dgrad"]) / 1e3) / 1e12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_res / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world, step_mode, args.dp_mode if world > 1 else None),
        "e2e": {"value": e2e_value, "unit": UNIT,
                "h2d_bytes_per_step": int(feeder.h2d_bytes_per_batch), "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"kernel": kernel_name,
                     "bound": "hbm", "achieved": adamw_gbs, "peak": peaks["hbm"], "unit": "GB/s",
                     "frac": adamw_gbs / peaks["hbm"], "traffic": traffic,
                     "algorithmic_bytes_per_launch": adamw_bytes // launches_per_step,
                     "launches_per_step": launches_per_step, "ms_per_launch": adamw_ms_per_launch,
                     "alone_ms_per_sweep": adamw_iso_ms,
                     "alone_frac": (adamw_bytes / (adamw_iso_ms / 1e3) / 1e9 / peaks["hbm"]) if adamw_iso_ms else None,
                     "tensor_flop_per_launch": (2 * B * K_FEAT * P_PIX // launches_per_step) if fused else 0,
                     "nvlink_egress_bytes_per_step": ((2 if args.dp_mode.endswith("-bf16") else 4)
                                                      * (N_PARAMS_W // world) * (world - 1)) if world > 1 else 0,
                     "peak_source": peaks["source"]},
        "gemm": {"tflops_incl_frontend_and_epilogues": gemm_tf,
                 "frac_of_bf16_sustained_peak": gemm_tf / peaks["tf_sustained"],
                 "frac_of_bf16_burst_peak": gemm_tf / peaks["tf_burst"],
                 "flop_per_glyph": GEMM_FLOP_PER_GLYPH,
                 # the tensor-bound GEMM kernels alone (CUDA events around each launch): 2*B*K*P flop
                 "forward_gemm": {"ms": fwd_gemm_ms, "tflops": one_gemm_tf(fwd_gemm_ms),
                                  "frac_of_bf16_burst_peak": one_gemm_tf(fwd_gemm_ms) / peaks["tf_burst"]},
                 "dgrad_gemm": {"ms": dgrad_gemm_ms, "tflops": one_gemm_tf(dgrad_gemm_ms),
                                "frac_of_bf16_burst_peak": one_gemm_tf(dgrad_gemm_ms) / peaks["tf_burst"]}},
        "phase_ms": phase_ms,
        "train_tflops_whole_step": value * GEMM_FLOP_PER_GLYPH / 1e12 / world,
        "final_loss": final_loss,
    }
    if render is not None:
        line["render"] = render
        line["render_bmp_set"] = render_bmp
    if dp_parity is not None:
        line["dp_parity"] = dp_parity
    if world == 1 and not args.no_gpu_eager:
        del model, opt, feeder, tok_d, tgt_d
        torch.cuda.empty_cache()
        line["gpu_eager_baseline"] = gpu_eager_baseline(device)
    if world == 1 and not args.no_cpu_baseline:
        cpu_value, cpu_sec = cpu_train_glyphs_per_sec(steps=6, warmup=2)
        line["cpu_baseline"] = {
            "value": cpu_value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "host_cpus": os.cpu_count(),
            "sample": f"6 train steps of the oracle port at the reference CPU batch {CPU_BATCH} "
                      f"({cpu_sec:.2f} s/step, 2 warm-up steps)"}
    emit_line(line)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
