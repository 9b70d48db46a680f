#!/usr/bin/env python
"""Round-2 overlap probe: does the BACKGROUND AdamW sweep (afr_adamw_rows_bg: 128-thread CTAs,
<= 40 registers, bulk-copy ring in a few KB of shared memory) share the SMs with the compute
kernels of the step, and what does a chunk-pipelined step (wgrad chunk k -> L2-resident dW chunk ->
AdamW chunk k on a second stream, dgrad / front-end backward on the compute stream) cost?

  E1  the sweep alone: plain kernel vs ring kernel at several (ctas, stages)
  E2  each compute kernel alone, then with the ring sweep running beside it
  E3  the pipelined step against the current fused step (and bit-equality of the weights)
"""
import argparse
import ctypes as C
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ai_font_renderer_b200.data import fast_synthetic_batch  # noqa: E402
from ai_font_renderer_b200.optim import FusedAdamW  # noqa: E402
from ai_font_renderer_b200.renderer import AttentionFontRenderer, _stream_ptr  # noqa: E402
from ai_font_renderer_b200.training import backward_and_step, row_buckets  # noqa: E402

P, K = 19200, 6400


def build(dev, fuse=True):
    torch.manual_seed(42)
    model = AttentionFontRenderer().to(dev).train()
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), fuse_wgrad=fuse)
    return model, opt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    B = args.batch
    dev = torch.device("cuda", 0)
    model, opt = build(dev)
    tok, tgt = fast_synthetic_batch(B, seed=1234)
    tok, tgt = tok.to(dev), tgt.to(dev)
    buckets = row_buckets(P, 1)
    for _ in range(2):
        model.fused_forward_loss(tok, tgt)
        backward_and_step(model, opt, buckets, 1)
    ctx = model._ctx
    lib, h = ctx.lib, ctx.handle
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    wstream = torch.cuda.Stream()
    st = lambda: _stream_ptr(dev)  # noqa: E731

    def ring(t, r0, r1, gptr, ctas, stages):
        ctx.check(lib.afr_adamw_rows_bg(h, *opt._hyper(), t, r0, r1, gptr, ctas, stages, st()))

    def timed(fn_main, fn_side=None, reps=7, side_first=True):
        res = []
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            m1 = torch.cuda.Event(enable_timing=True)
            e0.record()

            def run_side():
                side.wait_event(e0)
                with torch.cuda.stream(side):
                    s0.record()
                    fn_side()
                    s1.record()
            if fn_side is not None and side_first:
                run_side()
            if fn_main is not None:
                fn_main()
            m1.record()
            if fn_side is not None and not side_first:
                run_side()
            if fn_side is not None:
                main_s.wait_stream(side)
            e1.record()
            torch.cuda.synchronize()
            res.append((e0.elapsed_time(e1), e0.elapsed_time(m1),
                        s0.elapsed_time(s1) if fn_side is not None else 0.0))
        res.sort()
        return res[len(res) // 2]

    # make sure fc_output.weight.grad holds a real gradient for the stand-alone sweeps
    model.fused_forward_loss(tok, tgt)
    ctx.check(lib.afr_train_wgrad(h, 0, P, st()))
    ctx.check(lib.afr_train_dgrad(h, st()))

    def k_plain():
        t = opt.begin_step()
        opt.step_rows(t, 0, P)
        opt._bucket = None

    def k_ring(ctas, stages):
        def f():
            t = opt.begin_step()
            ring(t, 0, P, None, ctas, stages)
            opt._bucket = None
        return f

    if "e1" not in args.skip:
        print("E1 sweep alone (3.686 GB algorithmic incl. gradient read)")
        print("  plain adamw_kernel        %.3f ms" % timed(k_plain)[0])
        for ctas, stages in ((148, 4), (148, 6), (148, 8), (148, 12), (296, 3), (296, 4), (296, 6), (444, 4)):
            ms = timed(k_ring(ctas, stages))[0]
            print("  ring ctas=%3d stages=%2d    %.3f ms  (%.2f TB/s)" % (ctas, stages, ms, 3.686 / ms))
        sys.stdout.flush()

    def k_fwd():
        model.fused_forward_loss(tok, tgt)

    def k_fe_fwd():
        drop = model.make_dropout(B, 100)
        ctx.check(lib.afr_train_frontend(h, tok.data_ptr(), tok.stride(0), B, 100, C.byref(drop), st()))

    def k_wgrad():
        ctx.check(lib.afr_train_wgrad(h, 0, P, st()))

    def k_dgrad():
        ctx.check(lib.afr_train_dgrad_gemm(h, st()))

    def k_febwd():
        ctx.check(lib.afr_train_frontend_backward(h, st()))

    if "e2" not in args.skip:
        for reserve in (0, 33 * 1024):
            ctx.check(lib.afr_set_smem_reserve(h, reserve))
            print("E2 co-residency, GEMM smem reserve = %d KB" % (reserve // 1024))
            k_fwd()
            for name, fn in (("fe_fwd+gemm+loss", k_fwd), ("wgrad(full)", k_wgrad), ("dgrad gemm", k_dgrad),
                             ("fe_bwd", k_febwd)):
                alone = timed(fn)[0]
                line = "  %-18s alone %.3f |" % (name, alone)
                for ctas, stages in ((148, 4), (296, 3)):
                    ra = timed(k_ring(ctas, stages))[0]
                    tot, mn, sd = timed(fn, k_ring(ctas, stages), side_first=True)
                    line += " ring(%d,%d) alone %.3f: total %.3f (main %.3f ring %.3f) serial %.3f |" % (
                        ctas, stages, ra, tot, mn, sd, alone + ra)
                print(line)
                sys.stdout.flush()
            k_fwd()   # leave the context in the forward-done state
        ctx.check(lib.afr_set_smem_reserve(h, 0))

    # ------------------------------------------------------------------ E3: background-sweep step
    def step_with(o):
        def f():
            model.fused_forward_loss(tok, tgt)
            backward_and_step(model, o, buckets, 1)
        return f

    def time_steps(step, n=30, warm=5):
        for _ in range(warm):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            step()
        model.join_pending()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    if "e3" not in args.skip:
        print("E3 fused step (wgrad+AdamW epilogue): %.3f ms/step" % time_steps(step_with(opt)))
        sys.stdout.flush()
        for chunks, ctas, stages in ((1, 148, 4), (2, 148, 4), (4, 148, 4), (8, 148, 4), (15, 148, 4),
                                     (4, 148, 3), (4, 296, 3), (4, 296, 4), (4, 444, 3)):
            o2 = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), background=True,
                            bg_chunks=chunks, bg_ctas=ctas, bg_stages=stages)
            o2.state = opt.state
            try:
                ms = time_steps(step_with(o2))
                print("E3 background chunks=%2d ctas=%3d stages=%d: %.3f ms/step" % (chunks, ctas, stages, ms))
            except Exception as exc:   # noqa: BLE001
                print("E3 background chunks=%d failed: %s" % (chunks, exc))
            sys.stdout.flush()
        model.set_smem_reserve(0)


if __name__ == "__main__":
    main()
