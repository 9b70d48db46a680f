#!/usr/bin/env python
"""How well does the HBM-bound AdamW sweep share the GPU with each compute kernel of the step?
Times, with CUDA events: each kernel alone, then the pair launched on two streams."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from ai_font_renderer_b200.data import fast_synthetic_batch  # noqa: E402
from ai_font_renderer_b200.optim import FusedAdamW  # noqa: E402
from ai_font_renderer_b200.renderer import AttentionFontRenderer, _stream_ptr  # noqa: E402

P = 19200


def main():
    B = 1024
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = AttentionFontRenderer().to(dev).train()
    opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99))
    tok, tgt = fast_synthetic_batch(B, seed=1234)
    tok, tgt = tok.to(dev), tgt.to(dev)
    for _ in range(2):
        model.fused_train_step(tok, tgt)
        opt.step()
    ctx = model._ctx
    main_s = torch.cuda.current_stream()
    side = torch.cuda.Stream()
    drop = model.make_dropout(B, 100)
    dfeat = torch.randn(B, 6400, device=dev) * 1e-3

    def k_fwd():
        model.fused_forward_loss(tok, tgt)

    def k_wgrad():
        ctx.check(ctx.lib.afr_train_wgrad(ctx.handle, 0, P, _stream_ptr(dev)))

    def k_dgrad_febwd():
        ctx.check(ctx.lib.afr_train_dgrad(ctx.handle, _stream_ptr(dev)))

    def k_adam():
        t = opt.begin_step()
        opt.step_rows(t, 0, P)
        opt.end_step()

    def timed(fn_main, fn_side=None, reps=5):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if fn_side is not None:
                side.wait_event(e0)
                with torch.cuda.stream(side):
                    s0.record()
                    fn_side()
                    s1.record()
            if fn_main is not None:
                fn_main()
            if fn_side is not None:
                main_s.wait_stream(side)
            e1.record()
            torch.cuda.synchronize()
            r = (e0.elapsed_time(e1), s0.elapsed_time(s1) if fn_side is not None else 0.0)
            best = r if best is None or r[0] < best[0] else best
        return best

    k_fwd()
    print("alone: adamw %.3f ms" % timed(k_adam)[0])
    for name, fn in (("forward(front-end+GEMM+loss)", k_fwd), ("wgrad", k_wgrad),
                     ("dgrad+front-end backward", k_dgrad_febwd)):
        a = timed(fn)[0]
        both, adam_in = timed(fn, k_adam)
        print(f"{name:32s} alone {a:.3f} ms | with adamw on the side stream: total {both:.3f} ms "
              f"(adamw took {adam_in:.3f}); serial would be {a + timed(k_adam)[0]:.3f}")


if __name__ == "__main__":
    main()
