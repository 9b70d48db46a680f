"""Which front-end kernel costs the background AdamW sweep how much, and vice versa (B = 1024).
For each kernel X in {front-end forward, backward head / attention / tail}: the sweep on a side
stream and X launched back to back on the compute stream until the sweep ends; prints the sweep's
duration, X's average duration beside it, and both alone."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_font_renderer_b200.data import fast_synthetic_batch      # noqa: E402
from ai_font_renderer_b200.optim import FusedAdamW               # noqa: E402
from ai_font_renderer_b200.renderer import AttentionFontRenderer  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = AttentionFontRenderer().to(dev).train()
stages = int(sys.argv[1]) if len(sys.argv) > 1 else 8
opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), background=True, bg_stages=stages)
B = 1024
tok, tgt = fast_synthetic_batch(B)
tok, tgt = tok.to(dev), tgt.to(dev)
model.set_smem_reserve(stages * 8192 + 1024)
loss = model.fused_forward_loss(tok, tgt)
model._param_grads()
model.wgrad_rows(0, 19200)
model.dgrad_gemm()
model.frontend_backward()
torch.cuda.synchronize()
side = model.side_stream()
main = torch.cuda.current_stream()


def ev():
    return torch.cuda.Event(enable_timing=True)


def sweep(t):
    opt.step_rows_bg(t, 0, 19200, 0, stages)


def run(name, fn, n_beside):
    # alone
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a0, a1 = ev(), ev()
    a0.record()
    for _ in range(5):
        fn()
    a1.record()
    torch.cuda.synchronize()
    alone = a0.elapsed_time(a1) / 5
    res = []
    for rep in range(3):
        t = opt.begin_step()
        torch.cuda.synchronize()
        s0, s1, k0, k1 = ev(), ev(), ev(), ev()
        start = ev()
        start.record(main)
        with torch.cuda.stream(side):
            side.wait_event(start)
            s0.record(side)
            sweep(t)
            s1.record(side)
        k0.record(main)
        for _ in range(n_beside):
            fn()
        k1.record(main)
        torch.cuda.synchronize()
        model.join_pending()
        opt.end_step()
        res.append((s0.elapsed_time(s1), k0.elapsed_time(k1) / n_beside))
    sw = min(r[0] for r in res)
    kx = min(r[1] for r in res)
    print(f"{name:28s} alone {alone:.4f} ms | beside the sweep {kx:.4f} ms (x{kx / alone:.2f}) | sweep {sw:.4f} ms "
          f"({n_beside} launches = {n_beside * kx:.3f} ms)")


t = opt.begin_step()
torch.cuda.synchronize()
s0, s1 = ev(), ev()
s0.record()
sweep(t)
s1.record()
torch.cuda.synchronize()
model.join_pending()
opt.end_step()
print(f"sweep alone ({stages} stages): {s0.elapsed_time(s1):.4f} ms")

lib, ctx = model._ctx.lib, model._ctx


def bwd_only(k):
    def f():
        os.environ["AFR_FE_BWD_ONLY"] = str(k)
        model.frontend_backward()
        os.environ.pop("AFR_FE_BWD_ONLY")
    return f


import ctypes as C  # noqa: E402
from ai_font_renderer_b200.renderer import _stream_ptr  # noqa: E402
drop = model.make_dropout(B, tok.shape[1], enabled=True)


def fwd():
    ctx.check(lib.afr_train_frontend(ctx.handle, tok.data_ptr(), tok.stride(0), B, tok.shape[1], C.byref(drop),
                                     _stream_ptr(dev)))


run("backward head (K1)", bwd_only(1), 8)
run("backward attention (K2)", bwd_only(2), 4)
run("backward tail (K3)", bwd_only(3), 8)
run("front-end forward", fwd, 3)
