"""Stand-alone CUDA-event timing of the front-end kernels at the bench shape (B = 1024):
forward (training) and the three backward kernels together, plain and register-capped
('shared', the variants used beside the background AdamW sweep)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_font_renderer_b200.data import fast_synthetic_batch      # noqa: E402
from ai_font_renderer_b200.renderer import AttentionFontRenderer  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = AttentionFontRenderer().to(dev).train()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tok, tgt = fast_synthetic_batch(B)
tok, tgt = tok.to(dev), tgt.to(dev)


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for reserve in (0, 4 * 8192 + 1024):
    model.set_smem_reserve(reserve)
    loss = model.fused_forward_loss(tok, tgt)
    model._param_grads()
    model.dgrad_gemm()
    t_b = timed(model.frontend_backward)
    t_f = timed(lambda: model.fused_forward_loss(tok, tgt))
    print(f"smem reserve {reserve:6d}: forward+loss GEMM {t_f:.4f} ms, front-end backward (3 kernels + reduce) {t_b:.4f} ms")
