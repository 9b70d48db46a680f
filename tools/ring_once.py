"""One training step at the bench shape, then the plain AdamW sweep and the background (ring) sweep
alone, for profiling with ncu:
    ncu -k regex:adamw_ring -c 1 --set full --import-source on ... python tools/ring_once.py [ctas] [stages]
"""
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
opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99), fuse_wgrad=False)
ctas = int(sys.argv[1]) if len(sys.argv) > 1 else 148
stages = int(sys.argv[2]) if len(sys.argv) > 2 else 4
tok, tgt = fast_synthetic_batch(1024)
model.fused_train_step(tok.to(dev), tgt.to(dev))
for _ in range(2):
    t = opt.begin_step()
    opt.step_rows(t, 0, 19200)
    opt.end_step()
    t = opt.begin_step()
    opt.step_rows_bg(t, 0, 19200, ctas, stages)
    opt.end_step()
torch.cuda.synchronize()
print("ok")
