"""Two training steps at the bench shape (B = 1024), for profiling single kernels with ncu:
    ncu -k regex:gemm_bf16 -s 3 -c 3 ... python tools/step_once.py     # forward, wgrad+AdamW, dgrad of step 2
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_font_renderer_b200.data import fast_synthetic_batch      # noqa: E402
from ai_font_renderer_b200.optim import FusedAdamW               # noqa: E402
from ai_font_renderer_b200.renderer import AttentionFontRenderer  # noqa: E402
from ai_font_renderer_b200.training import backward_and_step, row_buckets  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(42)
model = AttentionFontRenderer().to(dev).train()
opt = FusedAdamW(model, lr=1e-3, weight_decay=5e-4, betas=(0.9, 0.99))
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tok, tgt = fast_synthetic_batch(B)
tok, tgt = tok.to(dev), tgt.to(dev)
for _ in range(2):
    loss = model.fused_forward_loss(tok, tgt)
    backward_and_step(model, opt, row_buckets(19200, 1), 1)
torch.cuda.synchronize()
print("loss", float(loss))
