#!/usr/bin/env python
"""Per-phase cycle breakdown of the front-end forward kernel (profiling build; the backward was
instrumented the same way while it was one kernel: profiles/r02/frontend_phase_cycles_before_split.log).

    AFR_EXTRA_NVCC_FLAGS=-DAFR_PHASE_TIMING python tools/phase_timing.py [B]

Builds libafr_sm100.so with the phase counters compiled in, runs a few training steps at batch B
(default 1024) and prints, per kernel, the average cycles one CTA spent between consecutive marks
(thread 0's view: compute of its own warp, then the wait at the following barrier).
Rebuild without the flag afterwards (python -m ai_font_renderer_b200.build)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("AFR_EXTRA_NVCC_FLAGS", "-DAFR_PHASE_TIMING")

import torch  # noqa: E402

from ai_font_renderer_b200 import _lib  # noqa: E402
from ai_font_renderer_b200.data import fast_synthetic_batch  # noqa: E402
from ai_font_renderer_b200.renderer import AttentionFontRenderer  # noqa: E402

FWD = ["(loop top)", "1 embed", "  sync", "2 in-proj", "  sync", "3 attention", "  sync", "4a out-proj+LN",
       "4b fc1", "  sync"]

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda", 0)
    torch.manual_seed(42)
    model = AttentionFontRenderer().to(dev).train()
    tok, tgt = fast_synthetic_batch(B, seed=1234)
    tok, tgt = tok.to(dev), tgt.to(dev)
    lib = _lib.load()
    buf = (C.c_ulonglong * 32)()
    steps = 5
    for i in range(2 + steps):
        if i == 2:
            _lib.check(lib.afr_debug_phase_cycles(buf, 1))
        model.fused_train_step(tok, tgt)
    _lib.check(lib.afr_debug_phase_cycles(buf, 0))
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    mhz = 1965.0
    for name, labels, ctas in (("frontend_forward_kernel", FWD, min(B, 2 * sms)),):
        base = 0
        vals = [buf[base + k] / (steps * ctas) for k in range(16)]
        tot = sum(vals)
        print(f"{name}: {tot:.0f} cycles per CTA per launch = {tot / mhz:.1f} us at {mhz:.0f} MHz")
        for k, lab in enumerate(labels):
            print(f"   {lab:28s} {vals[k]:10.0f} cyc  {100 * vals[k] / tot:5.1f}%")


if __name__ == "__main__":
    main()
