"""tcgen05 GEMM rate vs tile width and issue mode (afr_gemm_bf16, K-major operands, fp32 out
through TMA stores) on the forward GEMM shape of the training step.
Used to separate the MMA's operand-fetch cost from its compute floor."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ai_font_renderer_b200 import _lib  # noqa: E402

lib = _lib.load()
dev = torch.device("cuda", 0)
K = 6400
for bn in (128, 192, 224, 256):
    for cta2 in (0, 1):             # flag bit 1 = CTA pairs
        M, N = 1024, 19200                    # the forward GEMM of the training step (compute-bound)
        A = torch.randn(M, K, device=dev).to(torch.bfloat16)
        B = torch.randn(N, K, device=dev).to(torch.bfloat16)
        D = torch.empty(M, N, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        def run():
            _lib.check(lib.afr_gemm_bf16(0, A.data_ptr(), K, 0, B.data_ptr(), K, 0, D.data_ptr(), N, M, N, K, bn,
                                         1.0, 1 | (cta2 << 1), st))
        for _ in range(3):
            run()
        best = 1e9
        for _ in range(10):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tf = 2.0 * M * N * K / best / 1e9
        model = (bn / 2) / (bn / 2 + 64)
        print(f"BN={bn:3d} mode={cta2}: {best:.4f} ms {tf:7.0f} TFLOP/s   (N/2)/(N/2+64) = {model:.2f}", flush=True)
