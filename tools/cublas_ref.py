"""Library reference for the three GEMM shapes of one training step at B = 1024 (torch.matmul ->
cuBLASLt, bf16 operands): how fast does the vendor library run the same contractions?
Not part of the product path; prints ms and TFLOP/s per shape (best of 20, CUDA events)."""
import torch

dev = torch.device("cuda", 0)
B, K, P = 1024, 6400, 19200
shapes = {"forward  feats[B,K] x W[P,K]^T": ((B, K), (P, K), True),
          "dgrad    dZ[B,P]   x W[P,K]": ((B, P), (P, K), False),
          "wgrad    dZ[B,P]^T x feats[B,K]": ((B, P), (B, K), "tn")}
for name, (sa, sb, mode) in shapes.items():
    a = torch.randn(sa, device=dev, dtype=torch.bfloat16)
    b = torch.randn(sb, device=dev, dtype=torch.bfloat16)
    fn = (lambda: a @ b.t()) if mode is True else ((lambda: a @ b) if mode is False else (lambda: a.t() @ b))
    for _ in range(3):
        fn()
    best = 1e9
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    print(f"{name:36s} {best:.4f} ms  {2 * B * K * P / best / 1e9:.0f} TFLOP/s (bf16 out)")
