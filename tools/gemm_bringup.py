"""GPU bring-up check of the tcgen05 GEMM (afr_gemm_bf16) against torch.matmul in fp32.

Run on a B200: python tools/gemm_bringup.py
Covers K-major / MN-major operands, TMA store vs direct stores, ragged M / N / K, tile widths.
"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ai_font_renderer_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
torch.manual_seed(0)


def run(M, N, K, a_mn, b_mn, bn, tma, alpha=1.0):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    ref = alpha * (A.float() @ B.float().t())
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    D = torch.full((M, N), float("nan"), device=dev)
    st = torch.cuda.current_stream().cuda_stream
    rc = lib.afr_gemm_bf16(0, A_st.data_ptr(), A_st.stride(0), int(a_mn), B_st.data_ptr(),
                           B_st.stride(0), int(b_mn), D.data_ptr(), D.stride(0), M, N, K, bn,
                           alpha, int(tma), st)
    if rc != 0:
        return f"rc={rc} {lib.afr_last_error(None)}"
    try:
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        return f"CUDA error: {e}"
    err = (D - ref).abs().max().item()
    scale = ref.abs().max().item()
    nan = torch.isnan(D).sum().item()
    ok = nan == 0 and err <= 2e-3 * max(scale, 1.0)
    return ("ok " if ok else "BAD") + f" max_err={err:.3e} ref_max={scale:.2f} nan={nan}"


cases = []
for a_mn in (0, 1):
    for b_mn in (0, 1):
        cases.append((256, 512, 256, a_mn, b_mn, 256, 0))
for tma in (0, 1):
    cases += [
        (128, 256, 64, 0, 0, 256, tma),
        (1024, 1024, 512, 0, 0, 256, tma),
        (200, 320, 200, 0, 0, 128, tma),      # ragged M, K; N tail tile
        (192, 640, 19200 // 8, 0, 1, 192, tma),  # dgrad-like (B MN-major)
        (640, 640, 192, 1, 1, 256, tma),      # wgrad-like (both MN-major, K = batch)
        (304, 1280, 304, 1, 1, 224, tma),
        (1, 256, 640, 0, 0, 224, tma),        # batch 1 render
    ]
bad = 0
for c in cases:
    res = run(*c)
    print(c, res, flush=True)
    if not res.startswith("ok"):
        bad += 1
        if "CUDA error" in res:
            print("aborting after CUDA error"); sys.exit(2)

# timing at the real shapes (B = 1024)
def bench(M, N, K, a_mn, b_mn, bn, tma, iters=10):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    B = torch.randn(N, K, device=dev).to(torch.bfloat16)
    A_st = A.t().contiguous() if a_mn else A
    B_st = B.t().contiguous() if b_mn else B
    D = torch.empty((M, N), device=dev)
    st = torch.cuda.current_stream().cuda_stream
    args = (0, A_st.data_ptr(), A_st.stride(0), int(a_mn), B_st.data_ptr(), B_st.stride(0),
            int(b_mn), D.data_ptr(), D.stride(0), M, N, K, bn, 1.0, int(tma), st)
    for _ in range(3):
        lib.afr_gemm_bf16(*args)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters):
        lib.afr_gemm_bf16(*args)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    tf = 2.0 * M * N * K / ms / 1e9
    # cuBLAS for context
    for _ in range(3): torch.matmul(A, B.t())
    torch.cuda.synchronize(); e0.record()
    for _ in range(iters): torch.matmul(A, B.t())
    e1.record(); torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / iters
    return ms, tf, ms2, 2.0 * M * N * K / ms2 / 1e9

if bad == 0:
    for name, shp in [("fwd", (1024, 19200, 6400, 0, 0)), ("dgrad", (1024, 6400, 19200, 0, 1)),
                      ("wgrad", (19200, 6400, 1024, 1, 1))]:
        for bn in (256, 224, 192, 128):
            for tma in (1, 0):
                ms, tf, ms2, tf2 = bench(*shp, bn, tma)
                print(f"{name} BN={bn} tma_store={tma}: {ms:.3f} ms {tf:.0f} TFLOP/s | cuBLAS {ms2:.3f} ms {tf2:.0f} TFLOP/s", flush=True)
print("bad cases:", bad)
sys.exit(1 if bad else 0)
