// Host side of the tcgen05 GEMM: TMA tensor-map encoding, tile-shape bookkeeping, launch.
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "afr_gemm.cuh"
#include "afr_internal.h"

namespace afr {
namespace {

using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                   const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                   const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

// 2-D row-major tensor [rows, cols] with leading dimension ld (elements); box = {box_cols, box_rows}.
bool encode_2d(CUtensorMap* tm, CUtensorMapDataType dt, int elem_bytes, const void* base,
               long long rows, long long cols, long long ld, int box_cols, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (fn == nullptr) return false;
  cuuint64_t gdim[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t gstride[1] = {static_cast<cuuint64_t>(ld) * elem_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, dt, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

template <int EPI, bool A_MN, bool B_MN, bool CTA2>
cudaError_t launch_impl(const GemmParams& p, int grid, cudaStream_t stream) {
  auto kern = gemm_bf16_tcgen05_kernel<EPI, A_MN, B_MN, CTA2>;
  const int smem_bytes = gemm_smem_bytes(p.stages, p.b_stage_bytes, p.epi_bytes, p.compact != 0);
  const int threads = EPI == kEpiAdamW ? 64 + 128 * p.adam_sub : kGemmThreads;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    cudaError_t e =
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemOptinMax);
    if (e != cudaSuccess) return e;
    // every kernel of the step asks for the largest shared-memory carveout: an SM whose carveout
    // had to change could not take a co-resident CTA of another kernel before it drained
    e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  if constexpr (!CTA2) {
    kern<<<grid, threads, smem_bytes, stream>>>(p);
    return cudaGetLastError();
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(grid));
    cfg.blockDim = dim3(static_cast<unsigned>(threads));
    cfg.dynamicSmemBytes = static_cast<size_t>(smem_bytes);
    cfg.stream = stream;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, p);
  }
}

template <int EPI, bool A_MN, bool B_MN>
cudaError_t launch_one(const GemmParams& p, int grid, cudaStream_t stream) {
  if (p.cta2) return launch_impl<EPI, A_MN, B_MN, true>(p, grid, stream);
  return launch_impl<EPI, A_MN, B_MN, false>(p, grid, stream);
}

}  // namespace

int gemm_num_tiles(int M, int N, int BN, bool cta2) {
  const int tile_m = cta2 ? 2 * kBM : kBM;
  return ((M + tile_m - 1) / tile_m) * ((N + BN - 1) / BN);
}

cudaError_t launch_gemm_bf16(const __nv_bfloat16* A, long long lda, bool a_mn,
                             const __nv_bfloat16* B, long long ldb, bool b_mn, int M, int N, int K,
                             int BN, const GemmEpilogue& epi, int num_sms, cudaStream_t stream,
                             int* num_tiles_out, const char** err_msg, int* k_splits_out) {
  static const char* kBadShape = "gemm: need M,N,K > 0, N % 32 == 0, BN % 32 == 0, 32 <= BN <= 256";
  static const char* kBadAlign = "gemm: operand pointers / leading dimensions must be 16-byte aligned";
  static const char* kBadMap = "gemm: cuTensorMapEncodeTiled failed";
  static const char* kBadEpi = "gemm: unsupported epilogue / operand-major combination";
  if (M <= 0 || N <= 0 || K <= 0 || (N % 32) != 0 || (BN % 32) != 0 || BN < 32 || BN > kMaxBN) {
    if (err_msg) *err_msg = kBadShape;
    return cudaErrorInvalidValue;
  }
  if ((reinterpret_cast<uintptr_t>(A) & 15) || (reinterpret_cast<uintptr_t>(B) & 15) ||
      (lda % 8) != 0 || (ldb % 8) != 0) {
    if (err_msg) *err_msg = kBadAlign;
    return cudaErrorInvalidValue;
  }
  if ((epi.kind == kEpiU8 || epi.kind == kEpiLoss) && (reinterpret_cast<uintptr_t>(epi.bias) & 15)) {
    if (err_msg) *err_msg = "gemm: the bias of the uint8 / loss epilogues must be 16-byte aligned";
    return cudaErrorInvalidValue;
  }
  GemmParams p{};
  bool ok = true;
  // CTA pairs: 256 x BN tiles, each CTA loads half of the B tile (box of BN / 2 rows)
  const bool cta2 = epi.cta2 != 0 && !epi.compact && num_sms >= 2;
  p.cta2 = cta2 ? 1 : 0;
  const int kctas = cta2 ? 2 : 1;
  if (!a_mn) ok &= encode_2d(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, M, K, lda, kBK, kBM);
  else       ok &= encode_2d(&p.tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, A, K, M, lda, 64, kBK);
  if (!b_mn) ok &= encode_2d(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, N, K, ldb, kBK, BN / kctas);
  else       ok &= encode_2d(&p.tm_b, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, B, K, N, ldb, 64, kBK);
  // split-K: whole 128-row tiles only (a tile must not reach into the next piece's rows)
  int k_splits = epi.k_splits > 1 ? epi.k_splits : 1;
  const int num_kb_all = (K + kBK - 1) / kBK;
  if (k_splits > num_kb_all) k_splits = num_kb_all;
  int kb_per_split = (num_kb_all + k_splits - 1) / k_splits;
  k_splits = (num_kb_all + kb_per_split - 1) / kb_per_split;
  if (k_splits > 1 && (epi.kind != kEpiF32 || cta2 || epi.bias != nullptr)) {
    if (err_msg) *err_msg = "gemm: split-K needs the fp32 epilogue without bias and single CTAs";
    return cudaErrorInvalidValue;
  }
  p.k_splits = k_splits;
  p.kb_per_split = kb_per_split;
  p.split_rows = ((M + kBM - 1) / kBM) * kBM;
  int use_tma_store = epi.use_tma_store;
  p.out_bf16 = 0;
  if (epi.kind == kEpiF32 && epi.out_bf16) {
    if ((reinterpret_cast<uintptr_t>(epi.out) & 15) || (epi.ldo % 8) != 0 || k_splits > 1) {
      if (err_msg) *err_msg = "gemm: bf16 output needs a 16-byte aligned pointer, ldo % 8 == 0 and no split-K";
      return cudaErrorInvalidValue;
    }
    p.out_bf16 = 1;
    use_tma_store = 0;
  }
  if (epi.kind == kEpiF32 && use_tma_store) {
    if ((reinterpret_cast<uintptr_t>(epi.out) & 15) || (epi.ldo % 4) != 0) {
      use_tma_store = 0;  // unaligned output: fall back to direct stores
    } else {
      ok &= encode_2d(&p.tm_c, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, epi.out,
                      k_splits > 1 ? static_cast<long long>(p.split_rows) * k_splits : M, N, epi.ldo, 32, 32);
    }
  }
  // ---- shared / tensor memory footprint
  // default: 4 x (16 + 32) KB operand ring, 32 KB of TMA-store slabs, all 512 TMEM columns.
  // compact (co-resident): tile at most 128 wide, 2 stages of 16 + 16 KB, direct stores, 256 TMEM
  // columns -- two such CTAs (this GEMM under the HBM-bound AdamW GEMM) share one SM.
  p.compact = epi.compact ? 1 : 0;
  // shared memory left to a co-resident background kernel (afr_set_smem_reserve): shallower ring
  const int smem_cap = kSmemOptinMax - (epi.smem_reserve > 0 ? epi.smem_reserve : 0);
  p.stages = kStages;
  while (p.stages > 2 && gemm_smem_bytes(p.stages, kBStageBytes, kEpiBytes, false) > smem_cap) --p.stages;
  p.b_stage_bytes = kBStageBytes;
  p.epi_bytes = kEpiBytes;
  p.tmem_cols = kTmemCols;
  if (cta2) {
    // half a B tile per CTA: the ring gets as deep as shared memory allows (6 stages of 32 KB at BN = 256)
    p.b_stage_bytes = b_mn ? ((BN / 2 + 63) / 64) * 8192 : (BN / 2) * kBK * 2;
    int stages = kMaxStages;
    while (stages > 2 && gemm_smem_bytes(stages, p.b_stage_bytes, p.epi_bytes, false) > smem_cap) --stages;
    p.stages = stages;
  }
  if (epi.compact) {
    if (BN > 128) {
      if (err_msg) *err_msg = "gemm: the co-resident footprint needs a tile at most 128 wide";
      return cudaErrorInvalidValue;
    }
    p.stages = 2;
    p.b_stage_bytes = b_mn ? ((BN + 63) / 64) * 8192 : BN * kBK * 2;
    p.tmem_cols = 256;
    if (epi.kind == kEpiF32) { use_tma_store = 0; p.epi_bytes = 0; }
  }
  if (epi.kind == kEpiAdamW) {
    static const char* kBadAdam =
        "gemm: fused AdamW epilogue needs MN-major operands, M % 32 == 0, 16-byte aligned p / m / v / "
        "bf16 copy and ldo % 8 == 0";
    if (!a_mn || !b_mn || (M % 32) != 0 || epi.adam_p == nullptr || epi.adam_m == nullptr ||
        epi.adam_v == nullptr || epi.out == nullptr || (epi.ldo % 8) != 0 ||
        ((reinterpret_cast<uintptr_t>(epi.adam_p) | reinterpret_cast<uintptr_t>(epi.adam_m) |
          reinterpret_cast<uintptr_t>(epi.adam_v) | reinterpret_cast<uintptr_t>(epi.out)) & 15)) {
      if (err_msg) *err_msg = kBadAdam;
      return cudaErrorInvalidValue;
    }
    ok &= encode_2d(&p.tm_p, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, epi.adam_p, M, N, epi.ldo, 32, 32);
    ok &= encode_2d(&p.tm_m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, epi.adam_m, M, N, epi.ldo, 32, 32);
    ok &= encode_2d(&p.tm_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, epi.adam_v, M, N, epi.ldo, 32, 32);
    p.hyper = epi.hyper;
    p.adam_ptr[0] = epi.adam_p; p.adam_ptr[1] = epi.adam_m; p.adam_ptr[2] = epi.adam_v;
    p.b_stage_bytes = ((BN / kctas + 63) / 64) * 8192;
    p.adam_sub = epi.adam_sub >= 1 && epi.adam_sub <= kMaxAdamSub ? epi.adam_sub : 2;
    p.adam_sets = epi.adam_sets >= 1 && epi.adam_sets <= kMaxAdamSets ? epi.adam_sets : 1;
    p.epi_bytes = 4 * p.adam_sub * p.adam_sets * kAdamSlabBytes;
    // deepest operand ring that still fits beside the slab sets
    int stages = epi.adam_stages > 0 ? epi.adam_stages : (epi.compact ? 2 : kMaxStages);
    if (stages > kMaxStages) stages = kMaxStages;
    while (stages > 1 &&
           gemm_smem_bytes(stages, p.b_stage_bytes, p.epi_bytes, p.compact != 0) > smem_cap)
      --stages;
    if (stages < 2) {
      if (err_msg) *err_msg = "gemm: fused AdamW epilogue: tile width / slab sets do not fit in shared memory";
      return cudaErrorInvalidValue;
    }
    p.stages = stages;
  }
  if (!ok) {
    if (err_msg) *err_msg = kBadMap;
    return cudaErrorInvalidValue;
  }
  p.M = M; p.N = N; p.K = K; p.BN = BN;
  p.num_m_tiles = (M + kBM * kctas - 1) / (kBM * kctas);
  p.num_n_tiles = (N + BN - 1) / BN;
  p.idesc = ptx::make_idesc_bf16(kBM * kctas, BN, a_mn, b_mn);
  p.out = epi.out; p.ldo = epi.ldo; p.bias = epi.bias; p.alpha = epi.alpha;
  p.clamp01 = epi.clamp01; p.use_tma_store = use_tma_store;
  p.target = epi.target; p.target_is_f32 = epi.target_is_f32; p.loss_partials = epi.loss_partials;
  const int num_tiles = p.num_m_tiles * p.num_n_tiles * p.k_splits;
  if (num_tiles_out) *num_tiles_out = num_tiles;
  if (k_splits_out) *k_splits_out = p.k_splits;
  int grid = num_tiles * kctas < num_sms ? num_tiles * kctas : num_sms;
  if (cta2) grid &= ~1;

  if (epi.kind == kEpiF32) {
    if (!a_mn && !b_mn) return launch_one<kEpiF32, false, false>(p, grid, stream);
    if (!a_mn && b_mn)  return launch_one<kEpiF32, false, true>(p, grid, stream);
    if (a_mn && !b_mn)  return launch_one<kEpiF32, true, false>(p, grid, stream);
    return launch_one<kEpiF32, true, true>(p, grid, stream);
  }
  if (epi.kind == kEpiU8 && !a_mn && !b_mn && epi.bias != nullptr)
    return launch_one<kEpiU8, false, false>(p, grid, stream);
  if (epi.kind == kEpiLoss && !a_mn && !b_mn && epi.bias != nullptr)
    return launch_one<kEpiLoss, false, false>(p, grid, stream);
  if (epi.kind == kEpiAdamW) return launch_one<kEpiAdamW, true, true>(p, grid, stream);
  if (err_msg) *err_msg = kBadEpi;
  return cudaErrorInvalidValue;
}

}  // namespace afr
