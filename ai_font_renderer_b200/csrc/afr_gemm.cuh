// Persistent, warp-specialised tcgen05 GEMM for the fc_output contraction and its two
// gradients (reference: model.py:196 forward, autograd of it at model.py:309).
//
//   D[M,N] = A[M,K] * B[N,K]^T      bf16 operands, fp32 accumulation in TMEM
//
// Each operand can be "K-major" (K contiguous in global memory) or "MN-major" (M resp. N
// contiguous), so the three GEMMs of a training step all read the natural row-major
// tensors with no transposed copies:
//   forward : A = feats [B,6400]  (K-major)   B = W   [19200,6400] (K-major)
//   dgrad   : A = dZ    [B,19200] (K-major)   B = W   [19200,6400] (MN-major, K = pixel)
//   wgrad   : A = dZ    [B,19200] (MN-major)  B = feats [B,6400]   (MN-major, K = batch)
//
// Structure (one CTA per SM, 192 threads):
//   warp 0      : TMA producer  (cp.async.bulk.tensor -> 128B-swizzled smem ring)
//   warp 1      : TMEM allocator + single-thread tcgen05.mma issuer
//   warps 2..5  : epilogue (tcgen05.ld -> registers -> fused epilogue -> global / TMA store)
// Three mbarrier pipelines: smem full/empty, TMEM full/empty (two accumulator buffers so
// the epilogue of tile i overlaps the MMAs of tile i+1), and a static persistent tile loop.
//
// kEpiAdamW is the wgrad GEMM with optimizer.step() of fc_output.weight (model.py:310) folded
// into its epilogue: the accumulator IS the gradient, so instead of writing 491 MB of dW and
// reading it back in a separate sweep, each epilogue warp streams the matching 32x32 tiles of
// p / exp_avg / exp_avg_sq through shared memory with TMA (loads issued two chunks ahead, so
// they run under the MMAs of the tile), applies torch's AdamW arithmetic in place and stores
// p, m, v (TMA) and the bf16 weight copy. That kernel is HBM-bound (26 B/parameter); the operand
// ring shrinks to two stages to make room for the staging slabs.
#pragma once
#include "afr_internal.h"
#include "afr_ptx.cuh"

namespace afr {

constexpr int kBM = 128;                       // UMMA M (cta_group::1)
constexpr int kBK = 64;                        // k-block: 64 bf16 (one 128B swizzle row)
constexpr int kMaxBN = 256;                    // UMMA N upper bound
constexpr int kStages = 4;
constexpr int kAStageBytes = kBM * kBK * 2;    // 16 KB
constexpr int kBStageBytes = kMaxBN * kBK * 2; // 32 KB
constexpr int kStageBytes = kAStageBytes + kBStageBytes;
constexpr int kEpiWarpBufBytes = 32 * 128;     // one 32-row x 128-byte staging slab
constexpr int kEpiBytes = 4 * 2 * kEpiWarpBufBytes;  // 4 warps x 2 buffers = 32 KB
constexpr int kBarrierBytes = 512;
constexpr int kGemmSmemBytes = kStages * kStageBytes + kEpiBytes + kBarrierBytes + 1024;
// kEpiAdamW: per epilogue warp `adam_sets` (2..4) slab sets of three 32x32 fp32 tiles (p, exp_avg,
// exp_avg_sq); the operand ring has `stages` (2..8) stages of 16 KB + b_stage_bytes. Both are
// chosen by the host so that the total fits the 227 KB of shared memory (adam_smem_bytes).
constexpr int kAdamSlabBytes = 3 * kEpiWarpBufBytes;             // 12 KB
constexpr int kMaxStages = 8;
constexpr int kMaxAdamSets = 2;
constexpr int kMaxAdamSub = 3;                                   // epilogue warps per lane quadrant
constexpr int kAdamMaxThreads = 64 + 128 * kMaxAdamSub;
// dynamic shared memory of a launch: operand ring + epilogue region + barriers (+ 1 KB alignment
// slack unless the co-resident footprint is requested)
__host__ __device__ constexpr int gemm_smem_bytes(int stages, int b_stage_bytes, int epi_bytes,
                                                  bool compact) {
  return stages * (kAStageBytes + b_stage_bytes) + epi_bytes + kBarrierBytes + (compact ? 0 : 1024);
}
constexpr int kSmemOptinMax = 232448;                            // 227 KB per CTA on sm_100
constexpr int kGemmThreads = 192;
constexpr int kTmemCols = 512;                 // 2 accumulator buffers x 256 columns

enum EpiKind : int {
  kEpiF32 = 0,   // out_f32 = [clamp01](alpha*acc + bias[n])        (logits, dA, dW, eval sheet)
  kEpiU8 = 1,    // out_u8  = trunc(clamp01(acc + bias[n]) * 255)   (helpers.py:33 quantisation)
  kEpiLoss = 2,  // clamp + MSE partial sums + masked residual dZ   (model.py:202,270 + backward)
  kEpiAdamW = 3, // g = alpha*acc never leaves the SM: AdamW on p/m/v tiles + bf16 copy (model.py:310)
};

struct GemmParams {
  CUtensorMap tm_a;
  CUtensorMap tm_b;
  CUtensorMap tm_c;      // fp32 store map (kEpiF32 with use_tma_store)
  int M, N, K;
  int BN;                // tile N, multiple of 32, <= 256
  int num_m_tiles, num_n_tiles;
  uint32_t idesc;
  // epilogue
  void* out;             // f32 / u8 / bf16 (dZ) row-major [M, ldo]
  long long ldo;
  const float* bias;     // [N] or nullptr
  float alpha;
  int clamp01;
  int use_tma_store;
  const void* target;    // kEpiLoss: [M, N] u8 (k/255) or f32
  int target_is_f32;
  float* loss_partials;  // kEpiLoss: [num_tiles * 4] per-warp sums of (y - t)^2
  // kEpiAdamW: fp32 [M, N] maps (32x32 boxes, 128B swizzle) of the parameter and its moments,
  // used for both the loads and the in-place stores; out = bf16 copy [M, ldo]
  CUtensorMap tm_p, tm_m, tm_v;
  AdamHyper hyper;
  float* adam_ptr[3];    // p, exp_avg, exp_avg_sq base pointers (the epilogue's direct stores)
  // shared / tensor memory footprint (set by the host, launch_gemm_bf16)
  int stages;            // operand ring depth (2..kMaxStages)
  int b_stage_bytes;     // bytes of one B-operand stage
  int epi_bytes;         // epilogue staging region (TMA-store slabs / AdamW load slabs; may be 0)
  int tmem_cols;         // TMEM columns to allocate: 512, or 256 in the co-resident footprint
  int cta2;              // 1: launched as CTA pairs (template CTA2)
  int compact;           // 1: no alignment slack in the dynamic shared memory (base must be 1 KB aligned)
  int adam_sets;         // slab sets per epilogue warp (threads = 64 + 128 * warps per quadrant)
  int adam_sub;          // epilogue warps per TMEM lane quadrant (1..kMaxAdamSub)
  // split-K (kEpiF32 only; 1 = off): the K range is cut into k_splits pieces of kb_per_split
  // k-blocks, piece ks of tile (m, n) accumulates into rows [ks * M, ks * M + M) of an output of
  // k_splits * M rows (summed by a second kernel in a fixed order). For the weight gradients of
  // the wide front-end, whose M x N is one to three tiles and whose K is batch x positions.
  int k_splits;
  int kb_per_split;
  int split_rows;        // rows per K piece in the output: M rounded up to a whole tile
  int out_bf16;          // kEpiF32: the output is bf16 [M, ldo] (direct 64-byte row stores), not fp32
};

// CTA2: the kernel runs as clusters of two CTAs (one TPC) that execute 256-row MMAs together
// (tcgen05 cta_group::2). A tile is then 256 x BN per pair; CTA rank r loads rows [128 r, 128 r +
// 128) of A and columns [r BN/2, (r+1) BN/2) of B, so a stage costs 16 KB + BN/2 x 128 B of shared
// memory per CTA instead of 16 KB + BN x 128 B for the same tensor work per SM: a third less
// L2 -> shared-memory traffic and room for a deeper ring. Only the leader (rank 0) issues MMAs; its
// full barriers count the bytes of both CTAs' TMA loads, its commits arrive on the empty /
// accumulator-full barriers of both CTAs, and both CTAs' epilogue threads arrive on the leader's
// accumulator-empty barriers. Each CTA's epilogue drains its own 128 TMEM lanes.
template <int EPI, bool A_MN, bool B_MN, bool CTA2>
// (the AdamW instantiation is bounded as if it had 512 threads: that caps it at 128 registers, so
// that its 320 threads and a 192-thread GEMM CTA fit the register file of one SM together)
__global__ void __launch_bounds__(EPI == kEpiAdamW ? 512 : kGemmThreads, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  // 128B swizzle atoms are 1024-byte aligned. The co-resident footprint has no room for
  // alignment slack: the dynamic window of a kernel without static shared memory starts 1 KB
  // aligned, which is checked instead of assumed.
  if (p.compact && (ptx::smem_u32(smem_raw) & 1023u) != 0u) __trap();
  uint8_t* smem = p.compact ? smem_raw
                            : reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                                         ~static_cast<uintptr_t>(1023));
  const int kStages = p.stages;
  const int kBStage = p.b_stage_bytes;
  const int epi_region = p.epi_bytes;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * kAStageBytes;
  uint8_t* smem_epi = smem_b + kStages * kBStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + epi_region);
  uint64_t* full_bar = bars;                          // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;            // [kMaxStages]
  uint64_t* tmem_full_bar = bars + 2 * kMaxStages;    // [4]
  uint64_t* tmem_empty_bar = tmem_full_bar + 4;       // [4]
  uint64_t* adam_ld_bar = tmem_empty_bar + 4;         // [kMaxAdamSub * 4 warps][kMaxAdamSets]   (kEpiAdamW)
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(adam_ld_bar + 4 * kMaxAdamSub * kMaxAdamSets);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  constexpr int kCtas = CTA2 ? 2 : 1;
  const int cta_rank = CTA2 ? static_cast<int>(ptx::cluster_ctarank()) : 0;
  const int first_tile = static_cast<int>(blockIdx.x) / kCtas;     // pairs walk the tile list
  const int tile_step = static_cast<int>(gridDim.x) / kCtas;
  const int row_base = cta_rank * kBM;                             // this CTA's rows inside a tile
  constexpr int kTileM = kBM * kCtas;
  const int tiles_mn = p.num_m_tiles * p.num_n_tiles;
  const int num_tiles = tiles_mn * p.k_splits;
  const int num_kb_all = (p.K + kBK - 1) / kBK;
  // k-blocks of the K piece a tile works on (the last piece may be shorter, never empty)
  auto kb_count = [&](int tile) {
    if (p.k_splits == 1) return num_kb_all;
    const int ks = tile / tiles_mn;
    return min(p.kb_per_split, num_kb_all - ks * p.kb_per_split);
  };
  const int BN = p.BN;
  // TMEM accumulator buffers of 256 columns, or 128 when the tile is at most 128 wide: 2 or 4 of
  // them in the full 512-column allocation, 2 x 128 in the co-resident footprint (256 columns, so
  // that two kernels can share an SM's tensor memory)
  const int acc_stride = BN <= 128 ? 128 : 256;
  const int n_acc = p.tmem_cols / acc_stride;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&p.tm_a);
    ptx::prefetch_tensormap(&p.tm_b);
    if (EPI == kEpiF32 && p.use_tma_store) ptx::prefetch_tensormap(&p.tm_c);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 4; ++a) {
      ptx::mbar_init(&tmem_full_bar[a], 1);
      ptx::mbar_init(&tmem_empty_bar[a], (blockDim.x - 64) * kCtas);   // every epilogue thread (of the pair)
    }
    if (EPI == kEpiAdamW) {
      ptx::prefetch_tensormap(&p.tm_p);
      ptx::prefetch_tensormap(&p.tm_m);
      ptx::prefetch_tensormap(&p.tm_v);
      for (int i = 0; i < 4 * kMaxAdamSub * kMaxAdamSets; ++i) ptx::mbar_init(&adam_ld_bar[i], 1);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (CTA2) {
      ptx::tmem_alloc_pair(tmem_base_slot, static_cast<uint32_t>(p.tmem_cols));
      ptx::tmem_relinquish_pair();
    } else {
      ptx::tmem_alloc(tmem_base_slot, static_cast<uint32_t>(p.tmem_cols));
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (CTA2) ptx::cluster_sync();   // the peer's barriers are initialised before anyone signals them
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // an epilogue thread is done with accumulator buffer `a`: tell the (leader's) MMA warp
  auto release_acc = [&](int a) {
    if constexpr (CTA2) ptx::mbar_arrive_cluster(ptx::map_to_cta(&tmem_empty_bar[a], 0));
    else ptx::mbar_arrive(&tmem_empty_bar[a]);
  };

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (whole warp, one lane issues)
    {
      const bool leader = ptx::elect_one();
      const int bn_cta = BN / kCtas;        // B columns this CTA loads
      const uint32_t a_bytes = kAStageBytes;
      const uint32_t b_boxes = B_MN ? static_cast<uint32_t>((bn_cta + 63) / 64) : 0u;
      const uint32_t b_bytes = B_MN ? b_boxes * 8192u : static_cast<uint32_t>(bn_cta) * kBK * 2u;
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        const int m0 = (tile % p.num_m_tiles) * kTileM + row_base;
        const int n0 = ((tile / p.num_m_tiles) % p.num_n_tiles) * BN + cta_rank * bn_cta;
        const int kb0 = (tile / tiles_mn) * p.kb_per_split;
        const int num_kb = kb_count(tile);
        for (int kbl = 0; kbl < num_kb; ++kbl) {
          const int kb = kb0 + kbl;
          ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem_a + stage * kAStageBytes;
          uint8_t* sb = smem_b + stage * kBStage;
          // the fused-AdamW kernel streams 3 GB through L2 next to its operands: keep them
          constexpr uint64_t pol = EPI == kEpiAdamW ? ptx::kL2EvictLast : ptx::kL2EvictNormal;
          auto load = [&](const CUtensorMap* tm, void* dst, int c0, int c1) {
            if constexpr (CTA2) ptx::tma_load_2d_pair(tm, ptx::map_to_cta(&full_bar[stage], 0), dst, c0, c1, pol);
            else ptx::tma_load_2d_hint(tm, &full_bar[stage], dst, c0, c1, pol);
          };
          if (leader) {
            // the leader CTA's barrier counts the bytes of both CTAs
            if (cta_rank == 0) ptx::mbar_arrive_expect_tx(&full_bar[stage], (a_bytes + b_bytes) * kCtas);
            if constexpr (!A_MN) {
              load(&p.tm_a, sa, kb * kBK, m0);
            } else {
#pragma unroll
              for (int i = 0; i < kBM / 64; ++i) load(&p.tm_a, sa + i * 8192, m0 + i * 64, kb * kBK);
            }
            if constexpr (!B_MN) {
              load(&p.tm_b, sb, kb * kBK, n0);
            } else {
              for (uint32_t i = 0; i < b_boxes; ++i) load(&p.tm_b, sb + i * 8192, n0 + i * 64, kb * kBK);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (whole warp, one lane issues)
    if (cta_rank == 0) {
      const bool leader = ptx::elect_one();
      // K-major  : 8-row groups 1024 B apart (SBO); 16 K-elements = 32 B inside the swizzle row.
      // MN-major : 64-element MN chunks 8192 B apart (LBO); 8-row K groups 1024 B apart (SBO);
      //            16 K-rows = 2048 B.
      const uint32_t a_lbo = A_MN ? 8192u : 16u, b_lbo = B_MN ? 8192u : 16u;
      const uint32_t a_kstep = (A_MN ? 2048u : 32u) >> 4, b_kstep = (B_MN ? 2048u : 32u) >> 4;
      int stage = 0;
      uint32_t phase = 0;
      int acc = 0;
      uint32_t acc_phase = 0;
      for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
        ptx::mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * acc_stride);
        const int num_kb = kb_count(tile);
        for (int kb = 0; kb < num_kb; ++kb) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint64_t da = ptx::make_smem_desc_sw128(
              ptx::smem_u32(smem_a + stage * kAStageBytes), a_lbo, 1024u);
          const uint64_t db = ptx::make_smem_desc_sw128(
              ptx::smem_u32(smem_b + stage * kBStage), b_lbo, 1024u);
          if (leader) {
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k) {
              if constexpr (CTA2)
                ptx::umma_bf16_pair(d_tmem, da + static_cast<uint64_t>(k * a_kstep),
                                    db + static_cast<uint64_t>(k * b_kstep), p.idesc,
                                    (kb | k) != 0 ? 1u : 0u);
              else
                ptx::umma_bf16(d_tmem, da + static_cast<uint64_t>(k * a_kstep),
                               db + static_cast<uint64_t>(k * b_kstep), p.idesc,
                               (kb | k) != 0 ? 1u : 0u);
            }
            // smem slot reusable (in both CTAs) once these MMAs retire
            if constexpr (CTA2) {
              ptx::umma_commit_pair(&empty_bar[stage]);
              if (kb == num_kb - 1) ptx::umma_commit_pair(&tmem_full_bar[acc]);
            } else {
              ptx::umma_commit(&empty_bar[stage]);
              if (kb == num_kb - 1) ptx::umma_commit(&tmem_full_bar[acc]);
            }
          }
          if (++stage == kStages) { stage = 0; phase ^= 1u; }
        }
        if (++acc == n_acc) { acc = 0; acc_phase ^= 1u; }
      }
    }
  } else if constexpr (EPI == kEpiAdamW) {
    // ------------------------------------------------------------ epilogue warps, fused AdamW
    // 4 * nsub warps: warp (q, sub) owns the 32 rows of TMEM lane quadrant q = warp % 4 and the
    // 32-column chunks c = sub (mod nsub) of every tile. Per chunk the p / exp_avg / exp_avg_sq
    // tiles arrive by TMA in one of the warp's slab sets (128B swizzle, row = lane); the warp
    // moves them to registers and immediately refills the set with a later chunk, so the sets are
    // load buffers that are in flight almost all the time. The update runs in registers and
    // leaves through 32-byte streaming stores (every lane writes full sectors of its own row).
    // Loads never depend on the MMAs, so they also run while the warp waits for an accumulator.
    const int q = warp & 3;
    const int sub = (warp - 2) >> 2;
    const int nsub = (static_cast<int>(blockDim.x) - 64) >> 7;
    const int sets = p.adam_sets;
    uint8_t* slabs = smem_epi + (sub * 4 + q) * (sets * kAdamSlabBytes);
    uint64_t* ld_bar = adam_ld_bar + (sub * 4 + q) * kMaxAdamSets;
    const AdamPairConst hc(p.hyper);
    auto chunks_of = [&](int t) {
      const int n0 = (t / p.num_m_tiles) * BN;
      return (min(BN, p.N - n0)) >> 5;
    };
    // next tile >= t in which this warp has work: its 32 rows inside M, a chunk index == sub
    auto next_active = [&](int t) {
      while (t < num_tiles &&
             ((t % p.num_m_tiles) * kTileM + row_base + q * 32 >= p.M || chunks_of(t) <= sub))
        t += tile_step;
      return t;
    };
    // loader cursor (lane 0 issues; every lane tracks it)
    int lt = next_active(first_tile), lc = sub, s_issue = 0;
    int l_chunks = 0, l_col0 = 0, l_row = 0;
    auto load_tile_coords = [&]() {
      if (lt >= num_tiles) return;
      l_chunks = chunks_of(lt);
      l_col0 = (lt / p.num_m_tiles) * BN;
      l_row = (lt % p.num_m_tiles) * kTileM + row_base + q * 32;
    };
    load_tile_coords();
    auto issue_next = [&](uint32_t z) {     // z: 0, from ptx::warp_reads_done (orders the refill)
      if (lt >= num_tiles) return;
      if (lane == 0) {
        const int col = l_col0 + lc * 32;
        uint8_t* dst = slabs + s_issue * kAdamSlabBytes + z;
        ptx::mbar_arrive_expect_tx(&ld_bar[s_issue], kAdamSlabBytes);
        ptx::tma_load_2d_hint(&p.tm_p, &ld_bar[s_issue], dst, col, l_row, ptx::kL2EvictFirst);
        ptx::tma_load_2d_hint(&p.tm_m, &ld_bar[s_issue], dst + kEpiWarpBufBytes, col, l_row,
                              ptx::kL2EvictFirst);
        ptx::tma_load_2d_hint(&p.tm_v, &ld_bar[s_issue], dst + 2 * kEpiWarpBufBytes, col, l_row,
                              ptx::kL2EvictFirst);
      }
      if (++s_issue == sets) s_issue = 0;
      lc += nsub;
      if (lc >= l_chunks) { lc = sub; lt = next_active(lt + tile_step); load_tile_coords(); }
    };
    for (int i = 0; i < sets; ++i) issue_next(0u);
    int acc = 0;
    uint32_t acc_phase = 0;
    int s = 0;                              // slab set of the chunk being consumed
    uint32_t s_phase = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int m0 = (tile % p.num_m_tiles) * kTileM + row_base;
      const int n0 = (tile / p.num_m_tiles) * BN;
      const long long m = m0 + q * 32 + lane;
      const int n_chunks = chunks_of(tile);
      const bool active = m0 + q * 32 < p.M && n_chunks > sub;
      bool acc_ready = false;
      for (int c = sub; active && c < n_chunks; c += nsub) {
        // p, m, v of this chunk wait in slab set s; they are moved to registers and updated in
        // two halves of 16 columns (48 + 16 live values per thread instead of 96 + 32), and the
        // set is refilled as soon as the second half has been read
        const uint32_t sp = ptx::smem_u32(slabs + s * kAdamSlabBytes + lane * 128);
        ptx::mbar_wait(&ld_bar[s], s_phase);
        if (!acc_ready) {
          ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
          ptx::tc_fence_after();
          acc_ready = true;
        }
        const long long e = m * p.ldo + n0 + c * 32;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          float4 pv[4], mv[4], vv[4];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const int off = ((half * 4 + ch) ^ (lane & 7)) << 4;
            pv[ch] = ptx::lds_f4(sp + off);
            mv[ch] = ptx::lds_f4(sp + kEpiWarpBufBytes + off);
            vv[ch] = ptx::lds_f4(sp + 2 * kEpiWarpBufBytes + off);
          }
          if (half == 1) {
            // the refill is an async-proxy write into memory this warp has just read through the
            // generic proxy: every lane's reads must have been PERFORMED (not merely issued)
            // before the TMA load is launched. A vote over the loaded registers guarantees that
            // (ptx::warp_reads_done); the proxy fence used here before also did, but its
            // MEMBAR.ALL.CTA made every chunk wait for the previous chunk's streaming stores.
            uint32_t dep = 0;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              dep ^= __float_as_uint(pv[ch].w) ^ __float_as_uint(mv[ch].w) ^ __float_as_uint(vv[ch].w);
            issue_next(ptx::warp_reads_done(dep));
            if (++s == sets) { s = 0; s_phase ^= 1u; }
          }
          uint32_t r[16];
          ptx::tmem_ld_16(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                              static_cast<uint32_t>(acc * acc_stride + c * 32 + half * 16),
                          r);
          ptx::tmem_ld_wait();
          if (half == 1 && c + nsub >= n_chunks) {   // this warp's last read of the accumulator buffer
            ptx::tc_fence_before();
            release_acc(acc);
          }
          uint32_t packed[8];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            const float4 g4 = make_float4(__fmul_rn(__uint_as_float(r[4 * ch]), p.alpha),
                                          __fmul_rn(__uint_as_float(r[4 * ch + 1]), p.alpha),
                                          __fmul_rn(__uint_as_float(r[4 * ch + 2]), p.alpha),
                                          __fmul_rn(__uint_as_float(r[4 * ch + 3]), p.alpha));
            adamw_quad(pv[ch], g4, mv[ch], vv[ch], hc);
            const __nv_bfloat162 lo = __floats2bfloat162_rn(pv[ch].x, pv[ch].y);
            const __nv_bfloat162 hi = __floats2bfloat162_rn(pv[ch].z, pv[ch].w);
            packed[2 * ch] = *reinterpret_cast<const uint32_t*>(&lo);
            packed[2 * ch + 1] = *reinterpret_cast<const uint32_t*>(&hi);
          }
          float* op = p.adam_ptr[0] + e + half * 16;
          float* om = p.adam_ptr[1] + e + half * 16;
          float* ov = p.adam_ptr[2] + e + half * 16;
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            ptx::stg_256(op + 8 * k, pv[2 * k], pv[2 * k + 1]);
            ptx::stg_256(om + 8 * k, mv[2 * k], mv[2 * k + 1]);
            ptx::stg_256(ov + 8 * k, vv[2 * k], vv[2 * k + 1]);
          }
          // bf16 copy: one full 32-byte sector per half
          ptx::stg_256(reinterpret_cast<__nv_bfloat16*>(p.out) + e + half * 16, packed);
        }
      }
      if (!acc_ready) {                     // no chunk of this tile is ours: just release the buffer
        ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
        ptx::tc_fence_after();
        ptx::tc_fence_before();
        release_acc(acc);
      }
      if (++acc == n_acc) { acc = 0; acc_phase ^= 1u; }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int q = warp & 3;                 // TMEM lane quadrant this warp may read
    const int row_in_tile = q * 32 + lane;
    uint8_t* my_stage = smem_epi + q * 2 * kEpiWarpBufBytes;
    int acc = 0;
    uint32_t acc_phase = 0;
    int store_buf = 0;
    for (int tile = first_tile; tile < num_tiles; tile += tile_step) {
      const int m0 = (tile % p.num_m_tiles) * kTileM + row_base;
      const int n0 = ((tile / p.num_m_tiles) % p.num_n_tiles) * BN;
      const int m = m0 + row_in_tile;
      const bool row_ok = m < p.M;
      const int split_row = (tile / tiles_mn) * p.split_rows;   // output row offset of this K piece
      ptx::mbar_wait(&tmem_full_bar[acc], acc_phase);
      ptx::tc_fence_after();
      float loss_acc = 0.f;
      const int n_chunks = BN / 32;
      for (int c = 0; c < n_chunks; ++c) {
        const int n = n0 + c * 32;
        if (n >= p.N) break;  // warp-uniform
        uint32_t r[32];
        ptx::tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) +
                               static_cast<uint32_t>(acc * acc_stride + c * 32),
                           r);
        ptx::tmem_ld_wait();
        if (c == n_chunks - 1 || n + 32 >= p.N) {
          // last read of this accumulator buffer: hand it back to the MMA warp early
          ptx::tc_fence_before();
          release_acc(acc);
        }
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);

        if (EPI == kEpiF32) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float z = v[j] * p.alpha;
            if (p.bias != nullptr) z += __ldg(p.bias + n + j);
            if (p.clamp01) z = fminf(fmaxf(z, 0.f), 1.f);
            v[j] = z;
          }
          if (p.out_bf16) {
            // bf16 output (the data-parallel gradient that crosses NVLink): a lane owns a row,
            // its 32 columns are 64 contiguous bytes = two full sectors
            if (row_ok) {
              __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldo + n;
#pragma unroll
              for (int ch = 0; ch < 4; ++ch) {
                const __nv_bfloat162 a = __floats2bfloat162_rn(v[8 * ch], v[8 * ch + 1]);
                const __nv_bfloat162 b = __floats2bfloat162_rn(v[8 * ch + 2], v[8 * ch + 3]);
                const __nv_bfloat162 c = __floats2bfloat162_rn(v[8 * ch + 4], v[8 * ch + 5]);
                const __nv_bfloat162 d = __floats2bfloat162_rn(v[8 * ch + 6], v[8 * ch + 7]);
                *reinterpret_cast<uint4*>(o + 8 * ch) =
                    make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                               *reinterpret_cast<const uint32_t*>(&c), *reinterpret_cast<const uint32_t*>(&d));
              }
            }
          } else if (p.use_tma_store) {
            uint8_t* buf = my_stage + store_buf * kEpiWarpBufBytes;
            if (lane == 0) ptx::tma_store_wait_read<1>();
            __syncwarp();
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
              float4 val = make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
              *reinterpret_cast<float4*>(buf + lane * 128 + ((ch ^ (lane & 7)) << 4)) = val;
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) {
              ptx::tma_store_2d(&p.tm_c, buf, n, split_row + m0 + q * 32);
              ptx::tma_store_commit();
            }
            store_buf ^= 1;
          } else if (row_ok) {
            float* o = reinterpret_cast<float*>(p.out) + static_cast<long long>(split_row + m) * p.ldo + n;
#pragma unroll
            for (int ch = 0; ch < 8; ++ch)
              *reinterpret_cast<float4*>(o + 4 * ch) =
                  make_float4(v[4 * ch], v[4 * ch + 1], v[4 * ch + 2], v[4 * ch + 3]);
          }
        } else if (EPI == kEpiU8) {
          if (row_ok) {
            float bias[32];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {   // n is a multiple of 32: 16-byte aligned groups
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + ch);
              bias[4 * ch] = b4.x; bias[4 * ch + 1] = b4.y; bias[4 * ch + 2] = b4.z; bias[4 * ch + 3] = b4.w;
            }
            uint32_t packed[8];
#pragma unroll
            for (int w = 0; w < 8; ++w) {
              uint32_t word = 0;
#pragma unroll
              for (int b = 0; b < 4; ++b) {
                const int j = 4 * w + b;
                float z = v[j] + bias[j];
                z = fminf(fmaxf(z, 0.f), 1.f);
                const uint32_t u = static_cast<uint32_t>(__fmul_rn(z, 255.0f));  // truncation
                word |= (u & 0xFFu) << (8 * b);
              }
              packed[w] = word;
            }
            uint8_t* o = reinterpret_cast<uint8_t*>(p.out) + static_cast<long long>(m) * p.ldo + n;
            *reinterpret_cast<uint4*>(o) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
            *reinterpret_cast<uint4*>(o + 16) =
                make_uint4(packed[4], packed[5], packed[6], packed[7]);
          }
        } else {  // kEpiLoss
          if (row_ok) {
            float t[32];
            if (p.target_is_f32) {
              const float4* tp = reinterpret_cast<const float4*>(
                  reinterpret_cast<const float*>(p.target) + static_cast<long long>(m) * p.N + n);
#pragma unroll
              for (int ch = 0; ch < 8; ++ch) {
                const float4 tv = __ldg(tp + ch);
                t[4 * ch] = tv.x; t[4 * ch + 1] = tv.y; t[4 * ch + 2] = tv.z; t[4 * ch + 3] = tv.w;
              }
            } else {
              const uint4* tp = reinterpret_cast<const uint4*>(
                  reinterpret_cast<const uint8_t*>(p.target) + static_cast<long long>(m) * p.N + n);
              const uint4 t0 = __ldg(tp), t1 = __ldg(tp + 1);
              const uint32_t words[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
              for (int w = 0; w < 8; ++w)
#pragma unroll
                for (int b = 0; b < 4; ++b)
                  // exactly np.float32(u8) / 255.0 (helpers.py:121): round-to-nearest division
                  // without the range-check branch (operands are 0 or normal)
                  t[4 * w + b] = div_rn_nobranch(static_cast<float>((words[w] >> (8 * b)) & 0xFFu), 255.0f);
            }
            float bias[32];
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {   // n is a multiple of 32: 16-byte aligned groups
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + n) + ch);
              bias[4 * ch] = b4.x; bias[4 * ch + 1] = b4.y; bias[4 * ch + 2] = b4.z; bias[4 * ch + 3] = b4.w;
            }
            uint32_t packed[16];
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float g2[2];
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const float z = v[j + e] + bias[j + e];
                const float y = fminf(fmaxf(z, 0.f), 1.f);
                const float d = y - t[j + e];
                loss_acc = fmaf(d, d, loss_acc);
                // d clamp / dz is 1 on the closed interval [0,1] (torch.clamp backward)
                g2[e] = (z >= 0.f && z <= 1.f) ? d : 0.f;
              }
              const __nv_bfloat162 h = __floats2bfloat162_rn(g2[0], g2[1]);
              packed[j >> 1] = *reinterpret_cast<const uint32_t*>(&h);
            }
            __nv_bfloat16* o =
                reinterpret_cast<__nv_bfloat16*>(p.out) + static_cast<long long>(m) * p.ldo + n;
#pragma unroll
            for (int ch = 0; ch < 4; ++ch)
              *reinterpret_cast<uint4*>(o + 8 * ch) = make_uint4(
                  packed[4 * ch], packed[4 * ch + 1], packed[4 * ch + 2], packed[4 * ch + 3]);
          }
        }
      }
      if (EPI == kEpiLoss) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1)
          loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
        if (lane == 0) p.loss_partials[(tile * kCtas + cta_rank) * 4 + q] = loss_acc;
      }
      if (++acc == n_acc) { acc = 0; acc_phase ^= 1u; }
    }
    if (EPI == kEpiF32 && p.use_tma_store && lane == 0) ptx::tma_store_wait_all<0>();
  }

  ptx::tc_fence_before();
  if constexpr (CTA2) ptx::cluster_sync();   // nobody signals a CTA that has exited
  else __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    if constexpr (CTA2) ptx::tmem_dealloc_pair(tmem_base, static_cast<uint32_t>(p.tmem_cols));
    else ptx::tmem_dealloc(tmem_base, static_cast<uint32_t>(p.tmem_cols));
  }
}

}  // namespace afr
