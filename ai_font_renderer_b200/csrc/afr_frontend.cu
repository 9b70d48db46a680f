// Front-end of AttentionFontRenderer.forward (reference model.py:167-193) and its backward:
//   e = dropout(Emb[x]) + Pos ; a = MHA(e,e,e) ; h = LayerNorm(e + a) ;
//   f = dropout(relu(fc1(h))) ; feats = [f.reshape(B, S*64) | zeros]   (bf16, K-major GEMM operand)
// MHA follows torch.nn.functional.multi_head_attention_forward (packed in-proj, q scaled by
// sqrt(1/head_dim), softmax over all S keys with no padding mask, dropout on probabilities).
//
// 2.5 MFLOP/sample forward, S = 100, E = 32, head_dim = 8: far below a tcgen05 tile, and it has to
// hold the fp32 tolerance (1e-5 forward / 5e-5 backward against the oracle). Two tools:
//   * warp-level TF32 MMAs with 3-way split products (ptx::mma_3xtf32: fp32-equivalent; one HMMA
//     issues per 2.2 clocks per SM, tools/microbench/mma_rates.cu) for every linear layer of the
//     forward and the backward and for the attention backward: a fragment load replaces eight
//     broadcast LDS and the accumulators of the weight gradients live in the MMA accumulator
//     layout across all samples of a CTA;
//   * packed FFMA2 (118 FMA/clk/SM in half the issue slots of FFMA, tools/microbench/
//     fp32_rates.cu) fed by warp-uniform LDS.128 where the shapes do not pay for fragments: the
//     forward attention (one thread per (query, head), head-uniform warps, online soft-max over
//     blocks of 8 keys in the log2 domain, MUFU.EX2).
// A whole sample sits in shared memory per CTA.
//   * The training forward leaves a record per sample (e, q, k, v, context, normalised residual,
//     soft-max statistics and every dropout decision as bit masks; FrontStateLayout). The backward
//     (three kernels, see below) stages it back in and hands its intermediates from kernel to
//     kernel through the same record: no random number is drawn twice.
//   * The ten small weight gradients accumulate in registers across all samples of a CTA and
//     leave as per-CTA partials that a last kernel sums in a fixed order: deterministic, no
//     float atomics. The embedding scatter-add is a per-CTA shared-memory table walked in
//     position order by one warp.
#include <cstdlib>
#include <initializer_list>
#include <type_traits>

#include "afr_internal.h"
#include "afr_philox.cuh"
#include "afr_ptx.cuh"

namespace afr {
namespace {

using ptx::ex2;
using ptx::fma2;
using ptx::mul2;

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;
constexpr float kInvSqrtDh = 0.35355339059327373f;              // math.sqrt(1.0 / 8)
constexpr int kEmbSmemMaxVocab = 256;                           // dEmb table in smem up to here

// Optional per-phase cycle counters (nvcc -DAFR_PHASE_TIMING, see tools/phase_timing.py): thread 0
// of every CTA accumulates clock64() deltas between the marks below; summed over CTAs at exit.
#ifdef AFR_PHASE_TIMING
__device__ unsigned long long g_phase_cycles[2][16];
#define AFR_TICK_DECL long long tacc_[16] = {0}; long long tlast_ = clock64();
#define AFR_TICK(k)                                                   \
  do {                                                                \
    if (threadIdx.x == 0) {                                           \
      const long long now_ = clock64();                               \
      tacc_[k] += now_ - tlast_;                                      \
      tlast_ = now_;                                                  \
    }                                                                 \
  } while (0)
#define AFR_TICK_FLUSH(kernel)                                        \
  do {                                                                \
    if (threadIdx.x == 0)                                             \
      for (int k_ = 0; k_ < 16; ++k_)                                 \
        atomicAdd(&g_phase_cycles[kernel][k_], static_cast<unsigned long long>(tacc_[k_])); \
  } while (0)
#else
#define AFR_TICK_DECL
#define AFR_TICK(k)
#define AFR_TICK_FLUSH(kernel)
#endif

struct FrontArgs {
  Tensors w;
  const long long* tokens;
  long long token_stride;
  int B, S, L, vocab;
  Dropout drop;
  uint32_t thr_e, thr_a, thr_f;   // keep iff u16 >= thr
  float inv_e, inv_a, inv_f;      // 1 / (1 - p), or 1 with dropout off
  __nv_bfloat16* feats;           // forward
  float* feats_f32;               // forward, optional fp32 copy (test hook)
  float* state;                   // forward: written when non-null; backward: read
  FrontStateLayout sl;
  const float* dfeat;             // backward: [B, L*F]
  float* partials;                // backward: [grid, lay.total]
  SmallLayout lay;
  int* err_flag;
  FontCond font;                  // optional font conditioning (ids == nullptr: none)
};

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 f2(float x, float y) { return make_float2(x, y); }

// Slot i in [0, 4S) -> (head, row). Full 32-row chunks are head-uniform per warp (K/V loads
// become broadcasts); the S % 32 tail rows of all four heads are packed behind them.
__device__ __forceinline__ void slot_to_pair(int i, int S, int& h, int& r) {
  const int full = (S >> 5) << 7;
  if (i < full) {
    h = (i >> 5) & 3;
    r = ((i >> 7) << 5) + (i & 31);
  } else {
    const int rem = S & 31, x = i - full;
    h = x / rem;
    r = (S & ~31) + x % rem;
  }
}

// dot of the 8 head channels: a (4 packed pairs) . row (two float4)
__device__ __forceinline__ float dot8(const float2 (&a)[4], const float4& r0, const float4& r1) {
  float2 s = mul2(a[0], f2(r0.x, r0.y));
  s = fma2(a[1], f2(r0.z, r0.w), s);
  s = fma2(a[2], f2(r1.x, r1.y), s);
  s = fma2(a[3], f2(r1.z, r1.w), s);
  return s.x + s.y;
}
// acc += w * row
__device__ __forceinline__ void axpy8(float2 (&acc)[4], float w, const float4& r0, const float4& r1) {
  const float2 ww = f2(w, w);
  acc[0] = fma2(ww, f2(r0.x, r0.y), acc[0]);
  acc[1] = fma2(ww, f2(r0.z, r0.w), acc[1]);
  acc[2] = fma2(ww, f2(r1.x, r1.y), acc[2]);
  acc[3] = fma2(ww, f2(r1.z, r1.w), acc[3]);
}

// ============================================================================ forward kernel
// 14 warps. The linear layers (in-projection, out-projection, fc1) run on warp-level TF32 MMAs
// with 3-way split products (ptx::mma_3xtf32, fp32-equivalent: the 1e-5 tolerance holds): a warp
// owns a 16-row tile and half (or all) of the output columns, the row tile's A fragment of one
// k-step is split once and reused for all its column tiles. The FFMA2 forms (a lane per output
// channel, the row broadcast by LDS.128) cost 43 k warp instructions per sample, these 11 k.
// Activations and weights sit in shared memory as rows of 32 floats with their 4-float groups
// XOR-swizzled by the row (group ^ (row & 7)): every fragment load hits 32 distinct banks with no
// padding, and the attention phase's LDS.128 rows stay 16-byte aligned.
constexpr int kFwdWarps = 14;
constexpr int kFwdThreads = kFwdWarps * 32;

__device__ __forceinline__ int swz(int row, int col) {   // word offset of (row, col) in a swizzled [rows][32] array
  return row * kE + ((((col >> 2) ^ (row & 7)) << 2) | (col & 3));
}
__device__ __forceinline__ void load_matrix_swz(const float* __restrict__ g, float* sm, int rows) {
  for (int i = threadIdx.x; i < rows * kE; i += kFwdThreads) sm[swz(i / kE, i % kE)] = g[i];
}
// A fragment (rows r0, r1; columns 8 ks + {t, t + 4}) of a swizzled activation array, split
__device__ __forceinline__ void load_a_frag(const float* x, int r0, int r1, int ks, int t, uint32_t (&ah)[4],
                                            uint32_t (&al)[4]) {
  ptx::split_tf32(x[swz(r0, 8 * ks + t)], ah[0], al[0]);
  ptx::split_tf32(x[swz(r1, 8 * ks + t)], ah[1], al[1]);
  ptx::split_tf32(x[swz(r0, 8 * ks + t + 4)], ah[2], al[2]);
  ptx::split_tf32(x[swz(r1, 8 * ks + t + 4)], ah[3], al[3]);
}
// acc += A . W^T for output channels n0 .. n0 + 7 (W row-major [out][in], swizzled), one k-step
__device__ __forceinline__ void mma_w(float (&acc)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4],
                                      const float* w, int n0, int ks, int g, int t) {
  uint32_t bh0, bl0, bh1, bl1;
  ptx::split_tf32(w[swz(n0 + g, 8 * ks + t)], bh0, bl0);
  ptx::split_tf32(w[swz(n0 + g, 8 * ks + t + 4)], bh1, bl1);
  ptx::mma_3xtf32(acc, ah, al, bh0, bh1, bl0, bl1);
}

struct FwdSmem {
  int win, wo, w1, bin, bo, lnw, lnb, b1, e, q, k, v, total;
};
__host__ __device__ inline FwdSmem make_fwd_smem(int L) {
  FwdSmem s{};
  int o = 0;
  s.win = o; o += 3 * kE * kE;
  s.wo = o;  o += kE * kE;
  s.w1 = o;  o += kF * kE;
  s.bin = o; o += 3 * kE;
  s.bo = o;  o += kE;
  s.lnw = o; o += kE;
  s.lnb = o; o += kE;
  s.b1 = o;  o += kF;
  o = (o + 3) & ~3;
  s.e = o; o += L * kE;
  s.q = o; o += L * kE;
  s.k = o; o += L * kE;
  s.v = o; o += (L + 8) * kE;   // the last key block may read (never use) up to 7 rows past S
  s.total = o;
  return s;
}

__device__ __forceinline__ void frontend_forward_body(const FrontArgs& a) {
  extern __shared__ __align__(16) float sm[];
  const FwdSmem o = make_fwd_smem(a.L);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int S = a.S, KF = a.L * kF;
  const int mode = a.drop.mode;

  load_matrix_swz(a.w.win, sm + o.win, 3 * kE);
  load_matrix_swz(a.w.wo, sm + o.wo, kE);
  load_matrix_swz(a.w.w1, sm + o.w1, kF);
  for (int i = tid; i < 3 * kE; i += kFwdThreads) sm[o.bin + i] = a.w.bin[i];
  for (int i = tid; i < kF; i += kFwdThreads) sm[o.b1 + i] = a.w.b1[i];
  if (tid < kE) {
    sm[o.bo + tid] = a.w.bo[tid];
    sm[o.lnw + tid] = a.w.lnw[tid];
    sm[o.lnb + tid] = a.w.lnb[tid];
  }
  __syncthreads();

  float* se = sm + o.e;
  float* sq = sm + o.q;
  float* sk = sm + o.k;
  float* sv = sm + o.v;
  const int ntile_rows = (S + 15) >> 4;     // 16-row tiles of this batch

  AFR_TICK_DECL
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    AFR_TICK(0);
    const Rng rng{static_cast<uint32_t>(a.drop.seed), static_cast<uint32_t>(a.drop.seed >> 32),
                  static_cast<uint32_t>(a.drop.sample_offset + b), static_cast<uint32_t>(a.drop.step)};
    const long long* tok = a.tokens + static_cast<long long>(b) * a.token_stride;
    float* st = a.state != nullptr ? a.state + static_cast<long long>(b) * a.sl.stride : nullptr;

    // ---- (1) e = dropout(Emb[tok]) + Pos        model.py:167-172 (dropout BEFORE positions)
    // one thread per 8 channels = one Philox block
    for (int base = 0; base < S * 4; base += kFwdThreads) {
      const int i8 = base + tid;
      const bool valid = i8 < S * 4;
      uint32_t keep8 = 0xFFu;
      if (valid) {
        const int s = i8 >> 2, c0 = (i8 & 3) * 8;
        long long tk = tok[s];
        if (tk < 0 || tk >= a.vocab) { atomicOr(a.err_flag, 1); tk = 0; }
        const float4 e0 = __ldg(reinterpret_cast<const float4*>(a.w.emb + tk * kE + c0));
        const float4 e1 = __ldg(reinterpret_cast<const float4*>(a.w.emb + tk * kE + c0 + 4));
        const float4 p0 = __ldg(reinterpret_cast<const float4*>(a.w.pos + s * kE + c0));
        const float4 p1 = __ldg(reinterpret_cast<const float4*>(a.w.pos + s * kE + c0 + 4));
        if (mode == 1) {
          const uint4 r = rng.block(0u, static_cast<uint32_t>(i8));
          keep8 = (u16_of<0>(r) >= a.thr_e ? 1u : 0u) | (u16_of<1>(r) >= a.thr_e ? 2u : 0u) |
                  (u16_of<2>(r) >= a.thr_e ? 4u : 0u) | (u16_of<3>(r) >= a.thr_e ? 8u : 0u) |
                  (u16_of<4>(r) >= a.thr_e ? 16u : 0u) | (u16_of<5>(r) >= a.thr_e ? 32u : 0u) |
                  (u16_of<6>(r) >= a.thr_e ? 64u : 0u) | (u16_of<7>(r) >= a.thr_e ? 128u : 0u);
        } else if (mode == 2) {
          const uint2 mk = *reinterpret_cast<const uint2*>(
              a.drop.mask_embed + (static_cast<long long>(b) * S + s) * kE + c0);
          keep8 = 0;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            keep8 |= ((mk.x >> (8 * u)) & 0xFFu) ? (1u << u) : 0u;
            keep8 |= ((mk.y >> (8 * u)) & 0xFFu) ? (16u << u) : 0u;
          }
        }
        float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
        const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
        if (a.font.ids != nullptr) {   // config 3: + font_embedding[font of this sample], before the dropout
          const float* fr = a.font.table + static_cast<long long>(a.font.ids[b]) * kE + c0;
          const float4 f0 = __ldg(reinterpret_cast<const float4*>(fr)), f1 = __ldg(reinterpret_cast<const float4*>(fr + 4));
          ev[0] += f0.x; ev[1] += f0.y; ev[2] += f0.z; ev[3] += f0.w;
          ev[4] += f1.x; ev[5] += f1.y; ev[6] += f1.z; ev[7] += f1.w;
        }
        float ov[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          ov[u] = ((keep8 >> u) & 1u) ? fmaf(ev[u], a.inv_e, pv[u]) : pv[u];
        *reinterpret_cast<float4*>(se + swz(s, c0)) = make_float4(ov[0], ov[1], ov[2], ov[3]);
        *reinterpret_cast<float4*>(se + swz(s, c0 + 4)) = make_float4(ov[4], ov[5], ov[6], ov[7]);
        if (st != nullptr) {     // the record's copy (backward: dWin), rows of kLdT floats
          *reinterpret_cast<float4*>(st + a.sl.e + s * kLdT + c0) = make_float4(ov[0], ov[1], ov[2], ov[3]);
          *reinterpret_cast<float4*>(st + a.sl.e + s * kLdT + c0 + 4) = make_float4(ov[4], ov[5], ov[6], ov[7]);
        }
      }
      if (st != nullptr) {   // 32 keep bits per position, assembled from the 4 threads of the row
        uint32_t wbits = keep8 << (8 * (tid & 3));
        wbits |= __shfl_xor_sync(0xffffffffu, wbits, 1);
        wbits |= __shfl_xor_sync(0xffffffffu, wbits, 2);
        if (valid && (tid & 3) == 0) reinterpret_cast<uint32_t*>(st + a.sl.ebits)[i8 >> 2] = wbits;
      }
    }
    AFR_TICK(1);
    __syncthreads();
    AFR_TICK(2);

    // ---- (2) packed in-projection q|k|v = e Win^T + bin  (torch functional.py:5836) -----------
    // warp = (row tile, half of the 96 output channels)
    if ((warp >> 1) < ntile_rows) {
      const int mt = warp >> 1, nb = 48 * (warp & 1);
      const int r0 = 16 * mt + g, r1 = r0 + 8, q0 = min(r0, S - 1), q1 = min(r1, S - 1);
      float acc[6][4];
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[4], al[4];
        load_a_frag(se, q0, q1, ks, t, ah, al);
#pragma unroll
        for (int nt = 0; nt < 6; ++nt) mma_w(acc[nt], ah, al, sm + o.win, nb + 8 * nt, ks, g, t);
      }
#pragma unroll
      for (int nt = 0; nt < 6; ++nt) {
        const int n = nb + 8 * nt + 2 * t, j = n >> 5, ch = n & 31;     // j: 0 = q, 1 = k, 2 = v
        const float scale = j == 0 ? kInvSqrtDh * kLog2e : 1.f;          // q also carries log2(e): ex2 soft-max
        const float b0 = sm[o.bin + n], b1 = sm[o.bin + n + 1];
        float* dst = j == 0 ? sq : (j == 1 ? sk : sv);
        float* gdst = st != nullptr ? st + (j == 0 ? a.sl.q : (j == 1 ? a.sl.k : a.sl.v)) : nullptr;
        // q (later: the context, A operand of the out-projection) is swizzled; k and v are only read
        // row by row by the attention phase and stay plain
        if (r0 < S) {
          const float2 v = make_float2((acc[nt][0] + b0) * scale, (acc[nt][1] + b1) * scale);
          *reinterpret_cast<float2*>(dst + (j == 0 ? swz(r0, ch) : r0 * kE + ch)) = v;
          if (gdst != nullptr) *reinterpret_cast<float2*>(gdst + r0 * kE + ch) = v;
        }
        if (r1 < S) {
          const float2 v = make_float2((acc[nt][2] + b0) * scale, (acc[nt][3] + b1) * scale);
          *reinterpret_cast<float2*>(dst + (j == 0 ? swz(r1, ch) : r1 * kE + ch)) = v;
          if (gdst != nullptr) *reinterpret_cast<float2*>(gdst + r1 * kE + ch) = v;
        }
      }
    }
    AFR_TICK(3);
    __syncthreads();
    AFR_TICK(4);

    // ---- (3) per (query s, head h): ctx = dropout(softmax(q k^T)) v   (functional.py:6642-6647)
    for (int i = tid; i < S * kHeads; i += kFwdThreads) {
      int h, s;
      slot_to_pair(i, S, h, s);
      float2 q2[4];
      {
        const float4 q0 = lds4(sq + swz(s, h * kDh)), q1 = lds4(sq + swz(s, h * kDh + 4));
        q2[0] = f2(q0.x, q0.y); q2[1] = f2(q0.z, q0.w); q2[2] = f2(q1.x, q1.y); q2[3] = f2(q1.z, q1.w);
      }
      float m = -INFINITY, l = 0.f;
      float2 acc[4] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
      const uint32_t row = static_cast<uint32_t>(h * S + s);
      uint32_t* gbits = st != nullptr ? reinterpret_cast<uint32_t*>(st + a.sl.abits) + row * 4 : nullptr;
      const uint8_t* mrow = mode == 2
          ? a.drop.mask_attn + ((static_cast<long long>(b) * kHeads + h) * S + s) * S : nullptr;
      uint32_t bits = 0;
      const float* kh = sk + h * kDh;
      const float* vh = sv + h * kDh;
      const int nblk = (S + 7) >> 3;
      // full blocks of 8 keys run without the `t < S` selects; only the last block may be partial
      auto key_block = [&](const int blk, auto tail_tag) {
        constexpr bool kTail = decltype(tail_tag)::value;
        const int t0 = blk * 8;
        float sc[8];
        float bm = m;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const float4 k0 = lds4(kh + (t0 + u) * kE), k1 = lds4(kh + (t0 + u) * kE + 4);
          const float d = dot8(q2, k0, k1);
          sc[u] = (!kTail || t0 + u < S) ? d : -INFINITY;
          bm = fmaxf(bm, sc[u]);
        }
        const float corr = ex2(m - bm);      // first block: ex2(-inf) = 0 (l and acc are 0 anyway)
        m = bm;
        l *= corr;
        {
          const float2 c2 = f2(corr, corr);
          acc[0] = mul2(acc[0], c2); acc[1] = mul2(acc[1], c2);
          acc[2] = mul2(acc[2], c2); acc[3] = mul2(acc[3], c2);
        }
        uint4 rnd = make_uint4(0, 0, 0, 0);
        if (mode == 1) rnd = rng.block(1u | (row << 2), static_cast<uint32_t>(blk));
        uint32_t keep8 = 0xFFu;
        if (mode == 1) {
          keep8 = (u16_of<0>(rnd) >= a.thr_a ? 1u : 0u) | (u16_of<1>(rnd) >= a.thr_a ? 2u : 0u) |
                  (u16_of<2>(rnd) >= a.thr_a ? 4u : 0u) | (u16_of<3>(rnd) >= a.thr_a ? 8u : 0u) |
                  (u16_of<4>(rnd) >= a.thr_a ? 16u : 0u) | (u16_of<5>(rnd) >= a.thr_a ? 32u : 0u) |
                  (u16_of<6>(rnd) >= a.thr_a ? 64u : 0u) | (u16_of<7>(rnd) >= a.thr_a ? 128u : 0u);
        } else if (mode == 2) {
          keep8 = 0;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if ((!kTail || t0 + u < S) && mrow[t0 + u] != 0) keep8 |= 1u << u;
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (!kTail || t0 + u < S) {     // warp-uniform
            const float p = ex2(sc[u] - m);
            l += p;             // the soft-max denominator counts dropped keys too
            const float pk = ((keep8 >> u) & 1u) ? p : 0.f;
            const float4 v0 = lds4(vh + (t0 + u) * kE), v1 = lds4(vh + (t0 + u) * kE + 4);
            axpy8(acc, pk, v0, v1);
          }
        }
        bits |= keep8 << (8 * (blk & 3));
        if ((blk & 3) == 3 || blk == nblk - 1) {
          if (gbits != nullptr) gbits[blk >> 2] = bits;
          bits = 0;
        }
      };
      const int nfull = S >> 3;
      for (int blk = 0; blk < nfull; ++blk) key_block(blk, std::false_type{});
      if (nfull < nblk) key_block(nfull, std::true_type{});
      const float linv = 1.f / l;
      const float scale = a.inv_a * linv;
      // q of this (s, h) is dead: the context vector takes its place
      *reinterpret_cast<float4*>(sq + swz(s, h * kDh)) =
          make_float4(acc[0].x * scale, acc[0].y * scale, acc[1].x * scale, acc[1].y * scale);
      *reinterpret_cast<float4*>(sq + swz(s, h * kDh + 4)) =
          make_float4(acc[2].x * scale, acc[2].y * scale, acc[3].x * scale, acc[3].y * scale);
      if (st != nullptr)
        *reinterpret_cast<float4*>(st + a.sl.stat + (s * kHeads + h) * 4) = make_float4(m, linv, 0.f, 0.f);
    }
    AFR_TICK(5);
    __syncthreads();
    AFR_TICK(6);

    // ---- (4a) out-projection + residual + LayerNorm (model.py:180); h overwrites e ----------
    // warp = row tile with all 32 output channels: the row statistics stay inside a quad
    if (warp < ntile_rows) {
      const int r0 = 16 * warp + g, r1 = r0 + 8, q0 = min(r0, S - 1), q1 = min(r1, S - 1);
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[4], al[4];
        load_a_frag(sq, q0, q1, ks, t, ah, al);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_w(acc[nt], ah, al, sm + o.wo, 8 * nt, ks, g, t);
      }
      // residual uses the dropped + positioned e; rows r0 (elements 0, 1) and r1 (2, 3)
      float ev[4][4];
      float sa = 0.f, sb = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = 8 * nt + 2 * t;
        const float2 ea = *reinterpret_cast<const float2*>(se + swz(q0, c));
        const float2 eb = *reinterpret_cast<const float2*>(se + swz(q1, c));
        const float bo0 = sm[o.bo + c], bo1 = sm[o.bo + c + 1];
        ev[nt][0] = ea.x; ev[nt][1] = ea.y; ev[nt][2] = eb.x; ev[nt][3] = eb.y;
        acc[nt][0] += ea.x + bo0; acc[nt][1] += ea.y + bo1;
        acc[nt][2] += eb.x + bo0; acc[nt][3] += eb.y + bo1;
        sa += acc[nt][0] + acc[nt][1];
        sb += acc[nt][2] + acc[nt][3];
      }
      sa += __shfl_xor_sync(0xffffffffu, sa, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 1);
      sa += __shfl_xor_sync(0xffffffffu, sa, 2); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
      const float mean_a = sa * (1.f / kE), mean_b = sb * (1.f / kE);
      float va = 0.f, vb = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        acc[nt][0] -= mean_a; acc[nt][1] -= mean_a; acc[nt][2] -= mean_b; acc[nt][3] -= mean_b;
        va = fmaf(acc[nt][0], acc[nt][0], va); va = fmaf(acc[nt][1], acc[nt][1], va);
        vb = fmaf(acc[nt][2], acc[nt][2], vb); vb = fmaf(acc[nt][3], acc[nt][3], vb);
      }
      va += __shfl_xor_sync(0xffffffffu, va, 1); vb += __shfl_xor_sync(0xffffffffu, vb, 1);
      va += __shfl_xor_sync(0xffffffffu, va, 2); vb += __shfl_xor_sync(0xffffffffu, vb, 2);
      const float rstd_a = 1.f / sqrtf(va * (1.f / kE) + 1e-5f);    // biased variance, eps = 1e-5
      const float rstd_b = 1.f / sqrtf(vb * (1.f / kE) + 1e-5f);
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = 8 * nt + 2 * t;
        const float g0 = sm[o.lnw + c], g1 = sm[o.lnw + c + 1], be0 = sm[o.lnb + c], be1 = sm[o.lnb + c + 1];
        const float xa0 = acc[nt][0] * rstd_a, xa1 = acc[nt][1] * rstd_a;
        const float xb0 = acc[nt][2] * rstd_b, xb1 = acc[nt][3] * rstd_b;
        if (r0 < S) {
          if (st != nullptr) {
            *reinterpret_cast<float2*>(st + a.sl.ctx + r0 * kE + c) = *reinterpret_cast<const float2*>(sq + swz(r0, c));
            *reinterpret_cast<float2*>(st + a.sl.xhat + r0 * kE + c) = make_float2(xa0, xa1);
          }
          *reinterpret_cast<float2*>(se + swz(r0, c)) = make_float2(fmaf(xa0, g0, be0), fmaf(xa1, g1, be1));
        }
        if (r1 < S) {
          if (st != nullptr) {
            *reinterpret_cast<float2*>(st + a.sl.ctx + r1 * kE + c) = *reinterpret_cast<const float2*>(sq + swz(r1, c));
            *reinterpret_cast<float2*>(st + a.sl.xhat + r1 * kE + c) = make_float2(xb0, xb1);
          }
          *reinterpret_cast<float2*>(se + swz(r1, c)) = make_float2(fmaf(xb0, g0, be0), fmaf(xb1, g1, be1));
        }
      }
      if (st != nullptr && t == 0) {
        if (r0 < S) st[a.sl.rstd + r0] = rstd_a;
        if (r1 < S) st[a.sl.rstd + r1] = rstd_b;
      }
      (void)ev;
    }
    __syncthreads();   // (4b) reads rows other warps normalised
    AFR_TICK(7);

    // ---- (4b) f = dropout(relu(fc1(h)))  (model.py:183-184). warp = (row tile, half of the 64 features)
    if ((warp >> 1) < ntile_rows) {
      const int mt = warp >> 1, nh = warp & 1;
      const int r0 = 16 * mt + g, r1 = r0 + 8, q0 = min(r0, S - 1), q1 = min(r1, S - 1);
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) {
        uint32_t ah[4], al[4];
        load_a_frag(se, q0, q1, ks, t, ah, al);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) mma_w(acc[nt], ah, al, sm + o.w1, 32 * nh + 8 * nt, ks, g, t);
      }
      // keep decisions: Philox block (row, 8-feature group) = one column tile of one row. Lane t of a
      // quad draws the blocks of row (t & 1 ? r1 : r0) for column tiles (t >> 1) and (t >> 1) + 2; the
      // 8 keep bits of a block reach the other lanes by shuffle.
      uint32_t keepm[2] = {0xFFu, 0xFFu};
      if (mode == 1) {
        const int my_row = (t & 1) ? q1 : q0;
#pragma unroll
        for (int w = 0; w < 2; ++w) {
          const int jt = 4 * nh + (t >> 1) + 2 * w;     // 8-feature group of the row
          const uint4 r = rng.block(2u, static_cast<uint32_t>(my_row * (kF / 8) + jt));
          keepm[w] = (u16_of<0>(r) >= a.thr_f ? 1u : 0u) | (u16_of<1>(r) >= a.thr_f ? 2u : 0u) |
                     (u16_of<2>(r) >= a.thr_f ? 4u : 0u) | (u16_of<3>(r) >= a.thr_f ? 8u : 0u) |
                     (u16_of<4>(r) >= a.thr_f ? 16u : 0u) | (u16_of<5>(r) >= a.thr_f ? 32u : 0u) |
                     (u16_of<6>(r) >= a.thr_f ? 64u : 0u) | (u16_of<7>(r) >= a.thr_f ? 128u : 0u);
        }
      }
      __nv_bfloat16* out = a.feats + static_cast<long long>(b) * KF;
      const int quad = lane & ~3;
      uint32_t fb0a = 0, fb1a = 0, fb0b = 0, fb1b = 0;   // (ReLU' & keep) bits of rows r0 / r1: even / odd features
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int n = 32 * nh + 8 * nt + 2 * t;
        const float b0 = sm[o.b1 + n], b1 = sm[o.b1 + n + 1];
        float fa0 = fmaxf(acc[nt][0] + b0, 0.f), fa1 = fmaxf(acc[nt][1] + b1, 0.f);   // ReLU
        float fc0 = fmaxf(acc[nt][2] + b0, 0.f), fc1 = fmaxf(acc[nt][3] + b1, 0.f);
        bool ka0 = true, ka1 = true, kc0 = true, kc1 = true;
        if (mode == 1) {
          // column tile nt of row r0 was drawn by quad lane 2 (nt & 1), of row r1 by lane 2 (nt & 1) + 1, as block nt >> 1
          const uint32_t ma = __shfl_sync(0xffffffffu, keepm[nt >> 1], quad + 2 * (nt & 1)) >> (2 * t);
          const uint32_t mc = __shfl_sync(0xffffffffu, keepm[nt >> 1], quad + 2 * (nt & 1) + 1) >> (2 * t);
          ka0 = ma & 1u; ka1 = ma & 2u; kc0 = mc & 1u; kc1 = mc & 2u;
        } else if (mode == 2) {
          const uint8_t* mka = a.drop.mask_fc1 + (static_cast<long long>(b) * S + q0) * kF + n;
          const uint8_t* mkc = a.drop.mask_fc1 + (static_cast<long long>(b) * S + q1) * kF + n;
          ka0 = mka[0] != 0; ka1 = mka[1] != 0; kc0 = mkc[0] != 0; kc1 = mkc[1] != 0;
        }
        const int bit = 4 * nt + t;      // feature 2 (16 nh + bit) (+ 1): bit `bit` of this warp's 16-bit half
        fb0a |= (ka0 && fa0 > 0.f ? 1u : 0u) << bit; fb1a |= (ka1 && fa1 > 0.f ? 1u : 0u) << bit;
        fb0b |= (kc0 && fc0 > 0.f ? 1u : 0u) << bit; fb1b |= (kc1 && fc1 > 0.f ? 1u : 0u) << bit;
        fa0 = ka0 ? fa0 * a.inv_f : 0.f; fa1 = ka1 ? fa1 * a.inv_f : 0.f;
        fc0 = kc0 ? fc0 * a.inv_f : 0.f; fc1 = kc1 ? fc1 * a.inv_f : 0.f;
        if (r0 < S) {
          *reinterpret_cast<__nv_bfloat162*>(out + r0 * kF + n) = __floats2bfloat162_rn(fa0, fa1);
          if (a.feats_f32 != nullptr)
            *reinterpret_cast<float2*>(a.feats_f32 + static_cast<long long>(b) * KF + r0 * kF + n) = make_float2(fa0, fa1);
        }
        if (r1 < S) {
          *reinterpret_cast<__nv_bfloat162*>(out + r1 * kF + n) = __floats2bfloat162_rn(fc0, fc1);
          if (a.feats_f32 != nullptr)
            *reinterpret_cast<float2*>(a.feats_f32 + static_cast<long long>(b) * KF + r1 * kF + n) = make_float2(fc0, fc1);
        }
      }
      if (st != nullptr) {
        // fbits[s] = (word of the even features, word of the odd features), bit j = feature 2j / 2j + 1:
        // this warp owns bits 16 nh .. 16 nh + 15 of both words
        fb0a |= __shfl_xor_sync(0xffffffffu, fb0a, 1); fb1a |= __shfl_xor_sync(0xffffffffu, fb1a, 1);
        fb0b |= __shfl_xor_sync(0xffffffffu, fb0b, 1); fb1b |= __shfl_xor_sync(0xffffffffu, fb1b, 1);
        fb0a |= __shfl_xor_sync(0xffffffffu, fb0a, 2); fb1a |= __shfl_xor_sync(0xffffffffu, fb1a, 2);
        fb0b |= __shfl_xor_sync(0xffffffffu, fb0b, 2); fb1b |= __shfl_xor_sync(0xffffffffu, fb1b, 2);
        if (t == 0) {
          uint16_t* fw = reinterpret_cast<uint16_t*>(st + a.sl.fbits);
          if (r0 < S) { fw[4 * r0 + nh] = static_cast<uint16_t>(fb0a); fw[4 * r0 + 2 + nh] = static_cast<uint16_t>(fb1a); }
          if (r1 < S) { fw[4 * r1 + nh] = static_cast<uint16_t>(fb0b); fw[4 * r1 + 2 + nh] = static_cast<uint16_t>(fb1b); }
        }
      }
    }
    {
      __nv_bfloat16* out = a.feats + static_cast<long long>(b) * KF;
      // zero features for positions >= S (model.py:190-193)
      for (int i = S * kF / 2 + tid; i < KF / 2; i += kFwdThreads) {
        reinterpret_cast<__nv_bfloat162*>(out)[i] = __floats2bfloat162_rn(0.f, 0.f);
        if (a.feats_f32 != nullptr)
          reinterpret_cast<float2*>(a.feats_f32 + static_cast<long long>(b) * KF)[i] = make_float2(0.f, 0.f);
      }
    }
    AFR_TICK(8);
    __syncthreads();
    AFR_TICK(9);
  }
  AFR_TICK_FLUSH(0);
}

__global__ void __launch_bounds__(kFwdThreads, 2) frontend_forward_kernel(const FrontArgs a) {
  frontend_forward_body(a);
}
// Capped at 64 registers: two CTAs of 14 warps put 7 warps on each of the SM's four scheduler
// partitions, 7 x 72 x 32 registers leave no room there for a warp of another kernel; at 64 the
// background AdamW sweep's CTA (4 warps x 40 registers) fits beside both.
__global__ void __maxnreg__(64) frontend_forward_kernel_shared(const FrontArgs a) {
  frontend_forward_body(a);
}

// =========================================================================== backward kernels
// Three kernels per batch, each a loop over samples with the sample's operands in shared memory:
//   K1 frontend_backward_head_kernel : fc1 / ReLU / dropout / LayerNorm backward, out-projection
//      backward (B1, B1b) and dW1, db1, dWo, dbo (B2) on TF32 MMAs; hands d(residual), d(ctx) and D
//      to the record
//   K2 frontend_backward_attn_kernel : soft-max attention backward on warp-level TF32 MMAs
//   K3 frontend_backward_tail_kernel : in-projection backward, dPos, d(embedding) (B4) and dWin,
//      dbin (B5)
// A single kernel holding all of it (round 1) needed 190 KB of shared memory per sample: one
// 13-warp CTA per SM, 40 % of the issue slots and 61 % of the shared-memory pipe busy (ncu), half
// of its time in the attention phase. Split, the attention part takes two CTAs of 14 warps per SM.

// ---------------------------------------------------------------- K1: head of the backward
// dh = df W1 with the LayerNorm backward on the accumulator fragments (B1), d(ctx) = dr Wo with
// D = d(ctx) . ctx per head (B1b), dW1 += df^T h, dWo += dr^T ctx, db1, dbo (B2): all as 3xTF32
// warp MMAs like the tail kernel. W1 and Wo are split into TF32 hi / lo words once per CTA. Rows
// of 40 floats (xhat / h, dr, ctx) and 72 floats (df): 8t + g is a distinct bank for every lane of
// the B / A^T fragment loads, row-major A loads are 2-way. Rows beyond S are zero and pad the
// reductions over positions. 14 warps: 7 row tiles for B1 / B1b (the row statistics of the
// LayerNorm stay inside a quad), 8 + 4 tiles of the weight gradients for B2.
constexpr int kHeadWarps = 14;
constexpr int kHeadThreads = kHeadWarps * 32;
constexpr int kLdF = kF + 8;      // row stride of df

struct HeadSmem {
  int w1_hi, w1_lo, wo_hi, wo_lo, lnw, lnb, xh, df, dr, ctx, fbits, rstd, red, rows, total;
};
__host__ __device__ inline HeadSmem make_head_smem(int L) {
  const int L4 = (L + 3) & ~3;
  HeadSmem s{};
  s.rows = (L + 15) & ~15;
  int o = 0;
  s.w1_hi = o; o += kF * kLdT;
  s.w1_lo = o; o += kF * kLdT;
  s.wo_hi = o; o += kE * kLdT;
  s.wo_lo = o; o += kE * kLdT;
  s.lnw = o; o += kE;
  s.lnb = o; o += kE;
  s.xh = o;  o += s.rows * kLdT;     // xhat -> h (B1)
  s.dr = o;  o += s.rows * kLdT;     // d(residual)
  s.ctx = o; o += s.rows * kLdT;
  s.df = o;  o += s.rows * kLdF;     // dfeat -> df
  s.fbits = o; o += L4 * 2;
  s.rstd = o; o += L4;
  s.red = o;  o += 2 * 7 * kE;       // d(LayerNorm weight / bias) partials of the 7 row-tile warps
  s.total = (o + 3) & ~3;
  return s;
}

__device__ __forceinline__ void frontend_backward_head_body(const FrontArgs& a) {
  extern __shared__ __align__(16) float sm[];
  const HeadSmem o = make_head_smem(a.L);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int S = a.S, S4 = (S + 3) & ~3, KF = a.L * kF;
  float* part = a.partials + static_cast<long long>(blockIdx.x) * a.lay.total;

  uint32_t* w1_hi = reinterpret_cast<uint32_t*>(sm + o.w1_hi);
  uint32_t* w1_lo = reinterpret_cast<uint32_t*>(sm + o.w1_lo);
  uint32_t* wo_hi = reinterpret_cast<uint32_t*>(sm + o.wo_hi);
  uint32_t* wo_lo = reinterpret_cast<uint32_t*>(sm + o.wo_lo);
  for (int i = tid; i < kF * kE; i += kHeadThreads) {
    uint32_t hi, lo;
    ptx::split_tf32(a.w.w1[i], hi, lo);
    w1_hi[(i / kE) * kLdT + (i % kE)] = hi;
    w1_lo[(i / kE) * kLdT + (i % kE)] = lo;
  }
  for (int i = tid; i < kE * kE; i += kHeadThreads) {
    uint32_t hi, lo;
    ptx::split_tf32(a.w.wo[i], hi, lo);
    wo_hi[(i / kE) * kLdT + (i % kE)] = hi;
    wo_lo[(i / kE) * kLdT + (i % kE)] = lo;
  }
  if (tid < kE) {
    sm[o.lnw + tid] = a.w.lnw[tid];
    sm[o.lnb + tid] = a.w.lnb[tid];
  }
  // rows >= S of xhat / h, dr, ctx, df stay zero (xh, dr, ctx, df are contiguous)
  for (int i = tid; i < 3 * o.rows * kLdT + o.rows * kLdF; i += kHeadThreads) sm[o.xh + i] = 0.f;
  __syncthreads();

  float* s_xh = sm + o.xh;
  float* s_df = sm + o.df;
  float* s_dr = sm + o.dr;
  float* s_ctx = sm + o.ctx;
  const uint32_t* s_fbits = reinterpret_cast<const uint32_t*>(sm + o.fbits);
  const float* s_rstd = sm + o.rstd;

  // accumulators that live in registers for the whole kernel (MMA accumulator layout)
  float g_w[2][4];     // B2 unit: dW1 (warps 0-7) / dWo (warps 8-11) rows 16 mt + {g, g+8}, cols 16 nh + 8 nt + {2t, 2t+1}
  float g_b[4];        // B2, nh == 0: db1 / dbo of those rows (every column of the ones-tile holds the sum)
  float g_gam[4][2], g_bet[4][2];   // warps 0-6: d(LayerNorm weight / bias) of columns 8 nt + {2t, 2t+1}, this lane's rows
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    g_w[0][i] = g_w[1][i] = g_b[i] = 0.f;
    g_gam[i][0] = g_gam[i][1] = g_bet[i][0] = g_bet[i][1] = 0.f;
  }

  const float inv_a = a.inv_a, inv_f = a.inv_f;
  const uint32_t one = __float_as_uint(1.f);
  const int n_row_tiles = (S + 15) >> 4;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    float* st = a.state + static_cast<long long>(b) * a.sl.stride;
    __syncthreads();     // every warp is done with the previous sample's operands
    {
      const float* gx = st + a.sl.xhat;
      const float* gc = st + a.sl.ctx;
      const float* gd = a.dfeat + static_cast<long long>(b) * KF;
      for (int i = tid; i < S * (kE / 4); i += kHeadThreads) {
        const int row = i >> 3, c4 = (i & 7) * 4;
        ptx::cp_async_16(ptx::smem_u32(s_xh + row * kLdT + c4), gx + row * kE + c4);
        ptx::cp_async_16(ptx::smem_u32(s_ctx + row * kLdT + c4), gc + row * kE + c4);
      }
      for (int i = tid; i < S * (kF / 4); i += kHeadThreads) {
        const int row = i >> 4, c4 = (i & 15) * 4;
        ptx::cp_async_16(ptx::smem_u32(s_df + row * kLdF + c4), gd + row * kF + c4);
      }
      if (tid < S4 / 2) ptx::cp_async_16(ptx::smem_u32(sm + o.fbits + 4 * tid), st + a.sl.fbits + 4 * tid);
      if (tid < S4 / 4) ptx::cp_async_16(ptx::smem_u32(sm + o.rstd + 4 * tid), st + a.sl.rstd + 4 * tid);
      ptx::cp_async_commit();
      ptx::cp_async_wait_all();
    }
    __syncthreads();

    // ---- df = dfeat * ReLU' * dropout: thread = (position, feature pair 2j, 2j + 1) --------------
    for (int i = tid; i < S * (kF / 2); i += kHeadThreads) {
      const int s = i >> 5, j = i & 31;
      float2 d = *reinterpret_cast<const float2*>(s_df + s * kLdF + 2 * j);
      d.x = ((s_fbits[2 * s] >> j) & 1u) ? d.x * inv_f : 0.f;
      d.y = ((s_fbits[2 * s + 1] >> j) & 1u) ? d.y * inv_f : 0.f;
      *reinterpret_cast<float2*>(s_df + s * kLdF + 2 * j) = d;
    }
    __syncthreads();

    // ---- B1: dh = df W1 ; LayerNorm backward -> dr ; h.  B1b: dctx = dr Wo, D ------------------
    if (warp < n_row_tiles) {
      const int r0 = 16 * warp + g, r1 = r0 + 8;
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll 2
      for (int ks = 0; ks < kF / 8; ++ks) {
        uint32_t ah[4], al[4];
        ptx::split_tf32(s_df[r0 * kLdF + 8 * ks + t], ah[0], al[0]);
        ptx::split_tf32(s_df[r1 * kLdF + 8 * ks + t], ah[1], al[1]);
        ptx::split_tf32(s_df[r0 * kLdF + 8 * ks + t + 4], ah[2], al[2]);
        ptx::split_tf32(s_df[r1 * kLdF + 8 * ks + t + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int wi = (8 * ks + t) * kLdT + 8 * nt + g;       // W1[feature 8 ks + t (+ 4)][channel]
          ptx::mma_3xtf32(acc[nt], ah, al, w1_hi[wi], w1_hi[wi + 4 * kLdT], w1_lo[wi], w1_lo[wi + 4 * kLdT]);
        }
      }
      // LayerNorm backward on the fragments: rows r0 (elements 0, 1) and r1 (2, 3), 8 columns per lane
      float xa[4][2], xb[4][2];
      float m1a = 0.f, m2a = 0.f, m1b = 0.f, m2b = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = 8 * nt + 2 * t;
        const float2 x0 = *reinterpret_cast<const float2*>(s_xh + r0 * kLdT + c);
        const float2 x1 = *reinterpret_cast<const float2*>(s_xh + r1 * kLdT + c);
        const float gm0 = sm[o.lnw + c], gm1 = sm[o.lnw + c + 1];
        xa[nt][0] = x0.x; xa[nt][1] = x0.y; xb[nt][0] = x1.x; xb[nt][1] = x1.y;
        // rows beyond S: df = 0 -> dh = 0, xhat = 0: they add nothing
        g_gam[nt][0] = fmaf(acc[nt][0], x0.x, g_gam[nt][0]); g_gam[nt][1] = fmaf(acc[nt][1], x0.y, g_gam[nt][1]);
        g_gam[nt][0] = fmaf(acc[nt][2], x1.x, g_gam[nt][0]); g_gam[nt][1] = fmaf(acc[nt][3], x1.y, g_gam[nt][1]);
        g_bet[nt][0] += acc[nt][0] + acc[nt][2];
        g_bet[nt][1] += acc[nt][1] + acc[nt][3];
        acc[nt][0] *= gm0; acc[nt][1] *= gm1; acc[nt][2] *= gm0; acc[nt][3] *= gm1;   // dh * gamma
        m1a += acc[nt][0] + acc[nt][1];
        m2a = fmaf(acc[nt][0], x0.x, m2a); m2a = fmaf(acc[nt][1], x0.y, m2a);
        m1b += acc[nt][2] + acc[nt][3];
        m2b = fmaf(acc[nt][2], x1.x, m2b); m2b = fmaf(acc[nt][3], x1.y, m2b);
      }
#pragma unroll
      for (int msk = 1; msk <= 2; msk <<= 1) {
        m1a += __shfl_xor_sync(0xffffffffu, m1a, msk); m2a += __shfl_xor_sync(0xffffffffu, m2a, msk);
        m1b += __shfl_xor_sync(0xffffffffu, m1b, msk); m2b += __shfl_xor_sync(0xffffffffu, m2b, msk);
      }
      m1a *= (1.f / kE); m2a *= (1.f / kE); m1b *= (1.f / kE); m2b *= (1.f / kE);
      const float rs_a = r0 < S ? s_rstd[r0] : 0.f, rs_b = r1 < S ? s_rstd[r1] : 0.f;
      float* g_dr = st + a.sl.dr;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
        const int c = 8 * nt + 2 * t;
        const float gm0 = sm[o.lnw + c], gm1 = sm[o.lnw + c + 1], be0 = sm[o.lnb + c], be1 = sm[o.lnb + c + 1];
        if (r0 < S) {
          const float2 d = make_float2(rs_a * (acc[nt][0] - m1a - xa[nt][0] * m2a), rs_a * (acc[nt][1] - m1a - xa[nt][1] * m2a));
          *reinterpret_cast<float2*>(s_dr + r0 * kLdT + c) = d;
          *reinterpret_cast<float2*>(g_dr + r0 * kLdT + c) = d;                 // for the tail kernel
          *reinterpret_cast<float2*>(s_xh + r0 * kLdT + c) = make_float2(fmaf(xa[nt][0], gm0, be0), fmaf(xa[nt][1], gm1, be1));   // h, for dW1
        }
        if (r1 < S) {
          const float2 d = make_float2(rs_b * (acc[nt][2] - m1b - xb[nt][0] * m2b), rs_b * (acc[nt][3] - m1b - xb[nt][1] * m2b));
          *reinterpret_cast<float2*>(s_dr + r1 * kLdT + c) = d;
          *reinterpret_cast<float2*>(g_dr + r1 * kLdT + c) = d;
          *reinterpret_cast<float2*>(s_xh + r1 * kLdT + c) = make_float2(fmaf(xb[nt][0], gm0, be0), fmaf(xb[nt][1], gm1, be1));
        }
      }
      __syncwarp();
      // B1b: d(ctx) = dr Wo for the same rows; D[s][head] = sum over the head's 8 channels of d(ctx) * ctx
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < kE / 8; ++ks) {
        uint32_t ah[4], al[4];
        ptx::split_tf32(s_dr[r0 * kLdT + 8 * ks + t], ah[0], al[0]);
        ptx::split_tf32(s_dr[r1 * kLdT + 8 * ks + t], ah[1], al[1]);
        ptx::split_tf32(s_dr[r0 * kLdT + 8 * ks + t + 4], ah[2], al[2]);
        ptx::split_tf32(s_dr[r1 * kLdT + 8 * ks + t + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          const int wi = (8 * ks + t) * kLdT + 8 * nt + g;       // Wo[out channel 8 ks + t (+ 4)][ctx channel]
          ptx::mma_3xtf32(acc[nt], ah, al, wo_hi[wi], wo_hi[wi + 4 * kLdT], wo_lo[wi], wo_lo[wi + 4 * kLdT]);
        }
      }
      float* g_dc = st + a.sl.dctx;
      float* g_stat = st + a.sl.stat;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {          // column tile nt = head nt
        const int c = 8 * nt + 2 * t;
        const float2 ca = *reinterpret_cast<const float2*>(s_ctx + r0 * kLdT + c);
        const float2 cb = *reinterpret_cast<const float2*>(s_ctx + r1 * kLdT + c);
        float pa = acc[nt][0] * ca.x + acc[nt][1] * ca.y, pb = acc[nt][2] * cb.x + acc[nt][3] * cb.y;
        pa += __shfl_xor_sync(0xffffffffu, pa, 1); pb += __shfl_xor_sync(0xffffffffu, pb, 1);
        pa += __shfl_xor_sync(0xffffffffu, pa, 2); pb += __shfl_xor_sync(0xffffffffu, pb, 2);
        if (r0 < S) {
          *reinterpret_cast<float2*>(g_dc + r0 * kE + c) = make_float2(acc[nt][0] * inv_a, acc[nt][1] * inv_a);   // d(P_dropped) carries 1 / (1 - p)
          if (t == 0) g_stat[(r0 * kHeads + nt) * 4 + 2] = pa;
        }
        if (r1 < S) {
          *reinterpret_cast<float2*>(g_dc + r1 * kE + c) = make_float2(acc[nt][2] * inv_a, acc[nt][3] * inv_a);
          if (t == 0) g_stat[(r1 * kHeads + nt) * 4 + 2] = pb;
        }
      }
    }
    __syncthreads();

    // ---- B2: dW1, db1 (warps 0-7) | dWo, dbo (warps 8-11) --------------------------------------
    if (warp < 12) {
      const bool w1part = warp < 8;
      const int mt = w1part ? (warp >> 1) : ((warp - 8) >> 1), nh = warp & 1;
      // A^T: rows of the tile = features (df) / out-projection outputs (dr) 16 mt + {g, g+8}
      const float* src = w1part ? s_df + 16 * mt + g : s_dr + 16 * mt + g;
      const int lds = w1part ? kLdF : kLdT;
      const float* bsrc = (w1part ? s_xh : s_ctx) + 16 * nh + g;       // h (dW1) / ctx (dWo)
      const int nk = (S + 7) >> 3;
#pragma unroll 2
      for (int ks = 0; ks < nk; ++ks) {
        const int k0 = 8 * ks + t;
        uint32_t ah[4], al[4];
        ptx::split_tf32(src[k0 * lds], ah[0], al[0]);
        ptx::split_tf32(src[k0 * lds + 8], ah[1], al[1]);
        ptx::split_tf32(src[(k0 + 4) * lds], ah[2], al[2]);
        ptx::split_tf32(src[(k0 + 4) * lds + 8], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          uint32_t bh0, bl0, bh1, bl1;
          ptx::split_tf32(bsrc[k0 * kLdT + 8 * nt], bh0, bl0);
          ptx::split_tf32(bsrc[(k0 + 4) * kLdT + 8 * nt], bh1, bl1);
          ptx::mma_3xtf32(g_w[nt], ah, al, bh0, bh1, bl0, bl1);
        }
        if (nh == 0) {   // warp-uniform: column sums through a tile of ones
          ptx::mma_tf32(g_b, al, one, one);
          ptx::mma_tf32(g_b, ah, one, one);
        }
      }
    }
  }
  __syncthreads();

  // ---- flush this CTA's partial sums ------------------------------------------------------
  if (warp < 7) {
    // d(LayerNorm weight / bias): sum this lane's columns over the 8 row groups g, then over the warps
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float x = g_gam[nt][e], y = g_bet[nt][e];
#pragma unroll
        for (int msk = 4; msk <= 16; msk <<= 1) {
          x += __shfl_xor_sync(0xffffffffu, x, msk);
          y += __shfl_xor_sync(0xffffffffu, y, msk);
        }
        if (g == 0) {
          sm[o.red + warp * kE + 8 * nt + 2 * t + e] = x;
          sm[o.red + (7 + warp) * kE + 8 * nt + 2 * t + e] = y;
        }
      }
  }
  __syncthreads();
  if (tid < 2 * kE) {
    const int which = tid >> 5, c = tid & 31;
    float acc = 0.f;
    for (int w = 0; w < 7; ++w) acc += sm[o.red + (which * 7 + w) * kE + c];
    part[(which == 0 ? a.lay.off_lnw : a.lay.off_lnb) + c] = acc;
  }
  if (warp < 12) {
    const bool w1part = warp < 8;
    const int mt = w1part ? (warp >> 1) : ((warp - 8) >> 1), nh = warp & 1;
    float* wbase = part + (w1part ? a.lay.off_w1 : a.lay.off_wo);
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      float* w0 = wbase + (16 * mt + g) * kE + 16 * nh + 8 * nt + 2 * t;
      w0[0] = g_w[nt][0]; w0[1] = g_w[nt][1];
      w0[8 * kE] = g_w[nt][2]; w0[8 * kE + 1] = g_w[nt][3];
    }
    if (nh == 0 && t == 0) {
      float* bbase = part + (w1part ? a.lay.off_b1 : a.lay.off_bo);
      bbase[16 * mt + g] = g_b[0];
      bbase[16 * mt + g + 8] = g_b[2];
    }
  }
}

__global__ void __launch_bounds__(kHeadThreads, 1) frontend_backward_head_kernel(const FrontArgs a) {
  frontend_backward_head_body(a);
}
// Register-capped variants of the backward kernels (a few spills): 14 warps of 128 registers fill
// the scheduler partitions' register files and no warp of another kernel fits beside them; capped,
// the background AdamW sweep (afr_adamw_rows_bg, 4 warps x 40 registers) shares the SM.
__global__ void __maxnreg__(112) frontend_backward_head_kernel_shared(const FrontArgs a) {
  frontend_backward_head_body(a);
}

// ---------------------------------------------------------------- K2: attention backward
// Per (sample, head): P = softmax(q k^T) recomputed from the forward's row statistics,
//   dP = d(ctx) V^T / (1-p),  dS = P o (keep ? dP : 0  -  D),  dq = dS k,  dk = dS^T q,  dv = (P o keep)^T d(ctx) / (1-p)
// on mma.sync.m16n8k8 TF32 with every product as three MMAs (hi/lo split of both operands,
// ptx::mma_3xtf32: fp32-equivalent, the 5e-5 gradient tolerance holds). head_dim = 8 is exactly one
// k-step. A warp owns a (head, 16-row tile): pass 1 (rows = queries) walks the key tiles and
// accumulates dq; pass 2 (rows = keys) walks the query tiles and accumulates dk, dv. The 16 x 8
// tile of dS / P comes out of the score MMA in the accumulator layout and goes back in as the A
// operand of the next MMA by renaming the k index (k = t <-> column 2t, k = t+4 <-> column 2t+1;
// the B fragment reads rows 2t, 2t+1 to match): no shuffle, no shared-memory round trip.
// Operands sit in shared memory as fp32 rows of kE + 4 floats (every fragment load hits 32
// distinct banks), staged with cp.async; rows / statistics beyond S are zero, which makes the
// padding of the last tiles inert (P = 0 there). Nothing is written to shared memory after the
// staging, so the two passes need no barrier between them.
constexpr int kAttWarps = 14;
constexpr int kAttThreads = kAttWarps * 32;
constexpr int kLdA = kE + 4;

struct AttSmem {
  int q, k, v, d, stat, abits, rows, total;
};
__host__ __device__ inline AttSmem make_att_smem(int L) {
  AttSmem s{};
  s.rows = (L + 15) & ~15;
  int o = 0;
  s.q = o; o += s.rows * kLdA;
  s.k = o; o += s.rows * kLdA;
  s.v = o; o += s.rows * kLdA;
  s.d = o; o += s.rows * kLdA;
  s.stat = o; o += kHeads * 3 * s.rows;   // [head][m | 1/l | D][row]
  s.abits = o; o += kHeads * L * 4;
  s.total = o;
  return s;
}

__device__ __forceinline__ void frontend_backward_attn_body(const FrontArgs& a) {
  extern __shared__ __align__(16) float sm[];
  const AttSmem o = make_att_smem(a.L);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int S = a.S, RP = o.rows;
  float* sq = sm + o.q;
  float* sk = sm + o.k;
  float* sv = sm + o.v;
  float* sd = sm + o.d;
  float* sst = sm + o.stat;
  uint32_t* sab = reinterpret_cast<uint32_t*>(sm + o.abits);
  for (int i = tid; i < o.abits; i += kAttThreads) sm[i] = 0.f;   // padding rows / statistics stay zero
  __syncthreads();

  const int nmt = (S + 15) >> 4, nt8 = (S + 7) >> 3, units = nmt * kHeads;
  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    float* st = a.state + static_cast<long long>(b) * a.sl.stride;
    for (int i = tid; i < S * 8; i += kAttThreads) {
      const int row = i >> 3, c4 = (i & 7) * 4;
      const int dst = row * kLdA + c4, src = row * kE + c4;
      ptx::cp_async_16(ptx::smem_u32(sq + dst), st + a.sl.q + src);
      ptx::cp_async_16(ptx::smem_u32(sk + dst), st + a.sl.k + src);
      ptx::cp_async_16(ptx::smem_u32(sv + dst), st + a.sl.v + src);
      ptx::cp_async_16(ptx::smem_u32(sd + dst), st + a.sl.dctx + src);
    }
    for (int i = tid; i < kHeads * S; i += kAttThreads)
      ptx::cp_async_16(ptx::smem_u32(sab + 4 * i), st + a.sl.abits + 4 * i);
    ptx::cp_async_commit();
    for (int i = tid; i < S * kHeads; i += kAttThreads) {
      const float4 x = *reinterpret_cast<const float4*>(st + a.sl.stat + 4 * i);   // (m, 1/l, D, -)
      const int s = i >> 2, h = i & 3;
      sst[(h * 3 + 0) * RP + s] = x.x;
      sst[(h * 3 + 1) * RP + s] = x.y;
      sst[(h * 3 + 2) * RP + s] = x.z;
    }
    ptx::cp_async_wait_all();
    __syncthreads();

    for (int u = warp; u < 2 * units; u += kAttWarps) {
      const bool pass1 = u < units;
      const int uu = pass1 ? u : u - units;
      const int h = uu & 3, mt = uu >> 2, hc = kDh * h;
      const int r0 = 16 * mt + g, r1 = r0 + 8;
      const float* sth = sst + h * 3 * RP;
      if (pass1) {
        // ---- rows = queries r0, r1; stream over key tiles -> dq
        uint32_t qh[4], ql[4], dh[4], dl[4];
        {
          ptx::split_tf32(sq[r0 * kLdA + hc + t], qh[0], ql[0]);
          ptx::split_tf32(sq[r1 * kLdA + hc + t], qh[1], ql[1]);
          ptx::split_tf32(sq[r0 * kLdA + hc + t + 4], qh[2], ql[2]);
          ptx::split_tf32(sq[r1 * kLdA + hc + t + 4], qh[3], ql[3]);
          ptx::split_tf32(sd[r0 * kLdA + hc + t], dh[0], dl[0]);
          ptx::split_tf32(sd[r1 * kLdA + hc + t], dh[1], dl[1]);
          ptx::split_tf32(sd[r0 * kLdA + hc + t + 4], dh[2], dl[2]);
          ptx::split_tf32(sd[r1 * kLdA + hc + t + 4], dh[3], dl[3]);
        }
        const float m0 = sth[r0], m1 = sth[r1], l0 = sth[RP + r0], l1 = sth[RP + r1];
        const float D0 = sth[2 * RP + r0], D1 = sth[2 * RP + r1];
        const uint32_t* bp0 = sab + (h * S + min(r0, S - 1)) * 4;
        const uint32_t* bp1 = sab + (h * S + min(r1, S - 1)) * 4;
        float dq[4] = {0.f, 0.f, 0.f, 0.f};
        for (int w = 0; 32 * w < S; ++w) {
          const uint32_t w0 = bp0[w] >> (2 * t), w1 = bp1[w] >> (2 * t);   // this lane's key pair, tile 0
#pragma unroll
          for (int j4 = 0; j4 < 4; ++j4) {
            const int j = 4 * w + j4;
            if (8 * j < S) {   // warp-uniform
              uint32_t bh0, bh1, bl0, bl1;
              const float* kr = sk + (8 * j + g) * kLdA + hc + t;
              ptx::split_tf32(kr[0], bh0, bl0);
              ptx::split_tf32(kr[4], bh1, bl1);
              float c[4] = {0.f, 0.f, 0.f, 0.f};
              ptx::mma_3xtf32(c, qh, ql, bh0, bh1, bl0, bl1);        // scores (log2 domain)
              const float* vr = sv + (8 * j + g) * kLdA + hc + t;
              ptx::split_tf32(vr[0], bh0, bl0);
              ptx::split_tf32(vr[4], bh1, bl1);
              float e[4] = {0.f, 0.f, 0.f, 0.f};
              ptx::mma_3xtf32(e, dh, dl, bh0, bh1, bl0, bl1);        // dP
              const uint32_t k0 = w0 >> (8 * j4), k1 = w1 >> (8 * j4);
              const float p0 = ex2(c[0] - m0) * l0, p1 = ex2(c[1] - m0) * l0;
              const float p2 = ex2(c[2] - m1) * l1, p3 = ex2(c[3] - m1) * l1;
              float ds0 = p0 * (((k0 & 1u) ? e[0] : 0.f) - D0);
              float ds1 = p1 * (((k0 & 2u) ? e[1] : 0.f) - D0);
              float ds2 = p2 * (((k1 & 1u) ? e[2] : 0.f) - D1);
              float ds3 = p3 * (((k1 & 2u) ? e[3] : 0.f) - D1);
              if (8 * j + 8 > S) {   // keys beyond S in the last tile
                const int key = 8 * j + 2 * t;
                if (key >= S) { ds0 = 0.f; ds2 = 0.f; }
                if (key + 1 >= S) { ds1 = 0.f; ds3 = 0.f; }
              }
              uint32_t ah[4], al[4];
              ptx::split_tf32(ds0, ah[0], al[0]);
              ptx::split_tf32(ds2, ah[1], al[1]);
              ptx::split_tf32(ds1, ah[2], al[2]);
              ptx::split_tf32(ds3, ah[3], al[3]);
              const float* kt = sk + (8 * j + 2 * t) * kLdA + hc + g;
              ptx::split_tf32(kt[0], bh0, bl0);
              ptx::split_tf32(kt[kLdA], bh1, bl1);
              ptx::mma_3xtf32(dq, ah, al, bh0, bh1, bl0, bl1);       // dq += dS k
            }
          }
        }
        // d(q) of the unscaled projection: d(score) * k / sqrt(head_dim)
        float* gq = st + a.sl.dq + hc + 2 * t;
        if (r0 < S) *reinterpret_cast<float2*>(gq + r0 * kLdT) = make_float2(dq[0] * kInvSqrtDh, dq[1] * kInvSqrtDh);
        if (r1 < S) *reinterpret_cast<float2*>(gq + r1 * kLdT) = make_float2(dq[2] * kInvSqrtDh, dq[3] * kInvSqrtDh);
      } else {
        // ---- rows = keys r0, r1; stream over query tiles -> dk, dv
        uint32_t kh[4], kl[4], vh[4], vl[4];
        ptx::split_tf32(sk[r0 * kLdA + hc + t], kh[0], kl[0]);
        ptx::split_tf32(sk[r1 * kLdA + hc + t], kh[1], kl[1]);
        ptx::split_tf32(sk[r0 * kLdA + hc + t + 4], kh[2], kl[2]);
        ptx::split_tf32(sk[r1 * kLdA + hc + t + 4], kh[3], kl[3]);
        ptx::split_tf32(sv[r0 * kLdA + hc + t], vh[0], vl[0]);
        ptx::split_tf32(sv[r1 * kLdA + hc + t], vh[1], vl[1]);
        ptx::split_tf32(sv[r0 * kLdA + hc + t + 4], vh[2], vl[2]);
        ptx::split_tf32(sv[r1 * kLdA + hc + t + 4], vh[3], vl[3]);
        float dk[4] = {0.f, 0.f, 0.f, 0.f}, dv[4] = {0.f, 0.f, 0.f, 0.f};
        const int wsel = mt >> 1, sh0 = 16 * (mt & 1) + g;   // key r0 = bit sh0, key r1 = bit sh0 + 8 of word wsel
#pragma unroll 2
        for (int j = 0; j < nt8; ++j) {
          uint32_t bh0, bh1, bl0, bl1;
          const float* qr = sq + (8 * j + g) * kLdA + hc + t;
          ptx::split_tf32(qr[0], bh0, bl0);
          ptx::split_tf32(qr[4], bh1, bl1);
          float c[4] = {0.f, 0.f, 0.f, 0.f};
          ptx::mma_3xtf32(c, kh, kl, bh0, bh1, bl0, bl1);            // scores^T
          const float* dr = sd + (8 * j + g) * kLdA + hc + t;
          ptx::split_tf32(dr[0], bh0, bl0);
          ptx::split_tf32(dr[4], bh1, bl1);
          float e[4] = {0.f, 0.f, 0.f, 0.f};
          ptx::mma_3xtf32(e, vh, vl, bh0, bh1, bl0, bl1);            // dP^T
          const int q0 = 8 * j + 2 * t;                              // this lane's query pair
          const float2 mm = *reinterpret_cast<const float2*>(sth + q0);
          const float2 ll = *reinterpret_cast<const float2*>(sth + RP + q0);
          const float2 DD = *reinterpret_cast<const float2*>(sth + 2 * RP + q0);
          const uint32_t ba = sab[(h * S + min(q0, S - 1)) * 4 + wsel] >> sh0;
          const uint32_t bb = sab[(h * S + min(q0 + 1, S - 1)) * 4 + wsel] >> sh0;
          const float p0 = ex2(c[0] - mm.x) * ll.x, p1 = ex2(c[1] - mm.y) * ll.y;
          const float p2 = ex2(c[2] - mm.x) * ll.x, p3 = ex2(c[3] - mm.y) * ll.y;
          const bool e0 = ba & 1u, e1 = bb & 1u, e2 = ba & 0x100u, e3 = bb & 0x100u;
          const float ds0 = p0 * ((e0 ? e[0] : 0.f) - DD.x), ds1 = p1 * ((e1 ? e[1] : 0.f) - DD.y);
          const float ds2 = p2 * ((e2 ? e[2] : 0.f) - DD.x), ds3 = p3 * ((e3 ? e[3] : 0.f) - DD.y);
          uint32_t ah[4], al[4];
          ptx::split_tf32(ds0, ah[0], al[0]);
          ptx::split_tf32(ds2, ah[1], al[1]);
          ptx::split_tf32(ds1, ah[2], al[2]);
          ptx::split_tf32(ds3, ah[3], al[3]);
          const float* qt = sq + (8 * j + 2 * t) * kLdA + hc + g;
          ptx::split_tf32(qt[0], bh0, bl0);
          ptx::split_tf32(qt[kLdA], bh1, bl1);
          ptx::mma_3xtf32(dk, ah, al, bh0, bh1, bl0, bl1);           // dk += dS^T q
          ptx::split_tf32(e0 ? p0 : 0.f, ah[0], al[0]);
          ptx::split_tf32(e2 ? p2 : 0.f, ah[1], al[1]);
          ptx::split_tf32(e1 ? p1 : 0.f, ah[2], al[2]);
          ptx::split_tf32(e3 ? p3 : 0.f, ah[3], al[3]);
          const float* dt = sd + (8 * j + 2 * t) * kLdA + hc + g;
          ptx::split_tf32(dt[0], bh0, bl0);
          ptx::split_tf32(dt[kLdA], bh1, bl1);
          ptx::mma_3xtf32(dv, ah, al, bh0, bh1, bl0, bl1);           // dv += (P o keep)^T d(ctx)/(1-p)
        }
        // the stored q carries log2(e)/sqrt(head_dim): d(k) = sum ds * q / sqrt(head_dim)
        float* gk = st + a.sl.dk + hc + 2 * t;
        float* gv = st + a.sl.dv + hc + 2 * t;
        if (r0 < S) {
          *reinterpret_cast<float2*>(gk + r0 * kLdT) = make_float2(dk[0] * kLn2, dk[1] * kLn2);
          *reinterpret_cast<float2*>(gv + r0 * kLdT) = make_float2(dv[0], dv[1]);
        }
        if (r1 < S) {
          *reinterpret_cast<float2*>(gk + r1 * kLdT) = make_float2(dk[2] * kLn2, dk[3] * kLn2);
          *reinterpret_cast<float2*>(gv + r1 * kLdT) = make_float2(dv[2], dv[3]);
        }
      }
    }
    __syncthreads();   // every warp is done with this sample's operands
  }
}

__global__ void __launch_bounds__(kAttThreads, 2) frontend_backward_attn_kernel(const FrontArgs a) {
  frontend_backward_attn_body(a);
}
__global__ void __maxnreg__(64) frontend_backward_attn_kernel_shared(const FrontArgs a) {
  frontend_backward_attn_body(a);
}

// ---------------------------------------------------------------- K3: tail of the backward
// de = dr + [dq dk dv] Win (B4) and dWin += [dq dk dv]^T e, dbin += column sums (B5) as 3xTF32 warp
// MMAs (the FFMA2 forms were one LDS per 1-2 FMAs: 31 k warp instructions per sample); the
// embedding scatter-add stays a position-ordered walk by one warp. All operands are rows of kLdT =
// 40 floats (the producers write them that way into the record): 8t + g is a distinct bank for
// every lane of a B / A^T fragment load. 14 warps: B4 = 7 row tiles x 2 column halves, B5 = 6
// channel tiles x 2 column halves (+ a ones-tile for dbin) beside the scatter warp.
constexpr int kTailWarps = 14;
constexpr int kTailThreads = kTailWarps * 32;

struct TailSmem {
  int win_hi, win_lo, dq, dk, dv, dr, e, de, ebits, tok, hist, fonth, rows, total;
};
__host__ __device__ inline TailSmem make_tail_smem(int L, int vocab) {
  const int L4 = (L + 3) & ~3;
  TailSmem s{};
  s.rows = (L + 15) & ~15;
  int o = 0;
  s.win_hi = o; o += 3 * kE * kLdT;     // Win split once per CTA: TF32 hi / lo words
  s.win_lo = o; o += 3 * kE * kLdT;
  s.dq = o;  o += s.rows * kLdT;
  s.dk = o;  o += s.rows * kLdT;
  s.dv = o;  o += s.rows * kLdT;
  s.e = o;   o += s.rows * kLdT;
  s.dr = o;  o += L * kLdT;
  s.de = o;  o += L * kE;               // d(embedding rows) after the dropout mask, for the scatter walk
  s.ebits = o; o += L4;
  s.tok = o;  o += L4;
  s.hist = o; o += vocab <= kEmbSmemMaxVocab ? vocab * kE : 0;
  s.fonth = o; o += kMaxFonts * kE;   // d(font_embedding) of this CTA's samples
  s.total = (o + 3) & ~3;
  return s;
}

__device__ __forceinline__ void frontend_backward_tail_body(const FrontArgs& a) {
  extern __shared__ __align__(16) float sm[];
  const TailSmem o = make_tail_smem(a.L, a.vocab);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int S = a.S, S4 = (S + 3) & ~3;
  const bool hist_smem = a.vocab <= kEmbSmemMaxVocab;
  float* part = a.partials + static_cast<long long>(blockIdx.x) * a.lay.total;

  uint32_t* win_hi = reinterpret_cast<uint32_t*>(sm + o.win_hi);
  uint32_t* win_lo = reinterpret_cast<uint32_t*>(sm + o.win_lo);
  for (int i = tid; i < 3 * kE * kE; i += kTailThreads) {
    uint32_t hi, lo;
    ptx::split_tf32(a.w.win[i], hi, lo);
    win_hi[(i / kE) * kLdT + (i % kE)] = hi;
    win_lo[(i / kE) * kLdT + (i % kE)] = lo;
  }
  // rows >= S of dq, dk, dv, e stay zero: they pad the reduction over positions of B5
  for (int i = tid; i < 4 * o.rows * kLdT; i += kTailThreads) sm[o.dq + i] = 0.f;
  if (hist_smem) {
    for (int i = tid; i < a.vocab * kE; i += kTailThreads) sm[o.hist + i] = 0.f;
  } else {
    for (int i = tid; i < a.vocab * kE; i += kTailThreads) part[a.lay.off_emb + i] = 0.f;
  }
  for (int i = tid; i < kMaxFonts * kE; i += kTailThreads) sm[o.fonth + i] = 0.f;
  __syncthreads();

  float* s_dq = sm + o.dq;
  float* s_e = sm + o.e;
  float* s_dr = sm + o.dr;
  float* s_de = sm + o.de;
  const uint32_t* s_ebits = reinterpret_cast<const uint32_t*>(sm + o.ebits);
  int* s_tok = reinterpret_cast<int*>(sm + o.tok);

  // accumulators that live in registers for the whole kernel (MMA accumulator layout)
  float g_pos[2][4];   // B4 unit (mt = warp >> 1, nh = warp & 1): d(pos) rows 16 mt + {g, g+8}, cols 16 nh + 8 nt + {2t, 2t+1}
  float g_win[2][4];   // B5 unit (warps 0-11): dWin rows 16 mt + {g, g+8}, same columns
  float g_bin[4];      // B5, nh == 0: dbin rows 16 mt + {g, g+8} (every column of the ones-tile holds the sum)
#pragma unroll
  for (int i = 0; i < 4; ++i) { g_pos[0][i] = g_pos[1][i] = g_win[0][i] = g_win[1][i] = g_bin[i] = 0.f; }

  const float inv_e = a.inv_e;
  const int mt = warp >> 1, nh = warp & 1;
  const uint32_t one = __float_as_uint(1.f);

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    __syncthreads();     // every warp is done with the previous sample's operands
    {
      // cp.async by all threads: group 0 = dq, dk, dv, dr, keep bits; group 1 = e (needed by B5 only).
      // (cp.async.bulk + mbarrier, a double-buffered variant and an L2 prefetch of the next sample
      // all measure the same 0.071-0.073 ms: the wait at the top is not what bounds this kernel.)
      const float* st = a.state + static_cast<long long>(b) * a.sl.stride;
      const int chunks = S * kLdT / 4;
      for (int i = tid; i < chunks; i += kTailThreads) {
        ptx::cp_async_16(ptx::smem_u32(s_dq + 4 * i), st + a.sl.dq + 4 * i);
        ptx::cp_async_16(ptx::smem_u32(sm + o.dk + 4 * i), st + a.sl.dk + 4 * i);
        ptx::cp_async_16(ptx::smem_u32(sm + o.dv + 4 * i), st + a.sl.dv + 4 * i);
        ptx::cp_async_16(ptx::smem_u32(s_dr + 4 * i), st + a.sl.dr + 4 * i);
      }
      if (tid < S4 / 4) ptx::cp_async_16(ptx::smem_u32(sm + o.ebits + 4 * tid), st + a.sl.ebits + 4 * tid);
      ptx::cp_async_commit();
      for (int i = tid; i < chunks; i += kTailThreads)
        ptx::cp_async_16(ptx::smem_u32(s_e + 4 * i), st + a.sl.e + 4 * i);
      ptx::cp_async_commit();
    }
    if (tid < S) {
      long long tk = a.tokens[static_cast<long long>(b) * a.token_stride + tid];
      if (tk < 0 || tk >= a.vocab) tk = 0;
      s_tok[tid] = static_cast<int>(tk);
    }
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncthreads();

    // ---- B4: de = dr + [dq dk dv] Win ; dPos ; d(embedding rows) ------------------------------
    if (16 * mt < S) {
      const int r0 = 16 * mt + g, r1 = r0 + 8;
      float acc[2][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[0][i] = acc[1][i] = 0.f;
#pragma unroll 4
      for (int ks = 0; ks < 12; ++ks) {       // reduction index 32 j + kk: channel kk of dq (j = 0), dk, dv
        const float* src = s_dq + (ks >> 2) * (o.rows * kLdT) + 8 * (ks & 3) + t;
        uint32_t ah[4], al[4];
        ptx::split_tf32(src[r0 * kLdT], ah[0], al[0]);
        ptx::split_tf32(src[r1 * kLdT], ah[1], al[1]);
        ptx::split_tf32(src[r0 * kLdT + 4], ah[2], al[2]);
        ptx::split_tf32(src[r1 * kLdT + 4], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int wi = (8 * ks + t) * kLdT + 16 * nh + 8 * nt + g;     // Win[32 j + kk + t][column]
          ptx::mma_3xtf32(acc[nt], ah, al, win_hi[wi], win_hi[wi + 4 * kLdT], win_lo[wi], win_lo[wi + 4 * kLdT]);
        }
      }
      const uint32_t eb0 = r0 < S ? s_ebits[r0] : 0u, eb1 = r1 < S ? s_ebits[r1] : 0u;
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int col = 16 * nh + 8 * nt + 2 * t;
        if (r0 < S) {
          const float2 dr = *reinterpret_cast<const float2*>(s_dr + r0 * kLdT + col);
          const float d0 = acc[nt][0] + dr.x, d1 = acc[nt][1] + dr.y;
          g_pos[nt][0] += d0;
          g_pos[nt][1] += d1;
          // through the embedding dropout: scale or zero
          *reinterpret_cast<float2*>(s_de + r0 * kE + col) =
              make_float2(((eb0 >> col) & 1u) ? d0 * inv_e : 0.f, ((eb0 >> (col + 1)) & 1u) ? d1 * inv_e : 0.f);
        }
        if (r1 < S) {
          const float2 dr = *reinterpret_cast<const float2*>(s_dr + r1 * kLdT + col);
          const float d2 = acc[nt][2] + dr.x, d3 = acc[nt][3] + dr.y;
          g_pos[nt][2] += d2;
          g_pos[nt][3] += d3;
          *reinterpret_cast<float2*>(s_de + r1 * kE + col) =
              make_float2(((eb1 >> col) & 1u) ? d2 * inv_e : 0.f, ((eb1 >> (col + 1)) & 1u) ? d3 * inv_e : 0.f);
        }
      }
    }
    ptx::cp_async_wait_all();     // e
    __syncthreads();

    // ---- B5: dWin, dbin (warps 0-11) | embedding scatter-add (warp 12) ------------------------
    if (warp < 12) {
      // rows of the tile = channels 16 (mt & 1) + {g, g+8} of dq (mt = 0, 1), dk (2, 3), dv (4, 5)
      const float* src = s_dq + (mt >> 1) * (o.rows * kLdT) + 16 * (mt & 1) + g;
      const int nk = (S + 7) >> 3;
#pragma unroll 2
      for (int ks = 0; ks < nk; ++ks) {
        const int k0 = 8 * ks + t;
        uint32_t ah[4], al[4];
        ptx::split_tf32(src[k0 * kLdT], ah[0], al[0]);
        ptx::split_tf32(src[k0 * kLdT + 8], ah[1], al[1]);
        ptx::split_tf32(src[(k0 + 4) * kLdT], ah[2], al[2]);
        ptx::split_tf32(src[(k0 + 4) * kLdT + 8], ah[3], al[3]);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const float* er = s_e + k0 * kLdT + 16 * nh + 8 * nt + g;
          uint32_t bh0, bl0, bh1, bl1;
          ptx::split_tf32(er[0], bh0, bl0);
          ptx::split_tf32(er[4 * kLdT], bh1, bl1);
          ptx::mma_3xtf32(g_win[nt], ah, al, bh0, bh1, bl0, bl1);
        }
        if (nh == 0) {   // warp-uniform: column sums through a tile of ones
          ptx::mma_tf32(g_bin, al, one, one);
          ptx::mma_tf32(g_bin, ah, one, one);
        }
      }
    } else if (warp == 12) {
      // embedding scatter-add. Rows hit by several positions are summed in position order:
      // deterministic, no atomics. Operands are fetched four positions ahead of the dependent
      // read-modify-write chain on the table.
      float fsum = 0.f;    // d(font_embedding row of this sample)[lane] = sum over positions
      if (hist_smem) {
        float* hist = sm + o.hist + lane;
        int s = 0;
        for (; s + 4 <= S; s += 4) {
          const int t0 = s_tok[s], t1 = s_tok[s + 1], t2 = s_tok[s + 2], t3 = s_tok[s + 3];
          const float v0 = s_de[s * kE + lane], v1 = s_de[(s + 1) * kE + lane];
          const float v2 = s_de[(s + 2) * kE + lane], v3 = s_de[(s + 3) * kE + lane];
          hist[t0 * kE] += v0;
          hist[t1 * kE] += v1;
          hist[t2 * kE] += v2;
          hist[t3 * kE] += v3;
          fsum += (v0 + v1) + (v2 + v3);
        }
        for (; s < S; ++s) {
          const float v0 = s_de[s * kE + lane];
          hist[s_tok[s] * kE] += v0;
          fsum += v0;
        }
      } else {
        for (int s = 0; s < S; ++s) {
          float* pe = part + a.lay.off_emb + static_cast<long long>(s_tok[s]) * kE + lane;
          const float v0 = s_de[s * kE + lane];
          __stcg(pe, __ldcg(pe) + v0);
          fsum += v0;
        }
      }
      if (a.font.ids != nullptr) sm[o.fonth + a.font.ids[b] * kE + lane] += fsum;   // one writer per (font, lane)
    }
  }
  __syncthreads();

  // ---- flush this CTA's partial sums ------------------------------------------------------
  if (warp < 12) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const int col = 16 * nh + 8 * nt + 2 * t;
      float* w0 = part + a.lay.off_win + (16 * mt + g) * kE + col;
      w0[0] = g_win[nt][0]; w0[1] = g_win[nt][1];
      w0[8 * kE] = g_win[nt][2]; w0[8 * kE + 1] = g_win[nt][3];
    }
    if (nh == 0 && t == 0) {
      part[a.lay.off_bin + 16 * mt + g] = g_bin[0];
      part[a.lay.off_bin + 16 * mt + g + 8] = g_bin[2];
    }
  }
  // d(pos): rows of the tiles beyond L were never touched (stay zero) and are not part of the layout
#pragma unroll
  for (int nt = 0; nt < 2; ++nt) {
    const int col = 16 * nh + 8 * nt + 2 * t, r0 = 16 * mt + g, r1 = r0 + 8;
    if (r0 < a.L) { part[a.lay.off_pos + r0 * kE + col] = g_pos[nt][0]; part[a.lay.off_pos + r0 * kE + col + 1] = g_pos[nt][1]; }
    if (r1 < a.L) { part[a.lay.off_pos + r1 * kE + col] = g_pos[nt][2]; part[a.lay.off_pos + r1 * kE + col + 1] = g_pos[nt][3]; }
  }
  if (hist_smem)
    for (int i = tid; i < a.vocab * kE; i += kTailThreads) part[a.lay.off_emb + i] = sm[o.hist + i];
  for (int i = tid; i < kMaxFonts * kE; i += kTailThreads) part[a.lay.off_font + i] = sm[o.fonth + i];
}

__global__ void __launch_bounds__(kTailThreads, 1) frontend_backward_tail_kernel(const FrontArgs a) {
  frontend_backward_tail_body(a);
}
__global__ void __maxnreg__(112) frontend_backward_tail_kernel_shared(const FrontArgs a) {
  frontend_backward_tail_body(a);
}

// grads[i] = sum over CTAs of partials[cta][i], fixed order.
struct ReduceArgs {
  const float* partials;
  int grid, total;
  SmallLayout lay;
  Tensors g;
  float* font_grad;   // [n_fonts, kE] or nullptr
  int n_fonts;
};
__global__ void __launch_bounds__(256) small_grad_reduce_kernel(ReduceArgs r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r.total) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;   // four independent chains, fixed order
  int c = 0;
  for (; c + 4 <= r.grid; c += 4) {
    s0 += r.partials[static_cast<long long>(c) * r.total + i];
    s1 += r.partials[static_cast<long long>(c + 1) * r.total + i];
    s2 += r.partials[static_cast<long long>(c + 2) * r.total + i];
    s3 += r.partials[static_cast<long long>(c + 3) * r.total + i];
  }
  for (; c < r.grid; ++c) s0 += r.partials[static_cast<long long>(c) * r.total + i];
  const float s = (s0 + s1) + (s2 + s3);
  const SmallLayout& L = r.lay;
  float* dst;
  if (i < L.off_emb) dst = r.g.pos + (i - L.off_pos);
  else if (i < L.off_win) dst = r.g.emb + (i - L.off_emb);
  else if (i < L.off_bin) dst = r.g.win + (i - L.off_win);
  else if (i < L.off_wo) dst = r.g.bin + (i - L.off_bin);
  else if (i < L.off_bo) dst = r.g.wo + (i - L.off_wo);
  else if (i < L.off_lnw) dst = r.g.bo + (i - L.off_bo);
  else if (i < L.off_lnb) dst = r.g.lnw + (i - L.off_lnw);
  else if (i < L.off_w1) dst = r.g.lnb + (i - L.off_lnb);
  else if (i < L.off_b1) dst = r.g.w1 + (i - L.off_w1);
  else if (i < L.off_font) dst = r.g.b1 + (i - L.off_b1);
  else {
    if (r.font_grad == nullptr || i - L.off_font >= r.n_fonts * kE) return;
    dst = r.font_grad + (i - L.off_font);
  }
  *dst = s;
}

void fill_dropout(FrontArgs& a) {
  auto thr = [](double p) { return static_cast<uint32_t>(p * 65536.0 + 0.5); };
  auto inv = [](double p) { return 1.0f / static_cast<float>(1.0 - p); };
  a.thr_e = thr(a.drop.p_embed); a.thr_a = thr(a.drop.p_attn); a.thr_f = thr(a.drop.p_fc1);
  const bool on = a.drop.mode != 0;
  a.inv_e = on ? inv(a.drop.p_embed) : 1.f;
  a.inv_a = on ? inv(a.drop.p_attn) : 1.f;
  a.inv_f = on ? inv(a.drop.p_fc1) : 1.f;
}

int* g_err_flag = nullptr;  // device word: bit 0 = token id out of range
cudaError_t ensure_err_flag() {
  if (g_err_flag != nullptr) return cudaSuccess;
  cudaError_t e = cudaMalloc(&g_err_flag, sizeof(int));
  if (e != cudaSuccess) return e;
  return cudaMemset(g_err_flag, 0, sizeof(int));
}

}  // namespace

int* frontend_error_flag() { return g_err_flag; }
cudaError_t ensure_err_flag_public() { return ensure_err_flag(); }

cudaError_t read_phase_cycles(unsigned long long* host, int reset) {
#ifdef AFR_PHASE_TIMING
  cudaError_t e = cudaMemcpyFromSymbol(host, g_phase_cycles, sizeof(unsigned long long) * 32);
  if (e == cudaSuccess && reset) {
    static const unsigned long long zeros[32] = {0};
    e = cudaMemcpyToSymbol(g_phase_cycles, zeros, sizeof(zeros));
  }
  return e;
#else
  (void)host; (void)reset;
  return cudaErrorNotSupported;
#endif
}

size_t frontend_backward_smem_bytes(int L, int vocab) {   // the largest of the three kernels
  const size_t head = make_head_smem(L).total, att = make_att_smem(L).total, tail = make_tail_smem(L, vocab).total;
  return (head > tail ? (head > att ? head : att) : (tail > att ? tail : att)) * 4;
}

cudaError_t launch_frontend_forward(const Tensors& w, const long long* tokens, long long token_stride,
                                    int B, int S, int L, int vocab, const Dropout& drop,
                                    __nv_bfloat16* feats, float* state, int num_sms,
                                    cudaStream_t stream, float* feats_f32, bool shared_sm, const FontCond* font) {
  if (S < 1 || S > L || L > kMaxL) return cudaErrorInvalidValue;
  cudaError_t e = ensure_err_flag();
  if (e != cudaSuccess) return e;
  FrontArgs a{};
  a.w = w; a.tokens = tokens; a.token_stride = token_stride;
  a.B = B; a.S = S; a.L = L; a.vocab = vocab; a.drop = drop; a.feats = feats;
  a.feats_f32 = feats_f32;
  if (font != nullptr) a.font = *font;
  a.state = state; a.sl.init(L);
  a.err_flag = g_err_flag;
  fill_dropout(a);
  const size_t smem = static_cast<size_t>(make_fwd_smem(L).total) * 4;
  static size_t configured = 0;
  if (smem > configured) {
    for (auto kern : {frontend_forward_kernel, frontend_forward_kernel_shared}) {
      e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      if (e != cudaSuccess) return e;
    }
    configured = smem;
  }
  int grid = num_sms * 2;
  if (grid > B) grid = B;
  if (shared_sm) frontend_forward_kernel_shared<<<grid, kFwdThreads, smem, stream>>>(a);
  else frontend_forward_kernel<<<grid, kFwdThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_frontend_backward(const Tensors& w, const long long* tokens, long long token_stride,
                                     int B, int S, int L, int vocab, const Dropout& drop,
                                     const float* dfeat, const float* state, float* partials,
                                     int max_grid, int* grid_out, int num_sms, cudaStream_t stream,
                                     bool shared_sm, const FontCond* font) {
  if (S < 1 || S > L || L > kMaxL || state == nullptr) return cudaErrorInvalidValue;
  cudaError_t e = ensure_err_flag();
  if (e != cudaSuccess) return e;
  FrontArgs a{};
  a.w = w; a.tokens = tokens; a.token_stride = token_stride;
  a.B = B; a.S = S; a.L = L; a.vocab = vocab; a.drop = drop;
  a.dfeat = dfeat; a.partials = partials; a.lay.init(L, vocab);
  if (font != nullptr) a.font = *font;
  a.state = const_cast<float*>(state); a.sl.init(L);
  a.err_flag = g_err_flag;
  fill_dropout(a);
  const size_t smem_head = static_cast<size_t>(make_head_smem(L).total) * 4;
  const size_t smem_att = static_cast<size_t>(make_att_smem(L).total) * 4;
  const size_t smem_tail = static_cast<size_t>(make_tail_smem(L, vocab).total) * 4;
  static size_t configured[3] = {0, 0, 0};
  auto configure = [&](int which, size_t smem, std::initializer_list<const void*> kernels) -> cudaError_t {
    if (smem <= configured[which]) return cudaSuccess;
    for (const void* kern : kernels) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      if (err != cudaSuccess) return err;
      err = cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      if (err != cudaSuccess) return err;
    }
    configured[which] = smem;
    return cudaSuccess;
  };
  if ((e = configure(0, smem_head, {reinterpret_cast<const void*>(frontend_backward_head_kernel),
                                    reinterpret_cast<const void*>(frontend_backward_head_kernel_shared)})) != cudaSuccess) return e;
  if ((e = configure(1, smem_att, {reinterpret_cast<const void*>(frontend_backward_attn_kernel),
                                   reinterpret_cast<const void*>(frontend_backward_attn_kernel_shared)})) != cudaSuccess) return e;
  if ((e = configure(2, smem_tail, {reinterpret_cast<const void*>(frontend_backward_tail_kernel),
                                    reinterpret_cast<const void*>(frontend_backward_tail_kernel_shared)})) != cudaSuccess) return e;
  int grid = num_sms;
  if (grid > B) grid = B;
  if (grid > max_grid) grid = max_grid;
  *grid_out = grid;     // CTAs of the head / tail kernels = rows of `partials`
  int grid_att = 2 * num_sms;
  if (grid_att > B) grid_att = B;
  // AFR_FE_BWD_ONLY = 1 | 2 | 3 (profiling, tools/fe_contention.py): launch only that kernel
  const char* only_env = std::getenv("AFR_FE_BWD_ONLY");
  const int only = only_env != nullptr ? std::atoi(only_env) : 0;
  if (shared_sm) {
    if (only == 0 || only == 1) frontend_backward_head_kernel_shared<<<grid, kHeadThreads, smem_head, stream>>>(a);
    if (only == 0 || only == 2) frontend_backward_attn_kernel_shared<<<grid_att, kAttThreads, smem_att, stream>>>(a);
    if (only == 0 || only == 3) frontend_backward_tail_kernel_shared<<<grid, kTailThreads, smem_tail, stream>>>(a);
  } else {
    if (only == 0 || only == 1) frontend_backward_head_kernel<<<grid, kHeadThreads, smem_head, stream>>>(a);
    if (only == 0 || only == 2) frontend_backward_attn_kernel<<<grid_att, kAttThreads, smem_att, stream>>>(a);
    if (only == 0 || only == 3) frontend_backward_tail_kernel<<<grid, kTailThreads, smem_tail, stream>>>(a);
  }
  return cudaGetLastError();
}

cudaError_t launch_small_grad_reduce(const float* partials, int grid, const SmallLayout& lay,
                                     const Tensors& grads, cudaStream_t stream, float* font_grad, int n_fonts) {
  ReduceArgs r{partials, grid, lay.total, lay, grads, font_grad, n_fonts};
  small_grad_reduce_kernel<<<(lay.total + 255) / 256, 256, 0, stream>>>(r);
  return cudaGetLastError();
}

}  // namespace afr
