// Front-end of AttentionFontRenderer.forward (reference model.py:167-193) and its backward:
//   e = dropout(Emb[x]) + Pos ; a = MHA(e,e,e) ; h = LayerNorm(e + a) ;
//   f = dropout(relu(fc1(h))) ; feats = [f.reshape(B, S*64) | zeros]   (bf16, K-major GEMM operand)
// MHA follows torch.nn.functional.multi_head_attention_forward (packed in-proj, q scaled by
// sqrt(1/head_dim), softmax over all S keys with no padding mask, dropout on probabilities).
//
// This part is ~1 % of the FLOPs (2.5 MFLOP/sample), S = 100 and head_dim = 8 do not tile onto
// tensor cores, so it is fp32 SIMT: one CTA walks whole samples held in shared memory.
// The backward kernel recomputes the forward per sample (nothing but the tokens and the
// dropout seed is kept between forward and backward) and accumulates the ten small weight
// gradients into a per-CTA partial buffer that a second kernel sums in a fixed order, so the
// gradients are run-to-run deterministic.
#include "afr_internal.h"

namespace afr {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kLdW = kE + 1;  // padded leading dimension of weight matrices in smem

struct FrontArgs {
  Tensors w;
  const long long* tokens;
  long long token_stride;
  int B, S, L, vocab;
  Dropout drop;
  uint32_t thr_e, thr_a, thr_f;   // keep iff u16 >= thr
  float inv_e, inv_a, inv_f;      // 1 / (1 - p)
  __nv_bfloat16* feats;           // forward
  float* feats_f32;               // forward, optional fp32 copy (test hook)
  const float* dfeat;             // backward: [B, L*F]
  float* partials;                // backward: [grid, lay.total]
  SmallLayout lay;
  int* err_flag;
};

// ---- shared memory map (float offsets) ----------------------------------------------------
struct Smem {
  int win, bin, wo, bo, lnw, lnb, w1, b1;
  int e, q, k, v, ctx, hbuf;
  int xhat, df, dr, mrow, lrow, drow, rstd, maskbits, G;
  int total;
};
__host__ __device__ inline Smem make_smem(int L, bool bwd) {
  Smem s{};
  int o = 0;
  s.win = o; o += 3 * kE * kLdW;
  s.bin = o; o += 3 * kE;
  s.wo = o;  o += kE * kLdW;
  s.bo = o;  o += kE;
  s.lnw = o; o += kE;
  s.lnb = o; o += kE;
  s.w1 = o;  o += kF * kLdW;
  s.b1 = o;  o += kF;
  o = (o + 3) & ~3;
  s.e = o;   o += L * kE;
  s.q = o;   o += L * kE;
  s.k = o;   o += L * kE;
  s.v = o;   o += L * kE;
  s.ctx = o; o += L * kE;
  s.hbuf = o; o += kWarps * kE;
  if (bwd) {
    s.xhat = o; o += L * kE;
    s.df = o;   o += L * kF;
    s.dr = o;   o += L * kE;
    s.mrow = o; o += L * kHeads;
    s.lrow = o; o += L * kHeads;
    s.drow = o; o += L * kHeads;
    s.rstd = o; o += (L + 3) & ~3;
    s.maskbits = o; o += L * kHeads * 4;
    s.G = o;    o += 3 * kE * kE + 3 * kE + kE * kE + kE + kE + kE + kF * kE + kF;
  }
  s.total = o;
  return s;
}
// offsets inside the smem gradient accumulator G (dense, unpadded)
constexpr int kG_win = 0;
constexpr int kG_bin = kG_win + 3 * kE * kE;
constexpr int kG_wo = kG_bin + 3 * kE;
constexpr int kG_bo = kG_wo + kE * kE;
constexpr int kG_lnw = kG_bo + kE;
constexpr int kG_lnb = kG_lnw + kE;
constexpr int kG_w1 = kG_lnb + kE;
constexpr int kG_b1 = kG_w1 + kF * kE;
constexpr int kG_total = kG_b1 + kF;

// ---- counter-based dropout RNG (Philox4x32-10) ---------------------------------------------
// One call yields 8 x 16-bit uniforms for elements [8*block, 8*block+8) of one site of one
// sample at one step. Keyed by (seed); counter = (block, site, global sample index, step).
__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint32_t philox_u16(const uint4& r, int sub) {
  const uint32_t word = (sub >> 1) == 0 ? r.x : (sub >> 1) == 1 ? r.y : (sub >> 1) == 2 ? r.z : r.w;
  return (word >> ((sub & 1) * 16)) & 0xFFFFu;
}
struct Rng {
  uint32_t k0, k1, sample, step;
  __device__ __forceinline__ uint4 block(uint32_t site, uint32_t blk) const {
    return philox4x32_10(blk, site, sample, step, k0, k1);
  }
};

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// ---- weights -> smem (once per CTA) --------------------------------------------------------
__device__ void load_weights(const Tensors& w, float* sm, const Smem& o) {
  for (int i = threadIdx.x; i < 3 * kE * kE; i += kThreads)
    sm[o.win + (i / kE) * kLdW + (i % kE)] = w.win[i];
  for (int i = threadIdx.x; i < kE * kE; i += kThreads)
    sm[o.wo + (i / kE) * kLdW + (i % kE)] = w.wo[i];
  for (int i = threadIdx.x; i < kF * kE; i += kThreads)
    sm[o.w1 + (i / kE) * kLdW + (i % kE)] = w.w1[i];
  for (int i = threadIdx.x; i < 3 * kE; i += kThreads) sm[o.bin + i] = w.bin[i];
  for (int i = threadIdx.x; i < kF; i += kThreads) sm[o.b1 + i] = w.b1[i];
  if (threadIdx.x < kE) {
    sm[o.bo + threadIdx.x] = w.bo[threadIdx.x];
    sm[o.lnw + threadIdx.x] = w.lnw[threadIdx.x];
    sm[o.lnb + threadIdx.x] = w.lnb[threadIdx.x];
  }
}

// ---- forward of one sample up to the LayerNorm output --------------------------------------
// Leaves in smem: e, q (pre-scaled), k, v, ctx and, when BWD, xhat/rstd/softmax stats/mask bits.
// Returns through the callback-free convention: the caller runs the fc1 stage itself because
// forward and backward consume h differently.
template <bool BWD>
__device__ void forward_to_ctx(const FrontArgs& a, float* sm, const Smem& o, int b, const Rng& rng) {
  const int S = a.S;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long* tok = a.tokens + static_cast<long long>(b) * a.token_stride;

  // (1) e = dropout(Emb[tok]) + Pos           model.py:167-172 (dropout BEFORE positions)
  for (int i = tid; i < S * kE; i += kThreads) {
    const int s = i / kE, c = i % kE;
    long long t = tok[s];
    if (t < 0 || t >= a.vocab) { atomicOr(a.err_flag, 1); t = 0; }
    float val = a.w.emb[t * kE + c];
    if (a.drop.mode == 1) {
      const uint4 r = rng.block(0u, static_cast<uint32_t>(i >> 3));
      val = philox_u16(r, i & 7) >= a.thr_e ? val * a.inv_e : 0.f;
    } else if (a.drop.mode == 2) {
      val = a.drop.mask_embed[(static_cast<long long>(b) * S + s) * kE + c] ? val * a.inv_e : 0.f;
    }
    sm[o.e + i] = val + a.w.pos[i];
  }
  __syncthreads();

  // (2) packed in-projection: q|k|v = e Win^T + bin ; q *= sqrt(1/head_dim)
  {
    float wr[3][kE];
#pragma unroll
    for (int j = 0; j < 3; ++j)
#pragma unroll
      for (int c = 0; c < kE; ++c) wr[j][c] = sm[o.win + (lane + 32 * j) * kLdW + c];
    const float bq = sm[o.bin + lane], bk = sm[o.bin + 32 + lane], bv = sm[o.bin + 64 + lane];
    const float qscale = 0.35355339059327373f;  // math.sqrt(1.0 / 8)
    for (int s = warp; s < S; s += kWarps) {
      float aq = bq, ak = bk, av = bv;
      const float4* er = reinterpret_cast<const float4*>(sm + o.e + s * kE);
#pragma unroll
      for (int c4 = 0; c4 < kE / 4; ++c4) {
        const float4 x = er[c4];
        aq = fmaf(x.x, wr[0][4 * c4], aq); aq = fmaf(x.y, wr[0][4 * c4 + 1], aq);
        aq = fmaf(x.z, wr[0][4 * c4 + 2], aq); aq = fmaf(x.w, wr[0][4 * c4 + 3], aq);
        ak = fmaf(x.x, wr[1][4 * c4], ak); ak = fmaf(x.y, wr[1][4 * c4 + 1], ak);
        ak = fmaf(x.z, wr[1][4 * c4 + 2], ak); ak = fmaf(x.w, wr[1][4 * c4 + 3], ak);
        av = fmaf(x.x, wr[2][4 * c4], av); av = fmaf(x.y, wr[2][4 * c4 + 1], av);
        av = fmaf(x.z, wr[2][4 * c4 + 2], av); av = fmaf(x.w, wr[2][4 * c4 + 3], av);
      }
      sm[o.q + s * kE + lane] = aq * qscale;
      sm[o.k + s * kE + lane] = ak;
      sm[o.v + s * kE + lane] = av;
    }
  }
  __syncthreads();

  // (3) per (query s, head h): softmax(q k^T) [dropout] v
  for (int i = tid; i < S * kHeads; i += kThreads) {
    const int s = i / kHeads, h = i % kHeads;
    float qr[kDh];
    {
      const float4* qp = reinterpret_cast<const float4*>(sm + o.q + s * kE + h * kDh);
      const float4 q0 = qp[0], q1 = qp[1];
      qr[0] = q0.x; qr[1] = q0.y; qr[2] = q0.z; qr[3] = q0.w;
      qr[4] = q1.x; qr[5] = q1.y; qr[6] = q1.z; qr[7] = q1.w;
    }
    float mx = -INFINITY;
    for (int t = 0; t < S; ++t) {
      const float4* kp = reinterpret_cast<const float4*>(sm + o.k + t * kE + h * kDh);
      const float4 k0 = kp[0], k1 = kp[1];
      float sc = qr[0] * k0.x;
      sc = fmaf(qr[1], k0.y, sc); sc = fmaf(qr[2], k0.z, sc); sc = fmaf(qr[3], k0.w, sc);
      sc = fmaf(qr[4], k1.x, sc); sc = fmaf(qr[5], k1.y, sc); sc = fmaf(qr[6], k1.z, sc);
      sc = fmaf(qr[7], k1.w, sc);
      mx = fmaxf(mx, sc);
    }
    float l = 0.f, acc[kDh];
#pragma unroll
    for (int j = 0; j < kDh; ++j) acc[j] = 0.f;
    uint32_t bits = 0;
    uint4 rnd = make_uint4(0, 0, 0, 0);
    const uint32_t row_elem0 = static_cast<uint32_t>((h * S + s) * S);
    for (int t = 0; t < S; ++t) {
      const float4* kp = reinterpret_cast<const float4*>(sm + o.k + t * kE + h * kDh);
      const float4 k0 = kp[0], k1 = kp[1];
      float sc = qr[0] * k0.x;
      sc = fmaf(qr[1], k0.y, sc); sc = fmaf(qr[2], k0.z, sc); sc = fmaf(qr[3], k0.w, sc);
      sc = fmaf(qr[4], k1.x, sc); sc = fmaf(qr[5], k1.y, sc); sc = fmaf(qr[6], k1.z, sc);
      sc = fmaf(qr[7], k1.w, sc);
      const float p = expf(sc - mx);
      l += p;
      bool keep = true;
      if (a.drop.mode == 1) {
        const uint32_t el = row_elem0 + static_cast<uint32_t>(t);
        if ((el & 7u) == 0u || t == 0) rnd = rng.block(1u, el >> 3);
        keep = philox_u16(rnd, el & 7u) >= a.thr_a;
      } else if (a.drop.mode == 2) {
        keep = a.drop.mask_attn[((static_cast<long long>(b) * kHeads + h) * S + s) * S + t] != 0;
      }
      if (BWD) {
        if (keep) bits |= 1u << (t & 31);
        if ((t & 31) == 31 || t == S - 1) {
          reinterpret_cast<uint32_t*>(sm + o.maskbits)[(h * S + s) * 4 + (t >> 5)] = bits;
          bits = 0;
        }
      }
      if (keep) {
        const float4* vp = reinterpret_cast<const float4*>(sm + o.v + t * kE + h * kDh);
        const float4 v0 = vp[0], v1 = vp[1];
        acc[0] = fmaf(p, v0.x, acc[0]); acc[1] = fmaf(p, v0.y, acc[1]);
        acc[2] = fmaf(p, v0.z, acc[2]); acc[3] = fmaf(p, v0.w, acc[3]);
        acc[4] = fmaf(p, v1.x, acc[4]); acc[5] = fmaf(p, v1.y, acc[5]);
        acc[6] = fmaf(p, v1.z, acc[6]); acc[7] = fmaf(p, v1.w, acc[7]);
      }
    }
    const float scale = (a.drop.mode != 0 ? a.inv_a : 1.f) / l;
#pragma unroll
    for (int j = 0; j < kDh; ++j) sm[o.ctx + s * kE + h * kDh + j] = acc[j] * scale;
    if (BWD) { sm[o.mrow + i] = mx; sm[o.lrow + i] = l; }
  }
  __syncthreads();
}

// out-projection + residual + LayerNorm for position s (warp-wide, lane = channel).
// Returns h[s][lane]; optionally xhat and rstd.
__device__ __forceinline__ float attn_out_layernorm(const float* sm, const Smem& o, int s, int lane,
                                                    const float (&wo_row)[kE], float& xhat,
                                                    float& rstd) {
  float acc = sm[o.bo + lane];
  const float4* cr = reinterpret_cast<const float4*>(sm + o.ctx + s * kE);
#pragma unroll
  for (int j4 = 0; j4 < kE / 4; ++j4) {
    const float4 x = cr[j4];
    acc = fmaf(x.x, wo_row[4 * j4], acc); acc = fmaf(x.y, wo_row[4 * j4 + 1], acc);
    acc = fmaf(x.z, wo_row[4 * j4 + 2], acc); acc = fmaf(x.w, wo_row[4 * j4 + 3], acc);
  }
  const float r = sm[o.e + s * kE + lane] + acc;        // residual, model.py:180
  const float mean = warp_sum(r) * (1.f / kE);
  const float d = r - mean;
  const float var = warp_sum(d * d) * (1.f / kE);       // biased variance, eps = 1e-5
  rstd = 1.f / sqrtf(var + 1e-5f);
  xhat = d * rstd;
  return fmaf(xhat, sm[o.lnw + lane], sm[o.lnb + lane]);
}

// ============================================================================ forward kernel
__global__ void __launch_bounds__(kThreads, 2) frontend_forward_kernel(FrontArgs a) {
  extern __shared__ __align__(16) float sm[];
  const Smem o = make_smem(a.L, false);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  load_weights(a.w, sm, o);
  __syncthreads();

  const int S = a.S, KF = a.L * kF;

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    Rng rng{static_cast<uint32_t>(a.drop.seed), static_cast<uint32_t>(a.drop.seed >> 32),
            static_cast<uint32_t>(a.drop.sample_offset + b), static_cast<uint32_t>(a.drop.step)};
    forward_to_ctx<false>(a, sm, o, b, rng);
    float wo_row[kE], w1a[kE], w1b[kE];
#pragma unroll
    for (int c = 0; c < kE; ++c) {
      wo_row[c] = sm[o.wo + lane * kLdW + c];
      w1a[c] = sm[o.w1 + lane * kLdW + c];
      w1b[c] = sm[o.w1 + (lane + 32) * kLdW + c];
    }
    const float b1a = sm[o.b1 + lane], b1b = sm[o.b1 + 32 + lane];
    __nv_bfloat16* out = a.feats + static_cast<long long>(b) * KF;
    for (int s = warp; s < S; s += kWarps) {
      float xhat, rstd;
      const float hval = attn_out_layernorm(sm, o, s, lane, wo_row, xhat, rstd);
      sm[o.hbuf + warp * kE + lane] = hval;
      __syncwarp();
      float fa = b1a, fb = b1b;
      const float4* hr = reinterpret_cast<const float4*>(sm + o.hbuf + warp * kE);
#pragma unroll
      for (int c4 = 0; c4 < kE / 4; ++c4) {
        const float4 x = hr[c4];
        fa = fmaf(x.x, w1a[4 * c4], fa); fa = fmaf(x.y, w1a[4 * c4 + 1], fa);
        fa = fmaf(x.z, w1a[4 * c4 + 2], fa); fa = fmaf(x.w, w1a[4 * c4 + 3], fa);
        fb = fmaf(x.x, w1b[4 * c4], fb); fb = fmaf(x.y, w1b[4 * c4 + 1], fb);
        fb = fmaf(x.z, w1b[4 * c4 + 2], fb); fb = fmaf(x.w, w1b[4 * c4 + 3], fb);
      }
      __syncwarp();
      fa = fmaxf(fa, 0.f); fb = fmaxf(fb, 0.f);          // ReLU, model.py:183
      if (a.drop.mode == 1) {                            // dropout1, model.py:184
        const uint32_t ea = static_cast<uint32_t>(s * kF + lane), eb = ea + 32u;
        fa = philox_u16(rng.block(2u, ea >> 3), ea & 7u) >= a.thr_f ? fa * a.inv_f : 0.f;
        fb = philox_u16(rng.block(2u, eb >> 3), eb & 7u) >= a.thr_f ? fb * a.inv_f : 0.f;
      } else if (a.drop.mode == 2) {
        const uint8_t* mk = a.drop.mask_fc1 + (static_cast<long long>(b) * S + s) * kF;
        fa = mk[lane] ? fa * a.inv_f : 0.f;
        fb = mk[lane + 32] ? fb * a.inv_f : 0.f;
      }
      out[s * kF + lane] = __float2bfloat16_rn(fa);
      out[s * kF + 32 + lane] = __float2bfloat16_rn(fb);
      if (a.feats_f32 != nullptr) {
        a.feats_f32[static_cast<long long>(b) * KF + s * kF + lane] = fa;
        a.feats_f32[static_cast<long long>(b) * KF + s * kF + 32 + lane] = fb;
      }
    }
    // zero features for positions >= S (model.py:190-193)
    for (int i = S * kF + tid; i < KF; i += kThreads) {
      out[i] = __float2bfloat16_rn(0.f);
      if (a.feats_f32 != nullptr) a.feats_f32[static_cast<long long>(b) * KF + i] = 0.f;
    }
    __syncthreads();
  }
}

// =========================================================================== backward kernel
__global__ void __launch_bounds__(kThreads) frontend_backward_kernel(FrontArgs a) {
  extern __shared__ __align__(16) float sm[];
  const Smem o = make_smem(a.L, true);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int S = a.S, KF = a.L * kF;
  load_weights(a.w, sm, o);
  for (int i = tid; i < kG_total; i += kThreads) sm[o.G + i] = 0.f;
  float* part = a.partials + static_cast<long long>(blockIdx.x) * a.lay.total;
  for (int i = tid; i < a.L * kE; i += kThreads) part[a.lay.off_pos + i] = 0.f;
  for (int i = tid; i < a.vocab * kE; i += kThreads) part[a.lay.off_emb + i] = 0.f;
  __syncthreads();

  float dgamma = 0.f, dbeta = 0.f;  // per (warp, lane = channel), reduced over warps at the end

  for (int b = blockIdx.x; b < a.B; b += gridDim.x) {
    Rng rng{static_cast<uint32_t>(a.drop.seed), static_cast<uint32_t>(a.drop.seed >> 32),
            static_cast<uint32_t>(a.drop.sample_offset + b), static_cast<uint32_t>(a.drop.step)};
    forward_to_ctx<true>(a, sm, o, b, rng);
    const float* dfe = a.dfeat + static_cast<long long>(b) * KF;

    // ---- B1: LayerNorm output, fc1 recompute, d(fc1), d(LayerNorm) ------------------------
    {
      float wo_row[kE], w1a[kE], w1b[kE];
#pragma unroll
      for (int c = 0; c < kE; ++c) {
        wo_row[c] = sm[o.wo + lane * kLdW + c];
        w1a[c] = sm[o.w1 + lane * kLdW + c];
        w1b[c] = sm[o.w1 + (lane + 32) * kLdW + c];
      }
      const float b1a = sm[o.b1 + lane], b1b = sm[o.b1 + 32 + lane];
      const float gam = sm[o.lnw + lane];
      for (int s = warp; s < S; s += kWarps) {
        float xhat, rstd;
        const float hval = attn_out_layernorm(sm, o, s, lane, wo_row, xhat, rstd);
        sm[o.xhat + s * kE + lane] = xhat;
        if (lane == 0) sm[o.rstd + s] = rstd;
        sm[o.hbuf + warp * kE + lane] = hval;
        __syncwarp();
        float fa = b1a, fb = b1b;
        const float4* hr = reinterpret_cast<const float4*>(sm + o.hbuf + warp * kE);
#pragma unroll
        for (int c4 = 0; c4 < kE / 4; ++c4) {
          const float4 x = hr[c4];
          fa = fmaf(x.x, w1a[4 * c4], fa); fa = fmaf(x.y, w1a[4 * c4 + 1], fa);
          fa = fmaf(x.z, w1a[4 * c4 + 2], fa); fa = fmaf(x.w, w1a[4 * c4 + 3], fa);
          fb = fmaf(x.x, w1b[4 * c4], fb); fb = fmaf(x.y, w1b[4 * c4 + 1], fb);
          fb = fmaf(x.z, w1b[4 * c4 + 2], fb); fb = fmaf(x.w, w1b[4 * c4 + 3], fb);
        }
        bool ka = true, kb = true;
        if (a.drop.mode == 1) {
          const uint32_t ea = static_cast<uint32_t>(s * kF + lane), eb = ea + 32u;
          ka = philox_u16(rng.block(2u, ea >> 3), ea & 7u) >= a.thr_f;
          kb = philox_u16(rng.block(2u, eb >> 3), eb & 7u) >= a.thr_f;
        } else if (a.drop.mode == 2) {
          const uint8_t* mk = a.drop.mask_fc1 + (static_cast<long long>(b) * S + s) * kF;
          ka = mk[lane] != 0; kb = mk[lane + 32] != 0;
        }
        const float sc = a.drop.mode != 0 ? a.inv_f : 1.f;
        const float dfa = (fa > 0.f && ka) ? dfe[s * kF + lane] * sc : 0.f;
        const float dfb = (fb > 0.f && kb) ? dfe[s * kF + 32 + lane] * sc : 0.f;
        sm[o.df + s * kF + lane] = dfa;
        sm[o.df + s * kF + 32 + lane] = dfb;
        __syncwarp();
        // dh[c] = sum_j df[j] W1[j][c]     (lane = c)
        float dh = 0.f;
        const float* dfr = sm + o.df + s * kF;
#pragma unroll 8
        for (int j = 0; j < kF; ++j) dh = fmaf(dfr[j], sm[o.w1 + j * kLdW + lane], dh);
        dgamma = fmaf(dh, xhat, dgamma);
        dbeta += dh;
        const float dhg = dh * gam;
        const float m1 = warp_sum(dhg) * (1.f / kE);
        const float m2 = warp_sum(dhg * xhat) * (1.f / kE);
        sm[o.dr + s * kE + lane] = rstd * (dhg - m1 - xhat * m2);
        __syncwarp();
      }
    }
    __syncthreads();

    // ---- B1w: dW1 += df^T h ; db1 += sum_s df -------------------------------------------
    {
      const int j = tid >> 2, c0 = (tid & 3) * 8;
      float acc[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) acc[u] = 0.f;
      float g8[8], bt8[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) { g8[u] = sm[o.lnw + c0 + u]; bt8[u] = sm[o.lnb + c0 + u]; }
      for (int s = 0; s < S; ++s) {
        const float d = sm[o.df + s * kF + j];
        const float4* xr = reinterpret_cast<const float4*>(sm + o.xhat + s * kE + c0);
        const float4 x0 = xr[0], x1 = xr[1];
        const float xs[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = fmaf(d, fmaf(xs[u], g8[u], bt8[u]), acc[u]);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) sm[o.G + kG_w1 + j * kE + c0 + u] += acc[u];
      if (tid < kF) {
        float sacc = 0.f;
        for (int s = 0; s < S; ++s) sacc += sm[o.df + s * kF + tid];
        sm[o.G + kG_b1 + tid] += sacc;
      }
    }
    __syncthreads();

    // ---- B2: dctx = dr Wo (into df[0 : S*E]) ; D = dctx . ctx ; dWo, dbo ------------------
    float* dctx = sm + o.df;
    for (int s = warp; s < S; s += kWarps) {
      float acc = 0.f;
      const float* drr = sm + o.dr + s * kE;
#pragma unroll 8
      for (int c = 0; c < kE; ++c) acc = fmaf(drr[c], sm[o.wo + c * kLdW + lane], acc);
      dctx[s * kE + lane] = acc;
      // D[s][h] = sum_{j in head h} dctx[s][j] * ctx[s][j]   (= sum_t P_dropped dP)
      float prod = acc * sm[o.ctx + s * kE + lane];
      prod += __shfl_xor_sync(0xffffffffu, prod, 1);
      prod += __shfl_xor_sync(0xffffffffu, prod, 2);
      prod += __shfl_xor_sync(0xffffffffu, prod, 4);
      if ((lane & 7) == 0) sm[o.drow + s * kHeads + (lane >> 3)] = prod;
    }
    {
      const int c = tid >> 3, j0 = (tid & 7) * 4;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
      for (int s = 0; s < S; ++s) {
        const float d = sm[o.dr + s * kE + c];
        const float4 x = *reinterpret_cast<const float4*>(sm + o.ctx + s * kE + j0);
        acc[0] = fmaf(d, x.x, acc[0]); acc[1] = fmaf(d, x.y, acc[1]);
        acc[2] = fmaf(d, x.z, acc[2]); acc[3] = fmaf(d, x.w, acc[3]);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) sm[o.G + kG_wo + c * kE + j0 + u] += acc[u];
      if (tid < kE) {
        float sacc = 0.f;
        for (int s = 0; s < S; ++s) sacc += sm[o.dr + s * kE + tid];
        sm[o.G + kG_bo + tid] += sacc;
      }
    }
    __syncthreads();

    // ---- B3: attention backward. dq -> xhat buffer, dk -> ctx buffer, dv -> df[S*E : 2*S*E]
    float* dq = sm + o.xhat;
    float* dk = sm + o.ctx;
    float* dv = sm + o.df + a.L * kE;
    const float inv_a = a.drop.mode != 0 ? a.inv_a : 1.f;
    const uint32_t* mbits = reinterpret_cast<const uint32_t*>(sm + o.maskbits);
    // pass A: per (query s, head h) -> dq
    float dq_out[2][kDh];
#pragma unroll
    for (int npairs = 0; npairs < 2; ++npairs) {
      const int i = tid + npairs * kThreads;
      if (i >= S * kHeads) break;
      const int s = i / kHeads, h = i % kHeads;
      float qr[kDh], dc[kDh], acc[kDh];
#pragma unroll
      for (int j = 0; j < kDh; ++j) {
        qr[j] = sm[o.q + s * kE + h * kDh + j];
        dc[j] = dctx[s * kE + h * kDh + j];
        acc[j] = 0.f;
      }
      const float mx = sm[o.mrow + i], linv = 1.f / sm[o.lrow + i], D = sm[o.drow + i];
      for (int t = 0; t < S; ++t) {
        const float4* kp = reinterpret_cast<const float4*>(sm + o.k + t * kE + h * kDh);
        const float4* vp = reinterpret_cast<const float4*>(sm + o.v + t * kE + h * kDh);
        const float4 k0 = kp[0], k1 = kp[1], v0 = vp[0], v1 = vp[1];
        float sc = qr[0] * k0.x;
        sc = fmaf(qr[1], k0.y, sc); sc = fmaf(qr[2], k0.z, sc); sc = fmaf(qr[3], k0.w, sc);
        sc = fmaf(qr[4], k1.x, sc); sc = fmaf(qr[5], k1.y, sc); sc = fmaf(qr[6], k1.z, sc);
        sc = fmaf(qr[7], k1.w, sc);
        const float p = expf(sc - mx) * linv;
        float dp = dc[0] * v0.x;
        dp = fmaf(dc[1], v0.y, dp); dp = fmaf(dc[2], v0.z, dp); dp = fmaf(dc[3], v0.w, dp);
        dp = fmaf(dc[4], v1.x, dp); dp = fmaf(dc[5], v1.y, dp); dp = fmaf(dc[6], v1.z, dp);
        dp = fmaf(dc[7], v1.w, dp);
        const bool keep = (mbits[(h * S + s) * 4 + (t >> 5)] >> (t & 31)) & 1u;
        const float ds = p * ((keep ? dp * inv_a : 0.f) - D);
        acc[0] = fmaf(ds, k0.x, acc[0]); acc[1] = fmaf(ds, k0.y, acc[1]);
        acc[2] = fmaf(ds, k0.z, acc[2]); acc[3] = fmaf(ds, k0.w, acc[3]);
        acc[4] = fmaf(ds, k1.x, acc[4]); acc[5] = fmaf(ds, k1.y, acc[5]);
        acc[6] = fmaf(ds, k1.z, acc[6]); acc[7] = fmaf(ds, k1.w, acc[7]);
      }
#pragma unroll
      for (int j = 0; j < kDh; ++j) dq_out[npairs][j] = acc[j] * 0.35355339059327373f;
    }
    // pass B: per (key t, head h) -> dk, dv   (reads q, k, v, dctx; writes after the barrier)
    float dk_out[2][kDh], dv_out[2][kDh];
#pragma unroll
    for (int npairs = 0; npairs < 2; ++npairs) {
      const int i = tid + npairs * kThreads;
      if (i >= S * kHeads) break;
      const int t = i / kHeads, h = i % kHeads;
      float kr[kDh], vr[kDh], ak[kDh], av[kDh];
#pragma unroll
      for (int j = 0; j < kDh; ++j) {
        kr[j] = sm[o.k + t * kE + h * kDh + j];
        vr[j] = sm[o.v + t * kE + h * kDh + j];
        ak[j] = 0.f; av[j] = 0.f;
      }
      for (int s = 0; s < S; ++s) {
        const float4* qp = reinterpret_cast<const float4*>(sm + o.q + s * kE + h * kDh);
        const float4* cp = reinterpret_cast<const float4*>(dctx + s * kE + h * kDh);
        const float4 q0 = qp[0], q1 = qp[1], c0 = cp[0], c1 = cp[1];
        float sc = q0.x * kr[0];
        sc = fmaf(q0.y, kr[1], sc); sc = fmaf(q0.z, kr[2], sc); sc = fmaf(q0.w, kr[3], sc);
        sc = fmaf(q1.x, kr[4], sc); sc = fmaf(q1.y, kr[5], sc); sc = fmaf(q1.z, kr[6], sc);
        sc = fmaf(q1.w, kr[7], sc);
        const int sh = s * kHeads + h;
        const float p = expf(sc - sm[o.mrow + sh]) / sm[o.lrow + sh];
        float dp = c0.x * vr[0];
        dp = fmaf(c0.y, vr[1], dp); dp = fmaf(c0.z, vr[2], dp); dp = fmaf(c0.w, vr[3], dp);
        dp = fmaf(c1.x, vr[4], dp); dp = fmaf(c1.y, vr[5], dp); dp = fmaf(c1.z, vr[6], dp);
        dp = fmaf(c1.w, vr[7], dp);
        const bool keep = (mbits[(h * S + s) * 4 + (t >> 5)] >> (t & 31)) & 1u;
        const float pd = keep ? p * inv_a : 0.f;
        const float ds = p * ((keep ? dp * inv_a : 0.f) - sm[o.drow + sh]);
        av[0] = fmaf(pd, c0.x, av[0]); av[1] = fmaf(pd, c0.y, av[1]);
        av[2] = fmaf(pd, c0.z, av[2]); av[3] = fmaf(pd, c0.w, av[3]);
        av[4] = fmaf(pd, c1.x, av[4]); av[5] = fmaf(pd, c1.y, av[5]);
        av[6] = fmaf(pd, c1.z, av[6]); av[7] = fmaf(pd, c1.w, av[7]);
        ak[0] = fmaf(ds, q0.x, ak[0]); ak[1] = fmaf(ds, q0.y, ak[1]);
        ak[2] = fmaf(ds, q0.z, ak[2]); ak[3] = fmaf(ds, q0.w, ak[3]);
        ak[4] = fmaf(ds, q1.x, ak[4]); ak[5] = fmaf(ds, q1.y, ak[5]);
        ak[6] = fmaf(ds, q1.z, ak[6]); ak[7] = fmaf(ds, q1.w, ak[7]);
      }
#pragma unroll
      for (int j = 0; j < kDh; ++j) { dk_out[npairs][j] = ak[j]; dv_out[npairs][j] = av[j]; }
    }
    __syncthreads();  // every reader of ctx / xhat / dctx is done: now overwrite the aliases
#pragma unroll
    for (int npairs = 0; npairs < 2; ++npairs) {
      const int i = tid + npairs * kThreads;
      if (i >= S * kHeads) break;
      const int s = i / kHeads, h = i % kHeads;
#pragma unroll
      for (int j = 0; j < kDh; ++j) {
        dq[s * kE + h * kDh + j] = dq_out[npairs][j];
        dk[s * kE + h * kDh + j] = dk_out[npairs][j];
        dv[s * kE + h * kDh + j] = dv_out[npairs][j];
      }
    }
    __syncthreads();

    // ---- B4: de = dr + dqkv Win ; dPos, dEmb ; dWin, dbin --------------------------------
    for (int s = warp; s < S; s += kWarps) {
      float acc = sm[o.dr + s * kE + lane];
      const float* r0 = dq + s * kE;
      const float* r1 = dk + s * kE;
      const float* r2 = dv + s * kE;
#pragma unroll 8
      for (int oo = 0; oo < kE; ++oo) {
        acc = fmaf(r0[oo], sm[o.win + oo * kLdW + lane], acc);
        acc = fmaf(r1[oo], sm[o.win + (32 + oo) * kLdW + lane], acc);
        acc = fmaf(r2[oo], sm[o.win + (64 + oo) * kLdW + lane], acc);
      }
      // the same (warp, lane) owns (s, c) for every sample this CTA processes: plain RMW
      float* pp = part + a.lay.off_pos + s * kE + lane;
      __stcg(pp, __ldcg(pp) + acc);
      // gradient wrt the embedding row: through the embedding dropout (scale or zero)
      float de = acc;
      const int i = s * kE + lane;
      if (a.drop.mode == 1) {
        const uint4 r = rng.block(0u, static_cast<uint32_t>(i >> 3));
        de = philox_u16(r, i & 7) >= a.thr_e ? de * a.inv_e : 0.f;
      } else if (a.drop.mode == 2) {
        de = a.drop.mask_embed[(static_cast<long long>(b) * S + s) * kE + lane] ? de * a.inv_e : 0.f;
      }
      sm[o.dr + s * kE + lane] = de;   // dr is dead from here on: reuse it as d(embedding rows)
    }
    __syncthreads();
    {
      // dWin[o][c] += sum_s dqkv[s][o] e[s][c]: 96 x 4 (o, c-group) pairs of 8 outputs each
      for (int idx = tid; idx < 3 * kE * 4; idx += kThreads) {
        const int oo = idx >> 2, c0 = (idx & 3) * 8;
        const float* src = oo < 32 ? dq : (oo < 64 ? dk : dv);
        const int col = oo & 31;
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.f;
        for (int s = 0; s < S; ++s) {
          const float d = src[s * kE + col];
          const float4* er = reinterpret_cast<const float4*>(sm + o.e + s * kE + c0);
          const float4 x0 = er[0], x1 = er[1];
          acc[0] = fmaf(d, x0.x, acc[0]); acc[1] = fmaf(d, x0.y, acc[1]);
          acc[2] = fmaf(d, x0.z, acc[2]); acc[3] = fmaf(d, x0.w, acc[3]);
          acc[4] = fmaf(d, x1.x, acc[4]); acc[5] = fmaf(d, x1.y, acc[5]);
          acc[6] = fmaf(d, x1.z, acc[6]); acc[7] = fmaf(d, x1.w, acc[7]);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) sm[o.G + kG_win + oo * kE + c0 + u] += acc[u];
      }
      if (tid < 3 * kE) {
        const float* src = tid < 32 ? dq : (tid < 64 ? dk : dv);
        float sacc = 0.f;
        for (int s = 0; s < S; ++s) sacc += src[s * kE + (tid & 31)];
        sm[o.G + kG_bin + tid] += sacc;
      }
      // dEmb: scatter-add over token ids. One warp walks the positions in order (lane = c), so
      // rows hit by several positions are summed in a fixed order (deterministic, no atomics).
      if (warp == kWarps - 1) {
        const long long* tok = a.tokens + static_cast<long long>(b) * a.token_stride;
        for (int s = 0; s < S; ++s) {
          long long t = tok[s];
          if (t < 0 || t >= a.vocab) t = 0;
          float* pe = part + a.lay.off_emb + t * kE + lane;
          __stcg(pe, __ldcg(pe) + sm[o.dr + s * kE + lane]);
        }
      }
    }
    __syncthreads();
  }

  // ---- flush this CTA's partial sums ---------------------------------------------------
  sm[o.hbuf + warp * kE + lane] = dgamma;
  __syncthreads();
  if (tid < kE) {
    float sacc = 0.f;
    for (int w = 0; w < kWarps; ++w) sacc += sm[o.hbuf + w * kE + tid];
    sm[o.G + kG_lnw + tid] = sacc;
  }
  __syncthreads();
  sm[o.hbuf + warp * kE + lane] = dbeta;
  __syncthreads();
  if (tid < kE) {
    float sacc = 0.f;
    for (int w = 0; w < kWarps; ++w) sacc += sm[o.hbuf + w * kE + tid];
    sm[o.G + kG_lnb + tid] = sacc;
  }
  __syncthreads();
  for (int i = tid; i < 3 * kE * kE; i += kThreads) part[a.lay.off_win + i] = sm[o.G + kG_win + i];
  for (int i = tid; i < 3 * kE; i += kThreads) part[a.lay.off_bin + i] = sm[o.G + kG_bin + i];
  for (int i = tid; i < kE * kE; i += kThreads) part[a.lay.off_wo + i] = sm[o.G + kG_wo + i];
  for (int i = tid; i < kF * kE; i += kThreads) part[a.lay.off_w1 + i] = sm[o.G + kG_w1 + i];
  for (int i = tid; i < kF; i += kThreads) part[a.lay.off_b1 + i] = sm[o.G + kG_b1 + i];
  if (tid < kE) {
    part[a.lay.off_bo + tid] = sm[o.G + kG_bo + tid];
    part[a.lay.off_lnw + tid] = sm[o.G + kG_lnw + tid];
    part[a.lay.off_lnb + tid] = sm[o.G + kG_lnb + tid];
  }
}

// grads[i] = sum over CTAs of partials[cta][i], fixed order.
struct ReduceArgs {
  const float* partials;
  int grid, total;
  SmallLayout lay;
  Tensors g;
};
__global__ void __launch_bounds__(256) small_grad_reduce_kernel(ReduceArgs r) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r.total) return;
  float s = 0.f;
  for (int c = 0; c < r.grid; ++c) s += r.partials[static_cast<long long>(c) * r.total + i];
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
  else dst = r.g.b1 + (i - L.off_b1);
  *dst = s;
}

void fill_dropout(FrontArgs& a) {
  auto thr = [](double p) { return static_cast<uint32_t>(p * 65536.0 + 0.5); };
  auto inv = [](double p) { return 1.0f / static_cast<float>(1.0 - p); };
  a.thr_e = thr(a.drop.p_embed); a.thr_a = thr(a.drop.p_attn); a.thr_f = thr(a.drop.p_fc1);
  a.inv_e = inv(a.drop.p_embed); a.inv_a = inv(a.drop.p_attn); a.inv_f = inv(a.drop.p_fc1);
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

size_t frontend_backward_smem_bytes(int L) { return static_cast<size_t>(make_smem(L, true).total) * 4; }

cudaError_t launch_frontend_forward(const Tensors& w, const long long* tokens, long long token_stride,
                                    int B, int S, int L, int vocab, const Dropout& drop,
                                    __nv_bfloat16* feats, int num_sms, cudaStream_t stream,
                                    float* feats_f32) {
  if (S < 1 || S > L || L > kMaxL) return cudaErrorInvalidValue;
  cudaError_t e = ensure_err_flag();
  if (e != cudaSuccess) return e;
  FrontArgs a{};
  a.w = w; a.tokens = tokens; a.token_stride = token_stride;
  a.B = B; a.S = S; a.L = L; a.vocab = vocab; a.drop = drop; a.feats = feats;
  a.feats_f32 = feats_f32;
  a.err_flag = g_err_flag;
  fill_dropout(a);
  const size_t smem = static_cast<size_t>(make_smem(L, false).total) * 4;
  static size_t configured = 0;
  if (smem > configured) {
    e = cudaFuncSetAttribute(frontend_forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  int grid = num_sms * 2;
  if (grid > B) grid = B;
  frontend_forward_kernel<<<grid, kThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_frontend_backward(const Tensors& w, const long long* tokens, long long token_stride,
                                     int B, int S, int L, int vocab, const Dropout& drop,
                                     const float* dfeat, float* partials, int max_grid,
                                     int* grid_out, int num_sms, cudaStream_t stream) {
  if (S < 1 || S > L || L > kMaxL) return cudaErrorInvalidValue;
  cudaError_t e = ensure_err_flag();
  if (e != cudaSuccess) return e;
  FrontArgs a{};
  a.w = w; a.tokens = tokens; a.token_stride = token_stride;
  a.B = B; a.S = S; a.L = L; a.vocab = vocab; a.drop = drop;
  a.dfeat = dfeat; a.partials = partials; a.lay.init(L, vocab);
  a.err_flag = g_err_flag;
  fill_dropout(a);
  const size_t smem = frontend_backward_smem_bytes(L);
  static size_t configured = 0;
  if (smem > configured) {
    e = cudaFuncSetAttribute(frontend_backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  int grid = num_sms;
  if (grid > B) grid = B;
  if (grid > max_grid) grid = max_grid;
  *grid_out = grid;
  frontend_backward_kernel<<<grid, kThreads, smem, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_small_grad_reduce(const float* partials, int grid, const SmallLayout& lay,
                                     const Tensors& grads, cudaStream_t stream) {
  ReduceArgs r{partials, grid, lay.total, lay, grads};
  small_grad_reduce_kernel<<<(lay.total + 255) / 256, 256, 0, stream>>>(r);
  return cudaGetLastError();
}

}  // namespace afr
