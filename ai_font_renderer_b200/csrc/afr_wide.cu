// Front-end of AttentionFontRenderer.forward (reference model.py:167-193) for nets WIDER than the
// reference's (BASELINE config 4: embed_dim 128, 8 heads, fc1 width 128, 64-char strings, 64 x 64
// sheets; any embed_dim % 32 == 0, head_dim in {8, 16, 32}, hidden % 32 == 0).
//
// At the reference's size (E = 32, 100 x 32 activations per sample) the whole front-end of a
// sample fits in one SM's shared memory and afr_frontend.cu fuses it into one kernel per
// direction. At E = 128 the weights alone (in_proj 196 KB) no longer fit, and the linear layers
// are real GEMMs over all B x S token rows: here they run on the tcgen05 GEMM of afr_gemm.cuh
// (bf16 operands, fp32 accumulation; the weight gradients -- one to three output tiles with
// K = B x S -- as split-K), and this file holds what is left between them: embedding + dropout +
// positions, soft-max attention per (sample, head), residual + LayerNorm, ReLU + dropout, and
// their backward passes. Arithmetic follows the oracle's restatement (oracle/afr_oracle.py
// features(); torch functional.py:5832-5846,6623-6654 for the attention), dropout decisions come
// from the same counter-based generator as the narrow path (afr_philox.cuh), so the oracle's
// builtin_masks() reproduces them. Parity for this path is "restated-oracle parity" at the bf16
// tolerance (2e-2): the reference has no such configuration.
#include <cstdlib>

#include "afr_internal.h"
#include "afr_philox.cuh"
#include "afr_ptx.cuh"

namespace afr {
namespace {

using ptx::ex2;

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

__device__ __forceinline__ uint32_t keep_bits8(const uint4& r, uint32_t thr) {
  return (u16_of<0>(r) >= thr ? 1u : 0u) | (u16_of<1>(r) >= thr ? 2u : 0u) | (u16_of<2>(r) >= thr ? 4u : 0u) |
         (u16_of<3>(r) >= thr ? 8u : 0u) | (u16_of<4>(r) >= thr ? 16u : 0u) | (u16_of<5>(r) >= thr ? 32u : 0u) |
         (u16_of<6>(r) >= thr ? 64u : 0u) | (u16_of<7>(r) >= thr ? 128u : 0u);
}
__device__ __forceinline__ uint4 pack_bf16x8(const float (&v)[8]) {
  uint4 o;
  __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
  __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
  o.x = *reinterpret_cast<uint32_t*>(&a); o.y = *reinterpret_cast<uint32_t*>(&b);
  o.z = *reinterpret_cast<uint32_t*>(&c); o.w = *reinterpret_cast<uint32_t*>(&d);
  return o;
}
// 8 consecutive channels of a SPLIT bf16 row [hi | lo | hi] (3 x width columns): hi = bf16(x),
// lo = bf16(x - hi). Multiplied with weight rows [hi | hi | lo] a GEMM over the 3 x width columns
// computes x_hi w_hi + x_lo w_hi + x_hi w_lo, the product to ~2^-16 -- the forward GEMMs of this
// path run that way, so the features (and with them every ReLU / dropout decision and the clamp
// at the output) agree with an fp32 forward; the backward GEMMs use the hi parts only.
__device__ __forceinline__ void store_split8(__nv_bfloat16* row, int width, const float (&v)[8]) {
  float lo[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) lo[u] = v[u] - __bfloat162float(__float2bfloat16_rn(v[u]));
  const uint4 hi = pack_bf16x8(v);
  *reinterpret_cast<uint4*>(row) = hi;
  *reinterpret_cast<uint4*>(row + width) = pack_bf16x8(lo);
  *reinterpret_cast<uint4*>(row + 2 * width) = hi;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}
__device__ __forceinline__ Rng make_rng(const WideDrop& d, int b) {
  return Rng{d.k0, d.k1, static_cast<uint32_t>(d.sample_offset + b), d.step};
}

// ---------------------------------------------------------------------------- embedding forward
// e = dropout(Emb[tok]) + Pos   (model.py:167-172: dropout BEFORE the positions are added).
// One thread per 8 channels = one Philox block. Writes the fp32 rows (residual, dWin) and their
// bf16 copy (A operand of the in-projection GEMM).
__global__ void __launch_bounds__(256)
wide_embed_kernel(WideDims d, WideDrop dr, const long long* __restrict__ tokens, long long stride,
                  const float* __restrict__ emb, const float* __restrict__ pos, float* __restrict__ e32,
                  __nv_bfloat16* __restrict__ e16, uint8_t* __restrict__ ebits, int* err_flag) {
  const int g8 = d.E / 8;
  const long long total = static_cast<long long>(d.B) * d.S * g8;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int c0 = static_cast<int>(i % g8) * 8;
    const long long r = i / g8;
    const int s = static_cast<int>(r % d.S), b = static_cast<int>(r / d.S);
    long long t = tokens[b * stride + s];
    if (t < 0 || t >= d.vocab) { atomicOr(err_flag, 1); t = 0; }
    uint32_t keep = 0xFFu;
    if (dr.mode == 1) keep = keep_bits8(make_rng(dr, b).block(0u, static_cast<uint32_t>((s * d.E + c0) >> 3)), dr.thr_e);
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(emb + t * d.E + c0));
    const float4 e1 = __ldg(reinterpret_cast<const float4*>(emb + t * d.E + c0 + 4));
    const float4 p0 = __ldg(reinterpret_cast<const float4*>(pos + s * d.E + c0));
    const float4 p1 = __ldg(reinterpret_cast<const float4*>(pos + s * d.E + c0 + 4));
    const float ev[8] = {e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w};
    const float pv[8] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) o[u] = ((keep >> u) & 1u) ? fmaf(ev[u], dr.inv_e, pv[u]) : pv[u];
    float* dst = e32 + r * d.E + c0;
    *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
    store_split8(e16 + r * 3 * d.E + c0, d.E, o);
    if (ebits != nullptr) ebits[i] = static_cast<uint8_t>(keep);     // keep bits of these 8 channels
  }
}

// ---------------------------------------------------------------------------- attention forward
// One CTA per sample; thread = (head h, query s), warps head-uniform (S padded to 32 per head) so
// every K / V row read is a shared-memory broadcast. Online soft-max over blocks of 8 keys in the
// log2 domain (MUFU.EX2), one Philox block per 8 keys. Leaves the context rows as bf16 (A operand
// of the out-projection GEMM) and, in training, (row max, 1 / row sum) and the keep bits.
template <int DH>
__global__ void __launch_bounds__(512)
wide_attention_fwd_kernel(WideDims d, WideDrop dr, const float* __restrict__ qkv, __nv_bfloat16* __restrict__ ctx16,
                          float2* __restrict__ stat, uint32_t* __restrict__ abits) {
  extern __shared__ __align__(16) float smem[];
  const int S = d.S, E = d.E, b = blockIdx.x;
  float* sk = smem;                   // [S + 8][E]
  float* sv = smem + (S + 8) * E;     // [S + 8][E]  (the last key block may read up to 7 rows past S)
  const float* base = qkv + static_cast<long long>(b) * S * 3 * E;
  for (int i = threadIdx.x; i < S * (E / 4); i += blockDim.x) {
    const int s = i / (E / 4), c4 = (i % (E / 4)) * 4;
    *reinterpret_cast<float4*>(sk + s * E + c4) = *reinterpret_cast<const float4*>(base + s * 3 * E + E + c4);
    *reinterpret_cast<float4*>(sv + s * E + c4) = *reinterpret_cast<const float4*>(base + s * 3 * E + 2 * E + c4);
  }
  for (int i = threadIdx.x; i < 8 * E; i += blockDim.x) { sk[S * E + i] = 0.f; sv[S * E + i] = 0.f; }
  __syncthreads();
  const int s_pad = (S + 31) & ~31;
  const float qscale = rsqrtf(static_cast<float>(DH)) * kLog2e;
  const Rng rng = make_rng(dr, b);
  for (int i = threadIdx.x; i < d.H * s_pad; i += blockDim.x) {
    const int h = i / s_pad, s = i % s_pad;
    if (s >= S) continue;
    // packed fp32 (FFMA2): q and the context accumulator as DH / 2 register pairs
    float2 q2[DH / 2], acc2[DH / 2];
#pragma unroll
    for (int j = 0; j < DH; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(base + s * 3 * E + h * DH + j);
      q2[j / 2] = make_float2(v.x * qscale, v.y * qscale);
      q2[j / 2 + 1] = make_float2(v.z * qscale, v.w * qscale);
    }
#pragma unroll
    for (int j = 0; j < DH / 2; ++j) acc2[j] = make_float2(0.f, 0.f);
    float m = -INFINITY, l = 0.f;
    const uint32_t row = static_cast<uint32_t>(h * S + s);
    const float* kh = sk + h * DH;
    const float* vh = sv + h * DH;
    const int nblk = (S + 7) >> 3;
    uint32_t bits = 0;
    for (int blk = 0; blk < nblk; ++blk) {
      const int t0 = blk * 8;
      float sc[8];
      float bm = m;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        float2 d2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < DH; j += 4) {
          const float4 kv = *reinterpret_cast<const float4*>(kh + (t0 + u) * E + j);
          d2 = ptx::fma2(q2[j / 2], make_float2(kv.x, kv.y), d2);
          d2 = ptx::fma2(q2[j / 2 + 1], make_float2(kv.z, kv.w), d2);
        }
        const float dsum = d2.x + d2.y;
        sc[u] = (t0 + u < S) ? dsum : -INFINITY;
        bm = fmaxf(bm, sc[u]);
      }
      const float corr = ex2(m - bm);
      m = bm;
      l *= corr;
      {
        const float2 c2 = make_float2(corr, corr);
#pragma unroll
        for (int j = 0; j < DH / 2; ++j) acc2[j] = ptx::mul2(acc2[j], c2);
      }
      uint32_t keep = 0xFFu;
      if (dr.mode == 1) keep = keep_bits8(rng.block(1u | (row << 2), static_cast<uint32_t>(blk)), dr.thr_a);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (t0 + u < S) {
          const float p = ex2(sc[u] - m);
          l += p;                                   // the denominator counts dropped keys too
          const float pk = ((keep >> u) & 1u) ? p : 0.f;
          const float2 p2 = make_float2(pk, pk);
#pragma unroll
          for (int j = 0; j < DH; j += 4) {
            const float4 vv = *reinterpret_cast<const float4*>(vh + (t0 + u) * E + j);
            acc2[j / 2] = ptx::fma2(p2, make_float2(vv.x, vv.y), acc2[j / 2]);
            acc2[j / 2 + 1] = ptx::fma2(p2, make_float2(vv.z, vv.w), acc2[j / 2 + 1]);
          }
        }
      }
      bits |= keep << (8 * (blk & 3));
      if ((blk & 3) == 3 || blk == nblk - 1) {
        if (abits != nullptr) abits[((static_cast<long long>(b) * d.H + h) * S + s) * 4 + (blk >> 2)] = bits;
        bits = 0;
      }
    }
    float acc[DH];
#pragma unroll
    for (int j = 0; j < DH / 2; ++j) { acc[2 * j] = acc2[j].x; acc[2 * j + 1] = acc2[j].y; }
    const float linv = 1.f / l;
    const float scale = dr.inv_a * linv;
    __nv_bfloat16* out = ctx16 + (static_cast<long long>(b) * S + s) * 3 * E + h * DH;
#pragma unroll
    for (int j = 0; j < DH; j += 8) {
      const float o[8] = {acc[j] * scale, acc[j + 1] * scale, acc[j + 2] * scale, acc[j + 3] * scale,
                          acc[j + 4] * scale, acc[j + 5] * scale, acc[j + 6] * scale, acc[j + 7] * scale};
      store_split8(out + j, E, o);
    }
    if (stat != nullptr) stat[(static_cast<long long>(b) * d.H + h) * S + s] = make_float2(m, linv);
  }
}

// ---------------------------------------------------------------------------- LayerNorm forward
// h = LayerNorm(e + a) (model.py:180): one warp per token row, lane owns channels lane + 32 i.
// Keeps the normalised residual and 1/std for the backward. h goes out as the fc1 GEMM's A operand
// in SPLIT bf16: row = [hi | lo | hi] (3E columns, hi = bf16(h), lo = bf16(h - hi)), multiplied with
// weight rows [hi | hi | lo] (wide_split_weight_kernel): h_hi W_hi + h_lo W_hi + h_hi W_lo, i.e. the
// fc1 pre-activation to ~2^-16 instead of 2^-8. It is the one place of the front-end where bf16
// rounding is not smooth: ReLU's derivative flips for every pre-activation within rounding error
// of zero (0.2-0.3 % of them with plain bf16 operands, which alone moves the gradients by ~4 %).
template <int NPL>   // channels per lane = E / 32
__global__ void __launch_bounds__(256)
wide_ln_fwd_kernel(long long rows, const float* __restrict__ e32, const float* __restrict__ a32,
                   const float* __restrict__ gamma, const float* __restrict__ beta, float* __restrict__ xhat,
                   float* __restrict__ rstd_out, __nv_bfloat16* __restrict__ h16) {
  constexpr int E = NPL * 32;
  const int lane = threadIdx.x & 31;
  float g[NPL], bt[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) { g[i] = gamma[lane + 32 * i]; bt[i] = beta[lane + 32 * i]; }
  for (long long r = blockIdx.x * 8ll + (threadIdx.x >> 5); r < rows; r += gridDim.x * 8ll) {
    float x[NPL], sum = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      x[i] = e32[r * E + lane + 32 * i] + a32[r * E + lane + 32 * i];
      sum += x[i];
    }
    const float mean = warp_sum(sum) * (1.f / E);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) { x[i] -= mean; sq = fmaf(x[i], x[i], sq); }
    const float var = warp_sum(sq) * (1.f / E);         // biased variance, eps = 1e-5
    const float rstd = 1.f / sqrtf(var + 1e-5f);
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const float xh = x[i] * rstd;
      if (xhat != nullptr) xhat[r * E + lane + 32 * i] = xh;
      const float hv = fmaf(xh, g[i], bt[i]);
      const __nv_bfloat16 hi = __float2bfloat16_rn(hv);
      const __nv_bfloat16 lo = __float2bfloat16_rn(hv - __bfloat162float(hi));
      __nv_bfloat16* row = h16 + r * 3 * E + lane + 32 * i;
      row[0] = hi; row[E] = lo; row[2 * E] = hi;
    }
    if (rstd_out != nullptr && lane == 0) rstd_out[r] = rstd;
  }
}

// weight [rows, E] fp32 -> [rows, 3E] bf16 rows [hi | hi | lo]   (see store_split8)
__global__ void __launch_bounds__(256)
wide_split_weight_kernel(const float* __restrict__ w, int rows, int E, __nv_bfloat16* __restrict__ out) {
  const int i = blockIdx.x * 256 + threadIdx.x;
  if (i >= rows * E) return;
  const int r = i / E, c = i % E;
  const float v = w[i];
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  const __nv_bfloat16 lo = __float2bfloat16_rn(v - __bfloat162float(hi));
  __nv_bfloat16* row = out + static_cast<long long>(r) * 3 * E + c;
  row[0] = hi; row[E] = hi; row[2 * E] = lo;
}

// ---------------------------------------------------------------------------- ReLU + dropout forward
// feats[b, s*F + j] = dropout(relu(fc1 pre-activation)) as bf16, the K-major operand of fc_output.
__global__ void __launch_bounds__(256)
wide_act_fwd_kernel(WideDims d, WideDrop dr, const float* __restrict__ f32, __nv_bfloat16* __restrict__ feats,
                    float* __restrict__ feats_f32) {
  const int g8 = d.F / 8;
  const long long total = static_cast<long long>(d.B) * d.S * g8;
  const long long KF = static_cast<long long>(d.L) * d.F;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int j0 = static_cast<int>(i % g8) * 8;
    const long long r = i / g8;
    const int s = static_cast<int>(r % d.S), b = static_cast<int>(r / d.S);
    uint32_t keep = 0xFFu;
    if (dr.mode == 1) keep = keep_bits8(make_rng(dr, b).block(2u, static_cast<uint32_t>((s * d.F + j0) >> 3)), dr.thr_f);
    const float4 a = *reinterpret_cast<const float4*>(f32 + r * d.F + j0);
    const float4 c = *reinterpret_cast<const float4*>(f32 + r * d.F + j0 + 4);
    const float x[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) o[u] = ((keep >> u) & 1u) ? fmaxf(x[u], 0.f) * dr.inv_f : 0.f;
    *reinterpret_cast<uint4*>(feats + b * KF + s * d.F + j0) = pack_bf16x8(o);
    if (feats_f32 != nullptr) {
      float* fo = feats_f32 + b * KF + s * d.F + j0;
      *reinterpret_cast<float4*>(fo) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(fo + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
  }
}

// ---------------------------------------------------------------------------- ReLU + dropout backward
// df = dfeat * 1[pre-activation > 0] * keep / (1 - p)  ->  bf16 rows [B*S, F]
__global__ void __launch_bounds__(256)
wide_act_bwd_kernel(WideDims d, WideDrop dr, const float* __restrict__ f32, const float* __restrict__ dfeat,
                    __nv_bfloat16* __restrict__ df16) {
  const int g8 = d.F / 8;
  const long long total = static_cast<long long>(d.B) * d.S * g8;
  const long long KF = static_cast<long long>(d.L) * d.F;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < total; i += gridDim.x * 256ll) {
    const int j0 = static_cast<int>(i % g8) * 8;
    const long long r = i / g8;
    const int s = static_cast<int>(r % d.S), b = static_cast<int>(r / d.S);
    uint32_t keep = 0xFFu;
    if (dr.mode == 1) keep = keep_bits8(make_rng(dr, b).block(2u, static_cast<uint32_t>((s * d.F + j0) >> 3)), dr.thr_f);
    const float4 a = *reinterpret_cast<const float4*>(f32 + r * d.F + j0);
    const float4 c = *reinterpret_cast<const float4*>(f32 + r * d.F + j0 + 4);
    const float4 ga = *reinterpret_cast<const float4*>(dfeat + b * KF + s * d.F + j0);
    const float4 gc = *reinterpret_cast<const float4*>(dfeat + b * KF + s * d.F + j0 + 4);
    const float x[8] = {a.x, a.y, a.z, a.w, c.x, c.y, c.z, c.w};
    const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gc.x, gc.y, gc.z, gc.w};
    float o[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) o[u] = (((keep >> u) & 1u) && x[u] > 0.f) ? gg[u] * dr.inv_f : 0.f;
    *reinterpret_cast<uint4*>(df16 + r * d.F + j0) = pack_bf16x8(o);
  }
}

// ---------------------------------------------------------------------------- LayerNorm backward
// dr = rstd * (dh*gamma - mean(dh*gamma) - xhat * mean(dh*gamma*xhat)); d(gamma) = sum dh*xhat,
// d(beta) = sum dh. Warp per row; the per-channel sums accumulate in registers over the rows of a
// warp, then per CTA in shared memory, and leave as one partial row per CTA (summed in a fixed
// order by wide_colsum_partials_kernel).
template <int NPL>
__global__ void __launch_bounds__(256)
wide_ln_bwd_kernel(long long rows, const float* __restrict__ dh32, const float* __restrict__ xhat,
                   const float* __restrict__ rstd_in, const float* __restrict__ gamma, float* __restrict__ dr32,
                   __nv_bfloat16* __restrict__ dr16, float* __restrict__ partials /* [grid][2E] */) {
  constexpr int E = NPL * 32;
  __shared__ float red[8][2 * E];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float g[NPL], dg[NPL], db[NPL];
#pragma unroll
  for (int i = 0; i < NPL; ++i) { g[i] = gamma[lane + 32 * i]; dg[i] = 0.f; db[i] = 0.f; }
  for (long long r = blockIdx.x * 8ll + warp; r < rows; r += gridDim.x * 8ll) {
    float dhv[NPL], xh[NPL], m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      dhv[i] = dh32[r * E + lane + 32 * i];
      xh[i] = xhat[r * E + lane + 32 * i];
      dg[i] = fmaf(dhv[i], xh[i], dg[i]);
      db[i] += dhv[i];
      const float t = dhv[i] * g[i];
      m1 += t;
      m2 = fmaf(t, xh[i], m2);
    }
    m1 = warp_sum(m1) * (1.f / E);
    m2 = warp_sum(m2) * (1.f / E);
    const float rs = rstd_in[r];
#pragma unroll
    for (int i = 0; i < NPL; ++i) {
      const float v = rs * (dhv[i] * g[i] - m1 - xh[i] * m2);
      dr32[r * E + lane + 32 * i] = v;
      dr16[r * E + lane + 32 * i] = __float2bfloat16_rn(v);
    }
  }
#pragma unroll
  for (int i = 0; i < NPL; ++i) { red[warp][lane + 32 * i] = dg[i]; red[warp][E + lane + 32 * i] = db[i]; }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * E; c += 256) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += red[w][c];
    partials[static_cast<long long>(blockIdx.x) * 2 * E + c] = s;
  }
}

// out[c] = sum over `n` partial rows of partials[i][c]  (fixed order)
__global__ void __launch_bounds__(256)
wide_colsum_partials_kernel(const float* __restrict__ partials, int n, int width, float* __restrict__ out0,
                            int split, float* __restrict__ out1) {
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c >= width) return;
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int i = 0;
  for (; i + 4 <= n; i += 4) {
    s0 += partials[static_cast<long long>(i) * width + c];
    s1 += partials[static_cast<long long>(i + 1) * width + c];
    s2 += partials[static_cast<long long>(i + 2) * width + c];
    s3 += partials[static_cast<long long>(i + 3) * width + c];
  }
  for (; i < n; ++i) s0 += partials[static_cast<long long>(i) * width + c];
  const float s = (s0 + s1) + (s2 + s3);
  if (c < split) out0[c] = s;
  else out1[c - split] = s;
}

// ---------------------------------------------------------------------------- attention backward
// One CTA per sample. q (scaled by log2(e)/sqrt(dh) like the forward), k, v and d(ctx) of the
// sample sit in shared memory; P is recomputed from the saved (row max, 1/row sum).
//   pass A, thread = (head, query): D = d(ctx) . ctx ; dS = P (keep ? dP : 0 - D) ; dq = dS k / sqrt(dh)
//   pass B, thread = (head, key)  : dk = sum_q dS q / sqrt(dh) ; dv = sum_q P_kept d(ctx) / (1 - p)
// Output: d(q | k | v) rows as bf16 [B*S, 3E] (operand of the in-projection's dgrad / wgrad GEMMs).
template <int DH>
__global__ void __launch_bounds__(512)
wide_attention_bwd_kernel(WideDims d, WideDrop dr, const float* __restrict__ qkv, const float* __restrict__ dctx,
                          const __nv_bfloat16* __restrict__ ctx16, const float2* __restrict__ stat,
                          const uint32_t* __restrict__ abits, __nv_bfloat16* __restrict__ dqkv16) {
  extern __shared__ __align__(16) float smem[];
  const int S = d.S, E = d.E, H = d.H, b = blockIdx.x;
  float* sq = smem;                 // [S][E]
  float* sk = sq + S * E;
  float* sv = sk + S * E;
  float* sc = sv + S * E;           // d(ctx)
  float* sD = sc + S * E;           // [S][H]
  const float* base = qkv + static_cast<long long>(b) * S * 3 * E;
  const float qscale = rsqrtf(static_cast<float>(DH)) * kLog2e;
  for (int i = threadIdx.x; i < S * (E / 4); i += blockDim.x) {
    const int s = i / (E / 4), c4 = (i % (E / 4)) * 4;
    float4 qv = *reinterpret_cast<const float4*>(base + s * 3 * E + c4);
    qv.x *= qscale; qv.y *= qscale; qv.z *= qscale; qv.w *= qscale;
    *reinterpret_cast<float4*>(sq + s * E + c4) = qv;
    *reinterpret_cast<float4*>(sk + s * E + c4) = *reinterpret_cast<const float4*>(base + s * 3 * E + E + c4);
    *reinterpret_cast<float4*>(sv + s * E + c4) = *reinterpret_cast<const float4*>(base + s * 3 * E + 2 * E + c4);
    *reinterpret_cast<float4*>(sc + s * E + c4) =
        *reinterpret_cast<const float4*>(dctx + (static_cast<long long>(b) * S + s) * E + c4);
  }
  __syncthreads();
  const int s_pad = (S + 31) & ~31;
  const float inv_a = dr.inv_a, inv_sqrt = rsqrtf(static_cast<float>(DH));
  // ---- pass A
  for (int i = threadIdx.x; i < H * s_pad; i += blockDim.x) {
    const int h = i / s_pad, s = i % s_pad;
    if (s >= S) continue;
    float q[DH], dc[DH], acc[DH];
    float Dv = 0.f;
    const __nv_bfloat16* cx = ctx16 + (static_cast<long long>(b) * S + s) * 3 * E + h * DH;   // [hi | lo | hi]
#pragma unroll
    for (int j = 0; j < DH; ++j) {
      q[j] = sq[s * E + h * DH + j];
      const float c = sc[s * E + h * DH + j];
      Dv = fmaf(c, __bfloat162float(cx[j]) + __bfloat162float(cx[E + j]), Dv);
      dc[j] = c * inv_a;
      acc[j] = 0.f;
    }
    sD[s * H + h] = Dv;
    const float2 st = stat[(static_cast<long long>(b) * H + h) * S + s];
    const uint32_t* bw = abits + ((static_cast<long long>(b) * H + h) * S + s) * 4;
    const float* kh = sk + h * DH;
    const float* vh = sv + h * DH;
    for (int t = 0; t < S; ++t) {
      float dot = 0.f, dp = 0.f;
#pragma unroll
      for (int j = 0; j < DH; j += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(kh + t * E + j);
        const float4 vv = *reinterpret_cast<const float4*>(vh + t * E + j);
        dot = fmaf(q[j], kv.x, dot); dot = fmaf(q[j + 1], kv.y, dot);
        dot = fmaf(q[j + 2], kv.z, dot); dot = fmaf(q[j + 3], kv.w, dot);
        dp = fmaf(dc[j], vv.x, dp); dp = fmaf(dc[j + 1], vv.y, dp);
        dp = fmaf(dc[j + 2], vv.z, dp); dp = fmaf(dc[j + 3], vv.w, dp);
      }
      const float p = ex2(dot - st.x) * st.y;
      const bool keep = (bw[t >> 5] >> (t & 31)) & 1u;
      const float ds = p * ((keep ? dp : 0.f) - Dv);
#pragma unroll
      for (int j = 0; j < DH; j += 4) {
        const float4 kv = *reinterpret_cast<const float4*>(kh + t * E + j);
        acc[j] = fmaf(ds, kv.x, acc[j]); acc[j + 1] = fmaf(ds, kv.y, acc[j + 1]);
        acc[j + 2] = fmaf(ds, kv.z, acc[j + 2]); acc[j + 3] = fmaf(ds, kv.w, acc[j + 3]);
      }
    }
    __nv_bfloat16* out = dqkv16 + (static_cast<long long>(b) * S + s) * 3 * E + h * DH;
#pragma unroll
    for (int j = 0; j < DH; j += 8) {
      const float o[8] = {acc[j] * inv_sqrt, acc[j + 1] * inv_sqrt, acc[j + 2] * inv_sqrt, acc[j + 3] * inv_sqrt,
                          acc[j + 4] * inv_sqrt, acc[j + 5] * inv_sqrt, acc[j + 6] * inv_sqrt, acc[j + 7] * inv_sqrt};
      *reinterpret_cast<uint4*>(out + j) = pack_bf16x8(o);
    }
  }
  __syncthreads();
  // ---- pass B
  for (int i = threadIdx.x; i < H * s_pad; i += blockDim.x) {
    const int h = i / s_pad, t = i % s_pad;
    if (t >= S) continue;
    float k[DH], v[DH], ak[DH], av[DH];
#pragma unroll
    for (int j = 0; j < DH; ++j) {
      k[j] = sk[t * E + h * DH + j];
      v[j] = sv[t * E + h * DH + j] * inv_a;
      ak[j] = 0.f; av[j] = 0.f;
    }
    const float* qh = sq + h * DH;
    const float* ch = sc + h * DH;
    const float2* sth = stat + (static_cast<long long>(b) * H + h) * S;
    const uint32_t* bh = abits + (static_cast<long long>(b) * H + h) * S * 4 + (t >> 5);
    const int sh = t & 31;
    for (int s = 0; s < S; ++s) {
      float dot = 0.f, dp = 0.f;
#pragma unroll
      for (int j = 0; j < DH; j += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(qh + s * E + j);
        const float4 cv = *reinterpret_cast<const float4*>(ch + s * E + j);
        dot = fmaf(k[j], qv.x, dot); dot = fmaf(k[j + 1], qv.y, dot);
        dot = fmaf(k[j + 2], qv.z, dot); dot = fmaf(k[j + 3], qv.w, dot);
        dp = fmaf(v[j], cv.x, dp); dp = fmaf(v[j + 1], cv.y, dp);
        dp = fmaf(v[j + 2], cv.z, dp); dp = fmaf(v[j + 3], cv.w, dp);
      }
      const float2 st = sth[s];
      const bool keep = (bh[s * 4] >> sh) & 1u;
      const float p = ex2(dot - st.x) * st.y;
      const float pk = keep ? p : 0.f;
      const float ds = p * ((keep ? dp : 0.f) - sD[s * H + h]);
#pragma unroll
      for (int j = 0; j < DH; j += 4) {
        const float4 qv = *reinterpret_cast<const float4*>(qh + s * E + j);
        const float4 cv = *reinterpret_cast<const float4*>(ch + s * E + j);
        ak[j] = fmaf(ds, qv.x, ak[j]); ak[j + 1] = fmaf(ds, qv.y, ak[j + 1]);
        ak[j + 2] = fmaf(ds, qv.z, ak[j + 2]); ak[j + 3] = fmaf(ds, qv.w, ak[j + 3]);
        av[j] = fmaf(pk, cv.x, av[j]); av[j + 1] = fmaf(pk, cv.y, av[j + 1]);
        av[j + 2] = fmaf(pk, cv.z, av[j + 2]); av[j + 3] = fmaf(pk, cv.w, av[j + 3]);
      }
    }
    __nv_bfloat16* outk = dqkv16 + (static_cast<long long>(b) * S + t) * 3 * E + E + h * DH;
    __nv_bfloat16* outv = outk + E;
#pragma unroll
    for (int j = 0; j < DH; j += 8) {
      // the stored q carries log2(e)/sqrt(dh): d(k) = sum dS q_raw / sqrt(dh) = ak * ln 2
      const float ok[8] = {ak[j] * kLn2, ak[j + 1] * kLn2, ak[j + 2] * kLn2, ak[j + 3] * kLn2,
                           ak[j + 4] * kLn2, ak[j + 5] * kLn2, ak[j + 6] * kLn2, ak[j + 7] * kLn2};
      const float ov[8] = {av[j] * inv_a, av[j + 1] * inv_a, av[j + 2] * inv_a, av[j + 3] * inv_a,
                           av[j + 4] * inv_a, av[j + 5] * inv_a, av[j + 6] * inv_a, av[j + 7] * inv_a};
      *reinterpret_cast<uint4*>(outk + j) = pack_bf16x8(ok);
      *reinterpret_cast<uint4*>(outv + j) = pack_bf16x8(ov);
    }
  }
}


// ---------------------------------------------------------------------------- attention backward, tensor cores
// The same two passes on warp-level bf16 MMAs (mma.sync.m16n8k16, fp32 accumulate): 16-row tiles of
// one head per warp; S = Q K^T and dP = dO V^T per 8-key tile, soft-max recomputed from the saved
// row statistics, dS = P (keep ? dP : 0 - D) converted in registers into the A fragment of the
// next product (dQ += dS K; pass 2 with keys as rows: dK += dS^T Q, dV += P_kept^T dO). The score
// product uses split-bf16 q and k (q_hi k_hi + q_lo k_hi + q_hi k_lo) so that the recomputed P
// matches the forward's fp32 P; everything else is plain bf16 (the backward is smooth).
// Operands sit in shared memory as bf16, row-major with 8 elements of padding per row (conflict-
// free fragment loads) plus transposed copies of q, k, dO for the products whose reduction index
// is the row index. These are the legacy tensor instructions: a head is 64 x 64 x 16, far too
// small for a tcgen05 tile of 128 rows with its TMEM round trip per soft-max.
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t lds_u32(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }

struct AttnMmaSmem {
  int ldr, ldt, sp, nw;
  size_t q_hi, q_lo, k_hi, k_lo, v, dO, qT, kT, dOT, stat, D, ab, total;   // byte offsets
};
__host__ __device__ inline AttnMmaSmem attn_mma_smem(int S, int E, int H) {
  AttnMmaSmem m{};
  m.sp = (S + 15) & ~15;
  m.ldr = E + 8;
  m.ldt = m.sp + 8;
  m.nw = (S + 31) >> 5;
  size_t o = 0;
  const size_t row = static_cast<size_t>(m.sp) * m.ldr * 2, tr = static_cast<size_t>(E) * m.ldt * 2;
  m.q_hi = o; o += row; m.q_lo = o; o += row; m.k_hi = o; o += row; m.k_lo = o; o += row;
  m.v = o; o += row; m.dO = o; o += row;
  m.qT = o; o += tr; m.kT = o; o += tr; m.dOT = o; o += tr;
  m.stat = o; o += static_cast<size_t>(H) * m.sp * 8;
  m.D = o; o += static_cast<size_t>(H) * m.sp * 4;
  m.ab = o; o += static_cast<size_t>(H) * m.sp * m.nw * 4;
  m.total = o;
  return m;
}

template <int DH>
__global__ void __launch_bounds__(512, 1)
wide_attention_bwd_mma_kernel(WideDims d, WideDrop dr, const float* __restrict__ qkv, const float* __restrict__ dctx,
                              const __nv_bfloat16* __restrict__ ctx16, const float2* __restrict__ stat,
                              const uint32_t* __restrict__ abits, __nv_bfloat16* __restrict__ dqkv16) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int KK = DH / 16, NT = DH / 8;
  const int S = d.S, E = d.E, H = d.H, b = blockIdx.x;
  const AttnMmaSmem L = attn_mma_smem(S, E, H);
  const int SP = L.sp, LDR = L.ldr, LDT = L.ldt, NW = L.nw;
  __nv_bfloat16* q_hi = reinterpret_cast<__nv_bfloat16*>(smraw + L.q_hi);
  __nv_bfloat16* q_lo = reinterpret_cast<__nv_bfloat16*>(smraw + L.q_lo);
  __nv_bfloat16* k_hi = reinterpret_cast<__nv_bfloat16*>(smraw + L.k_hi);
  __nv_bfloat16* k_lo = reinterpret_cast<__nv_bfloat16*>(smraw + L.k_lo);
  __nv_bfloat16* v_s = reinterpret_cast<__nv_bfloat16*>(smraw + L.v);
  __nv_bfloat16* do_s = reinterpret_cast<__nv_bfloat16*>(smraw + L.dO);
  __nv_bfloat16* qT = reinterpret_cast<__nv_bfloat16*>(smraw + L.qT);
  __nv_bfloat16* kT = reinterpret_cast<__nv_bfloat16*>(smraw + L.kT);
  __nv_bfloat16* doT = reinterpret_cast<__nv_bfloat16*>(smraw + L.dOT);
  float2* st_s = reinterpret_cast<float2*>(smraw + L.stat);
  float* D_s = reinterpret_cast<float*>(smraw + L.D);
  uint32_t* ab_s = reinterpret_cast<uint32_t*>(smraw + L.ab);
  const float* base = qkv + static_cast<long long>(b) * S * 3 * E;
  const float qscale = rsqrtf(static_cast<float>(DH)) * kLog2e;
  const float inv_a = dr.inv_a, inv_sqrt = rsqrtf(static_cast<float>(DH));
  const __nv_bfloat16 zero = __float2bfloat16_rn(0.f);

  // ---- stage the sample: bf16 (split for q, k), row-major + transposed; zero padding.
  // Four channels per thread and iteration (128-bit global loads, 64-bit row-major stores).
  for (int i = threadIdx.x; i < SP * (E / 4); i += blockDim.x) {
    // a warp covers 8 rows x 16 channels: 64 contiguous bytes per row from global, and the
    // transposed 2-byte stores of 8 consecutive rows share words (2-way conflicts instead of 16-way)
    const int wi = i >> 5, li = i & 31;
    const int s = 8 * (wi / (E / 16)) + (li & 7), c = 16 * (wi % (E / 16)) + 4 * (li >> 3);
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv, vv = qv, dv = qv;
    if (s < S) {
      qv = *reinterpret_cast<const float4*>(base + s * 3 * E + c);
      kv = *reinterpret_cast<const float4*>(base + s * 3 * E + E + c);
      vv = *reinterpret_cast<const float4*>(base + s * 3 * E + 2 * E + c);
      dv = *reinterpret_cast<const float4*>(dctx + (static_cast<long long>(b) * S + s) * E + c);
    }
    const float q4[4] = {qv.x * qscale, qv.y * qscale, qv.z * qscale, qv.w * qscale};
    const float k4[4] = {kv.x, kv.y, kv.z, kv.w};
    const float d4[4] = {dv.x * inv_a, dv.y * inv_a, dv.z * inv_a, dv.w * inv_a};
    __nv_bfloat16 qh[4], kh[4], dh4[4];
    float ql[4], kl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      qh[u] = __float2bfloat16_rn(q4[u]); kh[u] = __float2bfloat16_rn(k4[u]); dh4[u] = __float2bfloat16_rn(d4[u]);
      ql[u] = q4[u] - __bfloat162float(qh[u]);
      kl[u] = k4[u] - __bfloat162float(kh[u]);
      qT[(c + u) * LDT + s] = qh[u];
      kT[(c + u) * LDT + s] = kh[u];
      doT[(c + u) * LDT + s] = dh4[u];
    }
    auto st4 = [&](__nv_bfloat16* dst, float a0, float a1, float a2, float a3) {
      *reinterpret_cast<uint2*>(dst + s * LDR + c) = make_uint2(pack2(a0, a1), pack2(a2, a3));
    };
    st4(q_hi, q4[0], q4[1], q4[2], q4[3]);
    st4(q_lo, ql[0], ql[1], ql[2], ql[3]);
    st4(k_hi, k4[0], k4[1], k4[2], k4[3]);
    st4(k_lo, kl[0], kl[1], kl[2], kl[3]);
    st4(v_s, vv.x, vv.y, vv.z, vv.w);
    st4(do_s, d4[0], d4[1], d4[2], d4[3]);
  }
  for (int i = threadIdx.x; i < E * 8; i += blockDim.x) {     // the 8 padding columns of the transposed copies
    const int c = i >> 3, s = SP + (i & 7);
    qT[c * LDT + s] = zero; kT[c * LDT + s] = zero; doT[c * LDT + s] = zero;
  }
  for (int i = threadIdx.x; i < H * SP; i += blockDim.x) {
    const int h = i / SP, s = i % SP;
    float2 stv = make_float2(0.f, 0.f);
    float Dv = 0.f;
    if (s < S) {
      stv = stat[(static_cast<long long>(b) * H + h) * S + s];
      const float* cg = dctx + (static_cast<long long>(b) * S + s) * E + h * DH;
      const __nv_bfloat16* cx = ctx16 + (static_cast<long long>(b) * S + s) * 3 * E + h * DH;   // [hi | lo | hi]
      float4 gv[DH / 4];
      uint4 hv[DH / 8], lv[DH / 8];     // all loads first: the chain below then waits once, not DH times
#pragma unroll
      for (int j = 0; j < DH / 4; ++j) gv[j] = *reinterpret_cast<const float4*>(cg + 4 * j);
#pragma unroll
      for (int j = 0; j < DH / 8; ++j) {
        hv[j] = *reinterpret_cast<const uint4*>(cx + 8 * j);
        lv[j] = *reinterpret_cast<const uint4*>(cx + E + 8 * j);
      }
#pragma unroll
      for (int j = 0; j < DH / 8; ++j) {
        const uint32_t hw[4] = {hv[j].x, hv[j].y, hv[j].z, hv[j].w}, lw[4] = {lv[j].x, lv[j].y, lv[j].z, lv[j].w};
        const float ga[8] = {gv[2 * j].x, gv[2 * j].y, gv[2 * j].z, gv[2 * j].w,
                             gv[2 * j + 1].x, gv[2 * j + 1].y, gv[2 * j + 1].z, gv[2 * j + 1].w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {     // a bf16 is the upper half of its fp32
          const float c0 = __uint_as_float(hw[u] << 16) + __uint_as_float(lw[u] << 16);
          const float c1 = __uint_as_float(hw[u] & 0xffff0000u) + __uint_as_float(lw[u] & 0xffff0000u);
          Dv = fmaf(ga[2 * u], c0, Dv);
          Dv = fmaf(ga[2 * u + 1], c1, Dv);
        }
      }
    }
    st_s[i] = stv;
    D_s[i] = Dv;
    for (int w = 0; w < NW; ++w)
      ab_s[i * NW + w] = s < S ? abits[((static_cast<long long>(b) * H + h) * S + s) * 4 + w] : 0u;
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int tiles = SP >> 4, tasks = H * tiles;

  // ---- pass 1: rows = queries -> dq
  for (int task = warp; task < tasks; task += nwarps) {
    const int h = task / tiles, i0 = (task % tiles) << 4;
    const int rA = i0 + g, rB = rA + 8;
    uint32_t aqh[KK][4], aql[KK][4], ado[KK][4];
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      const int c0 = h * DH + kk * 16 + 2 * t;
      aqh[kk][0] = lds_u32(q_hi + rA * LDR + c0); aqh[kk][1] = lds_u32(q_hi + rB * LDR + c0);
      aqh[kk][2] = lds_u32(q_hi + rA * LDR + c0 + 8); aqh[kk][3] = lds_u32(q_hi + rB * LDR + c0 + 8);
      aql[kk][0] = lds_u32(q_lo + rA * LDR + c0); aql[kk][1] = lds_u32(q_lo + rB * LDR + c0);
      aql[kk][2] = lds_u32(q_lo + rA * LDR + c0 + 8); aql[kk][3] = lds_u32(q_lo + rB * LDR + c0 + 8);
      ado[kk][0] = lds_u32(do_s + rA * LDR + c0); ado[kk][1] = lds_u32(do_s + rB * LDR + c0);
      ado[kk][2] = lds_u32(do_s + rA * LDR + c0 + 8); ado[kk][3] = lds_u32(do_s + rB * LDR + c0 + 8);
    }
    const float2 stA = st_s[h * SP + rA], stB = st_s[h * SP + rB];
    const float DA = D_s[h * SP + rA], DB = D_s[h * SP + rB];
    const uint32_t* wA = ab_s + (h * SP + rA) * NW;
    const uint32_t* wB = ab_s + (h * SP + rB) * NW;
    float dq[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { dq[nt][0] = dq[nt][1] = dq[nt][2] = dq[nt][3] = 0.f; }
    for (int jj = 0; jj < tiles; ++jj) {
      float ds[2][4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j0 = jj * 16 + half * 8;
        float c[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          const int c0 = h * DH + kk * 16 + 2 * t;
          const uint32_t bh0 = lds_u32(k_hi + (j0 + g) * LDR + c0), bh1 = lds_u32(k_hi + (j0 + g) * LDR + c0 + 8);
          const uint32_t bl0 = lds_u32(k_lo + (j0 + g) * LDR + c0), bl1 = lds_u32(k_lo + (j0 + g) * LDR + c0 + 8);
          mma_bf16_16816(c, aqh[kk], bh0, bh1);
          mma_bf16_16816(c, aql[kk], bh0, bh1);
          mma_bf16_16816(c, aqh[kk], bl0, bl1);
          const uint32_t bv0 = lds_u32(v_s + (j0 + g) * LDR + c0), bv1 = lds_u32(v_s + (j0 + g) * LDR + c0 + 8);
          mma_bf16_16816(e, ado[kk], bv0, bv1);
        }
        const int key0 = j0 + 2 * t, key1 = key0 + 1;
        const uint32_t wa = wA[key0 >> 5], wb = wB[key0 >> 5];       // key0, key1 share a word (key0 is even)
        const float pA0 = key0 < S ? ex2(c[0] - stA.x) * stA.y : 0.f, pA1 = key1 < S ? ex2(c[1] - stA.x) * stA.y : 0.f;
        const float pB0 = key0 < S ? ex2(c[2] - stB.x) * stB.y : 0.f, pB1 = key1 < S ? ex2(c[3] - stB.x) * stB.y : 0.f;
        ds[half][0] = pA0 * ((((wa >> (key0 & 31)) & 1u) ? e[0] : 0.f) - DA);
        ds[half][1] = pA1 * ((((wa >> (key1 & 31)) & 1u) ? e[1] : 0.f) - DA);
        ds[half][2] = pB0 * ((((wb >> (key0 & 31)) & 1u) ? e[2] : 0.f) - DB);
        ds[half][3] = pB1 * ((((wb >> (key1 & 31)) & 1u) ? e[3] : 0.f) - DB);
      }
      const uint32_t ads[4] = {pack2(ds[0][0], ds[0][1]), pack2(ds[0][2], ds[0][3]),
                               pack2(ds[1][0], ds[1][1]), pack2(ds[1][2], ds[1][3])};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const __nv_bfloat16* kp = kT + (h * DH + nt * 8 + g) * LDT + jj * 16 + 2 * t;
        mma_bf16_16816(dq[nt], ads, lds_u32(kp), lds_u32(kp + 8));
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = h * DH + nt * 8 + 2 * t;
      if (rA < S)
        *reinterpret_cast<uint32_t*>(dqkv16 + (static_cast<long long>(b) * S + rA) * 3 * E + col) =
            pack2(dq[nt][0] * inv_sqrt, dq[nt][1] * inv_sqrt);
      if (rB < S)
        *reinterpret_cast<uint32_t*>(dqkv16 + (static_cast<long long>(b) * S + rB) * 3 * E + col) =
            pack2(dq[nt][2] * inv_sqrt, dq[nt][3] * inv_sqrt);
    }
  }

  // ---- pass 2: rows = keys -> dk, dv
  for (int task = warp; task < tasks; task += nwarps) {
    const int h = task / tiles, j0 = (task % tiles) << 4;
    const int kA = j0 + g, kB = kA + 8;
    uint32_t akh[KK][4], akl[KK][4], av[KK][4];
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      const int c0 = h * DH + kk * 16 + 2 * t;
      akh[kk][0] = lds_u32(k_hi + kA * LDR + c0); akh[kk][1] = lds_u32(k_hi + kB * LDR + c0);
      akh[kk][2] = lds_u32(k_hi + kA * LDR + c0 + 8); akh[kk][3] = lds_u32(k_hi + kB * LDR + c0 + 8);
      akl[kk][0] = lds_u32(k_lo + kA * LDR + c0); akl[kk][1] = lds_u32(k_lo + kB * LDR + c0);
      akl[kk][2] = lds_u32(k_lo + kA * LDR + c0 + 8); akl[kk][3] = lds_u32(k_lo + kB * LDR + c0 + 8);
      av[kk][0] = lds_u32(v_s + kA * LDR + c0); av[kk][1] = lds_u32(v_s + kB * LDR + c0);
      av[kk][2] = lds_u32(v_s + kA * LDR + c0 + 8); av[kk][3] = lds_u32(v_s + kB * LDR + c0 + 8);
    }
    const bool okA = kA < S, okB = kB < S;
    const int wsel = kA >> 5, shA = kA & 31, shB = kB & 31;      // kA and kB lie in the same 32-key word
    float dk[NT][4], dvv[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      dk[nt][0] = dk[nt][1] = dk[nt][2] = dk[nt][3] = 0.f;
      dvv[nt][0] = dvv[nt][1] = dvv[nt][2] = dvv[nt][3] = 0.f;
    }
    for (int ii = 0; ii < tiles; ++ii) {
      float dsT[2][4], pT[2][4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int i0 = ii * 16 + half * 8;
        float c[4] = {0.f, 0.f, 0.f, 0.f}, e[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          const int c0 = h * DH + kk * 16 + 2 * t;
          const uint32_t bh0 = lds_u32(q_hi + (i0 + g) * LDR + c0), bh1 = lds_u32(q_hi + (i0 + g) * LDR + c0 + 8);
          const uint32_t bl0 = lds_u32(q_lo + (i0 + g) * LDR + c0), bl1 = lds_u32(q_lo + (i0 + g) * LDR + c0 + 8);
          mma_bf16_16816(c, akh[kk], bh0, bh1);
          mma_bf16_16816(c, akl[kk], bh0, bh1);
          mma_bf16_16816(c, akh[kk], bl0, bl1);
          const uint32_t bd0 = lds_u32(do_s + (i0 + g) * LDR + c0), bd1 = lds_u32(do_s + (i0 + g) * LDR + c0 + 8);
          mma_bf16_16816(e, av[kk], bd0, bd1);
        }
        const int q0 = i0 + 2 * t, q1 = q0 + 1;
        const float2 st0 = st_s[h * SP + q0], st1 = st_s[h * SP + q1];     // (0, 0) for padded queries: p = 0
        const float D0 = D_s[h * SP + q0], D1 = D_s[h * SP + q1];
        const uint32_t w0 = ab_s[(h * SP + q0) * NW + wsel], w1 = ab_s[(h * SP + q1) * NW + wsel];
        const float p00 = okA ? ex2(c[0] - st0.x) * st0.y : 0.f, p01 = okA ? ex2(c[1] - st1.x) * st1.y : 0.f;
        const float p10 = okB ? ex2(c[2] - st0.x) * st0.y : 0.f, p11 = okB ? ex2(c[3] - st1.x) * st1.y : 0.f;
        const bool k00 = (w0 >> shA) & 1u, k01 = (w1 >> shA) & 1u, k10 = (w0 >> shB) & 1u, k11 = (w1 >> shB) & 1u;
        dsT[half][0] = p00 * ((k00 ? e[0] : 0.f) - D0); dsT[half][1] = p01 * ((k01 ? e[1] : 0.f) - D1);
        dsT[half][2] = p10 * ((k10 ? e[2] : 0.f) - D0); dsT[half][3] = p11 * ((k11 ? e[3] : 0.f) - D1);
        pT[half][0] = k00 ? p00 : 0.f; pT[half][1] = k01 ? p01 : 0.f;
        pT[half][2] = k10 ? p10 : 0.f; pT[half][3] = k11 ? p11 : 0.f;
      }
      const uint32_t ads[4] = {pack2(dsT[0][0], dsT[0][1]), pack2(dsT[0][2], dsT[0][3]),
                               pack2(dsT[1][0], dsT[1][1]), pack2(dsT[1][2], dsT[1][3])};
      const uint32_t ap[4] = {pack2(pT[0][0], pT[0][1]), pack2(pT[0][2], pT[0][3]),
                              pack2(pT[1][0], pT[1][1]), pack2(pT[1][2], pT[1][3])};
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int off = (h * DH + nt * 8 + g) * LDT + ii * 16 + 2 * t;
        mma_bf16_16816(dk[nt], ads, lds_u32(qT + off), lds_u32(qT + off + 8));
        mma_bf16_16816(dvv[nt], ap, lds_u32(doT + off), lds_u32(doT + off + 8));
      }
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = h * DH + nt * 8 + 2 * t;
      // the stored q carries log2(e)/sqrt(dh): d(k) = sum dS q_raw / sqrt(dh) = dk * ln 2
      if (okA) {
        __nv_bfloat16* o = dqkv16 + (static_cast<long long>(b) * S + kA) * 3 * E + E + col;
        *reinterpret_cast<uint32_t*>(o) = pack2(dk[nt][0] * kLn2, dk[nt][1] * kLn2);
        *reinterpret_cast<uint32_t*>(o + E) = pack2(dvv[nt][0], dvv[nt][1]);
      }
      if (okB) {
        __nv_bfloat16* o = dqkv16 + (static_cast<long long>(b) * S + kB) * 3 * E + E + col;
        *reinterpret_cast<uint32_t*>(o) = pack2(dk[nt][2] * kLn2, dk[nt][3] * kLn2);
        *reinterpret_cast<uint32_t*>(o + E) = pack2(dvv[nt][2], dvv[nt][3]);
      }
    }
  }
}

// ---------------------------------------------------------------------------- attention forward, tensor cores
// The forward on the same warp-level bf16 MMAs, every product with split operands so that the
// context keeps the fp32 accuracy the ReLU downstream needs (see wide_ln_fwd_kernel):
//   scores = q_hi k_hi + q_lo k_hi + q_hi k_lo,   ctx = p_hi v_hi + p_lo v_hi + p_hi v_lo.
// A warp owns a (head, 16-query tile) and walks the keys 16 at a time: two 16 x 8 score tiles,
// online soft-max on the accumulator fragments (row max / sum reduced over the four lanes that
// share a row), P_kept re-packed in registers as the A fragment of the context product (two
// adjacent 8-column score tiles = one 16-wide k-step), V read from a transposed bf16 copy.
// The dropout decisions are the SIMT kernel's (one Philox block per row and 8 keys): lane t of a
// quad draws the block of (row g or g + 8, first or second tile) and its 8 keep bits reach the
// other three lanes by shuffle. One CTA per sample, q / k (row-major) and v^T as hi / lo bf16 in
// shared memory: 106 KB at config 4, two CTAs per SM.
struct AttnFwdSmem {
  int ldr, ldt, sp;
  size_t q_hi, q_lo, k_hi, k_lo, vT_hi, vT_lo, total;   // byte offsets
};
__host__ __device__ inline AttnFwdSmem attn_fwd_smem(int S, int E) {
  AttnFwdSmem m{};
  m.sp = (S + 15) & ~15;
  m.ldr = E + 8;
  m.ldt = m.sp + 8;
  size_t o = 0;
  const size_t row = static_cast<size_t>(m.sp) * m.ldr * 2, tr = static_cast<size_t>(E) * m.ldt * 2;
  m.q_hi = o; o += row; m.q_lo = o; o += row; m.k_hi = o; o += row; m.k_lo = o; o += row;
  m.vT_hi = o; o += tr; m.vT_lo = o; o += tr;
  m.total = o;
  return m;
}

template <int DH>
__global__ void __launch_bounds__(512, 2)
wide_attention_fwd_mma_kernel(WideDims d, WideDrop dr, const float* __restrict__ qkv, __nv_bfloat16* __restrict__ ctx16,
                              float2* __restrict__ stat, uint32_t* __restrict__ abits) {
  extern __shared__ __align__(16) unsigned char smraw[];
  constexpr int KK = DH / 16, NT = DH / 8;
  const int S = d.S, E = d.E, H = d.H, b = blockIdx.x;
  const AttnFwdSmem L = attn_fwd_smem(S, E);
  const int SP = L.sp, LDR = L.ldr, LDT = L.ldt;
  __nv_bfloat16* q_hi = reinterpret_cast<__nv_bfloat16*>(smraw + L.q_hi);
  __nv_bfloat16* q_lo = reinterpret_cast<__nv_bfloat16*>(smraw + L.q_lo);
  __nv_bfloat16* k_hi = reinterpret_cast<__nv_bfloat16*>(smraw + L.k_hi);
  __nv_bfloat16* k_lo = reinterpret_cast<__nv_bfloat16*>(smraw + L.k_lo);
  __nv_bfloat16* vT_hi = reinterpret_cast<__nv_bfloat16*>(smraw + L.vT_hi);
  __nv_bfloat16* vT_lo = reinterpret_cast<__nv_bfloat16*>(smraw + L.vT_lo);
  const float* base = qkv + static_cast<long long>(b) * S * 3 * E;
  const float qscale = rsqrtf(static_cast<float>(DH)) * kLog2e;

  // ---- stage the sample as split bf16 (rows >= S: zero)
  for (int i = threadIdx.x; i < SP * (E / 4); i += blockDim.x) {
    const int wi = i >> 5, li = i & 31;     // a warp covers 8 rows x 16 channels (see the backward kernel)
    const int s = 8 * (wi / (E / 16)) + (li & 7), c = 16 * (wi % (E / 16)) + 4 * (li >> 3);
    float4 qv = make_float4(0.f, 0.f, 0.f, 0.f), kv = qv, vv = qv;
    if (s < S) {
      qv = *reinterpret_cast<const float4*>(base + s * 3 * E + c);
      kv = *reinterpret_cast<const float4*>(base + s * 3 * E + E + c);
      vv = *reinterpret_cast<const float4*>(base + s * 3 * E + 2 * E + c);
    }
    const float q4[4] = {qv.x * qscale, qv.y * qscale, qv.z * qscale, qv.w * qscale};
    const float k4[4] = {kv.x, kv.y, kv.z, kv.w};
    const float v4[4] = {vv.x, vv.y, vv.z, vv.w};
    float ql[4], kl[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      ql[u] = q4[u] - __bfloat162float(__float2bfloat16_rn(q4[u]));
      kl[u] = k4[u] - __bfloat162float(__float2bfloat16_rn(k4[u]));
      const __nv_bfloat16 vh = __float2bfloat16_rn(v4[u]);
      vT_hi[(c + u) * LDT + s] = vh;
      vT_lo[(c + u) * LDT + s] = __float2bfloat16_rn(v4[u] - __bfloat162float(vh));
    }
    auto st4 = [&](__nv_bfloat16* dst, float a0, float a1, float a2, float a3) {
      *reinterpret_cast<uint2*>(dst + s * LDR + c) = make_uint2(pack2(a0, a1), pack2(a2, a3));
    };
    st4(q_hi, q4[0], q4[1], q4[2], q4[3]);
    st4(q_lo, ql[0], ql[1], ql[2], ql[3]);
    st4(k_hi, k4[0], k4[1], k4[2], k4[3]);
    st4(k_lo, kl[0], kl[1], kl[2], kl[3]);
  }
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int tiles = SP >> 4, tasks = H * tiles;
  const Rng rng = make_rng(dr, b);

  for (int task = warp; task < tasks; task += nwarps) {
    const int h = task / tiles, i0 = (task % tiles) << 4;
    const int rA = i0 + g, rB = rA + 8;
    uint32_t aqh[KK][4], aql[KK][4];
#pragma unroll
    for (int kk = 0; kk < KK; ++kk) {
      const int c0 = h * DH + kk * 16 + 2 * t;
      aqh[kk][0] = lds_u32(q_hi + rA * LDR + c0); aqh[kk][1] = lds_u32(q_hi + rB * LDR + c0);
      aqh[kk][2] = lds_u32(q_hi + rA * LDR + c0 + 8); aqh[kk][3] = lds_u32(q_hi + rB * LDR + c0 + 8);
      aql[kk][0] = lds_u32(q_lo + rA * LDR + c0); aql[kk][1] = lds_u32(q_lo + rB * LDR + c0);
      aql[kk][2] = lds_u32(q_lo + rA * LDR + c0 + 8); aql[kk][3] = lds_u32(q_lo + rB * LDR + c0 + 8);
    }
    float o[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) { o[nt][0] = o[nt][1] = o[nt][2] = o[nt][3] = 0.f; }
    float mA = -INFINITY, mB = -INFINITY, lA = 0.f, lB = 0.f;   // l: this lane's share of the row sum
    uint32_t bitsA = 0, bitsB = 0;
    // the Philox block this lane draws per 16 keys: row (t odd ? rB : rA), tile (t >> 1)
    const uint32_t my_row = static_cast<uint32_t>(h * S + min((t & 1) ? rB : rA, S - 1));
    for (int jj = 0; jj < tiles; ++jj) {
      float c[2][4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        const int j0 = jj * 16 + half * 8;
        c[half][0] = c[half][1] = c[half][2] = c[half][3] = 0.f;
#pragma unroll
        for (int kk = 0; kk < KK; ++kk) {
          const int c0 = h * DH + kk * 16 + 2 * t;
          const uint32_t bh0 = lds_u32(k_hi + (j0 + g) * LDR + c0), bh1 = lds_u32(k_hi + (j0 + g) * LDR + c0 + 8);
          const uint32_t bl0 = lds_u32(k_lo + (j0 + g) * LDR + c0), bl1 = lds_u32(k_lo + (j0 + g) * LDR + c0 + 8);
          mma_bf16_16816(c[half], aql[kk], bh0, bh1);
          mma_bf16_16816(c[half], aqh[kk], bl0, bl1);
          mma_bf16_16816(c[half], aqh[kk], bh0, bh1);
        }
      }
      if (jj * 16 + 16 > S) {   // keys beyond S in the last block (warp-uniform)
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          const int key = jj * 16 + half * 8 + 2 * t;
          if (key >= S) { c[half][0] = -INFINITY; c[half][2] = -INFINITY; }
          if (key + 1 >= S) { c[half][1] = -INFINITY; c[half][3] = -INFINITY; }
        }
      }
      // online soft-max: rows rA (elements 0, 1) and rB (elements 2, 3) of both tiles
      float bmA = fmaxf(fmaxf(c[0][0], c[0][1]), fmaxf(c[1][0], c[1][1]));
      float bmB = fmaxf(fmaxf(c[0][2], c[0][3]), fmaxf(c[1][2], c[1][3]));
      bmA = fmaxf(bmA, __shfl_xor_sync(0xffffffffu, bmA, 1));
      bmB = fmaxf(bmB, __shfl_xor_sync(0xffffffffu, bmB, 1));
      bmA = fmaxf(bmA, __shfl_xor_sync(0xffffffffu, bmA, 2));
      bmB = fmaxf(bmB, __shfl_xor_sync(0xffffffffu, bmB, 2));
      bmA = fmaxf(bmA, mA);
      bmB = fmaxf(bmB, mB);
      const float corrA = ex2(mA - bmA), corrB = ex2(mB - bmB);   // first block: ex2(-inf) = 0
      mA = bmA; mB = bmB;
      lA *= corrA; lB *= corrB;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) { o[nt][0] *= corrA; o[nt][1] *= corrA; o[nt][2] *= corrB; o[nt][3] *= corrB; }
      float p[2][4];
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        p[half][0] = ex2(c[half][0] - mA); p[half][1] = ex2(c[half][1] - mA);
        p[half][2] = ex2(c[half][2] - mB); p[half][3] = ex2(c[half][3] - mB);
        lA += p[half][0] + p[half][1];     // the denominator counts dropped keys too
        lB += p[half][2] + p[half][3];
      }
      if (dr.mode == 1) {
        const uint32_t mine = keep_bits8(rng.block(1u | (my_row << 2), static_cast<uint32_t>(2 * jj + (t >> 1))), dr.thr_a);
        const int quad = lane & ~3;
        const uint32_t kA0 = __shfl_sync(0xffffffffu, mine, quad) >> (2 * t);       // row A, first tile
        const uint32_t kB0 = __shfl_sync(0xffffffffu, mine, quad + 1) >> (2 * t);   // row B, first tile
        const uint32_t kA1 = __shfl_sync(0xffffffffu, mine, quad + 2) >> (2 * t);   // row A, second tile
        const uint32_t kB1 = __shfl_sync(0xffffffffu, mine, quad + 3) >> (2 * t);
        if (!(kA0 & 1u)) p[0][0] = 0.f;
        if (!(kA0 & 2u)) p[0][1] = 0.f;
        if (!(kB0 & 1u)) p[0][2] = 0.f;
        if (!(kB0 & 2u)) p[0][3] = 0.f;
        if (!(kA1 & 1u)) p[1][0] = 0.f;
        if (!(kA1 & 2u)) p[1][1] = 0.f;
        if (!(kB1 & 1u)) p[1][2] = 0.f;
        if (!(kB1 & 2u)) p[1][3] = 0.f;
        const int sh = (jj & 1) * 16 + 2 * t;
        bitsA |= ((kA0 & 3u) << sh) | ((kA1 & 3u) << (sh + 8));
        bitsB |= ((kB0 & 3u) << sh) | ((kB1 & 3u) << (sh + 8));
      } else {
        const int sh = (jj & 1) * 16 + 2 * t;
        bitsA |= (3u << sh) | (3u << (sh + 8));
        bitsB |= (3u << sh) | (3u << (sh + 8));
      }
      if ((jj & 1) == 1 || jj == tiles - 1) {   // a 32-key word of keep bits is complete
        bitsA |= __shfl_xor_sync(0xffffffffu, bitsA, 1); bitsB |= __shfl_xor_sync(0xffffffffu, bitsB, 1);
        bitsA |= __shfl_xor_sync(0xffffffffu, bitsA, 2); bitsB |= __shfl_xor_sync(0xffffffffu, bitsB, 2);
        if (abits != nullptr && t == 0) {
          if (rA < S) abits[((static_cast<long long>(b) * H + h) * S + rA) * 4 + (jj >> 1)] = bitsA;
          if (rB < S) abits[((static_cast<long long>(b) * H + h) * S + rB) * 4 + (jj >> 1)] = bitsB;
        }
        bitsA = 0; bitsB = 0;
      }
      // context: the two score tiles are one k-step of 16 keys; P_kept as split bf16
      uint32_t ah[4], al[4];
      {
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(p[0][0], p[0][1]), h1 = __floats2bfloat162_rn(p[0][2], p[0][3]);
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(p[1][0], p[1][1]), h3 = __floats2bfloat162_rn(p[1][2], p[1][3]);
        ah[0] = *reinterpret_cast<const uint32_t*>(&h0); ah[1] = *reinterpret_cast<const uint32_t*>(&h1);
        ah[2] = *reinterpret_cast<const uint32_t*>(&h2); ah[3] = *reinterpret_cast<const uint32_t*>(&h3);
        al[0] = pack2(p[0][0] - __low2float(h0), p[0][1] - __high2float(h0));
        al[1] = pack2(p[0][2] - __low2float(h1), p[0][3] - __high2float(h1));
        al[2] = pack2(p[1][0] - __low2float(h2), p[1][1] - __high2float(h2));
        al[3] = pack2(p[1][2] - __low2float(h3), p[1][3] - __high2float(h3));
      }
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int n0 = h * DH + nt * 8 + g, k0 = jj * 16 + 2 * t;
        const uint32_t bh0 = lds_u32(vT_hi + n0 * LDT + k0), bh1 = lds_u32(vT_hi + n0 * LDT + k0 + 8);
        const uint32_t bl0 = lds_u32(vT_lo + n0 * LDT + k0), bl1 = lds_u32(vT_lo + n0 * LDT + k0 + 8);
        mma_bf16_16816(o[nt], al, bh0, bh1);
        mma_bf16_16816(o[nt], ah, bl0, bl1);
        mma_bf16_16816(o[nt], ah, bh0, bh1);
      }
    }
    lA += __shfl_xor_sync(0xffffffffu, lA, 1); lB += __shfl_xor_sync(0xffffffffu, lB, 1);
    lA += __shfl_xor_sync(0xffffffffu, lA, 2); lB += __shfl_xor_sync(0xffffffffu, lB, 2);
    const float linvA = 1.f / lA, linvB = 1.f / lB;
    const float scA = dr.inv_a * linvA, scB = dr.inv_a * linvB;
    auto put = [&](int r, int col, float x0, float x1) {   // split bf16 row [hi | lo | hi]
      __nv_bfloat16* out = ctx16 + (static_cast<long long>(b) * S + r) * 3 * E + col;
      const __nv_bfloat162 hi = __floats2bfloat162_rn(x0, x1);
      const uint32_t hiw = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint32_t*>(out) = hiw;
      *reinterpret_cast<uint32_t*>(out + E) = pack2(x0 - __low2float(hi), x1 - __high2float(hi));
      *reinterpret_cast<uint32_t*>(out + 2 * E) = hiw;
    };
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int col = h * DH + nt * 8 + 2 * t;
      if (rA < S) put(rA, col, o[nt][0] * scA, o[nt][1] * scA);
      if (rB < S) put(rB, col, o[nt][2] * scB, o[nt][3] * scB);
    }
    if (stat != nullptr && t == 0) {
      if (rA < S) stat[(static_cast<long long>(b) * H + h) * S + rA] = make_float2(mA, linvA);
      if (rB < S) stat[(static_cast<long long>(b) * H + h) * S + rB] = make_float2(mB, linvB);
    }
  }
}

// ---------------------------------------------------------------------------- embedding backward
// de = d(residual) + d(through q|k|v); d(Pos)[s] = sum_b de; d(Emb)[tok] += keep ? de/(1-p) : 0.
// CTA c handles samples c, c + grid, ...; thread owns a fixed set of (position, channel) pairs, so
// its d(Pos) partials live in registers and reach memory once (per-CTA partial rows, summed in a
// fixed order afterwards). The embedding rows accumulate in a shared-memory table per CTA when the
// vocabulary fits (<= 128 KB), else directly in global memory; both with float atomics.
__global__ void __launch_bounds__(512)
wide_embed_bwd_kernel(WideDims d, WideDrop dr, const long long* __restrict__ tokens, long long stride,
                      const float* __restrict__ dr32, const float* __restrict__ de32,
                      float* __restrict__ pos_partials /* [grid][S*E] */, float* __restrict__ emb_partials
                      /* [grid][vocab*E] or nullptr */, float* __restrict__ demb_global) {
  extern __shared__ __align__(16) float hist[];
  const int S = d.S, E = d.E;
  const int n = S * E;
  const bool use_hist = emb_partials != nullptr;
  if (use_hist)
    for (int i = threadIdx.x; i < d.vocab * E; i += blockDim.x) hist[i] = 0.f;
  __syncthreads();
  constexpr int kMaxPer = 32;                      // (S*E) / 512 <= 32  (S*E <= 16384)
  float pacc[kMaxPer];
#pragma unroll
  for (int u = 0; u < kMaxPer; ++u) pacc[u] = 0.f;
  for (int b = blockIdx.x; b < d.B; b += gridDim.x) {
    const Rng rng = make_rng(dr, b);
    const float* a = dr32 + static_cast<long long>(b) * n;
    const float* c = de32 + static_cast<long long>(b) * n;
#pragma unroll
    for (int u = 0; u < kMaxPer; ++u) {
      const int i = threadIdx.x + 512 * u;
      if (i < n) {
        const float v = a[i] + c[i];
        pacc[u] += v;
        const int s = i / E, ch = i % E;
        bool keep = true;
        if (dr.mode == 1) {
          const uint4 r = rng.block(0u, static_cast<uint32_t>(i >> 3));
          const uint32_t word = word_of(r, (i & 7) >> 1);
          keep = ((i & 1) ? (word >> 16) : (word & 0xFFFFu)) >= dr.thr_e;
        }
        if (keep) {
          long long t = tokens[b * stride + s];
          if (t < 0 || t >= d.vocab) t = 0;
          const float g = v * dr.inv_e;
          if (use_hist) atomicAdd(&hist[t * E + ch], g);
          else atomicAdd(&demb_global[t * E + ch], g);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kMaxPer; ++u) {
    const int i = threadIdx.x + 512 * u;
    if (i < n) pos_partials[static_cast<long long>(blockIdx.x) * n + i] = pacc[u];
  }
  __syncthreads();
  if (use_hist)
    for (int i = threadIdx.x; i < d.vocab * E; i += blockDim.x)
      emb_partials[static_cast<long long>(blockIdx.x) * d.vocab * E + i] = hist[i];
}

// Column sums of a bf16 matrix [rows, width] (bias gradients: d(in_proj / out_proj / fc1 bias) =
// sum over all token rows). CTA c sums rows c, c + grid, ... for every column (a thread owns a pair
// of adjacent columns) and leaves one partial row; wide_colsum_partials_kernel adds them in order.
__global__ void __launch_bounds__(256)
wide_colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, long long rows, int width, float* __restrict__ partials) {
  const int pairs = width / 2;
  const int rows_per_iter = 256 / pairs > 0 ? 256 / pairs : 1;      // several rows at once when narrow
  const int lane_pair = threadIdx.x % pairs, lane_row = threadIdx.x / pairs;
  __shared__ float2 red[256];
  float2 acc = make_float2(0.f, 0.f);
  if (threadIdx.x < pairs * rows_per_iter) {
    for (long long r = static_cast<long long>(blockIdx.x) * rows_per_iter + lane_row; r < rows;
         r += static_cast<long long>(gridDim.x) * rows_per_iter) {
      const float2 v = __bfloat1622float2(reinterpret_cast<const __nv_bfloat162*>(x + r * width)[lane_pair]);
      acc.x += v.x; acc.y += v.y;
    }
  }
  red[threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.x < pairs) {
    float2 s = red[threadIdx.x];
    for (int k = 1; k < rows_per_iter; ++k) { s.x += red[threadIdx.x + k * pairs].x; s.y += red[threadIdx.x + k * pairs].y; }
    partials[static_cast<long long>(blockIdx.x) * width + 2 * threadIdx.x] = s.x;
    partials[static_cast<long long>(blockIdx.x) * width + 2 * threadIdx.x + 1] = s.y;
  }
}

// Embedding backward, column-owner form (vocabulary table in shared memory): thread = channel c,
// CTA = samples c, c + grid, ...; the thread walks the positions of a sample in order and adds into
// ITS column of the per-CTA tables d(Pos)[s][c] and d(Emb)[tok[s]][c] -- no two threads share an
// address, so plain read-modify-write, a fixed summation order, no atomics (the atomic version
// above spent 1.0 ms per 8192-glyph step on shared-memory atomics: ~2 cycles per lane). Keep bits
// come from the bytes the forward saved (one per 8 channels).
__global__ void __launch_bounds__(256)
wide_embed_bwd_cols_kernel(WideDims d, float inv_e, const long long* __restrict__ tokens, long long stride,
                           const float* __restrict__ dr32, const float* __restrict__ de32,
                           const uint8_t* __restrict__ ebits, float* __restrict__ pos_partials,
                           float* __restrict__ emb_partials) {
  extern __shared__ __align__(16) float sm[];
  const int S = d.S, E = d.E, c = threadIdx.x;
  float* hist = sm;                          // [vocab][E]
  float* posacc = sm + d.vocab * E;          // [S][E]
  int* stok = reinterpret_cast<int*>(posacc + S * E);
  for (int i = c; i < d.vocab * E + S * E; i += E) sm[i] = 0.f;
  __syncthreads();
  for (int b = blockIdx.x; b < d.B; b += gridDim.x) {
    for (int s = c; s < S; s += E) {
      long long t = tokens[b * stride + s];
      stok[s] = (t < 0 || t >= d.vocab) ? 0 : static_cast<int>(t);
    }
    __syncthreads();
    const float* a = dr32 + static_cast<long long>(b) * S * E + c;
    const float* g = de32 + static_cast<long long>(b) * S * E + c;
    const uint8_t* kb = ebits != nullptr ? ebits + static_cast<long long>(b) * S * (E / 8) + (c >> 3) : nullptr;
    const int bit = c & 7;
    int s = 0;
    for (; s + 8 <= S; s += 8) {
      float v[8];
      uint32_t keep[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        v[u] = a[(s + u) * E] + g[(s + u) * E];
        keep[u] = kb != nullptr ? (kb[(s + u) * (E / 8)] >> bit) & 1u : 1u;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        posacc[(s + u) * E + c] += v[u];
        if (keep[u]) hist[stok[s + u] * E + c] += v[u] * inv_e;
      }
    }
    for (; s < S; ++s) {
      const float v = a[s * E] + g[s * E];
      const uint32_t keep = kb != nullptr ? (kb[s * (E / 8)] >> bit) & 1u : 1u;
      posacc[s * E + c] += v;
      if (keep) hist[stok[s] * E + c] += v * inv_e;
    }
    __syncthreads();
  }
  for (int i = c; i < S * E; i += E) pos_partials[static_cast<long long>(blockIdx.x) * S * E + i] = posacc[i];
  for (int i = c; i < d.vocab * E; i += E)
    emb_partials[static_cast<long long>(blockIdx.x) * d.vocab * E + i] = hist[i];
}

// out[m][n] = sum over the K pieces of partials[ks][m][n]; a piece has rows_pad >= M rows (split-K
// weight gradients, fixed summation order)
__global__ void __launch_bounds__(256)
wide_splitk_reduce_kernel(const float* __restrict__ partials, int splits, int M, int N, int rows_pad,
                          float* __restrict__ out) {
  const long long i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= static_cast<long long>(M) * N) return;
  const long long piece = static_cast<long long>(rows_pad) * N;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += partials[k * piece + i];
  out[i] = s;
}

int grid_for(long long work, int per_block, int num_sms, int waves = 8) {
  long long blocks = (work + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms) * waves;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

template <typename K>
cudaError_t set_smem(K kern, size_t bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes));
  return e;
}

}  // namespace

bool wide_shape_supported(int E, int H, int F, int L, const char** why) {
  const int dh = H > 0 ? E / H : 0;
  if (E < 32 || E > 256 || (E % 32) != 0) { *why = "embed_dim must be a multiple of 32 in [32, 256]"; return false; }
  if (H < 1 || E % H != 0 || (dh != 8 && dh != 16 && dh != 32)) { *why = "embed_dim / num_heads must be 8, 16 or 32"; return false; }
  if (F < 32 || (F % 32) != 0) { *why = "hidden (fc1 width) must be a multiple of 32"; return false; }
  if (L > 128) { *why = "max_length <= 128"; return false; }
  if (static_cast<long long>(L) * E > 16384) { *why = "max_length * embed_dim <= 16384"; return false; }
  if ((4ll * L * E + static_cast<long long>(L) * H) * 4 > 220 * 1024) { *why = "max_length * embed_dim too large for the attention backward (4 x L x E floats of shared memory)"; return false; }
  if (static_cast<long long>(H) * ((L + 31) & ~31) > 1024 * 4) { *why = "too many (head, position) pairs"; return false; }
  return true;
}

cudaError_t launch_wide_embed(const WideDims& d, const WideDrop& dr, const long long* tokens, long long stride,
                              const float* emb, const float* pos, float* e32, __nv_bfloat16* e16, uint8_t* ebits,
                              int* err_flag, int num_sms, cudaStream_t st) {
  const long long work = static_cast<long long>(d.B) * d.S * (d.E / 8);
  wide_embed_kernel<<<grid_for(work, 256, num_sms), 256, 0, st>>>(d, dr, tokens, stride, emb, pos, e32, e16, ebits,
                                                                err_flag);
  return cudaGetLastError();
}

cudaError_t launch_wide_attention_fwd(const WideDims& d, const WideDrop& dr, const float* qkv,
                                      __nv_bfloat16* ctx16, float2* stat, uint32_t* abits, cudaStream_t st) {
  const size_t smem = static_cast<size_t>(2) * (d.S + 8) * d.E * 4;
  int threads = d.H * ((d.S + 31) & ~31);
  if (threads > 512) threads = 512;      // 128 registers per thread: q / acc rows stay in registers
  cudaError_t e;
  static const bool no_mma = std::getenv("AFR_WIDE_ATTN_SIMT") != nullptr;      // diagnostic: force the SIMT kernel
  const AttnFwdSmem ml = attn_fwd_smem(d.S, d.E);
  if (!no_mma && (d.dh == 16 || d.dh == 32) && ml.total <= 220 * 1024) {
    if (d.dh == 16) {
      if ((e = set_smem(wide_attention_fwd_mma_kernel<16>, ml.total)) != cudaSuccess) return e;
      wide_attention_fwd_mma_kernel<16><<<d.B, 512, ml.total, st>>>(d, dr, qkv, ctx16, stat, abits);
    } else {
      if ((e = set_smem(wide_attention_fwd_mma_kernel<32>, ml.total)) != cudaSuccess) return e;
      wide_attention_fwd_mma_kernel<32><<<d.B, 512, ml.total, st>>>(d, dr, qkv, ctx16, stat, abits);
    }
    return cudaGetLastError();
  }
  switch (d.dh) {
    case 8:
      if ((e = set_smem(wide_attention_fwd_kernel<8>, smem)) != cudaSuccess) return e;
      wide_attention_fwd_kernel<8><<<d.B, threads, smem, st>>>(d, dr, qkv, ctx16, stat, abits);
      break;
    case 16:
      if ((e = set_smem(wide_attention_fwd_kernel<16>, smem)) != cudaSuccess) return e;
      wide_attention_fwd_kernel<16><<<d.B, threads, smem, st>>>(d, dr, qkv, ctx16, stat, abits);
      break;
    case 32:
      if ((e = set_smem(wide_attention_fwd_kernel<32>, smem)) != cudaSuccess) return e;
      wide_attention_fwd_kernel<32><<<d.B, threads, smem, st>>>(d, dr, qkv, ctx16, stat, abits);
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_wide_attention_bwd(const WideDims& d, const WideDrop& dr, const float* qkv, const float* dctx,
                                      const __nv_bfloat16* ctx16, const float2* stat, const uint32_t* abits,
                                      __nv_bfloat16* dqkv16, cudaStream_t st) {
  cudaError_t e;
  static const bool no_mma = std::getenv("AFR_WIDE_ATTN_SIMT") != nullptr;      // diagnostic: force the SIMT kernel
  const AttnMmaSmem ml = attn_mma_smem(d.S, d.E, d.H);
  if (!no_mma && (d.dh == 16 || d.dh == 32) && ml.total <= 220 * 1024) {
    if (d.dh == 16) {
      if ((e = set_smem(wide_attention_bwd_mma_kernel<16>, ml.total)) != cudaSuccess) return e;
      wide_attention_bwd_mma_kernel<16><<<d.B, 512, ml.total, st>>>(d, dr, qkv, dctx, ctx16, stat, abits, dqkv16);
    } else {
      if ((e = set_smem(wide_attention_bwd_mma_kernel<32>, ml.total)) != cudaSuccess) return e;
      wide_attention_bwd_mma_kernel<32><<<d.B, 512, ml.total, st>>>(d, dr, qkv, dctx, ctx16, stat, abits, dqkv16);
    }
    return cudaGetLastError();
  }
  const size_t smem = (static_cast<size_t>(4) * d.S * d.E + static_cast<size_t>(d.S) * d.H) * 4;
  int threads = d.H * ((d.S + 31) & ~31);
  if (threads > 512) threads = 512;
  switch (d.dh) {
    case 8:
      if ((e = set_smem(wide_attention_bwd_kernel<8>, smem)) != cudaSuccess) return e;
      wide_attention_bwd_kernel<8><<<d.B, threads, smem, st>>>(d, dr, qkv, dctx, ctx16, stat, abits, dqkv16);
      break;
    case 16:
      if ((e = set_smem(wide_attention_bwd_kernel<16>, smem)) != cudaSuccess) return e;
      wide_attention_bwd_kernel<16><<<d.B, threads, smem, st>>>(d, dr, qkv, dctx, ctx16, stat, abits, dqkv16);
      break;
    case 32:
      if ((e = set_smem(wide_attention_bwd_kernel<32>, smem)) != cudaSuccess) return e;
      wide_attention_bwd_kernel<32><<<d.B, threads, smem, st>>>(d, dr, qkv, dctx, ctx16, stat, abits, dqkv16);
      break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

#define AFR_NPL_DISPATCH(E, CALL)             \
  switch ((E) / 32) {                         \
    case 1: { constexpr int NPL = 1; CALL; break; } \
    case 2: { constexpr int NPL = 2; CALL; break; } \
    case 3: { constexpr int NPL = 3; CALL; break; } \
    case 4: { constexpr int NPL = 4; CALL; break; } \
    case 5: { constexpr int NPL = 5; CALL; break; } \
    case 6: { constexpr int NPL = 6; CALL; break; } \
    case 7: { constexpr int NPL = 7; CALL; break; } \
    case 8: { constexpr int NPL = 8; CALL; break; } \
    default: return cudaErrorInvalidValue;    \
  }

cudaError_t launch_wide_ln_fwd(long long rows, int E, const float* e32, const float* a32, const float* gamma,
                               const float* beta, float* xhat, float* rstd, __nv_bfloat16* h16, int num_sms,
                               cudaStream_t st) {
  const int grid = grid_for(rows, 8, num_sms, 16);
  AFR_NPL_DISPATCH(E, (wide_ln_fwd_kernel<NPL><<<grid, 256, 0, st>>>(rows, e32, a32, gamma, beta, xhat, rstd, h16)));
  return cudaGetLastError();
}

cudaError_t launch_wide_ln_bwd(long long rows, int E, const float* dh32, const float* xhat, const float* rstd,
                               const float* gamma, float* dr32, __nv_bfloat16* dr16, float* partials,
                               int max_partials, float* dgamma, float* dbeta, int num_sms, cudaStream_t st) {
  int grid = grid_for(rows, 8, num_sms, 4);
  if (grid > max_partials) grid = max_partials;
  AFR_NPL_DISPATCH(E, (wide_ln_bwd_kernel<NPL><<<grid, 256, 0, st>>>(rows, dh32, xhat, rstd, gamma, dr32, dr16, partials)));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  wide_colsum_partials_kernel<<<(2 * E + 255) / 256, 256, 0, st>>>(partials, grid, 2 * E, dgamma, E, dbeta);
  return cudaGetLastError();
}

cudaError_t launch_wide_act_fwd(const WideDims& d, const WideDrop& dr, const float* f32, __nv_bfloat16* feats,
                                float* feats_f32, int num_sms, cudaStream_t st) {
  const long long KF = static_cast<long long>(d.L) * d.F;
  if (d.S < d.L) {     // zero features for positions >= S (model.py:190-193)
    cudaError_t e = cudaMemset2DAsync(feats + static_cast<long long>(d.S) * d.F, KF * 2, 0,
                                      static_cast<size_t>(d.L - d.S) * d.F * 2, d.B, st);
    if (e != cudaSuccess) return e;
    if (feats_f32 != nullptr) {
      e = cudaMemset2DAsync(feats_f32 + static_cast<long long>(d.S) * d.F, KF * 4, 0,
                            static_cast<size_t>(d.L - d.S) * d.F * 4, d.B, st);
      if (e != cudaSuccess) return e;
    }
  }
  const long long work = static_cast<long long>(d.B) * d.S * (d.F / 8);
  wide_act_fwd_kernel<<<grid_for(work, 256, num_sms), 256, 0, st>>>(d, dr, f32, feats, feats_f32);
  return cudaGetLastError();
}

cudaError_t launch_wide_act_bwd(const WideDims& d, const WideDrop& dr, const float* f32, const float* dfeat,
                                __nv_bfloat16* df16, int num_sms, cudaStream_t st) {
  const long long work = static_cast<long long>(d.B) * d.S * (d.F / 8);
  wide_act_bwd_kernel<<<grid_for(work, 256, num_sms), 256, 0, st>>>(d, dr, f32, dfeat, df16);
  return cudaGetLastError();
}

cudaError_t launch_wide_embed_bwd(const WideDims& d, const WideDrop& dr, const long long* tokens, long long stride,
                                  const float* dr32, const float* de32, const uint8_t* ebits, float* pos_partials,
                                  float* emb_partials, int max_partials, float* dpos, float* demb, int num_sms,
                                  cudaStream_t st) {
  const int n = d.S * d.E;
  const long long emb_elems = static_cast<long long>(d.vocab) * d.E;
  const size_t cols_smem = (static_cast<size_t>(emb_elems) + n) * 4 + static_cast<size_t>(d.S) * 4;
  if (emb_partials != nullptr && cols_smem <= 110 * 1024 && (dr.mode == 0 || ebits != nullptr)) {
    int grid = 2 * num_sms < d.B ? 2 * num_sms : d.B;
    if (grid > max_partials) grid = max_partials;
    cudaError_t e = set_smem(wide_embed_bwd_cols_kernel, cols_smem);
    if (e != cudaSuccess) return e;
    wide_embed_bwd_cols_kernel<<<grid, d.E, cols_smem, st>>>(d, dr.inv_e, tokens, stride, dr32, de32,
                                                          dr.mode == 1 ? ebits : nullptr, pos_partials, emb_partials);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    wide_colsum_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos_partials, grid, n, dpos, n, nullptr);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (d.S < d.L) {
      e = cudaMemsetAsync(dpos + n, 0, static_cast<size_t>(d.L - d.S) * d.E * 4, st);
      if (e != cudaSuccess) return e;
    }
    wide_colsum_partials_kernel<<<static_cast<int>((emb_elems + 255) / 256), 256, 0, st>>>(
        emb_partials, grid, static_cast<int>(emb_elems), demb, static_cast<int>(emb_elems), nullptr);
    return cudaGetLastError();
  }
  int grid = num_sms < d.B ? num_sms : d.B;
  if (grid > max_partials) grid = max_partials;
  const bool use_hist = emb_partials != nullptr;
  const size_t smem = use_hist ? static_cast<size_t>(d.vocab) * d.E * 4 : 0;
  cudaError_t e = set_smem(wide_embed_bwd_kernel, smem > 0 ? smem : 1024);
  if (e != cudaSuccess) return e;
  const long long emb_n = static_cast<long long>(d.vocab) * d.E;
  if (!use_hist) {
    e = cudaMemsetAsync(demb, 0, emb_n * 4, st);
    if (e != cudaSuccess) return e;
  }
  wide_embed_bwd_kernel<<<grid, 512, smem, st>>>(d, dr, tokens, stride, dr32, de32, pos_partials,
                                                 emb_partials, demb);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  wide_colsum_partials_kernel<<<(n + 255) / 256, 256, 0, st>>>(pos_partials, grid, n, dpos, n, nullptr);
  if ((e = cudaGetLastError()) != cudaSuccess) return e;
  if (d.S < d.L) {      // positions the batch never reached get a zero gradient
    e = cudaMemsetAsync(dpos + n, 0, static_cast<size_t>(d.L - d.S) * d.E * 4, st);
    if (e != cudaSuccess) return e;
  }
  if (use_hist) {
    wide_colsum_partials_kernel<<<static_cast<int>((emb_n + 255) / 256), 256, 0, st>>>(
        emb_partials, grid, static_cast<int>(emb_n), demb, static_cast<int>(emb_n), nullptr);
    e = cudaGetLastError();
  }
  return e;
}

cudaError_t launch_wide_colsum_bf16(const __nv_bfloat16* x, long long rows, int width, float* partials,
                                    int max_partials, float* out, int num_sms, cudaStream_t st) {
  if (width < 2 || (width % 2) != 0 || width > 512) return cudaErrorInvalidValue;
  int grid = num_sms * 4;
  if (grid > max_partials) grid = max_partials;
  if (grid > rows) grid = static_cast<int>(rows);
  wide_colsum_bf16_kernel<<<grid, 256, 0, st>>>(x, rows, width, partials);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  wide_colsum_partials_kernel<<<(width + 255) / 256, 256, 0, st>>>(partials, grid, width, out, width, nullptr);
  return cudaGetLastError();
}

cudaError_t launch_wide_split_weight(const float* w, int rows, int E, __nv_bfloat16* out, cudaStream_t st) {
  wide_split_weight_kernel<<<(rows * E + 255) / 256, 256, 0, st>>>(w, rows, E, out);
  return cudaGetLastError();
}

cudaError_t launch_wide_splitk_reduce(const float* partials, int splits, int M, int N, float* out,
                                      cudaStream_t st) {
  const long long n = static_cast<long long>(M) * N;
  wide_splitk_reduce_kernel<<<static_cast<int>((n + 255) / 256), 256, 0, st>>>(partials, splits, M, N,
                                                                             (M + 127) / 128 * 128, out);
  return cudaGetLastError();
}

}  // namespace afr
