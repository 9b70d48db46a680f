// HBM-bound kernels of the training step: the fused AdamW sweep (reference: optim.AdamW at
// model.py:273, stepped at model.py:310), the bf16 shadow refresh, the loss finalisation
// (mean of model.py:270), the fc_output bias gradient and the clamp backward of the generic
// autograd path. All are 128-bit vectorised, coalesced, grid sized in multiples of the SM count.
#include <cstdlib>

#include "afr_internal.h"
#include "afr_ptx.cuh"

namespace afr {
namespace {

__device__ __forceinline__ float4 ld_stream(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// Grid-stride sweep, two independent 16-byte groups per thread and iteration (8 loads in flight
// per thread before the first use).
__device__ __forceinline__ void adamw_store(float* p, float* m, float* v, __nv_bfloat16* shadow,
                                            long long i, const float4& pv, const float4& mv,
                                            const float4& vv) {
  reinterpret_cast<float4*>(p)[i] = pv;
  reinterpret_cast<float4*>(m)[i] = mv;
  reinterpret_cast<float4*>(v)[i] = vv;
  if (shadow != nullptr) {
    const __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y);
    const __nv_bfloat162 hi = __floats2bfloat162_rn(pv.z, pv.w);
    uint2 packed;
    packed.x = *reinterpret_cast<const uint32_t*>(&lo);
    packed.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(shadow)[i] = packed;
  }
}

__global__ void __launch_bounds__(256)
adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
             float* __restrict__ v, long long n4, long long n, AdamHyper h,
             __nv_bfloat16* __restrict__ shadow) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const AdamPairConst c(h);
  for (; i + stride < n4; i += 2 * stride) {
    const long long j = i + stride;
    float4 pa = reinterpret_cast<float4*>(p)[i], pb = reinterpret_cast<float4*>(p)[j];
    const float4 ga = ld_stream(reinterpret_cast<const float4*>(g) + i);
    const float4 gb = ld_stream(reinterpret_cast<const float4*>(g) + j);
    float4 ma = reinterpret_cast<float4*>(m)[i], mb = reinterpret_cast<float4*>(m)[j];
    float4 va = reinterpret_cast<float4*>(v)[i], vb = reinterpret_cast<float4*>(v)[j];
    adamw_quad(pa, ga, ma, va, c);
    adamw_store(p, m, v, shadow, i, pa, ma, va);
    adamw_quad(pb, gb, mb, vb, c);
    adamw_store(p, m, v, shadow, j, pb, mb, vb);
  }
  if (i < n4) {
    float4 pv = reinterpret_cast<float4*>(p)[i];
    const float4 gv = ld_stream(reinterpret_cast<const float4*>(g) + i);
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    adamw_quad(pv, gv, mv, vv, c);
    adamw_store(p, m, v, shadow, i, pv, mv, vv);
  }
  // scalar tail (n not a multiple of 4)
  if (blockIdx.x == 0) {
    for (long long j = n4 * 4 + threadIdx.x; j < n; j += blockDim.x) {
      float pj = p[j], mj = m[j], vj = v[j];
      adamw_elem(pj, g[j], mj, vj, h);
      p[j] = pj; m[j] = mj; v[j] = vj;
      if (shadow != nullptr) shadow[j] = __float2bfloat16_rn(pj);
    }
  }
}

// Background form of the sweep (afr_adamw_rows_bg): a small-footprint persistent kernel meant to
// share every SM with the compute kernels of the step (the wgrad / dgrad GEMMs, the front-end
// backward), which leave few registers and little shared memory but plenty of issue slots and
// all of the HBM bandwidth. The memory-level parallelism therefore cannot live in registers (a
// thread of the plain sweep holds 8 x 16 B in flight): every thread streams its own 16-byte
// groups of p, g, m, v through a private slot of a shared-memory ring with cp.async (LDGSTS,
// L2 only), kStages - 1 groups of 4 x 16 B in flight per thread, and only does LDS.128 ->
// adamw_quad -> streaming stores. No barrier of any kind: a thread reads only what it copied
// itself (cp.async.wait_group). 128 threads, <= 40 registers (5 K per CTA), 8 KB per stage.
// (A first version filled the ring with 2 KB cp.async.bulk copies + mbarriers: a CTA then
// sustained only ~8 B/clk whatever the ring depth -- see profiles/r02_notes.md.)
constexpr int kRingThreads = 128;
constexpr int kRingSegFloats = kRingThreads * 4;          // one float4 per thread and array
constexpr int kRingStageBytes = 4 * kRingSegFloats * 4;   // p | g | m | v

// (no L2::cache_hint: LDGSTS faulted with "illegal instruction" on the hand-encoded policy words
// the TMA loads accept)
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_dst), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void stg_stream_f4(float* dst, const float4& v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void stg_stream_u2(void* dst, uint32_t a, uint32_t b) {
  asm volatile("st.global.cs.v2.b32 [%0], {%1, %2};" ::"l"(dst), "r"(a), "r"(b)
               : "memory");
}

struct RingArgs {
  float* p;
  const float* g;
  float* m;
  float* v;
  __nv_bfloat16* shadow;
  long long n;        // floats, multiple of 4
  long long nseg;     // ceil(n / kRingSegFloats)
  AdamHyper h;
};

template <int kStages>
__global__ void __launch_bounds__(kRingThreads, 12) adamw_ring_kernel(const RingArgs a) {
  extern __shared__ __align__(128) unsigned char ring_smem[];
  const int tid = threadIdx.x;
  const uint32_t slot = ptx::smem_u32(ring_smem) + 16u * tid;
  const long long first = blockIdx.x, step = gridDim.x;
  // one commit group per segment of this CTA's sequence (empty past the end: uniform counting)
  auto issue = [&](long long seg, int s) {
    const long long i = seg * kRingSegFloats + 4 * tid;
    if (seg < a.nseg && i < a.n) {
      const uint32_t dst = slot + static_cast<uint32_t>(s) * kRingStageBytes;
      cp_async_16(dst, a.p + i);
      cp_async_16(dst + 4 * kRingSegFloats, a.g + i);
      cp_async_16(dst + 8 * kRingSegFloats, a.m + i);
      cp_async_16(dst + 12 * kRingSegFloats, a.v + i);
    }
    cp_async_commit();
  };
  {
    long long seg = first;
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s, seg += step) issue(seg, s);
  }
  int s = 0;
  for (long long seg = first; seg < a.nseg; seg += step) {
    // refill the slot this thread read in the previous iteration (its values were consumed by
    // that iteration's arithmetic, so the loads have long returned)
    issue(seg + static_cast<long long>(kStages - 1) * step, s == 0 ? kStages - 1 : s - 1);
    cp_async_wait<kStages - 1>();
    const long long i = seg * kRingSegFloats + 4 * tid;
    if (i < a.n) {
      const uint32_t src = slot + static_cast<uint32_t>(s) * kRingStageBytes;
      float4 pv = ptx::lds_f4(src);
      const float4 gv = ptx::lds_f4(src + 4 * kRingSegFloats);
      float4 mv = ptx::lds_f4(src + 8 * kRingSegFloats);
      float4 vv = ptx::lds_f4(src + 12 * kRingSegFloats);
      adamw_quad(pv, gv, mv, vv, AdamPairConst(a.h));   // constants straight from the constant bank
      stg_stream_f4(a.p + i, pv);
      stg_stream_f4(a.m + i, mv);
      stg_stream_f4(a.v + i, vv);
      if (a.shadow != nullptr) {
        const __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y);
        const __nv_bfloat162 hi = __floats2bfloat162_rn(pv.z, pv.w);
        stg_stream_u2(a.shadow + i, *reinterpret_cast<const uint32_t*>(&lo),
                      *reinterpret_cast<const uint32_t*>(&hi));
      }
    }
    if (++s == kStages) s = 0;
  }
  cp_async_wait<0>();
}

// Row-sharded data parallel AdamW with the two collectives folded in (training.py): this rank
// owns `n4` float4 groups of fc_output.weight. The gradient of an owned element is the sum over
// ranks of the local gradients, read straight out of every peer's dW buffer over NVLink (the
// reduce-scatter); the updated weight goes, rounded to bf16, into the inactive shadow copy of
// EVERY rank (the all-gather). All buffers are symmetric-memory allocations; `g` / `sh` hold the
// peer-mapped pointers already offset to the owned rows, in rank order (own rank included). The
// summation order is the rank order on every rank, and every element has exactly one owner, so
// the result does not depend on timing. Launched on a handful of SMs (the rest keep running the
// GEMMs / front-end backward) with 1024 threads each: remote loads are latency-bound, every
// thread keeps world + 3 16-byte loads in flight.
constexpr int kMaxPeers = 8;
struct GatherPeers {
  const float* g[kMaxPeers];
  __nv_bfloat16* sh[kMaxPeers];
  int world;
};

__device__ __forceinline__ float4 ld_peer(const float* base, long long i) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(reinterpret_cast<const float4*>(base) + i));
  return r;
}

__global__ void __launch_bounds__(1024, 1)
adamw_gather_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                    long long n4, AdamHyper h, GatherPeers peers) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int W = peers.world;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 g[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < W) g[q] = ld_peer(peers.g[q], i);
    float4 pv = reinterpret_cast<float4*>(p)[i];
    float4 mv = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float4 s = g[0];
#pragma unroll
    for (int q = 1; q < kMaxPeers; ++q)
      if (q < W) { s.x += g[q].x; s.y += g[q].y; s.z += g[q].z; s.w += g[q].w; }
    adamw_quad(pv, s, mv, vv, AdamPairConst(h));
    reinterpret_cast<float4*>(p)[i] = pv;
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    const __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
    uint2 packed;
    packed.x = *reinterpret_cast<const uint32_t*>(&lo);
    packed.y = *reinterpret_cast<const uint32_t*>(&hi);
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < W) reinterpret_cast<uint2*>(peers.sh[q])[i] = packed;
  }
}

// The same step through the NVSwitch (NVLS multicast objects of the symmetric allocations): ONE
// multimem.ld_reduce returns the gradient already summed over all ranks inside the switch, ONE
// multimem.st delivers the bf16 result to every rank -- the kernel's NVLink traffic drops from
// (world - 1) x (16 + 8) bytes per float4 group to 16 + 8, so it is HBM-bound instead of link-bound.
__device__ __forceinline__ float4 ld_reduce_mc(const float* g_mc, long long i) {
  float4 s;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(s.x), "=f"(s.y), "=f"(s.z), "=f"(s.w)
               : "l"(reinterpret_cast<const float4*>(g_mc) + i)
               : "memory");
  return s;
}
__device__ __forceinline__ void adamw_group_nvls(float* p, float* m, float* v, long long i, const float4& s,
                                                 const AdamHyper& h, __nv_bfloat16* sh_mc) {
  float4 pv = reinterpret_cast<float4*>(p)[i];
  float4 mv = reinterpret_cast<float4*>(m)[i];
  float4 vv = reinterpret_cast<float4*>(v)[i];
  adamw_quad(pv, s, mv, vv, AdamPairConst(h));
  reinterpret_cast<float4*>(p)[i] = pv;
  reinterpret_cast<float4*>(m)[i] = mv;
  reinterpret_cast<float4*>(v)[i] = vv;
  const __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
  asm volatile("multimem.st.relaxed.sys.global.v2.bf16x2 [%0], {%1, %2};" ::"l"(
                   reinterpret_cast<uint2*>(sh_mc) + i),
               "r"(*reinterpret_cast<const uint32_t*>(&lo)), "r"(*reinterpret_cast<const uint32_t*>(&hi))
               : "memory");
}
// The in-switch reduction has a long round trip: every thread keeps four of them in flight.
__global__ void __launch_bounds__(1024, 1)
adamw_gather_nvls_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v,
                         long long n4, AdamHyper h, const float* g_mc, __nv_bfloat16* sh_mc) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n4; i += 4 * stride) {
    float4 s[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) s[u] = ld_reduce_mc(g_mc, i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) adamw_group_nvls(p, m, v, i + u * stride, s[u], h, sh_mc);
  }
  for (; i < n4; i += stride) adamw_group_nvls(p, m, v, i, ld_reduce_mc(g_mc, i), h, sh_mc);
}

// ---- bf16 gradients over NVLink: half the egress of the fp32 forms above. Groups of 8 elements.
__device__ __forceinline__ void unpack_bf16x8(const uint4& u, float4& lo, float4& hi) {
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.y));
  const float2 c = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.z));
  const float2 d = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u.w));
  lo = make_float4(a.x, a.y, b.x, b.y);
  hi = make_float4(c.x, c.y, d.x, d.y);
}
__device__ __forceinline__ uint4 adamw_group8(float* p, float* m, float* v, long long i8, const float4& glo,
                                              const float4& ghi, const AdamPairConst& c) {
  float4 p0 = reinterpret_cast<float4*>(p)[2 * i8], p1 = reinterpret_cast<float4*>(p)[2 * i8 + 1];
  float4 m0 = reinterpret_cast<float4*>(m)[2 * i8], m1 = reinterpret_cast<float4*>(m)[2 * i8 + 1];
  float4 v0 = reinterpret_cast<float4*>(v)[2 * i8], v1 = reinterpret_cast<float4*>(v)[2 * i8 + 1];
  adamw_quad(p0, glo, m0, v0, c);
  adamw_quad(p1, ghi, m1, v1, c);
  reinterpret_cast<float4*>(p)[2 * i8] = p0; reinterpret_cast<float4*>(p)[2 * i8 + 1] = p1;
  reinterpret_cast<float4*>(m)[2 * i8] = m0; reinterpret_cast<float4*>(m)[2 * i8 + 1] = m1;
  reinterpret_cast<float4*>(v)[2 * i8] = v0; reinterpret_cast<float4*>(v)[2 * i8 + 1] = v1;
  const __nv_bfloat162 a = __floats2bfloat162_rn(p0.x, p0.y), b = __floats2bfloat162_rn(p0.z, p0.w);
  const __nv_bfloat162 cc = __floats2bfloat162_rn(p1.x, p1.y), d = __floats2bfloat162_rn(p1.z, p1.w);
  return make_uint4(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b),
                    *reinterpret_cast<const uint32_t*>(&cc), *reinterpret_cast<const uint32_t*>(&d));
}
struct GatherPeers16 {
  const __nv_bfloat16* g[kMaxPeers];
  __nv_bfloat16* sh[kMaxPeers];
  int world;
};
__global__ void __launch_bounds__(1024, 1)
adamw_gather_bf16_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, long long n8,
                         AdamHyper h, GatherPeers16 peers) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const int W = peers.world;
  const AdamPairConst c(h);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8; i += stride) {
    uint4 g[kMaxPeers];
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < W)
        asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                     : "=r"(g[q].x), "=r"(g[q].y), "=r"(g[q].z), "=r"(g[q].w)
                     : "l"(reinterpret_cast<const uint4*>(peers.g[q]) + i));
    float4 slo, shi;
    unpack_bf16x8(g[0], slo, shi);
#pragma unroll
    for (int q = 1; q < kMaxPeers; ++q)
      if (q < W) {
        float4 lo, hi;
        unpack_bf16x8(g[q], lo, hi);
        slo.x += lo.x; slo.y += lo.y; slo.z += lo.z; slo.w += lo.w;
        shi.x += hi.x; shi.y += hi.y; shi.z += hi.z; shi.w += hi.w;
      }
    const uint4 packed = adamw_group8(p, m, v, i, slo, shi, c);
#pragma unroll
    for (int q = 0; q < kMaxPeers; ++q)
      if (q < W) reinterpret_cast<uint4*>(peers.sh[q])[i] = packed;
  }
}
// in-switch sum of the bf16 gradients with fp32 accumulation
__device__ __forceinline__ uint4 ld_reduce_mc_bf16(const __nv_bfloat16* g_mc, long long i8) {
  uint4 s;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(s.x), "=r"(s.y), "=r"(s.z), "=r"(s.w)
               : "l"(reinterpret_cast<const uint4*>(g_mc) + i8)
               : "memory");
  return s;
}
constexpr int kGatherNvlsThreads = 512;   // 128 registers per thread: 16 reductions of 16 B in flight each
__global__ void __launch_bounds__(kGatherNvlsThreads, 1)
adamw_gather_nvls_bf16_kernel(float* __restrict__ p, float* __restrict__ m, float* __restrict__ v, long long n8,
                              AdamHyper h, const __nv_bfloat16* g_mc, __nv_bfloat16* sh_mc) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  const AdamPairConst c(h);
  long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  auto one = [&](long long j, const uint4& s) {
    float4 lo, hi;
    unpack_bf16x8(s, lo, hi);
    const uint4 packed = adamw_group8(p, m, v, j, lo, hi, c);
    asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(
                     reinterpret_cast<uint4*>(sh_mc) + j),
                 "r"(packed.x), "r"(packed.y), "r"(packed.z), "r"(packed.w)
                 : "memory");
  };
  // 16 in-switch reductions (16 B each) in flight per thread: the round trip through the NVSwitch is
  // ~2.5 us, so a CTA's link throughput is its bytes in flight over that (12 CTAs x 1024 threads x
  // 64 B = 0.79 MB gave 336 GB/s at 8 GPUs: 0.64 ms for the exchange, the longest chain of the step)
  constexpr int kDepth = 16;
  for (; i + (kDepth - 1) * stride < n8; i += kDepth * stride) {
    uint4 s[kDepth];
#pragma unroll
    for (int u = 0; u < kDepth; ++u) s[u] = ld_reduce_mc_bf16(g_mc, i + u * stride);
#pragma unroll
    for (int u = 0; u < kDepth; ++u) one(i + u * stride, s[u]);
  }
  for (; i < n8; i += stride) one(i, ld_reduce_mc_bf16(g_mc, i));
}

constexpr int kMaxSmallJobs = 16;
struct SmallJobs { SmallAdamJob j[kMaxSmallJobs]; int n; };

__global__ void __launch_bounds__(256) adamw_small_kernel(SmallJobs jobs, AdamHyper h) {
  const SmallAdamJob job = jobs.j[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < job.n; i += gridDim.x * blockDim.x) {
    float pj = job.p[i], mj = job.m[i], vj = job.v[i];
    adamw_elem(pj, job.g[i], mj, vj, h);
    job.p[i] = pj; job.m[i] = mj; job.v[i] = vj;
  }
}

__global__ void __launch_bounds__(256)
f32_to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n4,
                   long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += stride) {
    const float4 x = ld_stream(reinterpret_cast<const float4*>(src) + i);
    const __nv_bfloat162 lo = __floats2bfloat162_rn(x.x, x.y);
    const __nv_bfloat162 hi = __floats2bfloat162_rn(x.z, x.w);
    uint2 packed;
    packed.x = *reinterpret_cast<const uint32_t*>(&lo);
    packed.y = *reinterpret_cast<const uint32_t*>(&hi);
    reinterpret_cast<uint2*>(dst)[i] = packed;
  }
  if (blockIdx.x == 0)
    for (long long j = n4 * 4 + threadIdx.x; j < n; j += blockDim.x)
      dst[j] = __float2bfloat16_rn(src[j]);
}

// One block; fixed summation order -> run-to-run deterministic loss.
__global__ void __launch_bounds__(256)
loss_finalize_kernel(const float* __restrict__ partials, int n, double count,
                     float* __restrict__ loss_out) {
  __shared__ double red[256];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) acc += static_cast<double>(partials[i]);
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *loss_out = static_cast<float>(red[0] / count);
}

// Stage 1: scratch[slice, p] = sum over the slice's rows of dz[b, p]  (bf16x2 per thread).
__global__ void __launch_bounds__(128)
bias_grad_stage1(const __nv_bfloat16* __restrict__ dz, int B, int P, long long ld,
                 int rows_per_slice, float* __restrict__ scratch) {
  const int pair = blockIdx.x * blockDim.x + threadIdx.x;
  if (pair * 2 >= P) return;
  const int b0 = blockIdx.y * rows_per_slice;
  const int b1 = min(B, b0 + rows_per_slice);
  float s0 = 0.f, s1 = 0.f;
  for (int b = b0; b < b1; ++b) {
    const __nv_bfloat162 x =
        reinterpret_cast<const __nv_bfloat162*>(dz + static_cast<long long>(b) * ld)[pair];
    const float2 f = __bfloat1622float2(x);
    s0 += f.x; s1 += f.y;
  }
  float* out = scratch + static_cast<long long>(blockIdx.y) * P;
  out[pair * 2] = s0;
  out[pair * 2 + 1] = s1;
}
__global__ void __launch_bounds__(256)
bias_grad_stage2(const float* __restrict__ scratch, int slices, int P, float alpha,
                 float* __restrict__ dbias) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  float s = 0.f;
  for (int k = 0; k < slices; ++k) s += scratch[static_cast<long long>(k) * P + p];
  dbias[p] = s * alpha;
}

__global__ void __launch_bounds__(256)
clamp_backward_kernel(const float* __restrict__ dy, const float* __restrict__ z,
                      __nv_bfloat16* __restrict__ dz, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += stride) {
    const float zz = z[i];
    dz[i] = __float2bfloat16_rn((zz >= 0.f && zz <= 1.f) ? dy[i] : 0.f);
  }
}

__global__ void __launch_bounds__(256)
clamp01_kernel(const float* __restrict__ z, float* __restrict__ y, long long n4, long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += stride) {
    float4 x = reinterpret_cast<const float4*>(z)[i];
    x.x = fminf(fmaxf(x.x, 0.f), 1.f); x.y = fminf(fmaxf(x.y, 0.f), 1.f);
    x.z = fminf(fmaxf(x.z, 0.f), 1.f); x.w = fminf(fmaxf(x.w, 0.f), 1.f);
    reinterpret_cast<float4*>(y)[i] = x;
  }
  if (blockIdx.x == 0)
    for (long long j = n4 * 4 + threadIdx.x; j < n; j += blockDim.x)
      y[j] = fminf(fmaxf(z[j], 0.f), 1.f);
}

__global__ void __launch_bounds__(256)
div_sqrt_check_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ q,
                      float* __restrict__ s, float* __restrict__ q_ieee, float* __restrict__ s_ieee,
                      long long n) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float x = a[i], y = b[i];
    q[i] = div_rn_nobranch(x, y);
    q_ieee[i] = __fdiv_rn(x, y);
    s[i] = sqrt_rn_nobranch(fabsf(x));
    s_ieee[i] = __fsqrt_rn(fabsf(x));
  }
}

int stream_grid(long long work_items, int threads, int num_sms) {
  long long blocks = (work_items + threads - 1) / threads;
  const long long cap = static_cast<long long>(num_sms) * 8;  // 8 resident CTAs of 256 thr / SM
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace

cudaError_t launch_adamw(float* p, const float* g, float* m, float* v, long long n,
                         const AdamHyper& h, __nv_bfloat16* shadow, int num_sms, cudaStream_t s) {
  const long long n4 = n / 4;
  static const int ctas_per_sm = [] {
    const char* e = std::getenv("AFR_ADAMW_CTAS");   // tuning knob; 8 x 256 threads per SM by default
    const int v = e ? std::atoi(e) : 0;
    return v >= 1 && v <= 8 ? v : 8;
  }();
  long long blocks = (n4 + 511) / 512;
  if (blocks > static_cast<long long>(num_sms) * ctas_per_sm) blocks = static_cast<long long>(num_sms) * ctas_per_sm;
  if (blocks < 1) blocks = 1;
  adamw_kernel<<<static_cast<int>(blocks), 256, 0, s>>>(p, g, m, v, n4, n, h, shadow);
  return cudaGetLastError();
}

template <int kStages>
cudaError_t launch_ring_impl(const RingArgs& a, int ctas, cudaStream_t s) {
  const size_t smem = static_cast<size_t>(kStages) * kRingStageBytes;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(adamw_ring_kernel<kStages>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(adamw_ring_kernel<kStages>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  adamw_ring_kernel<kStages><<<ctas, kRingThreads, smem, s>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_adamw_ring(float* p, const float* g, float* m, float* v, long long n,
                              const AdamHyper& h, __nv_bfloat16* shadow, int ctas, int stages,
                              cudaStream_t s) {
  if ((n % 4) != 0 || n < 4 || ctas < 1) return cudaErrorInvalidValue;
  RingArgs a{};
  a.p = p; a.g = g; a.m = m; a.v = v; a.shadow = shadow; a.n = n;
  a.nseg = (n + kRingSegFloats - 1) / kRingSegFloats;
  a.h = h;
  if (ctas > a.nseg) ctas = static_cast<int>(a.nseg);
  // ring depth: 2, 3, 4, 6, 8 or 12 stages of 8 KB (rounded down to the next one built)
  if (stages >= 12) return launch_ring_impl<12>(a, ctas, s);
  if (stages >= 8) return launch_ring_impl<8>(a, ctas, s);
  if (stages >= 6) return launch_ring_impl<6>(a, ctas, s);
  if (stages >= 4) return launch_ring_impl<4>(a, ctas, s);
  if (stages == 3) return launch_ring_impl<3>(a, ctas, s);
  return launch_ring_impl<2>(a, ctas, s);
}

cudaError_t launch_adamw_gather(float* p, float* m, float* v, long long n, const AdamHyper& h,
                                const float* const* peer_g, __nv_bfloat16* const* peer_shadow, int world,
                                int ctas, cudaStream_t s) {
  if (world < 1 || world > kMaxPeers || (n % 4) != 0 || ctas < 1) return cudaErrorInvalidValue;
  GatherPeers peers{};
  peers.world = world;
  for (int q = 0; q < world; ++q) {
    peers.g[q] = peer_g[q];
    peers.sh[q] = peer_shadow[q];
  }
  adamw_gather_kernel<<<ctas, 1024, 0, s>>>(p, m, v, n / 4, h, peers);
  return cudaGetLastError();
}

cudaError_t launch_adamw_gather_nvls(float* p, float* m, float* v, long long n, const AdamHyper& h,
                                     const float* g_mc, __nv_bfloat16* sh_mc, int ctas, cudaStream_t s) {
  if ((n % 4) != 0 || ctas < 1 || g_mc == nullptr || sh_mc == nullptr) return cudaErrorInvalidValue;
  adamw_gather_nvls_kernel<<<ctas, 1024, 0, s>>>(p, m, v, n / 4, h, g_mc, sh_mc);
  return cudaGetLastError();
}

cudaError_t launch_adamw_gather_bf16(float* p, float* m, float* v, long long n, const AdamHyper& h,
                                     const __nv_bfloat16* const* peer_g, __nv_bfloat16* const* peer_shadow,
                                     int world, int ctas, cudaStream_t s) {
  if (world < 1 || world > kMaxPeers || (n % 8) != 0 || ctas < 1) return cudaErrorInvalidValue;
  GatherPeers16 peers{};
  peers.world = world;
  for (int q = 0; q < world; ++q) { peers.g[q] = peer_g[q]; peers.sh[q] = peer_shadow[q]; }
  adamw_gather_bf16_kernel<<<ctas, 1024, 0, s>>>(p, m, v, n / 8, h, peers);
  return cudaGetLastError();
}

cudaError_t launch_adamw_gather_nvls_bf16(float* p, float* m, float* v, long long n, const AdamHyper& h,
                                          const __nv_bfloat16* g_mc, __nv_bfloat16* sh_mc, int ctas,
                                          cudaStream_t s) {
  if ((n % 8) != 0 || ctas < 1 || g_mc == nullptr || sh_mc == nullptr) return cudaErrorInvalidValue;
  adamw_gather_nvls_bf16_kernel<<<ctas, kGatherNvlsThreads, 0, s>>>(p, m, v, n / 8, h, g_mc, sh_mc);
  return cudaGetLastError();
}

cudaError_t launch_adamw_small(const SmallAdamJob* jobs, int njobs, const AdamHyper& h,
                               cudaStream_t s) {
  if (njobs <= 0 || njobs > kMaxSmallJobs) return cudaErrorInvalidValue;
  SmallJobs sj{};
  int max_n = 0;
  for (int i = 0; i < njobs; ++i) {
    sj.j[i] = jobs[i];
    if (jobs[i].n > max_n) max_n = jobs[i].n;
  }
  sj.n = njobs;
  int gx = (max_n + 255) / 256;
  if (gx > 32) gx = 32;
  adamw_small_kernel<<<dim3(gx, njobs), 256, 0, s>>>(sj, h);
  return cudaGetLastError();
}

cudaError_t launch_f32_to_bf16(const float* src, __nv_bfloat16* dst, long long n, cudaStream_t s) {
  const long long n4 = n / 4;
  f32_to_bf16_kernel<<<stream_grid(n4, 256, 148), 256, 0, s>>>(src, dst, n4, n);
  return cudaGetLastError();
}

cudaError_t launch_loss_finalize(const float* partials, int n, double count, float* loss_out,
                                 cudaStream_t s) {
  loss_finalize_kernel<<<1, 256, 0, s>>>(partials, n, count, loss_out);
  return cudaGetLastError();
}

cudaError_t launch_bias_grad(const __nv_bfloat16* dz, int B, int P, float alpha, float* scratch,
                             float* dbias, cudaStream_t s, long long ld) {
  int slices = (B + 31) / 32;
  if (slices > 32) slices = 32;
  const int rows = (B + slices - 1) / slices;
  slices = (B + rows - 1) / rows;
  const int pairs = (P + 1) / 2;
  bias_grad_stage1<<<dim3((pairs + 127) / 128, slices), 128, 0, s>>>(dz, B, P, ld, rows, scratch);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  bias_grad_stage2<<<(P + 255) / 256, 256, 0, s>>>(scratch, slices, P, alpha, dbias);
  return cudaGetLastError();
}

cudaError_t launch_div_sqrt_check(const float* a, const float* b, float* q, float* s, float* q_ieee,
                                  float* s_ieee, long long n, cudaStream_t st) {
  div_sqrt_check_kernel<<<stream_grid(n, 256, 148), 256, 0, st>>>(a, b, q, s, q_ieee, s_ieee, n);
  return cudaGetLastError();
}

cudaError_t launch_clamp01(const float* z, float* y, long long n, cudaStream_t s) {
  const long long n4 = (reinterpret_cast<uintptr_t>(y) & 15) ? 0 : n / 4;
  clamp01_kernel<<<stream_grid(n4 > 0 ? n4 : n, 256, 148), 256, 0, s>>>(z, y, n4, n);
  return cudaGetLastError();
}

cudaError_t launch_clamp_backward(const float* dy, const float* z, __nv_bfloat16* dz, long long n,
                                  cudaStream_t s) {
  clamp_backward_kernel<<<stream_grid(n, 256, 148), 256, 0, s>>>(dy, z, dz, n);
  return cudaGetLastError();
}

}  // namespace afr
