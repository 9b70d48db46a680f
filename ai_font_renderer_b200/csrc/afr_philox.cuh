// Counter-based dropout generator shared by the front-end kernels (afr_frontend.cu: the fused
// per-sample kernels of the reference shape; afr_wide.cu: the GEMM-based front-end of wider nets).
#pragma once
#include <cstdint>

namespace afr {

// ---- counter-based dropout RNG (Philox4x32-10) ---------------------------------------------
// One call yields 8 x 16-bit uniforms. Key = seed; counter = (block, site | row << 2, global
// sample index, step):
//   site 0 (embedding)  row = 0,         block = (s*E + c) / 8
//   site 1 (attention)  row = h*S + s,   block = t / 8
//   site 2 (fc1)        row = 0,         block = (s*F + j) / 8
// and element i uses the (i % 8)-th 16-bit lane (low half of word i%8/2 first).
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
struct Rng {
  uint32_t k0, k1, sample, step;
  __device__ __forceinline__ uint4 block(uint32_t site_row, uint32_t blk) const {
    return philox4x32_10(blk, site_row, sample, step, k0, k1);
  }
};
template <int SUB>
__device__ __forceinline__ uint32_t u16_of(const uint4& r) {
  const uint32_t word = (SUB >> 1) == 0 ? r.x : (SUB >> 1) == 1 ? r.y : (SUB >> 1) == 2 ? r.z : r.w;
  return (SUB & 1) ? (word >> 16) : (word & 0xFFFFu);
}
__device__ __forceinline__ uint32_t word_of(const uint4& r, int i) {
  return i == 0 ? r.x : i == 1 ? r.y : i == 2 ? r.z : r.w;
}

}  // namespace afr
