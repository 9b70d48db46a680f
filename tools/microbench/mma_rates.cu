// Microbenchmark: issue rate of the warp-level (legacy) tensor instructions on sm_100a, which
// decides whether the narrow front-end (head_dim 8, E = 32) is worth moving off FFMA2.
//   mode 0: mma.sync.m16n8k8  tf32 x tf32 + f32, 4 independent accumulators per warp
//   mode 1: mma.sync.m16n8k16 bf16 x bf16 + f32, 4 independent accumulators per warp
//   mode 2: mma.sync.m16n8k8  tf32, ONE dependent accumulator chain (latency)
//   mode 3: mma.sync.m16n8k4  tf32
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rates mma_rates.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_tf32_k4(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k4.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(b[0]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__global__ void k(float* out, int iters, uint32_t seed) {
  uint32_t a[4] = {seed, seed + 1, seed + 2, seed + 3}, b[2] = {seed * 3, seed * 5};
  float c[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (MODE == 0) mma_tf32(c[i], a, b);
        if (MODE == 1) mma_bf16(c[i], a, b);
        if (MODE == 2) mma_tf32(c[0], a, b);
        if (MODE == 3) mma_tf32_k4(c[i], a, b);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) s += c[i][j];
  if (s == 123.456f) out[threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int warps, int ctas_per_sm, int sms, float* out, double fma_per_mma) {
  const int iters = 4096;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<sms * ctas_per_sm, warps * 32>>>(out, 16, 0);
  cudaEventRecord(e0);
  k<MODE><<<sms * ctas_per_sm, warps * 32>>>(out, iters, 0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double mmas_per_sm = double(iters) * 16 * warps * ctas_per_sm;
  const double clk = ms * 1e-3 * khz * 1e3;
  printf("%-22s warps/SM %2d: %.3f ms, %.2f clk per MMA per SM (at %d MHz nominal), %.1f TFLOP/s\n", name,
         warps * ctas_per_sm, ms, clk / mmas_per_sm, khz / 1000,
         2.0 * fma_per_mma * mmas_per_sm * sms / (ms * 1e-3) * 1e-12);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  cudaMalloc(&out, 4096);
  for (int w : {4, 8, 16, 32}) {
    run<0>("tf32 m16n8k8 x4 indep", w > 16 ? 16 : w, w > 16 ? 2 : 1, sms, out, 16 * 8 * 8);
    run<1>("bf16 m16n8k16 x4 indep", w > 16 ? 16 : w, w > 16 ? 2 : 1, sms, out, 16 * 8 * 16);
    run<3>("tf32 m16n8k4 x4 indep", w > 16 ? 16 : w, w > 16 ? 2 : 1, sms, out, 16 * 8 * 4);
  }
  run<2>("tf32 m16n8k8 dependent", 4, 1, sms, out, 16 * 8 * 8);
  run<2>("tf32 m16n8k8 dependent", 16, 1, sms, out, 16 * 8 * 8);
  printf("done: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
