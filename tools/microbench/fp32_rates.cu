// Microbenchmark: FP32 issue rates on sm_100a that decide the front-end kernel design.
//   ffma   : 3-register FFMA, 8 independent chains per thread
//   ffma2  : packed fma.rn.f32x2, 8 independent chains (16 FMAs) per thread
//   ex2    : MUFU.EX2
//   lds128 : broadcast LDS.128 + 4 FFMA per load
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32_rates fp32_rates.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& d, const float2& a, const float2& b) {
  asm volatile("{.reg .b64 ra, rb, rc;\n mov.b64 ra, {%2, %3};\n mov.b64 rb, {%4, %5};\n"
               "mov.b64 rc, {%0, %1};\n fma.rn.f32x2 rc, ra, rb, rc;\n mov.b64 {%0, %1}, rc;}\n"
               : "+f"(d.x), "+f"(d.y) : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
}

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float x, float y) {
  __shared__ float4 sm[256];
  sm[threadIdx.x] = make_float4(x, y, x, y);
  __syncthreads();
  float a[8], b[8];
  float2 a2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = x + i; b[i] = y * (i + 1); a2[i] = make_float2(x + i, y - i); }
  const float2 m2 = make_float2(y, x), n2 = make_float2(x * 0.5f, y * 0.25f);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b[i], b[(i + 1) & 7]);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) ffma2(a2[i], (r & 1) ? m2 : n2, a2[(i + 1) & 7]);
    } else if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = exp2f(a[i]);
    } else if (MODE == 3) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 v = sm[(it + r) & 255];   // warp-uniform address: broadcast
        a[0] = fmaf(v.x, b[0], a[0]); a[1] = fmaf(v.y, b[1], a[1]);
        a[2] = fmaf(v.z, b[2], a[2]); a[3] = fmaf(v.w, b[3], a[3]);
      }
    } else if (MODE == 4) {
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 v = sm[(it + r) & 255];
        a[r] += v.x + v.y + v.z + v.w;
      }
    } else if (MODE == 5) {   // LDS.128, 4 distinct addresses per warp (lane & 3)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 v = sm[(it + r * 4 + (threadIdx.x & 3)) & 255];
        a[r] += v.x + v.y + v.z + v.w;
      }
    } else if (MODE == 6) {   // LDS.128, 2 distinct addresses per warp (lane >> 4)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 v = sm[(it + r * 2 + ((threadIdx.x >> 4) & 1)) & 255];
        a[r] += v.x + v.y + v.z + v.w;
      }
    } else if (MODE == 7) {   // LDS.32, conflict-free (lane-consecutive words)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float v = reinterpret_cast<const float*>(sm)[(it * 32 + r * 32 + (threadIdx.x & 31)) & 1023];
        a[r] += v;
      }
    } else if (MODE == 8) {   // LDS.128, every lane its own 16 bytes (512 B per warp)
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float4 v = sm[(it * 32 + r * 32 + (threadIdx.x & 31)) & 255];
        a[r] += v.x + v.y + v.z + v.w;
      }
    } else if (MODE == 9) {   // LDS.64 uniform
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const float2 v = reinterpret_cast<const float2*>(sm)[(it + r) & 511];
        a[r] += v.x + v.y;
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i] + a2[i].x + a2[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, double ops_per_iter_per_thread, int ctas_per_sm) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int grid = sms * ctas_per_sm, iters = 20000;
  float* out;
  cudaMalloc(&out, sizeof(float) * grid * 256);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(out, 100, 1.0001f, 0.9999f);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  k<MODE><<<grid, 256>>>(out, iters, 1.0001f, 0.9999f);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = ops_per_iter_per_thread * iters * grid * 256.0;
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-8s ctas/sm=%d  %.3f ms  %.2f Tops/s  = %.1f ops/clk/SM at %d MHz max\n", name, ctas_per_sm,
         ms, ops / ms / 1e9, ops / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000);
  cudaFree(out);
}

int main() {
  for (int c : {2, 8}) {
    run<0>("ffma", 32, c);
    run<1>("ffma2", 64, c);
    run<2>("ex2", 32, c);
    run<3>("lds+4fma", 8, c);      // counts LDS.128 per iter
    run<4>("lds+4add", 8, c);
    run<5>("lds128x4a", 8, c);
    run<6>("lds128x2a", 8, c);
    run<7>("lds32", 8, c);
    run<8>("lds128all", 8, c);
    run<9>("lds64uni", 8, c);
  }
  return 0;
}
