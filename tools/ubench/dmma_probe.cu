// Does the FP64 tensor-core path (DMMA m8n8k4) share the issue/execution resources of the FP64 FMA pipe on
// sm_100a?  Three kernels with the same loop structure: DFMA only, DMMA only, both interleaved.  If the
// interleaved time is close to max(a, b) the two are independent pipes; close to a + b: one pipe.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu && ./dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}

template <int MODE>  // 0: dfma, 1: dmma, 2: both
__global__ void k(double* out, int iters, double seed) {
  double f[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    f[i] = seed + i;
    c[i] = seed - i;
  }
  const double x = 0.999999, y = 1e-9, a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (MODE != 1) {
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = fma(f[i], x, y);
      }
      if (MODE != 0) {
#pragma unroll
        for (int i = 0; i < 8; i += 2) dmma(c[i], c[i + 1], a, b);
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += f[i] + c[i];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
float run(double* d, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.0);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 256>>>(d, iters, 1.0);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  double* d;
  cudaMalloc(&d, 8);
  const int iters = 20000;
  const float a = run<0>(d, iters), b = run<1>(d, iters), c = run<2>(d, iters);
  const double thr = 148.0 * 8 * 256;
  // per iteration and thread: 32 DFMA (64 flop); 16 DMMA per warp = 16*512 flop / 32 threads = 256 flop per thread
  printf("dfma only : %8.3f ms  %7.2f TFLOP/s\n", a, thr * iters * 64.0 / a / 1e9);
  printf("dmma only : %8.3f ms  %7.2f TFLOP/s\n", b, thr * iters * 256.0 / b / 1e9);
  printf("both      : %8.3f ms  (sum %.3f, max %.3f)\n", c, a + b, a > b ? a : b);
  return 0;
}
