// DFMA throughput as a function of how many DISTINCT 64-bit register operands an instruction reads.
//   mode 0:  a_i = fma(a_i, x, y)            x, y shared by every DFMA (operand reuse cache) -- the peak probe
//   mode 1:  a_i = fma(u_j, w, a_i)          two distinct per-instruction registers (a_i, u_j), one shared
//   mode 2:  a_i = fma(u_j, w_k, a_i)        three distinct registers per instruction (the adjoint's accumulate)
//   mode 3:  like the adjoint's inner loop: 4 accumulators x 8 rings, a_k += u_r * w_{r,k}
// Build: nvcc -arch=sm_100a -O3 -o dfma_operands dfma_operands.cu ; run: ./dfma_operands
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double* out, int iters, double seed) {
  double a[16], u[8], w[32];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + i;
#pragma unroll
  for (int i = 0; i < 8; ++i) u[i] = 1.0 + 1e-9 * (seed + i + threadIdx.x);
#pragma unroll
  for (int i = 0; i < 32; ++i) w[i] = 1e-9 * (seed + 3 * i + threadIdx.x);
  const double x = 0.999999, y = 1e-9;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int rep = 0; rep < 8; ++rep)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(a[i], x, y);
    } else if (MODE == 1) {
#pragma unroll
      for (int rep = 0; rep < 8; ++rep)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(u[(i + rep) & 7], x, a[i]);
    } else if (MODE == 2) {
#pragma unroll
      for (int rep = 0; rep < 8; ++rep)
#pragma unroll
        for (int i = 0; i < 16; ++i) a[i] = fma(u[(i + rep) & 7], w[(i * 5 + rep * 3) & 31], a[i]);
    } else {
#pragma unroll
      for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
          for (int q = 0; q < 4; ++q) a[4 * s + q] = fma(u[r], w[4 * r + q], a[4 * s + q]);
      // keep u moving like a recurrence would (cheap, shared operands)
#pragma unroll
      for (int r = 0; r < 8; ++r) u[r] = fma(u[r], x, y);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 12345.678) out[0] = s;
}

template <int MODE>
double run(int blocks, int threads, int iters, double* d) {
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  k<MODE><<<blocks, threads>>>(d, 10, 0.5);
  float best = 1e30f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(a);
    k<MODE><<<blocks, threads>>>(d, iters, 0.5);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    best = ms < best ? ms : best;
  }
  const double nf = MODE == 3 ? 128.0 + 8.0 : 128.0;
  return 2.0 * nf * iters * (double)threads * blocks / (best * 1e-3) / 1e12;
}

int main() {
  double* d;
  cudaMalloc(&d, 8);
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  for (int threads : {128, 256, 512}) {
    const int blocks = p.multiProcessorCount * (1024 / threads);
    printf("threads/CTA %d (x%d CTAs/SM): reuse %.2f  two-distinct %.2f  three-distinct %.2f  adjoint-like %.2f TFLOP/s\n",
           threads, 1024 / threads, run<0>(blocks, threads, 4000, d), run<1>(blocks, threads, 4000, d),
           run<2>(blocks, threads, 4000, d), run<3>(blocks, threads, 4000, d));
  }
  // low occupancy like the adjoint kernel: 12 warps / SM
  for (int warps : {4, 8, 12, 16}) {
    const int blocks = p.multiProcessorCount * warps;
    printf("%2d single-warp CTAs/SM: reuse %.2f  three-distinct %.2f  adjoint-like %.2f TFLOP/s\n", warps,
           run<0>(blocks, 32, 8000, d), run<2>(blocks, 32, 8000, d), run<3>(blocks, 32, 8000, d));
  }
  return 0;
}
