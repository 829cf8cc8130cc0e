#include <cstdio>
#include <cuda_runtime.h>
template <int NCH>
__global__ void k(double* out, int iters, double x, double y) {
  double a[NCH];
#pragma unroll
  for (int i = 0; i < NCH; ++i) a[i] = x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u)
#pragma unroll
      for (int i = 0; i < NCH; ++i) a[i] = fma(a[i], x, y);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < NCH; ++i) s += a[i];
  if (threadIdx.x == 0) { out[blockIdx.x * 2] = (double)(t1 - t0) / (double)(iters * 4 * NCH); out[blockIdx.x*2+1] = s; }
}
template <int NCH> void run(int warps_per_block, int blocks) {
  double* d; cudaMalloc(&d, 1 << 20);
  k<NCH><<<blocks, warps_per_block * 32>>>(d, 2000, 0.999, 1e-9);
  cudaDeviceSynchronize();
  k<NCH><<<blocks, warps_per_block * 32>>>(d, 20000, 0.999, 1e-9);
  double h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("chains/thread %2d  warps/block %2d : %.2f cycles per DFMA (per warp)\n", NCH, warps_per_block, h[0]);
  cudaFree(d);
}
int main() {
  // one warp per SMSP -> exposes latency / ILP behaviour
  run<1>(4, 148); run<2>(4, 148); run<4>(4, 148); run<8>(4, 148); run<12>(4, 148); run<16>(4, 148); run<24>(4, 148); run<32>(4, 148);
  run<1>(8, 148); run<4>(8, 148); run<8>(8, 148); run<16>(8, 148);
  run<1>(16, 148); run<4>(16, 148); run<8>(16, 148);
  return 0;
}
