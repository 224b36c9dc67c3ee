#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k(double* out, long long* cyc, int iters, double m, double c) {
  double a[ILP];
  for (int i = 0; i < ILP; ++i) a[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int ILP> void run(int threads, double* out, long long* cyc) {
  int iters = 4096;
  k<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  k<ILP><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / iters;   // cycles per iteration (ILP dfma per warp)
  int warps_per_smsp = threads / 32 / 4; if (warps_per_smsp < 1) warps_per_smsp = 1;
  printf("ILP %d threads %4d (warps/SMSP %d): %.2f cycles/iter -> %.2f cycles per DFMA per SMSP (peak 2.0)\n", ILP, threads, warps_per_smsp, per, per / (ILP * (threads >= 128 ? threads / 128.0 : 1.0)));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int th : {32, 128, 256, 512, 768, 1024}) { run<1>(th, out, cyc); run<2>(th, out, cyc); run<4>(th, out, cyc); run<8>(th, out, cyc); }
  return 0;
}
