#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: a=fma(a,m,c) shared m,c | 1: a=fma(a,b,c) distinct regs | 2: a=fma(b,c,a) | 3: a=a+b | 4: a=a*b | 5: mix dadd,dmul,dfma distinct
// 6: fma(a,b,c) + 1 FSEL-type int op per fma
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, double m, double c0) {
  double a[8], b[8], c[8];
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x + i; b[i] = 1.0 + 1e-9 * (threadIdx.x + i); c[i] = 1e-9 * i + c0; }
  unsigned u[8]; for (int i = 0; i < 8; ++i) u[i] = threadIdx.x + i;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) a[i] = fma(a[i], m, c0);
      if (MODE == 1 || MODE == 6) a[i] = fma(a[i], b[i], c[i]);
      if (MODE == 2) a[i] = fma(b[i], c[(i + 3) & 7], a[i]);
      if (MODE == 3) a[i] = a[i] + b[i];
      if (MODE == 4) a[i] = a[i] * b[i];
      if (MODE == 5) { if (i % 3 == 0) a[i] = a[i] + b[i]; else if (i % 3 == 1) a[i] = a[i] * b[i]; else a[i] = fma(a[i], b[i], c[i]); }
      if (MODE == 6) u[i] = max(u[i], (unsigned)__double2hiint(a[i]));
    }
  }
  long long t1 = clock64();
  double s = 0; for (int i = 0; i < 8; ++i) s += a[i] + b[i] + c[i] + u[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(int threads, double* out, long long* cyc) {
  int iters = 2048;
  k<MODE><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  k<MODE><<<148, threads>>>(out, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("mode %d threads %4d: %.2f cycles per FP64 op per SMSP\n", MODE, threads, (double)h / iters / 8.0 / (threads / 128.0));
}
int main() {
  double* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8);
  for (int th : {512, 768}) { run<0>(th, out, cyc); run<1>(th, out, cyc); run<2>(th, out, cyc); run<3>(th, out, cyc); run<4>(th, out, cyc); run<5>(th, out, cyc); run<6>(th, out, cyc); }
  return 0;
}
