#include <cstdio>
#include <cuda_runtime.h>
// MODE 0: 8 DFMA; 1: 8 DFMA + 8 int ops; 2: 8 DFMA + 8 LDG.64 (L1 hits); 3: 8 DFMA + 2 F2F pairs; 4: 8 DFMA + 8 FFMA; 5: 8 DFMA+16 int
template <int MODE>
__global__ void k(double* out, const double* __restrict__ tbl, long long* cyc, int iters, double m, double c) {
  double a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i;
  unsigned u[16]; for (int i = 0; i < 16; ++i) u[i] = threadIdx.x * 7 + i;
  float f[8]; for (int i = 0; i < 8; ++i) f[i] = threadIdx.x + i;
  double acc = 0;
  unsigned off = threadIdx.x & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = fma(a[i], m, c);
      if (MODE == 1 || MODE == 5) u[i] = u[i] * 3u + (unsigned)it;
      if (MODE == 5) u[i + 8] = (u[i + 8] ^ (unsigned)it) + 5u;
      if (MODE == 2) { acc += __ldg(tbl + ((off + i * 37 + it) & 1023)); }
      if (MODE == 3 && (i & 3) == 0) { float t = __double2float_rn(a[i]); a[i] = (double)t; }
      if (MODE == 4) f[i] = fmaf(f[i], 1.0001f, 0.5f);
    }
  }
  long long t1 = clock64();
  double s = acc; for (int i = 0; i < 8; ++i) s += a[i] + u[i] + u[i+8] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(int threads, double* out, double* tbl, long long* cyc) {
  int iters = 2048;
  k<MODE><<<148, threads>>>(out, tbl, cyc, iters, 1.0000001, 1e-9);
  k<MODE><<<148, threads>>>(out, tbl, cyc, iters, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / iters / 8.0 / (threads / 128.0);
  printf("mode %d threads %4d: %.2f cycles per DFMA per SMSP (peak 2.0)\n", MODE, threads, per);
}
int main() {
  double* out; long long* cyc; double* tbl;
  cudaMalloc(&out, 148 * 1024 * 8); cudaMalloc(&cyc, 8); cudaMalloc(&tbl, 1024 * 8); cudaMemset(tbl, 0, 8192);
  for (int th : {256, 768}) { run<0>(th, out, tbl, cyc); run<1>(th, out, tbl, cyc); run<5>(th, out, tbl, cyc); run<2>(th, out, tbl, cyc); run<3>(th, out, tbl, cyc); run<4>(th, out, tbl, cyc); }
  return 0;
}
