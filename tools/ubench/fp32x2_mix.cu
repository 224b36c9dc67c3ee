// Does a packed FP32 instruction (FFMA2, sm_100) cost one dispatch cycle or two?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_mix fp32x2_mix.cu && ./fp32x2_mix
// MODE 0: 8 FFMA; 1: 8 FFMA2; 2: 8 FFMA2 + 8 int ops; 3: 8 FFMA + 8 int ops; 4: 8 FFMA2 + 4 DFMA; 5: 8 FFMA + 4 DFMA;
// 6: 16 FFMA + 8 int ops (same FP32 work as mode 2 in scalar form)
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters, float m, float c, double dm, double dc) {
  float2 a[8];
  for (int i = 0; i < 8; ++i) a[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
  float f[16]; for (int i = 0; i < 16; ++i) f[i] = threadIdx.x + i;
  unsigned u[8]; for (int i = 0; i < 8; ++i) u[i] = threadIdx.x * 7 + i;
  double d[4]; for (int i = 0; i < 4; ++i) d[i] = threadIdx.x + i;
  const float2 m2 = make_float2(m, m), c2 = make_float2(c, c);
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 1 || MODE == 2 || MODE == 4) a[i] = __ffma2_rn(a[i], m2, c2);
      if (MODE == 0 || MODE == 3 || MODE == 5 || MODE == 6) f[i] = fmaf(f[i], m, c);
      if (MODE == 6) f[i + 8] = fmaf(f[i + 8], m, c);
      if (MODE == 2 || MODE == 3 || MODE == 6) u[i] = u[i] * 3u + (unsigned)it;
      if ((MODE == 4 || MODE == 5) && (i & 1) == 0) d[i >> 1] = fma(d[i >> 1], dm, dc);
    }
  }
  long long t1 = clock64();
  float s = 0; for (int i = 0; i < 8; ++i) s += a[i].x + a[i].y + f[i] + f[i + 8] + u[i];
  for (int i = 0; i < 4; ++i) s += (float)d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE> void run(int threads, float* out, long long* cyc) {
  int iters = 4096;
  k<MODE><<<148, threads>>>(out, cyc, iters, 1.0000001f, 1e-9f, 1.0000001, 1e-9);
  k<MODE><<<148, threads>>>(out, cyc, iters, 1.0000001f, 1e-9f, 1.0000001, 1e-9);
  cudaDeviceSynchronize();
  long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  double per = (double)h / iters / (threads / 128.0);
  printf("mode %d threads %4d: %.2f cycles per loop body per warp and SMSP\n", MODE, threads, per);
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  for (int th : {512, 1024}) { run<0>(th, out, cyc); run<1>(th, out, cyc); run<2>(th, out, cyc); run<3>(th, out, cyc); run<6>(th, out, cyc); run<4>(th, out, cyc); run<5>(th, out, cyc); }
  return 0;
}
