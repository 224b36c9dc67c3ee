// coreg_core.cu -- library-wide state (error string, profiling events), device queries, the FP64 issue-rate microbenchmark.
#include "coreg_common.cuh"

namespace coreg {
thread_local char g_err[512] = "";
thread_local bool g_prof_on = false;
thread_local ProfPair g_prof[4096];
thread_local int g_prof_n = 0;

// ---------------------------------------------------------------------------------------------------------
// FP64 issue-rate microbenchmark
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double seed) {
  double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6,
         a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace coreg

using namespace coreg;

extern "C" {

const char* coreg_last_error(void) { return g_err; }
int coreg_version(void) { return 100; }

int coreg_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return COREG_ECUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return COREG_ECUDA;
  return sms;
}

int coreg_profile_begin(void) {
  for (int i = 0; i < g_prof_n; ++i) {
    cudaEventDestroy(g_prof[i].a);
    cudaEventDestroy(g_prof[i].b);
  }
  g_prof_n = 0;
  g_prof_on = true;
  return COREG_OK;
}

int coreg_profile_end(double* lag_kernel_ms_total, int* launches) {
  g_prof_on = false;
  double tot = 0.0;
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f;
    CK(cudaEventSynchronize(g_prof[i].b));
    CK(cudaEventElapsedTime(&ms, g_prof[i].a, g_prof[i].b));
    tot += ms;
    cudaEventDestroy(g_prof[i].a);
    cudaEventDestroy(g_prof[i].b);
  }
  if (lag_kernel_ms_total) *lag_kernel_ms_total = tot;
  if (launches) *launches = g_prof_n;
  g_prof_n = 0;
  return COREG_OK;
}

int coreg_fp64_peak(double* fma_per_s, int iters, void* stream) {
  if (!fma_per_s || iters <= 0) return fail(COREG_EINVAL, "coreg_fp64_peak: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  int sms = coreg_device_sm_count();
  if (sms <= 0) return fail(COREG_ECUDA, "no device");
  const int blocks = sms * 8, threads = 256;
  double* out = nullptr;
  CK(cudaMalloc(&out, (size_t)blocks * threads * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  fp64_peak_kernel<<<blocks, threads, 0, s>>>(out, iters / 8 + 1, 1.0);  // warm-up
  CK(cudaEventRecord(e0, s));
  fp64_peak_kernel<<<blocks, threads, 0, s>>>(out, iters, 1.0);
  CK(cudaEventRecord(e1, s));
  CK(cudaEventSynchronize(e1));
  CK_LAUNCH("fp64_peak_kernel");
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  *fma_per_s = (double)blocks * threads * 8.0 * (double)iters / ((double)ms * 1e-3);
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return COREG_OK;
}

}  // extern "C"
