// coreg_host.cu -- host-buffer entry points: whole searches behind one C call (H2D, one-time preparation, search, D2H).
#include "coreg_common.cuh"

namespace coreg {
// CoregLagTan rows of the generic kernel from the candidate headers (same formulas as the host's
// hdrshift/engine.py:tan_lag_table): used by coreg_hpc_search_host when the homography kernel does not apply.
__global__ void tan_lag_from_wcs_kernel(const CoregTanWcs* __restrict__ lag_wcs, int n, double alpha_ref_deg,
                                        double grid_lonpole_deg, CoregLagTan* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const CoregTanWcs w = lag_wcs[idx];
  const double f11 = w.cdelt1 * w.pc11, f12 = w.cdelt1 * w.pc12;
  const double f21 = w.cdelt2 * w.pc21, f22 = w.cdelt2 * w.pc22;
  const double det = f11 * f22 - f12 * f21;
  const double i11 = f22 / det, i12 = -f12 / det, i21 = -f21 / det, i22 = f11 / det;
  double sp, cp;
  sincos(grid_lonpole_deg * kD2R, &sp, &cp);
  if (grid_lonpole_deg == 180.0) { sp = 0.0; cp = -1.0; }
  CoregLagTan L;
  sincos((w.crval1 - alpha_ref_deg) * kD2R, &L.sin_da, &L.cos_da);
  sincos(w.crval2 * kD2R, &L.sin_d0, &L.cos_d0);
  L.m11 = (i11 * -cp + i12 * -sp) * kR2D;
  L.m12 = (i11 * sp + i12 * -cp) * kR2D;
  L.m21 = (i21 * -cp + i22 * -sp) * kR2D;
  L.m22 = (i21 * sp + i22 * -cp) * kR2D;
  L.x0 = w.crpix1 - 1.0;
  L.y0 = w.crpix2 - 1.0;
  out[idx] = L;
}
}  // namespace coreg

using namespace coreg;

extern "C" {

int coreg_hpc_search_host(const void* large, int large_dtype, int lnx, int lny, const CoregTanWcs* wcs_large,
                          const void* small, int small_dtype, int snx, int sny, const CoregTanWcs* wcs_small,
                          const CoregTanWcs* lag_wcs, int64_t n_lags, int order, int flags, double* corr,
                          int64_t* nvalid) {
  if (!large || !small || !wcs_large || !wcs_small || !lag_wcs || !corr)
    return fail(COREG_EINVAL, "coreg_hpc_search_host: null pointer");
  if (lnx <= 0 || lny <= 0 || snx <= 0 || sny <= 0 || n_lags <= 0)
    return fail(COREG_EINVAL, "coreg_hpc_search_host: empty input");
  if ((large_dtype != COREG_F32 && large_dtype != COREG_F64) || (small_dtype != COREG_F32 && small_dtype != COREG_F64))
    return fail(COREG_EINVAL, "coreg_hpc_search_host: dtype must be COREG_F32 or COREG_F64");
  const int64_t ns = (int64_t)snx * sny, nl = (int64_t)lnx * lny;
  const size_t lsz = large_dtype == COREG_F32 ? 4 : 8;
  const size_t work_bytes = coreg_lag_corr_workspace_bytes(snx, sny, n_lags);
  const bool fast = (order == 2) && !(flags & (COREG_FLAG_STRICT | COREG_FLAG_NO_FAST)) && snx >= 3 && sny >= 3;
  void *d_large = nullptr, *d_small_in = nullptr, *d_scr = nullptr;
  float* d_small32c = nullptr;
  int* d_flag = nullptr;
  bool mixed_done = false;
  double *d_small = nullptr, *d_lng = nullptr, *d_lat = nullptr, *d_x = nullptr, *d_y = nullptr, *d_planes = nullptr,
         *d_piv = nullptr, *d_corr = nullptr;
  float* d_ref = nullptr;
  CoregTanWcs* d_lagw = nullptr;
  CoregLagTan* d_lags = nullptr;
  int64_t* d_nv = nullptr;
  void* d_work = nullptr;
  cudaStream_t s = nullptr;
  int rc = COREG_OK;
#define TRY(call)                       \
  do {                                  \
    cudaError_t _e = (call);            \
    if (_e != cudaSuccess) {            \
      rc = cuda_fail(_e, #call);        \
      goto done;                        \
    }                                   \
  } while (0)
#define TRYRC(call)      \
  do {                   \
    rc = (call);         \
    if (rc) goto done;   \
  } while (0)
  TRY(cudaMalloc(&d_large, nl * lsz));
  TRY(cudaMalloc(&d_small, ns * sizeof(double)));
  TRY(cudaMalloc(&d_lng, ns * sizeof(double)));
  TRY(cudaMalloc(&d_lat, ns * sizeof(double)));
  TRY(cudaMalloc(&d_x, ns * sizeof(double)));
  TRY(cudaMalloc(&d_y, ns * sizeof(double)));
  TRY(cudaMalloc(&d_ref, ns * sizeof(float)));
  TRY(cudaMalloc(&d_piv, 8 * sizeof(double)));   // [4][2] statistics block; its first row is the pivots
  TRY(cudaMalloc(&d_scr, coreg_image_stats_scratch_bytes()));
  TRY(cudaMemsetAsync(d_scr, 0, coreg_image_stats_scratch_bytes(), s));
  TRY(cudaMalloc(&d_corr, n_lags * sizeof(double)));
  TRY(cudaMalloc(&d_nv, n_lags * sizeof(int64_t)));
  TRY(cudaMalloc(&d_lagw, n_lags * sizeof(CoregTanWcs)));
  TRY(cudaMalloc(&d_work, work_bytes));
  TRY(cudaMemcpyAsync(d_large, large, nl * lsz, cudaMemcpyHostToDevice, s));
  if (small_dtype == COREG_F64) {
    TRY(cudaMemcpyAsync(d_small, small, ns * sizeof(double), cudaMemcpyHostToDevice, s));
    TRYRC(coreg_image_stats(d_small, COREG_F64, ns, nullptr, d_piv + 1, 2, d_scr, s));
  } else {
    // the lag kernels run fastest on float64 storage (no per-tap conversion): widen once on the device, in the same
    // pass that takes the pivot
    TRY(cudaMalloc(&d_small_in, ns * sizeof(float)));
    TRY(cudaMemcpyAsync(d_small_in, small, ns * sizeof(float), cudaMemcpyHostToDevice, s));
    TRYRC(coreg_image_stats(d_small_in, COREG_F32, ns, d_small, d_piv + 1, 2, d_scr, s));
  }
  TRY(cudaMemcpyAsync(d_lagw, lag_wcs, n_lags * sizeof(CoregTanWcs), cudaMemcpyHostToDevice, s));
  TRYRC(coreg_tan_pix2world(wcs_small, snx, sny, 1, d_lng, d_lat, s));
  TRYRC(coreg_tan_world2pix(wcs_large, d_lng, d_lat, ns, d_x, d_y, s));
  TRYRC(coreg_map_coordinates(d_large, large_dtype, lny, lnx, d_y, d_x, ns, order, (double)NAN, d_ref, COREG_F32, s));
  TRYRC(coreg_image_stats(d_ref, COREG_F32, ns, nullptr, d_piv, 2, d_scr, s));
  if (fast && (flags & COREG_FLAG_MIXED) && d_small_in) {
    // opt-in mixed arithmetic: centred float32 payload, guarded per lag; any flagged lag -> the whole search in FP64
    double h_stats[8];
    TRY(cudaMalloc(&d_small32c, ns * sizeof(float)));
    TRY(cudaMalloc(&d_flag, n_lags * sizeof(int)));
    TRYRC(coreg_center_f32((const float*)d_small_in, ns, d_small32c, d_piv + 1, 2, d_scr, s));
    TRY(cudaMemcpyAsync(h_stats, d_piv, sizeof(h_stats), cudaMemcpyDeviceToHost, s));
    TRY(cudaStreamSynchronize(s));
    // float32 headroom of the segment sums (same rule as the Python engine's MIXED_ABS_RANGE)
    if (h_stats[4] > 1e-9 && h_stats[4] < 1e12 && h_stats[5] > 1e-9 && h_stats[5] < 1e12) {
      TRYRC(coreg_hpc_lag_corr_wcs_mixed(d_ref, d_small, d_small32c, snx, sny, snx, sny, wcs_small, d_lagw, n_lags, order,
                                         d_piv, d_work, work_bytes, d_corr, d_nv, d_flag, flags, s));
      int* h_flag = (int*)malloc(n_lags * sizeof(int));
      if (!h_flag) { rc = fail(COREG_ENOMEM, "coreg_hpc_search_host: out of host memory"); goto done; }
      cudaError_t e = cudaMemcpyAsync(h_flag, d_flag, n_lags * sizeof(int), cudaMemcpyDeviceToHost, s);
      if (e == cudaSuccess) e = cudaStreamSynchronize(s);
      int64_t tripped = 0;
      for (int64_t i = 0; i < n_lags; ++i) tripped += h_flag[i];
      free(h_flag);
      if (e != cudaSuccess) { rc = cuda_fail(e, "flag readback"); goto done; }
      mixed_done = tripped == 0;
    }
  }
  if (mixed_done) {
    // cube complete
  } else if (fast) {
    TRYRC(coreg_hpc_lag_corr_wcs(d_ref, d_small, snx, sny, snx, sny, wcs_small, d_lagw, n_lags, order, d_piv, d_work,
                                 work_bytes, d_corr, d_nv, flags, s));
  } else {
    TRY(cudaMalloc(&d_planes, 3 * ns * sizeof(double)));
    TRY(cudaMalloc(&d_lags, n_lags * sizeof(CoregLagTan)));
    TRYRC(coreg_tan_trig_planes(d_lng, d_lat, ns, wcs_small->crval1, d_planes, s));
    tan_lag_from_wcs_kernel<<<((int)n_lags + 127) / 128, 128, 0, s>>>(d_lagw, (int)n_lags, wcs_small->crval1,
                                                                     wcs_small->lonpole, d_lags);
    TRYRC(coreg_hpc_lag_corr(d_ref, d_small, COREG_F64, snx, sny, snx, sny, d_planes, d_lags, n_lags, order, d_piv,
                             d_work, work_bytes, d_corr, d_nv, flags, s));
  }
  TRY(cudaMemcpyAsync(corr, d_corr, n_lags * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (nvalid) TRY(cudaMemcpyAsync(nvalid, d_nv, n_lags * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  TRY(cudaStreamSynchronize(s));
done:
  cudaFree(d_large); cudaFree(d_small_in); cudaFree(d_small); cudaFree(d_lng); cudaFree(d_lat); cudaFree(d_x);
  cudaFree(d_y); cudaFree(d_planes); cudaFree(d_ref); cudaFree(d_piv); cudaFree(d_corr); cudaFree(d_nv);
  cudaFree(d_lagw); cudaFree(d_lags); cudaFree(d_work); cudaFree(d_scr); cudaFree(d_small32c); cudaFree(d_flag);
#undef TRY
#undef TRYRC
  return rc;
}

}  // extern "C"
