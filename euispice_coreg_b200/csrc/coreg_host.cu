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

int hpc_search_host_impl(const void* large, int large_dtype, int lnx, int lny, const CoregTanWcs* wcs_large,
                         const void* small, int small_dtype, int snx, int sny, const CoregTanWcs* wcs_small,
                         const CoregTanWcs* lag_wcs, int64_t n_lags, int order, int flags, double* corr,
                         int64_t* nvalid);

// CoregLagTanEdge rows (coreg_hpc_lag_corr_edge) from CoregLagTan rows + the bounds of reproject's edge rule
__global__ void tan_edge_rows_kernel(const CoregLagTan* __restrict__ in, int n, double xhi, double yhi,
                                     CoregLagTanEdge* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  CoregLagTanEdge e;
  e.t = in[idx];
  e.xhi = xhi;
  e.yhi = yhi;
  out[idx] = e;
}

// x[i] += dx, y[i] += dy  (detector coordinates of the large image = offset + plane)
__global__ void shift_planes_kernel(double* __restrict__ x, double* __restrict__ y, int64_t n, double dx, double dy) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    x[i] += dx;
    y[i] += dy;
  }
}
// `np.where(image == -32762, np.nan, image)` (hdrshift/alignment.py:900-901)
__global__ void fill_to_nan_kernel(double* __restrict__ img, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    if (img[i] == -32762.0) img[i] = CUDART_NAN;
}
}  // namespace coreg

using namespace coreg;

extern "C" {

int coreg_hpc_search_host(const void* large, int large_dtype, int lnx, int lny, const CoregTanWcs* wcs_large,
                          const void* small, int small_dtype, int snx, int sny, const CoregTanWcs* wcs_small,
                          const CoregTanWcs* lag_wcs, int64_t n_lags, int order, int flags, double* corr,
                          int64_t* nvalid) {
  return hpc_search_host_impl(large, large_dtype, lnx, lny, wcs_large, small, small_dtype, snx, sny, wcs_small, lag_wcs,
                              n_lags, order, flags, corr, nvalid);
}

}  // extern "C"

namespace coreg {
int hpc_search_host_impl(const void* large, int large_dtype, int lnx, int lny, const CoregTanWcs* wcs_large,
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
  // the one-time cut, fused: world grid -> large-image coordinates -> spline sample -> float32
  TRYRC(coreg_hpc_cut(wcs_small, snx, sny, wcs_large, d_large, large_dtype, lny, lnx, 0, 0, order, d_ref, s));
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
    // the generic kernel works from per-pixel trig planes of the world grid
    TRY(cudaMalloc(&d_lng, ns * sizeof(double)));
    TRY(cudaMalloc(&d_lat, ns * sizeof(double)));
    TRY(cudaMalloc(&d_planes, 3 * ns * sizeof(double)));
    TRY(cudaMalloc(&d_lags, n_lags * sizeof(CoregLagTan)));
    TRYRC(coreg_tan_pix2world(wcs_small, snx, sny, 1, d_lng, d_lat, s));
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
}  // namespace coreg

#include <algorithm>
#include <string>
#include <thread>
#include <vector>

namespace {
// RAII bundle of device allocations for the host entry points
struct DevBufs {
  std::vector<void*> p;
  template <typename T>
  cudaError_t get(T** out, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = cudaMalloc(&q, bytes ? bytes : 1);
    if (e == cudaSuccess) p.push_back(q);
    *out = static_cast<T*>(q);
    return e;
  }
  ~DevBufs() {
    for (void* q : p) cudaFree(q);
  }
};
#define HTRY(call)                                   \
  do {                                               \
    cudaError_t _e = (call);                         \
    if (_e != cudaSuccess) return cuda_fail(_e, #call); \
  } while (0)
#define HRC(call)        \
  do {                   \
    int _rc = (call);    \
    if (_rc) return _rc; \
  } while (0)

// Lag order the Carrington kernel wants (csrc/coreg_lag_offset.cu): 256 consecutive lags = neighbours in the detector
// plane. Without knowledge of the caller's grid structure the offsets themselves are binned: cells of 32 x 32 detector
// pixels, inside a cell bands of 4 pixels in y, x ascending; every cell is padded to a multiple of 256 slots with NaN
// dummies so that a block never straddles two cells. slot_of[k] = row of lag k in the padded table.
int64_t patch_order_from_offsets(const CoregLagOffset* lags, int64_t n, std::vector<int64_t>* slot_of) {
  struct Key { long long cy, cx, band; double x; int64_t k; };
  std::vector<Key> keys((size_t)n);
  for (int64_t k = 0; k < n; ++k) {
    const double x = lags[k].x0, y = lags[k].y0;
    const bool ok = std::isfinite(x) && std::isfinite(y) && fabs(x) < 1e9 && fabs(y) < 1e9;
    keys[(size_t)k] = ok ? Key{(long long)floor(y / 32.0), (long long)floor(x / 32.0), (long long)floor(y / 4.0), x, k}
                         : Key{LLONG_MAX, LLONG_MAX, 0, 0.0, k};
  }
  std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
    if (a.cy != b.cy) return a.cy < b.cy;
    if (a.cx != b.cx) return a.cx < b.cx;
    if (a.band != b.band) return a.band < b.band;
    if (a.x != b.x) return a.x < b.x;
    return a.k < b.k;
  });
  slot_of->assign((size_t)n, 0);
  int64_t slot = 0;
  for (size_t i = 0; i < keys.size(); ++i) {
    if (i > 0 && (keys[i].cy != keys[i - 1].cy || keys[i].cx != keys[i - 1].cx)) slot = (slot + 255) / 256 * 256;
    (*slot_of)[(size_t)keys[i].k] = slot++;
  }
  return (slot + 255) / 256 * 256;
}
}  // namespace

extern "C" {

int coreg_hpc_search_host_multi(const int* devices, int n_devices, const void* large, int large_dtype, int lnx, int lny,
                                const CoregTanWcs* wcs_large, const void* small, int small_dtype, int snx, int sny,
                                const CoregTanWcs* wcs_small, const CoregTanWcs* lag_wcs, int64_t n_lags, int order,
                                int flags, double* corr, int64_t* nvalid) {
  if (!devices || n_devices <= 0) return fail(COREG_EINVAL, "coreg_hpc_search_host_multi: no device given");
  if (!lag_wcs || !corr || n_lags <= 0) return fail(COREG_EINVAL, "coreg_hpc_search_host_multi: empty lag list");
  int prev = 0;
  cudaGetDevice(&prev);
  // contiguous equal slices of the flat lag list, like np.array_split over workers (hdrshift/alignment.py:677-687)
  const int64_t chunk = (n_lags + n_devices - 1) / n_devices;
  std::vector<int> rcs((size_t)n_devices, COREG_OK);
  std::vector<std::string> msgs((size_t)n_devices);
  std::vector<std::thread> workers;
  for (int r = 0; r < n_devices; ++r) {
    const int64_t lo = std::min<int64_t>(r * chunk, n_lags), hi = std::min<int64_t>((r + 1) * chunk, n_lags);
    if (hi <= lo) continue;
    workers.emplace_back([=, &rcs, &msgs]() {
      cudaError_t e = cudaSetDevice(devices[r]);
      int rc = (e == cudaSuccess) ? hpc_search_host_impl(large, large_dtype, lnx, lny, wcs_large, small, small_dtype,
                                                         snx, sny, wcs_small, lag_wcs + lo, hi - lo, order, flags,
                                                         corr + lo, nvalid ? nvalid + lo : nullptr)
                                  : cuda_fail(e, "cudaSetDevice");
      rcs[(size_t)r] = rc;
      if (rc) msgs[(size_t)r] = coreg_last_error();   // the error string is per thread
    });
  }
  for (auto& w : workers) w.join();
  cudaSetDevice(prev);
  for (int r = 0; r < n_devices; ++r)
    if (rcs[(size_t)r]) return fail(rcs[(size_t)r], "device slice failed: %s", msgs[(size_t)r].c_str());
  return COREG_OK;
}

int coreg_carrington_search_host(const void* large, int large_dtype, int lnx, int lny, const CoregCarrington* c_large,
                                 double x0_large, double y0_large, const void* small, int small_dtype, int snx, int sny,
                                 const CoregCarrington* c_small, const double* sinlon_large, const double* coslon_large,
                                 const double* sinlon_small, const double* coslon_small, int n_lon,
                                 const double* sinlat, const double* coslat, int n_lat, const CoregLagOffset* lags,
                                 int64_t n_lags, int order, int flags, double* corr, int64_t* nvalid) {
  if (!large || !small || !c_large || !c_small || !sinlon_large || !coslon_large || !sinlon_small || !coslon_small ||
      !sinlat || !coslat || !lags || !corr)
    return fail(COREG_EINVAL, "coreg_carrington_search_host: null pointer");
  if (lnx <= 0 || lny <= 0 || snx <= 0 || sny <= 0 || n_lon <= 0 || n_lat <= 0 || n_lags <= 0)
    return fail(COREG_EINVAL, "coreg_carrington_search_host: empty input");
  if ((large_dtype != COREG_F32 && large_dtype != COREG_F64) || (small_dtype != COREG_F32 && small_dtype != COREG_F64))
    return fail(COREG_EINVAL, "coreg_carrington_search_host: dtype must be COREG_F32 or COREG_F64");
  const int64_t ng = (int64_t)n_lon * n_lat, ns = (int64_t)snx * sny, nl = (int64_t)lnx * lny;
  const size_t lsz = large_dtype == COREG_F32 ? 4 : 8, ssz = small_dtype == COREG_F32 ? 4 : 8;
  // lags in detector-plane patches, padded with NaN dummies
  std::vector<int64_t> slot_of;
  const int64_t n_slots = patch_order_from_offsets(lags, n_lags, &slot_of);
  std::vector<CoregLagOffset> padded((size_t)n_slots, CoregLagOffset{NAN, NAN});
  for (int64_t k = 0; k < n_lags; ++k) padded[(size_t)slot_of[(size_t)k]] = lags[k];
  const size_t work_bytes = coreg_lag_corr_workspace_bytes(n_lon, n_lat, n_slots);
  DevBufs B;
  cudaStream_t s = nullptr;
  void *d_large, *d_small, *d_work, *d_scr;
  double *d_vec, *d_tx, *d_ty, *d_ref, *d_stats, *d_corr;
  int64_t* d_nv;
  CoregLagOffset* d_lags;
  HTRY(B.get(&d_large, nl * lsz));
  HTRY(B.get(&d_small, ns * ssz));
  HTRY(B.get(&d_vec, (size_t)(4 * n_lon + 2 * n_lat) * sizeof(double)));
  HTRY(B.get(&d_tx, ng * sizeof(double)));
  HTRY(B.get(&d_ty, ng * sizeof(double)));
  HTRY(B.get(&d_ref, ng * sizeof(double)));
  HTRY(B.get(&d_stats, 8 * sizeof(double)));
  HTRY(B.get(&d_scr, coreg_image_stats_scratch_bytes()));
  HTRY(B.get(&d_corr, n_slots * sizeof(double)));
  HTRY(B.get(&d_nv, n_slots * sizeof(int64_t)));
  HTRY(B.get(&d_lags, n_slots * sizeof(CoregLagOffset)));
  HTRY(B.get(&d_work, work_bytes));
  double *d_sll = d_vec, *d_cll = d_vec + n_lon, *d_sls = d_vec + 2 * n_lon, *d_cls = d_vec + 3 * n_lon,
         *d_slat = d_vec + 4 * n_lon, *d_clat = d_vec + 4 * n_lon + n_lat;
  HTRY(cudaMemcpyAsync(d_large, large, nl * lsz, cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_small, small, ns * ssz, cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_sll, sinlon_large, n_lon * sizeof(double), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_cll, coslon_large, n_lon * sizeof(double), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_sls, sinlon_small, n_lon * sizeof(double), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_cls, coslon_small, n_lon * sizeof(double), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_slat, sinlat, n_lat * sizeof(double), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_clat, coslat, n_lat * sizeof(double), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_lags, padded.data(), n_slots * sizeof(CoregLagOffset), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemsetAsync(d_scr, 0, coreg_image_stats_scratch_bytes(), s));
  // one-time: large image -> Carrington grid, float64, fill -32762 -> NaN (alignment.py:646-648, 889-901)
  HRC(coreg_carrington_planes(c_large, d_sll, d_cll, n_lon, d_slat, d_clat, n_lat, d_tx, d_ty, s));
  shift_planes_kernel<<<grid_for(ng), 256, 0, s>>>(d_tx, d_ty, ng, x0_large, y0_large);
  HRC(coreg_map_coordinates(d_large, large_dtype, lny, lnx, d_ty, d_tx, ng, order, -32762.0, d_ref, COREG_F64, s));
  fill_to_nan_kernel<<<grid_for(ng), 256, 0, s>>>(d_ref, ng);
  HTRY(cudaGetLastError());
  // the small image's planes, pivots, search
  HRC(coreg_carrington_planes(c_small, d_sls, d_cls, n_lon, d_slat, d_clat, n_lat, d_tx, d_ty, s));
  HRC(coreg_image_stats(d_ref, COREG_F64, ng, nullptr, d_stats, 2, d_scr, s));
  HRC(coreg_image_stats(d_small, small_dtype, ns, nullptr, d_stats + 1, 2, d_scr, s));
  HRC(coreg_offset_lag_corr(d_ref, d_small, small_dtype, snx, sny, n_lon, n_lat, d_tx, d_ty, d_lags, n_slots, order,
                            d_stats, d_work, work_bytes, d_corr, d_nv, flags, s));
  std::vector<double> h_corr((size_t)n_slots);
  std::vector<int64_t> h_nv((size_t)n_slots);
  HTRY(cudaMemcpyAsync(h_corr.data(), d_corr, n_slots * sizeof(double), cudaMemcpyDeviceToHost, s));
  HTRY(cudaMemcpyAsync(h_nv.data(), d_nv, n_slots * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  HTRY(cudaStreamSynchronize(s));
  for (int64_t k = 0; k < n_lags; ++k) {
    corr[k] = h_corr[(size_t)slot_of[(size_t)k]];
    if (nvalid) nvalid[k] = h_nv[(size_t)slot_of[(size_t)k]];
  }
  return COREG_OK;
}

int coreg_surface_search_host(const void* large, int large_dtype, int lnx, int lny, const CoregTanWcs* wcs_large,
                              const void* small, int small_dtype, int snx, int sny, const CoregTanWcs* wcs_small,
                              const CoregSurfaceFrames* frames, const CoregTanWcs* lag_wcs, int64_t n_lags, int flags,
                              double* corr, int64_t* nvalid) {
  if (!large || !small || !wcs_large || !wcs_small || !frames || !lag_wcs || !corr)
    return fail(COREG_EINVAL, "coreg_surface_search_host: null pointer");
  if (lnx <= 0 || lny <= 0 || snx <= 0 || sny <= 0 || n_lags <= 0)
    return fail(COREG_EINVAL, "coreg_surface_search_host: empty input");
  if ((large_dtype != COREG_F32 && large_dtype != COREG_F64) || (small_dtype != COREG_F32 && small_dtype != COREG_F64))
    return fail(COREG_EINVAL, "coreg_surface_search_host: dtype must be COREG_F32 or COREG_F64");
  const int64_t ns = (int64_t)snx * sny, nl = (int64_t)lnx * lny;
  const size_t lsz = large_dtype == COREG_F32 ? 4 : 8, ssz = small_dtype == COREG_F32 ? 4 : 8;
  const size_t work_bytes = coreg_lag_corr_workspace_bytes(snx, sny, n_lags);
  DevBufs B;
  cudaStream_t s = nullptr;
  void *d_large, *d_small, *d_work, *d_scr;
  double *d_large_pad, *d_small_pad, *d_ref, *d_lng, *d_lat, *d_planes, *d_stats, *d_corr;
  int64_t* d_nv;
  CoregTanWcs* d_lagw;
  CoregLagTan* d_lags;
  CoregLagTanEdge* d_edge;
  HTRY(B.get(&d_large, nl * lsz));
  HTRY(B.get(&d_small, ns * ssz));
  HTRY(B.get(&d_large_pad, (size_t)(lnx + 2) * (lny + 2) * sizeof(double)));
  HTRY(B.get(&d_small_pad, (size_t)(snx + 2) * (sny + 2) * sizeof(double)));
  HTRY(B.get(&d_ref, ns * sizeof(double)));
  HTRY(B.get(&d_lng, ns * sizeof(double)));
  HTRY(B.get(&d_lat, ns * sizeof(double)));
  HTRY(B.get(&d_planes, 3 * ns * sizeof(double)));
  HTRY(B.get(&d_stats, 8 * sizeof(double)));
  HTRY(B.get(&d_scr, coreg_image_stats_scratch_bytes()));
  HTRY(B.get(&d_corr, n_lags * sizeof(double)));
  HTRY(B.get(&d_nv, n_lags * sizeof(int64_t)));
  HTRY(B.get(&d_lagw, n_lags * sizeof(CoregTanWcs)));
  HTRY(B.get(&d_lags, n_lags * sizeof(CoregLagTan)));
  HTRY(B.get(&d_edge, n_lags * sizeof(CoregLagTanEdge)));
  HTRY(B.get(&d_work, work_bytes));
  HTRY(cudaMemcpyAsync(d_large, large, nl * lsz, cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_small, small, ns * ssz, cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_lagw, lag_wcs, n_lags * sizeof(CoregTanWcs), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemsetAsync(d_scr, 0, coreg_image_stats_scratch_bytes(), s));
  // once: the large image on the small grid through the solar-surface change of observer (alignment.py:939-956)
  HRC(coreg_pad_edge(d_large, large_dtype, lny, lnx, d_large_pad, s));
  HRC(coreg_surface_cut(wcs_small, snx, sny, wcs_large, d_large_pad, lny, lnx, frames, d_ref, s));
  // per lag: the bilinear helioprojective search with reproject's edge rule on the padded small image
  HRC(coreg_pad_edge(d_small, small_dtype, sny, snx, d_small_pad, s));
  HRC(coreg_tan_pix2world(wcs_small, snx, sny, 1, d_lng, d_lat, s));
  HRC(coreg_tan_trig_planes(d_lng, d_lat, ns, wcs_small->crval1, d_planes, s));
  tan_lag_from_wcs_kernel<<<((int)n_lags + 127) / 128, 128, 0, s>>>(d_lagw, (int)n_lags, wcs_small->crval1,
                                                                   wcs_small->lonpole, d_lags);
  tan_edge_rows_kernel<<<((int)n_lags + 127) / 128, 128, 0, s>>>(d_lags, (int)n_lags, (double)snx - 0.5,
                                                                (double)sny - 0.5, d_edge);
  HTRY(cudaGetLastError());
  HRC(coreg_image_stats(d_ref, COREG_F64, ns, nullptr, d_stats, 2, d_scr, s));
  HRC(coreg_image_stats(d_small, small_dtype, ns, nullptr, d_stats + 1, 2, d_scr, s));
  HRC(coreg_hpc_lag_corr_edge(d_ref, d_small_pad, snx, sny, snx, sny, d_planes, d_edge, n_lags, d_stats, d_work,
                              work_bytes, d_corr, d_nv, flags, s));
  HTRY(cudaMemcpyAsync(corr, d_corr, n_lags * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (nvalid) HTRY(cudaMemcpyAsync(nvalid, d_nv, n_lags * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  HTRY(cudaStreamSynchronize(s));
  return COREG_OK;
}

int coreg_synras_build_host(const void* frames, int frame_dtype, int n_frames, int fnx, int fny, const CoregTanWcs* wcs,
                            const int* frame_of_col, const double* lng, const double* lat, int n_rows, int n_cols,
                            int order, double* out) {
  if (!frames || !wcs || !frame_of_col || !lng || !lat || !out)
    return fail(COREG_EINVAL, "coreg_synras_build_host: null pointer");
  if (n_frames <= 0 || fnx <= 0 || fny <= 0 || n_rows <= 0 || n_cols <= 0)
    return fail(COREG_EINVAL, "coreg_synras_build_host: empty input");
  if (frame_dtype != COREG_F32 && frame_dtype != COREG_F64)
    return fail(COREG_EINVAL, "coreg_synras_build_host: frame_dtype must be COREG_F32 or COREG_F64");
  const size_t fsz = frame_dtype == COREG_F32 ? 4 : 8;
  const int64_t n = (int64_t)n_rows * n_cols;
  DevBufs B;
  cudaStream_t s = nullptr;
  void* d_frames;
  double *d_lng, *d_lat, *d_out;
  HTRY(B.get(&d_frames, (size_t)n_frames * fnx * fny * fsz));
  HTRY(B.get(&d_lng, n * sizeof(double)));
  HTRY(B.get(&d_lat, n * sizeof(double)));
  HTRY(B.get(&d_out, n * sizeof(double)));
  HTRY(cudaMemcpyAsync(d_frames, frames, (size_t)n_frames * fnx * fny * fsz, cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_lng, lng, n * sizeof(double), cudaMemcpyHostToDevice, s));
  HTRY(cudaMemcpyAsync(d_lat, lat, n * sizeof(double), cudaMemcpyHostToDevice, s));
  HRC(coreg_synras_build(d_frames, frame_dtype, n_frames, fnx, fny, wcs, frame_of_col, d_lng, d_lat, n_rows, n_cols,
                         order, d_out, s));
  HTRY(cudaMemcpyAsync(out, d_out, n * sizeof(double), cudaMemcpyDeviceToHost, s));
  HTRY(cudaStreamSynchronize(s));
  return COREG_OK;
}

}  // extern "C"
