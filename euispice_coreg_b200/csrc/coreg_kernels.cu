// coreg_kernels.cu -- hand-written sm_100a kernels + the C ABI of include/coreg_b200.h.
//
// Hot path of adolliou/euispice_coreg's pointing search (hdrshift/alignment.py:509-549, 613-797):
// per candidate header ("lag"), map every common-grid pixel into the small image, sample it with an order-k
// B-spline exactly as scipy.ndimage.map_coordinates(prefilter=False) does, and score the pair of images with a
// masked Pearson coefficient. Here all lags of a launch are evaluated by one kernel: a thread block owns a tile
// of the common grid (its lag-independent per-pixel constants live in registers), walks the lag list, and emits
// one 6-moment partial per (tile, lag); a second kernel folds the partials in a fixed order (deterministic,
// independent of how lags are sharded over GPUs) and turns moments into r.
//
// FP64 throughout: the work is gather + FP64 arithmetic + reduction, there is no dense contraction, so no tensor
// cores (tcgen05) are involved by design. Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo.
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>
#include <type_traits>
#include <math_constants.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/coreg_b200.h"

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char* fmt, const char* a = "", const char* b = "") {
  snprintf(g_err, sizeof(g_err), fmt, a, b);
  return code;
}
int cuda_fail(cudaError_t e, const char* where) {
  return fail(COREG_ECUDA, "%s: %s", where, cudaGetErrorString(e));
}
#define CK(call)                                         \
  do {                                                   \
    cudaError_t _e = (call);                             \
    if (_e != cudaSuccess) return cuda_fail(_e, #call);  \
  } while (0)
#define CK_LAUNCH(name)                                       \
  do {                                                        \
    cudaError_t _e = cudaGetLastError();                      \
    if (_e != cudaSuccess) return cuda_fail(_e, name);        \
  } while (0)

// optional per-launch timing of the fused lag kernel (bench.py): event pairs recorded around it while enabled
struct ProfPair { cudaEvent_t a, b; };
thread_local bool g_prof_on = false;
thread_local ProfPair g_prof[4096];
thread_local int g_prof_n = 0;

constexpr double kD2R = 0.017453292519943295769236907684886;
constexpr double kR2D = 57.295779513082320876798154814105;
constexpr double kMagic = 6755399441055744.0;  // 1.5 * 2^52

// ---------------------------------------------------------------------------------------------------------
// arithmetic helpers: `STRICT` keeps scipy's separate multiply / add (no FMA contraction)
// ---------------------------------------------------------------------------------------------------------
template <bool STRICT>
__device__ __forceinline__ double mul_(double a, double b) {
  return STRICT ? __dmul_rn(a, b) : a * b;
}
template <bool STRICT>
__device__ __forceinline__ double add_(double a, double b) {
  return STRICT ? __dadd_rn(a, b) : a + b;
}
template <bool STRICT>
__device__ __forceinline__ double sub_(double a, double b) {
  return STRICT ? __dsub_rn(a, b) : a - b;
}

// floor(s) for |s| < 2^31 on the FP64 pipe only (no F2F/F2I): add 1.5*2^52 rounding toward -inf, the integer
// lands in the low mantissa word. Exact, i.e. identical to floor().
__device__ __forceinline__ double floor_magic(double s, int& i) {
  const double m = __dadd_rd(s, kMagic);
  i = __double2loint(m);
  return __dsub_rn(m, kMagic);
}

// Spline start index and weights of scipy's map_coordinates without prefilter (ni_splines.c), orders 0..3.
template <int ORDER, bool STRICT>
__device__ __forceinline__ void spline_weights(double t, int& start, double (&w)[ORDER + 1]) {
  int i0;
  if (ORDER == 0) {
    floor_magic(__dadd_rn(t, 0.5), i0);
    start = i0;
    w[0] = 1.0;
  } else if (ORDER == 1) {
    const double fl = floor_magic(t, i0);
    const double d = __dsub_rn(t, fl);
    start = i0;
    w[0] = __dsub_rn(1.0, d);
    w[ORDER >= 1 ? 1 : 0] = __dsub_rn(1.0, w[0]);
  } else if (ORDER == 2) {
    const double fl = floor_magic(__dadd_rn(t, 0.5), i0);
    const double d = __dsub_rn(t, fl);
    start = i0 - 1;
    const double u = __dsub_rn(0.5, d);
    if (STRICT) {
      w[ORDER >= 2 ? 1 : 0] = __dsub_rn(0.75, __dmul_rn(d, d));
      w[0] = __dmul_rn(__dmul_rn(0.5, u), u);
    } else {
      w[ORDER >= 2 ? 1 : 0] = fma(-d, d, 0.75);
      w[0] = (0.5 * u) * u;
    }
    w[ORDER >= 2 ? 2 : 0] = __dsub_rn(__dsub_rn(1.0, w[0]), w[ORDER >= 2 ? 1 : 0]);
  } else {
    const double fl = floor_magic(t, i0);
    const double d = __dsub_rn(t, fl);
    start = i0 - 1;
    const double z = __dsub_rn(1.0, d);
    const double w1 = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(d, d), __dsub_rn(d, 2.0)), 3.0), 4.0), 6.0);
    const double w2 = __ddiv_rn(__dadd_rn(__dmul_rn(__dmul_rn(__dmul_rn(z, z), __dsub_rn(z, 2.0)), 3.0), 4.0), 6.0);
    const double w0 = __ddiv_rn(__dmul_rn(__dmul_rn(z, z), z), 6.0);
    w[0] = w0;
    w[ORDER >= 3 ? 1 : 0] = w1;
    w[ORDER >= 3 ? 2 : 0] = w2;
    w[ORDER >= 3 ? 3 : 0] = __dsub_rn(__dsub_rn(__dsub_rn(1.0, w0), w1), w2);
  }
}

__device__ __forceinline__ int mirror_index(int i, int n) {
  // scipy 'constant' mode keeps the full spline support near an edge by reflecting about the edge pixel centre
  if (n == 1) return 0;
  if (i < 0) i = -i;
  if (i > n - 1) i = 2 * (n - 1) - i;
  return min(max(i, 0), n - 1);
}

template <typename T>
__device__ __forceinline__ double ldval(const T* p) {
  return (double)__ldg(p);
}

// One sample of map_coordinates(img, (y, x), order=ORDER, mode='constant', prefilter=False).
// Returns false when the point is outside [0, n-1] on either axis (NaN coordinates included) -> caller uses cval.
template <int ORDER, bool STRICT, typename T>
__device__ __forceinline__ bool spline_sample(const T* __restrict__ img, int ny, int nx, double y, double x,
                                              double& out) {
  const bool inside = (y >= 0.0) && (y <= (double)(ny - 1)) && (x >= 0.0) && (x <= (double)(nx - 1));
  if (!inside) return false;
  int sy, sx;
  double wy[ORDER + 1], wx[ORDER + 1];
  spline_weights<ORDER, STRICT>(y, sy, wy);
  spline_weights<ORDER, STRICT>(x, sx, wx);
  double t = 0.0;
  const bool interior = (sy >= 0) && (sy + ORDER <= ny - 1) && (sx >= 0) && (sx + ORDER <= nx - 1);
  if (interior) {
    const T* p = img + (sy * nx + sx);  // callers guarantee ny*nx < 2^31
    if (STRICT) {
#pragma unroll
      for (int a = 0; a <= ORDER; ++a) {
#pragma unroll
        for (int b = 0; b <= ORDER; ++b) {
          t = __dadd_rn(t, __dmul_rn(__dmul_rn(ldval(p + b), wy[a]), wx[b]));
        }
        p += nx;
      }
    } else {
#pragma unroll
      for (int a = 0; a <= ORDER; ++a) {
        double row = ldval(p) * wx[0];
#pragma unroll
        for (int b = 1; b <= ORDER; ++b) row = fma(ldval(p + b), wx[b], row);
        t = fma(row, wy[a], t);
        p += nx;
      }
    }
  } else {
    int iy[ORDER + 1], ix[ORDER + 1];
#pragma unroll
    for (int a = 0; a <= ORDER; ++a) {
      iy[a] = mirror_index(sy + a, ny);
      ix[a] = mirror_index(sx + a, nx);
    }
#pragma unroll
    for (int a = 0; a <= ORDER; ++a) {
#pragma unroll
      for (int b = 0; b <= ORDER; ++b) {
        const double v = ldval(img + (iy[a] * nx + ix[b]));
        if (STRICT)
          t = __dadd_rn(t, __dmul_rn(__dmul_rn(v, wy[a]), wx[b]));
        else
          t = fma(v * wy[a], wx[b], t);
      }
    }
  }
  out = t;
  return true;
}

// ---------------------------------------------------------------------------------------------------------
// TAN (gnomonic) device math
// ---------------------------------------------------------------------------------------------------------
struct TanDev {
  double crpix1, crpix2;
  double f11, f12, f21, f22;  // cdelt_i * pc_ij * D2R : pixel offset -> projection plane [rad]
  double i11, i12, i21, i22;  // inverse, projection plane [rad] -> pixel offset
  double a0_deg, s0, c0;      // CRVAL1 [deg], sin/cos CRVAL2
  double lonpole_rad;
  double a0_rad;
};

int make_tan(const CoregTanWcs* w, TanDev* t) {
  if (!w) return fail(COREG_EINVAL, "null CoregTanWcs");
  const double f11 = w->cdelt1 * w->pc11, f12 = w->cdelt1 * w->pc12;
  const double f21 = w->cdelt2 * w->pc21, f22 = w->cdelt2 * w->pc22;
  const double det = f11 * f22 - f12 * f21;
  if (!(det != 0.0) || det != det) return fail(COREG_EINVAL, "singular CDELT*PC matrix");
  t->crpix1 = w->crpix1;
  t->crpix2 = w->crpix2;
  t->f11 = f11 * kD2R;
  t->f12 = f12 * kD2R;
  t->f21 = f21 * kD2R;
  t->f22 = f22 * kD2R;
  t->i11 = (f22 / det) * kR2D;
  t->i12 = (-f12 / det) * kR2D;
  t->i21 = (-f21 / det) * kR2D;
  t->i22 = (f11 / det) * kR2D;
  t->a0_deg = w->crval1;
  t->a0_rad = w->crval1 * kD2R;
  t->s0 = sin(w->crval2 * kD2R);
  t->c0 = cos(w->crval2 * kD2R);
  t->lonpole_rad = w->lonpole * kD2R;
  return COREG_OK;
}

__device__ __forceinline__ double wrap_pipi_deg(double a) {
  // -((-a + 180) % 360 - 180) with Python's floor-mod (utils/Util.py:76-80)
  double m = fmod(-a + 180.0, 360.0);
  if (m != 0.0 && m < 0.0) m += 360.0;
  return -(m - 180.0);
}

__global__ void tan_pix2world_kernel(TanDev w, int nx, int ny, int wrap, double* __restrict__ lng,
                                     double* __restrict__ lat) {
  const int64_t n = (int64_t)nx * ny;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % nx), j = (int)(idx / nx);
    const double u1 = ((double)i + 1.0) - w.crpix1;
    const double u2 = ((double)j + 1.0) - w.crpix2;
    const double px = w.f11 * u1 + w.f12 * u2;
    const double py = w.f21 * u1 + w.f22 * u2;
    const double r2 = px * px + py * py;
    const double r = sqrt(r2);
    const double phi = (r == 0.0) ? 0.0 : atan2(px, -py);
    const double st = rsqrt(1.0 + r2);  // sin(theta), theta = atan2(1, r)
    const double ct = r * st;
    double sp, cp;
    sincos(phi - w.lonpole_rad, &sp, &cp);
    const double xx = st * w.c0 - ct * w.s0 * cp;
    const double yy = -ct * sp;
    const double zz = st * w.s0 + ct * w.c0 * cp;
    double lo = w.a0_deg + atan2(yy, xx) * kR2D;
    if (w.a0_deg >= 0.0) {
      if (lo < 0.0) lo += 360.0;
    } else {
      if (lo > 0.0) lo -= 360.0;
    }
    double la = atan2(zz, sqrt(xx * xx + yy * yy)) * kR2D;
    if (wrap) {
      lo = wrap_pipi_deg(lo);
      la = wrap_pipi_deg(la);
    }
    lng[idx] = lo;
    lat[idx] = la;
  }
}

__device__ __forceinline__ void tan_world2pix_dev(const TanDev& w, double lng_deg, double lat_deg, double& x,
                                                  double& y) {
  double sl, cl, sa, ca;
  sincos(lat_deg * kD2R, &sl, &cl);
  sincos(lng_deg * kD2R - w.a0_rad, &sa, &ca);
  const double den = sl * w.s0 + cl * w.c0 * ca;
  const double xs = sl * w.c0 - cl * w.s0 * ca;
  const double ys = -cl * sa;
  // phi = lonpole + atan2(ys, xs); plane = (r sin phi, -r cos phi), r = hypot(xs, ys) / den
  double sp, cp;
  sincos(w.lonpole_rad, &sp, &cp);
  // sin(phi) * hypot = sp*xs + cp*ys ; cos(phi) * hypot = cp*xs - sp*ys
  const double inv = 1.0 / den;
  const double xi = (sp * xs + cp * ys) * inv;
  const double eta = -(cp * xs - sp * ys) * inv;
  x = w.i11 * xi + w.i12 * eta + (w.crpix1 - 1.0);
  y = w.i21 * xi + w.i22 * eta + (w.crpix2 - 1.0);
  if (!(den > 0.0)) {
    x = CUDART_NAN;
    y = CUDART_NAN;
  }
}

__global__ void tan_world2pix_kernel(TanDev w, const double* __restrict__ lng, const double* __restrict__ lat,
                                     int64_t n, double* __restrict__ x, double* __restrict__ y) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double xx, yy;
    tan_world2pix_dev(w, lng[idx], lat[idx], xx, yy);
    x[idx] = xx;
    y[idx] = yy;
  }
}

__global__ void tan_trig_planes_kernel(const double* __restrict__ lng, const double* __restrict__ lat, int64_t n,
                                       double alpha_ref_rad, double* __restrict__ planes) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double sl, cl, sa, ca;
    sincos(lat[idx] * kD2R, &sl, &cl);
    sincos(lng[idx] * kD2R - alpha_ref_rad, &sa, &ca);
    planes[idx] = sl;
    planes[n + idx] = cl * sa;
    planes[2 * n + idx] = cl * ca;
  }
}

// ---------------------------------------------------------------------------------------------------------
// map_coordinates at explicit coordinates (one-shot resampling: K2, K5 large image, host API interpol2d)
// ---------------------------------------------------------------------------------------------------------
template <int ORDER, typename TI, typename TO>
__global__ void map_coordinates_kernel(const TI* __restrict__ img, int ny, int nx, const double* __restrict__ yc,
                                       const double* __restrict__ xc, int64_t n, double cval, TO* __restrict__ out) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double v;
    if (!spline_sample<ORDER, true, TI>(img, ny, nx, yc[idx], xc[idx], v)) v = cval;
    out[idx] = (TO)v;
  }
}

template <typename TI, typename TO>
int launch_map_coordinates(const TI* img, int ny, int nx, const double* y, const double* x, int64_t n, int order,
                           double cval, TO* out, cudaStream_t s) {
  const int threads = 256;
  const int blocks = (int)std::min<int64_t>((n + threads - 1) / threads, 148 * 16);
  if (n == 0) return COREG_OK;
  switch (order) {
    case 0: map_coordinates_kernel<0, TI, TO><<<blocks, threads, 0, s>>>(img, ny, nx, y, x, n, cval, out); break;
    case 1: map_coordinates_kernel<1, TI, TO><<<blocks, threads, 0, s>>>(img, ny, nx, y, x, n, cval, out); break;
    case 2: map_coordinates_kernel<2, TI, TO><<<blocks, threads, 0, s>>>(img, ny, nx, y, x, n, cval, out); break;
    case 3: map_coordinates_kernel<3, TI, TO><<<blocks, threads, 0, s>>>(img, ny, nx, y, x, n, cval, out); break;
    default: return fail(COREG_EINVAL, "spline order must be 0..3");
  }
  CK_LAUNCH("map_coordinates_kernel");
  return COREG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// mean of finite values (pivot). One block, fixed traversal order -> deterministic.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void finite_mean_kernel(const T* __restrict__ img, int64_t n, double* __restrict__ mean) {
  __shared__ double ssum[1024];
  __shared__ unsigned long long scnt[1024];
  double s = 0.0;
  unsigned long long c = 0;
  // coarse sample (every 4th element) is plenty for a pivot and keeps this one-block kernel short
  for (int64_t i = (int64_t)threadIdx.x * 4; i < n; i += (int64_t)blockDim.x * 4) {
    const double v = (double)img[i];
    if (isfinite(v)) {
      s += v;
      ++c;
    }
  }
  ssum[threadIdx.x] = s;
  scnt[threadIdx.x] = c;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      ssum[threadIdx.x] += ssum[threadIdx.x + o];
      scnt[threadIdx.x] += scnt[threadIdx.x + o];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) mean[0] = scnt[0] ? ssum[0] / (double)scnt[0] : 0.0;
}

// ---------------------------------------------------------------------------------------------------------
// plate-carree (-CAR) device math: a sphere rotation followed by (phi, theta) -> pixel, see CoregLagCar
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void car_map_unit(const CoregLagCar& L, double cx, double cy, double cz, double& x,
                                             double& y) {
  const double vx = fma(L.r[0], cx, fma(L.r[1], cy, L.r[2] * cz));
  const double vy = fma(L.r[3], cx, fma(L.r[4], cy, L.r[5] * cz));
  const double vz = fma(L.r[6], cx, fma(L.r[7], cy, L.r[8] * cz));
  const double phi = atan2(vy, vx) * kR2D;
  const double theta = atan2(vz, sqrt(fma(vx, vx, vy * vy))) * kR2D;
  x = fma(L.m11, phi, fma(L.m12, theta, L.x0));
  y = fma(L.m21, phi, fma(L.m22, theta, L.y0));
}

__global__ void car_pix2world_kernel(CoregLagCar L, double f11, double f12, double f21, double f22, int nx, int ny,
                                     double* __restrict__ lng, double* __restrict__ lat) {
  const int64_t n = (int64_t)nx * ny;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % nx), j = (int)(idx / nx);
    const double u1 = (double)i - L.x0, u2 = (double)j - L.y0;
    const double phi = (f11 * u1 + f12 * u2) * kD2R, theta = (f21 * u1 + f22 * u2) * kD2R;
    double sp, cp, st, ct;
    sincos(phi, &sp, &cp);
    sincos(theta, &st, &ct);
    const double nx_ = ct * cp, ny_ = ct * sp, nz_ = st;
    // celestial = R^T native
    const double cx = L.r[0] * nx_ + L.r[3] * ny_ + L.r[6] * nz_;
    const double cy = L.r[1] * nx_ + L.r[4] * ny_ + L.r[7] * nz_;
    const double cz = L.r[2] * nx_ + L.r[5] * ny_ + L.r[8] * nz_;
    double lo = atan2(cy, cx) * kR2D;
    if (L.lng_ref >= 0.0) {
      if (lo < 0.0) lo += 360.0;
    } else {
      if (lo > 0.0) lo -= 360.0;
    }
    lng[idx] = lo;
    lat[idx] = atan2(cz, sqrt(cx * cx + cy * cy)) * kR2D;
  }
}

__global__ void car_world2pix_kernel(CoregLagCar L, const double* __restrict__ lng, const double* __restrict__ lat,
                                     int64_t n, double* __restrict__ x, double* __restrict__ y) {
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double sl, cl, sa, ca, xx, yy;
    sincos(lat[idx] * kD2R, &sl, &cl);
    sincos(lng[idx] * kD2R, &sa, &ca);
    car_map_unit(L, cl * ca, cl * sa, sl, xx, yy);
    x[idx] = xx;
    y[idx] = yy;
  }
}

// ---------------------------------------------------------------------------------------------------------
// fused lag search
// ---------------------------------------------------------------------------------------------------------
constexpr int kTileW = 64;
constexpr int kThreads = 256;
constexpr int kRowsPerPass = kThreads / kTileW;  // 4 grid rows per pass of the block
constexpr int kWarps = kThreads / 32;
constexpr int kLagSub = 64;   // lags staged in shared memory at a time
constexpr int kMom = 8;       // n, Sa, Sb, Saa, Sbb, Sab, pad, pad  (64 B per (tile, lag) partial)
constexpr int kMinTileH = 16; // smallest tile height of any variant (workspace sizing)

constexpr int kRollWRows = 12;  // smallest rows-per-thread of the rolling kernel (tile height 48): sizes the workspace
struct RollWLayout {
  size_t rows, rec, corr, cst, mask, total;   // byte offsets into the workspace
};
inline RollWLayout rollw_layout(int gnx, int gny, int64_t n_lags) {
  RollWLayout L;
  const size_t tiles = (size_t)((gnx + kTileW - 1) / kTileW) * ((gny + 4 * kRollWRows - 1) / (4 * kRollWRows));
  L.rows = tiles * (kThreads / 32);
  L.rec = 0;
  L.corr = L.rec + L.rows * (size_t)n_lags * 3 * sizeof(double);
  L.cst = L.corr + L.rows * (size_t)n_lags * 3 * sizeof(double);
  L.mask = L.cst + L.rows * 3 * sizeof(double);
  L.total = L.mask + ((tiles * (size_t)n_lags * sizeof(unsigned) + 15) / 16) * 16;
  return L;
}
inline size_t partials_bytes(int gnx, int gny, int64_t n_lags) {
  const size_t tiles = (size_t)((gnx + kTileW - 1) / kTileW) * ((gny + kMinTileH - 1) / kMinTileH);
  return std::max(tiles * (size_t)n_lags * kMom * sizeof(double), rollw_layout(gnx, gny, n_lags).total);
}

struct TanCoord {
  typedef CoregLagTan Lag;
  struct Planes {
    const double* p;  // [3][n]
    int64_t n;
  };
  struct Pix {
    double p0, p1, p2;
  };
  __device__ static __forceinline__ Pix load(const Planes& pl, int64_t idx) {
    Pix q;
    q.p0 = __ldg(pl.p + idx);
    q.p1 = __ldg(pl.p + pl.n + idx);
    q.p2 = __ldg(pl.p + 2 * pl.n + idx);
    return q;
  }
  __device__ static __forceinline__ Pix dead() {
    Pix q;
    q.p0 = q.p1 = q.p2 = CUDART_NAN;
    return q;
  }
  // world -> pixel of the lag's header; NaN when behind the tangent hemisphere
  __device__ static __forceinline__ void map(const Pix& q, const Lag& L, double& x, double& y) {
    const double qs = fma(q.p1, L.cos_da, -(q.p2 * L.sin_da));  // cos(lat) sin(dA)
    const double pc = fma(q.p2, L.cos_da, q.p1 * L.sin_da);     // cos(lat) cos(dA)
    const double den = fma(pc, L.cos_d0, q.p0 * L.sin_d0);
    const double en = fma(-pc, L.sin_d0, q.p0 * L.cos_d0);
    const double inv = 1.0 / den;
    const double xi = qs * inv, eta = en * inv;
    x = fma(L.m11, xi, fma(L.m12, eta, L.x0));
    y = fma(L.m21, xi, fma(L.m22, eta, L.y0));
    if (!(den > 0.0)) x = CUDART_NAN;
  }
};

struct OffsetCoord {
  typedef CoregLagOffset Lag;
  struct Planes {
    const double* tx;
    const double* ty;
  };
  struct Pix {
    double tx, ty;
  };
  __device__ static __forceinline__ Pix load(const Planes& pl, int64_t idx) {
    Pix q;
    q.tx = __ldg(pl.tx + idx);
    q.ty = __ldg(pl.ty + idx);
    return q;
  }
  __device__ static __forceinline__ Pix dead() {
    Pix q;
    q.tx = q.ty = CUDART_NAN;
    return q;
  }
  __device__ static __forceinline__ void map(const Pix& q, const Lag& L, double& x, double& y) {
    x = __dadd_rn(L.x0, q.tx);
    y = __dadd_rn(L.y0, q.ty);
  }
};

// plate-carree candidate headers: planes = (sin lat, cos lat sin lng, cos lat cos lng) of the common grid's pixels
// (coreg_tan_trig_planes with alpha_ref = 0), per lag one sphere rotation + two atan2
struct CarCoord {
  typedef CoregLagCar Lag;
  typedef TanCoord::Planes Planes;
  typedef TanCoord::Pix Pix;
  __device__ static __forceinline__ Pix load(const Planes& pl, int64_t idx) { return TanCoord::load(pl, idx); }
  __device__ static __forceinline__ Pix dead() { return TanCoord::dead(); }
  __device__ static __forceinline__ void map(const Pix& q, const Lag& L, double& x, double& y) {
    car_map_unit(L, q.p2, q.p1, q.p0, x, y);
  }
};

// butterfly that leaves, in every lane, the warp total of value index (lane >> 2) & 7 : 9 shuffles instead of 40
__device__ __forceinline__ double warp_transpose_reduce8(double (&v)[8], int lane) {
  double w4[4], w2[2], w1;
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const double send = up ? v[i] : v[i + 4];
      const double keep = up ? v[i + 4] : v[i];
      w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double send = up ? w4[i] : w4[i + 2];
      const double keep = up ? w4[i + 2] : w4[i];
      w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
  }
  {
    const bool up = lane & 4;
    const double send = up ? w2[0] : w2[1];
    const double keep = up ? w2[1] : w2[0];
    w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
  return w1;  // value index = 4*bit4 + 2*bit3 + bit2 = (lane >> 2) & 7
}

// One block = one tile of the common grid (64 x 4*PPT pixels, PPT pixels per thread, their lag-independent
// constants in registers) x one slice of the lag list. work layout: [tile][lag][kMom] doubles.
// Moments: the sums over the reference image (Sa, Saa) are taken once over the pixels whose reference value is
// finite and corrected, per lag, by the (rare) pixels whose small-image sample is missing; Sb, Sbb, Sab and the
// count are accumulated per lag.
template <class Coord, int ORDER, bool STRICT, typename SmallT, typename RefT, bool ROUND32, int PPT, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
lag_corr_kernel(const RefT* __restrict__ ref, const SmallT* __restrict__ small, int snx, int sny, int gnx, int gny,
                typename Coord::Planes planes, const typename Coord::Lag* __restrict__ lags, int n_lags,
                int lags_per_block, const double* __restrict__ pivots, double* __restrict__ work) {
  typedef typename Coord::Lag Lag;
  typedef typename Coord::Pix Pix;
  constexpr int TILE_H = kRowsPerPass * PPT;
  __shared__ Lag s_lag[kLagSub];
  __shared__ double s_part[kWarps][kLagSub][kMom];

  const int tiles_x = (gnx + kTileW - 1) / kTileW;
  const int tile = blockIdx.x;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & (kTileW - 1), ty0 = tid / kTileW;
  const int gx = tile_x * kTileW + tx;
  const double pivot_a = pivots[0], pivot_b = pivots[1];

  Pix pix[PPT];
  double a_c[PPT];       // ref - pivot (0 where the reference pixel is missing)
  unsigned a_ok = 0;     // bit k: reference pixel k is finite
  double sa_all = 0.0, saa_all = 0.0;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int gy = tile_y * TILE_H + ty0 + k * kRowsPerPass;
    a_c[k] = 0.0;
    pix[k] = Coord::dead();
    if (gx < gnx && gy < gny) {
      const int64_t idx = (int64_t)gy * gnx + gx;
      const double a = (double)ref[idx];
      if (isfinite(a)) {
        a_c[k] = a - pivot_a;
        a_ok |= 1u << k;
        pix[k] = Coord::load(planes, idx);
        sa_all += a_c[k];
        saa_all = fma(a_c[k], a_c[k], saa_all);
      }
    }
  }
  int n_all = __popc(a_ok);

  const int lag_begin = blockIdx.y * lags_per_block;
  const int lag_end = min(n_lags, lag_begin + lags_per_block);
  for (int l0 = lag_begin; l0 < lag_end; l0 += kLagSub) {
    const int cnt = min(kLagSub, lag_end - l0);
    __syncthreads();  // previous sub-chunk fully consumed
    {
      const double* src = reinterpret_cast<const double*>(lags + l0);
      double* dst = reinterpret_cast<double*>(s_lag);
      const int nd = cnt * (int)(sizeof(Lag) / sizeof(double));
      for (int i = tid; i < nd; i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    for (int l = 0; l < cnt; ++l) {
      const Lag L = s_lag[l];
      double sb = 0.0, sbb = 0.0, sab = 0.0, sa_miss = 0.0, saa_miss = 0.0;
      int n_miss = 0;
#pragma unroll
      for (int k = 0; k < PPT; ++k) {
        double x, y, v;
        Coord::map(pix[k], L, x, y);   // dead pixels carry NaN -> "outside"
        bool ok = spline_sample<ORDER, STRICT, SmallT>(small, sny, snx, y, x, v);
        double b;
        if (ROUND32) {
          const float bf = __double2float_rn(v);
          ok = ok && isfinite(bf);
          b = (double)bf;
        } else {
          ok = ok && isfinite(v) && (v != -32762.0);
          b = v;
        }
        if (ok) {
          const double bc = b - pivot_b;
          sb += bc;
          sbb = fma(bc, bc, sbb);
          sab = fma(a_c[k], bc, sab);
        } else if (a_ok & (1u << k)) {
          ++n_miss;
          sa_miss += a_c[k];
          saa_miss = fma(a_c[k], a_c[k], saa_miss);
        }
      }
      double m[8];
      m[0] = (double)(n_all - n_miss);
      m[1] = sa_all - sa_miss;
      m[2] = sb;
      m[3] = saa_all - saa_miss;
      m[4] = sbb;
      m[5] = sab;
      m[6] = 0.0;
      m[7] = 0.0;
      const double tot = warp_transpose_reduce8(m, lane);
      if ((lane & 3) == 0) s_part[warp][l][lane >> 2] = tot;
    }
    __syncthreads();
    // fold the warps in a fixed order and publish this tile's partials
    for (int i = tid; i < cnt * kMom; i += kThreads) {
      const int l = i / kMom, q = i % kMom;
      double s = s_part[0][l][q];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) s += s_part[w][l][q];
      work[((size_t)tile * n_lags + (l0 + l)) * kMom + q] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Fast variant of the fused lag kernel: order-2 spline, FMA arithmetic.
//
// Helioprojective frame = homography. The common grid and every candidate header are gnomonic (TAN) projections
// of the same sphere from its centre, so pixel (i, j) of the common grid maps to the candidate's pixel through a
// plane projective transformation, exactly:
//      (nx, ny, D) = H (i, j, 1)^T ,   x = x0 + nx / D ,   y = y0 + ny / D ,
// H = [plane'->pixel'] . E(lag)^T Rz(alpha0 - alpha') E(grid) . [pixel->plane]   (3x3, one per lag, built by
// tan_homography_kernel from the two CoregTanWcs). No per-pixel trig, no world-coordinate planes; per sample the
// map costs 3 FMA + the reciprocal. D = cos(angle to the lag's reference point) * sqrt(1 + r^2) is within 2^-7 of 1
// for fields smaller than ~6 deg, where 1/D is the product form of the geometric series (1+e)(1+e^2)(1+e^4),
// e = 1 - D, exact to 2^-56 (COREG_FLAG_SMALL_ANGLE, guaranteed by the caller); otherwise a true division is used.
//
// Both functors deliver coordinates already offset by +0.5, so floor(x + 0.5) is one magic-number add, and
// "strictly interior" (all 9 taps inside the image: no closed-bound test, no mirroring) is one unsigned integer
// compare per axis on the floor index. Groups of pixels take the branch-free path together; anything else (image
// borders, missing reference pixels) falls back to the exact generic sampler with the same coordinates.
// ---------------------------------------------------------------------------------------------------------
struct HomLag {
  double hx0, hx1, hx2, hy0, hy1, hy2, he0, he1, he2, x0h, y0h;  // he = (0,0,1) - (denominator row)
  double emax;  // max |e| over the common grid (e is linear in (i, j): attained at a corner); +inf if not finite
};

struct HomGrid {  // pixel (i, j, 1) -> native direction (-Y, X, 1), and the grid's Euler matrix
  double c[3][3];
  double e[3][3];
  double a0_rad;
  double xmax, ymax;  // gnx - 1, gny - 1
};

// E(delta0, lonpole): native unit vector -> celestial frame whose x axis points at longitude alpha0
__host__ __device__ inline void euler_matrix(double sin_d, double cos_d, double sin_lp, double cos_lp, double (&e)[3][3]) {
  e[0][0] = -sin_d * cos_lp; e[0][1] = -sin_d * sin_lp; e[0][2] = cos_d;
  e[1][0] = sin_lp;          e[1][1] = -cos_lp;         e[1][2] = 0.0;
  e[2][0] = cos_d * cos_lp;  e[2][1] = cos_d * sin_lp;  e[2][2] = sin_d;
}

int make_hom_grid(const CoregTanWcs* w, int gnx, int gny, HomGrid* g) {
  TanDev t;
  int rc = make_tan(w, &t);
  if (rc) return rc;
  const double cx = t.f11 * (1.0 - t.crpix1) + t.f12 * (1.0 - t.crpix2);
  const double cy = t.f21 * (1.0 - t.crpix1) + t.f22 * (1.0 - t.crpix2);
  // d_native = (-Y, X, 1) with X = f11 i + f12 j + cx, Y = f21 i + f22 j + cy
  g->c[0][0] = -t.f21; g->c[0][1] = -t.f22; g->c[0][2] = -cy;
  g->c[1][0] = t.f11;  g->c[1][1] = t.f12;  g->c[1][2] = cx;
  g->c[2][0] = 0.0;    g->c[2][1] = 0.0;    g->c[2][2] = 1.0;
  euler_matrix(t.s0, t.c0, sin(t.lonpole_rad), cos(t.lonpole_rad), g->e);
  g->a0_rad = t.a0_rad;
  g->xmax = (double)(gnx - 1);
  g->ymax = (double)(gny - 1);
  return COREG_OK;
}

__global__ void tan_homography_kernel(HomGrid g, const CoregTanWcs* __restrict__ lag_wcs, int n,
                                      HomLag* __restrict__ out) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n) return;
  const CoregTanWcs w = lag_wcs[idx];
  double sd, cd, sl, cl, sa, ca;
  sincos(w.crval2 * kD2R, &sd, &cd);
  sincos(w.lonpole * kD2R, &sl, &cl);
  sincos(g.a0_rad - w.crval1 * kD2R, &sa, &ca);  // Rz(alpha0 - alpha')
  double e2[3][3];
  euler_matrix(sd, cd, sl, cl, e2);
  // m = Rz * E(grid) * C   (celestial direction in the lag's alpha frame, as a function of (i, j, 1))
  double ec[3][3], m[3][3], r[3][3];
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) ec[a][b] = g.e[a][0] * g.c[0][b] + g.e[a][1] * g.c[1][b] + g.e[a][2] * g.c[2][b];
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    m[0][b] = ca * ec[0][b] - sa * ec[1][b];
    m[1][b] = sa * ec[0][b] + ca * ec[1][b];
    m[2][b] = ec[2][b];
  }
  // r = E(lag)^T m : native direction of the lag's projection
#pragma unroll
  for (int a = 0; a < 3; ++a)
#pragma unroll
    for (int b = 0; b < 3; ++b) r[a][b] = e2[0][a] * m[0][b] + e2[1][a] * m[1][b] + e2[2][a] * m[2][b];
  // plane' = (r1 / r2, -r0 / r2) [rad]; pixel' = inv(cdelt pc) plane' + crpix - 1
  const double f11 = w.cdelt1 * w.pc11 * kD2R, f12 = w.cdelt1 * w.pc12 * kD2R;
  const double f21 = w.cdelt2 * w.pc21 * kD2R, f22 = w.cdelt2 * w.pc22 * kD2R;
  const double det = f11 * f22 - f12 * f21;
  const double i11 = f22 / det, i12 = -f12 / det, i21 = -f21 / det, i22 = f11 / det;
  HomLag h;
  h.hx0 = i11 * r[1][0] - i12 * r[0][0]; h.hx1 = i11 * r[1][1] - i12 * r[0][1]; h.hx2 = i11 * r[1][2] - i12 * r[0][2];
  h.hy0 = i21 * r[1][0] - i22 * r[0][0]; h.hy1 = i21 * r[1][1] - i22 * r[0][1]; h.hy2 = i21 * r[1][2] - i22 * r[0][2];
  h.he0 = -r[2][0]; h.he1 = -r[2][1]; h.he2 = 1.0 - r[2][2];
  h.x0h = (w.crpix1 - 1.0) + 0.5;
  h.y0h = (w.crpix2 - 1.0) + 0.5;
  // e = he0 i + he1 j + he2 is linear over the grid: its extreme values sit at the four corners
  const double e00 = h.he2, e10 = fma(h.he0, g.xmax, h.he2), e01 = fma(h.he1, g.ymax, h.he2),
               e11 = fma(h.he0, g.xmax, fma(h.he1, g.ymax, h.he2));
  double em = fmax(fmax(fabs(e00), fabs(e10)), fmax(fabs(e01), fabs(e11)));
  if (!(em == em) || !isfinite(h.hx0 + h.hx1 + h.hx2 + h.hy0 + h.hy1 + h.hy2)) em = CUDART_INF;
  h.emax = em;
  out[idx] = h;
}

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

__global__ void f32_to_f64_kernel(const float* __restrict__ in, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)in[i];
}

// reciprocal of the projective denominator D = 1 - e, chosen per lag (block-uniform) from the lag's emax:
//   |e| <= 2^-18 : 1 + e + e^2                    (2 ops, truncation e^3 <= 2^-54)
//   |e| <= 2^-7  : (1 + e)(1 + e^2)(1 + e^4)      (5 ops, truncation e^8 <= 2^-56)
//   otherwise    : true division; D <= 0 (behind the tangent hemisphere) -> NaN
constexpr double kTinyE = 3.814697265625e-06;  // 2^-18
constexpr double kSmallE = 0.0078125;          // 2^-7

__device__ __forceinline__ double recip_1me_tiny(double e) { return fma(e, e, 1.0 + e); }
__device__ __forceinline__ double recip_1me_small(double e) {
  const double e2 = e * e;
  double inv = 1.0 + e;
  inv = fma(e2, inv, inv);
  const double e4 = e2 * e2;
  return fma(e4, inv, inv);
}
__device__ __forceinline__ double recip_1me_div(double e) {
  const double den = 1.0 - e;
  return (den > 0.0) ? 1.0 / den : CUDART_NAN;
}

// |v| < 2^30 (false for NaN / Inf): the range in which the magic-number floor of the fast kernels is exact
__device__ __forceinline__ bool small_magnitude(double v) {
  return (unsigned)(__double2hiint(v) & 0x7FFFFFFF) < 0x41D00000u;
}

// exact order-2 sample at (sx - 0.5, sy - 0.5) with float32 rounding / masking, kept out of line: only image
// borders, irregular columns and missing reference pixels come here
template <bool ROUND32>
__device__ __noinline__ bool sample_exact_half(const double* __restrict__ small, int sny, int snx, double sy,
                                               double sx, double* out) {
  double v;
  bool ok = spline_sample<2, false, double>(small, sny, snx, sy - 0.5, sx - 0.5, v);
  if (ROUND32) {
    const float bf = __double2float_rn(v);
    ok = ok && isfinite(bf);
    v = (double)bf;
  } else {
    ok = ok && isfinite(v) && (v != -32762.0);
  }
  *out = v;
  return ok;
}

struct OffsetFastLag {
  double x0h, y0h;
};

__global__ void offset_fast_table_kernel(const CoregLagOffset* __restrict__ lags, int n, OffsetFastLag* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  OffsetFastLag f;
  f.x0h = lags[i].x0 + 0.5;
  f.y0h = lags[i].y0 + 0.5;
  out[i] = f;
}

// per lag slice (blockIdx.y of the fast kernel): range of the offsets, so that a tile whose whole bounding box
// falls outside the small image for every lag of the slice can be skipped
__global__ void offset_lag_range_kernel(const OffsetFastLag* __restrict__ ft, int n_lags, int lags_per_block,
                                        double* __restrict__ ranges) {
  __shared__ double s[4][128];
  const int lo = blockIdx.x * lags_per_block, hi = min(n_lags, lo + lags_per_block);
  double x0 = CUDART_INF, x1 = -CUDART_INF, y0 = CUDART_INF, y1 = -CUDART_INF;
  bool bad = false;
  for (int i = lo + threadIdx.x; i < hi; i += blockDim.x) {
    const OffsetFastLag f = ft[i];
    bad = bad || !(f.x0h == f.x0h) || !(f.y0h == f.y0h);
    x0 = fmin(x0, f.x0h); x1 = fmax(x1, f.x0h);
    y0 = fmin(y0, f.y0h); y1 = fmax(y1, f.y0h);
  }
  if (bad) { x0 = y0 = -CUDART_INF; x1 = y1 = CUDART_INF; }   // NaN offsets: never skip
  s[0][threadIdx.x] = x0; s[1][threadIdx.x] = x1; s[2][threadIdx.x] = y0; s[3][threadIdx.x] = y1;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s[0][threadIdx.x] = fmin(s[0][threadIdx.x], s[0][threadIdx.x + o]);
      s[1][threadIdx.x] = fmax(s[1][threadIdx.x], s[1][threadIdx.x + o]);
      s[2][threadIdx.x] = fmin(s[2][threadIdx.x], s[2][threadIdx.x + o]);
      s[3][threadIdx.x] = fmax(s[3][threadIdx.x], s[3][threadIdx.x + o]);
    }
    __syncthreads();
  }
  if (threadIdx.x < 4) ranges[blockIdx.x * 4 + threadIdx.x] = s[threadIdx.x][0];
}

struct OffsetFast {
  typedef OffsetFastLag LagC;
  typedef OffsetCoord::Planes Planes;
  struct Thread {};
  typedef OffsetCoord::Pix Pix;
  struct TL {};
  static constexpr bool kCoordsAlwaysFinite = false;  // dead pixels carry NaN offsets -> never interior
  __device__ static __forceinline__ Thread thread_init(int) { return Thread(); }
  __device__ static __forceinline__ Pix load(const Planes& pl, int64_t idx, int) { return OffsetCoord::load(pl, idx); }
  __device__ static __forceinline__ Pix dead() { return OffsetCoord::dead(); }
  __device__ static __forceinline__ TL thread_lag(const LagC&, const Thread&) { return TL(); }
  __device__ static __forceinline__ double plane_x(const Pix& q) { return q.tx; }
  __device__ static __forceinline__ double plane_y(const Pix& q) { return q.ty; }
  __device__ static __forceinline__ void map_half(const Pix& q, const TL&, const LagC& C, double& sx, double& sy) {
    sx = C.x0h + q.tx;
    sy = C.y0h + q.ty;
  }
};

// butterfly for 4 values: every lane ends with the warp total of value index (lane >> 3) & 3 (6 shuffles)
__device__ __forceinline__ double warp_transpose_reduce4(double (&v)[4], int lane) {
  double w2[2], w1;
  {
    const bool up = lane & 16;
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const double send = up ? v[i] : v[i + 2];
      const double keep = up ? v[i + 2] : v[i];
      w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
  }
  {
    const bool up = lane & 8;
    const double send = up ? w2[0] : w2[1];
    const double keep = up ? w2[1] : w2[0];
    w1 = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  w1 += __shfl_xor_sync(0xffffffffu, w1, 4);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
  w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
  return w1;  // value index = 2*bit4 + bit3
}

constexpr int kFastLagSub = 32;


template <class Fast, typename SmallT, typename RefT, bool ROUND32, int PPT, int MINB, int GROUP>
__global__ void __launch_bounds__(kThreads, MINB)
lag_corr_fast_kernel(const RefT* __restrict__ ref, const SmallT* __restrict__ small, int snx, int sny, int gnx, int gny,
                     typename Fast::Planes planes, const typename Fast::LagC* __restrict__ fast_lags, int n_lags,
                     int lags_per_block, const double* __restrict__ pivots, double* __restrict__ work,
                     const double* __restrict__ ranges) {
  typedef typename Fast::Pix Pix;
  typedef typename Fast::LagC LagC;
  constexpr int TILE_H = kRowsPerPass * PPT;
  __shared__ double s_box[4][kWarps];
  // GROUP = pixels whose dependency chains are interleaved (their coordinates / indices are live together)
  static_assert(PPT % GROUP == 0, "PPT must be a multiple of GROUP");
  __shared__ __align__(16) LagC s_lag[kFastLagSub];
  __shared__ double s_part[kWarps][kFastLagSub][kMom];
  __shared__ double s_wconst[kWarps][3];                 // per warp: n, Sa, Saa over its finite reference pixels
  __shared__ unsigned char s_miss[kWarps][kFastLagSub];  // 1 when the (warp, lag) slot carries its own n, Sa, Saa

  const int tiles_x = (gnx + kTileW - 1) / kTileW;
  const int tile = blockIdx.x;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & (kTileW - 1), ty0 = tid / kTileW;
  const int gx = tile_x * kTileW + tx;
  const double pivot_a = pivots[0], pivot_b = pivots[1];
  const unsigned ux = (unsigned)(snx - 2), uy = (unsigned)(sny - 2);  // launcher guarantees snx, sny >= 3
  const unsigned row_elems = (unsigned)snx;

  const typename Fast::Thread tstate = Fast::thread_init(gx);
  Pix pix[PPT];
  double a_c[PPT];
  unsigned a_ok = 0;
  double sa_all = 0.0, saa_all = 0.0;
#pragma unroll
  for (int k = 0; k < PPT; ++k) {
    const int gy = tile_y * TILE_H + ty0 + k * kRowsPerPass;
    a_c[k] = 0.0;
    pix[k] = Fast::dead();
    if (gx < gnx && gy < gny) {
      const int64_t idx = (int64_t)gy * gnx + gx;
      const double a = (double)ref[idx];
      if (isfinite(a)) {
        a_c[k] = a - pivot_a;
        a_ok |= 1u << k;
        pix[k] = Fast::load(planes, idx, gy);
        sa_all += a_c[k];
        saa_all = fma(a_c[k], a_c[k], saa_all);
      }
    }
  }
  const int n_all = __popc(a_ok);
  // warp totals of the lag-independent reference moments (used when no sample of the warp is missing)
  double wsa = sa_all, wsaa = saa_all;
  int wn = n_all;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    wsa += __shfl_xor_sync(0xffffffffu, wsa, o);
    wsaa += __shfl_xor_sync(0xffffffffu, wsaa, o);
    wn += __shfl_xor_sync(0xffffffffu, wn, o);
  }
  if (lane == 0) {
    s_wconst[warp][0] = (double)wn;
    s_wconst[warp][1] = wsa;
    s_wconst[warp][2] = wsaa;
  }

  const int lag_begin = blockIdx.y * lags_per_block;
  const int lag_end = min(n_lags, lag_begin + lags_per_block);
  if (ranges != nullptr) {
    // A Carrington grid is usually far larger than the small image's footprint: when the bounding box of this
    // tile's detector-plane offsets cannot reach the image under any lag of the slice (or the tile has no finite
    // reference pixel), every sample is missing, all six moments are zero, and the lag walk is skipped.
    double bx0 = CUDART_INF, bx1 = -CUDART_INF, by0 = CUDART_INF, by1 = -CUDART_INF;
#pragma unroll
    for (int k = 0; k < PPT; ++k)
      if (a_ok & (1u << k)) {
        const double px = Fast::plane_x(pix[k]), py = Fast::plane_y(pix[k]);
        bx0 = fmin(bx0, px); bx1 = fmax(bx1, px);
        by0 = fmin(by0, py); by1 = fmax(by1, py);
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      bx0 = fmin(bx0, __shfl_xor_sync(0xffffffffu, bx0, o));
      bx1 = fmax(bx1, __shfl_xor_sync(0xffffffffu, bx1, o));
      by0 = fmin(by0, __shfl_xor_sync(0xffffffffu, by0, o));
      by1 = fmax(by1, __shfl_xor_sync(0xffffffffu, by1, o));
    }
    if (lane == 0) { s_box[0][warp] = bx0; s_box[1][warp] = bx1; s_box[2][warp] = by0; s_box[3][warp] = by1; }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < kWarps; ++w) {
      bx0 = fmin(bx0, s_box[0][w]); bx1 = fmax(bx1, s_box[1][w]);
      by0 = fmin(by0, s_box[2][w]); by1 = fmax(by1, s_box[3][w]);
    }
    const double* rg = ranges + 4 * blockIdx.y;   // min / max of x0 + 0.5, y0 + 0.5 over the slice
    // sample valid <=> 0 <= x <= n - 1 <=> 0.5 <= x + 0.5 <= n - 0.5 (one pixel of slack for rounding)
    const bool reach = (bx1 + rg[1] >= -0.5) && (bx0 + rg[0] <= (double)snx + 0.5) &&
                       (by1 + rg[3] >= -0.5) && (by0 + rg[2] <= (double)sny + 0.5);
    if (!reach) {   // also taken when the tile has no live pixel (box stays empty: +inf / -inf)
      for (int i = tid; i < (lag_end - lag_begin) * kMom; i += kThreads)
        work[((size_t)tile * n_lags + lag_begin) * kMom + i] = 0.0;
      return;
    }
  }
  for (int l0 = lag_begin; l0 < lag_end; l0 += kFastLagSub) {
    const int cnt = min(kFastLagSub, lag_end - l0);
    __syncthreads();
    {
      const double* src = reinterpret_cast<const double*>(fast_lags + l0);
      double* dst = reinterpret_cast<double*>(s_lag);
      const int nd = cnt * (int)(sizeof(LagC) / sizeof(double));
      for (int i = tid; i < nd; i += kThreads) dst[i] = src[i];
    }
    __syncthreads();
    for (int l = 0; l < cnt; ++l) {
      const LagC C = s_lag[l];
      const typename Fast::TL tl = Fast::thread_lag(C, tstate);
      double sb = 0.0, sbb = 0.0, sab = 0.0, sa_miss = 0.0, saa_miss = 0.0;
      int n_miss = 0;
#pragma unroll
      for (int g = 0; g < PPT; g += GROUP) {
        // phase A: coordinates (+0.5), floor indices, fractional parts; branch-free
        double sxs[GROUP], sys[GROUP], vx[GROUP], vy[GROUP];
        int ix[GROUP], iy[GROUP];
        bool interior = true;
#pragma unroll
        for (int j = 0; j < GROUP; ++j) {
          Fast::map_half(pix[g + j], tl, C, sxs[j], sys[j]);
          const double mx = __dadd_rd(sxs[j], kMagic), my = __dadd_rd(sys[j], kMagic);
          ix[j] = __double2loint(mx);
          iy[j] = __double2loint(my);
          vx[j] = sxs[j] - (mx - kMagic);   // = d + 0.5 in [0, 1)
          vy[j] = sys[j] - (my - kMagic);
          interior = interior && ((unsigned)(ix[j] - 1) < ux) && ((unsigned)(iy[j] - 1) < uy) &&
                     small_magnitude(sxs[j]) && small_magnitude(sys[j]);
        }
        if (Fast::kCoordsAlwaysFinite) interior = interior && (((a_ok >> g) & ((1u << GROUP) - 1u)) == ((1u << GROUP) - 1u));
        if (interior) {
          // phase B: weights, 9 taps, float32 rounding, moments
#pragma unroll
          for (int j = 0; j < GROUP; ++j) {
            // order-2 B-spline weights from v = d + 0.5: w2 = v^2/2, w0 = w2 - d, w1 = 1 - w0 - w2
            const double wx2 = (0.5 * vx[j]) * vx[j];
            const double wx0 = (wx2 + 0.5) - vx[j];
            const double wx1 = fma(-2.0, wx2, vx[j] + 0.5);
            const double wy2 = (0.5 * vy[j]) * vy[j];
            const double wy0 = (wy2 + 0.5) - vy[j];
            const double wy1 = fma(-2.0, wy2, vy[j] + 0.5);
            // interior => 1 <= ix, iy, so the first tap index is a non-negative 32-bit number
            const unsigned tap0 = (unsigned)(iy[j] - 1) * row_elems + (unsigned)(ix[j] - 1);
            const SmallT* r0p = small + tap0;
            const SmallT* r1p = r0p + row_elems;
            const SmallT* r2p = r1p + row_elems;
            const double r0 = fma(ldval(r0p + 2), wx2, fma(ldval(r0p + 1), wx1, ldval(r0p) * wx0));
            const double r1 = fma(ldval(r1p + 2), wx2, fma(ldval(r1p + 1), wx1, ldval(r1p) * wx0));
            const double r2 = fma(ldval(r2p + 2), wx2, fma(ldval(r2p + 1), wx1, ldval(r2p) * wx0));
            const double t = fma(r2, wy2, fma(r1, wy1, r0 * wy0));
            double b;
            bool ok;
            if (ROUND32) {
              const float bf = __double2float_rn(t);
              ok = isfinite(bf);
              b = (double)bf;
            } else {
              ok = isfinite(t) && (t != -32762.0);
              b = t;
            }
            if (ok) {
              const double bc = b - pivot_b;
              sb += bc;
              sbb = fma(bc, bc, sbb);
              sab = fma(a_c[g + j], bc, sab);
            } else {
              ++n_miss;   // interior => the reference pixel is present
              sa_miss += a_c[g + j];
              saa_miss = fma(a_c[g + j], a_c[g + j], saa_miss);
            }
          }
        } else {
          // exact generic sampler with the same coordinates (image borders, missing reference pixels)
#pragma unroll
          for (int j = 0; j < GROUP; ++j) {
            if (!(a_ok & (1u << (g + j)))) continue;
            double v;
            bool ok = spline_sample<2, false, SmallT>(small, sny, snx, sys[j] - 0.5, sxs[j] - 0.5, v);
            double b;
            if (ROUND32) {
              const float bf = __double2float_rn(v);
              ok = ok && isfinite(bf);
              b = (double)bf;
            } else {
              ok = ok && isfinite(v) && (v != -32762.0);
              b = v;
            }
            if (ok) {
              const double bc = b - pivot_b;
              sb += bc;
              sbb = fma(bc, bc, sbb);
              sab = fma(a_c[g + j], bc, sab);
            } else {
              ++n_miss;
              sa_miss += a_c[g + j];
              saa_miss = fma(a_c[g + j], a_c[g + j], saa_miss);
            }
          }
        }
      }
      if (__any_sync(0xffffffffu, n_miss != 0)) {
        double m[8];
        m[0] = (double)(n_all - n_miss);
        m[1] = sa_all - sa_miss;
        m[2] = sb;
        m[3] = saa_all - saa_miss;
        m[4] = sbb;
        m[5] = sab;
        m[6] = 0.0;
        m[7] = 0.0;
        const double tot = warp_transpose_reduce8(m, lane);
        if ((lane & 3) == 0) s_part[warp][l][lane >> 2] = tot;
        if (lane == 0) s_miss[warp][l] = 1;
      } else {
        // common case: nothing missing in this warp -> only the three lag-dependent sums need the butterfly;
        // n, Sa, Saa are the warp constants
        double m[4];
        m[0] = sb;
        m[1] = sbb;
        m[2] = sab;
        m[3] = 0.0;
        const double tot = warp_transpose_reduce4(m, lane);
        // lanes 0, 8, 16 hold Sb, Sbb, Sab -> slots 2, 4, 5
        if ((lane & 7) == 0 && lane < 24) s_part[warp][l][(lane >> 3) + 2 + (lane != 0)] = tot;
        if (lane == 1) s_miss[warp][l] = 0;
      }
    }
    __syncthreads();
    for (int i = tid; i < cnt * kMom; i += kThreads) {
      const int l = i / kMom, q = i % kMom;
      double s = 0.0;
      if (q < 6) {
        const int c = (q == 0) ? 0 : ((q == 1) ? 1 : ((q == 3) ? 2 : -1));  // slot of a warp constant, or -1
#pragma unroll
        for (int w = 0; w < kWarps; ++w) s += (c >= 0 && !s_miss[w][l]) ? s_wconst[w][c] : s_part[w][l][q];
      }
      work[((size_t)tile * n_lags + (l0 + l)) * kMom + q] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// Column-rolling form of the fused helioprojective lag kernel (order-2 spline, FMA arithmetic, float64 small image).
//
// A thread owns P CONSECUTIVE rows of one common-grid column. Under a candidate header the homography moves that
// column segment almost rigidly: x is constant to a small fraction of a pixel and y advances by one pixel per row,
// so floor(x + 0.5) is shared by the P pixels and floor(y + 0.5) increases by exactly one per row ("regular"
// column). The 3x3 tap windows of consecutive pixels then overlap in two of their three rows: the thread keeps a
// rolling window of three tap rows in registers and loads 3 new taps per pixel instead of 9 (3 (P + 2) / P per pixel
// overall) -- the L1 data pipe, not FP64 issue, bounded the per-pixel kernel. The fractional parts come from the
// shared floors (x - floor_x0, y - (floor_y0 + p)); whether they really lie in [0, 1) is checked per pixel with one
// integer compare on the high word, and a pixel that fails (a floor changed inside the segment: rotated lags, or a
// coordinate within 1e-7 of a half-integer) is re-evaluated by the exact out-of-line sampler, as is every pixel of
// a thread whose window touches the image border or whose reference pixels are not all finite. Results therefore
// equal the per-pixel kernel's up to the rounding of the reciprocal series.
// ---------------------------------------------------------------------------------------------------------
// The P pixels of a regular window, fully unrolled and free of branches, selects and predicates so that the
// independent dependency chains of consecutive pixels interleave (the FP64 pipe has an 8-cycle dependent-issue
// latency). MODE selects the reciprocal series (0: 1 + e + e^2, 1: three-factor product). Validity is tracked for the
// segment as a whole with two integer maxima: `vmax` over the high words of the fractional parts (all in [0, 1) <=>
// vmax < 0x3FF00000) and `bmax` over the magnitude bits of the float32 samples (all finite <=> bmax < 0x7F800000);
// the caller discards the sums and re-evaluates the segment pixel by pixel when either test fails.
template <int MODE, bool ROUND32, int P>
__device__ __forceinline__ void roll_segment(const double* __restrict__ small, unsigned tap, unsigned row_elems,
                                             double be, double bnx, double bny, double he1, double hx1, double hy1,
                                             double inv0, double xoff, double yoff, double pivot_b,
                                             const double (&a_c)[P], double& sb, double& sbb, double& sab,
                                             unsigned& vmax, unsigned& bmax) {
  // A tap row (a, b, c) enters the result only through  a w0 + b w1 + c w2  with the order-2 B-spline weights
  // w0 = (1 - v)^2 / 2, w1 = 1/2 + v - v^2, w2 = v^2 / 2 of v = d + 0.5, i.e. through the quadratic
  //   A + v (B + v C),   A = (a + b) / 2,  B = b - a,  C = (a + c) / 2 - b.
  // A row serves three consecutive pixels of the column (each with its own v), so its coefficients are formed once
  // (5 operations per row) and every pixel evaluates three Horner forms in x (6 FMAs) and, the same way, one in y
  // (5 + 2): 13 + 5 (P + 2) / P operations per pixel instead of 12 for the weights plus 12 for the taps.
  // (Forming the coefficients once per image into three planes was measured slower: three times the L1 footprint.)
  auto row = [&](unsigned t, double& ca, double& cb, double& cc) {
    const double ta = __ldg(small + t), tb = __ldg(small + t + 1), tc = __ldg(small + t + 2);
    ca = 0.5 * (ta + tb);
    cb = tb - ta;
    cc = fma(0.5, ta + tc, -tb);
  };
  double r0a, r0b, r0c, r1a, r1b, r1c, r2a, r2b, r2c;
  row(tap, r0a, r0b, r0c);
  row(tap += row_elems, r1a, r1b, r1c);
  row(tap += row_elems, r2a, r2b, r2c);
#pragma unroll
  for (int p = 0; p < P; ++p) {
    // next row first: it is consumed one pixel later
    double r3a = 0.0, r3b = 0.0, r3c = 0.0;
    if (p + 1 < P) row(tap += row_elems, r3a, r3b, r3c);
    const double e = (p == 0) ? be : fma(he1, (double)p, be);
    const double inv = (p == 0) ? inv0 : ((MODE == 0) ? recip_1me_tiny(e) : recip_1me_small(e));
    const double nx = (p == 0) ? bnx : fma(hx1, (double)p, bnx);
    const double ny = (p == 0) ? bny : fma(hy1, (double)p, bny);
    // fractional parts (+0.5) relative to the shared floors: v = d + 0.5 in [0, 1) on a regular column
    const double vx = fma(nx, inv, xoff);
    const double vy = fma(ny, inv, yoff - (double)p);
    vmax = max(vmax, max((unsigned)__double2hiint(vx), (unsigned)__double2hiint(vy)));
    const double q0 = fma(fma(r0c, vx, r0b), vx, r0a);
    const double q1 = fma(fma(r1c, vx, r1b), vx, r1a);
    const double q2 = fma(fma(r2c, vx, r2b), vx, r2a);
    const double t = fma(fma(fma(0.5, q0 + q2, -q1), vy, q1 - q0), vy, 0.5 * (q0 + q1));
    double b;
    if (ROUND32) {
      // a non-finite float32 sample makes Sbb non-finite: the caller tests that once per segment
      b = (double)__double2float_rn(t);
    } else {
      // finite and not the -32762 fill: fold both into the same flag (the fill marks the sample as missing)
      const unsigned hi = (unsigned)__double2hiint(t) & 0x7FFFFFFFu;
      bmax = max(bmax, (t == -32762.0) ? 0x7F800000u : ((hi >= 0x7FF00000u) ? 0x7F800000u : 0u));
      b = t;
    }
    const double bc = b - pivot_b;
    sb += bc;
    sbb = fma(bc, bc, sbb);
    sab = fma(a_c[p], bc, sab);
    r0a = r1a; r0b = r1b; r0c = r1c;
    r1a = r2a; r1b = r2b; r1c = r2c;
    r2a = r3a; r2b = r3b; r2c = r3c;
  }
}

// The same segment in MIXED arithmetic: coordinates in FP64 (7 instructions per pixel), the spline and the segment's
// moments in FP32 on a float32 copy of the small image. The reference rounds every sample to float32 anyway
// (`alignment.py:1024`), so an FP32 spline differs from it by about one float32 ulp per sample, unbiased; the
// Pearson sums average that over 4e6 samples (|dr| ~ 1e-9 measured, bar 1e-6). A DFMA occupies the dispatch port
// for two cycles and an FFMA for one (profiles/r1_fp64_issue_model.md), so moving the 23 spline / moment
// instructions off the FP64 pipe is what shortens the issue-bound loop.
// The fractional parts leave the FP64 domain without a conversion instruction: the coordinate FMA adds
// kFracMagic = 1.5 * 2^29, whose ulp is 2^-23, so the low word of the result is round((x - floor_x0) * 2^23): bits
// 23.. must be zero for x (p for y: the row index inside the segment), the low 23 bits are the float32 mantissa of
// 1 + frac. `vbad` collects the bits that must be zero; the caller re-evaluates the segment pixel by pixel when
// vbad >= 2^23 (a floor changed inside the segment, or a fraction rounded up to 1).
constexpr double kFracMagic = 805306368.0;   // 1.5 * 2^29

template <int MODE, int P>
__device__ __forceinline__ void roll_segment_mixed(const float* __restrict__ small32, unsigned tap,
                                                   unsigned row_elems, double be, double bnx, double bny, double he1,
                                                   double hx1, double hy1, double inv0, double xoffm, double yoffm,
                                                   float pivot_b, const float (&a_c)[P], float& sb, float& sbb,
                                                   float& sab, unsigned& vbad) {
  // (Measured and dropped, profiles/r1_mixed_kernel.md: a per-image plane of these coefficients, one 16-byte load
  // per row and no arithmetic, is slower -- four times the L1 footprint; a 64-bit row pointer, walked or formed as
  // base + k * stride, compiles to four IADD3 per row where the 32-bit index + IMAD.WIDE form below costs two.)
  auto row = [&](unsigned t, float& ca, float& cb, float& cc) {
    const float ta = __ldg(small32 + t), tb = __ldg(small32 + t + 1), tc = __ldg(small32 + t + 2);
    ca = 0.5f * (ta + tb);
    cb = tb - ta;
    cc = fmaf(0.5f, ta + tc, -tb);
  };
  float r0a, r0b, r0c, r1a, r1b, r1c, r2a, r2b, r2c;
  row(tap, r0a, r0b, r0c);
  row(tap += row_elems, r1a, r1b, r1c);
  row(tap += row_elems, r2a, r2b, r2c);
  double qx0 = 0, qx1 = 0, qx2 = 0, qy0 = 0, qy1 = 0, qy2 = 0;
  if (MODE == 0) {
    const double dinv = fma(be + be, he1, he1);   // d(1 + e + e^2) / dp at the first pixel
    qx0 = fma(bnx, inv0, xoffm);
    qy0 = fma(bny, inv0, yoffm);
    qx1 = fma(hx1, inv0, bnx * dinv);
    qy1 = fma(hy1, inv0, bny * dinv);
    qx2 = hx1 * dinv;
    qy2 = hy1 * dinv;
  }
#pragma unroll
  for (int p = 0; p < P; ++p) {
    float r3a = 0.f, r3b = 0.f, r3c = 0.f;
    if (p + 1 < P) row(tap += row_elems, r3a, r3b, r3c);
    double cx, cy;   // coordinate - shared floor + kFracMagic
    if (MODE == 0) {
      // |e| <= 2^-18 over the whole grid and |he1| < 2.2e-8 (the caller's test): along the segment 1 / (1 - e) is
      // linear in p to p^2 he1^2 < 1.3e-13, so numerator x reciprocal is a quadratic in p whose coefficients are
      // formed once per lag: two FMAs per coordinate instead of seven instructions for both
      cx = (p == 0) ? qx0 : fma(fma(qx2, (double)p, qx1), (double)p, qx0);
      cy = (p == 0) ? qy0 : fma(fma(qy2, (double)p, qy1), (double)p, qy0);
    } else {
      const double e = (p == 0) ? be : fma(he1, (double)p, be);
      const double inv = (p == 0) ? inv0 : recip_1me_small(e);
      const double nx = (p == 0) ? bnx : fma(hx1, (double)p, bnx);
      const double ny = (p == 0) ? bny : fma(hy1, (double)p, bny);
      cx = fma(nx, inv, xoffm);
      cy = fma(ny, inv, yoffm);
    }
    const unsigned ux = (unsigned)__double2loint(cx);
    const unsigned uy = (unsigned)__double2loint(cy) ^ ((unsigned)p << 23);
    vbad |= ux | uy;
    const float vx = __uint_as_float(ux | 0x3F800000u) - 1.0f;
    const float vy = __uint_as_float(uy | 0x3F800000u) - 1.0f;
    const float q0 = fmaf(fmaf(r0c, vx, r0b), vx, r0a);
    const float q1 = fmaf(fmaf(r1c, vx, r1b), vx, r1a);
    const float q2 = fmaf(fmaf(r2c, vx, r2b), vx, r2a);
    // sample - pivot in one go: the pivot rides in the constant term of the y quadratic
    const float bc = fmaf(fmaf(fmaf(0.5f, q0 + q2, -q1), vy, q1 - q0), vy, fmaf(0.5f, q0 + q1, -pivot_b));
    sb += bc;
    sbb = fmaf(bc, bc, sbb);
    sab = fmaf(a_c[p], bc, sab);
    r0a = r1a; r0b = r1b; r0c = r1c;
    r1a = r2a; r1b = r2b; r1c = r2c;
    r2a = r3a; r2b = r3b; r2c = r3c;
  }
}

// One pixel by the per-pixel rules: own floors, 9 taps when they are all inside the image, otherwise the exact
// out-of-line sampler. (sx, sy) are the coordinates + 0.5. Returns false when the sample is missing.
template <bool ROUND32>
__device__ __forceinline__ bool sample_pixel_half(const double* __restrict__ small, int sny, int snx,
                                                  unsigned row_elems, double sx, double sy, double* out) {
  const double mx = __dadd_rd(sx, kMagic), my = __dadd_rd(sy, kMagic);
  const int ix = __double2loint(mx), iy = __double2loint(my);
  const bool interior = small_magnitude(sx) && small_magnitude(sy) && ((unsigned)(ix - 1) < (unsigned)(snx - 2)) &&
                        ((unsigned)(iy - 1) < (unsigned)(sny - 2));
  if (!interior) return sample_exact_half<ROUND32>(small, sny, snx, sy, sx, out);
  const double vx = sx - (mx - kMagic), vy = sy - (my - kMagic);
  const double wx2 = (0.5 * vx) * vx;
  const double wx0 = (wx2 + 0.5) - vx;
  const double wx1 = fma(-2.0, wx2, vx + 0.5);
  const double wy2 = (0.5 * vy) * vy;
  const double wy0 = (wy2 + 0.5) - vy;
  const double wy1 = fma(-2.0, wy2, vy + 0.5);
  const double* r0p = small + ((unsigned)(iy - 1) * row_elems + (unsigned)(ix - 1));
  const double* r1p = r0p + row_elems;
  const double* r2p = r1p + row_elems;
  const double q0 = fma(__ldg(r0p + 2), wx2, fma(__ldg(r0p + 1), wx1, __ldg(r0p) * wx0));
  const double q1 = fma(__ldg(r1p + 2), wx2, fma(__ldg(r1p + 1), wx1, __ldg(r1p) * wx0));
  const double q2 = fma(__ldg(r2p + 2), wx2, fma(__ldg(r2p + 1), wx1, __ldg(r2p) * wx0));
  const double t = fma(q2, wy2, fma(q1, wy1, q0 * wy0));
  if (ROUND32) {
    const float bf = __double2float_rn(t);
    *out = (double)bf;
    return isfinite(bf);
  }
  *out = t;
  return isfinite(t) && (t != -32762.0);
}

// One lag for one thread's column segment: first-pixel coordinates, shared floors, window test, the regular rolling
// segment or -- image borders, irregular columns (rotated lags), missing pixels, division-mode lags -- the segment
// pixel by pixel. Out: the thread's Sb, Sbb, Sab over its valid samples and the mask of pixels that have a finite
// reference value but no valid sample.
template <bool ROUND32, int P, bool MIXED, typename AT>
__device__ __forceinline__ void roll_lag(const HomLag& C, const double* __restrict__ small,
                                         const float* __restrict__ small32, int snx, int sny, unsigned row_elems,
                                         double di, double dj0, const AT (&a_c)[P], unsigned a_ok, bool all_ref,
                                         double pivot_b, double& sb, double& sbb, double& sab, unsigned& miss) {
  const double hx1 = C.hx1, hy1 = C.hy1, he1 = C.he1, x0h = C.x0h, y0h = C.y0h;
  const int mode = (C.emax <= kTinyE) ? 0 : ((C.emax <= kSmallE) ? 1 : 2);  // block-uniform
  // numerators and e = 1 - D of the segment's first pixel (row gy0); pixel p adds p times the row slopes
  const double bnx = fma(hx1, dj0, fma(C.hx0, di, C.hx2));
  const double bny = fma(hy1, dj0, fma(C.hy0, di, C.hy2));
  const double be = fma(he1, dj0, fma(C.he0, di, C.he2));
  const double inv0 = (mode == 0) ? recip_1me_tiny(be) : ((mode == 1) ? recip_1me_small(be) : recip_1me_div(be));
  const double sx0 = fma(bnx, inv0, x0h), sy0 = fma(bny, inv0, y0h);  // coordinates + 0.5
  // floors shared by the segment
  const double mx0 = __dadd_rd(sx0, kMagic), my0 = __dadd_rd(sy0, kMagic);
  const int ix0 = __double2loint(mx0), iy0 = __double2loint(my0);
  const double xoff = x0h - (mx0 - kMagic), yoff = y0h - (my0 - kMagic);
  sb = sbb = sab = 0.0;
  // whole window inside the image: columns ix0-1 .. ix0+1, rows iy0-1 .. iy0+P; |coordinate| < 2^30 keeps the
  // magic-number floor meaningful (NaN fails the compare as well); division-mode lags go pixel by pixel
  bool fast = all_ref && (mode != 2) && small_magnitude(sx0) && small_magnitude(sy0) &&
              ((unsigned)(ix0 - 1) < (unsigned)(snx - 2)) && (iy0 >= 1) && (iy0 + P <= sny - 1);
  if constexpr (MIXED) {
    // the low word of coordinate + kFracMagic holds the offset from the shared floor only while that offset stays
    // below 2^9 pixels: true for any sane lag, guaranteed here by bounding the per-row slopes (block-uniform test)
    fast = fast && (fabs(hx1) < 8.0) && (fabs(hy1) < 8.0);
    if (fast) {
      const unsigned tap = (unsigned)(iy0 - 1) * row_elems + (unsigned)(ix0 - 1);
      // xoff + kFracMagic rounds to 2^-23 pixel; the residual (exact) goes into the numerator, whose factor inv is
      // 1 + O(2^-7): what is lost is below 1e-9 pixel, the same for every pixel of the lag
      const double xoffm = xoff + kFracMagic, yoffm = yoff + kFracMagic;
      const double bnx2 = bnx + (xoff - (xoffm - kFracMagic)), bny2 = bny + (yoff - (yoffm - kFracMagic));
      float fsb = 0.f, fsbb = 0.f, fsab = 0.f;
      unsigned vbad = 0;
      // the quadratic form of the coordinates drops p^2 he1^2 of the reciprocal: below 1e-9 pixel for |he1| < 2.2e-8
      // (P <= 16 rows, numerators below 8192 pixels) -- any 2048-row grid in this mode has |he1| < 4e-9; a small grid
      // with a steep denominator takes the per-pixel series instead (block-uniform choice)
      if (mode == 0 && fabs(he1) < 2.2e-8)
        roll_segment_mixed<0, P>(small32, tap, row_elems, be, bnx2, bny2, he1, hx1, hy1, inv0, xoffm, yoffm,
                                 (float)pivot_b, a_c, fsb, fsbb, fsab, vbad);
      else
        roll_segment_mixed<1, P>(small32, tap, row_elems, be, bnx2, bny2, he1, hx1, hy1, inv0, xoffm, yoffm,
                                 (float)pivot_b, a_c, fsb, fsbb, fsab, vbad);
      // every offset inside its cell, every sample finite (a non-finite sample makes Sbb non-finite; so does a
      // finite sample beyond 1.8e19, which then just takes the exact path)
      fast = (vbad < 0x00800000u) && ((__float_as_uint(fsbb) & 0x7F800000u) != 0x7F800000u);
      sb = (double)fsb;
      sbb = (double)fsbb;
      sab = (double)fsab;
    }
  } else if (fast) {
    const unsigned tap = (unsigned)(iy0 - 1) * row_elems + (unsigned)(ix0 - 1);
    unsigned vmax = 0, bmax = 0;
    if (mode == 0)
      roll_segment<0, ROUND32, P>(small, tap, row_elems, be, bnx, bny, he1, hx1, hy1, inv0, xoff, yoff, pivot_b,
                                  a_c, sb, sbb, sab, vmax, bmax);
    else
      roll_segment<1, ROUND32, P>(small, tap, row_elems, be, bnx, bny, he1, hx1, hy1, inv0, xoff, yoff, pivot_b,
                                  a_c, sb, sbb, sab, vmax, bmax);
    // all fractional parts in [0, 1), every sample finite (|bc| < 2^129 keeps bc^2 finite, so Sbb is finite
    // exactly when all samples are)
    fast = (vmax < 0x3FF00000u) && (bmax < 0x7F800000u) &&
           (((unsigned)__double2hiint(sbb) & 0x7FF00000u) != 0x7FF00000u);
  }
  miss = 0;  // bit p: pixel p has a finite reference value but no valid sample
  if (!fast && a_ok && mode != 2) {
    // A segment that lies outside the small image as a whole has no sample at all. Along the segment each
    // coordinate is numerator / (1 - e) with both linear in p and 1 - e > 0, hence monotonic: it stays between its
    // values at the first and the last pixel. (1e-6 pixel covers the rounding of the reciprocal series; NaN
    // coordinates fail every compare and go pixel by pixel.)
    const double eL = fma(he1, (double)(P - 1), be);
    const double invL = (mode == 0) ? recip_1me_tiny(eL) : recip_1me_small(eL);
    const double sxL = fma(fma(hx1, (double)(P - 1), bnx), invL, x0h);
    const double syL = fma(fma(hy1, (double)(P - 1), bny), invL, y0h);
    const double lo = 0.5 - 1e-6, hix = (double)snx - 0.5 + 1e-6, hiy = (double)sny - 0.5 + 1e-6;
    if ((fmax(sx0, sxL) < lo) || (fmin(sx0, sxL) > hix) || (fmax(sy0, syL) < lo) || (fmin(sy0, syL) > hiy)) {
      sb = sbb = sab = 0.0;
      miss = a_ok;
      return;
    }
  }
  if (!fast && a_ok) {
    // image borders, irregular columns (rotated lags), missing pixels: the segment pixel by pixel
    sb = sbb = sab = 0.0;
#pragma unroll 1
    for (int p = 0; p < P; ++p) {
      if (!(a_ok & (1u << p))) continue;
      const double e = fma(he1, (double)p, be);
      const double inv = (mode == 0) ? recip_1me_tiny(e) : ((mode == 1) ? recip_1me_small(e) : recip_1me_div(e));
      const double sx = fma(fma(hx1, (double)p, bnx), inv, x0h);
      const double sy = fma(fma(hy1, (double)p, bny), inv, y0h);
      AT acs = a_c[0];
#pragma unroll
      for (int q = 1; q < P; ++q)
        if (q == p) acs = a_c[q];
      const double ac = (double)acs;
      double b;
      if (sample_pixel_half<ROUND32>(small, sny, snx, row_elems, sx, sy, &b)) {
        const double bc = b - pivot_b;
        sb += bc;
        sbb = fma(bc, bc, sbb);
        sab = fma(ac, bc, sab);
      } else {
        miss |= 1u << p;
      }
    }
  }
}

#ifndef COREG_ROLL_CHUNK
#define COREG_ROLL_CHUNK 8
#endif
constexpr int kRollChunk = COREG_ROLL_CHUNK;  // lags whose per-lane sums wait in a warp's shared-memory slice for one fold

// ---------------------------------------------------------------------------------------------------------
// The rolling kernel proper. Warps are independent: no block barrier inside the lag walk. Every warp keeps its own
// shared-memory slice (its lanes' Sb, Sbb, Sab for kRollChunk lags, and its own copy of the chunk's 3x3 matrices),
// folds it with __syncwarp only, and writes one 24-byte record per (warp, lag) straight to the workspace; the
// finalize kernel sums 8 records per tile instead of one. Corrections for missing samples (rare) go to a second
// array that is only read where a bit of the per-(tile, lag) mask is set, so it needs no initialisation; the
// mask itself (4 B per tile and lag) is cleared by the launcher.
// ---------------------------------------------------------------------------------------------------------
constexpr int kAccPad = 33;   // lane stride of the per-warp accumulators (bank-conflict-free transposed reads)

struct RollWShared {
  double acc[kWarps][kRollChunk][3][kAccPad];
  HomLag lag[kWarps][kRollChunk];
};

template <typename RefT, bool ROUND32, int P, int MINB, bool MIXED>
__global__ void __launch_bounds__(kThreads, MINB)
lag_corr_roll_kernel(const RefT* __restrict__ ref, const double* __restrict__ small,
                      const float* __restrict__ small32, int snx, int sny, int gnx,
                      int gny, const HomLag* __restrict__ lags, int n_lags, int lags_per_block,
                      const double* __restrict__ pivots, double* __restrict__ wrec, double* __restrict__ wcorr,
                      double* __restrict__ wconst, unsigned* __restrict__ wmask) {
  constexpr int TILE_H = kRowsPerPass * P;
  extern __shared__ __align__(16) unsigned char roll_smem[];
  RollWShared& S = *reinterpret_cast<RollWShared*>(roll_smem);

  const int tiles_x = (gnx + kTileW - 1) / kTileW;
  const int tile = blockIdx.x;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & (kTileW - 1), rg = tid / kTileW;
  const int gx = tile_x * kTileW + tx;
  const int gy0 = tile_y * TILE_H + rg * P;
  // MIXED: float32-representable pivots, so that the float32 segment sums and the FP64 fallback subtract the same
  // numbers and a - pivot_a is (nearly always) exact in float32. r does not depend on the pivots.
  const double pivot_a = MIXED ? (double)(float)pivots[0] : pivots[0];
  const double pivot_b = MIXED ? (double)(float)pivots[1] : pivots[1];
  const unsigned row_elems = (unsigned)snx;
  const double di = (double)gx, dj0 = (double)gy0;
  const size_t wid = (size_t)tile * kWarps + warp;   // this warp's record row

  using AT = typename std::conditional<MIXED, float, double>::type;
  AT a_c[P];
  unsigned a_ok = 0;
  double sa_all = 0.0, saa_all = 0.0;
#pragma unroll
  for (int p = 0; p < P; ++p) {
    const int gy = gy0 + p;
    a_c[p] = (AT)0;
    if (gx < gnx && gy < gny) {
      const double a = (double)ref[(int64_t)gy * gnx + gx];
      if (isfinite(a)) {
        a_c[p] = (AT)(a - pivot_a);
        a_ok |= 1u << p;
        // MIXED: every sum sees the float32 value of a - pivot_a, i.e. one consistent reference image
        const double ac = (double)a_c[p];
        sa_all += ac;
        saa_all = fma(ac, ac, saa_all);
      }
    }
  }
  const bool all_ref = a_ok == ((P >= 32) ? 0xFFFFFFFFu : ((1u << (P & 31)) - 1u));
  if (blockIdx.y == 0) {
    // lag-independent reference moments of this warp's pixels (n, Sa, Saa): one record per warp, written once
    double wsa = sa_all, wsaa = saa_all;
    int wn = __popc(a_ok);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      wsa += __shfl_xor_sync(0xffffffffu, wsa, o);
      wsaa += __shfl_xor_sync(0xffffffffu, wsaa, o);
      wn += __shfl_xor_sync(0xffffffffu, wn, o);
    }
    if (lane == 0) {
      wconst[wid * 3 + 0] = (double)wn;
      wconst[wid * 3 + 1] = wsa;
      wconst[wid * 3 + 2] = wsaa;
    }
  }

  const int lag_begin = blockIdx.y * lags_per_block;
  const int lag_end = min(n_lags, lag_begin + lags_per_block);
  for (int l0 = lag_begin; l0 < lag_end; l0 += kRollChunk) {
    const int cnt = min(kRollChunk, lag_end - l0);
    __syncwarp();   // previous chunk folded
    {
      const double* src = reinterpret_cast<const double*>(lags + l0);
      double* dst = reinterpret_cast<double*>(&S.lag[warp][0]);
      const int nd = cnt * (int)(sizeof(HomLag) / sizeof(double));
      for (int i = lane; i < nd; i += 32) dst[i] = __ldg(src + i);
    }
    __syncwarp();
    for (int l = 0; l < cnt; ++l) {
      double sb, sbb, sab;
      unsigned miss;
      roll_lag<ROUND32, P, MIXED, AT>(S.lag[warp][l], small, small32, snx, sny, row_elems, di, dj0, a_c, a_ok, all_ref,
                                      pivot_b, sb, sbb, sab, miss);
      S.acc[warp][l][0][lane] = sb;
      S.acc[warp][l][1][lane] = sbb;
      S.acc[warp][l][2][lane] = sab;
      if (__any_sync(0xffffffffu, miss != 0)) {
        double m[4] = {(double)__popc(miss), 0.0, 0.0, 0.0};
#pragma unroll
        for (int p = 0; p < P; ++p)
          if (miss & (1u << p)) {
            const double ac = (double)a_c[p];
            m[1] += ac;
            m[2] = fma(ac, ac, m[2]);
          }
        const double tot = warp_transpose_reduce4(m, lane);  // lanes 0, 8, 16: n, Sa, Saa of the missing pixels
        if ((lane & 7) == 0 && lane < 24) wcorr[(wid * n_lags + (l0 + l)) * 3 + (lane >> 3)] = tot;
        if (lane == 0) atomicOr(wmask + ((size_t)tile * n_lags + (l0 + l)), 1u << warp);
      }
    }
    __syncwarp();
    if (lane < cnt * 3) {
      const int l = lane / 3, v = lane - 3 * l;
      const double* src = &S.acc[warp][l][v][0];
      double t = src[0];
#pragma unroll
      for (int k = 1; k < 32; ++k) t += src[k];
      wrec[(wid * n_lags + (l0 + l)) * 3 + v] = t;
    }
  }
}

// one block per lag: sum the warp records in a fixed order, moments -> Pearson r
__global__ void __launch_bounds__(128)
lag_corr_finalize_w_kernel(const double* __restrict__ wrec, const double* __restrict__ wcorr,
                           const double* __restrict__ wconst, const unsigned* __restrict__ wmask, int n_rows,
                           int n_lags, double* __restrict__ corr, int64_t* __restrict__ nvalid) {
  __shared__ double s[128][6];
  const int lag = blockIdx.x;
  double m[6] = {0, 0, 0, 0, 0, 0};   // n, Sa, Sb, Saa, Sbb, Sab
  for (int r = threadIdx.x; r < n_rows; r += 128) {
    const double* p = wrec + ((size_t)r * n_lags + lag) * 3;
    m[2] += p[0];
    m[4] += p[1];
    m[5] += p[2];
    m[0] += wconst[(size_t)r * 3 + 0];
    m[1] += wconst[(size_t)r * 3 + 1];
    m[3] += wconst[(size_t)r * 3 + 2];
    if ((wmask[(size_t)(r / kWarps) * n_lags + lag] >> (r % kWarps)) & 1u) {
      const double* c = wcorr + ((size_t)r * n_lags + lag) * 3;
      m[0] -= c[0];
      m[1] -= c[1];
      m[3] -= c[2];
    }
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) s[threadIdx.x][q] = m[q];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int q = 0; q < 6; ++q) s[threadIdx.x][q] += s[threadIdx.x + o][q];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = s[0][0], sa = s[0][1], sb = s[0][2], saa = s[0][3], sbb = s[0][4], sab = s[0][5];
    double r = CUDART_NAN;
    if (n > 0.0) {
      const double cov = sab - sa * sb / n;
      const double va = saa - sa * sa / n;
      const double vb = sbb - sb * sb / n;
      r = cov / sqrt(va * vb);
    }
    corr[lag] = r;
    if (nvalid) nvalid[lag] = (int64_t)n;
  }
}

// one block per lag: sum the tile partials in a fixed order, moments -> Pearson r
__global__ void __launch_bounds__(128)
lag_corr_finalize_kernel(const double* __restrict__ work, int n_tiles, int n_lags, double* __restrict__ corr,
                         int64_t* __restrict__ nvalid) {
  __shared__ double s[128][6];
  const int lag = blockIdx.x;
  double m[6] = {0, 0, 0, 0, 0, 0};
  for (int t = threadIdx.x; t < n_tiles; t += 128) {
    const double* p = work + ((size_t)t * n_lags + lag) * kMom;
#pragma unroll
    for (int q = 0; q < 6; ++q) m[q] += p[q];
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) s[threadIdx.x][q] = m[q];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int q = 0; q < 6; ++q) s[threadIdx.x][q] += s[threadIdx.x + o][q];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = s[0][0], sa = s[0][1], sb = s[0][2], saa = s[0][3], sbb = s[0][4], sab = s[0][5];
    double r = CUDART_NAN;
    if (n > 0.0) {
      const double cov = sab - sa * sb / n;
      const double va = saa - sa * sa / n;
      const double vb = sbb - sb * sb / n;
      r = cov / sqrt(va * vb);
    }
    corr[lag] = r;
    if (nvalid) nvalid[lag] = (int64_t)n;
  }
}

// grid for a (rows per tile, resident blocks per SM) choice: blockIdx.x = tile, blockIdx.y = slice of the lag list,
// enough slices for a few waves of resident blocks; returns false when the lag list does not fit one launch
inline bool lag_grid(int tile_h, int minb, int gnx, int gny, int64_t n_lags, int sms, dim3* grid, int* lags_per_block,
                     int* tiles_out, int lag_sub, bool amortise_block_setup = false) {
  const int tiles = ((gnx + kTileW - 1) / kTileW) * ((gny + tile_h - 1) / tile_h);
  *tiles_out = tiles;
  // many more blocks than resident slots: border tiles take the per-pixel path and run longer, so a fine
  // granularity keeps the last wave short (COREG_WAVES overrides the default for tuning)
  static const int waves = []() { const char* e = getenv("COREG_WAVES"); const int v = e ? atoi(e) : 0; return v > 0 ? v : 128; }();
  const int want_blocks = sms * minb * waves;
  int splits = (want_blocks + tiles - 1) / tiles;
  const int max_splits = (int)((n_lags + lag_sub - 1) / lag_sub);
  splits = std::max(1, std::min(splits, max_splits));
  // ... but not so fine that a block's own set-up (its pixels of the reference image, their moments: worth about
  // two lags) stops being amortised. Tail ~ 1 / waves and set-up ~ splits / n_lags balance at
  // splits ~ sqrt(1.3 * resident blocks * n_lags / tiles): the 128-wave value for a 3600-lag search of a 2048^2 grid,
  // fewer slices for short lag lists (one rank's share of a sharded search). Measured on 450-lag slices of config 1
  // (tools/shard_lab.py): 29 slices of 16 lags 4.06 ms, 19 of 24 3.95, 12 of 40 3.86, 4 of 120 4.07.
  if (amortise_block_setup) {
    const double s = sqrt(1.315 * (double)(sms * minb) * (double)n_lags / (double)tiles);
    splits = std::max(1, std::min(splits, (int)lround(s)));
  }
  int lpb = (int)((n_lags + splits - 1) / splits);
  lpb = ((lpb + lag_sub - 1) / lag_sub) * lag_sub;
  splits = (int)((n_lags + lpb - 1) / lpb);
  if (splits > 65535) return false;
  *grid = dim3(tiles, splits);
  *lags_per_block = lpb;
  return true;
}

// tuning variants (flags bits 8..11) of the per-pixel fast kernel: (pixels per thread, resident CTAs per SM, group)
template <class Fast, typename SmallT, typename RefT, bool ROUND32>
int launch_lag_fast(int variant, int gnx, int gny, int64_t n_lags, int sms, cudaStream_t s, const RefT* ref,
                    const SmallT* small, int snx, int sny, typename Fast::Planes planes,
                    const typename Fast::LagC* ft, const double* pivots, double* w, int* tiles_out,
                    double* ranges) {
  static const int kVar[3][2] = {{4, 3}, {8, 2}, {4, 4}};
  if (variant < 0 || variant > 2) variant = 0;
  dim3 grid;
  int lpb;
  if (!lag_grid(kRowsPerPass * kVar[variant][0], kVar[variant][1], gnx, gny, n_lags, sms, &grid, &lpb, tiles_out,
                kFastLagSub))
    return fail(COREG_EINVAL, "lag grid too large for one launch");
  if (ranges) offset_lag_range_kernel<<<grid.y, 128, 0, s>>>(ft, (int)n_lags, lpb, ranges);
#define LF(PPT_, MINB_, G_)                                                                     \
  lag_corr_fast_kernel<Fast, SmallT, RefT, ROUND32, PPT_, MINB_, G_><<<grid, kThreads, 0, s>>>( \
      ref, small, snx, sny, gnx, gny, planes, ft, (int)n_lags, lpb, pivots, w, ranges)
  switch (variant) {
    case 1: LF(8, 2, 2); break;
    case 2: LF(4, 4, 2); break;
    default: LF(4, 3, 2); break;
  }
#undef LF
  return COREG_OK;
}

// rolling kernel + its finalize. Tuning variants (flags bits 8..11): rows per thread 12 (default), 16, 14; the
// workspace layout is sized for 12 (fewer rows per thread would need more record rows). small32 != nullptr selects
// the mixed-arithmetic kernel (FP64 coordinates, FP32 spline on the float32 copy of the small image), variants 0 / 1
// = 12 / 16 rows per thread. (Measured and dropped: 3 CTAs per SM at 80 registers -- spills; 24 and 32 rows per
// thread -- spills and partial tiles; a precomputed row-coefficient plane: tools/mixed_lab.py,
// profiles/r1_mixed_kernel.md.)
template <typename RefT, bool ROUND32>
int launch_lag_rollw(int variant, int gnx, int gny, int64_t n_lags, int sms, cudaStream_t s, const RefT* ref,
                     const double* small, const float* small32, int snx, int sny,
                     const HomLag* ft, const double* pivots, void* work, double* corr, int64_t* nvalid, bool prof) {
  const bool mixed = small32 != nullptr;
  const int minb = 2;
  const int rows_per_thread = (variant == 1) ? 16 : ((variant == 2 && !mixed) ? 14 : kRollWRows);
  const RollWLayout L = rollw_layout(gnx, gny, n_lags);
  char* base = static_cast<char*>(work);
  double* wrec = reinterpret_cast<double*>(base + L.rec);
  double* wcorr = reinterpret_cast<double*>(base + L.corr);
  double* wconst = reinterpret_cast<double*>(base + L.cst);
  unsigned* wmask = reinterpret_cast<unsigned*>(base + L.mask);
  dim3 grid;
  int lpb, tiles;
  if (!lag_grid(kRowsPerPass * rows_per_thread, minb, gnx, gny, n_lags, sms, &grid, &lpb, &tiles, kRollChunk, true))
    return fail(COREG_EINVAL, "lag grid too large for one launch");
  CK(cudaMemsetAsync(wmask, 0, (size_t)tiles * (size_t)n_lags * sizeof(unsigned), s));
  if (prof) CK(cudaEventRecord(g_prof[g_prof_n].a, s));
#define RW(P_, MIXED_)                                                                                               \
  {                                                                                                                  \
    auto kern = lag_corr_roll_kernel<RefT, ROUND32, P_, 2, MIXED_>;                                                  \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RollWShared));               \
    kern<<<grid, kThreads, sizeof(RollWShared), s>>>(ref, small, small32, snx, sny, gnx, gny, ft, (int)n_lags, lpb,  \
                                                     pivots, wrec, wcorr, wconst, wmask);                            \
  }
  if (mixed) {
    if (rows_per_thread == 16) RW(16, true) else RW(kRollWRows, true)
  } else {
    if (rows_per_thread == 16) RW(16, false) else if (rows_per_thread == 14) RW(14, false) else RW(kRollWRows, false)
  }
#undef RW
  CK_LAUNCH("lag_corr_roll_kernel");
  if (prof) {
    CK(cudaEventRecord(g_prof[g_prof_n].b, s));
    ++g_prof_n;
  }
  // record rows actually written: 8 warps per tile of this variant (<= L.rows)
  lag_corr_finalize_w_kernel<<<(unsigned)n_lags, 128, 0, s>>>(wrec, wcorr, wconst, wmask, tiles * kWarps, (int)n_lags,
                                                              corr, nvalid);
  CK_LAUNCH("lag_corr_finalize_w_kernel");
  return COREG_OK;
}

template <typename SmallT, typename RefT, bool ROUND32>
int launch_offset_fast(int variant, int gnx, int gny, int64_t n_lags, int sms, cudaStream_t s, const RefT* ref,
                       const SmallT* small, int snx, int sny, OffsetCoord::Planes planes, const CoregLagOffset* lags,
                       const double* pivots, double* w, void* work, int* tiles_out) {
  // the per-lag fast table lives in the tail of the workspace (after the [tiles][lags][8] partials)
  OffsetFastLag* ft = reinterpret_cast<OffsetFastLag*>(static_cast<char*>(work) + partials_bytes(gnx, gny, n_lags));
  offset_fast_table_kernel<<<((int)n_lags + 127) / 128, 128, 0, s>>>(lags, (int)n_lags, ft);
  // per-slice offset ranges right behind the table (the tail reserves 96 B per lag; the table uses 16)
  double* ranges = reinterpret_cast<double*>(ft + n_lags);
  return launch_lag_fast<OffsetFast, SmallT, RefT, ROUND32>(variant, gnx, gny, n_lags, sms, s, ref, small, snx, sny,
                                                            planes, ft, pivots, w, tiles_out, ranges);
}
template <typename SmallT, typename RefT, bool ROUND32>
int launch_offset_fast(int, int, int, int64_t, int, cudaStream_t, const RefT*, const SmallT*, int, int,
                       TanCoord::Planes, const CoregLagTan*, const double*, double*, void*, int*) {
  return fail(COREG_EINVAL, "internal: offset fast path requested for the TAN functor");
}

template <typename SmallT, typename RefT, bool ROUND32>
int launch_offset_fast(int, int, int, int64_t, int, cudaStream_t, const RefT*, const SmallT*, int, int,
                       TanCoord::Planes, const CoregLagCar*, const double*, double*, void*, int*) {
  return fail(COREG_EINVAL, "internal: offset fast path requested for the CAR functor");
}

// generic kernel: variant 1 = 8 pixels per thread, 2 CTAs / SM; anything else 4 pixels per thread, 4 CTAs / SM
template <class Coord, int ORDER, bool STRICT, typename SmallT, typename RefT, bool ROUND32>
int launch_lag_variant(int variant, dim3 grid_tiles_of, int gnx, int gny, int64_t n_lags, int sms, cudaStream_t s,
                       const RefT* ref, const SmallT* small, int snx, int sny, typename Coord::Planes planes,
                       const typename Coord::Lag* lags, const double* pivots, double* w, int* tiles_out) {
  (void)grid_tiles_of;
  const int ppt = (variant == 1) ? 8 : 4, minb = (variant == 1) ? 2 : 4;
  dim3 grid;
  int lags_per_block;
  if (!lag_grid(kRowsPerPass * ppt, minb, gnx, gny, n_lags, sms, &grid, &lags_per_block, tiles_out, kLagSub))
    return fail(COREG_EINVAL, "lag grid too large for one launch");
#define LV(PPT_, MINB_)                                                                                     \
  lag_corr_kernel<Coord, ORDER, STRICT, SmallT, RefT, ROUND32, PPT_, MINB_><<<grid, kThreads, 0, s>>>(      \
      ref, small, snx, sny, gnx, gny, planes, lags, (int)n_lags, lags_per_block, pivots, w)
  if (variant == 1) LV(8, 2); else LV(4, 4);
#undef LV
  return COREG_OK;
}

template <class Coord, typename SmallT, typename RefT, bool ROUND32>
int launch_lag_corr(const RefT* ref, const SmallT* small, int snx, int sny, int gnx, int gny,
                    typename Coord::Planes planes, const typename Coord::Lag* lags, int64_t n_lags, int order,
                    const double* pivots, void* work, size_t work_bytes, double* corr, int64_t* nvalid, int flags,
                    cudaStream_t s) {
  if (n_lags <= 0) return COREG_OK;
  if (n_lags > (int64_t)1 << 30) return fail(COREG_EINVAL, "too many lags in one call (max 2^30)");
  if (gnx <= 0 || gny <= 0 || snx <= 0 || sny <= 0) return fail(COREG_EINVAL, "empty image");
  if ((int64_t)snx * sny >= ((int64_t)1 << 31)) return fail(COREG_EINVAL, "small image too large (>= 2^31 pixels)");
  if (work_bytes < coreg_lag_corr_workspace_bytes(gnx, gny, n_lags)) return fail(COREG_ENOMEM, "workspace too small");
  int sms = coreg_device_sm_count();
  if (sms <= 0) sms = 148;
  const bool strict = (flags & COREG_FLAG_STRICT) != 0;
  const int variant = (flags >> 8) & 15;
  double* w = static_cast<double*>(work);
  const bool prof = g_prof_on && g_prof_n < 4096;
  if (prof) {
    CK(cudaEventCreate(&g_prof[g_prof_n].a));
    CK(cudaEventCreate(&g_prof[g_prof_n].b));
    CK(cudaEventRecord(g_prof[g_prof_n].a, s));
  }
  int tiles = 0, rc = COREG_OK;
  // fast kernel for the offset (Carrington) functor: order 2, FMA arithmetic, image of at least 3x3
  const bool fast_ok = std::is_same<Coord, OffsetCoord>::value && (order == 2) && !strict && snx >= 3 && sny >= 3 &&
                       !(flags & COREG_FLAG_NO_FAST);
  if (fast_ok) {
    rc = launch_offset_fast<SmallT, RefT, ROUND32>(variant, gnx, gny, n_lags, sms, s, ref, small, snx, sny, planes,
                                                   lags, pivots, w, work, &tiles);
    if (rc) return rc;
    CK_LAUNCH("lag_corr_fast_kernel");
    if (prof) {
      CK(cudaEventRecord(g_prof[g_prof_n].b, s));
      ++g_prof_n;
    }
    lag_corr_finalize_kernel<<<(unsigned)n_lags, 128, 0, s>>>(w, tiles, (int)n_lags, corr, nvalid);
    CK_LAUNCH("lag_corr_finalize_kernel");
    return COREG_OK;
  }
#define LAUNCH(ORD, STR)                                                                                          \
  rc = launch_lag_variant<Coord, ORD, STR, SmallT, RefT, ROUND32>(variant, dim3(), gnx, gny, n_lags, sms, s, ref, \
                                                                   small, snx, sny, planes, lags, pivots, w, &tiles)
  switch (order) {
    case 0: LAUNCH(0, true); break;
    case 1: LAUNCH(1, true); break;
    case 2: if (strict) LAUNCH(2, true); else LAUNCH(2, false); break;
    case 3: LAUNCH(3, true); break;
    default: return fail(COREG_EINVAL, "spline order must be 0..3");
  }
#undef LAUNCH
  if (rc) return rc;
  CK_LAUNCH("lag_corr_kernel");
  if (prof) {
    CK(cudaEventRecord(g_prof[g_prof_n].b, s));
    ++g_prof_n;
  }
  lag_corr_finalize_kernel<<<(unsigned)n_lags, 128, 0, s>>>(w, tiles, (int)n_lags, corr, nvalid);
  CK_LAUNCH("lag_corr_finalize_kernel");
  return COREG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Pixel-shift lag search (pxlshift/alignment_pixels.py:35-84): integer (dx, dy) displacements of a window of the
// large image against the (rotated) small image, NaN-masked Pearson per lag.
// One block = one 64 x 32 tile of the small image (8 pixels per thread, in registers, pivot-subtracted) x one chunk
// of up to 8 x 8 (dx, dy) lags x one rotation. The part of the large image the chunk can reach -- the tile grown by
// the chunk's dx / dy spans -- is staged once in shared memory; every lag then reads it at its own offset (LDS.64,
// conflict-free along x). The mask is pairwise, so in general all six moments depend on the lag; when the staged window
// holds no missing pixel (the common case) only Sa, Saa, Sab do and the loop is select-free. Warp butterfly, warps
// folded in a fixed order, one 64 B partial per (tile, lag).
// ---------------------------------------------------------------------------------------------------------
constexpr int kPxTileH = 32;
constexpr int kPxPPT = kPxTileH / kRowsPerPass;   // 8
constexpr int kPxChunk = 8;

__global__ void __launch_bounds__(kThreads)
pixel_shift_corr_kernel(const double* __restrict__ large, int lnx, int lny, const double* __restrict__ smalls, int snx,
                        int sny, int x0, int y0, const int* __restrict__ lag_dx, int n_dx,
                        const int* __restrict__ lag_dy, int n_dy, int n_rot, int chunk, int win_w,
                        const double* __restrict__ pivots, double* __restrict__ work) {
  extern __shared__ double s_win[];
  __shared__ double s_part[kWarps][kPxChunk * kPxChunk][kMom];
  const int tiles_x = (snx + kTileW - 1) / kTileW;
  const int tile = blockIdx.x, tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const int chunks_x = (n_dx + chunk - 1) / chunk;
  const int i0 = (blockIdx.y % chunks_x) * chunk, i1 = min(n_dx, i0 + chunk);
  const int j0 = (blockIdx.y / chunks_x) * chunk, j1 = min(n_dy, j0 + chunk);
  const int rot = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & (kTileW - 1), ty = tid / kTileW;
  const double pivot_a = pivots[0], pivot_b = pivots[1];
  int dxmin = lag_dx[i0], dxmax = dxmin, dymin = lag_dy[j0], dymax = dymin;
  for (int i = i0 + 1; i < i1; ++i) { dxmin = min(dxmin, lag_dx[i]); dxmax = max(dxmax, lag_dx[i]); }
  for (int j = j0 + 1; j < j1; ++j) { dymin = min(dymin, lag_dy[j]); dymax = max(dymax, lag_dy[j]); }
  // stage the reachable window of the large image (pivot-subtracted). Positions outside the image are only ever read
  // by dead pixels (the host checked every lag's slice against the image), they hold 0.
  const int need_w = kTileW + (dxmax - dxmin), need_h = kPxTileH + (dymax - dymin);
  const int gx0 = x0 + tile_x * kTileW + dxmin, gy0 = y0 + tile_y * kPxTileH + dymin;
  int saw_nan = 0;
  for (int idx = tid; idx < need_w * need_h; idx += kThreads) {
    const int wy = idx / need_w, wx = idx - wy * need_w;
    const int gy = gy0 + wy, gx = gx0 + wx;
    double v = 0.0;
    if (gx >= 0 && gx < lnx && gy >= 0 && gy < lny) {
      v = __ldg(large + (size_t)gy * lnx + gx) - pivot_a;
      saw_nan |= !(fabs(v) <= 1.7976931348623157e308);   // NaN or Inf: the general path keeps the reference's semantics
    }
    s_win[wy * win_w + wx] = v;
  }
  // the small tile: value (0 where missing) and 0 / 1 weight per pixel, lag-independent
  double bz[kPxPPT], wb[kPxPPT];
  const double* sm = smalls + (size_t)rot * snx * sny;
  double inv[4] = {0.0, 0.0, 0.0, 0.0};   // n, Sb, Sbb over the live pixels of this thread
#pragma unroll
  for (int k = 0; k < kPxPPT; ++k) {
    const int sx = tile_x * kTileW + tx, sy = tile_y * kPxTileH + ty + k * kRowsPerPass;
    const double b = (sx < snx && sy < sny) ? __ldg(sm + (size_t)sy * snx + sx) - pivot_b : CUDART_NAN;
    const bool live = (b == b);
    bz[k] = live ? b : 0.0;
    wb[k] = live ? 1.0 : 0.0;
    inv[0] += wb[k];
    inv[1] += bz[k];
    inv[2] = fma(bz[k], bz[k], inv[2]);
  }
  const int window_has_nan = __syncthreads_or(saw_nan);   // also the barrier behind the staging
  const int n_lags = n_dx * n_dy * n_rot;
  const int cj = j1 - j0;
  if (!window_has_nan) {
    // no missing pixel of the large image in reach: the mask is the small image's alone, so n, Sb, Sbb do not depend on
    // the lag and the loop carries three sums, without a compare or a select: 1 LDS.64 + 1 DMUL + 3 DFMA per pixel-sample
    for (int i = i0; i < i1; ++i) {
      const int ox = lag_dx[i] - dxmin + tx;
      for (int j = j0; j < j1; ++j) {
        const int oy = lag_dy[j] - dymin + ty;
        double m[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
        for (int k = 0; k < kPxPPT; ++k) {
          const double a = s_win[(oy + k * kRowsPerPass) * win_w + ox];
          const double aw = a * wb[k];
          m[0] += aw;
          m[1] = fma(a, aw, m[1]);
          m[2] = fma(a, bz[k], m[2]);
        }
        const double tot = warp_transpose_reduce4(m, lane);
        if ((lane & 7) == 0) s_part[warp][(i - i0) * cj + (j - j0)][lane >> 3] = tot;
      }
    }
    const double itot = warp_transpose_reduce4(inv, lane);
    if ((lane & 7) == 0) s_part[warp][kPxChunk * kPxChunk - 1][4 + (lane >> 3)] = itot;   // slots 4..7 of the last row
    __syncthreads();
    const int cnt = (i1 - i0) * cj;
    for (int e = tid; e < cnt * kMom; e += kThreads) {
      const int l = e / kMom, q = e % kMom;
      // q: n, Sa, Sb, Saa, Sbb, Sab, pad, pad  <-  invariant 0, varying 0, invariant 1, varying 1, invariant 2, varying 2
      const bool varying = (q == 1 || q == 3 || q == 5);
      const int src = (q == 1) ? 0 : (q == 3) ? 1 : (q == 5) ? 2 : (q == 0) ? 4 : (q == 2) ? 5 : (q == 4) ? 6 : 7;
      const int row = varying ? l : kPxChunk * kPxChunk - 1;
      double acc = 0.0;
      if (q < 6) {
        acc = s_part[0][row][src];
#pragma unroll
        for (int w = 1; w < kWarps; ++w) acc += s_part[w][row][src];
      }
      const int i = i0 + l / cj, j = j0 + l % cj;
      const size_t lag = ((size_t)i * n_dy + j) * n_rot + rot;
      work[((size_t)tile * n_lags + lag) * kMom + q] = acc;
    }
    return;
  }
  for (int i = i0; i < i1; ++i) {
    const int ox = lag_dx[i] - dxmin + tx;
    for (int j = j0; j < j1; ++j) {
      const int oy = lag_dy[j] - dymin + ty;
      double m[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
      for (int k = 0; k < kPxPPT; ++k) {
        const double a = s_win[(oy + k * kRowsPerPass) * win_w + ox];
        const bool ok = (a == a) && (wb[k] != 0.0);          // np.isnan on either side masks the pair
        const double av = ok ? a : 0.0, bv = ok ? bz[k] : 0.0;
        m[0] += ok ? 1.0 : 0.0;
        m[1] += av;
        m[2] += bv;
        m[3] = fma(av, av, m[3]);
        m[4] = fma(bv, bv, m[4]);
        m[5] = fma(av, bv, m[5]);
      }
      const double tot = warp_transpose_reduce8(m, lane);
      if ((lane & 3) == 0) s_part[warp][(i - i0) * cj + (j - j0)][lane >> 2] = tot;
    }
  }
  __syncthreads();
  const int cnt = (i1 - i0) * cj;
  for (int e = tid; e < cnt * kMom; e += kThreads) {
    const int l = e / kMom, q = e % kMom;
    double acc = s_part[0][l][q];
#pragma unroll
    for (int w = 1; w < kWarps; ++w) acc += s_part[w][l][q];
    const int i = i0 + l / cj, j = j0 + l % cj;
    const size_t lag = ((size_t)i * n_dy + j) * n_rot + rot;
    work[((size_t)tile * n_lags + lag) * kMom + q] = acc;
  }
}

// moments -> r with the reference's float32 numerator (pxlshift/c_correlate.py:52-60: the lag products are summed into
// np.zeros(len(lags), dtype="float32")); tile partials folded in a fixed order
__global__ void __launch_bounds__(128)
pixel_shift_finalize_kernel(const double* __restrict__ work, int n_tiles, int n_lags, double* __restrict__ corr,
                            int64_t* __restrict__ nvalid) {
  __shared__ double s[128][6];
  const int lag = blockIdx.x;
  double m[6] = {0, 0, 0, 0, 0, 0};
  for (int t = threadIdx.x; t < n_tiles; t += 128) {
    const double* p = work + ((size_t)t * n_lags + lag) * kMom;
#pragma unroll
    for (int q = 0; q < 6; ++q) m[q] += p[q];
  }
#pragma unroll
  for (int q = 0; q < 6; ++q) s[threadIdx.x][q] = m[q];
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
#pragma unroll
      for (int q = 0; q < 6; ++q) s[threadIdx.x][q] += s[threadIdx.x + o][q];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const double n = s[0][0], sa = s[0][1], sb = s[0][2], saa = s[0][3], sbb = s[0][4], sab = s[0][5];
    double r = CUDART_NAN;
    if (n > 0.0) {
      const double cov = sab - sa * sb / n;
      const double va = saa - sa * sa / n;
      const double vb = sbb - sb * sb / n;
      r = (double)__double2float_rn(cov) / sqrt(va * vb);
    }
    corr[lag] = r;
    if (nvalid) nvalid[lag] = (int64_t)n;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Carrington planes
// ---------------------------------------------------------------------------------------------------------
__global__ void carrington_planes_kernel(CoregCarrington c, double cosb0, double sinb0, double cosr, double sinr,
                                         const double* __restrict__ sinlon, const double* __restrict__ coslon,
                                         int n_lon, const double* __restrict__ sinlat,
                                         const double* __restrict__ coslat, int n_lat, double* __restrict__ tx,
                                         double* __restrict__ ty) {
  const int64_t n = (int64_t)n_lon * n_lat;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % n_lon), j = (int)(idx / n_lon);
    // utils/rectify.py:345-363, numpy evaluation order, no FMA
    const double X = __dmul_rn(coslat[j], sinlon[i]);
    const double Y = sinlat[j];
    const double Z = __dmul_rn(coslat[j], coslon[i]);
    const double zz = __dadd_rn(__dmul_rn(Z, cosb0), __dmul_rn(Y, sinb0));
    const double yy = __dsub_rn(__dmul_rn(Y, cosb0), __dmul_rn(Z, sinb0));
    double ox = CUDART_NAN, oy = CUDART_NAN;
    if (zz >= 0.0) {
      const double y2 = __dsub_rn(__dmul_rn(yy, cosr), __dmul_rn(X, sinr));
      const double x2 = __dadd_rn(__dmul_rn(X, cosr), __dmul_rn(yy, sinr));
      const double z2 = __dsub_rn(c.dist, zz);
      ox = __ddiv_rn(__dmul_rn(__dmul_rn(atan(__ddiv_rn(x2, z2)), kR2D), 3600.0), c.cdelt1);
      oy = __ddiv_rn(__dmul_rn(__dmul_rn(atan(__ddiv_rn(y2, z2)), kR2D), 3600.0), c.cdelt2);
    }
    tx[idx] = ox;
    ty[idx] = oy;
  }
}

// ---------------------------------------------------------------------------------------------------------
// synthetic raster
// ---------------------------------------------------------------------------------------------------------
constexpr int kMaxSynrasFrames = 64;
struct SynrasWcs {
  TanDev w[kMaxSynrasFrames];
};

template <int ORDER, typename T>
__global__ void synras_kernel(const T* __restrict__ frames, int fnx, int fny, const TanDev* __restrict__ wcs,
                              const int* __restrict__ frame_of_col, const double* __restrict__ lng,
                              const double* __restrict__ lat, int n_rows, int n_cols, double* __restrict__ out) {
  const int64_t n = (int64_t)n_rows * n_cols;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int col = (int)(idx % n_cols);
    const int f = frame_of_col[col];
    double v = CUDART_NAN;
    if (f >= 0) {
      double x, y, s;
      tan_world2pix_dev(wcs[f], lng[idx], lat[idx], x, y);
      // interpol2d(dst=None) returns the imager's dtype (utils/Util.py:95-97): float32 frames give float32-rounded
      // samples, which the reference then stores into its float64 raster
      if (spline_sample<ORDER, true, T>(frames + (size_t)f * fnx * fny, fny, fnx, y, x, s))
        v = (sizeof(T) == 4) ? (double)__double2float_rn(s) : s;
    }
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// FITS tiled-image decoder: RICE_1 tiles (+ de-quantisation of floating-point images), one thread per tile.
// Real Solar Orbiter L2 files are RICE tile-compressed; decoding on the device means only the compressed bytes
// (about a quarter of the pixels' size) cross PCIe and the image never exists on the host. Algorithm: cfitsio
// ricecomp.c (fits_rdecomp) and imcompress.c (unquantize_i4r4 / i4r8), restated in oracle/rice.py.
// A tile is sequential by construction (differences + a running bit position), tiles are independent: a 2048^2
// image with row tiles gives 2048 threads of ~2048 pixels each -- about 0.2 ms, against the ~20 ms host decode.
// ---------------------------------------------------------------------------------------------------------
constexpr int kNRandom = 10000;
constexpr int kRiceZeroValue = -2147483646;

struct RiceBits {
  const unsigned char* p;
  const unsigned char* end;
  unsigned long long acc;   // the low n bits are valid, most significant first
  int n;
  // top the accumulator up: single bytes until the pointer is 4-byte aligned, then whole big-endian words
  __device__ __forceinline__ void refill() {
    while (n <= 56 && p < end && (reinterpret_cast<uintptr_t>(p) & 3u)) {
      acc = (acc << 8) | (unsigned long long)*p++;
      n += 8;
    }
    if (n <= 32 && p + 4 <= end) {
      const unsigned w = __byte_perm(*reinterpret_cast<const unsigned*>(p), 0u, 0x0123);
      acc = (acc << 32) | (unsigned long long)w;
      n += 32;
      p += 4;
    } else {
      while (n <= 56 && p < end) {
        acc = (acc << 8) | (unsigned long long)*p++;
        n += 8;
      }
    }
  }
  __device__ __forceinline__ unsigned take(int k) {   // k <= 32; bits past the end of the stream read as zero
    if (n < k) refill();
    if (n < k) {
      acc <<= (k - n);
      n = k;
    }
    n -= k;
    const unsigned v = (unsigned)((acc >> n) & ((1ull << k) - 1ull));
    acc &= (1ull << n) - 1ull;
    return v;
  }
  __device__ __forceinline__ int zeros_then_one() {   // number of zero bits before the next one bit (consumed too)
    int z = 0;
    for (;;) {
      if (n == 0 || acc == 0) {
        z += n;
        n = 0;
        acc = 0;
        refill();
        if (n == 0) return z;   // truncated stream: stop (caller decodes garbage, never reads out of bounds)
        continue;
      }
      const int top = 63 - __clzll((long long)acc);    // position of the highest set bit, < n
      z += n - 1 - top;
      n = top;
      acc &= (1ull << n) - 1ull;
      return z;
    }
  }
};

template <typename TO>
__global__ void rice_tiles_kernel(const unsigned char* __restrict__ heap, const long long* __restrict__ offs,
                                  const int* __restrict__ cnts, int n_tiles, int tiles_x, int tw, int th, int nx, int ny,
                                  int blocksize, int bytepix, const double* __restrict__ zscale,
                                  const double* __restrict__ zzero, int method, int zdither0, int has_blank, int blank,
                                  const float* __restrict__ rnd, TO* __restrict__ out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_tiles) return;
  const int x0 = (t % tiles_x) * tw, y0 = (t / tiles_x) * th;
  const int w = min(tw, nx - x0), h = min(th, ny - y0), npix = w * h;
  const int fsbits = bytepix == 4 ? 5 : (bytepix == 2 ? 4 : 3), fsmax = bytepix == 4 ? 25 : (bytepix == 2 ? 14 : 6),
            bbits = 8 * bytepix;
  RiceBits br;
  br.p = heap + offs[t];
  br.end = br.p + cnts[t];
  br.acc = 0;
  br.n = 0;
  const bool quant = method >= 0;   // floating-point image: de-quantise
  double scale = 1.0, zero = 0.0;
  int iseed = 0, nextrand = 0;
  if (quant) {
    scale = zscale[t];
    zero = zzero[t];
    iseed = (int)(((long long)t + zdither0 - 1) % kNRandom);   // table row n = t + 1: (n + ZDITHER0 - 2) % 10000
    if (iseed < 0) iseed += kNRandom;
    nextrand = (int)(rnd[iseed] * 500.0f);
  }
  int last = (int)br.take(bbits);
  if (bbits < 32) last = (last << (32 - bbits)) >> (32 - bbits);   // sign-extend
  int i = 0;
  while (i < npix) {
    const int fs = (int)br.take(fsbits) - 1;
    const int imax = min(npix, i + blocksize);
    for (; i < imax; ++i) {
      unsigned diff;
      if (fs < 0) {
        diff = 0;
      } else if (fs == fsmax) {
        diff = br.take(bbits);
      } else {
        const int nz = br.zeros_then_one();
        diff = ((unsigned)nz << fs) | br.take(fs);
      }
      const int d = (diff & 1u) ? ~(int)(diff >> 1) : (int)(diff >> 1);
      last = d + last;
      if (bbits < 32) last = (last << (32 - bbits)) >> (32 - bbits);
      const size_t o = (size_t)(y0 + i / w) * nx + (x0 + i % w);
      if (!quant) {
        out[o] = (TO)last;
      } else {
        double v;
        if (has_blank && last == blank) v = CUDART_NAN;
        else if (method == 2 && last == kRiceZeroValue) v = 0.0;
        else if (method == 0) v = (double)last * scale + zero;
        else v = ((double)last - (double)rnd[nextrand] + 0.5) * scale + zero;
        out[o] = (TO)v;
        if (++nextrand == kNRandom) {
          if (++iseed == kNRandom) iseed = 0;
          nextrand = (int)(rnd[iseed] * 500.0f);
        }
      }
    }
  }
}

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

inline int grid_for(int64_t n, int threads = 256) {
  return (int)std::max<int64_t>(1, std::min<int64_t>((n + threads - 1) / threads, 148 * 8));
}

int hpc_lag_corr_wcs_impl(const float* ref, const double* small, const float* small32, int snx, int sny, int gnx,
                          int gny,
                          const CoregTanWcs* grid_wcs, const CoregTanWcs* lag_wcs, int64_t n_lags,
                          const double* pivots, void* work, double* corr, int64_t* nvalid, int flags, cudaStream_t s) {
  HomGrid g;
  int rc = make_hom_grid(grid_wcs, gnx, gny, &g);
  if (rc) return rc;
  int sms = coreg_device_sm_count();
  if (sms <= 0) sms = 148;
  HomLag* ft = reinterpret_cast<HomLag*>(static_cast<char*>(work) + partials_bytes(gnx, gny, n_lags));
  tan_homography_kernel<<<((int)n_lags + 127) / 128, 128, 0, s>>>(g, lag_wcs, (int)n_lags, ft);
  CK_LAUNCH("tan_homography_kernel");
  const bool prof = g_prof_on && g_prof_n < 4096;
  if (prof) {
    CK(cudaEventCreate(&g_prof[g_prof_n].a));
    CK(cudaEventCreate(&g_prof[g_prof_n].b));
  }
  return launch_lag_rollw<float, true>((flags >> 8) & 15, gnx, gny, n_lags, sms, s, ref, small, small32, snx, sny, ft,
                                       pivots, work, corr, nvalid, prof);
}

}  // namespace

// =============================================================================================================
// C ABI
// =============================================================================================================
extern "C" {

const char* coreg_last_error(void) { return g_err; }
int coreg_version(void) { return 100; }

int coreg_device_sm_count(void) {
  int dev = 0, sms = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return COREG_ECUDA;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return COREG_ECUDA;
  return sms;
}

int coreg_tan_pix2world(const CoregTanWcs* wcs, int nx, int ny, int wrap_pipi, double* lng, double* lat,
                        void* stream) {
  TanDev t;
  int rc = make_tan(wcs, &t);
  if (rc) return rc;
  if (nx <= 0 || ny <= 0) return COREG_OK;
  if (!lng || !lat) return fail(COREG_EINVAL, "null output plane");
  const int64_t n = (int64_t)nx * ny;
  tan_pix2world_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(t, nx, ny, wrap_pipi, lng, lat);
  CK_LAUNCH("tan_pix2world_kernel");
  return COREG_OK;
}

int coreg_tan_world2pix(const CoregTanWcs* wcs, const double* lng, const double* lat, int64_t n, double* x,
                        double* y, void* stream) {
  TanDev t;
  int rc = make_tan(wcs, &t);
  if (rc) return rc;
  if (n <= 0) return COREG_OK;
  if (!lng || !lat || !x || !y) return fail(COREG_EINVAL, "null pointer");
  tan_world2pix_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(t, lng, lat, n, x, y);
  CK_LAUNCH("tan_world2pix_kernel");
  return COREG_OK;
}

int coreg_map_coordinates(const void* img, int img_dtype, int img_ny, int img_nx, const double* y, const double* x,
                          int64_t n, int order, double cval, void* out, int out_dtype, void* stream) {
  if (n <= 0) return COREG_OK;
  if (!img || !y || !x || !out) return fail(COREG_EINVAL, "null pointer");
  if (img_ny <= 0 || img_nx <= 0) return fail(COREG_EINVAL, "empty image");
  cudaStream_t s = (cudaStream_t)stream;
  if (img_dtype == COREG_F32 && out_dtype == COREG_F32)
    return launch_map_coordinates((const float*)img, img_ny, img_nx, y, x, n, order, cval, (float*)out, s);
  if (img_dtype == COREG_F32 && out_dtype == COREG_F64)
    return launch_map_coordinates((const float*)img, img_ny, img_nx, y, x, n, order, cval, (double*)out, s);
  if (img_dtype == COREG_F64 && out_dtype == COREG_F32)
    return launch_map_coordinates((const double*)img, img_ny, img_nx, y, x, n, order, cval, (float*)out, s);
  if (img_dtype == COREG_F64 && out_dtype == COREG_F64)
    return launch_map_coordinates((const double*)img, img_ny, img_nx, y, x, n, order, cval, (double*)out, s);
  return fail(COREG_EINVAL, "dtype must be COREG_F32 or COREG_F64");
}

int coreg_tan_trig_planes(const double* lng, const double* lat, int64_t n, double alpha_ref_deg, double* planes,
                          void* stream) {
  if (n <= 0) return COREG_OK;
  if (!lng || !lat || !planes) return fail(COREG_EINVAL, "null pointer");
  tan_trig_planes_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(lng, lat, n, alpha_ref_deg * kD2R, planes);
  CK_LAUNCH("tan_trig_planes_kernel");
  return COREG_OK;
}

int coreg_widen_f32(const float* in, int64_t n, double* out, void* stream) {
  if (n <= 0) return COREG_OK;
  if (!in || !out) return fail(COREG_EINVAL, "coreg_widen_f32: null pointer");
  f32_to_f64_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(in, n, out);
  CK_LAUNCH("f32_to_f64_kernel");
  return COREG_OK;
}

int coreg_finite_mean(const void* img, int dtype, int64_t n, double* mean, void* stream) {
  if (!img || !mean || n <= 0) return fail(COREG_EINVAL, "coreg_finite_mean: bad argument");
  if (dtype == COREG_F32)
    finite_mean_kernel<float><<<1, 1024, 0, (cudaStream_t)stream>>>((const float*)img, n, mean);
  else if (dtype == COREG_F64)
    finite_mean_kernel<double><<<1, 1024, 0, (cudaStream_t)stream>>>((const double*)img, n, mean);
  else
    return fail(COREG_EINVAL, "dtype must be COREG_F32 or COREG_F64");
  CK_LAUNCH("finite_mean_kernel");
  return COREG_OK;
}

size_t coreg_lag_corr_workspace_bytes(int gnx, int gny, int64_t n_lags) {
  if (gnx <= 0 || gny <= 0 || n_lags <= 0) return 0;
  return partials_bytes(gnx, gny, n_lags) + (size_t)n_lags * sizeof(HomLag);
}

int coreg_hpc_lag_corr(const float* ref, const void* small, int small_dtype, int snx, int sny, int gnx, int gny,
                       const double* planes, const CoregLagTan* lags, int64_t n_lags, int order,
                       const double* pivots, void* work, size_t work_bytes, double* corr, int64_t* nvalid, int flags,
                       void* stream) {
  if (!ref || !small || !planes || !lags || !pivots || !work || !corr)
    return fail(COREG_EINVAL, "coreg_hpc_lag_corr: null pointer");
  TanCoord::Planes pl{planes, (int64_t)gnx * gny};
  cudaStream_t s = (cudaStream_t)stream;
  if (small_dtype == COREG_F64)
    return launch_lag_corr<TanCoord, double, float, true>(ref, (const double*)small, snx, sny, gnx, gny, pl, lags,
                                                          n_lags, order, pivots, work, work_bytes, corr, nvalid,
                                                          flags, s);
  if (small_dtype == COREG_F32)
    return launch_lag_corr<TanCoord, float, float, true>(ref, (const float*)small, snx, sny, gnx, gny, pl, lags,
                                                         n_lags, order, pivots, work, work_bytes, corr, nvalid, flags,
                                                         s);
  return fail(COREG_EINVAL, "small_dtype must be COREG_F32 or COREG_F64");
}

int coreg_hpc_lag_corr_wcs(const float* ref, const double* small, int snx, int sny, int gnx, int gny,
                           const CoregTanWcs* grid_wcs, const CoregTanWcs* lag_wcs, int64_t n_lags, int order,
                           const double* pivots, void* work, size_t work_bytes, double* corr, int64_t* nvalid,
                           int flags, void* stream) {
  if (!ref || !small || !grid_wcs || !lag_wcs || !pivots || !work || !corr)
    return fail(COREG_EINVAL, "coreg_hpc_lag_corr_wcs: null pointer");
  if (n_lags <= 0) return COREG_OK;
  if (order != 2 || (flags & COREG_FLAG_STRICT))
    return fail(COREG_EINVAL, "coreg_hpc_lag_corr_wcs: only order 2 with FMA arithmetic; use coreg_hpc_lag_corr");
  if (gnx <= 0 || gny <= 0 || snx < 3 || sny < 3) return fail(COREG_EINVAL, "image too small for the fast kernel");
  if ((int64_t)snx * sny >= ((int64_t)1 << 31)) return fail(COREG_EINVAL, "small image too large (>= 2^31 pixels)");
  if (n_lags > (int64_t)1 << 30) return fail(COREG_EINVAL, "too many lags in one call (max 2^30)");
  if (work_bytes < coreg_lag_corr_workspace_bytes(gnx, gny, n_lags)) return fail(COREG_ENOMEM, "workspace too small");
  return hpc_lag_corr_wcs_impl(ref, small, nullptr, snx, sny, gnx, gny, grid_wcs, lag_wcs, n_lags, pivots, work, corr,
                               nvalid, flags, (cudaStream_t)stream);
}

int coreg_hpc_lag_corr_wcs_mixed(const float* ref, const double* small, const float* small32, int snx, int sny,
                                 int gnx, int gny, const CoregTanWcs* grid_wcs, const CoregTanWcs* lag_wcs,
                                 int64_t n_lags, int order, const double* pivots, void* work, size_t work_bytes,
                                 double* corr, int64_t* nvalid, int flags, void* stream) {
  if (!ref || !small || !small32 || !grid_wcs || !lag_wcs || !pivots || !work || !corr)
    return fail(COREG_EINVAL, "coreg_hpc_lag_corr_wcs_mixed: null pointer");
  if (n_lags <= 0) return COREG_OK;
  if (order != 2 || (flags & COREG_FLAG_STRICT))
    return fail(COREG_EINVAL, "coreg_hpc_lag_corr_wcs_mixed: only order 2 with FMA arithmetic; use coreg_hpc_lag_corr");
  if (gnx <= 0 || gny <= 0 || snx < 3 || sny < 3) return fail(COREG_EINVAL, "image too small for the fast kernel");
  if ((int64_t)snx * sny >= ((int64_t)1 << 31)) return fail(COREG_EINVAL, "small image too large (>= 2^31 pixels)");
  if (n_lags > (int64_t)1 << 30) return fail(COREG_EINVAL, "too many lags in one call (max 2^30)");
  if (work_bytes < coreg_lag_corr_workspace_bytes(gnx, gny, n_lags)) return fail(COREG_ENOMEM, "workspace too small");
  return hpc_lag_corr_wcs_impl(ref, small, small32, snx, sny, gnx, gny, grid_wcs, lag_wcs, n_lags, pivots, work, corr,
                               nvalid, flags, (cudaStream_t)stream);
}

int coreg_tan_homography_emax(const CoregTanWcs* grid_wcs, int gnx, int gny, const CoregTanWcs* lag_wcs, int64_t n_lags,
                              void* scratch, double* emax, void* stream) {
  if (!grid_wcs || !lag_wcs || !scratch || !emax) return fail(COREG_EINVAL, "coreg_tan_homography_emax: null pointer");
  if (n_lags <= 0) return COREG_OK;
  HomGrid g;
  int rc = make_hom_grid(grid_wcs, gnx, gny, &g);
  if (rc) return rc;
  cudaStream_t s = (cudaStream_t)stream;
  HomLag* ft = static_cast<HomLag*>(scratch);
  tan_homography_kernel<<<((int)n_lags + 127) / 128, 128, 0, s>>>(g, lag_wcs, (int)n_lags, ft);
  CK_LAUNCH("tan_homography_kernel");
  CK(cudaMemcpy2DAsync(emax, sizeof(double), &ft[0].emax, sizeof(HomLag), sizeof(double), (size_t)n_lags,
                       cudaMemcpyDeviceToDevice, s));
  return COREG_OK;
}

int coreg_rice_decode(const unsigned char* heap, const long long* offsets, const int* counts, int n_tiles, int tile_w,
                      int tile_h, int nx, int ny, int blocksize, int bytepix, const double* zscale, const double* zzero,
                      int method, int zdither0, int has_blank, int blank, const float* rand_values, void* out,
                      int out_dtype, void* stream) {
  if (!heap || !offsets || !counts || !out) return fail(COREG_EINVAL, "coreg_rice_decode: null pointer");
  if (nx <= 0 || ny <= 0 || tile_w <= 0 || tile_h <= 0 || blocksize <= 0)
    return fail(COREG_EINVAL, "coreg_rice_decode: bad geometry");
  if (bytepix != 1 && bytepix != 2 && bytepix != 4) return fail(COREG_EINVAL, "coreg_rice_decode: BYTEPIX must be 1, 2 or 4");
  const int tiles_x = (nx + tile_w - 1) / tile_w, tiles_y = (ny + tile_h - 1) / tile_h;
  if (n_tiles != tiles_x * tiles_y) return fail(COREG_EINVAL, "coreg_rice_decode: tile count does not match the geometry");
  if (method >= 0 && (!zscale || !zzero || (method > 0 && !rand_values)))
    return fail(COREG_EINVAL, "coreg_rice_decode: quantised image needs ZSCALE, ZZERO and the dither sequence");
  if (method > 2) return fail(COREG_EINVAL, "coreg_rice_decode: unknown ZQUANTIZ method");
  cudaStream_t s = (cudaStream_t)stream;
  const int threads = 64, blocks = (n_tiles + threads - 1) / threads;
#define RICE(T)                                                                                                   \
  rice_tiles_kernel<T><<<blocks, threads, 0, s>>>(heap, offsets, counts, n_tiles, tiles_x, tile_w, tile_h, nx, ny, \
                                                  blocksize, bytepix, zscale, zzero, method, zdither0, has_blank,  \
                                                  blank, rand_values, (T*)out)
  if (out_dtype == COREG_F32 && method >= 0) RICE(float);
  else if (out_dtype == COREG_F64 && method >= 0) RICE(double);
  else if (out_dtype == COREG_I32 && method < 0) RICE(int);
  else return fail(COREG_EINVAL, "coreg_rice_decode: out_dtype must be COREG_I32 for integer images, COREG_F32 / F64 for quantised ones");
#undef RICE
  CK_LAUNCH("rice_tiles_kernel");
  return COREG_OK;
}

int coreg_carrington_planes(const CoregCarrington* c, const double* sinlon, const double* coslon, int n_lon,
                            const double* sinlat, const double* coslat, int n_lat, double* tx, double* ty,
                            void* stream) {
  if (!c || !sinlon || !coslon || !sinlat || !coslat || !tx || !ty)
    return fail(COREG_EINVAL, "coreg_carrington_planes: null pointer");
  if (n_lon <= 0 || n_lat <= 0) return COREG_OK;
  const int64_t n = (int64_t)n_lon * n_lat;
  carrington_planes_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(
      *c, cos(c->lat0), sin(c->lat0), cos(c->roll), sin(c->roll), sinlon, coslon, n_lon, sinlat, coslat, n_lat, tx,
      ty);
  CK_LAUNCH("carrington_planes_kernel");
  return COREG_OK;
}

int coreg_offset_lag_corr(const double* ref, const void* small, int small_dtype, int snx, int sny, int gnx, int gny,
                          const double* tx, const double* ty, const CoregLagOffset* lags, int64_t n_lags, int order,
                          const double* pivots, void* work, size_t work_bytes, double* corr, int64_t* nvalid,
                          int flags, void* stream) {
  if (!ref || !small || !tx || !ty || !lags || !pivots || !work || !corr)
    return fail(COREG_EINVAL, "coreg_offset_lag_corr: null pointer");
  OffsetCoord::Planes pl{tx, ty};
  cudaStream_t s = (cudaStream_t)stream;
  if (small_dtype == COREG_F64)
    return launch_lag_corr<OffsetCoord, double, double, false>(ref, (const double*)small, snx, sny, gnx, gny, pl,
                                                               lags, n_lags, order, pivots, work, work_bytes, corr,
                                                               nvalid, flags, s);
  if (small_dtype == COREG_F32)
    return launch_lag_corr<OffsetCoord, float, double, false>(ref, (const float*)small, snx, sny, gnx, gny, pl, lags,
                                                              n_lags, order, pivots, work, work_bytes, corr, nvalid,
                                                              flags, s);
  return fail(COREG_EINVAL, "small_dtype must be COREG_F32 or COREG_F64");
}

static int car_forward(const CoregLagCar* m, double* f) {
  if (!m) return fail(COREG_EINVAL, "null CoregLagCar");
  const double det = m->m11 * m->m22 - m->m12 * m->m21;
  if (!(det != 0.0) || det != det) return fail(COREG_EINVAL, "singular CDELT*PC matrix");
  f[0] = m->m22 / det;
  f[1] = -m->m12 / det;
  f[2] = -m->m21 / det;
  f[3] = m->m11 / det;
  return COREG_OK;
}

int coreg_car_pix2world(const CoregLagCar* map, int nx, int ny, double* lng, double* lat, void* stream) {
  double f[4];
  int rc = car_forward(map, f);
  if (rc) return rc;
  if (nx <= 0 || ny <= 0) return COREG_OK;
  if (!lng || !lat) return fail(COREG_EINVAL, "null output plane");
  const int64_t n = (int64_t)nx * ny;
  car_pix2world_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*map, f[0], f[1], f[2], f[3], nx, ny, lng, lat);
  CK_LAUNCH("car_pix2world_kernel");
  return COREG_OK;
}

int coreg_car_world2pix(const CoregLagCar* map, const double* lng, const double* lat, int64_t n, double* x, double* y,
                        void* stream) {
  if (!map) return fail(COREG_EINVAL, "null CoregLagCar");
  if (n <= 0) return COREG_OK;
  if (!lng || !lat || !x || !y) return fail(COREG_EINVAL, "null pointer");
  car_world2pix_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(*map, lng, lat, n, x, y);
  CK_LAUNCH("car_world2pix_kernel");
  return COREG_OK;
}

int coreg_car_lag_corr(const float* ref, const void* small, int small_dtype, int snx, int sny, int gnx, int gny,
                       const double* planes, const CoregLagCar* lags, int64_t n_lags, int order, const double* pivots,
                       void* work, size_t work_bytes, double* corr, int64_t* nvalid, int flags, void* stream) {
  if (!ref || !small || !planes || !lags || !pivots || !work || !corr)
    return fail(COREG_EINVAL, "coreg_car_lag_corr: null pointer");
  CarCoord::Planes pl{planes, (int64_t)gnx * gny};
  cudaStream_t s = (cudaStream_t)stream;
  if (small_dtype == COREG_F64)
    return launch_lag_corr<CarCoord, double, float, true>(ref, (const double*)small, snx, sny, gnx, gny, pl, lags,
                                                          n_lags, order, pivots, work, work_bytes, corr, nvalid,
                                                          flags, s);
  if (small_dtype == COREG_F32)
    return launch_lag_corr<CarCoord, float, float, true>(ref, (const float*)small, snx, sny, gnx, gny, pl, lags,
                                                         n_lags, order, pivots, work, work_bytes, corr, nvalid, flags,
                                                         s);
  return fail(COREG_EINVAL, "small_dtype must be COREG_F32 or COREG_F64");
}

static inline size_t px_tiles(int snx, int sny) {
  return (size_t)((snx + kTileW - 1) / kTileW) * ((sny + kPxTileH - 1) / kPxTileH);
}

size_t coreg_pixel_shift_workspace_bytes(int snx, int sny, int n_dx, int n_dy, int n_rot) {
  if (snx <= 0 || sny <= 0 || n_dx <= 0 || n_dy <= 0 || n_rot <= 0) return 0;
  const size_t n_lags = (size_t)n_dx * n_dy * n_rot;
  return px_tiles(snx, sny) * n_lags * kMom * sizeof(double) + (size_t)(n_dx + n_dy + 2) * sizeof(int);
}

int coreg_pixel_shift_corr(const double* large, int lnx, int lny, const double* smalls, int n_rot, int snx, int sny,
                           int x0, int y0, const int* lag_dx, int n_dx, const int* lag_dy, int n_dy,
                           const double* pivots, void* work, size_t work_bytes, double* corr, int64_t* nvalid,
                           void* stream) {
  if (!large || !smalls || !lag_dx || !lag_dy || !pivots || !work || !corr)
    return fail(COREG_EINVAL, "coreg_pixel_shift_corr: null pointer");
  if (n_dx <= 0 || n_dy <= 0 || n_rot <= 0) return COREG_OK;
  if (lnx <= 0 || lny <= 0 || snx <= 0 || sny <= 0) return fail(COREG_EINVAL, "empty image");
  if ((int64_t)n_dx * n_dy * n_rot > ((int64_t)1 << 30)) return fail(COREG_EINVAL, "too many lags in one call");
  if (n_rot > 65535) return fail(COREG_EINVAL, "too many rotation lags in one call");
  if (work_bytes < coreg_pixel_shift_workspace_bytes(snx, sny, n_dx, n_dy, n_rot))
    return fail(COREG_ENOMEM, "workspace too small");
  // `_check_boundaries` (pxlshift/alignment_pixels.py:150-156)
  for (int i = 0; i < n_dx; ++i)
    if (x0 + lag_dx[i] < 0 || x0 + lag_dx[i] + snx > lnx) return fail(COREG_EINVAL, "too large shift : outside FSI");
  for (int j = 0; j < n_dy; ++j)
    if (y0 + lag_dy[j] < 0 || y0 + lag_dy[j] + sny > lny) return fail(COREG_EINVAL, "too large shift : outside FSI");
  // chunks of 8 x 8 lags share one staged window unless the lag arrays are so sparse that it would not fit
  int chunk = kPxChunk, span_x = 0, span_y = 0;
  for (int pass = 0; pass < 2; ++pass) {
    span_x = span_y = 0;
    for (int i0 = 0; i0 < n_dx; i0 += chunk) {
      int lo = lag_dx[i0], hi = lo;
      for (int i = i0; i < std::min(n_dx, i0 + chunk); ++i) { lo = std::min(lo, lag_dx[i]); hi = std::max(hi, lag_dx[i]); }
      span_x = std::max(span_x, hi - lo);
    }
    for (int j0 = 0; j0 < n_dy; j0 += chunk) {
      int lo = lag_dy[j0], hi = lo;
      for (int j = j0; j < std::min(n_dy, j0 + chunk); ++j) { lo = std::min(lo, lag_dy[j]); hi = std::max(hi, lag_dy[j]); }
      span_y = std::max(span_y, hi - lo);
    }
    if ((size_t)(kTileW + span_x) * (kPxTileH + span_y) * sizeof(double) <= 160 * 1024) break;
    chunk = 1;
  }
  const int win_w = kTileW + span_x;
  const size_t smem = (size_t)win_w * (kPxTileH + span_y) * sizeof(double);
  const int tiles = (int)px_tiles(snx, sny);
  const int64_t n_lags = (int64_t)n_dx * n_dy * n_rot;
  const int64_t chunks = (int64_t)((n_dx + chunk - 1) / chunk) * ((n_dy + chunk - 1) / chunk);
  if (chunks > 65535) return fail(COREG_EINVAL, "too many lag chunks in one call");
  cudaStream_t s = (cudaStream_t)stream;
  double* w = static_cast<double*>(work);
  int* d_dx = reinterpret_cast<int*>(w + (size_t)tiles * n_lags * kMom);
  int* d_dy = d_dx + n_dx;
  CK(cudaMemcpyAsync(d_dx, lag_dx, (size_t)n_dx * sizeof(int), cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(d_dy, lag_dy, (size_t)n_dy * sizeof(int), cudaMemcpyHostToDevice, s));
  CK(cudaFuncSetAttribute(pixel_shift_corr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const bool prof = g_prof_on && g_prof_n < 4096;
  if (prof) {
    CK(cudaEventCreate(&g_prof[g_prof_n].a));
    CK(cudaEventCreate(&g_prof[g_prof_n].b));
    CK(cudaEventRecord(g_prof[g_prof_n].a, s));
  }
  pixel_shift_corr_kernel<<<dim3(tiles, (unsigned)chunks, n_rot), kThreads, smem, s>>>(
      large, lnx, lny, smalls, snx, sny, x0, y0, d_dx, n_dx, d_dy, n_dy, n_rot, chunk, win_w, pivots, w);
  CK_LAUNCH("pixel_shift_corr_kernel");
  if (prof) {
    CK(cudaEventRecord(g_prof[g_prof_n].b, s));
    ++g_prof_n;
  }
  pixel_shift_finalize_kernel<<<(unsigned)n_lags, 128, 0, s>>>(w, tiles, (int)n_lags, corr, nvalid);
  CK_LAUNCH("pixel_shift_finalize_kernel");
  return COREG_OK;
}

int coreg_synras_build(const void* frames, int frame_dtype, int n_frames, int fnx, int fny, const CoregTanWcs* wcs,
                       const int* frame_of_col, const double* lng, const double* lat, int n_rows, int n_cols,
                       int order, double* out, void* stream) {
  if (!frames || !wcs || !frame_of_col || !lng || !lat || !out)
    return fail(COREG_EINVAL, "coreg_synras_build: null pointer");
  if (n_frames <= 0 || n_frames > kMaxSynrasFrames) return fail(COREG_EINVAL, "n_frames must be in 1..64 per call");
  if (n_rows <= 0 || n_cols <= 0) return COREG_OK;
  if (order < 0 || order > 3) return fail(COREG_EINVAL, "spline order must be 0..3");
  cudaStream_t s = (cudaStream_t)stream;
  TanDev hw[kMaxSynrasFrames];
  for (int f = 0; f < n_frames; ++f) {
    int rc = make_tan(wcs + f, hw + f);
    if (rc) return rc;
  }
  for (int c = 0; c < n_cols; ++c)
    if (frame_of_col[c] >= n_frames) return fail(COREG_EINVAL, "frame_of_col entry out of range");
  TanDev* dw = nullptr;
  int* dcol = nullptr;
  CK(cudaMallocAsync(&dw, sizeof(TanDev) * n_frames, s));
  CK(cudaMallocAsync(&dcol, sizeof(int) * n_cols, s));
  CK(cudaMemcpyAsync(dw, hw, sizeof(TanDev) * n_frames, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dcol, frame_of_col, sizeof(int) * n_cols, cudaMemcpyHostToDevice, s));
  // hw / frame_of_col are pageable: the async copies above have consumed them when the call returns
  const int64_t n = (int64_t)n_rows * n_cols;
  const int g = grid_for(n);
#define SYN(ORD, T) \
  synras_kernel<ORD, T><<<g, 256, 0, s>>>((const T*)frames, fnx, fny, dw, dcol, lng, lat, n_rows, n_cols, out)
  if (frame_dtype == COREG_F32) {
    switch (order) { case 0: SYN(0, float); break; case 1: SYN(1, float); break; case 2: SYN(2, float); break; default: SYN(3, float); }
  } else if (frame_dtype == COREG_F64) {
    switch (order) { case 0: SYN(0, double); break; case 1: SYN(1, double); break; case 2: SYN(2, double); break; default: SYN(3, double); }
  } else {
    return fail(COREG_EINVAL, "frame_dtype must be COREG_F32 or COREG_F64");
  }
#undef SYN
  CK_LAUNCH("synras_kernel");
  CK(cudaFreeAsync(dw, s));
  CK(cudaFreeAsync(dcol, s));
  return COREG_OK;
}

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
  void *d_large = nullptr, *d_small_in = nullptr;
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
  TRY(cudaMalloc(&d_piv, 2 * sizeof(double)));
  TRY(cudaMalloc(&d_corr, n_lags * sizeof(double)));
  TRY(cudaMalloc(&d_nv, n_lags * sizeof(int64_t)));
  TRY(cudaMalloc(&d_lagw, n_lags * sizeof(CoregTanWcs)));
  TRY(cudaMalloc(&d_work, work_bytes));
  TRY(cudaMemcpyAsync(d_large, large, nl * lsz, cudaMemcpyHostToDevice, s));
  if (small_dtype == COREG_F64) {
    TRY(cudaMemcpyAsync(d_small, small, ns * sizeof(double), cudaMemcpyHostToDevice, s));
  } else {
    // the lag kernels run fastest on float64 storage (no per-tap conversion): widen once on the device
    TRY(cudaMalloc(&d_small_in, ns * sizeof(float)));
    TRY(cudaMemcpyAsync(d_small_in, small, ns * sizeof(float), cudaMemcpyHostToDevice, s));
    f32_to_f64_kernel<<<grid_for(ns), 256, 0, s>>>((const float*)d_small_in, ns, d_small);
  }
  TRY(cudaMemcpyAsync(d_lagw, lag_wcs, n_lags * sizeof(CoregTanWcs), cudaMemcpyHostToDevice, s));
  TRYRC(coreg_tan_pix2world(wcs_small, snx, sny, 1, d_lng, d_lat, s));
  TRYRC(coreg_tan_world2pix(wcs_large, d_lng, d_lat, ns, d_x, d_y, s));
  TRYRC(coreg_map_coordinates(d_large, large_dtype, lny, lnx, d_y, d_x, ns, order, (double)NAN, d_ref, COREG_F32, s));
  TRYRC(coreg_finite_mean(d_ref, COREG_F32, ns, d_piv, s));
  TRYRC(coreg_finite_mean(d_small, COREG_F64, ns, d_piv + 1, s));
  if (fast && (flags & COREG_FLAG_MIXED) && d_small_in) {
    TRYRC(coreg_hpc_lag_corr_wcs_mixed(d_ref, d_small, (const float*)d_small_in, snx, sny, snx, sny, wcs_small, d_lagw,
                                       n_lags, order, d_piv, d_work, work_bytes, d_corr, d_nv, flags, s));
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
  cudaFree(d_lagw); cudaFree(d_lags); cudaFree(d_work);
#undef TRY
#undef TRYRC
  return rc;
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
