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
// fused lag search
// ---------------------------------------------------------------------------------------------------------
constexpr int kTileW = 64;
constexpr int kThreads = 256;
constexpr int kRowsPerPass = kThreads / kTileW;  // 4 grid rows per pass of the block
constexpr int kWarps = kThreads / 32;
constexpr int kLagSub = 64;   // lags staged in shared memory at a time
constexpr int kMom = 8;       // n, Sa, Sb, Saa, Sbb, Sab, pad, pad  (64 B per (tile, lag) partial)
constexpr int kMinTileH = 8;  // smallest tile height of any variant (workspace sizing)

inline size_t partials_bytes(int gnx, int gny, int64_t n_lags) {
  const size_t tiles = (size_t)((gnx + kTileW - 1) / kTileW) * ((gny + kMinTileH - 1) / kMinTileH);
  return tiles * (size_t)n_lags * kMom * sizeof(double);
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
// Fast variant of the fused lag kernel: order-2 spline, FMA arithmetic, float64 small image.
//  * coordinates come out already offset by +0.5 so that floor(x + 0.5) is one magic-number add;
//  * TAN: the reciprocal of the gnomonic denominator D = cos(angular distance to the lag's reference point) is the
//    product form of the geometric series (1+e)(1+e^2)(1+e^4), e = 1 - D, exact to 2^-56 for |e| <= 2^-7
//    (7.1 deg) -- the caller guarantees that bound (COREG_FLAG_SMALL_ANGLE, checked on the host from the FOV);
//  * "strictly interior" (all taps inside the image, so no closed-bound test and no mirroring) is one unsigned
//    integer compare per axis on the floor index; the PPT pixels of a thread take the branch-free path together,
//    anything else (image borders, missing reference pixels, NaN coordinates) falls back to the exact generic
//    per-pixel code, so results differ from the generic kernel only by FMA-level rounding.
// ---------------------------------------------------------------------------------------------------------
// Fast-path lag constants of the helioprojective search: with u = (sin lat, cos lat sin A, cos lat cos A) the unit
// vector of a pixel's sky direction (the three trig planes), the gnomonic map of a lag's header is one 3x3 matrix
// per lag followed by a perspective divide:  (nx, ny, D) = R u,  x = x0 + nx / D,  y = y0 + ny / D.
// R is built once per lag from CoregLagTan by tan_fast_table_kernel (rows: nx, ny, D; then x0 + 0.5, y0 + 0.5).
struct TanFastLag {
  double a0, a1, a2, b0, b1, b2, d0, d1, d2, x0h, y0h, pad;
};

__global__ void tan_fast_table_kernel(const CoregLagTan* __restrict__ lags, int n, TanFastLag* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const CoregLagTan L = lags[i];
  TanFastLag f;
  // qs = p1 cda - p2 sda ; pc = p2 cda + p1 sda ; D = pc cd0 + p0 sd0 ; en = p0 cd0 - pc sd0
  // nx = m11 qs + m12 en ; ny = m21 qs + m22 en
  f.a0 = L.m12 * L.cos_d0;
  f.a1 = L.m11 * L.cos_da - L.m12 * L.sin_da * L.sin_d0;
  f.a2 = -L.m11 * L.sin_da - L.m12 * L.cos_da * L.sin_d0;
  f.b0 = L.m22 * L.cos_d0;
  f.b1 = L.m21 * L.cos_da - L.m22 * L.sin_da * L.sin_d0;
  f.b2 = -L.m21 * L.sin_da - L.m22 * L.cos_da * L.sin_d0;
  f.d0 = L.sin_d0;
  f.d1 = L.sin_da * L.cos_d0;
  f.d2 = L.cos_da * L.cos_d0;
  f.x0h = L.x0 + 0.5;
  f.y0h = L.y0 + 0.5;
  f.pad = 0.0;
  out[i] = f;
}

struct TanFast {
  typedef TanCoord Base;
  typedef TanFastLag LagC;
  __device__ static __forceinline__ void map_half(const TanCoord::Pix& q, const LagC& L, double& sx, double& sy) {
    const double den = fma(q.p2, L.d2, fma(q.p1, L.d1, q.p0 * L.d0));
    const double nx = fma(q.p2, L.a2, fma(q.p1, L.a1, q.p0 * L.a0));
    const double ny = fma(q.p2, L.b2, fma(q.p1, L.b1, q.p0 * L.b0));
    // 1/D for |1 - D| <= 2^-7: (1+e)(1+e^2)(1+e^4), e = 1 - D
    const double e = 1.0 - den;
    const double e2 = e * e;
    double inv = 1.0 + e;
    inv = fma(e2, inv, inv);
    const double e4 = e2 * e2;
    inv = fma(e4, inv, inv);
    sx = fma(nx, inv, L.x0h);
    sy = fma(ny, inv, L.y0h);
  }
};

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

struct OffsetFast {
  typedef OffsetCoord Base;
  typedef OffsetFastLag LagC;
  __device__ static __forceinline__ void map_half(const OffsetCoord::Pix& q, const LagC& L, double& sx, double& sy) {
    sx = L.x0h + q.tx;
    sy = L.y0h + q.ty;
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

constexpr int kFastLagSub = 16;  // lags per shared-memory stage of the fast kernel (small: leaves L1 to the taps)

template <class Fast, typename SmallT, typename RefT, bool ROUND32, int PPT, int MINB, int GROUP>
__global__ void __launch_bounds__(kThreads, MINB)
lag_corr_fast_kernel(const RefT* __restrict__ ref, const SmallT* __restrict__ small, int snx, int sny, int gnx, int gny,
                     typename Fast::Base::Planes planes, const typename Fast::Base::Lag* __restrict__ lags,
                     const typename Fast::LagC* __restrict__ fast_lags, int n_lags, int lags_per_block,
                     const double* __restrict__ pivots, double* __restrict__ work) {
  typedef typename Fast::Base Coord;
  typedef typename Coord::Lag Lag;
  typedef typename Coord::Pix Pix;
  typedef typename Fast::LagC LagC;
  constexpr int TILE_H = kRowsPerPass * PPT;
  // GROUP = pixels whose dependency chains are interleaved (their coordinates / indices are live together)
  static_assert(PPT % GROUP == 0, "PPT must be a multiple of GROUP");
  __shared__ __align__(16) LagC s_lag[kFastLagSub];
  __shared__ double s_part[kWarps][kFastLagSub][kMom];

  const int tiles_x = (gnx + kTileW - 1) / kTileW;
  const int tile = blockIdx.x;
  const int tile_x = tile % tiles_x, tile_y = tile / tiles_x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx = tid & (kTileW - 1), ty0 = tid / kTileW;
  const int gx = tile_x * kTileW + tx;
  const double pivot_a = pivots[0], pivot_b = pivots[1];
  const unsigned ux = (unsigned)(snx - 2), uy = (unsigned)(sny - 2);  // launcher guarantees snx, sny >= 3
  const size_t row_bytes = (size_t)snx * sizeof(SmallT);

  Pix pix[PPT];
  double a_c[PPT];
  unsigned a_ok = 0;
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

  const int lag_begin = blockIdx.y * lags_per_block;
  const int lag_end = min(n_lags, lag_begin + lags_per_block);
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
      double sb = 0.0, sbb = 0.0, sab = 0.0, sa_miss = 0.0, saa_miss = 0.0;
      int n_miss = 0;
#pragma unroll
      for (int g = 0; g < PPT; g += GROUP) {
        // phase A: coordinates (+0.5), floor indices, fractional parts; branch-free
        double vx[GROUP], vy[GROUP];
        int ix[GROUP], iy[GROUP];
        bool interior = true;
#pragma unroll
        for (int j = 0; j < GROUP; ++j) {
          double sx, sy;
          Fast::map_half(pix[g + j], C, sx, sy);
          const double mx = __dadd_rd(sx, kMagic), my = __dadd_rd(sy, kMagic);
          ix[j] = __double2loint(mx);
          iy[j] = __double2loint(my);
          vx[j] = sx - (mx - kMagic);   // = d + 0.5 in [0, 1)
          vy[j] = sy - (my - kMagic);
          interior = interior && ((unsigned)(ix[j] - 1) < ux) && ((unsigned)(iy[j] - 1) < uy);
        }
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
            const char* p0 = reinterpret_cast<const char*>(small + ((iy[j] - 1) * snx + (ix[j] - 1)));
            const SmallT* r0p = reinterpret_cast<const SmallT*>(p0);
            const SmallT* r1p = reinterpret_cast<const SmallT*>(p0 + row_bytes);
            const SmallT* r2p = reinterpret_cast<const SmallT*>(p0 + 2 * row_bytes);
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
              ++n_miss;   // interior => the reference pixel is present (dead pixels carry NaN coordinates)
              sa_miss += a_c[g + j];
              saa_miss = fma(a_c[g + j], a_c[g + j], saa_miss);
            }
          }
        } else {
          // generic exact path (image borders, missing reference pixels)
          const Lag L = lags[l0 + l];
#pragma unroll
          for (int j = 0; j < GROUP; ++j) {
            double x, y, v;
            Coord::map(pix[g + j], L, x, y);
            bool ok = spline_sample<2, false, SmallT>(small, sny, snx, y, x, v);
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
            } else if (a_ok & (1u << (g + j))) {
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
      } else {
        // common case: nothing missing in this warp -> only the three lag-dependent sums need the butterfly
        double m[4];
        m[0] = sb;
        m[1] = sbb;
        m[2] = sab;
        m[3] = 0.0;
        const double tot = warp_transpose_reduce4(m, lane);
        if ((lane & 7) == 0) {
          const int q = lane >> 3;  // 0: Sb, 1: Sbb, 2: Sab, 3: unused
          if (q < 3) s_part[warp][l][q == 0 ? 2 : (q == 1 ? 4 : 5)] = tot;
        }
        if (lane == 1) s_part[warp][l][0] = (double)wn;
        if (lane == 2) s_part[warp][l][1] = wsa;
        if (lane == 3) s_part[warp][l][3] = wsaa;
        if (lane == 4) s_part[warp][l][6] = 0.0;
        if (lane == 5) s_part[warp][l][7] = 0.0;
      }
    }
    __syncthreads();
    for (int i = tid; i < cnt * kMom; i += kThreads) {
      const int l = i / kMom, q = i % kMom;
      double s = s_part[0][l][q];
#pragma unroll
      for (int w = 1; w < kWarps; ++w) s += s_part[w][l][q];
      work[((size_t)tile * n_lags + (l0 + l)) * kMom + q] = s;
    }
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

// tuning variants (flags bits 8..11): tile height = 4*PPT, MINB resident blocks per SM
template <class Coord> struct FastOf;
template <> struct FastOf<TanCoord> { typedef TanFast type; };
template <> struct FastOf<OffsetCoord> { typedef OffsetFast type; };

// grid for a (ppt, minb) variant; returns false when the lag list does not fit one launch
inline bool lag_grid(int ppt, int minb, int gnx, int gny, int64_t n_lags, int sms, dim3* grid, int* lags_per_block,
                     int* tiles_out) {
  const int tile_h = kRowsPerPass * ppt;
  const int tiles = ((gnx + kTileW - 1) / kTileW) * ((gny + tile_h - 1) / tile_h);
  *tiles_out = tiles;
  const int want_blocks = sms * minb * 4;   // a few waves of resident blocks
  int splits = (want_blocks + tiles - 1) / tiles;
  const int max_splits = (int)((n_lags + kLagSub - 1) / kLagSub);
  splits = std::max(1, std::min(splits, max_splits));
  int lpb = (int)((n_lags + splits - 1) / splits);
  lpb = ((lpb + kLagSub - 1) / kLagSub) * kLagSub;
  splits = (int)((n_lags + lpb - 1) / lpb);
  if (splits > 65535) return false;
  *grid = dim3(tiles, splits);
  *lags_per_block = lpb;
  return true;
}

inline void build_fast_table(const CoregLagTan* lags, int n, TanFastLag* out, cudaStream_t s) {
  tan_fast_table_kernel<<<(n + 127) / 128, 128, 0, s>>>(lags, n, out);
}
inline void build_fast_table(const CoregLagOffset* lags, int n, OffsetFastLag* out, cudaStream_t s) {
  offset_fast_table_kernel<<<(n + 127) / 128, 128, 0, s>>>(lags, n, out);
}

template <class Coord, typename SmallT, typename RefT, bool ROUND32>
int launch_lag_fast(int variant, int gnx, int gny, int64_t n_lags, int sms, cudaStream_t s, const RefT* ref,
                    const SmallT* small, int snx, int sny, typename Coord::Planes planes,
                    const typename Coord::Lag* lags, const double* pivots, double* w, void* fast_table,
                    int* tiles_out) {
  typedef typename FastOf<Coord>::type Fast;
  typedef typename Fast::LagC LagC;
  // (pixels per thread, resident CTAs per SM, interleave group); 0 = fastest measured on config 1
  static const int kVar[10][3] = {{4, 3, 2}, {8, 2, 2}, {4, 2, 2}, {2, 3, 2}, {4, 4, 2},
                                  {2, 4, 2}, {6, 2, 2}, {6, 3, 2}, {4, 2, 4}, {4, 3, 1}};
  if (variant < 0 || variant > 9) variant = 0;
  const int ppt = kVar[variant][0], minb = kVar[variant][1];
  dim3 grid;
  int lpb;
  if (!lag_grid(ppt, minb, gnx, gny, n_lags, sms, &grid, &lpb, tiles_out))
    return fail(COREG_EINVAL, "lag grid too large for one launch");
  LagC* ft = static_cast<LagC*>(fast_table);
  build_fast_table(lags, (int)n_lags, ft, s);
#define LF(PPT_, MINB_, G_)                                                                     \
  lag_corr_fast_kernel<Fast, SmallT, RefT, ROUND32, PPT_, MINB_, G_><<<grid, kThreads, 0, s>>>( \
      ref, small, snx, sny, gnx, gny, planes, lags, ft, (int)n_lags, lpb, pivots, w)
  switch (variant) {
    case 1: LF(8, 2, 2); break;
    case 2: LF(4, 2, 2); break;
    case 3: LF(2, 3, 2); break;
    case 4: LF(4, 4, 2); break;
    case 5: LF(2, 4, 2); break;
    case 6: LF(6, 2, 2); break;
    case 7: LF(6, 3, 2); break;
    case 8: LF(4, 2, 4); break;
    case 9: LF(4, 3, 1); break;
    default: LF(4, 3, 2); break;
  }
#undef LF
  return COREG_OK;
}

template <class Coord, int ORDER, bool STRICT, typename SmallT, typename RefT, bool ROUND32>
int launch_lag_variant(int variant, dim3 grid_tiles_of, int gnx, int gny, int64_t n_lags, int sms, cudaStream_t s,
                       const RefT* ref, const SmallT* small, int snx, int sny, typename Coord::Planes planes,
                       const typename Coord::Lag* lags, const double* pivots, double* w, int* tiles_out) {
  (void)grid_tiles_of;
  int ppt, minb;
  switch (variant) {  // measured on config 1 (profiles/r1_k1_tuning.md): 0 is the fastest
    case 1: ppt = 8; minb = 2; break;
    case 2: ppt = 8; minb = 3; break;
    case 3: ppt = 4; minb = 3; break;
    default: ppt = 4; minb = 4; break;
  }
  const int tile_h = kRowsPerPass * ppt;
  const int tiles = ((gnx + kTileW - 1) / kTileW) * ((gny + tile_h - 1) / tile_h);
  *tiles_out = tiles;
  // split the lag list over blockIdx.y until the grid has a few waves of resident blocks
  const int want_blocks = sms * minb * 4;
  int splits = (want_blocks + tiles - 1) / tiles;
  const int max_splits = (int)((n_lags + kLagSub - 1) / kLagSub);
  splits = std::max(1, std::min(splits, max_splits));
  int lags_per_block = (int)((n_lags + splits - 1) / splits);
  lags_per_block = ((lags_per_block + kLagSub - 1) / kLagSub) * kLagSub;
  splits = (int)((n_lags + lags_per_block - 1) / lags_per_block);
  if (splits > 65535) return fail(COREG_EINVAL, "lag grid too large for one launch");
  dim3 grid(tiles, splits);
#define LV(PPT_, MINB_)                                                                                     \
  lag_corr_kernel<Coord, ORDER, STRICT, SmallT, RefT, ROUND32, PPT_, MINB_><<<grid, kThreads, 0, s>>>(      \
      ref, small, snx, sny, gnx, gny, planes, lags, (int)n_lags, lags_per_block, pivots, w)
  switch (variant) {
    case 1: LV(8, 2); break;
    case 2: LV(8, 3); break;
    case 3: LV(4, 3); break;
    default: LV(4, 4); break;
  }
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
  // fast kernel: order 2, FMA arithmetic, float64 small image of at least 3x3; TAN additionally needs the
  // caller's small-angle guarantee (the offset functor has no reciprocal)
  const bool fast_ok = (order == 2) && !strict && snx >= 3 && sny >= 3 &&
                       !(flags & COREG_FLAG_NO_FAST) &&
                       (std::is_same<Coord, OffsetCoord>::value || (flags & COREG_FLAG_SMALL_ANGLE));
  if (fast_ok) {
    // the per-lag fast table lives in the tail of the workspace (after the [tiles][lags][8] partials)
    char* tail = static_cast<char*>(work) + partials_bytes(gnx, gny, n_lags);
    rc = launch_lag_fast<Coord, SmallT, RefT, ROUND32>(variant, gnx, gny, n_lags, sms, s, ref, small, snx, sny, planes,
                                                       lags, pivots, w, tail, &tiles);
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
  return partials_bytes(gnx, gny, n_lags) + (size_t)n_lags * sizeof(TanFastLag);
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

int coreg_hpc_search_host(const double* large, int lnx, int lny, const CoregTanWcs* wcs_large, const double* small,
                          int snx, int sny, const CoregTanWcs* wcs_small, const CoregLagTan* lags, int64_t n_lags,
                          int order, int flags, double* corr, int64_t* nvalid) {
  if (!large || !small || !wcs_large || !wcs_small || !lags || !corr)
    return fail(COREG_EINVAL, "coreg_hpc_search_host: null pointer");
  if (lnx <= 0 || lny <= 0 || snx <= 0 || sny <= 0 || n_lags <= 0)
    return fail(COREG_EINVAL, "coreg_hpc_search_host: empty input");
  const int64_t ns = (int64_t)snx * sny, nl = (int64_t)lnx * lny;
  const size_t work_bytes = coreg_lag_corr_workspace_bytes(snx, sny, n_lags);
  double *d_large = nullptr, *d_small = nullptr, *d_lng = nullptr, *d_lat = nullptr, *d_x = nullptr, *d_y = nullptr,
         *d_planes = nullptr, *d_piv = nullptr, *d_corr = nullptr;
  float* d_ref = nullptr;
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
  TRY(cudaMalloc(&d_large, nl * sizeof(double)));
  TRY(cudaMalloc(&d_small, ns * sizeof(double)));
  TRY(cudaMalloc(&d_lng, ns * sizeof(double)));
  TRY(cudaMalloc(&d_lat, ns * sizeof(double)));
  TRY(cudaMalloc(&d_x, ns * sizeof(double)));
  TRY(cudaMalloc(&d_y, ns * sizeof(double)));
  TRY(cudaMalloc(&d_planes, 3 * ns * sizeof(double)));
  TRY(cudaMalloc(&d_ref, ns * sizeof(float)));
  TRY(cudaMalloc(&d_piv, 2 * sizeof(double)));
  TRY(cudaMalloc(&d_corr, n_lags * sizeof(double)));
  TRY(cudaMalloc(&d_nv, n_lags * sizeof(int64_t)));
  TRY(cudaMalloc(&d_lags, n_lags * sizeof(CoregLagTan)));
  TRY(cudaMalloc(&d_work, work_bytes));
  TRY(cudaMemcpyAsync(d_large, large, nl * sizeof(double), cudaMemcpyHostToDevice, s));
  TRY(cudaMemcpyAsync(d_small, small, ns * sizeof(double), cudaMemcpyHostToDevice, s));
  TRY(cudaMemcpyAsync(d_lags, lags, n_lags * sizeof(CoregLagTan), cudaMemcpyHostToDevice, s));
  TRYRC(coreg_tan_pix2world(wcs_small, snx, sny, 1, d_lng, d_lat, s));
  TRYRC(coreg_tan_world2pix(wcs_large, d_lng, d_lat, ns, d_x, d_y, s));
  TRYRC(coreg_map_coordinates(d_large, COREG_F64, lny, lnx, d_y, d_x, ns, order, (double)NAN, d_ref, COREG_F32, s));
  TRYRC(coreg_tan_trig_planes(d_lng, d_lat, ns, wcs_small->crval1, d_planes, s));
  TRYRC(coreg_finite_mean(d_ref, COREG_F32, ns, d_piv, s));
  TRYRC(coreg_finite_mean(d_small, COREG_F64, ns, d_piv + 1, s));
  TRYRC(coreg_hpc_lag_corr(d_ref, d_small, COREG_F64, snx, sny, snx, sny, d_planes, d_lags, n_lags, order, d_piv,
                           d_work, work_bytes, d_corr, d_nv, flags, s));
  TRY(cudaMemcpyAsync(corr, d_corr, n_lags * sizeof(double), cudaMemcpyDeviceToHost, s));
  if (nvalid) TRY(cudaMemcpyAsync(nvalid, d_nv, n_lags * sizeof(int64_t), cudaMemcpyDeviceToHost, s));
  TRY(cudaStreamSynchronize(s));
done:
  cudaFree(d_large); cudaFree(d_small); cudaFree(d_lng); cudaFree(d_lat); cudaFree(d_x); cudaFree(d_y);
  cudaFree(d_planes); cudaFree(d_ref); cudaFree(d_piv); cudaFree(d_corr); cudaFree(d_nv); cudaFree(d_lags);
  cudaFree(d_work);
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
