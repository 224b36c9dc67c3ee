// coreg_wcs.cu -- one-shot kernels: TAN / CAR pixel <-> world, map_coordinates, pivots, Carrington planes, synthetic raster.
#include "coreg_common.cuh"

namespace coreg {
__device__ __forceinline__ double wrap_pipi_deg(double a) {
  // -((-a + 180) % 360 - 180) with Python's floor-mod (utils/Util.py:76-80)
  double m = fmod(-a + 180.0, 360.0);
  if (m != 0.0 && m < 0.0) m += 360.0;
  return -(m - 180.0);
}

__device__ __forceinline__ void tan_pix2world_dev(const TanDev& w, int i, int j, int wrap, double& lo, double& la) {
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
  lo = w.a0_deg + atan2(yy, xx) * kR2D;
  if (w.a0_deg >= 0.0) {
    if (lo < 0.0) lo += 360.0;
  } else {
    if (lo > 0.0) lo -= 360.0;
  }
  la = atan2(zz, sqrt(xx * xx + yy * yy)) * kR2D;
  if (wrap) {
    lo = wrap_pipi_deg(lo);
    la = wrap_pipi_deg(la);
  }
}

__global__ void tan_pix2world_kernel(TanDev w, int nx, int ny, int wrap, double* __restrict__ lng,
                                     double* __restrict__ lat) {
  const int64_t n = (int64_t)nx * ny;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double lo, la;
    tan_pix2world_dev(w, (int)(idx % nx), (int)(idx / nx), wrap, lo, la);
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
// The one-time cut of a helioprojective search in ONE kernel (`_create_submap_of_large_data`, alignment.py:987-1016):
// world coordinates of the unshifted small grid (ang2pipi-wrapped), their pixel coordinates in the large image, the
// integer origin of the uploaded window taken off, the spline sample, float32 store. The same device functions as the
// three separate kernels (tan_pix2world / tan_world2pix / map_coordinates), so the same bits -- without the four
// float64 planes (134 MB written and read back for a 2048^2 grid) and five launches in between.
// ---------------------------------------------------------------------------------------------------------
template <int ORDER, typename TI>
__global__ void hpc_cut_kernel(TanDev ws, int nx, int ny, TanDev wl, const TI* __restrict__ large, int lny, int lnx,
                               double x0, double y0, double cval, float* __restrict__ out) {
  const int64_t n = (int64_t)nx * ny;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double lo, la, x, y, v;
    tan_pix2world_dev(ws, (int)(idx % nx), (int)(idx / nx), 1, lo, la);
    tan_world2pix_dev(wl, lo, la, x, y);
    x -= x0;
    y -= y0;
    if (!spline_sample<ORDER, true, TI>(large, lny, lnx, y, x, v)) v = cval;
    out[idx] = (float)v;
  }
}

template <typename TI>
int launch_hpc_cut(const TanDev& ws, int nx, int ny, const TanDev& wl, const TI* large, int lny, int lnx, int x0, int y0,
                   int order, float* out, cudaStream_t s) {
  const int64_t n = (int64_t)nx * ny;
  const unsigned g = grid_for(n);
  const double cval = (double)NAN;   // the quiet NaN of the host (`cval=np.nan`): the bits coreg_map_coordinates writes
  switch (order) {
    case 0: hpc_cut_kernel<0, TI><<<g, 256, 0, s>>>(ws, nx, ny, wl, large, lny, lnx, (double)x0, (double)y0, cval, out); break;
    case 1: hpc_cut_kernel<1, TI><<<g, 256, 0, s>>>(ws, nx, ny, wl, large, lny, lnx, (double)x0, (double)y0, cval, out); break;
    case 2: hpc_cut_kernel<2, TI><<<g, 256, 0, s>>>(ws, nx, ny, wl, large, lny, lnx, (double)x0, (double)y0, cval, out); break;
    case 3: hpc_cut_kernel<3, TI><<<g, 256, 0, s>>>(ws, nx, ny, wl, large, lny, lnx, (double)x0, (double)y0, cval, out); break;
    default: return fail(COREG_EINVAL, "spline order must be 0..3");
  }
  CK_LAUNCH("hpc_cut_kernel");
  return COREG_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Solar-surface reprojection (sunpy `reproject_to` under `propagate_with_solar_surface`, alignment.py:939-985), the
// one-time call. See include/coreg_b200.h and oracle/surface_reproject.py for the algorithm restated here.
// ---------------------------------------------------------------------------------------------------------
template <typename TI>
__global__ void pad_edge_kernel(const TI* __restrict__ img, int ny, int nx, double* __restrict__ out) {
  const int pnx = nx + 2;
  const int64_t n = (int64_t)(ny + 2) * pnx;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = min(max((int)(idx % pnx) - 1, 0), nx - 1), j = min(max((int)(idx / pnx) - 1, 0), ny - 1);
    out[idx] = (double)img[(int64_t)j * nx + i];
  }
}

struct SurfaceDev {
  double ex[3], ey[3], ez[3];      // heliocentric-cartesian axes of the grid's observer in Stonyhurst coordinates
  double fx[3], fy[3], fz[3];      // ... of the image's observer
  double d_grid, d_image, rsun, dt_days;
  int identical;                   // same observer, same time: the change of frame is the identity
};

__global__ void surface_cut_kernel(TanDev ws, int nx, int ny, TanDev wl, const double* __restrict__ large_pad, int lny,
                                   int lnx, SurfaceDev f, double cval, double* __restrict__ out) {
  const int64_t n = (int64_t)nx * ny;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    double lo, la;
    tan_pix2world_dev(ws, (int)(idx % nx), (int)(idx / nx), 0, lo, la);
    bool ok = true;
    if (!f.identical) {
      double stx, ctx, sty, cty;
      sincos(lo * kD2R, &stx, &ctx);
      sincos(la * kD2R, &sty, &cty);
      // make_3d: near intersection of the line of sight with the sphere of radius rsun (off-disc: NaN)
      const double D = f.d_grid, cosa = cty * ctx;
      const double d = D * cosa - sqrt(D * D * cosa * cosa - D * D + f.rsun * f.rsun);
      const double x = d * cty * stx, y = d * sty, z = D - d * cty * ctx;
      const double px = x * f.ex[0] + y * f.ey[0] + z * f.ez[0];
      const double py = x * f.ex[1] + y * f.ey[1] + z * f.ez[1];
      const double pz = x * f.ex[2] + y * f.ey[2] + z * f.ez[2];
      const double r = sqrt(px * px + py * py + pz * pz);
      const double lat = asin(pz / r);
      // differential rotation, Howard et al. (urad / s), in the synodic frame of the Stonyhurst longitude
      const double s2 = sin(lat) * sin(lat);
      const double rate = (2.894 + -0.428 * s2 + -0.370 * s2 * s2) * 1e-6 * 86400.0;   // rad / day
      const double lon = atan2(py, px) + (rate * f.dt_days * kR2D - 0.9856 * f.dt_days) * kD2R;
      double sl, cl;
      sincos(lon, &sl, &cl);
      const double cb = cos(lat);
      const double qx = r * cb * cl, qy = r * cb * sl, qz = r * sin(lat);
      const double x2 = qx * f.fx[0] + qy * f.fx[1] + qz * f.fx[2];
      const double y2 = qx * f.fy[0] + qy * f.fy[1] + qz * f.fy[2];
      const double z2 = qx * f.fz[0] + qy * f.fz[1] + qz * f.fz[2];
      const double D2 = f.d_image;
      const double dist = sqrt(x2 * x2 + y2 * y2 + (D2 - z2) * (D2 - z2));
      lo = atan2(x2, D2 - z2) * kR2D;
      la = asin(y2 / dist) * kR2D;
      ok = z2 * D2 > r * r;          // the surface point faces the image's observer (NaN compares false)
    }
    double x, y, v = cval;
    tan_world2pix_dev(wl, lo, la, x, y);
    ok = ok && (x >= -0.5) && (x <= (double)lnx - 0.5) && (y >= -0.5) && (y <= (double)lny - 0.5);
    if (ok) {
      if (!spline_sample<1, true, double>(large_pad, lny + 2, lnx + 2, y + 1.0, x + 1.0, v)) v = cval;
    }
    out[idx] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// Image statistics in one pass over the image (multi-block, deterministic): mean of the finite values (the pivot of
// the single-pass Pearson moments), their count and max |v|, optionally widening a float32 image to float64 on the
// way (the host-side `np.array(..., dtype=float64)` of hdrshift/alignment.py:299-316). A second pass, for the mixed-
// arithmetic kernel only, writes the float32 image centred on the float32-rounded pivot and takes the RMS of the
// centred values. Every block leaves one partial; the block that finishes last folds the partials in index order,
// so the result does not depend on scheduling. The grid is fixed (kStatBlocks x kStatThreads) on every device.
// ---------------------------------------------------------------------------------------------------------
constexpr int kStatBlocks = 592, kStatThreads = 256;
struct StatPart { double sum, cnt, maxabs, sumsq; };
struct StatScratch { StatPart part[kStatBlocks]; unsigned done; unsigned pad[3]; };

__device__ __forceinline__ void stat_block_fold(StatPart& v, StatPart* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int o = kStatThreads / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      sh[threadIdx.x].sum += sh[threadIdx.x + o].sum;
      sh[threadIdx.x].cnt += sh[threadIdx.x + o].cnt;
      sh[threadIdx.x].sumsq += sh[threadIdx.x + o].sumsq;
      sh[threadIdx.x].maxabs = fmax(sh[threadIdx.x].maxabs, sh[threadIdx.x + o].maxabs);
    }
    __syncthreads();
  }
  v = sh[0];
  __syncthreads();
}

// CENTER = false: sum / count / max|v| (+ optional widening);  CENTER = true (T = float): out32 = v - (float)mean, RMS
template <typename T, bool CENTER>
__global__ void __launch_bounds__(kStatThreads)
image_stats_kernel(const T* __restrict__ img, int64_t n, double* __restrict__ widen, float* __restrict__ out32,
                   StatScratch* __restrict__ scr, double* __restrict__ stats, int stride) {
  __shared__ StatPart sh[kStatThreads];
  __shared__ bool last;
  StatPart v{0.0, 0.0, 0.0, 0.0};
  const float p32 = CENTER ? (float)stats[0] : 0.f;
  for (int64_t i = (int64_t)blockIdx.x * kStatThreads + threadIdx.x; i < n; i += (int64_t)kStatBlocks * kStatThreads) {
    const T raw = img[i];
    if (CENTER) {
      const float c = (float)raw - p32;
      out32[i] = c;
      if (isfinite(c)) {
        v.sumsq = fma((double)c, (double)c, v.sumsq);
        v.cnt += 1.0;
      }
    } else {
      const double x = (double)raw;
      if (widen) widen[i] = x;
      if (isfinite(x)) {
        v.sum += x;
        v.cnt += 1.0;
        v.maxabs = fmax(v.maxabs, fabs(x));
      }
    }
  }
  stat_block_fold(v, sh);
  if (threadIdx.x == 0) {
    scr->part[blockIdx.x] = v;
    __threadfence();
    last = atomicAdd(&scr->done, 1u) == (unsigned)(kStatBlocks - 1);
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  StatPart t{0.0, 0.0, 0.0, 0.0};
  for (int b = threadIdx.x; b < kStatBlocks; b += kStatThreads) {
    const volatile StatPart* q = &scr->part[b];
    t.sum += q->sum;
    t.cnt += q->cnt;
    t.sumsq += q->sumsq;
    t.maxabs = fmax(t.maxabs, q->maxabs);
  }
  stat_block_fold(t, sh);
  if (threadIdx.x == 0) {
    if (CENTER) {
      stats[3 * stride] = (t.cnt > 0.0) ? sqrt(t.sumsq / t.cnt) : 0.0;
    } else {
      stats[0] = (t.cnt > 0.0) ? t.sum / t.cnt : 0.0;
      stats[stride] = t.cnt;
      stats[2 * stride] = t.maxabs;
      stats[3 * stride] = 0.0;
    }
    scr->done = 0;   // ready for the next call on this scratch
  }
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

// big-endian 32-bit words (a FITS BITPIX -32 / 32 payload as stored) -> native, in place or out of place
__global__ void bswap32_kernel(const unsigned* __restrict__ in, int64_t n, unsigned* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __byte_perm(in[i], 0u, 0x0123);
}

__global__ void f32_to_f64_kernel(const float* __restrict__ in, int64_t n, double* __restrict__ out) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = (double)in[i];
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

// SPICE L2 cube [n_lambda][ny][nx] (float32 as stored in the FITS file: big-endian) -> 2-D image: `np.nansum` over the
// selected wavelength planes in float64, planes added in ascending order (numpy reduces the leading axis of a C-ordered
// array plane by plane: the same additions in the same order, NaN planes skipped, a pixel without any finite plane gives
// 0.0), rows outside [ymin, ymax) set to NaN (`hdrshift/alignment_spice.py:250-276`).
__global__ void wave_nansum_kernel(const unsigned* __restrict__ cube, int big_endian, int n_lambda, int ny, int nx,
                                   const unsigned char* __restrict__ sel, int ymin, int ymax, double* __restrict__ out) {
  const int64_t n = (int64_t)ny * nx;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (int64_t)gridDim.x * blockDim.x) {
    const int row = (int)(idx / nx);
    double s = 0.0;
    for (int k = 0; k < n_lambda; ++k) {
      if (!sel[k]) continue;
      unsigned u = cube[(int64_t)k * n + idx];
      if (big_endian) u = __byte_perm(u, 0, 0x0123);
      const float v = __uint_as_float(u);
      if (v == v) s += (double)v;
    }
    out[idx] = (row < ymin || row >= ymax) ? CUDART_NAN : s;
  }
}

template <int ORDER, typename T>
__global__ void synras_kernel(const T* __restrict__ frames, int fnx, int fny, const TanDev* __restrict__ wcs,
                              const double* __restrict__ origin, const int* __restrict__ frame_of_col,
                              const double* __restrict__ lng,
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
      // frames may be windows [y0:, x0:] of the images their WCS describes: coordinates are computed in the full image
      // and the integer origin comes off exactly -- same taps, same weights, same bits as with the whole frame
      x -= origin[2 * f];
      y -= origin[2 * f + 1];
      // interpol2d(dst=None) returns the imager's dtype (utils/Util.py:95-97): float32 frames give float32-rounded
      // samples, which the reference then stores into its float64 raster
      if (spline_sample<ORDER, true, T>(frames + (size_t)f * fnx * fny, fny, fnx, y, x, s))
        v = (sizeof(T) == 4) ? (double)__double2float_rn(s) : s;
    }
    out[idx] = v;
  }
}
}  // namespace coreg

using namespace coreg;

extern "C" {

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

int coreg_hpc_cut(const CoregTanWcs* wcs_small, int nx, int ny, const CoregTanWcs* wcs_large, const void* large,
                  int large_dtype, int large_ny, int large_nx, int origin_x, int origin_y, int order, float* ref,
                  void* stream) {
  TanDev ts, tl;
  int rc = make_tan(wcs_small, &ts);
  if (rc) return rc;
  rc = make_tan(wcs_large, &tl);
  if (rc) return rc;
  if (nx <= 0 || ny <= 0) return COREG_OK;
  if (!large || !ref) return fail(COREG_EINVAL, "coreg_hpc_cut: null pointer");
  if (large_ny <= 0 || large_nx <= 0) return fail(COREG_EINVAL, "empty image");
  cudaStream_t s = (cudaStream_t)stream;
  if (large_dtype == COREG_F32)
    return launch_hpc_cut(ts, nx, ny, tl, (const float*)large, large_ny, large_nx, origin_x, origin_y, order, ref, s);
  if (large_dtype == COREG_F64)
    return launch_hpc_cut(ts, nx, ny, tl, (const double*)large, large_ny, large_nx, origin_x, origin_y, order, ref, s);
  return fail(COREG_EINVAL, "dtype must be COREG_F32 or COREG_F64");
}

int coreg_pad_edge(const void* img, int dtype, int ny, int nx, double* out, void* stream) {
  if (ny <= 0 || nx <= 0) return fail(COREG_EINVAL, "empty image");
  if (!img || !out) return fail(COREG_EINVAL, "coreg_pad_edge: null pointer");
  const int64_t n = (int64_t)(ny + 2) * (nx + 2);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == COREG_F32)
    pad_edge_kernel<float><<<grid_for(n), 256, 0, s>>>((const float*)img, ny, nx, out);
  else if (dtype == COREG_F64)
    pad_edge_kernel<double><<<grid_for(n), 256, 0, s>>>((const double*)img, ny, nx, out);
  else
    return fail(COREG_EINVAL, "dtype must be COREG_F32 or COREG_F64");
  CK_LAUNCH("pad_edge_kernel");
  return COREG_OK;
}

static void surface_axes(double lon, double lat, double* ex, double* ey, double* ez) {
  ez[0] = cos(lat) * cos(lon); ez[1] = cos(lat) * sin(lon); ez[2] = sin(lat);
  ex[0] = -sin(lon); ex[1] = cos(lon); ex[2] = 0.0;
  ey[0] = ez[1] * ex[2] - ez[2] * ex[1];
  ey[1] = ez[2] * ex[0] - ez[0] * ex[2];
  ey[2] = ez[0] * ex[1] - ez[1] * ex[0];
}

int coreg_surface_cut(const CoregTanWcs* wcs_small, int nx, int ny, const CoregTanWcs* wcs_large,
                      const double* large_pad, int large_ny, int large_nx, const CoregSurfaceFrames* fr, double* ref,
                      void* stream) {
  TanDev ts, tl;
  int rc = make_tan(wcs_small, &ts);
  if (rc) return rc;
  rc = make_tan(wcs_large, &tl);
  if (rc) return rc;
  if (!fr) return fail(COREG_EINVAL, "null CoregSurfaceFrames");
  if (nx <= 0 || ny <= 0) return COREG_OK;
  if (!large_pad || !ref) return fail(COREG_EINVAL, "coreg_surface_cut: null pointer");
  if (large_ny <= 0 || large_nx <= 0) return fail(COREG_EINVAL, "empty image");
  if (!(fr->rsun > 0.0) || !(fr->grid_dsun > fr->rsun) || !(fr->image_dsun > fr->rsun))
    return fail(COREG_EINVAL, "coreg_surface_cut: observers must be outside a sphere of positive radius");
  SurfaceDev f;
  surface_axes(fr->grid_lon, fr->grid_lat, f.ex, f.ey, f.ez);
  surface_axes(fr->image_lon, fr->image_lat, f.fx, f.fy, f.fz);
  f.d_grid = fr->grid_dsun;
  f.d_image = fr->image_dsun;
  f.rsun = fr->rsun;
  f.dt_days = fr->dt_days;
  f.identical = (fr->grid_lon == fr->image_lon && fr->grid_lat == fr->image_lat && fr->grid_dsun == fr->image_dsun &&
                 fr->dt_days == 0.0) ? 1 : 0;
  const int64_t n = (int64_t)nx * ny;
  surface_cut_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(ts, nx, ny, tl, large_pad, large_ny, large_nx, f,
                                                                    (double)NAN, ref);
  CK_LAUNCH("surface_cut_kernel");
  return COREG_OK;
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

int coreg_bswap32(const void* in, int64_t n, void* out, void* stream) {
  if (n <= 0) return COREG_OK;
  if (!in || !out) return fail(COREG_EINVAL, "coreg_bswap32: null pointer");
  bswap32_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>((const unsigned*)in, n, (unsigned*)out);
  CK_LAUNCH("bswap32_kernel");
  return COREG_OK;
}

int coreg_spice_wave_sum(const void* cube, int big_endian, int n_lambda, int ny, int nx, const unsigned char* sel_host,
                         int ymin, int ymax, double* out, void* stream) {
  if (!cube || !sel_host || !out) return fail(COREG_EINVAL, "coreg_spice_wave_sum: null pointer");
  if (n_lambda <= 0 || ny <= 0 || nx <= 0) return fail(COREG_EINVAL, "coreg_spice_wave_sum: empty cube");
  cudaStream_t s = (cudaStream_t)stream;
  unsigned char* dsel = nullptr;
  CK(cudaMallocAsync(&dsel, (size_t)n_lambda, s));
  CK(cudaMemcpyAsync(dsel, sel_host, (size_t)n_lambda, cudaMemcpyHostToDevice, s));
  const int64_t n = (int64_t)ny * nx;
  wave_nansum_kernel<<<grid_for(n), 256, 0, s>>>((const unsigned*)cube, big_endian, n_lambda, ny, nx, dsel, ymin, ymax,
                                                 out);
  CK_LAUNCH("wave_nansum_kernel");
  CK(cudaFreeAsync(dsel, s));
  CK(cudaStreamSynchronize(s));   // sel_host is the caller's temporary
  return COREG_OK;
}

size_t coreg_image_stats_scratch_bytes(void) { return sizeof(StatScratch); }

int coreg_image_stats(const void* img, int dtype, int64_t n, double* widen, double* stats, int stats_stride,
                      void* scratch, void* stream) {
  if (!img || !stats || !scratch || n <= 0 || stats_stride <= 0) return fail(COREG_EINVAL, "coreg_image_stats: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  StatScratch* scr = static_cast<StatScratch*>(scratch);
  CK(cudaMemsetAsync(&scr->done, 0, sizeof(unsigned), s));
  if (dtype == COREG_F32)
    image_stats_kernel<float, false><<<kStatBlocks, kStatThreads, 0, s>>>((const float*)img, n, widen, nullptr, scr, stats,
                                                                          stats_stride);
  else if (dtype == COREG_F64 && !widen)
    image_stats_kernel<double, false><<<kStatBlocks, kStatThreads, 0, s>>>((const double*)img, n, nullptr, nullptr, scr,
                                                                           stats, stats_stride);
  else
    return fail(COREG_EINVAL, "coreg_image_stats: dtype must be COREG_F32 or COREG_F64 (widening needs COREG_F32)");
  CK_LAUNCH("image_stats_kernel");
  return COREG_OK;
}

int coreg_center_f32(const float* img, int64_t n, float* out, double* stats, int stats_stride, void* scratch,
                     void* stream) {
  if (!img || !out || !stats || !scratch || n <= 0 || stats_stride <= 0)
    return fail(COREG_EINVAL, "coreg_center_f32: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  StatScratch* scr = static_cast<StatScratch*>(scratch);
  CK(cudaMemsetAsync(&scr->done, 0, sizeof(unsigned), s));
  image_stats_kernel<float, true><<<kStatBlocks, kStatThreads, 0, s>>>(img, n, nullptr, out, scr, stats, stats_stride);
  CK_LAUNCH("image_stats_kernel<center>");
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

int coreg_synras_build(const void* frames, int frame_dtype, int n_frames, int fnx, int fny, const CoregTanWcs* wcs,
                       const int* frame_of_col, const double* lng, const double* lat, int n_rows, int n_cols,
                       int order, double* out, void* stream) {
  return coreg_synras_build_windows(frames, frame_dtype, n_frames, fnx, fny, wcs, nullptr, frame_of_col, lng, lat,
                                    n_rows, n_cols, order, out, stream);
}

int coreg_synras_build_windows(const void* frames, int frame_dtype, int n_frames, int fnx, int fny,
                               const CoregTanWcs* wcs, const int* origin_xy, const int* frame_of_col, const double* lng,
                               const double* lat, int n_rows, int n_cols, int order, double* out, void* stream) {
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
  double horg[2 * kMaxSynrasFrames];
  for (int f = 0; f < 2 * n_frames; ++f) horg[f] = origin_xy ? (double)origin_xy[f] : 0.0;
  TanDev* dw = nullptr;
  double* dorg = nullptr;
  int* dcol = nullptr;
  CK(cudaMallocAsync(&dw, sizeof(TanDev) * n_frames, s));
  CK(cudaMallocAsync(&dorg, sizeof(double) * 2 * n_frames, s));
  CK(cudaMallocAsync(&dcol, sizeof(int) * n_cols, s));
  CK(cudaMemcpyAsync(dw, hw, sizeof(TanDev) * n_frames, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dorg, horg, sizeof(double) * 2 * n_frames, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dcol, frame_of_col, sizeof(int) * n_cols, cudaMemcpyHostToDevice, s));
  // hw / frame_of_col are pageable: the async copies above have consumed them when the call returns
  const int64_t n = (int64_t)n_rows * n_cols;
  const int g = grid_for(n);
#define SYN(ORD, T) \
  synras_kernel<ORD, T><<<g, 256, 0, s>>>((const T*)frames, fnx, fny, dw, dorg, dcol, lng, lat, n_rows, n_cols, out)
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
  CK(cudaFreeAsync(dorg, s));
  CK(cudaFreeAsync(dcol, s));
  return COREG_OK;
}

}  // extern "C"
