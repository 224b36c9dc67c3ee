// coreg_lag_offset.cu -- Carrington-frame lag kernel (order-2 spline, FMA arithmetic): per-pixel detector-plane offsets + per-lag shift.
#include "coreg_common.cuh"

namespace coreg {
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

int launch_offset_fast(int variant, int gnx, int gny, int64_t n_lags, int sms, cudaStream_t s, const double* ref,
                       const void* small, int small_dtype, int snx, int sny, const double* tx, const double* ty,
                       const CoregLagOffset* lags, const double* pivots, void* work, int* tiles_out) {
  // the per-lag fast table lives in the tail of the workspace (after the [tiles][lags][8] partials)
  OffsetFastLag* ft = reinterpret_cast<OffsetFastLag*>(static_cast<char*>(work) + partials_bytes(gnx, gny, n_lags));
  offset_fast_table_kernel<<<((int)n_lags + 127) / 128, 128, 0, s>>>(lags, (int)n_lags, ft);
  // per-slice offset ranges right behind the table (the tail reserves 96 B per lag; the table uses 16)
  double* ranges = reinterpret_cast<double*>(ft + n_lags);
  double* w = static_cast<double*>(work);
  OffsetCoord::Planes planes{tx, ty};
  int rc;
  if (small_dtype == COREG_F32)
    rc = launch_lag_fast<OffsetFast, float, double, false>(variant, gnx, gny, n_lags, sms, s, ref, (const float*)small, snx,
                                                           sny, planes, ft, pivots, w, tiles_out, ranges);
  else
    rc = launch_lag_fast<OffsetFast, double, double, false>(variant, gnx, gny, n_lags, sms, s, ref, (const double*)small,
                                                            snx, sny, planes, ft, pivots, w, tiles_out, ranges);
  if (rc) return rc;
  CK_LAUNCH("lag_corr_fast_kernel");
  return COREG_OK;
}
}  // namespace coreg
