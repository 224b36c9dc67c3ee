// coreg_pixel_shift.cu -- integer pixel-shift lag search (pxlshift.AlignmentPixels).
#include "coreg_common.cuh"

namespace coreg {
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
}  // namespace coreg

using namespace coreg;

extern "C" {

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

}  // extern "C"
