// coreg_rice.cu -- RICE_1 tile decoder for tile-compressed FITS image HDUs.
#include "coreg_common.cuh"

namespace coreg {
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
}  // namespace coreg

using namespace coreg;

extern "C" {

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

}  // extern "C"
