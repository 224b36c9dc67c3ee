// TMA 2-D tile load probe: which (box, coordinate, issue pattern) combinations run on this GPU.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/_tmp/tma_probe tools/ubench/tma_probe.cu && tools/_tmp/tma_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef CUresult (*EncFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                          const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                          CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

template <int ISSUE>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int bw, int bh, int x, int y, float* out) {
  extern __shared__ __align__(128) unsigned char smem[];
  float* win = reinterpret_cast<float*>(smem);
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem + (size_t)bw * bh * 4);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  bool issue = (ISSUE == 0) ? (threadIdx.x == 0) : (threadIdx.x < 32);
  if (issue) {
    bool leader = true;
    if (ISSUE == 1) {
      unsigned pred = 0;
      asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
      leader = pred != 0;
    }
    if (leader) {
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bw * bh * 4) : "memory");
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                       smem_u32(win)),
                   "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(smem_u32(bar)), "r"(x), "r"(y)
                   : "memory");
    }
  }
  unsigned done = 0;
  while (!done) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done) : "r"(smem_u32(bar)), "r"(0u) : "memory");
  }
  for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = win[i];
}

int main() {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncFn enc = (EncFn)p;
  if (!enc) { printf("no encoder\n"); return 1; }
  const int sizes[2][2] = {{2048, 2048}, {96, 96}};
  const int boxes[4][2] = {{64, 8}, {200, 64}, {256, 48}, {128, 64}};
  for (int si = 0; si < 2; ++si) {
    const int nx = sizes[si][0], ny = sizes[si][1];
    float* h = (float*)malloc((size_t)nx * ny * 4);
    for (int i = 0; i < nx * ny; ++i) h[i] = (float)(i % 100003);
    float* d; cudaMalloc(&d, (size_t)nx * ny * 4);
    cudaMemcpy(d, h, (size_t)nx * ny * 4, cudaMemcpyHostToDevice);
    for (int bi = 0; bi < 4; ++bi) {
      const int bw = boxes[bi][0], bh = boxes[bi][1];
      CUtensorMap map;
      const cuuint64_t dims[2] = {(cuuint64_t)nx, (cuuint64_t)ny};
      const cuuint64_t strides[1] = {(cuuint64_t)nx * 4};
      const cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh};
      const cuuint32_t es[2] = {1, 1};
      CUresult rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      printf("image %dx%d box %dx%d encode rc=%d\n", nx, ny, bw, bh, (int)rc);
      if (rc != CUDA_SUCCESS) continue;
      float* out; cudaMalloc(&out, (size_t)bw * bh * 4);
      const size_t smem = (size_t)bw * bh * 4 + 16;
      const int coords[4][2] = {{16, 8}, {-8, -3}, {nx - 20, ny - 10}, {13, 5}};
      for (int issue = 0; issue < 2; ++issue)
        for (int ci = 0; ci < (issue == 1 && bi == 3 ? 4 : 3); ++ci) {
          cudaError_t e;
          if (issue == 0) {
            cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            probe<0><<<1, 256, smem>>>(map, bw, bh, coords[ci][0], coords[ci][1], out);
          } else {
            cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            probe<1><<<1, 256, smem>>>(map, bw, bh, coords[ci][0], coords[ci][1], out);
          }
          e = cudaDeviceSynchronize();
          float* ho = (float*)malloc((size_t)bw * bh * 4);
          int bad = -1;
          if (e == cudaSuccess) {
            cudaMemcpy(ho, out, (size_t)bw * bh * 4, cudaMemcpyDeviceToHost);
            bad = 0;
            for (int yy = 0; yy < bh; ++yy)
              for (int xx = 0; xx < bw; ++xx) {
                const int gx = coords[ci][0] + xx, gy = coords[ci][1] + yy;
                const float want = (gx >= 0 && gx < nx && gy >= 0 && gy < ny) ? h[(size_t)gy * nx + gx] : 0.f;
                bad += ho[yy * bw + xx] != want;
              }
          }
          printf("  issue=%s coord=(%d,%d): %s mismatches=%d\n", issue ? "elect" : "tid0", coords[ci][0], coords[ci][1],
                 cudaGetErrorString(e), bad);
          free(ho);
          if (e != cudaSuccess) { printf("  (context lost, stopping)\n"); return 2; }
        }
      cudaFree(out);
    }
    cudaFree(d);
    free(h);
  }
  return 0;
}
