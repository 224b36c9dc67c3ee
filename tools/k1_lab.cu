// k1_lab.cu -- standalone timing / agreement harness for the fused helioprojective lag kernels (GPU box only).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -o tools/_tmp/k1_lab tools/k1_lab.cu
//   tools/_tmp/k1_lab [n=2048] [lags_per_axis=60] [variants=0,1,2,...] [crota_lag_deg=0]
// Builds a config-1-like problem (n x n grid, 0.492"/px, CROTA 3 deg, lags_per_axis^2 CRVAL lags at 1"), runs the
// generic kernel once as the yardstick and every requested variant of the column-rolling kernel, and prints the
// lag-kernel time (CUDA events around the launch, coreg_profile_*) and the largest |r - r_generic|.
#include "../euispice_coreg_b200/csrc/coreg_kernels.cu"

#include <cstdlib>
#include <vector>

#define LAB_CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA %s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)
#define LAB_RC(x) do { int r_ = (x); if (r_) { printf("%s -> %d: %s\n", #x, r_, coreg_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 2048;
  const int la = argc > 2 ? atoi(argv[2]) : 60;
  const char* vars = argc > 3 ? argv[3] : "0,1,2,3,4,5,6,7";
  const double dcrota = argc > 4 ? atof(argv[4]) : 0.0;
  const int64_t npix = (int64_t)n * n, n_lags = (int64_t)la * la;
  // images: smooth positive field + small-scale structure (deterministic)
  std::vector<double> h_small(npix);
  std::vector<float> h_ref(npix);
  uint64_t st = 88172645463325252ull;
  auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return (double)(st >> 11) / 9007199254740992.0; };
  for (int64_t i = 0; i < npix; ++i) {
    const int x = (int)(i % n), y = (int)(i / n);
    const double v = 500.0 + 200.0 * sin(0.013 * x) * cos(0.017 * y) + 120.0 * sin(0.21 * x + 0.13 * y) + 60.0 * rnd();
    h_small[i] = (double)(float)v;
    h_ref[i] = (float)(v * 0.8 + 30.0 * rnd());
  }
  const double as = 1.0 / 3600.0, rho = 3.0 * kD2R;
  CoregTanWcs g;
  g.crpix1 = g.crpix2 = (n + 1) / 2.0;
  g.cdelt1 = g.cdelt2 = 0.492 * as;
  g.pc11 = cos(rho); g.pc12 = -sin(rho); g.pc21 = sin(rho); g.pc22 = cos(rho);
  g.crval1 = -124.0 * as; g.crval2 = 44.0 * as; g.lonpole = 180.0;
  std::vector<CoregTanWcs> h_lag(n_lags, g);
  for (int a = 0; a < la; ++a)
    for (int b = 0; b < la; ++b) {
      CoregTanWcs& w = h_lag[(size_t)a * la + b];
      w.crval1 = (-124.0 + (a - la / 2)) * as;
      w.crval2 = (44.0 + (b - la / 2)) * as;
      if (dcrota != 0.0) {
        const double r2 = (3.0 + dcrota) * kD2R;
        w.pc11 = cos(r2); w.pc12 = -sin(r2); w.pc21 = sin(r2); w.pc22 = cos(r2);
      }
    }
  double *d_small, *d_piv, *d_corr, *d_lng, *d_lat, *d_planes;
  float* d_ref;
  CoregTanWcs* d_lagw;
  CoregLagTan* d_lagt;
  int64_t* d_nv;
  void* d_work;
  const size_t wb = coreg_lag_corr_workspace_bytes(n, n, n_lags);
  LAB_CK(cudaMalloc(&d_small, npix * 8)); LAB_CK(cudaMalloc(&d_ref, npix * 4)); LAB_CK(cudaMalloc(&d_piv, 16));
  LAB_CK(cudaMalloc(&d_corr, n_lags * 8)); LAB_CK(cudaMalloc(&d_nv, n_lags * 8)); LAB_CK(cudaMalloc(&d_work, wb));
  LAB_CK(cudaMalloc(&d_lagw, n_lags * sizeof(CoregTanWcs))); LAB_CK(cudaMalloc(&d_lagt, n_lags * sizeof(CoregLagTan)));
  LAB_CK(cudaMalloc(&d_lng, npix * 8)); LAB_CK(cudaMalloc(&d_lat, npix * 8)); LAB_CK(cudaMalloc(&d_planes, 3 * npix * 8));
  LAB_CK(cudaMemcpy(d_small, h_small.data(), npix * 8, cudaMemcpyHostToDevice));
  LAB_CK(cudaMemcpy(d_ref, h_ref.data(), npix * 4, cudaMemcpyHostToDevice));
  LAB_CK(cudaMemcpy(d_lagw, h_lag.data(), n_lags * sizeof(CoregTanWcs), cudaMemcpyHostToDevice));
  LAB_RC(coreg_finite_mean(d_ref, COREG_F32, npix, d_piv, nullptr));
  LAB_RC(coreg_finite_mean(d_small, COREG_F64, npix, d_piv + 1, nullptr));
  // yardstick: generic kernel (world-coordinate planes + per-lag trig constants)
  LAB_RC(coreg_tan_pix2world(&g, n, n, 1, d_lng, d_lat, nullptr));
  LAB_RC(coreg_tan_trig_planes(d_lng, d_lat, npix, g.crval1, d_planes, nullptr));
  tan_lag_from_wcs_kernel<<<((int)n_lags + 127) / 128, 128>>>(d_lagw, (int)n_lags, g.crval1, g.lonpole, d_lagt);
  std::vector<double> r_gen(n_lags), r(n_lags);
  std::vector<int64_t> nv_gen(n_lags), nv(n_lags);
  double ms; int nl;
  for (int rep = 0; rep < 2; ++rep) {
    coreg_profile_begin();
    LAB_RC(coreg_hpc_lag_corr(d_ref, d_small, COREG_F64, n, n, n, n, d_planes, d_lagt, n_lags, 2, d_piv, d_work, wb, d_corr, d_nv, 0, nullptr));
    LAB_RC(coreg_profile_end(&ms, &nl));
  }
  LAB_CK(cudaMemcpy(r_gen.data(), d_corr, n_lags * 8, cudaMemcpyDeviceToHost));
  LAB_CK(cudaMemcpy(nv_gen.data(), d_nv, n_lags * 8, cudaMemcpyDeviceToHost));
  printf("{\"kernel\": \"generic\", \"ms\": %.3f, \"n\": %d, \"lags\": %lld}\n", ms, n, (long long)n_lags);
  for (const char* p = vars; *p;) {
    const int v = atoi(p);
    while (*p && *p != ',') ++p;
    if (*p == ',') ++p;
    double best = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
      coreg_profile_begin();
      LAB_RC(coreg_hpc_lag_corr_wcs(d_ref, d_small, n, n, n, n, &g, d_lagw, n_lags, 2, d_piv, d_work, wb, d_corr, d_nv, COREG_FLAG_VARIANT(v), nullptr));
      LAB_RC(coreg_profile_end(&ms, &nl));
      if (rep > 0 && ms < best) best = ms;
    }
    LAB_CK(cudaMemcpy(r.data(), d_corr, n_lags * 8, cudaMemcpyDeviceToHost));
    LAB_CK(cudaMemcpy(nv.data(), d_nv, n_lags * 8, cudaMemcpyDeviceToHost));
    double dmax = 0.0; int64_t nvd = 0; int worst = -1;
    for (int64_t i = 0; i < n_lags; ++i) {
      const double d = fabs(r[i] - r_gen[i]);
      if (d > dmax || d != d) { dmax = d; worst = (int)i; }
      nvd += llabs((long long)(nv[i] - nv_gen[i]));
    }
    printf("{\"kernel\": \"roll\", \"variant\": %d, \"ms\": %.3f, \"max_abs_diff_vs_generic\": %.3e, \"worst_lag\": %d, \"nvalid_abs_diff_sum\": %lld, \"r_mid\": %.15f}\n",
           v, best, dmax, worst, (long long)nvd, r[n_lags / 2 + la / 2]);
  }
  return 0;
}
