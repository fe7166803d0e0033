// Stand-alone timing of k_eig's building blocks (eig_factor / eig_solve / eig_apply_red / eig_dot_red) on synthetic
// channel data: cycles per call for 1 CTA alone and for a full wave (4 CTAs per SM), red in global vs shared memory.
#define AKS_EIG_TIMING
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "../../atomickohnsham_b200/csrc/aks_device.cuh"
#include "../../atomickohnsham_b200/csrc/aks_types.cuh"
#include "../../atomickohnsham_b200/csrc/k_eig.cuh"

__global__ void __launch_bounds__(EIG_NT, 4) bench(const double* red_g, int C, int nb, int reps, int red_in_smem, int pf, long long* out, double* chk) {
  extern __shared__ __align__(16) double sm[];
  EigShared S;
  double* ptr = sm;
  S.mhh = ptr; ptr += 3 * C;
  S.fac = ptr; ptr += (size_t)4 * nb * C;
  S.x.t = ptr; ptr += (size_t)nb * C; S.x.h = ptr; ptr += C;
  S.r.t = ptr; ptr += (size_t)nb * C; S.r.h = ptr; ptr += C;
  S.y.t = ptr; ptr += (size_t)nb * C; S.y.h = ptr; ptr += C;
  S.hatpart = ptr; ptr += 2 * C;
  S.sc = ptr; ptr += EIG_NG * 3 * C;
  S.tri_di = ptr; ptr += C;
  S.tri_l = ptr; ptr += C;
  S.xh = ptr; ptr += C;
  S.red8 = ptr; ptr += AKS_NW + 8;
  S.negc = (int*)ptr;
  S.cntg = S.negc + EIG_NG * (C + 2);
  ptr += (EIG_NG * (C + 2) + EIG_NG + 2 + 1) / 2 + 1;
  const double* red = red_g + (size_t)blockIdx.x % 64 * (size_t)red_stride(nb) * C;
  if (red_in_smem) {
    for (int i = threadIdx.x; i < red_stride(nb) * C; i += blockDim.x) ptr[i] = red[i];
    S.red = ptr;
  } else S.red = red;
  for (int c = threadIdx.x; c < C; c += blockDim.x) { S.mhh[c] = 0.7; S.mhh[C + c] = 0.16; S.mhh[2 * C + c] = 0.7; }
  for (int i = threadIdx.x; i < nb * C; i += blockDim.x) { S.x.t[i] = 1.0 / (1 + i % 13); S.r.t[i] = 0.5 / (1 + i % 7); }
  for (int i = threadIdx.x; i < C; i += blockDim.x) { S.x.h[i] = 0.3; S.r.h[i] = 0.2; }
  __syncthreads();
  long long ft[4] = {0, 0, 0, 0};
  long long t_f1 = 0, t_f3 = 0, t_sol = 0, t_app = 0, t_dot = 0;
  long long c1[2] = {0, 0}, c3[2] = {0, 0};
  for (int r = 0; r < reps; ++r) {
    const double sigma = 52.3 + 37.0 * r;      // inside the spectrum: negative pivots in cells and hats
    long long t0 = clock64();
    ft[0] = ft[1] = 0;
    bool lay = false;
    if (pf) lay = eig_factor_pf(S, C, nb, 1, sigma, 1, ft); else eig_factor(S, C, nb, 1, sigma, 1, ft);
    t_f1 += clock64() - t0; c1[0] += ft[0]; c1[1] += ft[1];
    t0 = clock64();
    if (pf && lay) eig_solve_pf(S, C, nb); else eig_solve(S, C, nb);
    t_sol += clock64() - t0;
    if (blockIdx.x == 0 && chk) {   // results for the old-vs-new comparison: solve output and the inertia count
      for (int i = threadIdx.x; i < nb * C; i += blockDim.x) chk[(size_t)r * 1024 + i] = S.y.t[i];
      for (int i = threadIdx.x; i < C - 1; i += blockDim.x) chk[(size_t)r * 1024 + 800 + i] = S.y.h[i];
      if (threadIdx.x == 0) { chk[(size_t)r * 1024 + 900] = S.cntg[0]; chk[(size_t)r * 1024 + 901] = lay; }
      __syncthreads();
    }
    t0 = clock64();
    eig_apply_red(S, C, nb, 0, S.y, S.r);
    t_app += clock64() - t0;
    t0 = clock64();
    const double d = eig_dot_red(S, C, nb, S.y, S.r);
    t_dot += clock64() - t0;
    if (d == 123.456) S.x.t[0] = d;
    const int g = threadIdx.x / C;
    t0 = clock64();
    ft[0] = ft[1] = 0;
    if (pf) eig_factor_pf(S, C, nb, 3, sigma - 11.0 * g, 0, ft); else eig_factor(S, C, nb, 3, sigma - 11.0 * g, 0, ft);
    t_f3 += clock64() - t0; c3[0] += ft[0]; c3[1] += ft[1];
    if (blockIdx.x == 0 && chk && threadIdx.x < 3) chk[(size_t)r * 1024 + 902 + threadIdx.x] = S.cntg[threadIdx.x];
  }
  if (threadIdx.x == 0) {
    long long* o = out + (size_t)blockIdx.x * 12;
    o[0] = t_f1 / reps; o[1] = c1[0] / reps; o[2] = c1[1] / reps; o[3] = t_f3 / reps; o[4] = c3[0] / reps; o[5] = c3[1] / reps;
    o[6] = t_sol / reps; o[7] = t_app / reps; o[8] = t_dot / reps; o[9] = S.cntg[0];
  }
}

int main() {
  const int C = 40, nb = 19, stride = red_stride(nb);
  std::vector<double> h((size_t)64 * stride * C);
  for (int ch = 0; ch < 64; ++ch)
    for (int c = 0; c < C; ++c) {
      double* R = h.data() + (size_t)ch * stride * C + c;
      for (int j = 0; j < nb; ++j) {
        R[(size_t)j * C] = 5.0 + 40.0 * j * j + 0.1 * c + 0.01 * ch;        // td
        R[(size_t)(nb + j) * C] = 3.0 + 2.0 * j;                              // te
        R[(size_t)(2 * nb + j) * C] = 0.3 / (1 + j); R[(size_t)(3 * nb + j) * C] = -0.2 / (1 + j);
        R[(size_t)(4 * nb + j) * C] = 0.05 / (1 + j); R[(size_t)(5 * nb + j) * C] = 0.04 / (1 + j);
      }
      R[(size_t)(6 * nb) * C] = 30.0; R[(size_t)(6 * nb + 1) * C] = -7.0; R[(size_t)(6 * nb + 2) * C] = 30.0; R[(size_t)(6 * nb + 3) * C] = 0;
    }
  double* dred; cudaMalloc(&dred, h.size() * 8); cudaMemcpy(dred, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
  long long* dout; cudaMalloc(&dout, 8 * 12 * 2048);
  double* dchk; cudaMalloc(&dchk, 8 * 1024 * 20 * 2);
  std::vector<double> hchk[2];
  const size_t base = eig_smem_bytes(C, nb, 0) + 64;
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int pf = 0; pf < 2; ++pf)
  for (int smem_red = 0; smem_red < 2; ++smem_red)
    for (int grid : {1, 592}) {
      const size_t smem = base + (smem_red ? sizeof(double) * stride * C : 0);
      cudaMemset(dchk, 0, 8 * 1024 * 20);
      bench<<<grid, EIG_NT, smem>>>(dred, C, nb, 20, smem_red, pf, dout, dchk);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> o((size_t)grid * 12);
      cudaMemcpy(o.data(), dout, o.size() * 8, cudaMemcpyDeviceToHost);
      double a[10] = {0};
      for (int b = 0; b < grid; ++b) for (int k = 0; k < 10; ++k) a[k] += (double)o[(size_t)b * 12 + k] / grid;
      if (grid == 1 && smem_red == 0) { hchk[pf].resize(1024 * 20); cudaMemcpy(hchk[pf].data(), dchk, 8 * 1024 * 20, cudaMemcpyDeviceToHost); }
      printf("%s red %s grid %4d: factor(store) %6.0f [cells %6.0f hats %6.0f]  factor(3 shifts) %6.0f [cells %6.0f hats %6.0f]  solve %6.0f  apply %6.0f  dot %5.0f   (count %g)\n",
             pf ? "product" : "ratio  ", smem_red ? "smem  " : "global", grid, a[0], a[1], a[2], a[3], a[4], a[5], a[6], a[7], a[8], a[9]);
    }
  // old vs new: solve outputs and inertia counts of 20 shifts
  double worst = 0.0; int cntdiff = 0, fallback = 0;
  for (int r = 0; r < 20; ++r) {
    double nrm = 0.0, dif = 0.0;
    for (int i = 0; i < 840; ++i) { const double a = hchk[0][r * 1024 + i], b = hchk[1][r * 1024 + i]; nrm = fmax(nrm, fabs(a)); dif = fmax(dif, fabs(a - b)); }
    worst = fmax(worst, dif / nrm);
    for (int k = 900; k < 905; ++k) if (k != 901 && hchk[0][r * 1024 + k] != hchk[1][r * 1024 + k]) ++cntdiff;
    fallback += hchk[1][r * 1024 + 901] == 0.0;
    if (r < 4) printf("shift %d: counts ratio-form %g / %g %g %g   product-form %g / %g %g %g\n", r, hchk[0][r*1024+900], hchk[0][r*1024+902], hchk[0][r*1024+903], hchk[0][r*1024+904],
                      hchk[1][r*1024+900], hchk[1][r*1024+902], hchk[1][r*1024+903], hchk[1][r*1024+904]);
  }
  printf("old vs new: worst relative difference of the solve output %.2e, differing inertia counts %d of 80, fallbacks %d\n", worst, cntdiff, fallback);
  return 0;
}
