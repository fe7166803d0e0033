// Dependent-issue latencies on sm_100a that bound the sequential sweeps of k_eig (one warp, one thread active).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double pivot_rcp(double d) {
  double x;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(d));
  double err = fma(-d, x, 1.0);
  x = fma(x, err, x);
  err = fma(-d, x, 1.0);
  x = fma(x, err, x);
  return x;
}
template <int MODE>
__global__ void k(double* out, long long* cyc, int n, double a, double b) {
  __shared__ double sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = a + 1e-3 * i;
  __syncthreads();
  double x = a, y = b, z = a * 0.5;
  long long t0 = clock64();
  if (MODE == 0) for (int i = 0; i < n; ++i) x = fma(x, b, a);                       // dependent DFMA
  if (MODE == 1) for (int i = 0; i < n; ++i) x = pivot_rcp(x + a);                    // DADD + rcp + 4 DFMA
  if (MODE == 2) for (int i = 0; i < n; ++i) x = 1.0 / (x + a);                       // IEEE division
  if (MODE == 3) for (int i = 0; i < n; ++i) x = fma(x, sm[(i * 7) & 1023], a);       // smem load off-chain + DFMA
  if (MODE == 4) for (int i = 0; i < n; ++i) x = sm[((int)x + i) & 1023];             // dependent smem load (pointer chase + cvt)
  if (MODE == 5) for (int i = 0; i < n; ++i) { __syncthreads(); x = fma(x, b, a); }   // barrier + DFMA
  if (MODE == 6) for (int i = 0; i < n; ++i) { x = fma(x, b, a); y = fma(y, b, a); z = fma(z, b, a); }  // 3 chains
  if (MODE == 7) for (int i = 0; i < n; ++i) {   // LDL^T pivot step as in eig_factor
      const double te = sm[i & 1023];
      const double lj = te * y;            // y = dprev_inv
      double dj = (a - 0.1) - lj * te;
      if (fabs(dj) < 1e-300) dj = (dj < 0.0) ? -1e-300 : 1e-300;
      y = pivot_rcp(dj);
      x += y;
  }
  if (MODE == 8) for (int i = 0; i < n; ++i) {   // product-form Sturm step
      const double e2 = sm[i & 1023];
      const double pn = fma(a, x, -(e2 * y));
      y = x; x = pn * 1e-1;
  }
  if (MODE == 9) for (int i = 0; i < n; ++i) { double v = x; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o); x = v * 1e-3 + a; }  // warp_sum chain
  if (MODE == 10 || MODE == 11 || MODE == 12) {   // FP64 tensor-core tile update (mma.sync m8n8k4 -> DMMA.8x8x4)
    double c[8][2];
    for (int j = 0; j < 8; ++j) c[j][0] = c[j][1] = 0.0;
    constexpr int NCH = MODE == 10 ? 1 : (MODE == 11 ? 4 : 8);   // independent accumulator chains
    for (int i = 0; i < n; ++i)
#pragma unroll
      for (int j = 0; j < NCH; ++j)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    for (int j = 0; j < 8; ++j) x += c[j][0] + c[j][1];
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[blockIdx.x] = x + y + z; cyc[blockIdx.x] = t1 - t0; }
}
template <int MODE> void run(const char* name, int threads, int n) {
  double* out; long long* cyc;
  cudaMalloc(&out, 8 * 8); cudaMalloc(&cyc, 8 * 8);
  k<MODE><<<1, threads>>>(out, cyc, n, 0.9, 0.5);
  k<MODE><<<1, threads>>>(out, cyc, n, 0.9, 0.5);
  long long h;
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s threads %4d: %7.1f cycles per iteration\n", name, threads, (double)h / n);
  cudaFree(out); cudaFree(cyc);
}
int main() {
  const int n = 4096;
  run<0>("dependent DFMA", 32, n);
  run<6>("3 independent DFMA chains (per round)", 32, n);
  run<1>("DADD + rcp.approx + 2 Newton (pivot_rcp)", 32, n);
  run<2>("DADD + IEEE division", 32, n);
  run<3>("DFMA with smem operand (independent address)", 32, n);
  run<4>("dependent smem load (+cvt)", 32, n);
  run<5>("__syncthreads + DFMA", 128, n);
  run<5>("__syncthreads + DFMA", 512, n);
  run<7>("LDL^T pivot step (eig_factor critical path)", 32, n);
  run<8>("product-form Sturm step", 32, n);
  run<9>("warp_sum (5 shuffles of a double) chain", 32, n);
  run<10>("dependent DMMA m8n8k4 (1 chain, 1 warp)", 32, n);
  run<11>("4 DMMA chains, 1 warp (per round of 4)", 32, n);
  run<12>("8 DMMA chains, 1 warp (per round of 8)", 32, n);
  run<12>("8 DMMA chains, 4 warps (per round of 8)", 128, n);
  run<12>("8 DMMA chains, 16 warps (per round of 8)", 512, n);
  return 0;
}
