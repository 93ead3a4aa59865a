// Microbenchmark (development tool): the element kernel's operator products as FP64 tensor-core GEMMs with the operator
// held in registers.  out[(s,k), i] = sum_r A[(s,k), r] * C[r][i]: M = 32 rows (4 m-tiles of 8), K = 60 (15 steps of 4),
// N = 16 (2 n-tiles of 8).  Per warp and tile: 60 LDS.64 (A fragments from shared memory) + 120 DMMA m8n8k4; B fragments
// (30 doubles per lane) stay in registers -- nothing streams through the uniform datapath.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
constexpr int MT = 4, KS = 15, NT = 2, ROWW = 62;   // shared-memory row stride in doubles (bank spread)
__global__ void __launch_bounds__(1024, 1) k(const double* __restrict__ cg, double* out, int iters) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, gr = lane >> 2, gc = lane & 3;
  double* w = sm + warp * (32 * ROWW);
  for (int i = lane; i < 32 * ROWW; i += 32) w[i] = 1.0 + 1e-3 * (i % 17);
  __syncwarp();
  double b[KS][NT];
#pragma unroll
  for (int t = 0; t < KS; ++t)
#pragma unroll
    for (int n = 0; n < NT; ++n) b[t][n] = cg[(4 * t + gc) * 16 + 8 * n + gr];
  double c[MT][NT][2];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) c[m][n][0] = c[m][n][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int t = 0; t < KS; ++t) {
      double av[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m) av[m] = w[(8 * m + gr) * ROWW + 4 * t + gc];
#pragma unroll
      for (int m = 0; m < MT; ++m)
#pragma unroll
        for (int n = 0; n < NT; ++n)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(c[m][n][0]), "+d"(c[m][n][1]) : "d"(av[m]), "d"(b[t][n]));
    }
  }
  double s = 0;
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int n = 0; n < NT; ++n) s += c[m][n][0] + c[m][n][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  double *out, *cg; CK(cudaMalloc(&out, 148 * 1024 * sizeof(double))); CK(cudaMalloc(&cg, 60 * 16 * sizeof(double)));
  CK(cudaMemset(cg, 0, 60 * 16 * sizeof(double)));
  for (int warps : {4, 8, 11, 14, 20}) {
    int iters = 4000;
    size_t smem = (size_t)warps * 32 * ROWW * 8;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<<<148, warps * 32, smem>>>(cg, out, 10);
    CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    k<<<148, warps * 32, smem>>>(cg, out, iters);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double tiles = 148.0 * warps * iters;
    double useful = tiles * 30 * 57 * 11 * 2.0, padded = tiles * 32 * 60 * 16 * 2.0;
    printf("DMMA tile warps/SM=%2d %8.3f ms  %6.1f ns/tile/SM-warp  padded %6.2f TF/s  useful %6.2f TF/s  (DFMA+LDCU path: 627 DFMA/lane = useful flops at 41%% of 36.7 = 15 TF/s)\n",
           warps, ms, ms * 1e6 / iters, padded / ms * 1e-9, useful / ms * 1e-9);
  }
  return 0;
}
