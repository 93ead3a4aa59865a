// Microbenchmark (development tool): what feeds DFMA fastest on sm_100a?  The S3 product of the element kernel
// (acc[u] += C[r][u] * a[r], 24 r x 11 u per row) with the coefficients coming from
//   0: LDCU.128 (kernel-parameter bank), one row per lane           (the shape k_element_tma uses)
//   1: LDCU.128, two rows per lane (every coefficient pair feeds 4 DFMA)
//   2: LDS.128 broadcast from shared memory, one row per lane
//   3: LDS.128 broadcast, two rows per lane
//   4: registers (no coefficient traffic at all: the DFMA ceiling of the loop shape)
// at realistic register budgets (launch bounds = the resident warps), no spills.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
struct __align__(16) Tab { double C[24][12]; };

template <int MODE, int NW, int HSPLIT = 3>
__global__ void __launch_bounds__(32 * NW, 1) k(const __grid_constant__ Tab op, double* out, int iters) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* sc = sm;                               // coefficient table (modes 2, 3)
  double* w = sm + 24 * 12 + warp * 2 * 32 * 25;       // per warp: two row sets [32 lanes][24 (+1 pad)]
  for (int i = threadIdx.x; i < 24 * 12; i += blockDim.x) sc[i] = op.C[i / 12][i % 12];
  for (int i = lane; i < 2 * 32 * 25; i += 32) w[i] = 1.0 + 1e-3 * (i % 17);
  __syncthreads();
  const double* g0 = w + lane * 25;
  const double* g1 = w + 32 * 25 + lane * 25;
  double acc[12], acc1[12];
#pragma unroll
  for (int u = 0; u < 12; ++u) acc[u] = acc1[u] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 24; ++r) {
      const double a0 = g0[r];
      const double a1 = (MODE == 1 || MODE == 3 || MODE == 6) ? g1[r] : 0.0;
#pragma unroll
      for (int h = 0; h < 6; ++h) {
        double2 c;
        if (MODE == 0 || MODE == 1) c = reinterpret_cast<const double2*>(&op.C[r][0])[h];
        else if (MODE == 2 || MODE == 3) c = reinterpret_cast<const double2*>(sc + r * 12)[h];
        else if (MODE >= 5) c = (h < HSPLIT) ? reinterpret_cast<const double2*>(&op.C[r][0])[h] : reinterpret_cast<const double2*>(sc + r * 12)[h];
        else c = make_double2(1.0 + 1e-9 * (r + h), 1.0 - 1e-9 * (r - h));
        acc[2 * h] = fma(c.x, a0, acc[2 * h]);
        if (2 * h + 1 < 11) acc[2 * h + 1] = fma(c.y, a0, acc[2 * h + 1]);
        if (MODE == 1 || MODE == 3 || MODE == 6) {
          acc1[2 * h] = fma(c.x, a1, acc1[2 * h]);
          if (2 * h + 1 < 11) acc1[2 * h + 1] = fma(c.y, a1, acc1[2 * h + 1]);
        }
      }
    }
    asm volatile("" ::: "memory");
  }
  double t = 0;
#pragma unroll
  for (int u = 0; u < 11; ++u) t += acc[u] + acc1[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE, int NW, int HSPLIT = 3>
int run(const char* name, double* out, const Tab& t) {
  const int iters = 4000;
  size_t smem = (size_t)(24 * 12 + NW * 2 * 32 * 25) * 8;
  CK(cudaFuncSetAttribute(k<MODE, NW, HSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaFuncAttributes fa; CK(cudaFuncGetAttributes(&fa, k<MODE, NW, HSPLIT>));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE, NW, HSPLIT><<<148, NW * 32, smem>>>(t, out, 10);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k<MODE, NW, HSPLIT><<<148, NW * 32, smem>>>(t, out, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = ((MODE == 1 || MODE == 3 || MODE == 6) ? 2.0 : 1.0) * 264.0;
  const double flops = 148.0 * NW * 32 * iters * dfma * 2.0;
  printf("%-28s warps/SM=%2d regs=%3d spill=%3zu  %8.3f ms  %6.2f TF/s = %5.1f %% of 36.7\n", name, NW, fa.numRegs,
         fa.localSizeBytes, ms, flops / ms * 1e-9, 100 * flops / ms * 1e-9 / 36.7);
  return 0;
}
#define RUNALL(NW) \
  run<0, NW>("LDCU.128 1 row", out, t); run<1, NW>("LDCU.128 2 rows", out, t); run<2, NW>("LDS.128 1 row", out, t); \
  run<3, NW>("LDS.128 2 rows", out, t); run<4, NW>("registers (ceiling)", out, t); \
  run<5, NW, 3>("hybrid 3 LDCU + 3 LDS 1 row", out, t); run<5, NW, 4>("hybrid 4 LDCU + 2 LDS 1 row", out, t); \
  run<5, NW, 2>("hybrid 2 LDCU + 4 LDS 1 row", out, t); run<6, NW, 3>("hybrid 3+3 2 rows", out, t); run<6, NW, 4>("hybrid 4+2 2 rows", out, t);
int main() {
  double* out; CK(cudaMalloc(&out, 148 * 1024 * sizeof(double)));
  Tab t;
  for (int r = 0; r < 24; ++r) for (int i = 0; i < 12; ++i) t.C[r][i] = 1e-3 * (r - i);
  RUNALL(6) RUNALL(10) RUNALL(14)
  return 0;
}
