// Microbenchmark (development tool): the operator products of k_element_tma in isolation.  Every warp runs
// NodeSlice::run_s3 (264 DFMA fed by 132 LDCU.128 from the kernel-parameter bank and 24 LDS.64) or run_s2 (363 DFMA + the
// flux rebuild) in a loop on a resident shared-memory tile: the FP64 rate that code shape can reach on one SM, as a
// function of the number of resident warps -- the ceiling of the element kernel's S2 / S3 phases.
#include <cstdio>
#include <cuda_runtime.h>
#include "../pdesolver.jl_b200/csrc/element_tma.cuh"
using namespace pdes;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

using Cfg = ElemTmaCfg<3, 11, 6, false>;
template <int MODE>
__global__ void __launch_bounds__(1024, 1) k(const __grid_constant__ OpTabP<3, 11, 6> op, double* out, int iters) {
  extern __shared__ __align__(16) double sm[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* w = sm + warp * (Cfg::QW + Cfg::RW + Cfg::XW + Cfg::UW);
  for (int i = lane; i < Cfg::QW + Cfg::RW + Cfg::XW + Cfg::UW; i += 32) w[i] = 1.0 + 1e-3 * (i % 17);
  __syncwarp();
  const double *sQ = w, *sR = w + Cfg::QW, *sX = sR + Cfg::RW, *sU = sX + Cfg::XW;
  const int s = lane / 5, k_ = lane - s * 5, sc = s < 6 ? s : 0;
  double acc[12];
#pragma unroll
  for (int u = 0; u < 12; ++u) acc[u] = 0.0;
  double acc1[12];
#pragma unroll
  for (int u = 0; u < 12; ++u) acc1[u] = 0.0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) NodeSlice<3, 11, 6, 0, 12>::run_s3(op, sR, sc, k_, acc);
    else if (MODE == 1) NodeSlice<3, 11, 6, 0, 12>::template run_s2<false>(op, sQ, sX, sU, sc, k_, acc);
    else {
      // two rows per lane share every coefficient pair: 6 LDCU.128 feed 22 DFMA
      const double* g0 = sR + sc * 120 + k_;
      const double* g1 = sR + ((sc + 3) % 6) * 120 + k_;
#pragma unroll
      for (int f = 0; f < 4; ++f) {
        double a0[6], a1[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { a0[i] = g0[(f * 6 + i) * 5]; a1[i] = g1[(f * 6 + i) * 5]; }
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          const double2* crow = reinterpret_cast<const double2*>(&op.RfN[f * 6 + i][0]);
#pragma unroll
          for (int h = 0; h < 6; ++h) {
            const double2 c = crow[h];
            acc[2 * h] = fma(c.x, a0[i], acc[2 * h]);
            acc1[2 * h] = fma(c.x, a1[i], acc1[2 * h]);
            if (2 * h + 1 < 11) { acc[2 * h + 1] = fma(c.y, a0[i], acc[2 * h + 1]); acc1[2 * h + 1] = fma(c.y, a1[i], acc1[2 * h + 1]); }
          }
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 12; ++u) acc[u] += acc1[u];
  double t = 0;
#pragma unroll
  for (int u = 0; u < 11; ++u) t += acc[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE>
int run(const char* name, double* out, const OpTabP<3, 11, 6>& t, int warps) {
  int iters = 4000;
  size_t smem = (size_t)warps * (Cfg::QW + Cfg::RW + Cfg::XW + Cfg::UW) * 8;
  CK(cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, warps * 32, smem>>>(t, out, 10);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k<MODE><<<148, warps * 32, smem>>>(t, out, iters);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double dfma = MODE == 0 ? 264.0 : (MODE == 2 ? 528.0 : 363.0 + 33.0 * 2 + 11);      // per lane and iteration (S2: + rebuild)
  double flops = 148.0 * warps * 32 * iters * dfma * 2.0;
  printf("%-10s warps/SM=%2d  %8.3f ms  %7.2f TF/s (all 32 lanes counted)  = %5.1f %% of 36.7\n", name, warps, ms, flops / ms * 1e-9,
         100 * flops / ms * 1e-9 / 36.7);
  return 0;
}

int main() {
  double* out; CK(cudaMalloc(&out, 148 * 1024 * sizeof(double)));
  OpTabP<3, 11, 6> t;
  for (int r = 0; r < 33; ++r) for (int i = 0; i < 12; ++i) t.Qt[r][i] = 1e-3 * (r + i);
  for (int r = 0; r < 24; ++r) for (int i = 0; i < 12; ++i) t.RfN[r][i] = 1e-3 * (r - i);
  for (int w : {4, 6, 8, 11, 14}) { run<0>("run_s3", out, t, w); run<2>("s3 x2rows", out, t, w); }
  return 0;
}
