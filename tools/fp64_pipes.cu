// Microbenchmark (development tool, not product): FP64 FMA pipe vs FP64 tensor
// (DMMA, mma.sync m8n8k4 / m16n8k8 .f64) throughput on sm_100a, alone and mixed,
// to decide whether the SBP operator application should use DMMA (north_star:
// "only if ncu shows that stage is compute-bound").
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__constant__ double cc[64];

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(double* out, int iters, double b0) {
  // MODE 0: DFMA reg operands; 1: DMMA m8n8k4; 2: mixed per-warp (even warps DFMA, odd DMMA)
  // 3: DFMA with constant-bank operand; 4: DMMA m16n8k8 ; 5: same-warp interleave DFMA+DMMA
  int warp = threadIdx.x >> 5;
  double acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
  double a = b0 + threadIdx.x * 1e-9, b = 1.0 - 1e-9;
  bool do_fma = (MODE == 0) || (MODE == 3) || (MODE == 2 && (warp & 1) == 0) || MODE == 5;
  bool do_mma = (MODE == 1) || (MODE == 4) || (MODE == 2 && (warp & 1) == 1) || MODE == 5;
  double d[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int it = 0; it < iters; ++it) {
    if (do_fma) {
      if (MODE == 3) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], cc[i], cc[8 + i]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], cc[16 + i], cc[24 + i]);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], b, a);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], b, a);
      }
    }
    if (do_mma) {
      if (MODE == 4) {
        // m16n8k8: A 4 regs, B 2 regs, C/D 4 regs
#pragma unroll
        for (int i = 0; i < 2; ++i)
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+d"(d[4 * i]), "+d"(d[4 * i + 1]), "+d"(d[4 * i + 2]), "+d"(d[4 * i + 3])
                       : "d"(a), "d"(b), "d"(a), "d"(b), "d"(b), "d"(a));
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i)
          asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                       : "+d"(d[2 * i]), "+d"(d[2 * i + 1]) : "d"(a), "d"(b));
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += acc[i] + d[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
int run(const char* name, double* out, int ctas_per_sm) {
  int iters = 20000;
  int grid = 148 * ctas_per_sm;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k_pipe<MODE><<<grid, 256>>>(out, 100, 0.5);
  CK(cudaDeviceSynchronize());
  cudaEventRecord(e0);
  k_pipe<MODE><<<grid, 256>>>(out, iters, 0.5);
  cudaEventRecord(e1);
  CK(cudaDeviceSynchronize());
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double nthreads = (double)grid * 256, nwarps = nthreads / 32;
  double fma_flops = 0, mma_flops = 0;
  double fma_frac = (MODE == 0 || MODE == 3 || MODE == 5) ? 1.0 : (MODE == 2 ? 0.5 : 0.0);
  double mma_frac = (MODE == 1 || MODE == 4 || MODE == 5) ? 1.0 : (MODE == 2 ? 0.5 : 0.0);
  fma_flops = nthreads * fma_frac * iters * 16.0 * 2.0;
  double per_iter_mma = (MODE == 4) ? 2.0 * (16 * 8 * 8 * 2) : 4.0 * (8 * 8 * 4 * 2);
  mma_flops = nwarps * mma_frac * iters * per_iter_mma;
  printf("%-28s ctas/sm=%d  %8.3f ms  DFMA %7.2f TF/s  DMMA %7.2f TF/s  total %7.2f TF/s\n", name, ctas_per_sm, ms,
         fma_flops / ms * 1e-9, mma_flops / ms * 1e-9, (fma_flops + mma_flops) / ms * 1e-9);
  return 0;
}

int main() {
  double* out; CK(cudaMalloc(&out, 148 * 8 * 256 * sizeof(double)));
  double h[64]; for (int i = 0; i < 64; ++i) h[i] = 1.0 - 1e-9 * i;
  CK(cudaMemcpyToSymbol(cc, h, sizeof(h)));
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  printf("device %s sm_%d%d SMs %d clock %d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
  for (int c : {2, 4, 8}) {
    run<0>("DFMA reg", out, c);
    run<3>("DFMA const-bank", out, c);
    run<1>("DMMA m8n8k4", out, c);
    run<4>("DMMA m16n8k8", out, c);
    run<2>("mixed warps DFMA|DMMA", out, c);
    run<5>("interleaved DFMA+DMMA", out, c);
  }
  return 0;
}
