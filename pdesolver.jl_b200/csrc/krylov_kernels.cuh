// Device-resident vector kernels of the matrix-free Newton-Krylov solver (configuration 5).
//
// The reference solves dR/dq * delta_q = -R(q) with PETSc's GMRES(30) on a MatShell whose product is the complex-step
// residual (NonlinearSolvers/newton.jl:137-304, newton_setup.jl:632-662, linear solver defaults
// input/read_input.jl:493-496, 560-570).  Here the Krylov basis, the Newton update and every reduction stay in HBM;
// per GMRES iteration the host reads back one small vector (the Hessenberg column) to update its Givens rotations.
//
// All reductions are two-pass and deterministic: per-CTA partials, then one CTA per result sums them in a fixed order.
#pragma once
#include <stdint.h>

namespace pdes {

constexpr int KRY_T = 256;     // threads per CTA of the vector kernels
constexpr int KRY_VB = 8;      // basis vectors per CTA row of k_multi_dot

// partials[i * gridDim.x + blockIdx.x] = sum over this CTA's dofs of V_i[idx] * w[idx],  i in [8*blockIdx.y, +8)
__global__ void __launch_bounds__(KRY_T)
k_multi_dot(const double* __restrict__ V, int64_t ld, int nv, const double* __restrict__ w, int64_t n,
            double* __restrict__ partials, const int* done = nullptr) {
  if (done && *done) return;
  const int i0 = blockIdx.y * KRY_VB;
  double acc[KRY_VB];
#pragma unroll
  for (int u = 0; u < KRY_VB; ++u) acc[u] = 0.0;
  for (int64_t idx = (int64_t)blockIdx.x * KRY_T + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * KRY_T) {
    const double wv = w[idx];
#pragma unroll
    for (int u = 0; u < KRY_VB; ++u)
      if (i0 + u < nv) acc[u] = fma(V[(int64_t)(i0 + u) * ld + idx], wv, acc[u]);
  }
  __shared__ double sh[KRY_VB][KRY_T / 32];
#pragma unroll
  for (int u = 0; u < KRY_VB; ++u) {
    double s = acc[u];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sh[u][threadIdx.x >> 5] = s;
  }
  __syncthreads();
  if (threadIdx.x < KRY_VB && i0 + threadIdx.x < nv) {
    double s = 0.0;
    for (int wp = 0; wp < KRY_T / 32; ++wp) s += sh[threadIdx.x][wp];
    partials[(int64_t)(i0 + threadIdx.x) * gridDim.x + blockIdx.x] = s;
  }
}

// sum_j Minv[node(j)] * r[j]^2  (calcNorm with strongres=true, Utils/Utils.jl:427-449) as CTA partials (row 0)
__global__ void __launch_bounds__(KRY_T)
k_strong_norm_partials(const double* __restrict__ r, const double* __restrict__ minv, int nd, int64_t n,
                       double* __restrict__ partials) {
  double acc = 0.0;
  for (int64_t idx = (int64_t)blockIdx.x * KRY_T + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * KRY_T) {
    const double v = r[idx];
    acc = fma(v * v, minv[idx / nd], acc);
  }
  __shared__ double sh[KRY_T / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int wp = 0; wp < KRY_T / 32; ++wp) s += sh[wp];
    partials[blockIdx.x] = s;
  }
}

// out[i] = sum_b partials[i * B + b]   (one CTA per result, fixed summation order)
__global__ void __launch_bounds__(KRY_T)
k_reduce_rows(const double* __restrict__ partials, int B, double* __restrict__ out) {
  __shared__ double sh[KRY_T];
  const double* p = partials + (int64_t)blockIdx.x * B;
  double s = 0.0;
  for (int b = threadIdx.x; b < B; b += KRY_T) s += p[b];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = KRY_T / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = sh[0];
}

// w -= sum_i h[i] V_i   (Gram-Schmidt projection; h lives on the device: no host round trip between the dot and the update)
__global__ void __launch_bounds__(KRY_T)
k_multi_axpy(const double* __restrict__ V, int64_t ld, int nv, const double* __restrict__ h, double sign,
             double* __restrict__ w, int64_t n, const int* done = nullptr) {
  if (done && *done) return;
  for (int64_t idx = (int64_t)blockIdx.x * KRY_T + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * KRY_T) {
    double s = 0.0;
    for (int i = 0; i < nv; ++i) s = fma(h[i], V[(int64_t)i * ld + idx], s);
    w[idx] = fma(sign, s, w[idx]);
  }
}

// dst = src / sqrt(*norm_sq)   (next basis vector; a zero norm leaves zeros: happy breakdown is detected on the host)
__global__ void __launch_bounds__(KRY_T)
k_normalize(const double* __restrict__ src, const double* __restrict__ norm_sq, double* __restrict__ dst, int64_t n,
            const int* done = nullptr) {
  if (done && *done) return;
  const double nv = *norm_sq;
  const double s = nv > 0.0 ? 1.0 / sqrt(nv) : 0.0;
  for (int64_t idx = (int64_t)blockIdx.x * KRY_T + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * KRY_T)
    dst[idx] = s * src[idx];
}

// y = a*x + b*y  (b = 0: y = a*x without reading y)
__global__ void __launch_bounds__(KRY_T)
k_axpby(double a, const double* __restrict__ x, double b, double* __restrict__ y, int64_t n) {
  for (int64_t idx = (int64_t)blockIdx.x * KRY_T + threadIdx.x; idx < n; idx += (int64_t)gridDim.x * KRY_T)
    y[idx] = b == 0.0 ? a * x[idx] : fma(a, x[idx], b * y[idx]);
}

// ---- Hessenberg column + Givens rotations of one GMRES iteration, on the device ------------------------------------------
// Round 1 read the projection coefficients back every iteration to update the rotations on the host (one stream
// synchronisation per iteration, ~25 % of an iteration on the 60 k-DOF meshes).  Here one thread does that update right
// behind the reductions: column j of H from the two Gram-Schmidt passes, the previous rotations, the new one, the
// residual norm |g_{j+1}| and the convergence / divergence / breakdown / itermax tests (PETSc's default test as in
// gmres_dev).  When a test fires it raises Ctl::kry_done, which turns every Krylov and J*v kernel enqueued behind it into a
// no-op; the host enqueues several iterations per poll of this block.  H (column-major, (m+1) x m), cs, sn, g stay on the
// device until the end of the restart cycle.
struct GmresState {
  double rnorm, bnorm;
  double reltol, abstol, dtol;   // (in device memory, not kernel arguments: the captured graph of an iteration serves every solve)
  long long its, itermax;
  int32_t reason, jdone;      // jdone: columns of the current cycle that are complete
};
__global__ void k_gmres_update(int j, int m, const double* __restrict__ h1, const double* __restrict__ h2,
                               const double* __restrict__ nq, double* __restrict__ H, double* __restrict__ cs,
                               double* __restrict__ sn, double* __restrict__ g, GmresState* st, int* done) {
  if (*done) return;
  const double reltol = st->reltol, abstol = st->abstol, dtol = st->dtol;
  for (int i = 0; i <= j; ++i) H[(size_t)i * m + j] = h1[i] + h2[i];
  const double hn = sqrt(*nq);
  H[(size_t)(j + 1) * m + j] = hn;
  for (int i = 0; i < j; ++i) {
    const double t = cs[i] * H[(size_t)i * m + j] + sn[i] * H[(size_t)(i + 1) * m + j];
    H[(size_t)(i + 1) * m + j] = -sn[i] * H[(size_t)i * m + j] + cs[i] * H[(size_t)(i + 1) * m + j];
    H[(size_t)i * m + j] = t;
  }
  const double a0 = H[(size_t)j * m + j], a1 = hn, d = hypot(a0, a1);
  if (d == 0.0) { cs[j] = 1.0; sn[j] = 0.0; }
  else { cs[j] = a0 / d; sn[j] = a1 / d; }
  H[(size_t)j * m + j] = d;
  H[(size_t)(j + 1) * m + j] = 0.0;
  g[j + 1] = -sn[j] * g[j];
  g[j] = cs[j] * g[j];
  const double rnorm = fabs(g[j + 1]);
  st->rnorm = rnorm;
  st->its += 1;
  st->jdone = j + 1;
  int reason = 0;
  if (rnorm <= reltol * st->bnorm) reason = 1;
  else if (rnorm <= abstol) reason = 2;
  else if (hn == 0.0) reason = 3;
  else if (rnorm >= dtol * st->bnorm) reason = -2;
  else if (st->its >= st->itermax) reason = -1;
  if (reason) { st->reason = reason; *done = 1; }
}

// ---- element-block Jacobi right preconditioner --------------------------------------------------------------------
// The reference preconditions its Krylov solves from the right with PETSc's block Jacobi (read_input.jl:560-570:
// -pc_type bjacobi, -ksp_pc_side right).  Here the blocks are the element-diagonal blocks dR_e/dq_e of the DG Jacobian
// ((nd*nn)^2 doubles each), obtained matrix-free: elements are coloured so that no two face neighbours share a colour, and
// for every colour and every local dof c ONE Jacobian-vector product with the indicator of dof c on the elements of that
// colour yields column c of all their blocks.  The blocks are inverted in place (Gauss-Jordan with partial pivoting, one
// CTA per element) and applied as z_e = B_e^-1 r_e.

// v[e*EL + x] = 1 where x == c and colour[e] == col, else 0
__global__ void __launch_bounds__(KRY_T)
k_probe_set(double* __restrict__ v, const int32_t* __restrict__ colour, int col, int c, int EL, int64_t nE) {
  for (int64_t idx = (int64_t)blockIdx.x * KRY_T + threadIdx.x; idx < nE * EL; idx += (int64_t)gridDim.x * KRY_T) {
    const int64_t e = idx / EL;
    v[idx] = ((int)(idx - e * EL) == c && colour[e] == col) ? 1.0 : 0.0;
  }
}
// column c of the blocks of colour col: blocks[e][c*EL + r] = out[e*EL + r]   (column-major blocks)
__global__ void __launch_bounds__(KRY_T)
k_probe_get(const double* __restrict__ out, const int32_t* __restrict__ colour, int col, int c, int EL, int64_t nE,
            double* __restrict__ blocks) {
  for (int64_t idx = (int64_t)blockIdx.x * KRY_T + threadIdx.x; idx < nE * EL; idx += (int64_t)gridDim.x * KRY_T) {
    const int64_t e = idx / EL;
    if (colour[e] == col) blocks[e * EL * EL + (int64_t)c * EL + (idx - e * EL)] = out[idx];
  }
}
// in-place inverse of every block (shared memory: EL*EL doubles + EL ints)
__global__ void __launch_bounds__(64)
k_block_invert(double* __restrict__ blocks, int EL, int64_t nE) {
  extern __shared__ __align__(16) unsigned char smem_blk[];
  double* A = reinterpret_cast<double*>(smem_blk);                 // A[c*EL + r]
  int* piv = reinterpret_cast<int*>(A + EL * EL);
  __shared__ int s_p;
  const int tid = threadIdx.x;
  double* G = blocks + (int64_t)blockIdx.x * EL * EL;
  for (int i = tid; i < EL * EL; i += 64) A[i] = G[i];
  __syncthreads();
  for (int k = 0; k < EL; ++k) {
    if (tid == 0) {
      int p = k;
      double best = fabs(A[k * EL + k]);
      for (int i = k + 1; i < EL; ++i) {
        const double v = fabs(A[k * EL + i]);
        if (v > best) { best = v; p = i; }
      }
      s_p = p;
      piv[k] = p;
    }
    __syncthreads();
    const int p = s_p;
    if (p != k)
      for (int c = tid; c < EL; c += 64) { const double t = A[c * EL + k]; A[c * EL + k] = A[c * EL + p]; A[c * EL + p] = t; }
    __syncthreads();
    const double pinv = 1.0 / A[k * EL + k];
    __syncthreads();
    for (int c = tid; c < EL; c += 64) A[c * EL + k] = (c == k) ? pinv : A[c * EL + k] * pinv;      // row k
    __syncthreads();
    for (int i = tid; i < EL; i += 64) {
      if (i == k) continue;
      const double f = A[k * EL + i];
      A[k * EL + i] = 0.0;
      for (int c = 0; c < EL; ++c) A[c * EL + i] = fma(-f, A[c * EL + k], A[c * EL + i]);
    }
    __syncthreads();
  }
  for (int k = EL - 1; k >= 0; --k) {        // undo the row interchanges as column interchanges
    const int p = piv[k];
    if (p != k)
      for (int r = tid; r < EL; r += 64) { const double t = A[k * EL + r]; A[k * EL + r] = A[p * EL + r]; A[p * EL + r] = t; }
    __syncthreads();
  }
  for (int i = tid; i < EL * EL; i += 64) G[i] = A[i];
}
// z_e = Binv_e r_e: one thread per (element, row); consecutive rows read consecutive addresses of every column
__global__ void __launch_bounds__(KRY_T)
k_block_apply(const double* __restrict__ blocks, int EL, int64_t nE, const double* __restrict__ r, double* __restrict__ z,
              const int* done = nullptr) {
  if (done && *done) return;
  for (int64_t idx = (int64_t)blockIdx.x * KRY_T + threadIdx.x; idx < nE * EL; idx += (int64_t)gridDim.x * KRY_T) {
    const int64_t e = idx / EL;
    const int row = (int)(idx - e * EL);
    const double* B = blocks + e * EL * EL + row;
    const double* re = r + e * EL;
    double s = 0.0;
    for (int c = 0; c < EL; ++c) s = fma(B[(int64_t)c * EL], re[c], s);
    z[idx] = s;
  }
}

}  // namespace pdes
