// k_element_tma: the element half of one residual evaluation (+ fused RK stage) as a persistent, warp-autonomous
// bulk-copy (TMA) pipeline for sm_100a.
//
//   dataPrep checks, getEulerFlux, weakdifferentiate!          euler.jl:441-519, 543-611, 628-658; euler_funcs.jl:23-58
//   interiorfaceintegrate! / boundaryintegrate! (gather form)  euler.jl:669-690, 770-802 (records written by k_face_flux)
//   applySourceTerm (tabulated), pde_post_func, RK4 / LSERK54 stage update   source.jl:27-47, rk4.jl:244-319, 446-457
//
// Why this shape (profiles/r1_s4_element_rk_c3.txt): k_element_rk walked load -> flux -> products -> records
// -> epilogue serially inside 3-warp CTAs with 56 KB of single-use tiles (18 % of the warp slots, 0.5 eligible warps per
// cycle, 3.2 TB/s).  Every input of a tile is a CONTIGUOUS block, so here
//   * a GROUP of NS warps owns tiles of G elements (G*nd <= 32 lanes: one lane per (element, variable) row) and a two-stage
//     shared-memory ring; lane 0 of the group issues four cp.async.bulk copies (q tile, face-record tile, Minv, dxidx) for
//     the group's NEXT tile before the group waits on the mbarrier of the current one -- no block barrier anywhere, no
//     producer warp, no LDG/LDGSTS on the input side.  The NS warps of a group split the OUTPUT NODES of the operator
//     products (every warp: all rows, a slice of the nodes) and the items of S1 / the epilogue, and meet at a named
//     barrier of 32*NS threads: twice the resident warps for the same shared memory (profiles/r2_c_element_tma_v1.txt:
//     with one warp per tile 11 warps per SM, FP64 pipe 41 %, 125 us);
//   * the volume-flux tile (15 doubles per node) is gone: S1 leaves (U_1..U_dim, p) per node and the row lanes rebuild
//     F_d[k,j] = (q_kj + [k = E] p_j) U_dj + dxidx[d,k-1] p_j on the fly;
//   * one row per lane with the operator coefficients as 16-byte uniform loads from the (even-padded) kernel-parameter
//     table: 11 accumulators instead of 22, so ten such warps fit the register file and the shared memory of an SM;
//   * the results are staged in the (dead) record tile and leave through a coalesced 16-byte epilogue that takes its own
//     input rows (x_old of stage 1, q2 / q4 of the sum-free RK4 update) from the q tile in shared memory.
#pragma once
#include "residual_kernels.cuh"

namespace pdes {

// operator table of k_element_tma: rows padded to an even length so that coefficient pairs are 16-byte aligned
template <int DIM, int NN, int NFN>
struct __align__(16) OpTabP {
  static constexpr int NF = DIM + 1;
  static constexpr int NNP = (NN + 1) & ~1;
  double Qt[DIM * NN][NNP];    // Qt[d*NN+j][i] = sbp.Q[j,i,d]
  double RfN[NF * NFN][NNP];   // RfN[f*NFN+i][node] = sum_j interp[j,i] [perm[j,f]==node]
};

template <int DIM, int NN, int NFN, bool DXN>
struct ElemTmaCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1, EL = NN * ND, FL = NFN * ND, DD = DIM * DIM;
  static constexpr int G = ((32 / ND) / 2) * 2;              // elements per warp tile (even: every block a 16-byte multiple)
  static constexpr int DXE = DXN ? NN * DD : DD;             // doubles of dxidx per element
  static constexpr int QW = G * EL, RW = G * NF * FL, MW = G * NN, XW = G * DXE;
  static constexpr int UPN = 4;                              // (U_1..U_dim, p) per node, 32-byte slots
  static constexpr int UW = G * NN * UPN;
  static constexpr int STAGE = QW + RW + MW + XW;            // doubles per ring stage (records double-buffered with q)
  static constexpr int WS = 2 * STAGE + UW;                  // doubles per group
  // RSB layout (records single-buffered): [q | Minv | dxidx] x 2, one record tile, one (U_d, p) / staging tile
  static constexpr int QSTAGE = QW + MW + XW;
  static constexpr int UOW = UW > QW ? UW : QW;
  static constexpr int WS_RSB = 2 * QSTAGE + RW + UOW;
  // QSB layout (q single-buffered too: the next tile's q is requested when this tile's epilogue has read it; the other
  // warps of the SM cover the latency -- 16 instead of 14 warps per SM, four per scheduler)
  static constexpr int WS_QSB = QSTAGE + RW + UOW;
  static constexpr int HDR = 512;                            // mbarriers: 4 per group
  static constexpr int ws(bool rsb, bool qsb = false) { return qsb ? WS_QSB : (rsb ? WS_RSB : WS); }
  static constexpr int max_groups_l(int smem_budget, bool rsb, bool qsb = false) { return (smem_budget - HDR) / (ws(rsb, qsb) * 8); }
  static_assert(G >= 2 && QW % 2 == 0 && RW % 2 == 0 && MW % 2 == 0 && XW % 2 == 0, "bulk copies need 16-byte multiples");
  static_assert(RW >= QW, "the result rows are staged in the record tile");
  static constexpr int max_groups(int smem_budget) { return (smem_budget - 256) / (WS * 8); }
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long* bar, unsigned parity) {
  // try_wait suspends the thread for a hardware-chosen time slice before it returns false
  unsigned done;
  do {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, unsigned bytes, unsigned long long* bar,
                                              unsigned long long pol) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol) : "memory");
}

// the block [src, src + bytes) into L2 (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}

#ifndef PDES_TMA_DYN
#define PDES_TMA_DYN 1       // (0.890 -> 0.886 ms per step) tiles beyond a warp's first one are drawn from a device counter instead of a fixed stride
#endif
#ifndef PDES_TMA_L2PF
#define PDES_TMA_L2PF 1      // QSB: the inputs of a warp's NEXT tile are pulled into L2 while it works on the current one
#endif

template <int DIM, int NN, int NFN, int U0, int NC>
struct NodeSlice {
  // operator products of one warp: rows (element, variable) x output nodes [U0, U0 + NC)
  static constexpr int ND = DIM + 2, NF = DIM + 1, FL = NFN * ND;
  template <bool DXN>
  static __device__ __forceinline__ void run_s2(const OpTabP<DIM, NN, NFN>& op, const double* sQ, const double* sX,
                                                const double* sU, int sc, int k, double* acc) {
    constexpr int EL = NN * ND, DD = DIM * DIM, UPN = 4;
    static_assert(U0 % 2 == 0 && NC % 2 == 0, "coefficient pairs are 16-byte aligned");
    // S2: res[k,i] = sum_d sum_j Q[j,i,d] F_d[k,j] (weakdifferentiate!, trans=true) with F rebuilt from (q, U_d, p):
    //     F_d[k,j] = (q_kj + [k = E] p_j) U_dj + dxidx[d,k-1] p_j      (calcEulerFlux, euler_funcs.jl:512-536, 749-774)
    const double ek = (k == ND - 1) ? 1.0 : 0.0;
    const bool mom = k >= 1 && k <= DIM;
    double cd[DIM];
    if (!DXN) {
#pragma unroll
      for (int d = 0; d < DIM; ++d) cd[d] = mom ? sX[sc * DD + d + DIM * (k - 1)] : 0.0;
    }
    const double* qrow = sQ + sc * EL + k;
    const double2* uprow = reinterpret_cast<const double2*>(sU + sc * NN * UPN);
    // (fully unrolled: with immediate offsets the coefficients arrive as 16-byte uniform loads, LDCU.128 -> UR operands
    // of the DFMAs; a rolled loop keeps its counter in a vector register and falls back to per-thread LDC.64)
#pragma unroll
    for (int j = 0; j < NN; ++j) {
      const double qkj = qrow[j * ND];
      const double2 u01 = uprow[2 * j], u2p = uprow[2 * j + 1];
      const double Ud[3] = {u01.x, u01.y, u2p.x};
      const double pj = DIM == 2 ? u2p.x : u2p.y;
      if (DXN) {
#pragma unroll
        for (int d = 0; d < DIM; ++d) cd[d] = mom ? sX[(sc * NN + j) * DD + d + DIM * (k - 1)] : 0.0;
      }
      const double base = fma(ek, pj, qkj);
#pragma unroll
      for (int d = 0; d < DIM; ++d) {
        const double f = fma(base, Ud[d], cd[d] * pj);
        const double2* crow = reinterpret_cast<const double2*>(&op.Qt[d * NN + j][U0]);
#pragma unroll
        for (int h = 0; h < NC / 2; ++h) {
          if (U0 + 2 * h >= NN) continue;
          const double2 c = crow[h];
          acc[2 * h] = fma(c.x, f, acc[2 * h]);
          if (U0 + 2 * h + 1 < NN) acc[2 * h + 1] = fma(c.y, f, acc[2 * h + 1]);
        }
      }
    }
  }
  static __device__ __forceinline__ void run_s3(const OpTabP<DIM, NN, NFN>& op, const double* sR, int sc, int k, double* acc) {
    // S3: res[k,node] += sum_f sum_i RfN[f,i][node] * (-+ w_i f*[k,i])  (interiorfaceintegrate!, boundaryintegrate!)
    const double* grow = sR + sc * (NF * FL) + k;
#pragma unroll
    for (int f = 0; f < NF; ++f) {
      double g[NFN];
#pragma unroll
      for (int i = 0; i < NFN; ++i) g[i] = grow[(f * NFN + i) * ND];
#pragma unroll
      for (int i = 0; i < NFN; ++i) {
        const double2* crow = reinterpret_cast<const double2*>(&op.RfN[f * NFN + i][U0]);
#pragma unroll
        for (int h = 0; h < NC / 2; ++h) {
          if (U0 + 2 * h >= NN) continue;
          const double2 c = crow[h];
          acc[2 * h] = fma(c.x, g[i], acc[2 * h]);
          if (U0 + 2 * h + 1 < NN) acc[2 * h + 1] = fma(c.y, g[i], acc[2 * h + 1]);
        }
      }
    }
  }
};

// NP groups of NS warps per CTA; one CTA per SM (persistent), tiles strided over the groups of the grid
template <int DIM, int NN, int NFN, int MODE, bool DXN, int NP, int NS, bool RSB = false, bool HOIST = true, bool QSB = false>
__global__ void __launch_bounds__(32 * NP * NS, 1)
k_element_tma(const __grid_constant__ OpTabP<DIM, NN, NFN> op, const __grid_constant__ ElemArgs a) {
  using Cfg = ElemTmaCfg<DIM, NN, NFN, DXN>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, EL = Cfg::EL, FL = Cfg::FL, DD = Cfg::DD, G = Cfg::G, DXE = Cfg::DXE;
  constexpr int NNP = OpTabP<DIM, NN, NFN>::NNP, UPN = Cfg::UPN;
  constexpr int NC = NS == 1 ? NNP : ((NNP / 2 + NS - 1) / NS) * 2;      // output nodes per warp of a group (even)
  static_assert(NS == 1 || NS == 2, "one or two warps per tile");
  static_assert(NP <= 15 || NS == 1, "one named barrier per group");
  extern __shared__ __align__(128) unsigned char smem_tma[];
  if (a.ctl->stop) return;
  // fused halo: the face launch of this evaluation is complete (stream order) -- close the evaluation
  if (a.halo_epoch && blockIdx.x == 0 && threadIdx.x == 0) *a.halo_epoch += 1u;
  const int lane = threadIdx.x & 31;
  // (the shuffle tells the compiler that the warp index is warp-uniform)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int grp = warp / NS, half = warp - grp * NS;
  // mbarriers per group: [0,1] the two q stages, [2,3] the record tile(s)
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_tma) + 4 * grp;
  static_assert(!QSB || RSB, "QSB implies RSB");
  double* wbase = reinterpret_cast<double*>(smem_tma + Cfg::HDR) + (size_t)grp * (QSB ? Cfg::WS_QSB : (RSB ? Cfg::WS_RSB : Cfg::WS));
  constexpr int QST = QSB ? 0 : (RSB ? Cfg::QSTAGE : Cfg::STAGE);        // stride of a q stage
  double* sRbase = RSB ? wbase + (QSB ? 1 : 2) * Cfg::QSTAGE : wbase + Cfg::QW + Cfg::MW + Cfg::XW;   // record tile (of stage 0)
  double* sU = RSB ? sRbase + Cfg::RW : wbase + 2 * Cfg::STAGE;
  const int64_t ntiles = (a.nE - a.e_begin + G - 1) / G;
  const int64_t W = (int64_t)gridDim.x * NP;
  const int64_t gw = (int64_t)blockIdx.x * NP + grp;
  if (gw >= ntiles) return;                       // (both warps of a group leave together)
  auto group_sync = [&]() {
    if (NS == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(32 * NS) : "memory");
  };
  if (half == 0 && lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_init(&bars[2], 1);
    mbar_init(&bars[3], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  group_sync();
  const double gami = a.ph.gamma - 1.0;
  unsigned long long pol_first = 0;
  if (half == 0 && lane == 0) pol_first = policy_evict_first();

  // tile t of this launch: reverse sweep starts on the elements whose records k_face_flux wrote last (still in L2)
  auto tile_e0 = [&](int64_t t) { return a.e_begin + (a.reverse ? (ntiles - 1 - t) : t) * G; };
  auto tile_ne = [&](int64_t e0) { return (int)((a.nE - e0) < G ? (a.nE - e0) : G); };

  // stage layout: q stage st = [q | Minv | dxidx]; record tile rst (RSB: the only one; else the one of stage st)
  auto q_stage = [&](int st) { return wbase + st * QST; };
  auto r_tile = [&](int st) { return RSB ? sRbase : sRbase + st * QST; };
  // fills q stage st with tile t: bulk copies for a full tile, plain loads for the (single) ragged one
  auto issue_q = [&](int64_t t, int st) {
    const int64_t e0 = tile_e0(t);
    const int ne = tile_ne(e0);
    double* sQ = q_stage(st);
    double* sM = sQ + Cfg::QW;
    double* sX = sM + Cfg::MW;
    const double* gq = a.q + e0 * EL;
    const double* gm = a.minv + e0 * NN;
    const double* gx = a.dxidx + e0 * a.dx_el_stride;
    if (half == 0) {
      if (ne == G) {
        if (lane == 0) {
          // the stage was read / written through the generic proxy by this group (ordered by the barrier that ends a tile)
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          constexpr unsigned bytes = (unsigned)(Cfg::QW + (MODE == EPI_RK ? Cfg::MW : 0) + Cfg::XW) * 8u;
          mbar_expect_tx(&bars[st], bytes);
          bulk_g2s(sQ, gq, Cfg::QW * 8, &bars[st]);
          if (MODE == EPI_RK) bulk_g2s(sM, gm, Cfg::MW * 8, &bars[st]);
          bulk_g2s(sX, gx, Cfg::XW * 8, &bars[st]);
        }
      } else {
        for (int i = lane; i < ne * EL; i += 32) sQ[i] = __ldg(gq + i);
        if (MODE == EPI_RK) for (int i = lane; i < ne * NN; i += 32) sM[i] = __ldg(gm + i);
        for (int i = lane; i < ne * DXE; i += 32) sX[i] = __ldg(gx + i);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[st]);
      }
    }
    // the epilogue's global streams of that tile: in L2 by the time the tile is processed
    if (half == NS - 1) {
      const int64_t b0 = e0 * EL * 8;
      const int nb = ne * EL * 8;
      const int o = lane * 128;
      if (o < nb) {
        if (MODE == EPI_RES) {
          if (a.srcw) prefetch_l2(reinterpret_cast<const char*>(a.srcw) + b0 + o);
        } else {
          const bool s2 = a.scheme == 2;
          if (a.srcm && !(s2 && a.stage == 4)) prefetch_l2(reinterpret_cast<const char*>(a.srcm) + b0 + o);
          if (a.scheme != 1 && a.stage > 1 && !(s2 && a.stage == 4)) prefetch_l2(reinterpret_cast<const char*>(a.x_old) + b0 + o);
          if (a.scheme == 1 ? a.stage > 1 : (s2 ? a.stage == 4 : a.stage > 1))
            prefetch_l2(reinterpret_cast<const char*>(a.ksum) + b0 + o);
        }
      }
    }
  };
  // the face records of tile t into record tile rst
  auto issue_r = [&](int64_t t, int rst) {
    const int64_t e0 = tile_e0(t);
    const int ne = tile_ne(e0);
    double* sR = r_tile(rst);
    const double* gr = a.fluxe + e0 * (NF * FL);
    if (half == 0) {
      if (ne == G) {
        if (lane == 0) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_expect_tx(&bars[2 + rst], (unsigned)Cfg::RW * 8u);
          bulk_g2s_hint(sR, gr, Cfg::RW * 8, &bars[2 + rst], pol_first);       // records: read once, then discarded
        }
      } else {
        for (int i = lane; i < ne * NF * FL; i += 32) sR[i] = __ldg(gr + i);
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars[2 + rst]);
      }
    }
  };

  issue_q(gw, 0);
  issue_r(gw, 0);
  // identical tiles keep the warps of an SM in lockstep (all in the FP64-bound products, then all in the latency-bound
  // epilogue); a staggered start spreads the phases over the tile period
  if (a.stagger_ns > 0) __nanosleep((unsigned)(a.stagger_ns * (grp & 3)));
  int it = 0;
  constexpr bool DYN = PDES_TMA_DYN != 0 && NS == 1;
  // dynamic deal: the first tile of a warp is gw, every further one the next undealt tile (the draw for the tile after next is
  // requested at the top of an iteration and consumed at its end: the atomic's latency is never exposed)
  auto draw = [&]() -> int64_t {
    unsigned v = 0;
    if (lane == 0) v = atomicAdd(a.tile_ctr, 1u);
    return W + (int64_t)__shfl_sync(0xffffffffu, v, 0);
  };
  int64_t tn = DYN ? draw() : gw + W;
#pragma unroll 1
  for (int64_t t = gw; t < ntiles; ++it) {
    int st = QSB ? 0 : (it & 1);
    asm volatile("" : "+r"(st));          // opaque: keeps ONE copy of the (fully unrolled) tile body in the instruction cache
    const bool more = tn < ntiles;
    int64_t tn2 = tn + W;
    if (DYN && more) tn2 = draw();
    if (more) {
      if (!QSB) issue_q(tn, st ^ 1);
      if (!RSB) issue_r(tn, st ^ 1);
    }
    const int64_t e0 = tile_e0(t);
    const int ne = tile_ne(e0);
    double* sQ = q_stage(st);
    double* sM = sQ + Cfg::QW;
    double* sX = sM + Cfg::MW;
    double* sR = r_tile(st);
    if (QSB && PDES_TMA_L2PF && more && half == 0 && lane == 0) {
      // single-buffered tiles are requested late (when their predecessor's rows have been consumed): have DRAM deliver them to
      // L2 now, a whole tile period ahead, so that the bulk copies issued later are L2 hits
      const int64_t e1 = tile_e0(tn);
      if (tile_ne(e1) == G) {
        bulk_prefetch_l2(a.q + e1 * EL, Cfg::QW * 8);
        bulk_prefetch_l2(a.fluxe + e1 * (NF * FL), Cfg::RW * 8);
        if (MODE == EPI_RK) bulk_prefetch_l2(a.minv + e1 * NN, Cfg::MW * 8);
        bulk_prefetch_l2(a.dxidx + e1 * a.dx_el_stride, Cfg::XW * 8);
      }
    }
    mbar_wait_sleep(&bars[st], (unsigned)(QSB ? (it & 1) : ((it >> 1) & 1)));

    // ---- S1 (node items, split over the group): density / pressure checks, pressure, U_d = dxidx[d,:].u ---------------
    // (all items of a lane in one unrolled pass: their shared-memory loads and dependent FP64 chains interleave; one
    // item at a time left the warp idle for three chain latencies per tile -- 18 % of the samples of the first version)
    {
      constexpr int NI = (G * NN + 32 * NS - 1) / (32 * NS);
      double qn[NI][ND], dxv[NI][DD];
      bool on[NI];
#pragma unroll
      for (int r = 0; r < NI; ++r) {
        const int n = r * 32 * NS + half * 32 + lane;
        on[r] = n < ne * NN;
        const int nc = on[r] ? n : 0;
#pragma unroll
        for (int k = 0; k < ND; ++k) qn[r][k] = sQ[nc * ND + k];
        const double* dx = sX + (DXN ? nc * DD : (nc / NN) * DD);
#pragma unroll
        for (int m = 0; m < DD; ++m) dxv[r][m] = dx[m];
      }
#pragma unroll
      for (int r = 0; r < NI; ++r) {
        const int n = r * 32 * NS + half * 32 + lane;
        const double press = calc_pressure<DIM>(qn[r], gami);
        if (on[r] && ((a.ph.check_density && !(qn[r][0] > 0.0)) || (a.ph.check_pressure && !(press > 0.0)))) {
          const int code = (a.ph.check_density && !(qn[r][0] > 0.0)) ? 1 : 2;
          const int s = n / NN;
          const unsigned long long loc = ((unsigned long long)(e0 + s) << 8) | (unsigned)(n - s * NN);
          // density errors win over pressure errors (checkDensity runs first), lowest location wins
          atomicMin(&a.ctl->err_loc, ((unsigned long long)(code - 1) << 62) | loc);
          atomicExch(&a.ctl->err_code, 1);
          atomicExch(&a.ctl->stop, 1);
        }
        const double rinv = fast_rcp(qn[r][0]);
        double up[UPN];
#pragma unroll
        for (int d = 0; d < DIM; ++d) {
          double U = 0.0;
#pragma unroll
          for (int p = 0; p < DIM; ++p) U += qn[r][1 + p] * dxv[r][d + DIM * p];
          up[d] = U * rinv;
        }
        up[DIM] = press;
        if (DIM == 2) up[3] = 0.0;
        if (on[r]) {
          double2* dst = reinterpret_cast<double2*>(sU + n * UPN);
          dst[0] = make_double2(up[0], up[1]);
          dst[1] = make_double2(up[2], up[3]);
        }
      }
    }
    group_sync();

    // ---- epilogue streams of this tile: requested now (L2 hits after the prefetch that accompanied the bulk copies), held in
    // registers across the operator products, consumed by S4 -- the first version issued them at the top of S4 and waited
    // (12 % of its samples on that scoreboard)
    constexpr int NPAIR = (G * EL) / 2;
    constexpr int TG = 32 * NS;
    constexpr int CH = (NPAIR + TG - 1) / TG;
    const int gl = half * 32 + lane;               // thread index inside the group
    const int npair = (ne * EL) / 2;               // ne*EL is even except for odd ragged tiles (handled below)
    const int64_t base = e0 * EL;                  // even: 16-byte aligned in every array
    const double2* q2 = reinterpret_cast<const double2*>(sQ);
    const double* psrc = MODE == EPI_RES ? a.srcw : a.srcm;
    const bool sch2 = a.scheme == 2;
    const bool need_src = psrc && !(MODE == EPI_RK && sch2 && a.stage == 4);
    // x_old: stage 1 of rk4 and every lserk54 stage update the state they were evaluated at (the q tile); the sum-free
    // RK4 update takes q4 from the q tile in stage 4
    const bool xo_smem = MODE == EPI_RK && (a.scheme == 1 || a.stage == 1 || (sch2 && a.stage == 4));
    const bool ks_smem = MODE == EPI_RK && sch2 && a.stage == 2;          // q2: the stage's own input rows
    const bool need_ks = MODE == EPI_RK && (a.scheme == 1 ? a.stage > 1 : (sch2 ? a.stage == 4 : a.stage > 1));
    const bool need_mw = MODE == EPI_RK && a.stage == 1;
    double2 sv[CH], xo[CH], ks[CH], mw[CH];      // (separate registers: two predicated loads into one register serialise)
    auto load_streams = [&]() {
  #pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int i2 = c * TG + gl;
        sv[c] = xo[c] = ks[c] = mw[c] = make_double2(0.0, 0.0);
        if (i2 < npair) {
          const int64_t dof = base + 2 * i2;
          if (need_src) sv[c] = __ldg(reinterpret_cast<const double2*>(psrc + dof));
          if (MODE == EPI_RK) {
            if (need_ks) ks[c] = *reinterpret_cast<const double2*>(a.ksum + dof);
            if (need_mw) {
              // calcNorm weights M = w_j/jac_j (Utils.jl:427-449): node = dof / ND.  (Exclusive with the x_old load below:
              // two loads into the same registers would serialise on the write-after-write hazard.)
              const double* mt = a.mass + e0 * NN;
              mw[c] = make_double2(__ldg(mt + (2 * i2) / ND), __ldg(mt + (2 * i2 + 1) / ND));
            } else if (!xo_smem) {
              xo[c] = __ldg(reinterpret_cast<const double2*>(a.x_old + dof));
            }
          }
        }
      }
    };
    if (HOIST) load_streams();

    // ---- S2 + S3 (row lanes): lane = (element s, variable k); this warp's slice of the output nodes ---------------------
    const int s = lane / ND, k = lane - s * ND;
    const bool act = lane < G * ND && s < ne;
    const int sc = act ? s : 0;                 // idle lanes recompute row (0, k): the products stay warp-convergent,
    double acc[NC];                             // which keeps the operator coefficients on the uniform datapath (LDCU)
#pragma unroll
    for (int u = 0; u < NC; ++u) acc[u] = 0.0;
    const int u0 = half * NC;
    using Slice0 = NodeSlice<DIM, NN, NFN, 0, NC>;
    using Slice1 = NodeSlice<DIM, NN, NFN, (NS == 1 ? 0 : NC), NC>;
    if (NS == 1 || half == 0) Slice0::template run_s2<DXN>(op, sQ, sX, sU, sc, k, acc);
    else Slice1::template run_s2<DXN>(op, sQ, sX, sU, sc, k, acc);
    mbar_wait_sleep(&bars[2 + (RSB ? 0 : st)], (unsigned)(RSB ? (it & 1) : ((it >> 1) & 1)));
    if (NS == 1 || half == 0) Slice0::run_s3(op, sR, sc, k, acc);
    else Slice1::run_s3(op, sR, sc, k, acc);
    if (MODE == EPI_RK) {      // pde_post_func: res_vec *= Minv
#pragma unroll
      for (int u = 0; u < NC; ++u)
        if (u0 + u < NN) acc[u] *= sM[sc * NN + u0 + u];
    }
    group_sync();              // every lane of the group has consumed its records and (U_d, p)
    // RSB: the record tile is free -- the next tile's records start to arrive while this tile is staged and written out;
    // the results are staged in the (dead) (U_d, p) tile.  Otherwise the (dead) record tile is the staging tile.
    if (RSB && more) issue_r(tn, 0);
    double* sOut = RSB ? sU : sR;
    if (act) {
#pragma unroll
      for (int u = 0; u < NC; ++u)
        if (u0 + u < NN) sOut[s * EL + (u0 + u) * ND + k] = acc[u];
    }
    if (a.discard_records && half == NS - 1) {
      // drop the consumed record lines from L2 (nobody reads them again: no write-back)
      const uintptr_t b = reinterpret_cast<uintptr_t>(a.fluxe + e0 * (NF * FL));
      const uintptr_t lo = (b + 127) & ~(uintptr_t)127, hi = (b + (uintptr_t)ne * NF * FL * 8) & ~(uintptr_t)127;
      for (uintptr_t p = lo + (uintptr_t)lane * 128; p < hi; p += 32 * 128)
        asm volatile("discard.global.L2 [%0], 128;" ::"l"(p) : "memory");
    }
    group_sync();

    // ---- S4: coalesced epilogue, two dofs per access, items split over the group -----------------------------------------
    {
      const double2* out2 = reinterpret_cast<const double2*>(sOut);
      if (!HOIST) load_streams();
      double nrm2 = 0.0;
#pragma unroll
      for (int c = 0; c < CH; ++c) {
        const int i2 = c * TG + gl;
        if (i2 >= npair) continue;
        const int64_t dof = base + 2 * i2;
        const double2 v = out2[i2];
        const double2 mwc = mw[c];
        const double2 xoc = xo_smem ? q2[i2] : xo[c];
        const double2 ksc = ks_smem ? q2[i2] : ks[c];
        const double2 kk = make_double2(v.x + sv[c].x, v.y + sv[c].y);
        if (MODE == EPI_RES) {
          *reinterpret_cast<double2*>(a.res + dof) = kk;
          continue;
        }
        if (a.stage == 1) {
          nrm2 = fma(kk.x * mwc.x, kk.x, nrm2);
          nrm2 = fma(kk.y * mwc.y, kk.y, nrm2);
        }
        double2 o1 = kk, o2 = kk;
        bool st1 = false;
        if (a.scheme == 1) {
          // lserk54: dq = a_s*dq + delta_t*res ; q += b_s*dq   (lserk.jl:183-205)
          if (a.stage == 1) o1 = make_double2(a.hh * kk.x, a.hh * kk.y);
          else o1 = make_double2(a.ah * ksc.x + a.hh * kk.x, a.ah * ksc.y + a.hh * kk.y);
          o2 = make_double2(xoc.x + a.h6 * o1.x, xoc.y + a.h6 * o1.y);
          st1 = true;
        } else if (sch2) {
          // classical RK4 without the running sum (see epilogue_tile): stage 2 stores w' = q2 + 2 q3 - x + (h/2) srcm,
          // stage 4 forms x_new = (w' + q4)/3 + (h/6) Minv R(q4)
          if (a.stage == 4) {
            const double third = 1.0 / 3.0;
            o2 = make_double2(fma(a.h6, v.x, (ksc.x + xoc.x) * third), fma(a.h6, v.y, (ksc.y + xoc.y) * third));
          } else {
            o2 = make_double2(xoc.x + a.ah * kk.x, xoc.y + a.ah * kk.y);
            if (a.stage == 2) {
              o1 = make_double2(ksc.x + 2.0 * o2.x - xoc.x + a.ah * sv[c].x, ksc.y + 2.0 * o2.y - xoc.y + a.ah * sv[c].y);
              st1 = true;
            }
          }
        } else {
          // rk4.jl:244-319 with the running sum k1 + 2 k2 + 2 k3 in ksum
          if (a.stage == 1) {
            o1 = kk; st1 = true;
            o2 = make_double2(xoc.x + a.ah * kk.x, xoc.y + a.ah * kk.y);
          } else if (a.stage < 4) {
            o1 = make_double2(ksc.x + 2.0 * kk.x, ksc.y + 2.0 * kk.y); st1 = true;
            o2 = make_double2(xoc.x + a.ah * kk.x, xoc.y + a.ah * kk.y);
          } else {
            o2 = make_double2(xoc.x + a.h6 * (ksc.x + kk.x), xoc.y + a.h6 * (ksc.y + kk.y));
          }
        }
        if (st1) __stcs(reinterpret_cast<double2*>(a.ksum + dof), o1);
        *reinterpret_cast<double2*>(a.q_next + dof) = o2;
      }
      if ((ne * EL) & 1) {
        // odd ragged tile: the last dof of the tile, handled by one lane through the scalar form of the same update
        if (gl == 0) {
          const int i1 = ne * EL - 1;
          const int64_t dof = base + i1;
          const double v = sOut[i1];
          const double svs = need_src ? __ldg(psrc + dof) : 0.0;
          const double kk = v + svs;
          if (MODE == EPI_RES) {
            a.res[dof] = kk;
          } else {
            const double xos = xo_smem ? sQ[i1] : __ldg(a.x_old + dof);
            const double kss = ks_smem ? sQ[i1] : (need_ks ? a.ksum[dof] : 0.0);
            if (a.stage == 1) nrm2 = fma(kk * __ldg(a.mass + e0 * NN + i1 / ND), kk, nrm2);
            double o1 = kk, o2 = kk;
            bool st1 = false;
            if (a.scheme == 1) {
              o1 = a.stage == 1 ? a.hh * kk : a.ah * kss + a.hh * kk;
              o2 = xos + a.h6 * o1; st1 = true;
            } else if (sch2) {
              if (a.stage == 4) o2 = fma(a.h6, v, (kss + xos) * (1.0 / 3.0));
              else {
                o2 = xos + a.ah * kk;
                if (a.stage == 2) { o1 = kss + 2.0 * o2 - xos + a.ah * svs; st1 = true; }
              }
            } else {
              if (a.stage == 1) { o1 = kk; st1 = true; o2 = xos + a.ah * kk; }
              else if (a.stage < 4) { o1 = kss + 2.0 * kk; st1 = true; o2 = xos + a.ah * kk; }
              else o2 = xos + a.h6 * (kss + kk);
            }
            if (st1) a.ksum[dof] = o1;
            a.q_next[dof] = o2;
          }
        }
      }
      if (MODE == EPI_RK && a.stage == 1) {
        // stage-1 norm partials of this tile (deterministic: fixed lane order, one slot per warp of the group)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrm2 += __shfl_xor_sync(0xffffffffu, nrm2, o);
        if (lane == 0) a.norm_partials[(e0 / G) * NS + half] = nrm2;
      }
    }
    group_sync();              // the stage may be refilled by the next iteration's issue
    if (QSB && more) issue_q(tn, 0);
    t = tn;
    tn = tn2;
  }
  if (DYN && lane == 0) {
    // the last warp to finish re-arms the counters for the next launch
    const int64_t nwarps = ntiles < W ? ntiles : W;
    if (atomicAdd(a.tile_ctr + 1, 1u) + 1u == (unsigned)nwarps) { a.tile_ctr[0] = 0u; a.tile_ctr[1] = 0u; }
  }
}

}  // namespace pdes
