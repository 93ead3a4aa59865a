// k_face_tma: the face half of one residual evaluation as a persistent, warp-autonomous bulk-copy (TMA) pipeline.
//
//   interpolateFace / interiorFaceInterpolate!        flux.jl:613-641, 79-125
//   calcFaceFlux + RoeSolver + calcSAT                flux.jl:37-64, bc_solvers.jl:29-420
//   interpolateBoundary + getBCFluxes (BC functors)   bc.jl:49-80, 162-175, 251-284
//   calcSharedFaceIntegrals_nopre_inner (flux part)   flux.jl:264-308
//
// Same arithmetic and the same output (one record per (element, local face)) as k_face_flux.  What changes is how the two
// elements of a face reach the interpolation (profiles/r1_s4_face_flux_c3.txt: k_face_flux issues 22 forty-byte-strided
// LDG per variable thread, ~7 L1 wavefronts each -- the LSU data pipe is 70 % busy -- and meets four block barriers per
// 16-face tile):
//   * every warp owns tiles of FW faces (FW*nd variable lanes, FW*nfn node lanes) and a two-stage shared-memory ring; the
//     element blocks of the warp's NEXT tile are fetched by 2*FW cp.async.bulk copies (one per lane, 16-byte aligned
//     window around the 8*nd*nn-byte block) while the current tile is processed; completion through an mbarrier;
//   * the variable lanes read their columns from shared memory (one 8-byte LDS per node instead of a strided LDG), the
//     face-node permutation comes out of a 4-bit-packed register instead of a table lookup, the interpolation coefficients
//     are 16-byte uniform loads shared by both sides of the face;
//   * only warp-level barriers; the records leave as coalesced 16-byte stores.
#pragma once
#include "residual_kernels.cuh"
#include "element_tma.cuh"

namespace pdes {

template <int DIM, int NN, int NFN>
struct __align__(16) FaceTabP {
  static constexpr int NF = DIM + 1, NOR = (DIM == 2) ? 1 : 3;
  static constexpr int NFNP = (NFN + 1) & ~1;
  double interp[NN][NFNP];              // sbpface.interp[j,i], rows padded to an even length (16-byte coefficient pairs)
  double wface[NFNP];
  unsigned long long perm_pk[NF];       // sum_j perm[j,f] << 4j   (0-based volume node of stencil entry j on face f)
  unsigned long long nbr_pk[NOR];       // sum_i nbrperm[i,orient] << 4i
};

template <int DIM, int NN, int NFN>
struct FaceTmaWCfg {
  static constexpr int ND = DIM + 2, NF = DIM + 1, EL = NN * ND, FL = NFN * ND;
  static constexpr int PER = ND > NFN ? ND : NFN;
  static constexpr int FW = 32 / PER;                                  // faces per warp tile
  static constexpr bool ALIGNED = (EL % 2) == 0;                       // every element block starts 16-byte aligned
  static constexpr int CPY = ALIGNED ? EL * 8 : EL * 8 + 8;            // bytes per bulk copy (from the 16-byte floor)
  static constexpr int SLOT0 = CPY / 8;
  // slot stride in doubles: even (16-byte destinations) and not a multiple of 16 doubles (bank spread of the FW faces)
  static constexpr int SLOTD = (SLOT0 % 16 == 0 || SLOT0 % 16 == 8) ? SLOT0 + 2 : SLOT0;
  static constexpr int FS = (FL + 1) & ~1;                             // per-face stride of a face-state tile (even)
  static constexpr int STAGE = 2 * FW * SLOTD;                         // doubles per ring stage
  static constexpr int WS = 2 * STAGE + 2 * FW * FS;                   // doubles per warp: ring + (sL | sR)
  static_assert(NN <= 16 && NFN <= 16, "4-bit packed permutations");
  static constexpr int max_warps(int smem_budget) { return (smem_budget - 512) / (WS * 8); }
};

// HALO: the launch carries the halo exchange of a partitioned mesh (HaloArgs); the single-GPU instantiation has none of it
// LANECPY: the staged elements arrive by lane-parallel 16-byte cp.async copies (three lanes per element, ten LDGSTS each,
// one address computation per lane and tile) instead of one cp.async.bulk per element.  The bulk copy is a uniform-datapath
// instruction: with per-lane operands ptxas wraps it into an ELECT / R2UR loop over the active lanes, ~14 instructions per
// copy -- 142 of the ~1000 warp instructions of a tile and 15 % of the kernel's stall samples (profiles/r2_final_face_tma.txt).
// Measured (same box): 1.237 vs 0.891 ms per RK4 step -- the 280 LDGSTS of a tile cost the LSU more than the issue loop costs the
// schedulers, and the 128-register cap spills (424 bytes of stack).  Opt-in build switch, like the round-1 finding for
// cp.async staging.
#ifndef PDES_FACE_LANECPY
#define PDES_FACE_LANECPY 0
#endif
template <int DIM, int NN, int NFN, int NW, bool EXTBC, bool HALO = false, bool LANECPY = (PDES_FACE_LANECPY != 0)>
__global__ void __launch_bounds__(32 * NW, 1)
k_face_tma(const __grid_constant__ FaceTabP<DIM, NN, NFN> op, const __grid_constant__ FaceArgs a) {
  using Cfg = FaceTmaWCfg<DIM, NN, NFN>;
  constexpr int ND = Cfg::ND, NF = Cfg::NF, EL = Cfg::EL, FL = Cfg::FL, FW = Cfg::FW, SLOTD = Cfg::SLOTD, FS = Cfg::FS;
  constexpr int NFNP = FaceTabP<DIM, NN, NFN>::NFNP;
  extern __shared__ __align__(128) unsigned char smem_ftma[];
  const HaloArgs& hx = a.halo;
  if (a.ctl->stop) {
    // a rank stopped by an error never sends: tell the neighbours (abort slots) instead of letting them wait for the
    // time-out (a res_tol stop is taken by every rank at the same step head: nobody waits)
    if (HALO && a.ctl->err_code != 0 && blockIdx.x == 0 && (int)threadIdx.x < hx.npeers) {
      __threadfence_system();
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(hx.peer_flags[threadIdx.x] + 32), "r"(1u) : "memory");
    }
    return;
  }
  const int lane = threadIdx.x & 31;
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem_ftma) + 2 * warp;
  double* wbase = reinterpret_cast<double*>(smem_ftma + 512) + (size_t)warp * Cfg::WS;
  double* sL = wbase + 2 * Cfg::STAGE;
  double* sR = sL + FW * FS;
  const int64_t ntiles = (a.ng + FW - 1) / FW;
  const int64_t W = (int64_t)gridDim.x * NW;
  const int64_t gw = (int64_t)blockIdx.x * NW + warp;
  const unsigned hE = HALO ? ld_relaxed_u32(hx.ctr) + 1u : 0u;          // number of this evaluation
  const double* q_recv = HALO ? hx.recv_base + (size_t)(hE & 1u) * hx.nsend : a.q_recv;
  bool h_waited = false;
  if (HALO) {
    // ---- SEND pass (getSendDataFace, Utils/parallel.jl:249-258) before anything else: npk tiles of FW shared faces, dealt
    // to the warps from the END of the grid (the tiles of the main loop are dealt from its start: the warps that own one
    // tile less take the send tiles).  Variable lanes (face, k) interpolate this rank's side and store the states, in its
    // own face-node order, straight into the neighbour's receive buffer (peer memory over NVLink); the warp that completes
    // the pass publishes the evaluation number in every neighbour's flag slot (fence.sys by every storing warp, cumulative
    // through the counter, fence.sys + release by the last one).
    const int64_t npk = (hx.nS + FW - 1) / FW;
    for (int64_t j = W - 1 - gw; j < npk; j += W) {
      const int fi = lane / ND, k = lane - fi * ND;
      const int64_t u = j * FW + fi;
      if (lane < FW * ND && u < hx.nS) {
        const FaceRec r = a.faces[hx.s0 + u];
        const unsigned long long pk = op.perm_pk[r.fL];
        const double* b = a.q + (int64_t)r.elL * EL + k;
        double sv[NFN];
#pragma unroll
        for (int i = 0; i < NFN; ++i) sv[i] = 0.0;
#pragma unroll
        for (int jn = 0; jn < NN; ++jn) {
          const double ql = __ldg(b + (int)((pk >> (4 * jn)) & 15ull) * ND);
#pragma unroll
          for (int i = 0; i < NFN; ++i) sv[i] = fma(op.interp[jn][i], ql, sv[i]);
        }
        double* dst = hx.face_dst[(size_t)(hE & 1u) * hx.nS + u];
#pragma unroll
        for (int i = 0; i < NFN; ++i) dst[i * ND + k] = sv[i];
      }
      __syncwarp();
      if (lane == 0) {
        __threadfence_system();
        const unsigned old = atomicAdd(hx.ctr + 1, 1u);
        if (old + 1u == (unsigned)npk) {
          hx.ctr[1] = 0u;
          __threadfence_system();
          for (int p = 0; p < hx.npeers; ++p)
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(hx.peer_flags[p]), "r"(hE) : "memory");
        }
      }
      __syncwarp();
    }
  }
  if (gw >= ntiles) return;
  if (!LANECPY && lane == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  const int64_t gend = a.g0 + a.ng;

  // lane fi < FW carries the record of face fi of a tile as the raw 16 bytes {elL, elR, fL | fR<<8 | orient<<16 | kind<<24,
  // aux}; fields are unpacked where they are used, two iterations after the load was issued (unpacking at the load
  // exposed its full latency: 9 % of the samples of the first version)
  // (volatile: the load keeps its place at the top of an iteration -- left to the scheduler it sinks to the register moves
  // at the end of the loop body, where its full latency is exposed: 10 % of the samples of the halo instantiation)
  auto ld_rec = [](const FaceRec* p) {
    int4 v;
    asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
  };
  auto load_rec = [&](int64_t t) {
    int4 v = make_int4(0, 0, (int)(255u << 24), 0);
    const int64_t g = a.g0 + t * FW + lane;
    if (t < ntiles && lane < FW && g < gend) v = ld_rec(a.faces + g);
    return v;
  };
  // slot s of a stage = left (s even) / right (s odd) element of face s/2
  constexpr int LPS = 32 / (2 * FW);                               // lanes per slot (LANECPY)
  constexpr int NCH = Cfg::CPY / 16;                               // 16-byte chunks per element
  constexpr int CPL = (NCH + LPS - 1) / LPS;                       // chunks per lane
  auto issue = [&](const int4& rmine, int st) {
    if (LANECPY) {
      const int slot = lane / LPS, part = lane - slot * LPS;
      const int fic = slot < 2 * FW ? (slot >> 1) : 0;
      const int elL = __shfl_sync(0xffffffffu, rmine.x, fic);
      const int elR = __shfl_sync(0xffffffffu, rmine.y, fic);
      const int kind = (int)((unsigned)__shfl_sync(0xffffffffu, rmine.z, fic) >> 24);
      int el = -1;
      if (slot < 2 * FW && kind != 255) el = (slot & 1) ? (kind == FK_INTERIOR ? elR : -1) : elL;
      if (el >= 0) {
        const char* src = reinterpret_cast<const char*>(a.q) + (int64_t)el * (EL * 8) + part * (CPL * 16);
        if (!Cfg::ALIGNED) src -= (el & 1) * 8;
        char* dst = reinterpret_cast<char*>(wbase + st * Cfg::STAGE + slot * SLOTD) + part * (CPL * 16);
#pragma unroll
        for (int c = 0; c < CPL; ++c)
          if (part * CPL + c < NCH) cp_async16(dst + 16 * c, src + 16 * c);
      }
      cp_async_commit();
      return;
    }
    const int fi = lane >> 1;
    const int elL = __shfl_sync(0xffffffffu, rmine.x, fi);
    const int elR = __shfl_sync(0xffffffffu, rmine.y, fi);
    const int kind = (int)((unsigned)__shfl_sync(0xffffffffu, rmine.z, fi) >> 24);
    int el = -1;
    if (lane < 2 * FW && kind != 255) el = (lane & 1) ? (kind == FK_INTERIOR ? elR : -1) : elL;
    const unsigned have = __ballot_sync(0xffffffffu, el >= 0);
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_expect_tx(&bars[st], (unsigned)__popc(have) * (unsigned)Cfg::CPY);
    }
    __syncwarp();
    if (el >= 0) {
      const char* src = reinterpret_cast<const char*>(a.q) + (int64_t)el * (EL * 8);
      if (!Cfg::ALIGNED) src -= (el & 1) * 8;
      bulk_g2s(wbase + st * Cfg::STAGE + lane * SLOTD, src, Cfg::CPY, &bars[st]);
    }
  };

  // record stores (stage C): per pass this lane moves access c_off[p] of record c_rec[p] (tile independent)
  constexpr int PERREC = (FL % 2 == 0) ? FL / 2 : FL;              // accesses per record
  constexpr int NPASS = (2 * FW * PERREC + 31) / 32;
  int c_rec[NPASS], c_off[NPASS], c_src[NPASS];
#pragma unroll
  for (int p = 0; p < NPASS; ++p) {
    const int idx = p * 32 + lane;
    const int rec = idx / PERREC;
    c_off[p] = idx - rec * PERREC;
    c_rec[p] = rec < 2 * FW ? rec : 0;
    // offset (doubles, from sL) of the access inside the face-state tiles: sR follows sL
    c_src[p] = rec < 2 * FW ? ((rec & 1) * FW * FS + (rec >> 1) * FS + c_off[p] * (FL % 2 == 0 ? 2 : 1)) : -1;
  }
  int4 rc = load_rec(gw);               // current tile
  int4 rn = load_rec(gw + W);           // next tile
  issue(rc, 0);
  int it = 0;
#pragma unroll 1
  for (int64_t t = gw; t < ntiles; t += W, ++it) {
    int st = it & 1;
    asm volatile("" : "+r"(st));
    const bool more = t + W < ntiles;
    if (more) issue(rn, st ^ 1);
    const int4 rn2 = load_rec(t + 2 * W);         // records of the tile after next (consumed two iterations later)
    const int64_t g0 = a.g0 + t * FW;
    const int nf = (int)((gend - g0) < FW ? (gend - g0) : FW);
    const double* sQ = wbase + st * Cfg::STAGE;
    if (HALO && !h_waited && g0 + nf > hx.s0) {      // (the shared faces are the tail of the face list)
      // finishExchangeData: every neighbour's states of evaluation hE are in the local receive buffer
      if (lane < hx.npeers) {
        unsigned long long t0, t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
        for (unsigned spin = 0;; ++spin) {
          unsigned v, ab;
          asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(hx.flags + lane) : "memory");
          if ((int)(v - hE) >= 0) break;
          asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(ab) : "l"(hx.flags + 32 + lane) : "memory");
          Ctl* ctl = const_cast<Ctl*>(a.ctl);
          if (ab) { atomicCAS(&ctl->err_code, 0, 5); atomicExch(&ctl->stop, 1); break; }
          if ((spin & 63u) == 63u) {
            if (ld_relaxed_u32(&ctl->stop)) break;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
            if (t1 - t0 > 120000000000ull) { atomicCAS(&ctl->err_code, 0, 4); atomicExch(&ctl->stop, 1); break; }
          }
          __nanosleep(100);
        }
      }
      __syncwarp();
      h_waited = true;
    }

    // ---- node lanes: the normal of this lane's face node (depends on the face number only), requested before the wait
    const int nfi = lane / NFN, ni = lane - nfi * NFN;
    const bool nact = lane < nf * NFN;
    double nrm[DIM];
    if (nact) {
      const double* np_ = a.nrm + (g0 + nfi) * a.nrm_face_stride + ni * a.nrm_node_stride;
#pragma unroll
      for (int d = 0; d < DIM; ++d) nrm[d] = __ldg(np_ + d);
    }
    if (LANECPY) {
      // this lane's copies of the current tile have landed (the group of the next tile may still be in flight);
      // the warp barrier publishes every lane's part to the others
      if (more) cp_async_wait<1>(); else cp_async_wait<0>();
      __syncwarp();
    } else {
      mbar_wait_sleep(&bars[st], (unsigned)((it >> 1) & 1));
    }

    // ---- A: interpolate both sides to the face nodes (variable lanes: lane = (face fi, variable k)) ------------------
    {
      const int fi = lane / ND, k = lane - fi * ND;
      const bool vact = lane < nf * ND;
      const int fic = vact ? fi : 0;
      const int elL = __shfl_sync(0xffffffffu, rc.x, fic);
      const int elR = __shfl_sync(0xffffffffu, rc.y, fic);
      const int aux = __shfl_sync(0xffffffffu, rc.w, fic);
      const int pk4 = __shfl_sync(0xffffffffu, rc.z, fic);
      const int fL = pk4 & 0xff, fR = (pk4 >> 8) & 0xff, orient = (pk4 >> 16) & 0xff, kind = (pk4 >> 24) & 0xff;
      // packed permutations of this lane's faces (uniform table, per-lane select)
      unsigned long long pkL = op.perm_pk[0], pkR = op.perm_pk[0], pkN = op.nbr_pk[0];
#pragma unroll
      for (int f = 1; f < NF; ++f) {
        if (fL == f) pkL = op.perm_pk[f];
        if (fR == f) pkR = op.perm_pk[f];
      }
#pragma unroll
      for (int o = 1; o < FaceTabP<DIM, NN, NFN>::NOR; ++o)
        if (orient == o) pkN = op.nbr_pk[o];
      const bool interior = kind == FK_INTERIOR;
      const double* bL = sQ + (2 * fic) * SLOTD + (Cfg::ALIGNED ? 0 : (elL & 1)) + k;
      const double* bR = sQ + (2 * fic + 1) * SLOTD + (Cfg::ALIGNED ? 0 : (elR & 1)) + k;
      double sLv[NFN], sRv[NFN];
#pragma unroll
      for (int i = 0; i < NFN; ++i) { sLv[i] = 0.0; sRv[i] = 0.0; }
#pragma unroll
      for (int j = 0; j < NN; ++j) {
        const int nl = (int)((pkL >> (4 * j)) & 15ull), nr = (int)((pkR >> (4 * j)) & 15ull);
        const double ql = bL[nl * ND];
        const double qr = interior ? bR[nr * ND] : 0.0;
        const double2* crow = reinterpret_cast<const double2*>(&op.interp[j][0]);
        // both sides share every interpolation coefficient
#pragma unroll
        for (int h = 0; h < NFNP / 2; ++h) {
          const double2 c = crow[h];
          sLv[2 * h] = fma(c.x, ql, sLv[2 * h]);
          sRv[2 * h] = fma(c.x, qr, sRv[2 * h]);
          if (2 * h + 1 < NFN) {
            sLv[2 * h + 1] = fma(c.y, ql, sLv[2 * h + 1]);
            sRv[2 * h + 1] = fma(c.y, qr, sRv[2 * h + 1]);
          }
        }
      }
      if (vact) {
#pragma unroll
        for (int i = 0; i < NFN; ++i) {
          sL[fi * FS + i * ND + k] = sLv[i];
          // elementR's face node i coincides with elementL's face node nbrperm[i,orient] (involution)
          const int ir = (int)((pkN >> (4 * i)) & 15ull);
          if (interior) sR[fi * FS + ir * ND + k] = sRv[i];
        }
        if (kind == FK_SHARED) {
          // permuteinterface! (Utils/parallel.jl:198-201): received node i of the peer is own node nbrperm[i,orient]
          const double* b = q_recv + (int64_t)aux * (NFN * ND) + k;
#pragma unroll
          for (int i = 0; i < NFN; ++i) sR[fi * FS + (int)((pkN >> (4 * i)) & 15ull) * ND + k] = b[i * ND];
        }
      }
    }
    __syncwarp();

    // ---- B: numerical flux at every face node (node lanes) ----------------------------------------------------------
    // results overwrite the face-state tiles: sL <- -w f* in elementL's node order, sR <- +w f* in elementR's node order
    {
      const int fic = nact ? nfi : 0;
      const int elR = __shfl_sync(0xffffffffu, rc.y, fic);
      const int aux = __shfl_sync(0xffffffffu, rc.w, fic);
      const int pk4 = __shfl_sync(0xffffffffu, rc.z, fic);
      const int orient = (pk4 >> 16) & 0xff, kind = (pk4 >> 24) & 0xff;
      double qL[ND], qR[ND], flux[ND];
      if (nact) {
#pragma unroll
        for (int k = 0; k < ND; ++k) { qL[k] = sL[nfi * FS + ni * ND + k]; qR[k] = sR[nfi * FS + ni * ND + k]; }
      }
      __syncwarp();            // every node lane holds its inputs: the tiles may be overwritten
      if (nact) {
        if (kind == FK_BOUNDARY) {
          const double* xp = a.coords_bndry + ((int64_t)elR * NFN + ni) * DIM;   // elR of a boundary face: its index in bndryfaces
          double xb[DIM], nb_[DIM], qb[ND], fb[ND];
#pragma unroll
          for (int d = 0; d < DIM; ++d) { xb[d] = xp[d]; nb_[d] = nrm[d]; }
#pragma unroll
          for (int k = 0; k < ND; ++k) qb[k] = qL[k];
          if (EXTBC) bc_flux_any<DIM>(aux, qb, xb, nb_, a.ph, fb);
          else bc_flux<DIM>(aux, qb, xb, nb_, a.ph, fb);
#pragma unroll
          for (int k = 0; k < ND; ++k) flux[k] = fb[k];
        } else {
          roe_flux<DIM>(qL, qR, nrm, a.ph.gamma, flux);
        }
        const double w = op.wface[ni];
        unsigned long long pkN = op.nbr_pk[0];
#pragma unroll
        for (int o = 1; o < FaceTabP<DIM, NN, NFN>::NOR; ++o)
          if (orient == o) pkN = op.nbr_pk[o];
        const int ir = (kind == FK_INTERIOR) ? (int)((pkN >> (4 * ni)) & 15ull) : ni;
#pragma unroll
        for (int k = 0; k < ND; ++k) {
          const double wf = w * flux[k];
          sL[nfi * FS + ni * ND + k] = -wf;
          sR[nfi * FS + ir * ND + k] = wf;
        }
      }
    }
    __syncwarp();

    // ---- C: one record per (element, local face): 8*ND*NFN contiguous bytes each, coalesced 16-byte stores -----------
    {
      // destination of record `rec` (lane rec < 2*FW): element*NF + local face; -1: no such record
      int dst_mine = -1;
      {
        const int fi = lane < 2 * FW ? (lane >> 1) : 0;
        const int elL = __shfl_sync(0xffffffffu, rc.x, fi);
        const int elR = __shfl_sync(0xffffffffu, rc.y, fi);
        const int pk4 = __shfl_sync(0xffffffffu, rc.z, fi);
        const int fL = pk4 & 0xff, fR = (pk4 >> 8) & 0xff, kind = (pk4 >> 24) & 0xff;
        if (lane < 2 * nf) dst_mine = (lane & 1) ? (kind == FK_INTERIOR ? elR * NF + fR : -1) : elL * NF + fL;
      }
#pragma unroll
      for (int p = 0; p < NPASS; ++p) {
        const int di = __shfl_sync(0xffffffffu, dst_mine, c_rec[p]);
        if (c_src[p] >= 0 && di >= 0) {
          if (FL % 2 == 0)
            reinterpret_cast<double2*>(a.fluxe + (int64_t)di * FL)[c_off[p]] = reinterpret_cast<const double2*>(sL + c_src[p])[0];
          else
            (a.fluxe + (int64_t)di * FL)[c_off[p]] = sL[c_src[p]];
        }
      }
    }
    __syncwarp();              // the ring stage and the face-state tiles may be reused
    rc = rn;
    rn = rn2;
  }
}

}  // namespace pdes
