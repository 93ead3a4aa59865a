// libpdes_euler_b200.so -- C ABI of the B200 Euler residual / RK4 hot path (include/pdes_euler_b200.h).
//
// Host-side runtime: owns the device copies of the operator tables, metrics and connectivity, builds the
// element-centric face table the kernels gather through, sequences the fused residual/RK4-stage kernels
// on a compute stream and the halo exchange (NCCL send/recv) on a communication stream.  There is no CPU
// fallback: without a usable CUDA device every entry point fails.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/pdes_euler_b200.h"
#include "residual_kernels.cuh"
#include "element_tma.cuh"
#include "face_tma.cuh"
#include "es_kernels.cuh"
#include "jvp_kernels.cuh"
#include "generic_kernels.cuh"
#include "krylov_kernels.cuh"
#include "diag_kernels.cuh"

using namespace pdes;

namespace {

thread_local std::string g_last_error;

void set_err(PdesCtx* ctx, const char* fmt, ...);

#define CUDA_TRY(ctx, expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      set_err(ctx, "CUDA error %s at %s:%d (%s)", cudaGetErrorString(_e), __FILE__, __LINE__, #expr); \
      return PDES_ERR_CUDA;                                                                          \
    }                                                                                                \
  } while (0)

// ---- NCCL, resolved at run time so that single-GPU users need no libnccl ---------------------------
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string* why) {
    if (h) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) break;
    }
    if (!h) { *why = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define SYM(field, name) *(void**)(&field) = dlsym(h, name); if (!field) { *why = std::string("missing symbol ") + name; dlclose(h); h = nullptr; return false; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy")
    SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(AllReduce, "ncclAllReduce") SYM(AllGather, "ncclAllGather") SYM(GroupStart, "ncclGroupStart")
    SYM(GroupEnd, "ncclGroupEnd") SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
    return true;
  }
} g_nccl;

struct Peer {
  int32_t rank = -1;
  int64_t nfaces = 0, offset = 0;   // offset into the concatenated shared-face arrays
  std::vector<PdesBoundary> bndries_local;
  std::vector<PdesInterface> ifaces;
  std::vector<double> nrm;
  // element-data halo (face_integral_type 2, parallel_data = element): my elements the peer needs, in the peer's
  // remote-element order; the peer's elements in my halo are numbered shared_el_offset, shared_el_offset + 1, ...
  std::vector<int64_t> send_els;
  int64_t n_recv_el = 0, shared_el_offset = 0, el_send_off = 0, el_recv_off = 0;
  bool have_els = false;
};

int env_int(const char* name, int dflt) {
  const char* v = getenv(name);
  return v ? atoi(v) : dflt;
}

#ifndef PDES_TMA_RSB
#define PDES_TMA_RSB 1      // k_element_tma: face records single-buffered (more resident warp groups)
#endif
#ifndef PDES_TMA_QSB
#define PDES_TMA_QSB 1      // k_element_tma: the q / Minv / dxidx tile single-buffered as well (16 warps per SM: 0.908 -> 0.892 ms per step)
#endif
#ifndef PDES_TMA_HOIST
#define PDES_TMA_HOIST 0    // k_element_tma: epilogue streams requested before the operator products (held in registers)
#endif
#ifndef PDES_TMA_NS
#define PDES_TMA_NS 1       // warps per tile group of k_element_tma (2: measured slower, profiles/r2_d_element_tma_v2.txt)
#endif
#ifndef PDES_SPLIT_DEFAULT
#define PDES_SPLIT_DEFAULT 2     // split-form volume kernel: 0 k_element_split, 1 k_element_split_n, 2 k_element_split_r
#endif

// type-erased launcher for one (DIM, NN, NFN) operator family
struct Ops {
  virtual ~Ops() {}
  virtual void build_tables(const PdesConfig& c, const double* Q, const double* w, const double* interp,
                            const int64_t* perm, const int64_t* nbrperm, const double* wface, int base) = 0;
  virtual cudaError_t launch_faces(const FaceArgs& a, cudaStream_t s) = 0;
  virtual cudaError_t launch_elements(const ElemArgs& a, int mode, cudaStream_t s) = 0;
  virtual cudaError_t launch_pack(const double* q, const int32_t* sh_el, const uint8_t* sh_face, int64_t nS,
                                  double* q_send, double* const* face_dst, const Ctl* ctl, cudaStream_t s) = 0;
  virtual int64_t grid_for(int64_t nelems) const = 0;
  virtual int resident_element_ctas() = 0;   // CTAs of k_element_rk the device holds at once
  virtual int resident_face_ctas() = 0;
  virtual int tile_elems() const = 0;
  virtual int record_doubles() const { return 0; }   // doubles per (element, local face) record; 0: nd * nfn
  virtual int pipe_face_tile() const { return 0; }   // faces per CTA of the chunk-pipeline face kernel; 0: no pipeline
  virtual bool staged_epilogue() const { return false; }   // element kernel stages its epilogue streams (rk4 scheme 2)
  virtual bool fused_halo() const { return false; }        // the face kernel sends / receives the shared-face states itself (HaloArgs)
  virtual cudaError_t prepare() = 0;         // one-time function attributes (must not happen inside a graph capture)
  // fused face + element kernel (k_fused); families without one return 0 faces per group
  virtual int fused_group_faces() const { return 0; }
  virtual int fused_tile_elems() const { return 0; }
  virtual int resident_fused_ctas() { return 0; }
  virtual cudaError_t launch_fused(const FusedArgs& a, int mode, cudaStream_t s) {
    (void)a; (void)mode; (void)s;
    return cudaErrorNotSupported;
  }
  // J*v kernels (dual numbers); cudaErrorNotSupported when the operator family has none
  virtual cudaError_t launch_jvp(const FaceArgs& fa, const ElemArgs& a, const double* v, double* out, cudaStream_t s) {
    (void)fa; (void)a; (void)v; (void)out; (void)s;
    return cudaErrorNotSupported;
  }
};

template <int DIM, int NN, int NFN, int E, int MINB_E, int FT, int MINB_F, int WMINB = 4, int FFT = 16, int NSUB = 4>
struct OpsImpl : Ops {
  using Tab = OpTab<DIM, NN, NFN>;
  using Cfg = TileCfg<DIM, NN, NFN, E>;
  using FCfg = FaceCfg<DIM, NN, NFN, FT>;
  using UCfg = FusedCfg<DIM, NN, NFN, E, FFT, NSUB>;
  int fused_group_faces() const override { return FFT * NSUB; }
  int fused_tile_elems() const override { return E; }
  int pipe_face_tile() const override { return FT; }
  bool staged_epilogue() const override { return use_tma_elem || !use_warp_kernel; }
  bool fused_halo() const override { return use_tma_elem && use_tma_face; }
  // chunk pipeline: programmatic stream serialization lets the CTAs of this launch start while the last wave of the
  // previous launch is still running; the kernels synchronise through PipeArgs counters
  template <typename K, typename A>
  static cudaError_t launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t s, const Tab& tab, const A& a) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, tab, a);
  }
  int resident_fused_ctas() override {
    int per_sm = 0, dev = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_fused<DIM, NN, NFN, E, FFT, NSUB, EPI_RK, MINB_E>, UCfg::T,
                                                  UCfg::smem_bytes);
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return per_sm * sms;
  }
  cudaError_t launch_fused(const FusedArgs& a, int mode, cudaStream_t s) override {
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    const int64_t g = std::max<int64_t>(a.n_groups, (int64_t)a.n_tiles + a.lag);
    if (g <= 0) return cudaSuccess;
    dim3 grid((unsigned)g), block(UCfg::T);
    if (mode == EPI_RES)
      k_fused<DIM, NN, NFN, E, FFT, NSUB, EPI_RES, MINB_E><<<grid, block, UCfg::smem_bytes, s>>>(tab, a);
    else
      k_fused<DIM, NN, NFN, E, FFT, NSUB, EPI_RK, MINB_E><<<grid, block, UCfg::smem_bytes, s>>>(tab, a);
    return cudaGetLastError();
  }
  Tab tab;
  bool attr_set = false;
  bool use_tma = false;
  // k_element_tma (default): persistent warp-autonomous bulk-copy pipeline, G elements per warp tile
  using TabP = OpTabP<DIM, NN, NFN>;
  using PCfg0 = ElemTmaCfg<DIM, NN, NFN, false>;
  using PCfg1 = ElemTmaCfg<DIM, NN, NFN, true>;
  static constexpr int SMEM_MAX = 232448;      // 227 KB: the opt-in dynamic shared memory of an sm_100 CTA
  static constexpr int NSW = PDES_TMA_NS;                // warps per tile group (they split the output nodes of the operator products)
#ifndef PDES_TMA_NWCAP
#define PDES_TMA_NWCAP (PDES_TMA_NS == 1 ? 16 : 15)      // upper bound on the tile groups per CTA of k_element_tma (NS > 1: one named barrier each)
#endif
  static constexpr int nw_cap(int m) { return m > PDES_TMA_NWCAP ? PDES_TMA_NWCAP : (m < 1 ? 1 : m); }
  static constexpr bool RSB = PDES_TMA_RSB != 0, HOIST = PDES_TMA_HOIST != 0, QSB = PDES_TMA_QSB != 0;
  static constexpr int nw_thr(int m) { return m * NSW > 32 ? 32 / NSW : m; }        // <= 1024 threads per CTA
  static constexpr int NW0 = nw_thr(nw_cap(PCfg0::max_groups_l(SMEM_MAX, RSB, QSB))), NW1 = nw_thr(nw_cap(PCfg1::max_groups_l(SMEM_MAX, RSB, QSB)));
  TabP tabp;
  // k_face_tma (default): warp-autonomous face tiles, element blocks staged by bulk copies
  using TabF = FaceTabP<DIM, NN, NFN>;
  using FWCfg = FaceTmaWCfg<DIM, NN, NFN>;
#ifndef PDES_FACE_NW
#define PDES_FACE_NW 16       // warps per CTA of k_face_tma: 16 caps the kernel at 128 registers per thread
#endif
  static constexpr int NWF = FWCfg::max_warps(SMEM_MAX) > PDES_FACE_NW ? PDES_FACE_NW : FWCfg::max_warps(SMEM_MAX);
  static_assert(32 / (2 * FWCfg::FW) >= 1, "k_face_tma: at least one copy lane per staged element");
  TabF tabf;
  bool use_tma_face = true;
  bool face_small = true;
  template <bool EXTBC>
  cudaError_t launch_faces_tma(const FaceArgs& a, cudaStream_t s) {
    const int64_t ntiles = (a.ng + FWCfg::FW - 1) / FWCfg::FW;      // (ng includes the shared faces: >= the send tiles)
    const int64_t nblk = std::min<int64_t>((ntiles + NWF - 1) / NWF, (int64_t)sm_count);
    const size_t smem = 512 + (size_t)NWF * FWCfg::WS * sizeof(double);
    if (a.halo.on) k_face_tma<DIM, NN, NFN, NWF, EXTBC, true><<<dim3((unsigned)nblk), dim3(32 * NWF), smem, s>>>(tabf, a);
    else k_face_tma<DIM, NN, NFN, NWF, EXTBC, false><<<dim3((unsigned)nblk), dim3(32 * NWF), smem, s>>>(tabf, a);
    return cudaGetLastError();
  }
  bool use_tma_elem = true;
  int sm_count = 0;
  void build_tables(const PdesConfig& c, const double* Q, const double* w, const double* interp, const int64_t* perm,
                    const int64_t* nbrperm, const double* wface, int base) override {
    memset(&tab, 0, sizeof(tab));
    use_tma_elem = env_int("PDES_ELEM_TMA", 1) != 0 && env_int("PDES_FUSED", 0) == 0 && env_int("PDES_PIPE", 0) <= 1 &&
                   env_int("PDES_ELEM_W", 0) == 0 && env_int("PDES_MMA", 0) == 0;
    attr_set = false;                       // device copies of the tables are refreshed by the next prepare()
    if (d_ftab) { cudaFree(d_ftab); d_ftab = nullptr; }
    if (d_qr) { cudaFree(d_qr); d_qr = nullptr; }
    use_tma = env_int("PDES_FACE_TMA", 0) != 0;
    use_warp_kernel = env_int("PDES_ELEM_W", 0) != 0;
    const int ss = c.ss;
    for (int d = 0; d < DIM; ++d)
      for (int j = 0; j < NN; ++j)
        for (int i = 0; i < NN; ++i) tab.Qt[d * NN + j][i] = Q[j + NN * (i + NN * d)];
    for (int f = 0; f < DIM + 1; ++f)
      for (int j = 0; j < NN; ++j) tab.perm[f][j] = j < ss ? (int)(perm[j + (int64_t)ss * f] - base) : 0;
    for (int j = 0; j < NN; ++j)
      for (int i = 0; i < NFN; ++i) tab.interp[j][i] = j < ss ? interp[j + ss * i] : 0.0;
    for (int f = 0; f < DIM + 1; ++f)
      for (int i = 0; i < NFN; ++i)
        for (int j = 0; j < ss; ++j) tab.RfN[f * NFN + i][tab.perm[f][j]] += interp[j + ss * i];
    for (int i = 0; i < NFN; ++i) tab.wface[i] = wface[i];
    for (int o = 0; o < Tab::NOR; ++o)
      for (int i = 0; i < NFN; ++i) tab.nbrperm[o][i] = (int)(nbrperm[i + NFN * o] - base);
    use_tma_face = env_int("PDES_FACE_WTMA", 1) != 0 && env_int("PDES_FUSED", 0) == 0 && env_int("PDES_PIPE", 0) <= 1 &&
                   env_int("PDES_FACE_TMA", 0) == 0 && env_int("PDES_FACE_W", 0) == 0 && env_int("PDES_FACE_P", 0) == 0;
    face_small = env_int("PDES_FACE_SMALL", 1) != 0;
    memset(&tabf, 0, sizeof(tabf));
    for (int j = 0; j < NN; ++j)
      for (int i = 0; i < NFN; ++i) tabf.interp[j][i] = tab.interp[j][i];
    for (int i = 0; i < NFN; ++i) tabf.wface[i] = tab.wface[i];
    for (int f = 0; f < DIM + 1; ++f)
      for (int j = 0; j < NN; ++j) tabf.perm_pk[f] |= (unsigned long long)(tab.perm[f][j] & 15) << (4 * j);
    for (int o = 0; o < Tab::NOR; ++o)
      for (int i = 0; i < NFN; ++i) tabf.nbr_pk[o] |= (unsigned long long)(tab.nbrperm[o][i] & 15) << (4 * i);
    memset(&tabp, 0, sizeof(tabp));
    for (int r = 0; r < DIM * NN; ++r)
      for (int i = 0; i < NN; ++i) tabp.Qt[r][i] = tab.Qt[r][i];
    for (int r = 0; r < (DIM + 1) * NFN; ++r)
      for (int i = 0; i < NN; ++i) tabp.RfN[r][i] = tab.RfN[r][i];
    (void)w;
  }
  int64_t grid_for(int64_t nelems) const override {
    if (use_tma_elem) return NSW * ((nelems + PCfg0::G - 1) / PCfg0::G);        // one norm partial per warp of a tile group
    return use_warp_kernel ? (nelems + GW - 1) / GW : (nelems + E - 1) / E;     // norm partials per tile / per warp
  }
  int tile_elems() const override { return use_tma_elem ? PCfg0::G : (use_warp_kernel ? GW * WPC : E); }
  template <int MODE, bool DXN, int NW>
  cudaError_t launch_elements_tma(const ElemArgs& a, cudaStream_t s) {
    using C = ElemTmaCfg<DIM, NN, NFN, DXN>;
    const int64_t ntiles = (a.nE - a.e_begin + C::G - 1) / C::G;
    const int64_t nblk = std::min<int64_t>((ntiles + NW - 1) / NW, (int64_t)sm_count);
    const size_t smem = C::HDR + (size_t)NW * C::ws(RSB, QSB) * sizeof(double);
    k_element_tma<DIM, NN, NFN, MODE, DXN, NW, NSW, RSB, HOIST, QSB><<<dim3((unsigned)nblk), dim3(32 * NW * NSW), smem, s>>>(tabp, a);
    return cudaGetLastError();
  }
  template <int MODE, bool DXN, int NW>
  cudaError_t prepare_tma() {
    using C = ElemTmaCfg<DIM, NN, NFN, DXN>;
    return cudaFuncSetAttribute(k_element_tma<DIM, NN, NFN, MODE, DXN, NW, NSW, RSB, HOIST, QSB>,
                                cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(C::HDR + (size_t)NW * C::ws(RSB, QSB) * sizeof(double)));
  }
  int resident_element_ctas() override {
    int per_sm = 0, dev = 0, sms = 0;
    cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                         (int)Cfg::smem_bytes);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E>, Cfg::T,
                                                  Cfg::smem_bytes);
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return per_sm * sms;
  }
  int resident_face_ctas() override {
    int per_sm = 0, dev = 0, sms = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_face_flux<DIM, NN, NFN, FT, MINB_F>, FCfg::T, 0);
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return per_sm * sms;
  }
  // warp-autonomous element kernel (node-independent metrics only): G elements per warp, WPC warps per CTA
  static constexpr int GW = (32 / (DIM + 2)) * 2, WPC = 4;
  using WCfg = WarpCfg<DIM, NN, NFN, GW, WPC>;
  bool use_warp_kernel = false, w_attr = false;
  cudaError_t launch_elements_w(const ElemArgs& a, int mode, cudaStream_t s) {
    if (!w_attr) {
      cudaFuncSetAttribute(k_element_w<DIM, NN, NFN, GW, WPC, EPI_RES, WMINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)WCfg::smem_bytes);
      cudaFuncSetAttribute(k_element_w<DIM, NN, NFN, GW, WPC, EPI_RK, WMINB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)WCfg::smem_bytes);
      w_attr = true;
    }
    if (a.nE <= a.e_begin) return cudaSuccess;
    const int64_t ngroups = (a.nE - a.e_begin + GW - 1) / GW;
    dim3 grid((unsigned)((ngroups + WPC - 1) / WPC)), block(WCfg::T);
    if (mode == EPI_RES) k_element_w<DIM, NN, NFN, GW, WPC, EPI_RES, WMINB><<<grid, block, WCfg::smem_bytes, s>>>(tab, a);
    else k_element_w<DIM, NN, NFN, GW, WPC, EPI_RK, WMINB><<<grid, block, WCfg::smem_bytes, s>>>(tab, a);
    return cudaGetLastError();
  }
  int tma_grid = -1;
  int32_t* d_ftab = nullptr;       // perm | nbrperm
  // FP64 tensor-core operator products (PDES_MMA=1): the p=2 tet operator with the default tile only
  static constexpr bool HAS_MMA = DIM == 3 && NN == 11 && E == 32 && MINB_E == 4 && FT == 16 && MINB_F == 8 && WMINB == 4;
  bool use_mma = false;
  double* d_qr = nullptr;          // Qt | RfN
  ~OpsImpl() override { cudaFree(d_ftab); cudaFree(d_qr); }
  cudaError_t launch_faces(const FaceArgs& a_in, cudaStream_t s) override {
    FaceArgs a = a_in;
    a.tab_dev = d_ftab;
    if (a.pipe.on) {
      // (an empty chunk still launches one CTA: it has to publish its completion)
      const int64_t nt = std::max<int64_t>(1, (a.ng + FT - 1) / FT);
      return launch_pdl(k_face_flux<DIM, NN, NFN, FT, MINB_F, false, true>, dim3((unsigned)nt), dim3(FCfg::T), 0, s, tab, a);
    }
    if (a.ng <= 0) return cudaSuccess;
    if (use_tma_face) {
      { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
      // a launch with less than one tile per resident warp has nothing to pipeline: the plain tile kernel (same arithmetic,
      // bit-identical records) starts 1.5 us sooner -- C1 (7,600 faces): 51.2 instead of 57.4 us per RK4 step
      // (profiles/r2_c1_small_mesh_ab.txt).  PDES_FACE_SMALL=0 keeps the persistent kernel at every size (the tests do).
      const bool small = face_small && !a.halo.on && (a.ng + FWCfg::FW - 1) / FWCfg::FW < (int64_t)NWF * sm_count;
      if (!small) return a.ext_bc ? launch_faces_tma<true>(a, s) : launch_faces_tma<false>(a, s);
      a.tab_dev = d_ftab;          // (allocated by prepare())
    }
    const int64_t ntiles = (a.ng + FT - 1) / FT;
    if (a.ext_bc) {
      // a boundary functor outside the four scoped ones is in use: the instantiation with the extended dispatch
      dim3 gridx((unsigned)ntiles), blockx(FCfg::T);
      k_face_flux<DIM, NN, NFN, FT, MINB_F, true><<<gridx, blockx, 0, s>>>(tab, a);
      return cudaGetLastError();
    }
    if constexpr (FaceTmaCfg<DIM, NN, NFN, FT>::FITS) if (use_tma) {
      // persistent CTAs (one wave), elements staged by the bulk-copy engine one tile ahead
      if (tma_grid < 0) {
        int per_sm = 0, dev = 0, sms = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_face_flux_tma<DIM, NN, NFN, FT, MINB_F>,
                                                      FaceTmaCfg<DIM, NN, NFN, FT>::T, 0);
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        tma_grid = per_sm * sms > 0 ? per_sm * sms : sms;
      }
      dim3 grid((unsigned)(ntiles < tma_grid ? ntiles : tma_grid)), block(FaceTmaCfg<DIM, NN, NFN, FT>::T);
      k_face_flux_tma<DIM, NN, NFN, FT, MINB_F><<<grid, block, 0, s>>>(tab, a);
      return cudaGetLastError();
    }
    static const int face_w = env_int("PDES_FACE_W", 0);
    if (face_w > 0) {
      // warp-autonomous tiles: FW faces per warp so that FW * max(nd, nfn) <= 32
      constexpr int PER = (DIM + 2) > NFN ? (DIM + 2) : NFN;
      constexpr int FW = 32 / PER, WPC = 4;
      const int64_t nt = (a.ng + FW - 1) / FW;
      dim3 gridw((unsigned)((nt + WPC - 1) / WPC)), blockw(32 * WPC);
      if (face_w == 1) k_face_flux_w<DIM, NN, NFN, FW, WPC, 4><<<gridw, blockw, 0, s>>>(tab, a);
      else k_face_flux_w<DIM, NN, NFN, FW, WPC, 6><<<gridw, blockw, 0, s>>>(tab, a);
      return cudaGetLastError();
    }
    static const int face_p = env_int("PDES_FACE_P", 0);
    if (face_p > 0) {
      // persistent tiles with the next tile's records carried in registers; PDES_FACE_P = CTAs per SM (register cap)
      int dev = 0, sms = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const int per_sm = face_p >= 8 ? 8 : (face_p >= 6 ? 6 : 5);
      dim3 gridp((unsigned)std::min<int64_t>(ntiles, (int64_t)per_sm * sms)), blockp(FCfg::T);
      if (per_sm == 8) k_face_flux_p<DIM, NN, NFN, FT, 8><<<gridp, blockp, 0, s>>>(tab, a);
      else if (per_sm == 6) k_face_flux_p<DIM, NN, NFN, FT, 6><<<gridp, blockp, 0, s>>>(tab, a);
      else k_face_flux_p<DIM, NN, NFN, FT, 5><<<gridp, blockp, 0, s>>>(tab, a);
      return cudaGetLastError();
    }
    dim3 grid((unsigned)ntiles), block(FCfg::T);
    k_face_flux<DIM, NN, NFN, FT, MINB_F><<<grid, block, 0, s>>>(tab, a);
    return cudaGetLastError();
  }
  cudaError_t prepare() override {
    if (attr_set) return cudaSuccess;
    cudaError_t e;
    if (!d_ftab && env_int("PDES_FACE_TAB_DEV", 1)) {
      std::vector<int32_t> h((DIM + 1) * NN + Tab::NOR * NFN);
      memcpy(h.data(), &tab.perm[0][0], sizeof(int32_t) * (DIM + 1) * NN);
      memcpy(h.data() + (DIM + 1) * NN, &tab.nbrperm[0][0], sizeof(int32_t) * Tab::NOR * NFN);
      e = cudaMalloc((void**)&d_ftab, sizeof(int32_t) * h.size());
      if (e != cudaSuccess) return e;
      e = cudaMemcpy(d_ftab, h.data(), sizeof(int32_t) * h.size(), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) return e;
    }
    if constexpr (HAS_MMA) {
      use_mma = env_int("PDES_MMA", 0) != 0;
      if (use_mma && !d_qr) {
        std::vector<double> h((size_t)(DIM * NN + (DIM + 1) * NFN) * NN);
        memcpy(h.data(), &tab.Qt[0][0], sizeof(double) * DIM * NN * NN);
        memcpy(h.data() + DIM * NN * NN, &tab.RfN[0][0], sizeof(double) * (DIM + 1) * NFN * NN);
        e = cudaMalloc((void**)&d_qr, sizeof(double) * h.size());
        if (e != cudaSuccess) return e;
        e = cudaMemcpy(d_qr, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RES, MINB_E, false, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E, false, true>,
                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
        if (e != cudaSuccess) return e;
      }
    }
    {
      int dev = 0;
      cudaGetDevice(&dev);
      cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
      if (sm_count <= 0) sm_count = 148;
      {
        const int fsm = (int)(512 + (size_t)NWF * FWCfg::WS * sizeof(double));
        if ((e = cudaFuncSetAttribute(k_face_tma<DIM, NN, NFN, NWF, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_face_tma<DIM, NN, NFN, NWF, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_face_tma<DIM, NN, NFN, NWF, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm)) != cudaSuccess) return e;
        if ((e = cudaFuncSetAttribute(k_face_tma<DIM, NN, NFN, NWF, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsm)) != cudaSuccess) return e;
      }
      if ((e = prepare_tma<EPI_RES, false, NW0>()) != cudaSuccess) return e;
      if ((e = prepare_tma<EPI_RK, false, NW0>()) != cudaSuccess) return e;
      if ((e = prepare_tma<EPI_RES, true, NW1>()) != cudaSuccess) return e;
      if ((e = prepare_tma<EPI_RK, true, NW1>()) != cudaSuccess) return e;
    }
    e = cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RES, MINB_E>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RES, MINB_E, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E, true>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    {
      // shared-memory carve-out of the SM: k_element_rk needs the largest one, k_face_flux by itself would get the
      // smallest (largest L1).  An SM cannot change its carve-out while CTAs are resident, so launches that should
      // overlap (chunk pipeline) must agree on it: PDES_CARVEOUT_F / _E = percent of the maximum, -1 = driver default
      const int cf = env_int("PDES_CARVEOUT_F", -1), ce = env_int("PDES_CARVEOUT_E", -1);
      if (cf >= 0) {
        cudaFuncSetAttribute(k_face_flux<DIM, NN, NFN, FT, MINB_F>, cudaFuncAttributePreferredSharedMemoryCarveout, cf);
        cudaFuncSetAttribute(k_face_flux<DIM, NN, NFN, FT, MINB_F, false, true>, cudaFuncAttributePreferredSharedMemoryCarveout, cf);
      }
      if (ce >= 0) {
        cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E>, cudaFuncAttributePreferredSharedMemoryCarveout, ce);
        cudaFuncSetAttribute(k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E, true>, cudaFuncAttributePreferredSharedMemoryCarveout, ce);
      }
    }
    e = cudaFuncSetAttribute(k_fused<DIM, NN, NFN, E, FFT, NSUB, EPI_RES, MINB_E>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_fused<DIM, NN, NFN, E, FFT, NSUB, EPI_RK, MINB_E>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)UCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    attr_set = true;
    return cudaSuccess;
  }
  cudaError_t launch_elements(const ElemArgs& a, int mode, cudaStream_t s) override {
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    if (a.pipe.on) {
      dim3 gridp((unsigned)((a.nE - a.e_begin + E - 1) / E)), blockp(Cfg::T);
      if (mode == EPI_RES)
        return launch_pdl(k_element_rk<DIM, NN, NFN, E, EPI_RES, MINB_E, true>, gridp, blockp, Cfg::smem_bytes, s, tab, a);
      return launch_pdl(k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E, true>, gridp, blockp, Cfg::smem_bytes, s, tab, a);
    }
    if (use_warp_kernel && a.dx_node_stride == 0) return launch_elements_w(a, mode, s);
    if (a.nE <= a.e_begin) return cudaSuccess;
    if (use_tma_elem) {
      if (a.dx_node_stride == 0)
        return mode == EPI_RES ? launch_elements_tma<EPI_RES, false, NW0>(a, s) : launch_elements_tma<EPI_RK, false, NW0>(a, s);
      return mode == EPI_RES ? launch_elements_tma<EPI_RES, true, NW1>(a, s) : launch_elements_tma<EPI_RK, true, NW1>(a, s);
    }
    if constexpr (HAS_MMA) if (use_mma && d_qr) {
      // operator products on the FP64 tensor-core path (element_tile<..., MMA>)
      ElemArgs b = a;
      b.s2_dev = d_qr;
      dim3 gridm((unsigned)grid_for(a.nE - a.e_begin)), blockm(Cfg::T);
      if (mode == EPI_RES)
        k_element_rk<DIM, NN, NFN, E, EPI_RES, MINB_E, false, true><<<gridm, blockm, Cfg::smem_bytes, s>>>(tab, b);
      else
        k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E, false, true><<<gridm, blockm, Cfg::smem_bytes, s>>>(tab, b);
      return cudaGetLastError();
    }
    dim3 grid((unsigned)grid_for(a.nE - a.e_begin)), block(Cfg::T);
    if (mode == EPI_RES)
      k_element_rk<DIM, NN, NFN, E, EPI_RES, MINB_E><<<grid, block, Cfg::smem_bytes, s>>>(tab, a);
    else
      k_element_rk<DIM, NN, NFN, E, EPI_RK, MINB_E><<<grid, block, Cfg::smem_bytes, s>>>(tab, a);
    return cudaGetLastError();
  }
  cudaError_t launch_pack(const double* q, const int32_t* sh_el, const uint8_t* sh_face, int64_t nS, double* q_send,
                          double* const* face_dst, const Ctl* ctl, cudaStream_t s) override {
    if (nS <= 0) return cudaSuccess;
    int64_t n = nS * NFN * (DIM + 2);
    k_pack_send<DIM, NN, NFN><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tab, q, sh_el, sh_face, nS, q_send, face_dst, ctl);
    return cudaGetLastError();
  }
  cudaError_t launch_jvp(const FaceArgs& fa, const ElemArgs& a, const double* v, double* out, cudaStream_t s) override {
    const int64_t nfn = fa.ng * NFN, nen = a.nE * NN;
    if (nfn > 0) k_jvp_face<DIM, NN, NFN><<<(unsigned)((nfn + 127) / 128), 128, 0, s>>>(tab, fa, v);
    k_jvp_element<DIM, NN, NFN><<<(unsigned)((nen + 127) / 128), 128, 0, s>>>(tab, a, v, out);
    return cudaGetLastError();
  }
};

Ops* make_generic_ops(const PdesConfig& c);

// SBPDiagonalE operators (sparse faces) with the split-form entropy-stable volume integral (config 2)
template <int DIM, int NN, int NFN, int E>
struct OpsImplS : Ops {
  using Tab = OpTabS<DIM, NN, NFN>;
  using Cfg = SplitCfg<DIM, NN, NFN, E>;
  Tab tab;
  int flux_id = FLUX_IRSLF;
  bool attr_set = false;
  void build_tables(const PdesConfig& c, const double* Q, const double* w, const double* interp, const int64_t* perm,
                    const int64_t* nbrperm, const double* wface, int base) override {
    memset(&tab, 0, sizeof(tab));
    attr_set = false;
    flux_id = c.flux_id;
    for (int d = 0; d < DIM; ++d)
      for (int i = 0; i < NN; ++i)
        for (int m = 0; m < NN; ++m) tab.S2[d][i][m] = Q[i + NN * (m + NN * d)] - Q[m + NN * (i + NN * d)];
    for (int n = 0; n < NN; ++n)
      for (int u = 0; u < DIM; ++u) tab.inv[n][u] = -1;
    for (int f = 0; f < DIM + 1; ++f)
      for (int i = 0; i < NFN; ++i) {
        const int n = (int)(perm[i + (int64_t)NFN * f] - base);
        tab.perm[f][i] = n;
        for (int u = 0; u < DIM; ++u)
          if (tab.inv[n][u] < 0) { tab.inv[n][u] = f * NFN + i; break; }
      }
    for (int i = 0; i < NFN; ++i) tab.wface[i] = wface[i];
    for (int o = 0; o < Tab::NOR; ++o)
      for (int i = 0; i < NFN; ++i) tab.nbrperm[o][i] = (int)(nbrperm[i + NFN * o] - base);
    jv.reset(make_generic_ops(c));
    if (jv) jv->build_tables(c, Q, w, interp, perm, nbrperm, wface, base);
  }
  int64_t grid_for(int64_t nelems) const override { return (nelems + E - 1) / E; }
  int tile_elems() const override { return E; }
  int resident_element_ctas() override { return 0; }
  int resident_face_ctas() override { return 0; }
  // J*v of the entropy-stable configuration: the dual-number instantiations of the size-generic kernels
  std::unique_ptr<Ops> jv;
  cudaError_t launch_jvp(const FaceArgs& fa, const ElemArgs& a, const double* v, double* out, cudaStream_t s) override {
    return jv ? jv->launch_jvp(fa, a, v, out, s) : cudaErrorNotSupported;
  }
  cudaError_t launch_faces(const FaceArgs& a, cudaStream_t s) override {
    if (a.ng <= 0) return cudaSuccess;
    const int64_t n = a.ng * NFN;
    k_face_flux_sparse<DIM, NN, NFN><<<(unsigned)((n + 127) / 128), 128, 0, s>>>(tab, a, flux_id);
    return cudaGetLastError();
  }
  cudaError_t prepare() override {
    if (attr_set) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_element_split<DIM, NN, NFN, E, EPI_RES>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split<DIM, NN, NFN, E, EPI_RK>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split_n<DIM, NN, NFN, E, EPI_RES, false, NMINB>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split_n<DIM, NN, NFN, E, EPI_RK, false, NMINB>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split_r<DIM, NN, NFN, E, EPI_RES, false, NMINB>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split_r<DIM, NN, NFN, E, EPI_RK, false, NMINB>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    node_centric = env_int("PDES_SPLIT_N", PDES_SPLIT_DEFAULT);
    e = upload_tab();
    if (e != cudaSuccess) return e;
    attr_set = true;
    return cudaSuccess;
  }
  // device copy of S2 | inv for k_element_split_r
  double* d_s2 = nullptr;
  int32_t* d_inv = nullptr;
  cudaError_t upload_tab() {
    if (!d_s2) {
      cudaError_t e = cudaMalloc((void**)&d_s2, sizeof(tab.S2));
      if (e != cudaSuccess) return e;
      e = cudaMalloc((void**)&d_inv, sizeof(tab.inv));
      if (e != cudaSuccess) return e;
    }
    cudaError_t e = cudaMemcpy(d_s2, &tab.S2[0][0][0], sizeof(tab.S2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(d_inv, &tab.inv[0][0], sizeof(tab.inv), cudaMemcpyHostToDevice);
  }
  ~OpsImplS() override { cudaFree(d_s2); cudaFree(d_inv); }
  // node-centric split-form kernels: k_element_split_n (1: every two-point flux evaluated at both of its end points) and
  // k_element_split_r (2: every flux once, round-robin pair schedule, exchange through shared memory)
  using RCfg = SplitRCfg<DIM, NN, NFN, E>;
  using NCfg = SplitNCfg<DIM, NN, NFN, E>;
#ifdef PDES_SPLITN_MINB
  static constexpr int NMINB = PDES_SPLITN_MINB;
#else
  static constexpr int NMINB = NCfg::T <= 96 ? 8 : 4;      // 80 registers per thread either way
#endif
  int node_centric = 1;
  cudaError_t launch_elements(const ElemArgs& a, int mode, cudaStream_t s) override {
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    if (a.nE <= a.e_begin) return cudaSuccess;
    if (node_centric == 2) {
      dim3 gridr((unsigned)grid_for(a.nE - a.e_begin)), blockr(RCfg::T);
      ElemArgs b = a;
      b.s2_dev = d_s2; b.inv_dev = d_inv;
      if (mode == EPI_RES) k_element_split_r<DIM, NN, NFN, E, EPI_RES, false, NMINB><<<gridr, blockr, RCfg::smem_bytes, s>>>(tab, b);
      else k_element_split_r<DIM, NN, NFN, E, EPI_RK, false, NMINB><<<gridr, blockr, RCfg::smem_bytes, s>>>(tab, b);
      return cudaGetLastError();
    }
    if (node_centric) {
      dim3 gridn((unsigned)grid_for(a.nE - a.e_begin)), blockn(NCfg::T);
      if (mode == EPI_RES) k_element_split_n<DIM, NN, NFN, E, EPI_RES, false, NMINB><<<gridn, blockn, NCfg::smem_bytes, s>>>(tab, a);
      else k_element_split_n<DIM, NN, NFN, E, EPI_RK, false, NMINB><<<gridn, blockn, NCfg::smem_bytes, s>>>(tab, a);
      return cudaGetLastError();
    }
    dim3 grid((unsigned)grid_for(a.nE - a.e_begin)), block(Cfg::T);
    if (mode == EPI_RES) k_element_split<DIM, NN, NFN, E, EPI_RES><<<grid, block, Cfg::smem_bytes, s>>>(tab, a);
    else k_element_split<DIM, NN, NFN, E, EPI_RK><<<grid, block, Cfg::smem_bytes, s>>>(tab, a);
    return cudaGetLastError();
  }
  cudaError_t launch_pack(const double* q, const int32_t* sh_el, const uint8_t* sh_face, int64_t nS, double* q_send,
                          double* const* face_dst, const Ctl* ctl, cudaStream_t s) override {
    if (nS <= 0) return cudaSuccess;
    int64_t n = nS * NFN * (DIM + 2);
    k_pack_send_sparse<DIM, NN, NFN><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tab, q, sh_el, sh_face, nS, q_send, face_dst, ctl);
    return cudaGetLastError();
  }
};

// SBP-Omega (dense face) operators with the entropy-stable scheme of face_integral_type = 2: split-form volume
// integrals (k_element_split, dense records) + face-element integrals (k_face_element)
template <int DIM, int NN, int NFN, int E>
struct OpsImplE : Ops {
  using Tab = OpTab<DIM, NN, NFN>;
  using TabS = OpTabS<DIM, NN, NFN>;
  using Cfg = SplitCfg<DIM, NN, NFN, E>;
  Tab tab;
  TabS tabs;
  int fei = FEI_ESLF;
  bool attr_set = false;
  void build_tables(const PdesConfig& c, const double* Q, const double* w, const double* interp, const int64_t* perm,
                    const int64_t* nbrperm, const double* wface, int base) override {
    memset(&tab, 0, sizeof(tab));
    memset(&tabs, 0, sizeof(tabs));
    attr_set = false;
    fei = c.face_element_id;
    const int ss = c.ss;
    for (int d = 0; d < DIM; ++d)
      for (int i = 0; i < NN; ++i)
        for (int m = 0; m < NN; ++m) tabs.S2[d][i][m] = Q[i + NN * (m + NN * d)] - Q[m + NN * (i + NN * d)];
    for (int f = 0; f < DIM + 1; ++f)
      for (int j = 0; j < NN; ++j) tab.perm[f][j] = j < ss ? (int)(perm[j + (int64_t)ss * f] - base) : 0;
    for (int j = 0; j < NN; ++j)
      for (int i = 0; i < NFN; ++i) tab.interp[j][i] = j < ss ? interp[j + ss * i] : 0.0;
    for (int i = 0; i < NFN; ++i) tab.wface[i] = wface[i];
    for (int o = 0; o < Tab::NOR; ++o)
      for (int i = 0; i < NFN; ++i) tab.nbrperm[o][i] = (int)(nbrperm[i + NFN * o] - base);
    (void)w;
  }
  int64_t grid_for(int64_t nelems) const override { return (nelems + E - 1) / E; }
  int tile_elems() const override { return E; }
  int resident_element_ctas() override { return 0; }
  int resident_face_ctas() override { return 0; }
  int record_doubles() const override { return NN * (DIM + 2); }
  int32_t* d_ftab = nullptr;       // perm | nbrperm
  double* d_otab = nullptr;        // interp | wface
  // faces per CTA of k_face_element_b: as many as fit in ~48 KB of shared memory (4 CTAs per SM), at most 16.  Measured
  // (tools/r2_fei_fb.sh, ESLF, DOF-evals/s with 2 / 4 / 8 faces): p=2 tets 2.9 / 3.8 / 3.0e9 (8 faces = 94 KB: 2 CTAs per SM),
  // p=2 triangles 3.3 / 5.5 / 7.6e9, p=1 tets 1.8 / 3.2 / 4.8e9; one face per CTA (k_face_element): 2.7 / 2.4 / 1.3e9
#ifdef PDES_FEI_FB
  static constexpr int FEB = PDES_FEI_FB;
#else
  static constexpr int feb_fit = (int)(49152 / (FaceElemBCfg<DIM, NN, NFN, 1>::PER * sizeof(double)));
  static constexpr int FEB = feb_fit > 16 ? 16 : (feb_fit < 2 ? 2 : (feb_fit & ~1));
#endif
  cudaError_t launch_faces(const FaceArgs& a_in, cudaStream_t s) override {
    if (a_in.ng <= 0) return cudaSuccess;
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    FaceArgs a = a_in;
    a.tab_dev = d_ftab; a.optab_dev = d_otab;
    // FB faces per CTA (default; PDES_FEI_BATCH=0: the one-face-per-CTA kernel)
    static const int batch = env_int("PDES_FEI_BATCH", 1);
    if (batch) {
      using BC = FaceElemBCfg<DIM, NN, NFN, FEB>;
      k_face_element_b<DIM, NN, NFN, FEB><<<(unsigned)((a.ng + FEB - 1) / FEB), 128, BC::smem_bytes, s>>>(tab, a, fei);
      return cudaGetLastError();
    }
    k_face_element<DIM, NN, NFN><<<(unsigned)a.ng, 128, 0, s>>>(tab, a, fei);
    return cudaGetLastError();
  }
  cudaError_t prepare() override {
    if (attr_set) return cudaSuccess;
    cudaError_t e = cudaFuncSetAttribute(k_element_split<DIM, NN, NFN, E, EPI_RES, true>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_face_element_b<DIM, NN, NFN, FEB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)FaceElemBCfg<DIM, NN, NFN, FEB>::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split<DIM, NN, NFN, E, EPI_RK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)Cfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split_n<DIM, NN, NFN, E, EPI_RES, true, NMINB>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split_n<DIM, NN, NFN, E, EPI_RK, true, NMINB>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)NCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split_r<DIM, NN, NFN, E, EPI_RES, true, NMINB>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_element_split_r<DIM, NN, NFN, E, EPI_RK, true, NMINB>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RCfg::smem_bytes);
    if (e != cudaSuccess) return e;
    node_centric = env_int("PDES_SPLIT_N", PDES_SPLIT_DEFAULT);
    if (!d_s2) {
      e = cudaMalloc((void**)&d_s2, sizeof(tabs.S2));
      if (e != cudaSuccess) return e;
      e = cudaMalloc((void**)&d_inv, sizeof(tabs.inv));
      if (e != cudaSuccess) return e;
    }
    e = cudaMemcpy(d_s2, &tabs.S2[0][0][0], sizeof(tabs.S2), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    e = cudaMemcpy(d_inv, &tabs.inv[0][0], sizeof(tabs.inv), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return e;
    {
      std::vector<int32_t> hi((DIM + 1) * NN + Tab::NOR * NFN);
      memcpy(hi.data(), &tab.perm[0][0], sizeof(int32_t) * (DIM + 1) * NN);
      memcpy(hi.data() + (DIM + 1) * NN, &tab.nbrperm[0][0], sizeof(int32_t) * Tab::NOR * NFN);
      std::vector<double> hd(NN * NFN + NFN);
      memcpy(hd.data(), &tab.interp[0][0], sizeof(double) * NN * NFN);
      memcpy(hd.data() + NN * NFN, &tab.wface[0], sizeof(double) * NFN);
      if (!d_ftab) {
        e = cudaMalloc((void**)&d_ftab, sizeof(int32_t) * hi.size());
        if (e != cudaSuccess) return e;
        e = cudaMalloc((void**)&d_otab, sizeof(double) * hd.size());
        if (e != cudaSuccess) return e;
      }
      e = cudaMemcpy(d_ftab, hi.data(), sizeof(int32_t) * hi.size(), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) return e;
      e = cudaMemcpy(d_otab, hd.data(), sizeof(double) * hd.size(), cudaMemcpyHostToDevice);
      if (e != cudaSuccess) return e;
    }
    attr_set = true;
    return cudaSuccess;
  }
  double* d_s2 = nullptr;
  int32_t* d_inv = nullptr;
  ~OpsImplE() override { cudaFree(d_s2); cudaFree(d_inv); cudaFree(d_ftab); cudaFree(d_otab); }
  using NCfg = SplitNCfg<DIM, NN, NFN, E>;
  using RCfg = SplitRCfg<DIM, NN, NFN, E>;
  static constexpr int NMINB = 3;
  int node_centric = 1;
  cudaError_t launch_elements(const ElemArgs& a, int mode, cudaStream_t s) override {
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    if (a.nE <= a.e_begin) return cudaSuccess;
    if (node_centric == 2) {
      dim3 gridr((unsigned)grid_for(a.nE - a.e_begin)), blockr(RCfg::T);
      ElemArgs b = a;
      b.s2_dev = d_s2; b.inv_dev = d_inv;
      if (mode == EPI_RES) k_element_split_r<DIM, NN, NFN, E, EPI_RES, true, NMINB><<<gridr, blockr, RCfg::smem_bytes, s>>>(tabs, b);
      else k_element_split_r<DIM, NN, NFN, E, EPI_RK, true, NMINB><<<gridr, blockr, RCfg::smem_bytes, s>>>(tabs, b);
      return cudaGetLastError();
    }
    if (node_centric) {
      dim3 gridn((unsigned)grid_for(a.nE - a.e_begin)), blockn(NCfg::T);
      if (mode == EPI_RES) k_element_split_n<DIM, NN, NFN, E, EPI_RES, true, NMINB><<<gridn, blockn, NCfg::smem_bytes, s>>>(tabs, a);
      else k_element_split_n<DIM, NN, NFN, E, EPI_RK, true, NMINB><<<gridn, blockn, NCfg::smem_bytes, s>>>(tabs, a);
      return cudaGetLastError();
    }
    dim3 grid((unsigned)grid_for(a.nE - a.e_begin)), block(Cfg::T);
    if (mode == EPI_RES) k_element_split<DIM, NN, NFN, E, EPI_RES, true><<<grid, block, Cfg::smem_bytes, s>>>(tabs, a);
    else k_element_split<DIM, NN, NFN, E, EPI_RK, true><<<grid, block, Cfg::smem_bytes, s>>>(tabs, a);
    return cudaGetLastError();
  }
  cudaError_t launch_pack(const double*, const int32_t*, const uint8_t*, int64_t nS, double*, double* const*, const Ctl*,
                          cudaStream_t) override {
    return nS <= 0 ? cudaSuccess : cudaErrorNotSupported;
  }
};

// Any other operator createSBPOperator can build (solver/common.jl:276-390): run-time sizes, tables in global memory
// (generic_kernels.cuh).  dense faces + Roe flux (also J*v), or sparse faces + split form with the IR volume flux.
template <int DIM>
struct OpsGeneric : Ops {
  GenTab tab{};
  bool split = false, uploaded = false;
  int flux_id = FLUX_ROE;
  std::vector<double> hd;      // Qt | RfN | interp | wface | S2
  std::vector<int32_t> hi;     // perm | nbrperm | inv
  size_t o_Qt = 0, o_RfN = 0, o_interp = 0, o_wface = 0, o_S2 = 0, o_perm = 0, o_nbr = 0, o_inv = 0;
  double* d_d = nullptr;
  int32_t* d_i = nullptr;
  ~OpsGeneric() override { cudaFree(d_d); cudaFree(d_i); }
  void build_tables(const PdesConfig& c, const double* Q, const double* w, const double* interp, const int64_t* perm,
                    const int64_t* nbrperm, const double* wface, int base) override {
    const int nn = c.nn, nfn = c.nfn, ss = c.sparse_face ? 1 : c.ss, NF = DIM + 1, nor = c.norient;
    split = c.volume_integral_type == 2;
    flux_id = c.flux_id;
    tab.nn = nn; tab.nfn = nfn; tab.ss = ss; tab.nor = nor; tab.sparse = c.sparse_face ? 1 : 0;
    const int nperm = c.sparse_face ? nfn : ss;
    o_Qt = 0; o_RfN = o_Qt + (size_t)DIM * nn * nn; o_interp = o_RfN + (size_t)NF * nfn * nn;
    o_wface = o_interp + (size_t)ss * nfn; o_S2 = o_wface + nfn;
    hd.assign(o_S2 + (size_t)DIM * nn * nn, 0.0);
    o_perm = 0; o_nbr = o_perm + (size_t)NF * nperm; o_inv = o_nbr + (size_t)nor * nfn;
    hi.assign(o_inv + (size_t)nn * NF, -1);
    for (int d = 0; d < DIM; ++d)
      for (int j = 0; j < nn; ++j)
        for (int i = 0; i < nn; ++i) {
          hd[o_Qt + ((size_t)d * nn + j) * nn + i] = Q[j + (size_t)nn * (i + (size_t)nn * d)];
          hd[o_S2 + ((size_t)d * nn + j) * nn + i] = Q[j + (size_t)nn * (i + (size_t)nn * d)] - Q[i + (size_t)nn * (j + (size_t)nn * d)];
        }
    for (int f = 0; f < NF; ++f)
      for (int j = 0; j < nperm; ++j) hi[o_perm + (size_t)f * nperm + j] = (int32_t)(perm[j + (int64_t)nperm * f] - base);
    for (int o = 0; o < nor; ++o)
      for (int i = 0; i < nfn; ++i) hi[o_nbr + (size_t)o * nfn + i] = (int32_t)(nbrperm[i + (int64_t)nfn * o] - base);
    for (int i = 0; i < nfn; ++i) hd[o_wface + i] = wface[i];
    if (c.sparse_face) {
      for (int f = 0; f < NF; ++f)
        for (int i = 0; i < nfn; ++i) {
          const int n = hi[o_perm + (size_t)f * nfn + i];
          for (int u = 0; u < NF; ++u)
            if (hi[o_inv + (size_t)n * NF + u] < 0) { hi[o_inv + (size_t)n * NF + u] = f * nfn + i; break; }
        }
    } else {
      for (int j = 0; j < ss; ++j)
        for (int i = 0; i < nfn; ++i) hd[o_interp + (size_t)j * nfn + i] = interp[j + (size_t)ss * i];
      for (int f = 0; f < NF; ++f)
        for (int i = 0; i < nfn; ++i)
          for (int j = 0; j < ss; ++j)
            hd[o_RfN + ((size_t)f * nfn + i) * nn + hi[o_perm + (size_t)f * ss + j]] += interp[j + (size_t)ss * i];
    }
    uploaded = false;
    (void)w;
  }
  cudaError_t prepare() override {
    if (uploaded) return cudaSuccess;
    cudaFree(d_d); cudaFree(d_i); d_d = nullptr; d_i = nullptr;
    cudaError_t e = cudaMalloc((void**)&d_d, sizeof(double) * hd.size());
    if (e != cudaSuccess) return e;
    if ((e = cudaMalloc((void**)&d_i, sizeof(int32_t) * hi.size())) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d_d, hd.data(), sizeof(double) * hd.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    if ((e = cudaMemcpy(d_i, hi.data(), sizeof(int32_t) * hi.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
    tab.Qt = d_d + o_Qt; tab.RfN = d_d + o_RfN; tab.interp = d_d + o_interp; tab.wface = d_d + o_wface; tab.S2 = d_d + o_S2;
    tab.perm = d_i + o_perm; tab.nbrperm = d_i + o_nbr; tab.inv = d_i + o_inv;
    uploaded = true;
    return cudaSuccess;
  }
  int64_t grid_for(int64_t nelems) const override { return (nelems * tab.nn + 127) / 128; }
  int tile_elems() const override { return 1; }
  int resident_element_ctas() override { return 0; }
  int resident_face_ctas() override { return 0; }
  cudaError_t launch_faces(const FaceArgs& a, cudaStream_t s) override {
    if (a.ng <= 0) return cudaSuccess;
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    const unsigned nb = (unsigned)((a.ng * tab.nfn + 127) / 128);
    if (tab.sparse) k_gen_face_sparse<DIM, double><<<nb, 128, 0, s>>>(tab, a, flux_id, nullptr, nullptr);
    else k_gen_face<DIM, double><<<nb, 128, 0, s>>>(tab, a, nullptr, nullptr);
    return cudaGetLastError();
  }
  cudaError_t launch_elements(const ElemArgs& a, int mode, cudaStream_t s) override {
    if (a.nE <= a.e_begin) return cudaSuccess;
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    const unsigned nb = (unsigned)grid_for(a.nE - a.e_begin);
    if (split) {
      if (mode == EPI_RES) k_gen_element_split<DIM, EPI_RES><<<nb, 128, 0, s>>>(tab, a);
      else k_gen_element_split<DIM, EPI_RK><<<nb, 128, 0, s>>>(tab, a);
    } else {
      if (mode == EPI_RES) k_gen_element<DIM, EPI_RES><<<nb, 128, 0, s>>>(tab, a);
      else k_gen_element<DIM, EPI_RK><<<nb, 128, 0, s>>>(tab, a);
    }
    return cudaGetLastError();
  }
  cudaError_t launch_pack(const double* q, const int32_t* sh_el, const uint8_t* sh_face, int64_t nS, double* q_send,
                          double* const* face_dst, const Ctl* ctl, cudaStream_t s) override {
    if (nS <= 0) return cudaSuccess;
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    const int64_t n = nS * tab.nfn * (DIM + 2);
    k_gen_pack<DIM><<<(unsigned)((n + 255) / 256), 256, 0, s>>>(tab, q, sh_el, sh_face, nS, q_send, face_dst, ctl);
    return cudaGetLastError();
  }
  cudaError_t launch_jvp(const FaceArgs& fa, const ElemArgs& a, const double* v, double* out, cudaStream_t s) override {
    if (tab.sparse != (split ? 1 : 0)) return cudaErrorNotSupported;
    { cudaError_t e = prepare(); if (e != cudaSuccess) return e; }
    const int64_t nfn = fa.ng * tab.nfn, nen = a.nE * tab.nn;
    if (split) {
      // entropy-stable configuration: IR / IRSLF / Roe interface flux and the split-form volume terms on dual numbers
      if (nfn > 0) k_gen_face_sparse<DIM, Dual><<<(unsigned)((nfn + 127) / 128), 128, 0, s>>>(tab, fa, flux_id, v, fa.v_recv);
      k_gen_jvp_element_split<DIM><<<(unsigned)((nen + 127) / 128), 128, 0, s>>>(tab, a, v, out);
      return cudaGetLastError();
    }
    if (nfn > 0) k_gen_face<DIM, Dual><<<(unsigned)((nfn + 127) / 128), 128, 0, s>>>(tab, fa, v, fa.v_recv);
    k_gen_jvp_element<DIM><<<(unsigned)((nen + 127) / 128), 128, 0, s>>>(tab, a, v, out);
    return cudaGetLastError();
  }
};

Ops* make_generic_ops(const PdesConfig& c) {
  if (c.nn < 1 || c.nn > 255 || c.nfn < 1 || c.nfn > 255) return nullptr;      // (the error location keeps 8 bits for the node)
  if (c.sparse_face) {
    if (c.volume_integral_type != 2 || c.volume_flux_id != PDES_FLUX_IR || c.face_integral_type != 1) return nullptr;
    if (c.flux_id != PDES_FLUX_ROE && c.flux_id != PDES_FLUX_IR && c.flux_id != PDES_FLUX_IRSLF) return nullptr;
  } else {
    if (c.face_integral_type != 1 || c.volume_integral_type != 1 || c.flux_id != PDES_FLUX_ROE) return nullptr;
  }
  if (c.dim == 2) return new OpsGeneric<2>();
  if (c.dim == 3) return new OpsGeneric<3>();
  return nullptr;
}

Ops* make_ops_tuned(const PdesConfig& c);
Ops* make_ops(const PdesConfig& c) {
  // PDES_GENERIC=1: the size-generic kernels also for the operators that have tuned instantiations (tests)
  Ops* o = env_int("PDES_GENERIC", 0) ? nullptr : make_ops_tuned(c);
  return o ? o : make_generic_ops(c);
}

Ops* make_ops_tuned(const PdesConfig& c) {
#ifdef PDES_LEAN
  // development build (make EXTRA=-DPDES_LEAN): only the kernels of the headline workload, a fraction of the build time
  if (!c.sparse_face && c.face_integral_type == 1 && c.volume_integral_type == 1 && c.flux_id == PDES_FLUX_ROE &&
      c.dim == 3 && c.nn == 11 && c.nfn == 6)
#ifndef PDES_DEV_E
#define PDES_DEV_E 32
#define PDES_DEV_MINB_E 4
#define PDES_DEV_FT 16
#define PDES_DEV_MINB_F 8
#endif
    return new OpsImpl<3, 11, 6, PDES_DEV_E, PDES_DEV_MINB_E, PDES_DEV_FT, PDES_DEV_MINB_F>();
#ifndef PDES_DEV_ES_E
#define PDES_DEV_ES_E 8
#endif
  if (c.sparse_face && c.dim == 2 && c.nn == 12 && c.nfn == 4 && c.volume_integral_type == 2) return new OpsImplS<2, 12, 4, PDES_DEV_ES_E>();
  return nullptr;
#else
  if (c.sparse_face) {
    // entropy-stable configuration: diag-E operator, split-form IR volume flux, Roe / IR / IRSLF interface flux
    if (c.volume_integral_type != 2 || c.volume_flux_id != PDES_FLUX_IR) return nullptr;
    if (c.face_integral_type != 1) return nullptr;      // diagonal-E operators keep face_integral_type 1 (read_input.jl:742-755)
    if (c.flux_id != PDES_FLUX_ROE && c.flux_id != PDES_FLUX_IR && c.flux_id != PDES_FLUX_IRSLF) return nullptr;
    // 8 elements = 96 threads per CTA (C2, k_element_split_r: 13.9 ms per RK4 step; 16: 14.3, 24: 15.2, 32: 15.3)
    if (c.dim == 2 && c.nn == 12 && c.nfn == 4) return new OpsImplS<2, 12, 4, 8>();
    return nullptr;
  }
  if (c.face_integral_type == 2) {
    // entropy-stable scheme on SBP-Omega operators: split-form IR volume flux + face-element integrals with the IR flux
    if (c.volume_integral_type != 2 || c.volume_flux_id != PDES_FLUX_IR || c.flux_id != PDES_FLUX_IR) return nullptr;
    if (c.face_element_id < PDES_FEI_EC || c.face_element_id > PDES_FEI_ESLW2) return nullptr;
    if (c.dim == 2 && c.nn == 3 && c.nfn == 2) return new OpsImplE<2, 3, 2, 32>();
    if (c.dim == 2 && c.nn == 6 && c.nfn == 3) return new OpsImplE<2, 6, 3, 32>();
    if (c.dim == 3 && c.nn == 4 && c.nfn == 3) return new OpsImplE<3, 4, 3, 32>();
    if (c.dim == 3 && c.nn == 11 && c.nfn == 6) return new OpsImplE<3, 11, 6, 8>();
    return nullptr;
  }
  if (c.volume_integral_type != 1 || c.flux_id != PDES_FLUX_ROE) return nullptr;
  const int variant = env_int("PDES_VARIANT", 0);   // tuning knob (tools/bench_variants.sh)
  if (c.dim == 2 && c.nn == 3 && c.nfn == 2) return new OpsImpl<2, 3, 2, 64, 2, 64, 2, 4, 32, 3>();
  if (c.dim == 2 && c.nn == 6 && c.nfn == 3) return new OpsImpl<2, 6, 3, 32, 2, 32, 2, 4, 16, 3>();
  if (c.dim == 3 && c.nn == 4 && c.nfn == 3) return new OpsImpl<3, 4, 3, 32, 2, 32, 2, 4, 16, 4>();
  if (c.dim == 3 && c.nn == 11 && c.nfn == 6) {
    switch (variant) {
      case 1: return new OpsImpl<3, 11, 6, 32, 3, 16, 8>();
      case 2: return new OpsImpl<3, 11, 6, 32, 5, 16, 8>();
      case 3: return new OpsImpl<3, 11, 6, 64, 1, 16, 8>();
      case 4: return new OpsImpl<3, 11, 6, 64, 2, 16, 8>();
      case 5: return new OpsImpl<3, 11, 6, 38, 3, 16, 8>();
      case 6: return new OpsImpl<3, 11, 6, 24, 5, 16, 8>();
      case 9: return new OpsImpl<3, 11, 6, 32, 4, 16, 8, 5>();
      case 10: return new OpsImpl<3, 11, 6, 32, 4, 16, 8, 6>();
      case 11: return new OpsImpl<3, 11, 6, 32, 4, 16, 8, 3>();
      case 7: return new OpsImpl<3, 11, 6, 32, 4, 16, 4>();
      case 8: return new OpsImpl<3, 11, 6, 32, 4, 16, 5>();
      default: return new OpsImpl<3, 11, 6, 32, 4, 16, 8>();
    }
  }
  return nullptr;
#endif
}

// one-time uploads: stream-ordered with the kernels of the (non-blocking) compute stream, then synchronised so
// that the host buffer may be released
template <typename T>
cudaError_t dev_upload(cudaStream_t st, T** dst, const T* src, size_t n) {
  if (*dst) { cudaFree(*dst); *dst = nullptr; }
  if (n == 0) n = 1, src = nullptr;
  cudaError_t e = cudaMalloc((void**)dst, n * sizeof(T));
  if (e != cudaSuccess) return e;
  if (src) e = cudaMemcpyAsync(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice, st);
  else e = cudaMemsetAsync(*dst, 0, n * sizeof(T), st);
  if (e != cudaSuccess) return e;
  return cudaStreamSynchronize(st);
}

}  // namespace

struct PdesCtx {
  PdesConfig cfg;
  int nd = 0, nf = 0;
  int64_t ndof = 0;
  std::string err;
  int64_t err_element = -1, err_node = -1;
  std::unique_ptr<Ops> ops;
  bool have_op = false, have_mesh = false, finalized = false;
  cudaStream_t stream = nullptr, comm_stream = nullptr, face_stream = nullptr;
  // the face-flux and element kernels of one evaluation are cut into chunks that run concurrently on two streams
  // (k_face_flux is issue-bound, k_element_rk is HBM-bound): chunk c of the elements needs face chunks <= c
  static const int MAXC = 16;
  int nchunks = 1;
  int64_t chunk_e[MAXC + 1] = {0}, chunk_g[MAXC + 1] = {0};
  cudaEvent_t ev_face[MAXC] = {nullptr}, ev_elem = nullptr;
  cudaEvent_t ev_packed = nullptr, ev_recv = nullptr, ev_norm = nullptr, ev_t0 = nullptr, ev_t1 = nullptr, ev_q = nullptr;
  bool comm_overlap = true;     // shared-face branch (pack, exchange, flux) on the communication stream
  // peer-to-peer halo (default with a communicator): the packed face states go straight into the neighbour's receive buffer
  // (CUDA IPC mapping, copy engine over NVLink) and a flag; no NCCL kernel competes with the interior faces for SMs
  int p2p = -1;                 // -1: not set up yet, 0: NCCL send/recv, 1: peer-to-peer
  double* halo_buf = nullptr;   // [2][nsend] receive buffers (ping-pong by evaluation parity) | flags[npeers]
  unsigned* halo_flags = nullptr;
  size_t halo_nsend = 0;
  uint32_t halo_epoch = 0;
  bool halo_fused = false;      // send pass, flags and receive wait inside the one face launch (HaloArgs); epoch on the device
  unsigned* halo_ctr = nullptr; // device: {evaluations completed, pack tiles done, norms committed}
  // calcNorm's all-reduce through peer memory (NormX): every rank's buffer mapped, slot rings behind the flags
  std::vector<void*> rank_mapped;     // [nranks] imported mappings (nullptr for this rank)
  double** d_norm_slots = nullptr;    // device: [nranks] slot ring of every rank
  bool norm_p2p = false;
  bool norm_pending = false;          // a k_norm_reduce whose k_norm_commit has been deferred to the end of the step
  double pend_tol = -1.0; int pend_pseudo = 0;
  double* q_recv_eval = nullptr;
  struct PeerMap { double* base = nullptr; int64_t remote_nsend = 0, remote_off = 0; unsigned* flag = nullptr; };
  std::vector<PeerMap> pmap;
  unsigned** d_flag_ptrs = nullptr;   // device array of the neighbours' flag slots
  double** d_face_dst = nullptr;      // [2][nS] per shared face: its slot in the neighbour's receive buffer
  // state
  double* qbuf[3] = {nullptr, nullptr, nullptr};
  int cur = 0;
  double *ksum = nullptr, *res = nullptr;
  // mesh
  double *dxidx = nullptr, *minv = nullptr, *mass = nullptr, *srcw = nullptr, *coords_bndry = nullptr, *w_dev = nullptr;
  double* Q_dev = nullptr;          // sbp.Q [nn,nn,dim] as uploaded (calcVorticity differentiates with it)
  FaceRec* faces = nullptr;
  double *nrm_all = nullptr, *fluxe = nullptr, *srcm = nullptr;
  std::vector<EFace> h_efaces;      // interior + boundary part (shared faces added by finalize)
  std::vector<FaceRec> h_faces;
  std::vector<double> h_nrm;        // nrm_face | nrm_bndry
  bool dx_compact = false, nrm_compact = false;   // node-independent metrics detected at upload
  int prefetch_ahead = 0, prefetch_ahead_faces = 0;
  // k_fused schedule: plan A = interior + boundary faces with the element tiles that touch no shared face (all tiles
  // on one GPU); plan B = the shared faces and the remaining tiles, launched once the receive has completed
  struct FusedPlan {
    int32_t *tile_list = nullptr, *need = nullptr;
    int32_t n_tiles = 0, n_groups = 0, lag = 0;
    int64_t g0 = 0, ng = 0;
  } plan[2];
  bool fused = false, has_ext_bc = false;
  int prefetch_ahead_groups = 0, discard_records = 0, acquire_fence = 1;
  int rk4_nosum = 0;       // rk4 stages without the running sum of the k's (epilogue_tile scheme 2)
  int stagger_ns = 0;
  unsigned* tile_ctr = nullptr;   // k_element_tma dynamic tile deal (PDES_TMA_DYN builds)
  // pdes_eval_residual_host: evaluation pipelined with the upload of q and the download of res (HC chunks)
#ifndef PDES_HOST_CHUNKS
#define PDES_HOST_CHUNKS 8
#endif
  static const int HC = PDES_HOST_CHUNKS;
  int host_chunks = 0;            // 0: not possible for this context (partitioned mesh, too small, other schedule)
  int64_t hc_e[HC + 1] = {0}, hc_g[HC + 1] = {0};
  int hc_updep[HC] = {0};
  cudaStream_t up_stream = nullptr, down_stream = nullptr;
  cudaEvent_t ev_up[HC] = {nullptr}, ev_el[HC] = {nullptr}, ev_idle = nullptr;
  int reverse_elems = 0, discard_split = 0;   // split kernels: element tiles swept last-to-first; consumed records dropped from L2
  // chunk pipeline (PDES_PIPE = number of chunks): F0 F1 E0 F2 E1 ... with programmatic dependent launches
  bool pipe = false;
  int pipe_lag = 1, pipe_discard = 1, pipe_dbg = 0, chunk_dep[MAXC] = {0};
  uint32_t pipe_epoch = 0;
  unsigned* pipe_ctr = nullptr;     // [4][MAXC]: arriveF | doneF | arriveE | doneE
  Sched* sched = nullptr;
  unsigned* flags = nullptr;
  double* diag_buf = nullptr;
  int diag_B = 0;
  // Newton-Krylov workspace (allocated on first use): basis V[(restart+1)][ndof], work vectors, reduction scratch
  struct Krylov {
    int restart = 0, nblk = 0;
    double *V = nullptr, *w = nullptr, *b = nullptr, *x = nullptr, *partials = nullptr, *hdev = nullptr;
    double* hhost = nullptr;     // pinned: [3*(restart+2)]
    // Hessenberg matrix, Givens rotations, right-hand side and solver state on the device (k_gmres_update)
    double *gH = nullptr, *gcs = nullptr, *gsn = nullptr, *gg = nullptr;
    GmresState* gstate = nullptr;
    GmresState* hstate = nullptr;   // pinned
    // one CUDA graph per Krylov iteration j of a restart cycle (its ~12 launches differ only in j): captured on first use,
    // replayed by every later cycle and solve; key = (restart, preconditioner, state buffer)
    std::vector<cudaGraphExec_t> itg;
    int itg_pc = -1, itg_cur = -1;
    // element-block Jacobi right preconditioner (pdes_set_krylov_pc)
    int pc_type = 0, ncolours = 0;
    bool pc_ready = false;
    double *pc_blocks = nullptr, *pcz = nullptr, *pcu = nullptr;
    int32_t* colour = nullptr;
  } kry;
  // CUDA graphs of one RK4 step, one per state-buffer rotation; key = (h, norm?, res_tol, pseudo_time)
  cudaGraphExec_t step_graph[3] = {nullptr, nullptr, nullptr};
  double g_h = -1.0, g_tol = 0.0;
  bool g_norm = false, no_graph = false;
  bool capturing = false, norm_branch_open = false;   // graph capture of a multi-GPU step: the all-reduce branch joins at the end
  int g_pseudo = 0, g_launches = 0;
  std::vector<double> h_w;
  // partition
  std::vector<Peer> peers;
  int64_t nS = 0;
  double *q_send = nullptr, *q_recv = nullptr;
  double *v_send = nullptr, *v_recv = nullptr;     // J*v on a partitioned mesh: shared-face values of the direction
  // element-data halo of the type-2 face integrals
  bool elem_halo = false;
  int64_t n_send_el = 0, n_recv_el = 0;
  int32_t* el_send_list = nullptr;
  double *qel_send = nullptr, *qel_recv = nullptr;
  int32_t* sh_el = nullptr;
  uint8_t* sh_face = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
  // control
  Ctl* ctl = nullptr;
  Ctl* h_ctl = nullptr;   // pinned
  double *norm_partials = nullptr, *norm_sq = nullptr, *norms_dev = nullptr;
  int64_t norms_cap = 0;
  int64_t launches = 0, n_evals = 0;
  PdesTimings tm{};
};

namespace {

void set_err(PdesCtx* ctx, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (ctx) ctx->err = buf;
}

int usage(PdesCtx* ctx, const char* msg) {
  set_err(ctx, "%s", msg);
  return PDES_ERR_USAGE;
}

PhysPar phys_of(const PdesConfig& c) {
  PhysPar p;
  p.gamma = c.gamma; p.R = c.R; p.Ma = c.Ma; p.aoa = c.aoa; p.rho_free = c.rho_free; p.E_free = c.E_free;
  p.check_density = c.check_density; p.check_pressure = c.check_pressure;
  return p;
}

int reset_ctl(PdesCtx* ctx) {
  Ctl z;
  z.stop = 0; z.err_code = 0; z.err_loc = ~0ull; z.converged_step = -1; z.norm_count = 0; z.kry_done = 0;
  *ctx->h_ctl = z;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->ctl, ctx->h_ctl, sizeof(Ctl), cudaMemcpyHostToDevice, ctx->stream));
  if (ctx->pipe_ctr) {
    // an aborted evaluation leaves arrival counts and unpublished chunks behind: every chunk "complete at the current epoch"
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned h[4 * PdesCtx::MAXC];
    for (int i = 0; i < PdesCtx::MAXC; ++i) {
      h[i] = 0u; h[2 * PdesCtx::MAXC + i] = 0u;
      h[PdesCtx::MAXC + i] = ctx->pipe_epoch; h[3 * PdesCtx::MAXC + i] = ctx->pipe_epoch;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pipe_ctr, h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return PDES_OK;
}

// reads the control block back (synchronises the compute stream) and converts a device-side physics error into
// the reference's exception semantics (euler.jl:552-556, 598-603)
int fetch_ctl(PdesCtx* ctx) {
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_ctl, ctx->ctl, sizeof(Ctl), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->comm_stream) CUDA_TRY(ctx, cudaStreamSynchronize(ctx->comm_stream));
  if (ctx->h_ctl->err_code == 4) {
    set_err(ctx, "halo exchange: a neighbour's face states did not arrive within 120 s");
    return PDES_ERR_COMM;
  }
  if (ctx->h_ctl->err_code == 5) {
    set_err(ctx, "halo exchange: a neighbouring rank stopped (physics error there); re-establish the communicator "
                 "(pdes_set_comm) on every rank before the next evaluation");
    return PDES_ERR_COMM;
  }
  if (ctx->h_ctl->err_code == 3) {
    set_err(ctx, "k_fused: an element tile waited for face groups that never completed (scheduler time-out)");
    return PDES_ERR_CUDA;
  }
  if (ctx->h_ctl->err_code) {
    unsigned long long key = ctx->h_ctl->err_loc;
    int code = (int)(key >> 62) + 1;
    unsigned long long loc = key & ((1ull << 62) - 1);
    ctx->err_element = (int64_t)(loc >> 8) + ctx->cfg.index_base;
    ctx->err_node = (int64_t)(loc & 0xff) + ctx->cfg.index_base;
    set_err(ctx, "%s detected at element %lld, node %lld", code == 1 ? "Negative density" : "Negative pressure",
            (long long)ctx->err_element, (long long)ctx->err_node);
    return code == 1 ? PDES_ERR_NEG_DENSITY : PDES_ERR_NEG_PRESSURE;
  }
  return PDES_OK;
}

// builds the device copies that depend on both the mesh and the peer lists
int finalize(PdesCtx* ctx) {
  if (ctx->finalized) return PDES_OK;
  if (!ctx->have_op || !ctx->have_mesh) return usage(ctx, "pdes_set_operator and pdes_set_mesh must be called first");
  const PdesConfig& c = ctx->cfg;
  const int NF = ctx->nf, base = c.index_base;
  const size_t per_nrm = (size_t)c.nfn * c.dim;
  // shared faces are appended to the face list after the interfaces and the boundary faces
  ctx->nS = 0;
  for (auto& p : ctx->peers) { p.offset = ctx->nS; ctx->nS += p.nfaces; }
  ctx->elem_halo = c.face_integral_type == 2 && !ctx->peers.empty();
  ctx->n_send_el = ctx->n_recv_el = 0;
  std::vector<int32_t> el_send;
  if (ctx->elem_halo) {
    for (auto& p : ctx->peers) {
      if (!p.have_els) return usage(ctx, "face_integral_type 2 on a partitioned mesh: pdes_set_peer_elements was not called for every peer");
      p.el_send_off = ctx->n_send_el; p.el_recv_off = ctx->n_recv_el;
      ctx->n_send_el += (int64_t)p.send_els.size(); ctx->n_recv_el += p.n_recv_el;
      for (int64_t e : p.send_els) {
        if (e - base < 0 || e - base >= c.nE) return usage(ctx, "pdes_set_peer_elements: local element out of range");
        el_send.push_back((int32_t)(e - base));
      }
    }
  }
  const int64_t nG = c.nF + c.nB + ctx->nS;
  std::vector<int32_t> sh_el(ctx->nS);
  std::vector<uint8_t> sh_face(ctx->nS);
  std::vector<EFace> ef = ctx->h_efaces;
  std::vector<FaceRec> faces = ctx->h_faces;
  std::vector<double> nrm = ctx->h_nrm;
  faces.resize(nG);
  nrm.resize((size_t)nG * per_nrm);
  for (auto& p : ctx->peers) {
    if (p.rank < 0) return usage(ctx, "pdes_set_peer was not called for every peer");
    for (int64_t j = 0; j < p.nfaces; ++j) {
      int64_t el = (int64_t)p.ifaces[j].elementL - base;
      int f = (int)p.ifaces[j].faceL - base, o = (int)p.ifaces[j].orient - base;
      if (el < 0 || el >= c.nE || f < 0 || f >= NF || o < 0 || o >= c.norient)
        return usage(ctx, "shared interface out of range");
      if ((int64_t)p.bndries_local[j].element - base != el || (int)p.bndries_local[j].face - base != f)
        return usage(ctx, "bndries_local and shared_interfaces disagree");
      EFace& r = ef[el * NF + f];
      if (r.gface >= 0) return usage(ctx, "shared face already claimed by an interface or boundary face");
      const int64_t g = c.nF + c.nB + p.offset + j;
      r.gface = (int32_t)g; r.right = 0; r.orient = (uint8_t)o;
      FaceRec& fr = faces[g];
      memset(&fr, 0, sizeof(fr));
      fr.elL = (int32_t)el; fr.elR = -1; fr.fL = (uint8_t)f; fr.orient = (uint8_t)o; fr.kind = FK_SHARED;
      fr.fR = (uint8_t)((int)p.ifaces[j].faceR - base);
      fr.aux = (int32_t)(p.offset + j);
      if (ctx->elem_halo) {
        // the neighbour's element behind this face: slot in the element receive buffer
        const int64_t rel = (int64_t)p.ifaces[j].elementR - p.shared_el_offset;
        if (rel < 0 || rel >= p.n_recv_el || fr.fR >= NF) return usage(ctx, "shared interface: remote element / face out of range");
        fr.elR = (int32_t)(p.el_recv_off + rel);
      }
      sh_el[p.offset + j] = (int32_t)el;
      sh_face[p.offset + j] = (uint8_t)f;
    }
    if (p.nfaces)
      memcpy(nrm.data() + (size_t)(c.nF + c.nB + p.offset) * per_nrm, p.nrm.data(), sizeof(double) * p.nrm.size());
  }
  {
    // order the interior + boundary faces by the lowest element they touch: the faces an element chunk needs are then a
    // prefix of the list, which lets element chunk c start while face chunk c+1 is still running
    const int64_t nIB = c.nF + c.nB;
    std::vector<int64_t> order(nIB);
    for (int64_t g = 0; g < nIB; ++g) order[g] = g;
    auto key = [&](int64_t g) {
      const FaceRec& r = faces[g];
      return (r.kind == FK_INTERIOR && r.elR < r.elL) ? r.elR : r.elL;
    };
    std::stable_sort(order.begin(), order.end(), [&](int64_t x, int64_t y) { return key(x) < key(y); });
    std::vector<FaceRec> fs(faces);
    std::vector<double> ns(nrm);
    for (int64_t g = 0; g < nIB; ++g) {
      faces[g] = fs[order[g]];
      if (faces[g].kind == FK_BOUNDARY) faces[g].elR = (int32_t)(order[g] - c.nF);
      memcpy(nrm.data() + (size_t)g * per_nrm, ns.data() + (size_t)order[g] * per_nrm, sizeof(double) * per_nrm);
    }
    const int tile = ctx->ops->tile_elems();
    int nc = env_int("PDES_CHUNKS", 0);
    if (nc <= 0) nc = 1;   // measured on C3: more chunks only add tail waves
    const int want_pipe = env_int("PDES_PIPE", 0);
    const bool can_pipe = want_pipe > 1 && ctx->ops->pipe_face_tile() > 0 && ctx->nS == 0 && !ctx->has_ext_bc &&
                          env_int("PDES_FUSED", 0) == 0 && env_int("PDES_ELEM_W", 0) == 0 && env_int("PDES_FACE_TMA", 0) == 0;
    if (can_pipe) nc = want_pipe;
    if (nc > PdesCtx::MAXC) nc = PdesCtx::MAXC;
    const int64_t ntiles = (c.nE + tile - 1) / tile;
    if (nc > ntiles) nc = (int)ntiles;
    ctx->nchunks = nc;
    int64_t g = 0;
    for (int k = 0; k <= nc; ++k) {
      int64_t e = k == nc ? c.nE : (ntiles * k / nc) * tile;
      ctx->chunk_e[k] = e;
      while (g < nIB && key(g) < e) ++g;      // faces[] is sorted: key(g) uses the permuted list
      ctx->chunk_g[k] = k == nc ? nIB : g;
    }
    ctx->chunk_g[0] = 0;
    ctx->pipe = can_pipe && nc > 1;
    if (ctx->pipe) {
      // chunk_dep[k]: the highest element chunk touched by the faces of face chunk k (q they gather, records they write)
      for (int k = 0; k < nc; ++k) {
        int64_t emax = 0;
        for (int64_t gg = ctx->chunk_g[k]; gg < ctx->chunk_g[k + 1]; ++gg) {
          emax = std::max<int64_t>(emax, faces[gg].elL);
          if (faces[gg].kind == FK_INTERIOR) emax = std::max<int64_t>(emax, faces[gg].elR);
        }
        int d = k;
        while (d + 1 < nc && ctx->chunk_e[d + 1] <= emax) ++d;
        ctx->chunk_dep[k] = d;
      }
      ctx->pipe_lag = std::max(0, env_int("PDES_PIPE_LAG", 1));
      ctx->pipe_discard = env_int("PDES_PIPE_DISCARD", 1);
      ctx->pipe_dbg = env_int("PDES_PIPE_DBG", 0) & 6;      // measurement only: 2 = no acquire fence, 4 = no release fence
      if (!ctx->pipe_ctr) CUDA_TRY(ctx, cudaMalloc((void**)&ctx->pipe_ctr, sizeof(unsigned) * 4 * PdesCtx::MAXC));
      CUDA_TRY(ctx, cudaMemsetAsync(ctx->pipe_ctr, 0, sizeof(unsigned) * 4 * PdesCtx::MAXC, ctx->stream));
      ctx->pipe_epoch = 0;
    }

    // k_fused schedule (see residual_kernels.cuh): the faces an element tile integrates are a prefix of the sorted list
    const int FPG = ctx->ops->fused_group_faces();
    ctx->fused = FPG > 0 && nc == 1 && !ctx->has_ext_bc && env_int("PDES_FUSED", 0) != 0 && env_int("PDES_ELEM_W", 0) == 0;
    if (ctx->fused) {
      const int E = ctx->ops->fused_tile_elems();
      const int64_t nt = (c.nE + E - 1) / E;
      std::vector<char> touches(nt, 0);
      for (int64_t j = 0; j < ctx->nS; ++j) touches[sh_el[j] / E] = 1;
      std::vector<int32_t> list[2], need[2];
      int64_t gp = 0;
      for (int64_t t = 0; t < nt; ++t) {
        const int64_t e_end = std::min<int64_t>((t + 1) * E, c.nE);
        while (gp < nIB && key(gp) < e_end) ++gp;
        const int k = touches[t] ? 1 : 0;
        list[k].push_back((int32_t)t);
        need[k].push_back((int32_t)((gp + FPG - 1) / FPG));
      }
      const int32_t groupsB = (int32_t)((ctx->nS + FPG - 1) / FPG);
      for (auto& v : need[1]) v = groupsB;
      for (int k = 0; k < 2; ++k) {
        PdesCtx::FusedPlan& pl = ctx->plan[k];
        pl.n_tiles = (int32_t)list[k].size();
        pl.g0 = k == 0 ? 0 : nIB;
        pl.ng = k == 0 ? nIB : ctx->nS;
        pl.n_groups = (int32_t)((pl.ng + FPG - 1) / FPG);
        pl.lag = 0;
        for (size_t i = 0; i < need[k].size(); ++i) pl.lag = std::max<int32_t>(pl.lag, need[k][i] - (int32_t)i);
        // slack: a tile should wait for groups that are (almost surely) complete, i.e. older than one wave of CTAs
        if (k == 0) pl.lag += env_int("PDES_LAG_SLACK", ctx->ops->resident_fused_ctas() + 64);
        if (pl.tile_list) { cudaFree(pl.tile_list); pl.tile_list = nullptr; }
        if (ctx->nS > 0) CUDA_TRY(ctx, dev_upload(ctx->stream, &pl.tile_list, list[k].data(), list[k].size()));
        CUDA_TRY(ctx, dev_upload(ctx->stream, &pl.need, need[k].data(), need[k].size()));
      }
      const size_t nflags = (size_t)std::max(ctx->plan[0].n_groups, ctx->plan[1].n_groups) + 1;
      CUDA_TRY(ctx, dev_upload<unsigned>(ctx->stream, &ctx->flags, nullptr, nflags));
      Sched s0{0u, 0u, 0u, 1u};
      CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->sched, &s0, 1));
    }
  }
  {
    // chunks of pdes_eval_residual_host: HC element ranges of whole tiles; the faces whose lowest element lies in range k
    // (the face list is sorted by it) form face chunk k; hc_updep[k] = the last element chunk those faces read
    ctx->host_chunks = 0;
    const int tile = ctx->ops->tile_elems();
    const int64_t ntiles = (c.nE + tile - 1) / tile, nIB = c.nF + c.nB;
    if (ctx->nS == 0 && ctx->nchunks == 1 && !ctx->fused && !ctx->pipe && ntiles >= 64 * PdesCtx::HC) {
      auto key = [&](int64_t g) {
        const FaceRec& r = faces[g];
        return (r.kind == FK_INTERIOR && r.elR < r.elL) ? r.elR : r.elL;
      };
      int64_t g = 0;
      for (int k = 0; k <= PdesCtx::HC; ++k) {
        const int64_t e = k == PdesCtx::HC ? c.nE : (ntiles * k / PdesCtx::HC) * tile;
        ctx->hc_e[k] = e;
        while (g < nIB && key(g) < e) ++g;
        ctx->hc_g[k] = k == PdesCtx::HC ? nIB : g;
      }
      ctx->hc_g[0] = 0;
      for (int k = 0; k < PdesCtx::HC; ++k) {
        int64_t emax = 0;
        for (int64_t gg = ctx->hc_g[k]; gg < ctx->hc_g[k + 1]; ++gg) {
          emax = std::max<int64_t>(emax, faces[gg].elL);
          if (faces[gg].kind == FK_INTERIOR) emax = std::max<int64_t>(emax, faces[gg].elR);
        }
        int d = k;
        while (d + 1 < PdesCtx::HC && ctx->hc_e[d + 1] <= emax) ++d;
        ctx->hc_updep[k] = d;
      }
      ctx->host_chunks = PdesCtx::HC;
    }
  }
  for (int64_t e = 0; e < c.nE; ++e)
    for (int f = 0; f < NF; ++f)
      if (ef[e * NF + f].gface < 0) {
        set_err(ctx, "face %d of element %lld belongs to no interface, boundary face or shared face", f, (long long)e);
        return PDES_ERR_USAGE;
      }
  CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->faces, faces.data(), faces.size()));
  {
    bool compact = env_int("PDES_NO_COMPACT", 0) == 0;
    for (int64_t g = 0; g < nG && compact; ++g)
      for (int i = 1; i < c.nfn && compact; ++i)
        compact = memcmp(nrm.data() + (size_t)g * per_nrm + (size_t)i * c.dim, nrm.data() + (size_t)g * per_nrm,
                         sizeof(double) * c.dim) == 0;
    ctx->nrm_compact = compact;
    if (compact) {
      std::vector<double> nc((size_t)nG * c.dim);
      for (int64_t g = 0; g < nG; ++g) memcpy(nc.data() + (size_t)g * c.dim, nrm.data() + (size_t)g * per_nrm, sizeof(double) * c.dim);
      CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->nrm_all, nc.data(), nc.size()));
    } else {
      CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->nrm_all, nrm.data(), nrm.size()));
    }
  }
  {
    const size_t rec = ctx->ops->record_doubles() ? (size_t)ctx->ops->record_doubles() : (size_t)c.nfn * ctx->nd;
    CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->fluxe, nullptr, (size_t)c.nE * NF * rec + 2));
  }
  CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->sh_el, sh_el.data(), sh_el.size()));
  CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->sh_face, sh_face.data(), sh_face.size()));
  size_t nsend = (size_t)ctx->nS * c.nfn * ctx->nd;
  CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->q_send, nullptr, nsend));
  CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->q_recv, nullptr, nsend));
  if (ctx->elem_halo) {
    const size_t el_len = (size_t)c.nn * ctx->nd;
    CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->el_send_list, el_send.data(), el_send.size()));
    CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->qel_send, nullptr, (size_t)ctx->n_send_el * el_len));
    CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->qel_recv, nullptr, (size_t)ctx->n_recv_el * el_len));
  }
  CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->norm_partials, nullptr, (size_t)ctx->ops->grid_for(c.nE)));
  for (int i = 0; i < 3; ++i)
    if (ctx->step_graph[i]) { cudaGraphExecDestroy(ctx->step_graph[i]); ctx->step_graph[i] = nullptr; }
  for (cudaGraphExec_t ge : ctx->kry.itg) if (ge) cudaGraphExecDestroy(ge);      // (captured with the old device arrays)
  ctx->kry.itg.clear();
  ctx->no_graph = env_int("PDES_NO_GRAPH", 0) != 0;
  CUDA_TRY(ctx, ctx->ops->prepare());
  ctx->prefetch_ahead = env_int("PDES_PREFETCH_AHEAD", ctx->ops->resident_element_ctas() / 8);   // measured optimum: ~half a wave
  ctx->prefetch_ahead_faces = env_int("PDES_PREFETCH_AHEAD_F", ctx->ops->resident_face_ctas());
  // k_face_flux sweeps the (element-sorted) face list upwards, so when it ends the L2 holds the records and the gathered
  // q of the HIGHEST elements; k_element_rk sweeping downwards starts on those lines and ends on the lowest elements,
  // whose q_next the next stage's k_face_flux asks for first
  ctx->rk4_nosum = env_int("PDES_RK4_NOSUM", 1) != 0 && ctx->ops->staged_epilogue();
  ctx->comm_overlap = env_int("PDES_COMM_INLINE", 0) == 0 && !ctx->fused;
  ctx->reverse_elems = env_int("PDES_REV", 1);
  ctx->stagger_ns = env_int("PDES_STAGGER_NS", 0);
  ctx->discard_split = env_int("PDES_DISCARD_SPLIT", 1);
  if (ctx->fused) {
    const int res = ctx->ops->resident_fused_ctas();
    ctx->prefetch_ahead = env_int("PDES_PREFETCH_AHEAD", res / 8);
    ctx->prefetch_ahead_groups = env_int("PDES_PREFETCH_AHEAD_G", res);
    ctx->discard_records = env_int("PDES_DISCARD", 1);
    ctx->acquire_fence = env_int("PDES_ACQ_FENCE", 1);
  }
  ctx->finalized = true;
  return PDES_OK;
}

void fill_args(PdesCtx* ctx, ElemArgs* a, const double* q) {
  memset(a, 0, sizeof(*a));
  a->q = q; a->dxidx = ctx->dxidx; a->fluxe = ctx->fluxe;
  a->srcw = ctx->cfg.src_id == PDES_SRC_EXP ? ctx->srcw : nullptr;
  a->srcm = ctx->cfg.src_id == PDES_SRC_EXP ? ctx->srcm : nullptr;
  a->minv = ctx->minv; a->mass = ctx->mass; a->nE = ctx->cfg.nE; a->ctl = ctx->ctl; a->ph = phys_of(ctx->cfg);
  a->norm_partials = ctx->norm_partials;
  a->e_begin = 0;
  const int dd = ctx->cfg.dim * ctx->cfg.dim;
  a->dx_el_stride = ctx->dx_compact ? dd : ctx->cfg.nn * dd;
  a->dx_node_stride = ctx->dx_compact ? 0 : dd;
  a->prefetch_ahead = ctx->prefetch_ahead;
  a->reverse = ctx->reverse_elems;
  a->stagger_ns = ctx->stagger_ns;
  a->tile_ctr = ctx->tile_ctr;
  a->discard_records = ctx->pipe ? ctx->pipe_discard : ctx->discard_split;
}

// One-time set-up of the peer-to-peer halo: every rank exports its receive buffer (CUDA IPC), the handles and the layout of
// every rank's shared-face list travel by ncclAllGather, each neighbour's buffer is mapped.  All ranks take the same
// decision (ncclAllReduce of the local outcome): 1 = peer-to-peer, 0 = ncclSend/ncclRecv.
struct HaloRec {
  cudaIpcMemHandle_t handle;
  int64_t nsend;                 // doubles per receive buffer
  int32_t npeers, ok;
  int32_t peer_rank[32];
  int64_t peer_off[32], peer_n[32];   // in faces
};

int setup_p2p(PdesCtx* ctx) {
  ctx->p2p = 0;
  if (!ctx->comm) return PDES_OK;
  const size_t per_face = (size_t)ctx->cfg.nfn * ctx->nd;
  const size_t nsend = (size_t)ctx->nS * per_face;
  const int np = (int)ctx->peers.size();
  HaloRec mine;
  memset(&mine, 0, sizeof(mine));
  // (the element-data halo of the type-2 face integrals travels by ncclSend/ncclRecv)
  bool ok = env_int("PDES_HALO_NCCL", 0) == 0 && np <= 32 && g_nccl.AllGather != nullptr && !ctx->elem_halo;
  if (ok) {
    // + flags | abort | ctr | norm slot ring
    const size_t bytes = 2 * nsend * sizeof(double) + 256 + (32 + 32 + 4) * sizeof(unsigned) + NORM_RING * 32 * 2 * sizeof(double);
    ok = cudaMalloc((void**)&ctx->halo_buf, bytes) == cudaSuccess && cudaMemset(ctx->halo_buf, 0, bytes) == cudaSuccess;
    if (ok) {
      ctx->halo_flags = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(ctx->halo_buf) + 2 * nsend * sizeof(double) + 256);
      ok = cudaIpcGetMemHandle(&mine.handle, ctx->halo_buf) == cudaSuccess;
    }
  }
  cudaGetLastError();
  mine.nsend = (int64_t)nsend; mine.npeers = np; mine.ok = ok ? 1 : 0;
  for (int i = 0; i < np && i < 32; ++i) {
    mine.peer_rank[i] = ctx->peers[i].rank; mine.peer_off[i] = ctx->peers[i].offset; mine.peer_n[i] = ctx->peers[i].nfaces;
  }
  // gather every rank's record
  HaloRec* d_all = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void**)&d_all, sizeof(HaloRec) * ctx->nranks));
  CUDA_TRY(ctx, cudaMemcpyAsync(d_all + ctx->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->comm_stream));
  ncclResult_t r = g_nccl.AllGather ? g_nccl.AllGather(d_all + ctx->rank, d_all, sizeof(HaloRec), ncclChar, ctx->comm, ctx->comm_stream)
                                    : ncclSuccess;
  if (r != ncclSuccess) { set_err(ctx, "ncclAllGather failed: %s", g_nccl.GetErrorString(r)); return PDES_ERR_COMM; }
  std::vector<HaloRec> all(ctx->nranks);
  CUDA_TRY(ctx, cudaMemcpyAsync(all.data(), d_all, sizeof(HaloRec) * ctx->nranks, cudaMemcpyDeviceToHost, ctx->comm_stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->comm_stream));
  cudaFree(d_all);
  if (!g_nccl.AllGather) ok = false;
  ctx->pmap.assign(np, PdesCtx::PeerMap());
  std::vector<unsigned*> flag_ptrs(np > 0 ? np : 1, nullptr);
  // every rank's buffer is mapped once (the norm slots of all ranks, the receive buffers of the neighbours)
  ctx->rank_mapped.assign(ctx->nranks, nullptr);
  std::vector<double*> slot_ptrs(ctx->nranks, nullptr);
  auto tail_of = [](void* base, int64_t ns) { return static_cast<char*>(base) + 2 * (size_t)ns * sizeof(double) + 256; };
  bool all_mapped = ok && ctx->nranks <= 32;
  for (int rr = 0; rr < ctx->nranks && ok; ++rr) {
    void* mapped = ctx->halo_buf;
    if (rr != ctx->rank) {
      mapped = nullptr;
      if (!all[rr].ok || cudaIpcOpenMemHandle(&mapped, all[rr].handle, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        cudaGetLastError(); mapped = nullptr; all_mapped = false;
      }
      ctx->rank_mapped[rr] = mapped;
    }
    if (mapped) slot_ptrs[rr] = reinterpret_cast<double*>(tail_of(mapped, all[rr].nsend) + (32 + 32 + 4) * sizeof(unsigned));
  }
  for (int i = 0; i < np && ok; ++i) {
    const int rr = ctx->peers[i].rank;
    if (rr < 0 || rr >= ctx->nranks || rr == ctx->rank || !ctx->rank_mapped[rr]) { ok = false; break; }
    const HaloRec& o = all[rr];
    int slot = -1;
    for (int k = 0; k < o.npeers; ++k) if (o.peer_rank[k] == ctx->rank) slot = k;
    if (slot < 0 || o.peer_n[slot] != ctx->peers[i].nfaces) { ok = false; break; }
    void* mapped = ctx->rank_mapped[rr];
    PdesCtx::PeerMap& m = ctx->pmap[i];
    m.base = static_cast<double*>(mapped); m.remote_nsend = o.nsend; m.remote_off = o.peer_off[slot];
    m.flag = reinterpret_cast<unsigned*>(tail_of(mapped, o.nsend)) + slot;
    flag_ptrs[i] = m.flag;
  }
  // the same decision everywhere
  int* d_ok = nullptr;
  CUDA_TRY(ctx, cudaMalloc((void**)&d_ok, sizeof(int)));
  const int h_ok = ok ? (all_mapped ? 2 : 1) : 0;           // 2: every rank mapped every buffer (norm through peer memory)
  CUDA_TRY(ctx, cudaMemcpyAsync(d_ok, &h_ok, sizeof(int), cudaMemcpyHostToDevice, ctx->comm_stream));
  r = g_nccl.AllReduce(d_ok, d_ok, 1, ncclInt32, ncclMin, ctx->comm, ctx->comm_stream);
  if (r != ncclSuccess) { set_err(ctx, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r)); return PDES_ERR_COMM; }
  int all_ok = 0;
  CUDA_TRY(ctx, cudaMemcpyAsync(&all_ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost, ctx->comm_stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->comm_stream));
  cudaFree(d_ok);
  if (all_ok) {
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_flag_ptrs, sizeof(unsigned*) * flag_ptrs.size()));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_flag_ptrs, flag_ptrs.data(), sizeof(unsigned*) * flag_ptrs.size(), cudaMemcpyHostToDevice));
    ctx->halo_nsend = nsend;
    ctx->halo_ctr = ctx->halo_flags + 64;
    ctx->p2p = 1;
    if (all_ok == 2 && env_int("PDES_NORM_NCCL", 0) == 0) {
      CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_norm_slots, sizeof(double*) * ctx->nranks));
      CUDA_TRY(ctx, cudaMemcpy(ctx->d_norm_slots, slot_ptrs.data(), sizeof(double*) * ctx->nranks, cudaMemcpyHostToDevice));
      ctx->norm_p2p = true;
    }
    if (env_int("PDES_HALO_COPY", 0) == 0 && ctx->nS > 0) {
      // per shared face and evaluation parity: its slot in the neighbour's receive buffer (k_pack_send stores there)
      std::vector<double*> dst(2 * (size_t)ctx->nS, nullptr);
      for (int par = 0; par < 2; ++par)
        for (int i = 0; i < np; ++i) {
          const Peer& p = ctx->peers[i];
          const PdesCtx::PeerMap& m = ctx->pmap[i];
          for (int64_t j = 0; j < p.nfaces; ++j)
            dst[(size_t)par * ctx->nS + p.offset + j] = m.base + (size_t)par * m.remote_nsend + (size_t)(m.remote_off + j) * per_face;
        }
      CUDA_TRY(ctx, cudaMalloc((void**)&ctx->d_face_dst, sizeof(double*) * dst.size()));
      CUDA_TRY(ctx, cudaMemcpy(ctx->d_face_dst, dst.data(), sizeof(double*) * dst.size(), cudaMemcpyHostToDevice));
      ctx->halo_fused = env_int("PDES_HALO_FUSED", 1) != 0 && ctx->ops->fused_halo() && ctx->nchunks == 1 && !ctx->fused &&
                        !ctx->pipe;
    }
  }
  return PDES_OK;
}

// startSolutionExchange (Utils/parallel.jl:29-49).  With a communicator the whole shared-face branch runs on the
// communication stream, concurrently with the interior faces on the compute stream:
//     comm stream : [q ready] -> k_pack_send -> ncclSend/Recv -> k_face_flux over the shared faces -> [ev_recv]
// (PDES_COMM_INLINE=1: pack and shared-face flux on the compute stream, as before: 4 small kernels serialised per evaluation)
int start_exchange(PdesCtx* ctx, const double* q) {
  // (collective: every rank of the communicator passes here at its first evaluation, shared faces or not)
  if (ctx->comm && ctx->p2p < 0) { int rc = setup_p2p(ctx); if (rc) return rc; }
  if (ctx->nS == 0 || ctx->halo_fused) return PDES_OK;      // (fused halo: the face kernel is the exchange)
  const bool overlap = ctx->comm && ctx->comm_overlap;
  cudaStream_t ps = overlap ? ctx->comm_stream : ctx->stream;
  if (overlap) {
    // q of this evaluation is complete once everything enqueued so far on the compute stream has run
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_q, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_q, 0));
  }
  if (ctx->elem_halo) {
    // parallel_data = element (getSendDataElement, Utils/parallel.jl:276-293): whole elements of the negotiated lists
    const int el_len = ctx->cfg.nn * ctx->nd;
    const int64_t n = ctx->n_send_el * el_len;
    if (n > 0) {
      k_pack_send_element<<<(unsigned)((n + 255) / 256), 256, 0, ps>>>(q, ctx->el_send_list, ctx->n_send_el, el_len, ctx->qel_send, ctx->ctl);
      CUDA_TRY(ctx, cudaGetLastError());
      ctx->launches++;
    }
    if (!ctx->comm) return PDES_OK;   // test mode: receive buffer injected by hand
    if (!overlap) {
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_packed, ctx->stream));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_packed, 0));
    }
    ncclResult_t r = g_nccl.GroupStart();
    for (auto& p : ctx->peers) {
      if (r != ncclSuccess) break;
      r = g_nccl.Recv(ctx->qel_recv + p.el_recv_off * el_len, (size_t)p.n_recv_el * el_len, ncclFloat64, p.rank, ctx->comm, ctx->comm_stream);
      if (r != ncclSuccess) break;
      r = g_nccl.Send(ctx->qel_send + p.el_send_off * el_len, p.send_els.size() * (size_t)el_len, ncclFloat64, p.rank, ctx->comm, ctx->comm_stream);
    }
    ncclResult_t r2 = g_nccl.GroupEnd();
    if (r != ncclSuccess || r2 != ncclSuccess) {
      set_err(ctx, "NCCL send/recv failed: %s", g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
      return PDES_ERR_COMM;
    }
    if (!overlap) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_recv, ctx->comm_stream));
    return PDES_OK;
  }
  const bool put = ctx->comm && ctx->p2p == 1 && ctx->d_face_dst != nullptr;
  const uint32_t ep_next = ctx->halo_epoch + 1;
  CUDA_TRY(ctx, ctx->ops->launch_pack(q, ctx->sh_el, ctx->sh_face, ctx->nS, ctx->q_send,
                                      put ? ctx->d_face_dst + (size_t)(ep_next & 1u) * ctx->nS : nullptr, ctx->ctl, ps));
  ctx->launches++;
  if (!ctx->comm) return PDES_OK;   // test mode: receive buffer injected by hand
  if (!overlap) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_packed, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_packed, 0));
  }
  const size_t per_face = (size_t)ctx->cfg.nfn * ctx->nd;
  if (ctx->p2p == 1) {
    // copy engine -> the neighbour's receive buffer of this evaluation's parity (it consumed that buffer two evaluations
    // ago: it cannot be more than one evaluation behind, since this rank needed ITS data to finish the previous one),
    // then the flag; the wait for the neighbours' flags sits right in front of the shared-face kernel
    const uint32_t ep = ++ctx->halo_epoch;
    const size_t par = ep & 1u;
    for (size_t i = 0; i < ctx->peers.size() && !put; ++i) {      // (PDES_HALO_COPY=1: copy engine instead of remote stores)
      const Peer& p = ctx->peers[i];
      const PdesCtx::PeerMap& m = ctx->pmap[i];
      double* dst = m.base + par * (size_t)m.remote_nsend + (size_t)m.remote_off * per_face;
      CUDA_TRY(ctx, cudaMemcpyAsync(dst, ctx->q_send + p.offset * per_face, sizeof(double) * p.nfaces * per_face,
                                    cudaMemcpyDeviceToDevice, ctx->comm_stream));
    }
    const int np = (int)ctx->peers.size();
    k_halo_signal<<<1, 32, 0, ctx->comm_stream>>>(ctx->d_flag_ptrs, np, ep, ctx->ctl);
    CUDA_TRY(ctx, cudaGetLastError());
    k_halo_wait<<<1, 32, 0, ctx->comm_stream>>>(ctx->halo_flags, np, ep, ctx->ctl);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 2;
    ctx->q_recv_eval = ctx->halo_buf + par * ctx->halo_nsend;
    if (!overlap) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_recv, ctx->comm_stream));
    return PDES_OK;
  }
  ncclResult_t r = g_nccl.GroupStart();
  for (auto& p : ctx->peers) {
    if (r != ncclSuccess) break;
    r = g_nccl.Recv(ctx->q_recv + p.offset * per_face, p.nfaces * per_face, ncclFloat64, p.rank, ctx->comm, ctx->comm_stream);
    if (r != ncclSuccess) break;
    r = g_nccl.Send(ctx->q_send + p.offset * per_face, p.nfaces * per_face, ncclFloat64, p.rank, ctx->comm, ctx->comm_stream);
  }
  ncclResult_t r2 = g_nccl.GroupEnd();
  if (r != ncclSuccess || r2 != ncclSuccess) {
    set_err(ctx, "NCCL send/recv failed: %s", g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
    return PDES_ERR_COMM;
  }
  if (!overlap) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_recv, ctx->comm_stream));
  return PDES_OK;
}

// one residual evaluation:
//   compute stream : pack -> [event] ............ shared-face fluxes (after the receive) -> element chunk 1..C
//   comm stream    :          send/recv
//   face stream    : face chunk 1..C (interior + boundary faces; overlaps the exchange and the element chunks)
// element chunk c waits for face chunk c; the next evaluation's face chunks wait for this evaluation's elements
int enqueue_residual(PdesCtx* ctx, ElemArgs& a, int mode) {
  int rc = start_exchange(ctx, a.q);
  if (rc) return rc;
  const PdesConfig& c = ctx->cfg;
  FaceArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.q = a.q; fa.faces = ctx->faces; fa.nrm = ctx->nrm_all; fa.coords_bndry = ctx->coords_bndry;
  fa.q_recv = ctx->elem_halo ? ctx->qel_recv : ((ctx->comm && ctx->p2p == 1) ? ctx->q_recv_eval : ctx->q_recv);
  fa.fluxe = ctx->fluxe; fa.ctl = ctx->ctl; fa.ph = a.ph;
  fa.nrm_face_stride = ctx->nrm_compact ? c.dim : c.nfn * c.dim;
  fa.nrm_node_stride = ctx->nrm_compact ? 0 : c.dim;
  fa.prefetch_ahead = ctx->prefetch_ahead_faces;
  fa.ext_bc = ctx->has_ext_bc ? 1 : 0;
  if (ctx->fused) {
    FusedArgs fu;
    memset(&fu, 0, sizeof(fu));
    fu.f = fa; fu.e = a; fu.sched = ctx->sched; fu.flags = ctx->flags;
    fu.e.discard_records = ctx->discard_records;
    fu.acquire_fence = ctx->acquire_fence;
    fu.prefetch_ahead_groups = ctx->prefetch_ahead_groups;
    for (int k = 0; k < 2; ++k) {
      const PdesCtx::FusedPlan& pl = ctx->plan[k];
      if (k == 1) {
        if (ctx->nS == 0) break;
        if (ctx->comm) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_recv, 0));
      }
      if (pl.n_groups == 0 && pl.n_tiles == 0) continue;
      fu.f.g0 = pl.g0; fu.f.ng = pl.ng;
      fu.tile_list = pl.tile_list; fu.need = pl.need;
      fu.n_tiles = pl.n_tiles; fu.n_groups = pl.n_groups; fu.lag = pl.lag;
      CUDA_TRY(ctx, ctx->ops->launch_fused(fu, mode, ctx->stream));
      ctx->launches++;
    }
    ctx->n_evals++;
    return PDES_OK;
  }
  const int nc = ctx->nchunks;
  if (ctx->pipe) {
    // chunk pipeline on ONE stream: F0 .. F_lag E0 F_lag+1 E1 ...  Every launch may start while its predecessor drains
    // (programmatic stream serialization); F_k waits for the element chunks <= chunk_dep[k] of the PREVIOUS evaluation
    // (its q, and the record slots it overwrites), E_k for the face chunks <= k of this one.  The records of a chunk
    // are consumed a launch or two after they were written: they stay in L2 and are discarded there (no write-back).
    const uint32_t ep = ++ctx->pipe_epoch;
    unsigned* ctr = ctx->pipe_ctr;
    const int M = PdesCtx::MAXC, FT = ctx->ops->pipe_face_tile(), ET = ctx->ops->tile_elems();
    for (int i = 0; i < nc + ctx->pipe_lag; ++i) {
      if (i < nc) {
        fa.g0 = ctx->chunk_g[i]; fa.ng = ctx->chunk_g[i + 1] - ctx->chunk_g[i];
        PipeArgs& pp = fa.pipe;
        pp.arrive = ctr; pp.done_self = ctr + M; pp.done_dep = ctr + 3 * M;
        pp.on = 1 | ctx->pipe_dbg; pp.chunk = i; pp.ncta = (int32_t)std::max<int64_t>(1, (fa.ng + FT - 1) / FT);
        pp.dep_chunk = ctx->chunk_dep[i]; pp.epoch = ep; pp.dep_epoch = ep - 1;
        CUDA_TRY(ctx, ctx->ops->launch_faces(fa, ctx->stream));
        ctx->launches++;
      }
      const int k = i - ctx->pipe_lag;
      if (k >= 0 && k < nc) {
        a.e_begin = ctx->chunk_e[k]; a.nE = ctx->chunk_e[k + 1];
        PipeArgs& pp = a.pipe;
        pp.arrive = ctr + 2 * M; pp.done_self = ctr + 3 * M; pp.done_dep = ctr + M;
        pp.on = 1 | ctx->pipe_dbg; pp.chunk = k; pp.ncta = (int32_t)((a.nE - a.e_begin + ET - 1) / ET);
        pp.dep_chunk = k; pp.epoch = ep; pp.dep_epoch = ep;
        CUDA_TRY(ctx, ctx->ops->launch_elements(a, mode, ctx->stream));
        ctx->launches++;
      }
    }
    {
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(1); cfg.blockDim = dim3(1); cfg.stream = ctx->stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      CUDA_TRY(ctx, cudaLaunchKernelEx(&cfg, k_pipe_join, (const unsigned*)(ctr + 3 * M + nc - 1), (unsigned)ep, ctx->ctl));
      ctx->launches++;
    }
    a.e_begin = 0; a.nE = c.nE;
    memset(&a.pipe, 0, sizeof(a.pipe));
    ctx->n_evals++;
    return PDES_OK;
  }
  if (ctx->halo_fused) {
    // ONE face launch: send pass over the shared faces, interior + boundary faces, then the shared faces once the
    // neighbours' states have arrived; the element kernel closes the evaluation (++epoch)
    HaloArgs& hx = fa.halo;
    hx.on = 1; hx.npeers = (int32_t)ctx->peers.size(); hx.nS = ctx->nS; hx.s0 = c.nF + c.nB;
    hx.face_dst = ctx->d_face_dst; hx.peer_flags = ctx->d_flag_ptrs; hx.flags = ctx->halo_flags; hx.ctr = ctx->halo_ctr;
    hx.recv_base = ctx->halo_buf; hx.nsend = (int64_t)ctx->halo_nsend;
    fa.g0 = 0; fa.ng = c.nF + c.nB + ctx->nS;
    CUDA_TRY(ctx, ctx->ops->launch_faces(fa, ctx->stream));
    a.halo_epoch = ctx->halo_ctr;
    a.e_begin = 0; a.nE = c.nE;
    CUDA_TRY(ctx, ctx->ops->launch_elements(a, mode, ctx->stream));
    a.halo_epoch = nullptr;
    ctx->launches += 2;
    ctx->n_evals++;
    return PDES_OK;
  }
  cudaStream_t fs = nc > 1 ? ctx->face_stream : ctx->stream;
  if (nc > 1) {
    // q of this evaluation is complete once everything enqueued so far on the compute stream has run
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_elem, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(fs, ctx->ev_elem, 0));
  }
  for (int k = 0; k < nc; ++k) {
    fa.g0 = ctx->chunk_g[k]; fa.ng = ctx->chunk_g[k + 1] - ctx->chunk_g[k];
    CUDA_TRY(ctx, ctx->ops->launch_faces(fa, fs));
    ctx->launches += fa.ng > 0;
    if (nc > 1) CUDA_TRY(ctx, cudaEventRecord(ctx->ev_face[k], fs));
  }
  if (ctx->nS > 0) {
    const bool overlap = ctx->comm && ctx->comm_overlap;
    if (ctx->comm && !overlap) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_recv, 0));
    fa.g0 = c.nF + c.nB; fa.ng = ctx->nS;
    // overlap: stream order on the communication stream puts the kernel behind the receive; it writes the record slots
    // of the shared faces only, while the interior-face kernel is still running on the compute stream
    CUDA_TRY(ctx, ctx->ops->launch_faces(fa, overlap ? ctx->comm_stream : ctx->stream));
    ctx->launches++;
    if (overlap) {
      CUDA_TRY(ctx, cudaEventRecord(ctx->ev_recv, ctx->comm_stream));
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_recv, 0));
    }
  }
  for (int k = 0; k < nc; ++k) {
    if (nc > 1) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_face[k], 0));
    a.e_begin = ctx->chunk_e[k]; a.nE = ctx->chunk_e[k + 1];
    CUDA_TRY(ctx, ctx->ops->launch_elements(a, mode, ctx->stream));
    ctx->launches += a.nE > a.e_begin;
  }
  a.e_begin = 0; a.nE = c.nE;
  ctx->n_evals++;
  return PDES_OK;
}

NormX norm_x(PdesCtx* ctx) {
  NormX nx;
  memset(&nx, 0, sizeof(nx));
  if (ctx->comm && ctx->nranks > 1 && ctx->norm_p2p) {
    nx.on = 1; nx.rank = ctx->rank; nx.nranks = ctx->nranks; nx.slots = ctx->d_norm_slots; nx.nctr = ctx->halo_ctr + 2;
  }
  return nx;
}

NormOut norm_out(PdesCtx* ctx, double res_tol, int pseudo_time, bool fuse) {
  NormOut o;
  // rk4.jl:451-453 reduces the already-reduced norm again: sqrt(P) too large in parallel runs
  o.quirk_scale = (ctx->comm && ctx->nranks > 1) ? (double)ctx->nranks : 1.0;
  o.norms = ctx->norms_dev; o.norms_cap = ctx->norms_cap; o.res_tol = res_tol; o.pseudo_time = pseudo_time; o.fuse = fuse ? 1 : 0;
  return o;
}

// second half of the stage-1 norm: with the peer-memory all-reduce it may be enqueued anywhere later in the stream
int enqueue_norm_commit(PdesCtx* ctx, double res_tol, int pseudo_time) {
  k_norm_commit<<<1, 1, 0, ctx->stream>>>(ctx->norm_sq, norm_out(ctx, res_tol, pseudo_time, false), ctx->ctl, norm_x(ctx));
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  ctx->norm_pending = false;
  return PDES_OK;
}

// defer: the caller enqueues the commit itself (end of the step: by then every rank's partial sum has arrived, the
// commit never waits).  Ignored when the res_tol test needs the norm before stage 2, and on the NCCL path.
int enqueue_norm(PdesCtx* ctx, double res_tol, int pseudo_time, bool defer = false) {
  int n1 = (int)ctx->ops->grid_for(ctx->cfg.nE);
  const bool parallel = ctx->comm && ctx->nranks > 1;
  if (!parallel || ctx->norm_p2p) {
    // one GPU: the reduction kernel commits the norm itself
    k_norm_reduce<<<1, 1024, 0, ctx->stream>>>(ctx->norm_partials, n1, ctx->norm_sq, ctx->ctl, norm_x(ctx),
                                               norm_out(ctx, res_tol, pseudo_time, !parallel));
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    if (!parallel) return PDES_OK;
    if (defer && !(pseudo_time && res_tol >= 0.0)) {
      ctx->norm_pending = true; ctx->pend_tol = res_tol; ctx->pend_pseudo = pseudo_time;
      return PDES_OK;
    }
    return enqueue_norm_commit(ctx, res_tol, pseudo_time);
  }
  // the norm only feeds back into the time loop through the res_tol test; when that test is off the
  // Allreduce + commit run on the communication stream and never stall the stage kernels
  const bool async = !(pseudo_time && res_tol >= 0.0);
  // norm_sq of the previous step consumed (inside a graph capture the step itself joins the branch before it ends)
  if (!ctx->capturing) CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_norm, 0));
  k_norm_reduce<<<1, 1024, 0, ctx->stream>>>(ctx->norm_partials, n1, ctx->norm_sq, ctx->ctl, NormX{},
                                             norm_out(ctx, res_tol, pseudo_time, false));
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  cudaStream_t st = ctx->stream;
  if (async) {
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_packed, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->comm_stream, ctx->ev_packed, 0));
    st = ctx->comm_stream;
  }
  // calcNorm's Allreduce (Utils.jl:443-448): 8 bytes, once per step
  ncclResult_t r = g_nccl.AllReduce(ctx->norm_sq, ctx->norm_sq, 1, ncclFloat64, ncclSum, ctx->comm, st);
  if (r != ncclSuccess) { set_err(ctx, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r)); return PDES_ERR_COMM; }
  k_norm_commit<<<1, 1, 0, st>>>(ctx->norm_sq, norm_out(ctx, res_tol, pseudo_time, false), ctx->ctl, NormX{});
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_norm, st));
  ctx->norm_branch_open = async;
  return PDES_OK;
}

// the four fused stages of one RK4 step (rk4.jl:238-319); stops after stage 1 when only_head is set
int enqueue_rk4_step(PdesCtx* ctx, double h, bool with_norm, double res_tol, int pseudo_time, bool only_head) {
  double* A = ctx->qbuf[ctx->cur];
  double* B = ctx->qbuf[(ctx->cur + 1) % 3];
  double* Cb = ctx->qbuf[(ctx->cur + 2) % 3];
  ElemArgs a;
  const double* in[4] = {A, B, Cb, B};
  double* out[4] = {B, Cb, B, Cb};
  const double ah[4] = {h / 2, h / 2, h, 0.0};
  for (int s = 0; s < 4; ++s) {
    fill_args(ctx, &a, in[s]);
    a.x_old = A; a.ksum = ctx->ksum; a.q_next = out[s]; a.ah = ah[s]; a.h6 = h / 6; a.stage = s + 1;
    a.scheme = ctx->rk4_nosum ? 2 : 0;
    int rc = enqueue_residual(ctx, a, EPI_RK);
    if (rc) return rc;
    if (s == 0) {
      if (with_norm) { rc = enqueue_norm(ctx, res_tol, pseudo_time, !only_head); if (rc) return rc; }
      if (only_head) { ctx->cur = (ctx->cur + 1) % 3; return PDES_OK; }
    }
  }
  if (ctx->norm_pending) { int rc = enqueue_norm_commit(ctx, ctx->pend_tol, ctx->pend_pseudo); if (rc) return rc; }
  ctx->cur = (ctx->cur + 2) % 3;
  return PDES_OK;
}

// the five stages of one lserk54 step (lserk.jl:183-205).  Two state buffers ping-pong (every stage reads the
// neighbours' q while it writes its own), ksum holds dq_vec.  head_only: stage-1 evaluation + norm only, q untouched
// (the reference tests res_tol / itermax BEFORE the stage-1 update, lserk.jl:161-181).
int enqueue_lserk_step(PdesCtx* ctx, double h, double res_tol, int pseudo_time, bool head_only) {
  static const double a_c[5] = {0, -567301805773.0 / 1357537059087.0, -2404267990393.0 / 2016746695238.0,
                                -3550918686646.0 / 2091501179385.0, -1275806237668.0 / 842570457699.0};
  static const double b_c[5] = {1432997174477.0 / 9575080441755.0, 5161836677717.0 / 13612068292357.0,
                                1720146321549.0 / 2090206949498.0, 3134564353537.0 / 4481467310338.0,
                                2277821191437.0 / 14882151754819.0};
  double* P0 = ctx->qbuf[ctx->cur];
  double* P1 = ctx->qbuf[(ctx->cur + 1) % 3];
  ElemArgs a;
  for (int s = 0; s < 5; ++s) {
    double* in = (s % 2 == 0) ? P0 : P1;
    double* out = (s % 2 == 0) ? P1 : P0;
    fill_args(ctx, &a, in);
    a.scheme = 1; a.stage = s + 1; a.x_old = in; a.ksum = ctx->ksum; a.q_next = out;
    a.ah = a_c[s]; a.h6 = b_c[s]; a.hh = h;
    int rc = enqueue_residual(ctx, a, EPI_RK);
    if (rc) return rc;
    if (s == 0) {
      rc = enqueue_norm(ctx, res_tol, pseudo_time);
      if (rc) return rc;
      if (head_only) return PDES_OK;
    }
  }
  ctx->cur = (ctx->cur + 1) % 3;     // five ping-pong stages end in P1
  return PDES_OK;
}

// One full RK4 step as a CUDA graph: the nine launches of a step (at N > 1 with the fused halo: ten) are
// captured once per buffer rotation (three graphs) and replayed; small meshes are launch-bound otherwise.
int launch_rk4_step(PdesCtx* ctx, double h, bool with_norm, double res_tol, int pseudo_time) {
  // (collective one-time set-up of the peer-to-peer halo: not inside a capture)
  if (ctx->comm && ctx->p2p < 0) { int rc = setup_p2p(ctx); if (rc) return rc; }
  // multi-GPU: the fused halo has no host-side state (evaluation number on the device), the step is two launches per stage
  // plus the norm branch (k_norm_reduce -> ncclAllReduce -> k_norm_commit), all capturable
  static const bool graph_mp = env_int("PDES_GRAPH_MP", 1) != 0;
  const bool graphable = ctx->nchunks == 1 && !ctx->no_graph &&
                         (ctx->comm ? (graph_mp && (ctx->nS == 0 || ctx->halo_fused)) : ctx->nS == 0);
  if (!graphable) return enqueue_rk4_step(ctx, h, with_norm, res_tol, pseudo_time, false);
  if (ctx->g_h != h || ctx->g_norm != with_norm || ctx->g_tol != res_tol || ctx->g_pseudo != pseudo_time) {
    for (int i = 0; i < 3; ++i)
      if (ctx->step_graph[i]) { cudaGraphExecDestroy(ctx->step_graph[i]); ctx->step_graph[i] = nullptr; }
    ctx->g_h = h; ctx->g_norm = with_norm; ctx->g_tol = res_tol; ctx->g_pseudo = pseudo_time;
  }
  const int slot = ctx->cur;
  if (!ctx->step_graph[slot]) {
    const int64_t l0 = ctx->launches, n0 = ctx->n_evals;
    cudaGraph_t g = nullptr;
    CUDA_TRY(ctx, cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
    ctx->capturing = true; ctx->norm_branch_open = false;
    int rc = enqueue_rk4_step(ctx, h, with_norm, res_tol, pseudo_time, false);
    if (ctx->norm_branch_open) cudaStreamWaitEvent(ctx->stream, ctx->ev_norm, 0);       // join the all-reduce branch
    ctx->capturing = false; ctx->norm_branch_open = false;
    cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
    ctx->cur = slot;                       // the capture only recorded the step; it is executed below
    ctx->g_launches = (int)(ctx->launches - l0);
    ctx->launches = l0; ctx->n_evals = n0;
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    CUDA_TRY(ctx, e);
    CUDA_TRY(ctx, cudaGraphInstantiate(&ctx->step_graph[slot], g, 0));
    cudaGraphDestroy(g);
  }
  CUDA_TRY(ctx, cudaGraphLaunch(ctx->step_graph[slot], ctx->stream));
  ctx->launches += ctx->g_launches;
  ctx->n_evals += 4;
  ctx->cur = (ctx->cur + 2) % 3;
  return PDES_OK;
}

// out = dR/dq(q) * v on device vectors (q = the resident state).  Partitioned mesh: the shared-face states AND the
// shared-face values of the direction travel first (the reference's complex-step product exchanges the perturbed complex
// state, newton_setup.jl:632-662 with parallel_data from read_input.jl:250-258): ncclSend/ncclRecv in stream order -- these
// products serve the Krylov loop, not the RK4 hot loop.
// halo_mode 0: exchange states and directions; 1: exchange the states only, the direction is zero on the neighbours' side
// (probing of the element-diagonal blocks); 2: no exchange at all (states and the zero direction of a previous mode-1 call)
int enqueue_jvp(PdesCtx* ctx, const double* vdev, double* odev, int halo_mode = 0) {
  const PdesConfig& c = ctx->cfg;
  ElemArgs a;
  fill_args(ctx, &a, ctx->qbuf[ctx->cur]);
  FaceArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.q = a.q; fa.faces = ctx->faces; fa.nrm = ctx->nrm_all; fa.coords_bndry = ctx->coords_bndry;
  fa.q_recv = ctx->q_recv; fa.fluxe = ctx->fluxe; fa.ctl = ctx->ctl; fa.ph = a.ph;
  fa.nrm_face_stride = ctx->nrm_compact ? c.dim : c.nfn * c.dim;
  fa.nrm_node_stride = ctx->nrm_compact ? 0 : c.dim;
  fa.g0 = 0; fa.ng = c.nF + c.nB;
  if (ctx->nS > 0) {
    if (!ctx->comm) {
      set_err(ctx, "J*v on a partitioned mesh needs a communicator (pdes_set_comm)");
      return PDES_ERR_UNSUPPORTED;
    }
    const size_t per_face = (size_t)c.nfn * ctx->nd, nsend = (size_t)ctx->nS * per_face;
    if (!ctx->v_send) {
      CUDA_TRY(ctx, cudaMalloc((void**)&ctx->v_send, sizeof(double) * nsend));
      CUDA_TRY(ctx, cudaMalloc((void**)&ctx->v_recv, sizeof(double) * nsend));
    }
    if (halo_mode < 2) {
      CUDA_TRY(ctx, ctx->ops->launch_pack(a.q, ctx->sh_el, ctx->sh_face, ctx->nS, ctx->q_send, nullptr, ctx->ctl, ctx->stream));
      if (halo_mode == 0)
        CUDA_TRY(ctx, ctx->ops->launch_pack(vdev, ctx->sh_el, ctx->sh_face, ctx->nS, ctx->v_send, nullptr, ctx->ctl, ctx->stream));
      else
        CUDA_TRY(ctx, cudaMemsetAsync(ctx->v_recv, 0, sizeof(double) * nsend, ctx->stream));
      ctx->launches += 2;
      ncclResult_t r = g_nccl.GroupStart();
      for (auto& p : ctx->peers) {
        if (r != ncclSuccess) break;
        const size_t off = (size_t)p.offset * per_face, cnt = (size_t)p.nfaces * per_face;
        r = g_nccl.Recv(ctx->q_recv + off, cnt, ncclFloat64, p.rank, ctx->comm, ctx->stream);
        if (r == ncclSuccess) r = g_nccl.Send(ctx->q_send + off, cnt, ncclFloat64, p.rank, ctx->comm, ctx->stream);
        if (halo_mode == 0) {
          if (r == ncclSuccess) r = g_nccl.Recv(ctx->v_recv + off, cnt, ncclFloat64, p.rank, ctx->comm, ctx->stream);
          if (r == ncclSuccess) r = g_nccl.Send(ctx->v_send + off, cnt, ncclFloat64, p.rank, ctx->comm, ctx->stream);
        }
      }
      ncclResult_t r2 = g_nccl.GroupEnd();
      if (r != ncclSuccess || r2 != ncclSuccess) {
        set_err(ctx, "NCCL send/recv failed: %s", g_nccl.GetErrorString(r != ncclSuccess ? r : r2));
        return PDES_ERR_COMM;
      }
    }
    fa.v_recv = ctx->v_recv;
    fa.ng = c.nF + c.nB + ctx->nS;
  }
  cudaError_t e = ctx->ops->launch_jvp(fa, a, vdev, odev, ctx->stream);
  if (e == cudaErrorNotSupported) {
    set_err(ctx, "pdes_eval_jvp is not implemented for face_integral_type 2 (face-element integrals)");
    return PDES_ERR_UNSUPPORTED;
  }
  CUDA_TRY(ctx, e);
  ctx->launches += 2;
  return PDES_OK;
}

}  // namespace

extern "C" {

const char* pdes_last_error(const PdesCtx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int pdes_last_error_location(const PdesCtx* ctx, int64_t* element, int64_t* node) {
  if (!ctx) return PDES_ERR_USAGE;
  if (element) *element = ctx->err_element;
  if (node) *node = ctx->err_node;
  return PDES_OK;
}

int pdes_create(const PdesConfig* cfg, PdesCtx** out) {
  if (!cfg || !out) return usage(nullptr, "pdes_create: null argument");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    set_err(nullptr, "no usable CUDA device (%s); this library has no CPU fallback", cudaGetErrorString(e));
    return PDES_ERR_CUDA;
  }
  if (cfg->device < 0 || cfg->device >= ndev) return usage(nullptr, "pdes_create: device ordinal out of range");
  cudaDeviceProp prop;
  CUDA_TRY(nullptr, cudaGetDeviceProperties(&prop, cfg->device));
  if (prop.major != 10) {
    set_err(nullptr, "device %d is sm_%d%d; this library contains sm_100a code only", cfg->device, prop.major, prop.minor);
    return PDES_ERR_CUDA;
  }
  if (cfg->dim != 2 && cfg->dim != 3) return usage(nullptr, "pdes_create: dim must be 2 or 3");
  if (cfg->face_integral_type != 1 && cfg->face_integral_type != 2) {
    // euler.jl:796: ErrorException("Unsupported face integral type")
    set_err(nullptr, "Unsupported face integral type = %d", cfg->face_integral_type);
    return PDES_ERR_UNSUPPORTED;
  }
  if (cfg->nE <= 0 || cfg->nE * (cfg->dim + 1) > 0x7fffff00ll)
    return usage(nullptr, "pdes_create: numEl out of range (32-bit element-face indices)");
  std::unique_ptr<PdesCtx> ctx(new PdesCtx());
  ctx->cfg = *cfg;
  ctx->nd = cfg->dim + 2;
  ctx->nf = cfg->dim + 1;
  ctx->ndof = (int64_t)ctx->nd * cfg->nn * cfg->nE;
  ctx->ops.reset(make_ops(*cfg));
  if (!ctx->ops) {
    set_err(nullptr,
            "unsupported operator/flux combination (dim=%d nn=%d nfn=%d sparse=%d volume_integral_type=%d flux=%d)",
            cfg->dim, cfg->nn, cfg->nfn, cfg->sparse_face, cfg->volume_integral_type, cfg->flux_id);
    return PDES_ERR_UNSUPPORTED;
  }
  CUDA_TRY(nullptr, cudaSetDevice(cfg->device));
  PdesCtx* c = ctx.get();
  {
    // the compute stream (element kernels) outranks the face stream: when the two kernel families run concurrently
    // (PDES_CHUNKS > 1) pending element CTAs are dispatched before the queued face CTAs
    int lo = 0, hi = 0;
    CUDA_TRY(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(c, cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, hi));
  }
  {
    // the shared-face branch must not queue behind the thousands of interior-face CTAs it overlaps with
    int lo = 0, hi = 0;
    CUDA_TRY(c, cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CUDA_TRY(c, cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi));
  }
  CUDA_TRY(c, cudaStreamCreateWithFlags(&c->face_stream, cudaStreamNonBlocking));
  for (int i = 0; i < PdesCtx::MAXC; ++i) CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_face[i], cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_elem, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_q, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_recv, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreateWithFlags(&c->ev_norm, cudaEventDisableTiming));
  CUDA_TRY(c, cudaEventCreate(&c->ev_t0));
  CUDA_TRY(c, cudaEventCreate(&c->ev_t1));
  for (int i = 0; i < 3; ++i) CUDA_TRY(c, dev_upload<double>(c->stream, &c->qbuf[i], nullptr, (size_t)c->ndof + 2));
  CUDA_TRY(c, dev_upload<double>(c->stream, &c->ksum, nullptr, (size_t)c->ndof));
  CUDA_TRY(c, dev_upload<double>(c->stream, &c->res, nullptr, (size_t)c->ndof));
  CUDA_TRY(c, cudaMalloc((void**)&c->ctl, sizeof(Ctl)));
  CUDA_TRY(c, cudaMallocHost((void**)&c->h_ctl, sizeof(Ctl)));
  CUDA_TRY(c, cudaMalloc((void**)&c->norm_sq, sizeof(double)));
  CUDA_TRY(c, cudaMalloc((void**)&c->tile_ctr, 4 * sizeof(unsigned)));
  CUDA_TRY(c, cudaMemset(c->tile_ctr, 0, 4 * sizeof(unsigned)));
  c->peers.resize(cfg->npeers > 0 ? cfg->npeers : 0);
  int rc = reset_ctl(c);
  if (rc) return rc;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  *out = ctx.release();
  return PDES_OK;
}

void pdes_destroy(PdesCtx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->cfg.device);
  cudaDeviceSynchronize();
  for (int i = 0; i < 3; ++i) if (ctx->step_graph[i]) cudaGraphExecDestroy(ctx->step_graph[i]);
  for (void* m : ctx->rank_mapped) if (m) cudaIpcCloseMemHandle(m);
  if (ctx->halo_buf && ctx->comm && ctx->p2p == 1 && g_nccl.AllReduce) {
    // the receive buffer is mapped by the neighbours (CUDA IPC): freeing it while a peer still holds the mapping is
    // undefined, so every rank closes its imported handles first (above) and the ranks meet here before the export
    // goes away -- pdes_destroy is collective over the communicator, like the ncclCommDestroy that follows
    int* d_tok = nullptr;
    if (cudaMalloc((void**)&d_tok, sizeof(int)) == cudaSuccess) {
      cudaMemsetAsync(d_tok, 0, sizeof(int), ctx->comm_stream);
      if (g_nccl.AllReduce(d_tok, d_tok, 1, ncclInt32, ncclSum, ctx->comm, ctx->comm_stream) == ncclSuccess)
        cudaStreamSynchronize(ctx->comm_stream);
      cudaFree(d_tok);
    }
    cudaGetLastError();
  }
  if (ctx->halo_buf) cudaFree(ctx->halo_buf);
  if (ctx->d_flag_ptrs) cudaFree(ctx->d_flag_ptrs);
  if (ctx->d_face_dst) cudaFree(ctx->d_face_dst);
  if (ctx->d_norm_slots) cudaFree(ctx->d_norm_slots);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  void* ptrs[] = {ctx->qbuf[0], ctx->qbuf[1], ctx->qbuf[2], ctx->ksum, ctx->res, ctx->dxidx, ctx->minv, ctx->srcw,
                  ctx->nrm_all, ctx->fluxe, ctx->srcm, ctx->faces, ctx->coords_bndry, ctx->w_dev, ctx->Q_dev,
                  ctx->q_send, ctx->q_recv, ctx->v_send, ctx->v_recv, ctx->el_send_list, ctx->qel_send, ctx->qel_recv, ctx->sh_el, ctx->sh_face, ctx->ctl, ctx->norm_partials,
                  ctx->norm_sq, ctx->tile_ctr, ctx->norms_dev, ctx->plan[0].tile_list, ctx->plan[0].need, ctx->plan[1].tile_list,
                  ctx->plan[1].need, ctx->flags, ctx->sched, ctx->mass, ctx->diag_buf, ctx->kry.V, ctx->kry.w, ctx->kry.b, ctx->kry.x,
                  ctx->kry.partials, ctx->kry.hdev, ctx->kry.gH, ctx->kry.gcs, ctx->kry.gsn, ctx->kry.gg, ctx->kry.gstate, ctx->kry.pc_blocks, ctx->kry.pcz, ctx->kry.pcu, ctx->kry.colour};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (ctx->kry.hhost) cudaFreeHost(ctx->kry.hhost);
  if (ctx->kry.hstate) cudaFreeHost(ctx->kry.hstate);
  for (cudaGraphExec_t ge : ctx->kry.itg) if (ge) cudaGraphExecDestroy(ge);
  if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
  if (ctx->ev_packed) cudaEventDestroy(ctx->ev_packed);
  if (ctx->ev_q) cudaEventDestroy(ctx->ev_q);
  if (ctx->ev_recv) cudaEventDestroy(ctx->ev_recv);
  if (ctx->ev_norm) cudaEventDestroy(ctx->ev_norm);
  if (ctx->ev_t0) cudaEventDestroy(ctx->ev_t0);
  if (ctx->ev_t1) cudaEventDestroy(ctx->ev_t1);
  for (int i = 0; i < PdesCtx::MAXC; ++i) if (ctx->ev_face[i]) cudaEventDestroy(ctx->ev_face[i]);
  if (ctx->ev_elem) cudaEventDestroy(ctx->ev_elem);
  for (int k = 0; k < PdesCtx::HC; ++k) {
    if (ctx->ev_up[k]) cudaEventDestroy(ctx->ev_up[k]);
    if (ctx->ev_el[k]) cudaEventDestroy(ctx->ev_el[k]);
  }
  if (ctx->ev_idle) cudaEventDestroy(ctx->ev_idle);
  if (ctx->up_stream) cudaStreamDestroy(ctx->up_stream);
  if (ctx->down_stream) cudaStreamDestroy(ctx->down_stream);
  if (ctx->face_stream) cudaStreamDestroy(ctx->face_stream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->comm_stream) cudaStreamDestroy(ctx->comm_stream);
  delete ctx;
}

int pdes_set_operator(PdesCtx* ctx, const double* Q, const double* w, const double* interp, const int64_t* perm,
                      const int64_t* nbrperm, const double* wface) {
  if (!ctx || !Q || !w || !interp || !perm || !nbrperm || !wface) return usage(ctx, "pdes_set_operator: null argument");
  const PdesConfig& c = ctx->cfg;
  if (!c.sparse_face && (c.ss < 1 || c.ss > c.nn)) return usage(ctx, "dense face operators need 1 <= stencilsize <= numnodes");
  const int nor = c.dim == 2 ? 1 : 3;
  if (c.norient != nor) return usage(ctx, "norient must be 1 (2D) or 3 (3D)");
  const int nperm = c.sparse_face ? c.nfn : c.ss;     // sbpface.perm is [nfn, numfaces] for a SparseFace
  for (int f = 0; f < ctx->nf; ++f)
    for (int j = 0; j < nperm; ++j) {
      int64_t p = perm[j + (int64_t)nperm * f] - c.index_base;
      if (p < 0 || p >= c.nn) return usage(ctx, "sbpface.perm entry out of range");
    }
  for (int o = 0; o < nor; ++o)
    for (int i = 0; i < c.nfn; ++i) {
      int64_t p = nbrperm[i + c.nfn * o] - c.index_base;
      if (p < 0 || p >= c.nfn) return usage(ctx, "sbpface.nbrperm entry out of range");
    }
  CUDA_TRY(ctx, cudaSetDevice(c.device));
  ctx->ops->build_tables(c, Q, w, interp, perm, nbrperm, wface, c.index_base);
  ctx->h_w.assign(w, w + c.nn);
  CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->w_dev, w, (size_t)c.nn));
  CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->Q_dev, Q, (size_t)c.nn * c.nn * c.dim));
  ctx->have_op = true;
  ctx->finalized = false;
  return PDES_OK;
}

int pdes_set_mesh(PdesCtx* ctx, const double* dxidx, const double* jac, const double* coords, const double* nrm_face,
                  const double* nrm_bndry, const double* coords_bndry, const PdesInterface* interfaces,
                  const PdesBoundary* bndryfaces, const int64_t* bndry_offsets, const int32_t* bc_ids) {
  if (!ctx || !dxidx || !jac) return usage(ctx, "pdes_set_mesh: null argument");
  if (!ctx->have_op) return usage(ctx, "pdes_set_operator must precede pdes_set_mesh");
  const PdesConfig& c = ctx->cfg;
  const int NF = ctx->nf, base = c.index_base;
  if (c.nF > 0 && (!interfaces || !nrm_face)) return usage(ctx, "pdes_set_mesh: interfaces missing");
  if (c.nB > 0 && (!bndryfaces || !nrm_bndry || !coords_bndry || !bndry_offsets || !bc_ids))
    return usage(ctx, "pdes_set_mesh: boundary arrays missing");
  if (c.src_id != PDES_SRC_NONE && !coords) return usage(ctx, "pdes_set_mesh: coords needed for the source term");
  CUDA_TRY(ctx, cudaSetDevice(c.device));
  std::vector<EFace>& ef = ctx->h_efaces;
  std::vector<FaceRec>& faces = ctx->h_faces;
  EFace blank;
  memset(&blank, 0, sizeof(blank));
  blank.gface = -1;
  ef.assign((size_t)c.nE * NF, blank);
  faces.assign((size_t)(c.nF + c.nB), FaceRec());
  for (int64_t f = 0; f < c.nF; ++f) {
    const PdesInterface& I = interfaces[f];
    int64_t eL = (int64_t)I.elementL - base, eR = (int64_t)I.elementR - base;
    int fL = (int)I.faceL - base, fR = (int)I.faceR - base, o = (int)I.orient - base;
    if (eL < 0 || eL >= c.nE || eR < 0 || eR >= c.nE || fL < 0 || fL >= NF || fR < 0 || fR >= NF || o < 0 ||
        o >= c.norient)
      return usage(ctx, "mesh.interfaces entry out of range");
    EFace& L = ef[eL * NF + fL];
    EFace& Rr = ef[eR * NF + fR];
    if (L.gface >= 0 || Rr.gface >= 0) return usage(ctx, "element face referenced by two interfaces");
    L.gface = (int32_t)f; L.right = 0; L.orient = (uint8_t)o;
    Rr.gface = (int32_t)f; Rr.right = 1; Rr.orient = (uint8_t)o;
    FaceRec& fr = faces[f];
    memset(&fr, 0, sizeof(fr));
    fr.elL = (int32_t)eL; fr.elR = (int32_t)eR; fr.fL = (uint8_t)fL; fr.fR = (uint8_t)fR; fr.orient = (uint8_t)o;
    fr.kind = FK_INTERIOR;
  }
  std::vector<char> bseen((size_t)c.nB, 0);
  ctx->has_ext_bc = false;
  for (int i = 0; i < c.numBC; ++i) {
    if (bc_ids[i] >= PDES_BC_RHO1E2U3) ctx->has_ext_bc = true;
    if (bc_ids[i] < PDES_BC_ISENTROPIC_VORTEX || bc_ids[i] > PDES_BC_NOPENETRATION_ES) {
      set_err(ctx, "BC id %d is not supported", bc_ids[i]);
      return PDES_ERR_UNSUPPORTED;
    }
    for (int64_t b = bndry_offsets[i] - base; b < bndry_offsets[i + 1] - base; ++b) {
      if (b < 0 || b >= c.nB) return usage(ctx, "bndry_offsets out of range");
      int64_t el = (int64_t)bndryfaces[b].element - base;
      int f = (int)bndryfaces[b].face - base;
      if (el < 0 || el >= c.nE || f < 0 || f >= NF) return usage(ctx, "mesh.bndryfaces entry out of range");
      EFace& r = ef[el * NF + f];
      if (r.gface >= 0) return usage(ctx, "boundary face already claimed by an interface");
      r.gface = (int32_t)(c.nF + b); r.right = 0; r.orient = 0;
      FaceRec& fr = faces[c.nF + b];
      memset(&fr, 0, sizeof(fr));
      fr.elL = (int32_t)el; fr.elR = -1; fr.fL = (uint8_t)f; fr.kind = FK_BOUNDARY; fr.aux = bc_ids[i];
      bseen[b] = 1;
    }
  }
  for (int64_t b = 0; b < c.nB; ++b)
    if (!bseen[b]) return usage(ctx, "boundary face not covered by bndry_offsets");
  {
    const size_t per_nrm = (size_t)c.nfn * c.dim;
    ctx->h_nrm.resize((size_t)(c.nF + c.nB) * per_nrm);
    if (c.nF) memcpy(ctx->h_nrm.data(), nrm_face, sizeof(double) * (size_t)c.nF * per_nrm);
    if (c.nB) memcpy(ctx->h_nrm.data() + (size_t)c.nF * per_nrm, nrm_bndry, sizeof(double) * (size_t)c.nB * per_nrm);
  }
  const size_t nnE = (size_t)c.nn * c.nE;
  {
    // straight-sided elements have node-independent dxidx: store one matrix per element (the reference stores it
    // per node, docs/src/interfaces.md:351-357); exact comparison, so curved meshes take the general path
    const int dd = c.dim * c.dim;
    bool compact = env_int("PDES_NO_COMPACT", 0) == 0;
    for (int64_t e = 0; e < c.nE && compact; ++e)
      for (int j = 1; j < c.nn && compact; ++j)
        compact = memcmp(dxidx + ((size_t)e * c.nn + j) * dd, dxidx + (size_t)e * c.nn * dd, sizeof(double) * dd) == 0;
    ctx->dx_compact = compact;
    if (compact) {
      std::vector<double> dc((size_t)c.nE * dd);
      for (int64_t e = 0; e < c.nE; ++e) memcpy(dc.data() + (size_t)e * dd, dxidx + (size_t)e * c.nn * dd, sizeof(double) * dd);
      CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->dxidx, dc.data(), dc.size()));
    } else {
      CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->dxidx, dxidx, nnE * c.dim * c.dim));
    }
  }
  CUDA_TRY(ctx, dev_upload(ctx->stream, &ctx->coords_bndry, coords_bndry, (size_t)c.nB * c.nfn * c.dim));
  // Minv and the tabulated source are produced on the device from jac / coords
  double* jac_dev = nullptr;
  CUDA_TRY(ctx, dev_upload(ctx->stream, &jac_dev, jac, nnE));
  CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->minv, nullptr, nnE));
  CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->mass, nullptr, nnE));
  unsigned nb = (unsigned)((nnE + 255) / 256);
  k_minv<<<nb, 256, 0, ctx->stream>>>(jac_dev, ctx->w_dev, c.nn, c.nE, ctx->minv, ctx->mass);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  if (c.src_id == PDES_SRC_EXP) {
    double* coords_dev = nullptr;
    CUDA_TRY(ctx, dev_upload(ctx->stream, &coords_dev, coords, nnE * c.dim));
    CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->srcw, nullptr, (size_t)ctx->ndof));
    if (c.dim == 2) k_tabulate_source<2><<<nb, 256, 0, ctx->stream>>>(coords_dev, jac_dev, ctx->w_dev, c.nn, c.nE, c.gamma, ctx->srcw);
    else k_tabulate_source<3><<<nb, 256, 0, ctx->stream>>>(coords_dev, jac_dev, ctx->w_dev, c.nn, c.nE, c.gamma, ctx->srcw);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    CUDA_TRY(ctx, dev_upload<double>(ctx->stream, &ctx->srcm, nullptr, (size_t)ctx->ndof));
    k_srcm<<<(unsigned)((ctx->ndof + 255) / 256), 256, 0, ctx->stream>>>(ctx->srcw, ctx->minv, ctx->nd, (int64_t)nnE, ctx->srcm);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(coords_dev);
  } else if (c.src_id != PDES_SRC_NONE) {
    set_err(ctx, "source id %d is not supported", c.src_id);
    return PDES_ERR_UNSUPPORTED;
  }
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  cudaFree(jac_dev);
  ctx->have_mesh = true;
  ctx->finalized = false;
  return PDES_OK;
}

int pdes_set_peer(PdesCtx* ctx, int32_t peer_idx, int32_t peer_rank, int64_t nfaces, const PdesBoundary* bndries_local,
                  const PdesInterface* shared_interfaces, const double* nrm_sharedface) {
  if (!ctx) return usage(ctx, "pdes_set_peer: null ctx");
  if (peer_idx < 0 || peer_idx >= (int)ctx->peers.size()) return usage(ctx, "pdes_set_peer: peer index out of range");
  if (nfaces < 0 || (nfaces > 0 && (!bndries_local || !shared_interfaces || !nrm_sharedface)))
    return usage(ctx, "pdes_set_peer: null argument");
  Peer& p = ctx->peers[peer_idx];
  p.rank = peer_rank;
  p.nfaces = nfaces;
  p.bndries_local.assign(bndries_local, bndries_local + nfaces);
  p.ifaces.assign(shared_interfaces, shared_interfaces + nfaces);
  p.nrm.assign(nrm_sharedface, nrm_sharedface + (size_t)nfaces * ctx->cfg.nfn * ctx->cfg.dim);
  ctx->finalized = false;
  return PDES_OK;
}

int pdes_set_peer_elements(PdesCtx* ctx, int32_t peer_idx, int64_t nsend, const int64_t* local_elements, int64_t nrecv,
                           int64_t shared_element_offset) {
  if (!ctx) return usage(ctx, "pdes_set_peer_elements: null ctx");
  if (peer_idx < 0 || peer_idx >= (int)ctx->peers.size()) return usage(ctx, "pdes_set_peer_elements: peer index out of range");
  if (nsend < 0 || nrecv < 0 || (nsend > 0 && !local_elements)) return usage(ctx, "pdes_set_peer_elements: bad argument");
  Peer& p = ctx->peers[peer_idx];
  p.send_els.assign(local_elements, local_elements + nsend);
  p.n_recv_el = nrecv;
  p.shared_el_offset = shared_element_offset;
  p.have_els = true;
  ctx->finalized = false;
  return PDES_OK;
}

int pdes_pack_send_elements(PdesCtx* ctx, int32_t peer_idx, double* q_send_out) {
  if (!ctx || !q_send_out) return usage(ctx, "pdes_pack_send_elements: null argument");
  int rc = finalize(ctx);
  if (rc) return rc;
  if (peer_idx < 0 || peer_idx >= (int)ctx->peers.size() || !ctx->elem_halo) return usage(ctx, "no element-data halo for this peer");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const int el_len = ctx->cfg.nn * ctx->nd;
  const int64_t n = ctx->n_send_el * el_len;
  if (n > 0) {
    k_pack_send_element<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->qbuf[ctx->cur], ctx->el_send_list, ctx->n_send_el,
                                                                              el_len, ctx->qel_send, ctx->ctl);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
  }
  const Peer& p = ctx->peers[peer_idx];
  CUDA_TRY(ctx, cudaMemcpyAsync(q_send_out, ctx->qel_send + p.el_send_off * el_len, sizeof(double) * p.send_els.size() * el_len,
                                cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

int pdes_inject_recv_elements(PdesCtx* ctx, int32_t peer_idx, const double* q_recv) {
  if (!ctx || !q_recv) return usage(ctx, "pdes_inject_recv_elements: null argument");
  int rc = finalize(ctx);
  if (rc) return rc;
  if (peer_idx < 0 || peer_idx >= (int)ctx->peers.size() || !ctx->elem_halo) return usage(ctx, "no element-data halo for this peer");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const Peer& p = ctx->peers[peer_idx];
  const int el_len = ctx->cfg.nn * ctx->nd;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->qel_recv + p.el_recv_off * el_len, q_recv, sizeof(double) * (size_t)p.n_recv_el * el_len,
                                cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

int pdes_get_unique_id(uint8_t id_out[128]) {
  std::string why;
  if (!g_nccl.load(&why)) { set_err(nullptr, "%s", why.c_str()); return PDES_ERR_COMM; }
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) { set_err(nullptr, "ncclGetUniqueId: %s", g_nccl.GetErrorString(r)); return PDES_ERR_COMM; }
  memcpy(id_out, &id, 128);
  return PDES_OK;
}

int pdes_set_comm(PdesCtx* ctx, const uint8_t id[128], int32_t rank, int32_t nranks) {
  if (!ctx || !id) return usage(ctx, "pdes_set_comm: null argument");
  std::string why;
  if (!g_nccl.load(&why)) { set_err(ctx, "%s", why.c_str()); return PDES_ERR_COMM; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  ncclUniqueId uid;
  memcpy(&uid, id, 128);
  ncclResult_t r = g_nccl.CommInitRank(&ctx->comm, nranks, uid, rank);
  if (r != ncclSuccess) { set_err(ctx, "ncclCommInitRank: %s", g_nccl.GetErrorString(r)); return PDES_ERR_COMM; }
  ctx->rank = rank;
  ctx->nranks = nranks;
  return PDES_OK;
}

int pdes_pack_send(PdesCtx* ctx, int32_t peer_idx, double* q_send_out) {
  if (!ctx || !q_send_out) return usage(ctx, "pdes_pack_send: null argument");
  int rc = finalize(ctx);
  if (rc) return rc;
  if (peer_idx < 0 || peer_idx >= (int)ctx->peers.size()) return usage(ctx, "peer index out of range");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  CUDA_TRY(ctx, ctx->ops->launch_pack(ctx->qbuf[ctx->cur], ctx->sh_el, ctx->sh_face, ctx->nS, ctx->q_send, nullptr, ctx->ctl,
                                      ctx->stream));
  ctx->launches++;
  const Peer& p = ctx->peers[peer_idx];
  const size_t per_face = (size_t)ctx->cfg.nfn * ctx->nd;
  CUDA_TRY(ctx, cudaMemcpyAsync(q_send_out, ctx->q_send + p.offset * per_face, sizeof(double) * p.nfaces * per_face,
                                cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

int pdes_inject_recv(PdesCtx* ctx, int32_t peer_idx, const double* q_recv) {
  if (!ctx || !q_recv) return usage(ctx, "pdes_inject_recv: null argument");
  int rc = finalize(ctx);
  if (rc) return rc;
  if (peer_idx < 0 || peer_idx >= (int)ctx->peers.size()) return usage(ctx, "peer index out of range");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const Peer& p = ctx->peers[peer_idx];
  const size_t per_face = (size_t)ctx->cfg.nfn * ctx->nd;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->q_recv + p.offset * per_face, q_recv, sizeof(double) * p.nfaces * per_face,
                                cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

int pdes_set_q(PdesCtx* ctx, const double* q) {
  if (!ctx || !q) return usage(ctx, "pdes_set_q: null argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->qbuf[ctx->cur], q, sizeof(double) * ctx->ndof, cudaMemcpyHostToDevice, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

int pdes_get_q(PdesCtx* ctx, double* q) {
  if (!ctx || !q) return usage(ctx, "pdes_get_q: null argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  CUDA_TRY(ctx, cudaMemcpyAsync(q, ctx->qbuf[ctx->cur], sizeof(double) * ctx->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

int pdes_get_res(PdesCtx* ctx, double* res) {
  if (!ctx || !res) return usage(ctx, "pdes_get_res: null argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  CUDA_TRY(ctx, cudaMemcpyAsync(res, ctx->res, sizeof(double) * ctx->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

int pdes_set_q_dev(PdesCtx* ctx, const double* q_dev) {
  if (!ctx || !q_dev) return usage(ctx, "pdes_set_q_dev: null argument");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->qbuf[ctx->cur], q_dev, sizeof(double) * ctx->ndof, cudaMemcpyDeviceToDevice, ctx->stream));
  return PDES_OK;
}

int pdes_pin_host(void* ptr, int64_t bytes) {
  if (!ptr || bytes <= 0) return usage(nullptr, "pdes_pin_host: bad argument");
  CUDA_TRY(nullptr, cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterDefault));
  return PDES_OK;
}

int pdes_unpin_host(void* ptr) {
  if (!ptr) return usage(nullptr, "pdes_unpin_host: null");
  CUDA_TRY(nullptr, cudaHostUnregister(ptr));
  return PDES_OK;
}

void* pdes_stream(PdesCtx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

double* pdes_q_dev(PdesCtx* ctx) { return ctx ? ctx->qbuf[ctx->cur] : nullptr; }
double* pdes_res_dev(PdesCtx* ctx) { return ctx ? ctx->res : nullptr; }

int pdes_eval_residual_async(PdesCtx* ctx, double t) {
  (void)t;  // the scoped BCs and SRCExp are time independent
  if (!ctx) return usage(ctx, "null ctx");
  int rc = finalize(ctx);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  ElemArgs a;
  fill_args(ctx, &a, ctx->qbuf[ctx->cur]);
  a.res = ctx->res;
  return enqueue_residual(ctx, a, EPI_RES);
}

int pdes_sync(PdesCtx* ctx) {
  if (!ctx) return usage(ctx, "null ctx");
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  int rc = fetch_ctl(ctx);
  if (rc > 0) reset_ctl(ctx);
  return rc;
}

int pdes_eval_residual(PdesCtx* ctx, double t) {
  if (!ctx) return usage(ctx, "null ctx");
  cudaEventRecord(ctx->ev_t0, ctx->stream);
  int rc = pdes_eval_residual_async(ctx, t);
  if (rc) return rc;
  cudaEventRecord(ctx->ev_t1, ctx->stream);
  rc = pdes_sync(ctx);
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1) == cudaSuccess) ctx->tm.t_func += ms * 1e-3;   // Timings.t_func
  return rc;
}

// evalResidual with host arrays in ONE call: the evaluation is cut into HC element / face chunks and pipelined with the
// copies -- chunk k of q goes up while the faces and elements of the chunks before it are evaluated, chunk k of res comes
// down (PCIe is full duplex) as soon as its elements are done.  A host-driven time integrator pays ~max(upload, download)
// instead of upload + evaluation + download per call.  Same kernels on sub-ranges: bit-identical to pdes_set_q +
// pdes_eval_residual + pdes_get_res, which is also the fall-back (partitioned meshes, small meshes).
int pdes_eval_residual_host(PdesCtx* ctx, const double* q, double* res, double t) {
  if (!ctx || !q || !res) return usage(ctx, "pdes_eval_residual_host: null argument");
  int rc = finalize(ctx);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  static const bool pipelined = env_int("PDES_HOST_PIPE", 1) != 0;
  if (!ctx->host_chunks || !pipelined) {
    if ((rc = pdes_set_q(ctx, q))) return rc;
    if ((rc = pdes_eval_residual(ctx, t))) return rc;
    return pdes_get_res(ctx, res);
  }
  const int HC = PdesCtx::HC;
  if (!ctx->up_stream) {
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->up_stream, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->down_stream, cudaStreamNonBlocking));
    CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_idle, cudaEventDisableTiming));
    for (int k = 0; k < HC; ++k) {
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_up[k], cudaEventDisableTiming));
      CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_el[k], cudaEventDisableTiming));
    }
  }
  const PdesConfig& c = ctx->cfg;
  const size_t el = (size_t)c.nn * ctx->nd;
  double* qd = ctx->qbuf[ctx->cur];
  cudaEventRecord(ctx->ev_t0, ctx->stream);
  // nothing enqueued earlier may still read the state buffer or write res
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_idle, ctx->stream));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->up_stream, ctx->ev_idle, 0));
  CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->down_stream, ctx->ev_idle, 0));
  for (int k = 0; k < HC; ++k) {
    const size_t o = (size_t)ctx->hc_e[k] * el, n = (size_t)(ctx->hc_e[k + 1] - ctx->hc_e[k]) * el;
    CUDA_TRY(ctx, cudaMemcpyAsync(qd + o, q + o, sizeof(double) * n, cudaMemcpyHostToDevice, ctx->up_stream));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_up[k], ctx->up_stream));
  }
  ElemArgs a;
  fill_args(ctx, &a, qd);
  a.res = ctx->res;
  FaceArgs fa;
  memset(&fa, 0, sizeof(fa));
  fa.q = a.q; fa.faces = ctx->faces; fa.nrm = ctx->nrm_all; fa.coords_bndry = ctx->coords_bndry;
  fa.q_recv = ctx->q_recv; fa.fluxe = ctx->fluxe; fa.ctl = ctx->ctl; fa.ph = a.ph;
  fa.nrm_face_stride = ctx->nrm_compact ? c.dim : c.nfn * c.dim;
  fa.nrm_node_stride = ctx->nrm_compact ? 0 : c.dim;
  fa.prefetch_ahead = ctx->prefetch_ahead_faces;
  fa.ext_bc = ctx->has_ext_bc ? 1 : 0;
  int up_done = -1;
  for (int k = 0; k < HC; ++k) {
    if (ctx->hc_updep[k] > up_done) {
      up_done = ctx->hc_updep[k];
      CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_up[up_done], 0));     // (uploads complete in order)
    }
    fa.g0 = ctx->hc_g[k]; fa.ng = ctx->hc_g[k + 1] - ctx->hc_g[k];
    CUDA_TRY(ctx, ctx->ops->launch_faces(fa, ctx->stream));
    a.e_begin = ctx->hc_e[k]; a.nE = ctx->hc_e[k + 1];
    CUDA_TRY(ctx, ctx->ops->launch_elements(a, EPI_RES, ctx->stream));
    ctx->launches += 2;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_el[k], ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->down_stream, ctx->ev_el[k], 0));
    const size_t o = (size_t)ctx->hc_e[k] * el, n = (size_t)(ctx->hc_e[k + 1] - ctx->hc_e[k]) * el;
    CUDA_TRY(ctx, cudaMemcpyAsync(res + o, ctx->res + o, sizeof(double) * n, cudaMemcpyDeviceToHost, ctx->down_stream));
  }
  ctx->n_evals++;
  cudaEventRecord(ctx->ev_t1, ctx->stream);
  // the compute stream must not run ahead of the copies it shares buffers with
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->up_stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->down_stream));
  rc = pdes_sync(ctx);
  float ms = 0.f;
  if (cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1) == cudaSuccess) ctx->tm.t_func += ms * 1e-3;
  return rc;
}

// evaldRdqProduct / applyLinearOperator (interface2.jl:454-498, newton_setup.jl:632-662): out = dR/dq(q) * v at the
// resident q, without Minv (the reference's physicsRhs residual), v and out in the layout of eqn.q
int pdes_eval_jvp(PdesCtx* ctx, const double* v, double* out) {
  if (!ctx || !v || !out) return usage(ctx, "pdes_eval_jvp: null argument");
  int rc = finalize(ctx);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  double* vdev = ctx->qbuf[(ctx->cur + 1) % 3];     // RK4 scratch buffers double as (v, out) storage
  double* odev = ctx->qbuf[(ctx->cur + 2) % 3];
  CUDA_TRY(ctx, cudaMemcpyAsync(vdev, v, sizeof(double) * ctx->ndof, cudaMemcpyHostToDevice, ctx->stream));
  rc = enqueue_jvp(ctx, vdev, odev);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(out, odev, sizeof(double) * ctx->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

// every step carries what the reference's rk4 does per step (rk4.jl:244-319, 446-457): four stages AND the stage-1
// residual norm (k_norm_reduce / k_norm_commit, at N > 1 the 8-byte all-reduce); the norms are not kept (no buffer)
int pdes_rk4_steps_async(PdesCtx* ctx, double h, int64_t nsteps) {
  if (!ctx) return usage(ctx, "null ctx");
  int rc = finalize(ctx);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  static const bool with_norm = env_int("PDES_STEPS_NO_NORM", 0) == 0;
  for (int64_t i = 0; i < nsteps; ++i) {
    rc = launch_rk4_step(ctx, h, with_norm, -1.0, 0);
    if (rc) return rc;
  }
  return PDES_OK;
}

int pdes_rk4(PdesCtx* ctx, double h, double t_max, int64_t itermax, double res_tol, int32_t real_time, double* t_out,
             double* norms_out, int64_t norms_cap, int64_t* nsteps_out) {
  if (!ctx) return usage(ctx, "null ctx");
  if (!(h > 0.0)) return usage(ctx, "pdes_rk4: h must be positive");
  int rc = finalize(ctx);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const int64_t t_steps = (int64_t)llround(t_max / h);          // rk4.jl:170
  int64_t max_heads = t_steps;
  if (itermax >= 0 && itermax < max_heads) max_heads = itermax > 0 ? itermax : 1;
  if (max_heads < 0) max_heads = 0;
  if (ctx->norms_cap < max_heads + 1) {
    if (ctx->norms_dev) cudaFree(ctx->norms_dev);
    ctx->norms_dev = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->norms_dev, sizeof(double) * (size_t)(max_heads + 1)));
    ctx->norms_cap = max_heads + 1;
    ctx->g_h = -1.0;      // captured graphs hold the old norms pointer
  }
  rc = reset_ctl(ctx);
  if (rc) return rc;
  const int pseudo = real_time ? 0 : 1;
  const int cur0 = ctx->cur;
  int64_t heads = 0;       // executed step heads (stage 1 + norm)
  int64_t full = 0;        // completed full steps
  double t = 0.0;
  int status = PDES_OK;

  const int64_t poll = 32;
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_t0, ctx->stream));
  for (int64_t i = 2; i <= t_steps + 1; ++i) {
    t = (double)(i - 2) * h;
    const bool head_only = (itermax >= 0 && i > itermax);       // rk4.jl:269-276, after stage 1
    rc = head_only ? enqueue_rk4_step(ctx, h, true, res_tol, pseudo, true) : launch_rk4_step(ctx, h, true, res_tol, pseudo);
    if (rc) return rc;
    ++heads;
    if (head_only) break;
    ++full;
    if ((full % poll) == 0 || i == t_steps + 1) {
      status = fetch_ctl(ctx);
      if (status || ctx->h_ctl->stop) break;
    }
  }
  CUDA_TRY(ctx, cudaEventRecord(ctx->ev_t1, ctx->stream));
  if (!status) status = fetch_ctl(ctx);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, ctx->ev_t0, ctx->ev_t1);
  ctx->tm.t_timemarch += ms * 1e-3;

  if (status == PDES_OK && ctx->h_ctl->stop && ctx->h_ctl->converged_step >= 0) {
    // norm < res_tol at step head c: the reference breaks right after stage 1 (rk4.jl:258-267); every later
    // kernel was a no-op, so the state is x_old + (h/2) k1 in buffer B of that step.  This also holds when the
    // itermax head-only exit was enqueued after c (solver/common.jl:543 passes res_tol together with use_itermax):
    // for c == heads-1 the correction reproduces the head-only bookkeeping.
    const int64_t c = ctx->h_ctl->converged_step;
    heads = c + 1;
    ctx->cur = (int)((cur0 + 2 * c + 1) % 3);
    t = (double)c * h;
  }
  t += h;                                                        // rk4.jl:323
  if (t_out) *t_out = t;
  if (nsteps_out) *nsteps_out = heads;
  if (norms_out && heads > 0) {
    int64_t n = heads < norms_cap ? heads : norms_cap;
    if (n > 0) CUDA_TRY(ctx, cudaMemcpy(norms_out, ctx->norms_dev, sizeof(double) * n, cudaMemcpyDeviceToHost));
  }
  if (ctx->h_ctl->stop) reset_ctl(ctx);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return status;
}

int pdes_lserk54(PdesCtx* ctx, double h, double t_max, int64_t itermax, double res_tol, int32_t real_time, double* t_out,
                 double* norms_out, int64_t norms_cap, int64_t* nsteps_out) {
  if (!ctx) return usage(ctx, "null ctx");
  if (!(h > 0.0)) return usage(ctx, "pdes_lserk54: h must be positive");
  int rc = finalize(ctx);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  const int64_t t_steps = (int64_t)llround(t_max / h);
  int64_t max_heads = t_steps;
  if (itermax >= 0 && itermax < max_heads) max_heads = itermax > 0 ? itermax : 1;
  if (max_heads < 0) max_heads = 0;
  if (ctx->norms_cap < max_heads + 1) {
    if (ctx->norms_dev) cudaFree(ctx->norms_dev);
    ctx->norms_dev = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->norms_dev, sizeof(double) * (size_t)(max_heads + 1)));
    ctx->norms_cap = max_heads + 1;
    ctx->g_h = -1.0;
  }
  rc = reset_ctl(ctx);
  if (rc) return rc;
  const int pseudo = real_time ? 0 : 1;
  const int cur0 = ctx->cur;
  int64_t heads = 0, full = 0;
  double t = 0.0;
  int status = PDES_OK;
  for (int64_t i = 2; i <= t_steps + 1; ++i) {
    t = (double)(i - 2) * h;
    const bool head_only = (itermax >= 0 && i > itermax);
    rc = enqueue_lserk_step(ctx, h, res_tol, pseudo, head_only);
    if (rc) return rc;
    ++heads;
    if (head_only) break;
    ++full;
    if ((full % 32) == 0 || i == t_steps + 1) {
      status = fetch_ctl(ctx);
      if (status || ctx->h_ctl->stop) break;
    }
  }
  if (!status) status = fetch_ctl(ctx);
  if (status == PDES_OK && ctx->h_ctl->stop && ctx->h_ctl->converged_step >= 0) {
    // norm < res_tol at step head c: q of that step head is untouched (buffer P0 of step c)
    const int64_t c = ctx->h_ctl->converged_step;
    heads = c + 1;
    ctx->cur = (int)((cur0 + c) % 3);
    t = (double)c * h;
  }
  t += h;
  if (t_out) *t_out = t;
  if (nsteps_out) *nsteps_out = heads;
  if (norms_out && heads > 0) {
    int64_t n = heads < norms_cap ? heads : norms_cap;
    if (n > 0) CUDA_TRY(ctx, cudaMemcpy(norms_out, ctx->norms_dev, sizeof(double) * n, cudaMemcpyDeviceToHost));
  }
  if (ctx->h_ctl->stop) reset_ctl(ctx);
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return status;
}

// ---- matrix-free Newton-Krylov (configuration 5) -------------------------------------------------------------
}  // extern "C"

namespace {

int kry_alloc(PdesCtx* ctx, int restart) {
  PdesCtx::Krylov& k = ctx->kry;
  if (k.restart >= restart && k.V) return PDES_OK;
  void* old[] = {k.V, k.w, k.b, k.x, k.partials, k.hdev, k.gH, k.gcs, k.gsn, k.gg, k.gstate};
  for (void* p : old) if (p) cudaFree(p);
  if (k.hhost) cudaFreeHost(k.hhost);
  if (k.hstate) cudaFreeHost(k.hstate);
  for (cudaGraphExec_t ge : k.itg) if (ge) cudaGraphExecDestroy(ge);     // (they hold the old buffers)
  k.itg.clear();
  {
    // (the preconditioner's buffers do not depend on the restart length)
    PdesCtx::Krylov fresh;
    fresh.pc_type = k.pc_type; fresh.ncolours = k.ncolours; fresh.pc_ready = k.pc_ready;
    fresh.pc_blocks = k.pc_blocks; fresh.pcz = k.pcz; fresh.pcu = k.pcu; fresh.colour = k.colour;
    k = fresh;
  }
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n = ctx->ndof;
  k.nblk = (int)std::max<int64_t>(1, std::min<int64_t>((n + KRY_T - 1) / KRY_T, (int64_t)sms * 4));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.V, sizeof(double) * (size_t)(restart + 1) * n));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.w, sizeof(double) * n));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.b, sizeof(double) * n));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.x, sizeof(double) * n));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.partials, sizeof(double) * (size_t)(restart + 2) * k.nblk));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.hdev, sizeof(double) * 3 * (size_t)(restart + 2)));
  CUDA_TRY(ctx, cudaMallocHost((void**)&k.hhost, sizeof(double) * 3 * (size_t)(restart + 2)));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.gH, sizeof(double) * (size_t)(restart + 1) * restart));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.gcs, sizeof(double) * restart));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.gsn, sizeof(double) * restart));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.gg, sizeof(double) * (restart + 1)));
  CUDA_TRY(ctx, cudaMalloc((void**)&k.gstate, sizeof(GmresState)));
  CUDA_TRY(ctx, cudaMallocHost((void**)&k.hstate, sizeof(GmresState)));
  k.restart = restart;
  return PDES_OK;
}

// inner products of distributed vectors: the rank-local sums are added over the communicator (the MPI.Allreduce of the
// reference's PETSc / calcNorm reductions); every rank then holds the same coefficients
int kry_allreduce(PdesCtx* ctx, double* dev, int count) {
  if (!ctx->comm || ctx->nranks <= 1) return PDES_OK;
  ncclResult_t r = g_nccl.AllReduce(dev, dev, (size_t)count, ncclFloat64, ncclSum, ctx->comm, ctx->stream);
  if (r != ncclSuccess) { set_err(ctx, "ncclAllReduce failed: %s", g_nccl.GetErrorString(r)); return PDES_ERR_COMM; }
  return PDES_OK;
}

// out[0..nv) = V[0..nv)^T w  (device results)
int kry_dots(PdesCtx* ctx, const double* V, int nv, const double* w, double* out, const int* done = nullptr) {
  PdesCtx::Krylov& k = ctx->kry;
  dim3 grid(k.nblk, (nv + KRY_VB - 1) / KRY_VB);
  k_multi_dot<<<grid, KRY_T, 0, ctx->stream>>>(V, ctx->ndof, nv, w, ctx->ndof, k.partials, done);
  k_reduce_rows<<<nv, KRY_T, 0, ctx->stream>>>(k.partials, k.nblk, out);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches += 2;
  return kry_allreduce(ctx, out, nv);
}

int kry_fetch(PdesCtx* ctx, int count) {
  PdesCtx::Krylov& k = ctx->kry;
  CUDA_TRY(ctx, cudaMemcpyAsync(k.hhost, k.hdev, sizeof(double) * count, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

// Element-block Jacobi preconditioner of the current state (see krylov_kernels.cuh): colours, probing products, inversion.
int build_block_pc(PdesCtx* ctx) {
  PdesCtx::Krylov& k = ctx->kry;
  const PdesConfig& c = ctx->cfg;
  const int EL = c.nn * ctx->nd;
  const int64_t nE = c.nE, n = ctx->ndof;
  const size_t smem = sizeof(double) * (size_t)EL * EL + sizeof(int) * (size_t)EL;
  if (smem > 200 * 1024) { set_err(ctx, "element-block preconditioner: %d dofs per element exceed the shared-memory inversion", EL); return PDES_ERR_UNSUPPORTED; }
  if (!k.colour) {
    // greedy colouring of the face-adjacency graph (<= dim + 2 colours on a simplex mesh)
    std::vector<std::vector<int32_t>> adj((size_t)nE);
    for (const FaceRec& f : ctx->h_faces)
      if (f.kind == FK_INTERIOR) { adj[f.elL].push_back(f.elR); adj[f.elR].push_back(f.elL); }
    std::vector<int32_t> col((size_t)nE, -1);
    int nc = 0;
    for (int64_t e = 0; e < nE; ++e) {
      unsigned used = 0;
      for (int32_t o : adj[e]) if (col[o] >= 0) used |= 1u << col[o];
      int cc = 0;
      while (used & (1u << cc)) ++cc;
      col[e] = cc;
      nc = std::max(nc, cc + 1);
    }
    k.ncolours = nc;
    CUDA_TRY(ctx, dev_upload(ctx->stream, &k.colour, col.data(), col.size()));
    CUDA_TRY(ctx, cudaMalloc((void**)&k.pc_blocks, sizeof(double) * (size_t)nE * EL * EL));
    CUDA_TRY(ctx, cudaMalloc((void**)&k.pcz, sizeof(double) * n));
    CUDA_TRY(ctx, cudaMalloc((void**)&k.pcu, sizeof(double) * n));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_block_invert, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  }
  const int nb = k.nblk;
  bool first = true;
  for (int col = 0; col < k.ncolours; ++col)
    for (int cdof = 0; cdof < EL; ++cdof) {
      k_probe_set<<<nb, KRY_T, 0, ctx->stream>>>(k.pcz, k.colour, col, cdof, EL, nE);
      int rc = enqueue_jvp(ctx, k.pcz, k.pcu, first ? 1 : 2);
      if (rc) return rc;
      first = false;
      k_probe_get<<<nb, KRY_T, 0, ctx->stream>>>(k.pcu, k.colour, col, cdof, EL, nE, k.pc_blocks);
      ctx->launches += 2;
    }
  k_block_invert<<<(unsigned)nE, 64, smem, ctx->stream>>>(k.pc_blocks, EL, nE);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches++;
  k.pc_ready = true;
  return PDES_OK;
}

// Restarted GMRES on device vectors: solves dR/dq(q) x = b, x0 = 0; classical Gram-Schmidt applied twice (CGS2): two
// batched dot kernels instead of j+1 dependent ones.  Convergence as PETSc's default test: rnorm <= max(reltol*|b|,
// abstol); divergence when rnorm >= dtol*|b|.  reason: 1 rtol, 2 abstol, 3 exact breakdown, -1 itermax, -2 dtol.
int gmres_dev(PdesCtx* ctx, const double* b, double* x, double reltol, double abstol, double dtol, int64_t itermax,
              int restart, int64_t* iters_out, double* rnorm_out, int* reason_out) {
  int rc = kry_alloc(ctx, restart);
  if (rc) return rc;
  PdesCtx::Krylov& k = ctx->kry;
  const int64_t n = ctx->ndof;
  const int m = restart, S = restart + 2;
  // right preconditioning (-ksp_pc_side right): J M^-1 y = b, x = M^-1 y; the residual norms are those of the true system
  const bool pc = k.pc_type == 1;
  if (pc && !k.pc_ready) { rc = build_block_pc(ctx); if (rc) return rc; }
  const int EL = ctx->cfg.nn * ctx->nd;
  double* h1 = k.hdev;           // first projection coefficients
  double* h2 = k.hdev + S;       // second pass
  double* nq = k.hdev + 2 * S;   // squared norms
  const int nb = k.nblk;
  cudaStream_t st = ctx->stream;
  std::vector<double> H((size_t)(m + 1) * m, 0.0), g(m + 1), y(m);
  int* done = &ctx->ctl->kry_done;
  // CUDA graph per iteration (one GPU: the partitioned J*v carries NCCL calls); cache valid for this (restart, pc, buffer)
  const bool use_graph = !ctx->no_graph && !ctx->comm && ctx->nS == 0;
  if (use_graph && ((int)k.itg.size() != m || k.itg_pc != (pc ? 1 : 0) || k.itg_cur != ctx->cur)) {
    for (cudaGraphExec_t ge : k.itg) if (ge) cudaGraphExecDestroy(ge);
    k.itg.assign(m, nullptr);
    k.itg_pc = pc ? 1 : 0; k.itg_cur = ctx->cur;
    // one product outside any capture: one-time set-up of the J*v kernels (table uploads, function attributes) must not
    // happen while a stream is capturing
    rc = enqueue_jvp(ctx, b, k.w);
    if (rc) return rc;
  }
  // iterations enqueued per poll of the device-side solver state (PDES_GMRES_POLL; 1 = a read-back per iteration)
  static const int poll = std::max(1, env_int("PDES_GMRES_POLL", 8));
  CUDA_TRY(ctx, cudaMemsetAsync(x, 0, sizeof(double) * n, st));
  CUDA_TRY(ctx, cudaMemsetAsync(done, 0, sizeof(int), st));
  rc = kry_dots(ctx, b, 1, b, nq);
  if (rc) return rc;
  rc = kry_fetch(ctx, 3 * S);
  if (rc) return rc;
  const double bnorm = sqrt(k.hhost[2 * S]);
  int64_t its = 0;
  int reason = 0;
  double rnorm = bnorm;
  if (bnorm == 0.0) reason = 3;
  const double tol = std::max(reltol * bnorm, abstol);
  bool first = true;
  while (!reason) {
    // r = b - J x  (x = 0 in the first cycle)
    if (first) {
      k_axpby<<<nb, KRY_T, 0, st>>>(1.0, b, 0.0, k.w, n);
    } else {
      rc = enqueue_jvp(ctx, x, k.w);
      if (rc) return rc;
      k_axpby<<<nb, KRY_T, 0, st>>>(1.0, b, -1.0, k.w, n);
    }
    ctx->launches++;
    rc = kry_dots(ctx, k.w, 1, k.w, nq);
    if (rc) return rc;
    k_normalize<<<nb, KRY_T, 0, st>>>(k.w, nq, k.V, n);
    ctx->launches++;
    rc = kry_fetch(ctx, 3 * S);
    if (rc) return rc;
    const double beta = sqrt(k.hhost[2 * S]);
    rnorm = beta;
    if (!first && beta <= tol) { reason = beta <= reltol * bnorm ? 1 : 2; break; }
    first = false;
    // the cycle's Hessenberg system lives on the device: g = beta e_1, state = {its so far, no reason, no column}
    {
      std::fill(g.begin(), g.end(), 0.0);
      g[0] = beta;
      GmresState hs;
      hs.rnorm = beta; hs.bnorm = bnorm; hs.its = its; hs.itermax = itermax; hs.reason = 0; hs.jdone = 0;
      hs.reltol = reltol; hs.abstol = abstol; hs.dtol = dtol;
      *k.hstate = hs;
      CUDA_TRY(ctx, cudaMemcpyAsync(k.gstate, k.hstate, sizeof(GmresState), cudaMemcpyHostToDevice, st));
      CUDA_TRY(ctx, cudaMemcpyAsync(k.gg, g.data(), sizeof(double) * (m + 1), cudaMemcpyHostToDevice, st));
      CUDA_TRY(ctx, cudaStreamSynchronize(st));       // g and hstate are host temporaries
    }
    int j = 0;
    while (j < m && !reason) {
      const int jb = std::min(poll, m - j);
      for (int u = 0; u < jb; ++u) {
        const int jj = j + u;
        auto enqueue_iteration = [&]() -> int {
          int rci;
          if (pc) {
            k_block_apply<<<nb, KRY_T, 0, st>>>(k.pc_blocks, EL, ctx->cfg.nE, k.V + (size_t)jj * n, k.pcz, done);
            ctx->launches++;
          }
          rci = enqueue_jvp(ctx, pc ? k.pcz : k.V + (size_t)jj * n, k.w);
          if (rci) return rci;
          rci = kry_dots(ctx, k.V, jj + 1, k.w, h1, done);
          if (rci) return rci;
          k_multi_axpy<<<nb, KRY_T, 0, st>>>(k.V, n, jj + 1, h1, -1.0, k.w, n, done);
          rci = kry_dots(ctx, k.V, jj + 1, k.w, h2, done);
          if (rci) return rci;
          k_multi_axpy<<<nb, KRY_T, 0, st>>>(k.V, n, jj + 1, h2, -1.0, k.w, n, done);
          rci = kry_dots(ctx, k.w, 1, k.w, nq, done);
          if (rci) return rci;
          k_normalize<<<nb, KRY_T, 0, st>>>(k.w, nq, k.V + (size_t)(jj + 1) * n, n, done);
          // Hessenberg column, rotations, residual norm and the stopping tests: on the device, no read-back
          k_gmres_update<<<1, 1, 0, st>>>(jj, m, h1, h2, nq, k.gH, k.gcs, k.gsn, k.gg, k.gstate, done);
          if (cudaGetLastError() != cudaSuccess) return PDES_ERR_CUDA;
          ctx->launches += 4;
          return PDES_OK;
        };
        if (!use_graph) {
          rc = enqueue_iteration();
          if (rc) { if (rc == PDES_ERR_CUDA) set_err(ctx, "GMRES iteration launch failed"); return rc; }
          continue;
        }
        if (!k.itg[jj]) {
          const int64_t l0 = ctx->launches;
          cudaGraph_t gr = nullptr;
          CUDA_TRY(ctx, cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
          rc = enqueue_iteration();
          cudaError_t ce = cudaStreamEndCapture(st, &gr);
          ctx->launches = l0;
          if (rc || ce != cudaSuccess) { if (gr) cudaGraphDestroy(gr); if (!rc) { CUDA_TRY(ctx, ce); } return rc; }
          CUDA_TRY(ctx, cudaGraphInstantiate(&k.itg[jj], gr, 0));
          cudaGraphDestroy(gr);
        }
        CUDA_TRY(ctx, cudaGraphLaunch(k.itg[jj], st));
        ctx->launches += pc ? 13 : 12;
      }
      CUDA_TRY(ctx, cudaMemcpyAsync(k.hstate, k.gstate, sizeof(GmresState), cudaMemcpyDeviceToHost, st));
      CUDA_TRY(ctx, cudaStreamSynchronize(st));
      j = k.hstate->jdone;
      its = k.hstate->its;
      rnorm = k.hstate->rnorm;
      reason = k.hstate->reason;
    }
    // x += V y with H y = g (back substitution over the j columns built in this cycle)
    const int jj = j;
    CUDA_TRY(ctx, cudaMemsetAsync(done, 0, sizeof(int), st));        // the update kernels below must run
    CUDA_TRY(ctx, cudaMemcpyAsync(H.data(), k.gH, sizeof(double) * (size_t)(m + 1) * m, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaMemcpyAsync(g.data(), k.gg, sizeof(double) * (m + 1), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(ctx, cudaStreamSynchronize(st));
    for (int i = jj - 1; i >= 0; --i) {
      double sacc = g[i];
      for (int c2 = i + 1; c2 < jj; ++c2) sacc -= H[(size_t)i * m + c2] * y[c2];
      y[i] = H[(size_t)i * m + i] != 0.0 ? sacc / H[(size_t)i * m + i] : 0.0;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(h1, y.data(), sizeof(double) * jj, cudaMemcpyHostToDevice, st));
    if (pc) {
      // x += M^-1 (V y)
      CUDA_TRY(ctx, cudaMemsetAsync(k.pcu, 0, sizeof(double) * n, st));
      k_multi_axpy<<<nb, KRY_T, 0, st>>>(k.V, n, jj, h1, 1.0, k.pcu, n);
      k_block_apply<<<nb, KRY_T, 0, st>>>(k.pc_blocks, EL, ctx->cfg.nE, k.pcu, k.pcz);
      k_axpby<<<nb, KRY_T, 0, st>>>(1.0, k.pcz, 1.0, x, n);
      ctx->launches += 2;
    } else
    k_multi_axpy<<<nb, KRY_T, 0, st>>>(k.V, n, jj, h1, 1.0, x, n);
    CUDA_TRY(ctx, cudaGetLastError());
    CUDA_TRY(ctx, cudaStreamSynchronize(st));     // y is a host temporary
    ctx->launches++;
  }
  *iters_out = its;
  *rnorm_out = rnorm;
  *reason_out = reason;
  return PDES_OK;
}

// physicsRhs (jacobian/residual_evaluation.jl:64-88): res = R(q) on the device, returns calcNorm(strongres=true)
int newton_rhs(PdesCtx* ctx, double* norm_out) {
  int rc = pdes_eval_residual_async(ctx, 0.0);
  if (rc) return rc;
  PdesCtx::Krylov& k = ctx->kry;
  k_strong_norm_partials<<<k.nblk, KRY_T, 0, ctx->stream>>>(ctx->res, ctx->minv, ctx->nd, ctx->ndof, k.partials);
  k_reduce_rows<<<1, KRY_T, 0, ctx->stream>>>(k.partials, k.nblk, k.hdev);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches += 2;
  rc = kry_allreduce(ctx, k.hdev, 1);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(k.hhost, k.hdev, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  rc = pdes_sync(ctx);                     // also turns a negative density / pressure into the reference's exception
  if (rc) return rc;
  *norm_out = sqrt(k.hhost[0]);
  return PDES_OK;
}

}  // namespace

extern "C" {

int pdes_set_krylov_pc(PdesCtx* ctx, int32_t pc_type) {
  if (!ctx) return usage(ctx, "pdes_set_krylov_pc: null ctx");
  if (pc_type != PDES_PC_NONE && pc_type != PDES_PC_ELEMENT_BLOCK_JACOBI) return usage(ctx, "pdes_set_krylov_pc: unknown preconditioner");
  ctx->kry.pc_type = pc_type;
  ctx->kry.pc_ready = false;
  return PDES_OK;
}

int pdes_gmres(PdesCtx* ctx, const double* b, double* x, double reltol, double abstol, double dtol, int64_t itermax,
               int32_t restart, int64_t* iters_out, double* rnorm_out, int32_t* reason_out) {
  if (!ctx || !b || !x || !iters_out || !rnorm_out || !reason_out) return usage(ctx, "pdes_gmres: null argument");
  if (restart < 1 || itermax < 1) return usage(ctx, "pdes_gmres: restart and itermax must be positive");
  int rc = finalize(ctx);
  if (rc) return rc;
  if (ctx->nS > 0 && !ctx->comm) { set_err(ctx, "pdes_gmres on a partitioned mesh needs a communicator (pdes_set_comm)"); return PDES_ERR_UNSUPPORTED; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  rc = kry_alloc(ctx, restart);
  if (rc) return rc;
  CUDA_TRY(ctx, cudaMemcpyAsync(ctx->kry.b, b, sizeof(double) * ctx->ndof, cudaMemcpyHostToDevice, ctx->stream));
  int reason = 0;
  ctx->kry.pc_ready = false;          // the blocks belong to the state of this call
  rc = gmres_dev(ctx, ctx->kry.b, ctx->kry.x, reltol, abstol, dtol, itermax, restart, iters_out, rnorm_out, &reason);
  if (rc) return rc;
  *reason_out = reason;
  CUDA_TRY(ctx, cudaMemcpyAsync(x, ctx->kry.x, sizeof(double) * ctx->ndof, cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
  return PDES_OK;
}

int pdes_newton_krylov(PdesCtx* ctx, const PdesNewtonOpts* o, double* res_norms_out, double* step_norms_out,
                       PdesNewtonResult* result) {
  if (!ctx || !o || !result) return usage(ctx, "pdes_newton_krylov: null argument");
  if (o->krylov_restart < 1 || o->krylov_itermax < 1 || o->itermax < 0) return usage(ctx, "pdes_newton_krylov: bad options");
  int rc = finalize(ctx);
  if (rc) return rc;
  if (ctx->nS > 0 && !ctx->comm) { set_err(ctx, "pdes_newton_krylov on a partitioned mesh needs a communicator (pdes_set_comm)"); return PDES_ERR_UNSUPPORTED; }
  CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
  rc = kry_alloc(ctx, o->krylov_restart);
  if (rc) return rc;
  PdesCtx::Krylov& k = ctx->kry;
  const int64_t n = ctx->ndof;
  memset(result, 0, sizeof(*result));
  double res_norm = 0.0, step_norm = 0.0;
  rc = newton_rhs(ctx, &res_norm);
  if (rc) return rc;
  result->residual_evals = 1;
  if (res_norms_out) res_norms_out[0] = res_norm;
  const double res_norm_rel = res_norm;                      // set_rel_norm: the first residual (newton.jl:198-200)
  auto converged = [&](int64_t itr) {                        // checkConvergence (newton.jl:402-445)
    bool c = res_norm < o->res_abstol;
    if (res_norm / res_norm_rel < o->res_reltol) c = true;
    if (step_norm <= o->step_tol && itr > 0) c = true;
    return c;
  };
  bool conv = converged(0);
  int64_t itr = 0;
  while (!conv && itr < o->itermax) {
    ++itr;
    // delta_q: J delta_q = -res
    k_axpby<<<k.nblk, KRY_T, 0, ctx->stream>>>(-1.0, ctx->res, 0.0, k.b, n);
    ctx->launches++;
    int64_t kits = 0;
    double krn = 0.0;
    int reason = 0;
    k.pc_ready = false;               // recalculated for every Newton iterate (recalc policy "always")
    rc = gmres_dev(ctx, k.b, k.x, o->krylov_reltol, o->krylov_abstol, o->krylov_dtol, o->krylov_itermax,
                   o->krylov_restart, &kits, &krn, &reason);
    if (rc) return rc;
    result->krylov_iters += kits;
    result->krylov_reason = reason;
    // step_norm = norm(delta_q_vec) (plain 2-norm, newton.jl:237), q += step_fac * delta_q
    rc = kry_dots(ctx, k.x, 1, k.x, k.hdev);
    if (rc) return rc;
    k_axpby<<<k.nblk, KRY_T, 0, ctx->stream>>>(o->step_fac, k.x, 1.0, ctx->qbuf[ctx->cur], n);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches++;
    rc = kry_fetch(ctx, 1);
    if (rc) return rc;
    step_norm = sqrt(k.hhost[0]);
    if (step_norms_out) step_norms_out[itr - 1] = step_norm;
    rc = newton_rhs(ctx, &res_norm);
    if (rc) return rc;
    result->residual_evals++;
    if (res_norms_out) res_norms_out[itr] = res_norm;
    conv = converged(itr);
  }
  result->converged = conv ? 1 : 0;
  result->newton_iters = itr;
  result->res_norm = res_norm;
  result->res_norm_rel = res_norm_rel;
  result->step_norm = step_norm;
  return PDES_OK;
}

// majorIterationCallback functionals (euler.jl:330-407) of the resident q: evaluates R(q) and reduces on the device.
// out[0..5+nd): entropy integral, w^T R, kinetic energy, d(kinetic energy)/dt, volume, integral of q[0..nd)
int pdes_diagnostics(PdesCtx* ctx, double* out) {
  if (!ctx || !out) return usage(ctx, "pdes_diagnostics: null argument");
  int rc = pdes_eval_residual_async(ctx, 0.0);
  if (rc) return rc;
  const int nv = 5 + ctx->nd;
  const int64_t n_nodes = (int64_t)ctx->cfg.nn * ctx->cfg.nE;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int B = (int)std::max<int64_t>(1, std::min<int64_t>((n_nodes + DIAG_T - 1) / DIAG_T, (int64_t)sms * 4));
  if (!ctx->diag_buf || ctx->diag_B != B) {
    if (ctx->diag_buf) cudaFree(ctx->diag_buf);
    ctx->diag_buf = nullptr;
    CUDA_TRY(ctx, cudaMalloc((void**)&ctx->diag_buf, sizeof(double) * (size_t)(nv + 1) * (B + 1)));
    ctx->diag_B = B;
  }
  double* partials = ctx->diag_buf;
  double* sums = ctx->diag_buf + (size_t)(nv + 1) * B;        // [nv + 1]: the last entry is the enstrophy
  if (ctx->cfg.dim == 2)
    k_diag_partials<2><<<B, DIAG_T, 0, ctx->stream>>>(ctx->qbuf[ctx->cur], ctx->res, ctx->mass, n_nodes, ctx->cfg.gamma, partials);
  else
    k_diag_partials<3><<<B, DIAG_T, 0, ctx->stream>>>(ctx->qbuf[ctx->cur], ctx->res, ctx->mass, n_nodes, ctx->cfg.gamma, partials);
  if (ctx->cfg.dim == 3) {
    const int dd = 9;
    k_enstrophy_partials<<<B, DIAG_T, 0, ctx->stream>>>(ctx->qbuf[ctx->cur], ctx->Q_dev, ctx->w_dev, ctx->dxidx,
                                                        ctx->dx_compact ? dd : (int64_t)ctx->cfg.nn * dd, ctx->mass, ctx->cfg.nn,
                                                        ctx->cfg.nE, partials + (size_t)nv * B);
    ctx->launches++;
  } else {
    CUDA_TRY(ctx, cudaMemsetAsync(partials + (size_t)nv * B, 0, sizeof(double) * B, ctx->stream));
  }
  k_reduce_rows<<<nv + 1, KRY_T, 0, ctx->stream>>>(partials, B, sums);
  CUDA_TRY(ctx, cudaGetLastError());
  ctx->launches += 2;
  CUDA_TRY(ctx, cudaMemcpyAsync(out, sums, sizeof(double) * (nv + 1), cudaMemcpyDeviceToHost, ctx->stream));
  rc = pdes_sync(ctx);
  if (rc) return rc;
  const double volume = out[4];
  out[2] = 0.5 * out[2] / volume;
  out[3] = out[3] / volume;
  out[nv] = 0.5 * out[nv] / volume;
  return PDES_OK;
}

int pdes_get_minv(PdesCtx* ctx, double* Minv) {
  if (!ctx || !Minv) return usage(ctx, "pdes_get_minv: null argument");
  if (!ctx->have_mesh) return usage(ctx, "pdes_set_mesh must be called first");
  const PdesConfig& c = ctx->cfg;
  std::vector<double> m((size_t)c.nn * c.nE);
  CUDA_TRY(ctx, cudaSetDevice(c.device));
  CUDA_TRY(ctx, cudaMemcpy(m.data(), ctx->minv, sizeof(double) * m.size(), cudaMemcpyDeviceToHost));
  for (int64_t e = 0; e < c.nE; ++e)
    for (int j = 0; j < c.nn; ++j)
      for (int k = 0; k < ctx->nd; ++k) Minv[k + ctx->nd * (j + (int64_t)c.nn * e)] = m[j + (size_t)c.nn * e];
  return PDES_OK;
}

int pdes_get_timings(PdesCtx* ctx, PdesTimings* out) {
  if (!ctx || !out) return usage(ctx, "pdes_get_timings: null argument");
  *out = ctx->tm;
  out->n_residual_evals = ctx->n_evals;
  out->n_kernel_launches = ctx->launches;
  return PDES_OK;
}

int64_t pdes_kernel_launch_count(const PdesCtx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"
