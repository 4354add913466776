// Host side of libcalamity_b200: plan construction, uploads, the fit loop and the C ABI of
// include/calamity_b200.h.  No torch types, no exceptions across the boundary.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>  // header-only: the ranges cost nothing unless a profiler injects itself

#include <algorithm>
#include <functional>
#include <queue>
#include <climits>
#include <cstdarg>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/calamity_b200.h"
#include "calfit_kernels.cuh"
#include "calfit_shared.cuh"
#include "calfit_tc.cuh"
#include "calfit_setup.cuh"
#include "calfit_generic.cuh"

namespace calb2 {

static thread_local std::string g_err;

static int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

static volatile unsigned int* g_tc_dbg = nullptr;  // last plan's tensor-core wait record (development aid)
#define CU(call)                                                                                         \
  do {                                                                                                   \
    cudaError_t e__ = (call);                                                                            \
    if (e__ != cudaSuccess)                                                                              \
      return fail(CALB2_ERR_CUDA, "%s failed at %s:%d: %s [tc wait record: code %u tile %u cta %u thread %u]", #call, __FILE__, __LINE__, \
                  cudaGetErrorString(e__), g_tc_dbg ? g_tc_dbg[0] : 0u, g_tc_dbg ? g_tc_dbg[1] : 0u, g_tc_dbg ? g_tc_dbg[2] : 0u,   \
                  g_tc_dbg ? g_tc_dbg[3] : 0u);                                                            \
  } while (0)

// Guard zones (CALB2_GUARD=1): compute-sanitizer is closed on the GPU pool this library is developed on, so
// out-of-bounds WRITES are caught by the library itself -- every device allocation gets 256 pattern bytes on either
// side, and calb2_debug_check_guards() (tests/test_gpu_guards.py) verifies that no kernel touched them.
static constexpr size_t GUARD_BYTES = 256;
static constexpr unsigned char GUARD_PATTERN = 0xA5;
struct GuardRegistry {
  std::mutex mu;
  std::unordered_map<void*, size_t> live;  // base pointer -> payload bytes
};
static GuardRegistry& guards() {
  static GuardRegistry g;
  return g;
}
static bool guards_enabled() {
  static const bool on = getenv("CALB2_GUARD") && atoi(getenv("CALB2_GUARD")) != 0;
  return on;
}

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  void* base = nullptr;  // != nullptr: allocation with guard zones, p = base + GUARD_BYTES
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    if (!guards_enabled()) return cudaMalloc(&p, count * sizeof(T));
    const size_t payload = (count * sizeof(T) + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&base, payload + 2 * GUARD_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaMemset(base, GUARD_PATTERN, payload + 2 * GUARD_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaDeviceSynchronize();  // the fill runs on the legacy stream; the plan's streams are non-blocking and would race it
    if (e != cudaSuccess) return e;
    p = reinterpret_cast<T*>(static_cast<unsigned char*>(base) + GUARD_BYTES);
    std::lock_guard<std::mutex> lk(guards().mu);
    guards().live[base] = payload;
    return cudaSuccess;
  }
  void release() {
    if (base) {
      {
        std::lock_guard<std::mutex> lk(guards().mu);
        guards().live.erase(base);
      }
      cudaFree(base);
    } else if (p) {
      cudaFree(p);
    }
    p = nullptr;
    base = nullptr;
    n = 0;
  }
  size_t bytes() const { return n * sizeof(T); }
};

// steps per warp: 11 -> 45 KB tiles, 2 CTAs/SM.  -DCALB2_RPT=7 -DCALB2_MINB=3 builds the 3 CTAs/SM experiment.
#ifndef CALB2_RPT
#define CALB2_RPT 11
#define CALB2_MINB 2
#endif
static constexpr int RPT_DEFAULT = CALB2_RPT;
static constexpr int SMAX = 8;

// ---- NCCL through dlopen (the library has no link-time dependency on it) -------------------------
struct IdBlob {
  char bytes[128];
};
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(void*) = nullptr;
  int (*CommInitRank)(void**, int, /*ncclUniqueId by value*/ IdBlob, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int load_nccl(const char* path) {
  if (g_nccl.handle) return 0;
  const char* cands[] = {path, "libnccl.so.2", "libnccl.so"};
  for (const char* c : cands) {
    if (!c || !*c) continue;
    g_nccl.handle = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (g_nccl.handle) break;
  }
  if (!g_nccl.handle) return fail(CALB2_ERR_NCCL, "cannot dlopen NCCL (%s)", dlerror());
  g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))dlsym(g_nccl.handle, "ncclGetUniqueId");
  g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))dlsym(g_nccl.handle, "ncclCommInitRank");
  g_nccl.AllReduce = (decltype(g_nccl.AllReduce))dlsym(g_nccl.handle, "ncclAllReduce");
  g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))dlsym(g_nccl.handle, "ncclCommDestroy");
  g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))dlsym(g_nccl.handle, "ncclGetErrorString");
  g_nccl.GroupStart = (decltype(g_nccl.GroupStart))dlsym(g_nccl.handle, "ncclGroupStart");
  g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))dlsym(g_nccl.handle, "ncclGroupEnd");
  if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.AllReduce)
    return fail(CALB2_ERR_NCCL, "NCCL symbols missing");
  return 0;
}
static constexpr int NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8, NCCL_SUM = 0;

}  // namespace calb2

using namespace calb2;
namespace calb2 {
struct GenericBase;
}

struct calb2_plan {
  calb2::GenericBase* gen = nullptr;  // float64 plans and oversized groups: generic unfused device path
  int dtype = 0;
  int device = 0, nants = 0, nf = 0, ngroups = 0;
  int FL = 0, G = 0, FT = 0, KMAX = 0, nfp = 0, ntiles = 0, RPT = RPT_DEFAULT, NW = 8;
  long long nbls = 0, nslots = 0, ncoef = 0, rows_total = 0, a_floats = 0, n_a_nz = 0;
  // host copies of the description
  std::vector<int> grp_ncomp, grp_nslots, grp_slot0, grp_coef0, slot_nbls, slot_grp, slot_row0, slot_bl0, slot_item,
      bl_ant0, bl_ant1, bl_slot;
  std::vector<ItemDesc> items;
  // shared-basis path (calfit_shared.cuh): single-slot groups whose basis block is shared by enough other groups.
  // Slots are numbered INTERNALLY: streaming-path slots first (canonical order), then the class slots, class by class.
  struct ClsInfo {
    long long a_off;  // float offset of the class's [ntiles][kp][32] swizzled block in A
    int kp, ncomp;
    int nmembers;
    bool uploaded;
    // tensor-core shape (calfit_tc.cuh): classes of <= 128 vectors keep four more copies of every tile (MN-major / K-major,
    // hi / lo) in At
    bool tc;
    int kpt;          // ncomp rounded up to 16
    long long tc_off; // float offset of the class's [tile][4][kpt][32] block in At
  };
  std::vector<ClsInfo> classes;
  std::vector<int> grp_cls;           // dense class index of the group, or -1: streaming path
  // CTAs of the shared-basis kernel: [v][shape], v = 0: plain chi^2 (NQ = 2), 1: 'sum' regulariser (NQ = 4);
  // shape 0: 256 threads, classes of <= 160 vectors, two CTAs per SM; shape 1: 512 threads, <= 208 vectors
  std::vector<MTileDesc> mtiles[2][2];
  DevBuf<MTileDesc> d_mtiles[2][2];
  // tensor-core shape: its CTAs (64 groups of a class of <= 128 vectors) and, when it is in use, the 256-thread CTAs of the
  // remaining small classes (129-160 vectors)
  std::vector<MTileDesc> mt_tc, mt_rest[2][2];  // tensor-core tiles; the classes it does not take, [plain / 'sum'][CUDA-core shape]
  DevBuf<MTileDesc> d_mt_tc, d_mt_rest[2][2];
  DevBuf<float> At;
  bool tc_enabled = false;
  bool tc_sum = true;                 // 'sum' on the tensor cores too (two launches); CALB2_TC_SUM=0: CUDA-core shapes
  size_t tc_smem_bytes = 0;           // dynamic shared memory of the tensor-core launch: the largest class's layout
  unsigned int* tc_dbg = nullptr;     // mapped host memory: record of a tensor-core wait that timed out
  DevBuf<long long> tc_prof;          // CALB2_TC_PROF=<cta>: clock stamps of one CTA of the tensor-core kernel
  int tc_prof_cta = -1;
  int nseg[2] = {1, 1};               // channel segments per class tile = planes of dcpart in use
  int nseg_tc = 1;                    // the same for the tensor-core tiles (chosen by a makespan model, below)
  int dc_mode = -1;                   // which pass wrote dcpart last (planes a pass does not write must be zero)
  long long dc_plane = 0;             // floats per plane of dcpart
  int first_class_row = 0;            // rows below it belong to the streaming path (plane 0 only)
  DevBuf<ClassSlot> d_cslots;
  DevBuf<int> d_cs_slot, d_slot_nb;
  long long nslots_heavy = 0, nslots_class = 0, a_class_floats = 0;
  int ntiles_c = 0;
  bool heavy_single_bl = true, cls_single_bl = true;
  // device
  DevBuf<float> A, d_r, d_i, w, g_r[2], g_i[2], gm_r, gu_r, gm_i, gu_i, gsnap_r, gsnap_i, ggrad_r, ggrad_i;
  DevBuf<float> c_r, c_i, cm_r, cu_r, cm_i, cu_i, csnap_r, csnap_i, cgrad_r, cgrad_i, dcpart, hist, scratch_f;
  DevBuf<float2> z, y, vout;
  DevBuf<double> partials, red_d;
  DevBuf<unsigned long long> dbg_out;
  std::vector<cudaEvent_t> stage_ev;  // -DCALB2_PROFILE + CALB2_DBG&1024: one event per stage boundary per step
  size_t stage_cursor = 0;
  DevBuf<ItemDesc> d_items;
  DevBuf<unsigned char> row_slot;
  DevBuf<int> row_coef, d_slot_row0, d_slot_bl0, d_bl_ant0, d_bl_ant1, d_bl_slot, ant_ptr, ant_ent, coef_row0, coef_grp, ant_partner,
      d_grp_nslots, d_grp_slot0, d_grp_coef0, d_grp_ncomp;
  DevBuf<FitState> state, state_eval;
  DevBuf<SlotGeom> slot_geom;
  DevBuf<float> sky_r, sky_i;
  DevBuf<float> staging;
  float* h_staging = nullptr;
  size_t staging_floats = 0;
  FitState* h_state = nullptr;  // pinned
  cudaStream_t stream = nullptr;
  int cur_buf = 0;
  bool have_data = false, have_gains = false, have_coeffs = false, basis_complete = false, all_single_slot = true, all_single_bl = true;
  int light_blocks = 0;
  DevBuf<double> light_partials;
  long long basis_groups_set = 0;
  // comm
  void* comm = nullptr;
  int rank = 0, nranks = 1;
  DevBuf<double> comm_scalars;
  // peer-memory exchange (calb2_comm_peer_*): own buffer + the peers' buffers opened through cudaIpc
  unsigned char* xbuf = nullptr;
  void* xpeer[CALB2_MAX_RANKS] = {nullptr};
  PeerView peers{};
  unsigned int xseq = 0;  // steps enqueued so far on the exchange (identical on all ranks)
  unsigned long long xtimeout_ns = 20000000000ull;  // bound on every wait for a peer's flag (CALB2_PEER_TIMEOUT_MS)
  bool peers_open = false;
  DevBuf<unsigned int> tail_counter;
  // LAMB (per-variable trust ratios): coefficient ranges of the reference's chunk variables, norm partials, ratios
  std::vector<long long> var_bounds;
  bool var_uploaded = false;
  DevBuf<long long> d_var_bounds;
  DevBuf<double> lamb_gain_partials, lamb_coef_partials;
  DevBuf<float> lamb_ratio;
  // the coefficient update runs on a second stream next to the gain update (they touch disjoint state)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  size_t device_bytes = 0;
};

namespace calb2 {

template <class T>
static int upload(DevBuf<T>& buf, const std::vector<T>& v, calb2_plan* pl) {
  CU(buf.alloc(v.size()));
  pl->device_bytes += buf.bytes();
  if (!v.empty()) CU(cudaMemcpy(buf.p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}
template <class T>
static int dalloc(DevBuf<T>& buf, size_t n, calb2_plan* pl, bool zero = true) {
  CU(buf.alloc(n));
  pl->device_bytes += buf.bytes();
  // stream-ordered: the plan's streams are non-blocking, a fill on the legacy stream would not be ordered with the kernels
  // that use the buffer next (found by the guard-zone run: a late fill overwrote freshly uploaded job descriptors)
  if (zero && n) {
    if (pl->stream)
      CU(cudaMemsetAsync(buf.p, 0, buf.bytes(), pl->stream));
    else {
      CU(cudaMemset(buf.p, 0, buf.bytes()));
      CU(cudaDeviceSynchronize());
    }
  }
  return 0;
}

// Rows the greedy in-order packing would leave unused, as a fraction of the staged tile capacity.
static double packing_fill(const calb2_plan_desc* d, int fl, int RPT, int NWARP) {
  const int G = 32 / fl, kmax = NWARP * RPT * G;
  long long rows = 0, items = 0;
  int cur_rows = 0, cur_slots = 0;
  for (int g = 0; g < d->ngroups; ++g) {
    const int rp = std::max(G, ((d->group_ncomp[g] + G - 1) / G) * G);
    for (int s = 0; s < d->group_nslots[g]; ++s) {
      if (cur_slots > 0 && (cur_rows + rp > kmax || cur_slots >= SMAX)) {
        ++items;
        cur_rows = cur_slots = 0;
      }
      cur_rows += rp;
      cur_slots += 1;
      rows += rp;
    }
  }
  if (cur_slots) ++items;
  return items ? (double)rows / ((double)items * kmax) : 0.0;
}

// Tile width: 64- and 32-channel tiles stream equally well, 16 is measurably worse (64-byte row segments), so
// 16 is only used when a group's basis does not fit otherwise; between 64 and 32 the better-filled packing wins
// (HERA-128: 0.77 vs 0.89 fill -> +10 % throughput at 32; HERA-37: 0.86 vs 0.43 -> 64).
static int choose_fl(const calb2_plan_desc* d, int RPT, int NWARP, int* fl_out) {
  int maxc = 0;
  for (int g = 0; g < d->ngroups; ++g) maxc = std::max(maxc, d->group_ncomp[g]);
  auto fits = [&](int fl) {
    const int G = 32 / fl;
    return ((maxc + G - 1) / G) * G <= NWARP * RPT * G;
  };
  if (d->tile_freqs) {
    const int fl = d->tile_freqs / 4;
    if (d->tile_freqs % 4 || (fl != 16 && fl != 8 && fl != 4)) return fail(CALB2_ERR_ARG, "tile_freqs must be 16, 32 or 64");
    if (!fits(fl))
      return fail(CALB2_ERR_UNSUPPORTED, "a group has %d basis vectors; tile_freqs=%d stages at most %d", maxc,
                  d->tile_freqs, NWARP * RPT * (32 / fl));
    *fl_out = fl;
    return 0;
  }
  if (fits(16)) {
    *fl_out = packing_fill(d, 8, RPT, NWARP) > packing_fill(d, 16, RPT, NWARP) + 0.02 ? 8 : 16;
    return 0;
  }
  if (fits(8)) {
    *fl_out = 8;
    return 0;
  }
  if (fits(4)) {
    *fl_out = 4;
    return 0;
  }
  return fail(CALB2_ERR_UNSUPPORTED, "a group has %d basis vectors; at most %d are supported", maxc, NWARP * RPT * 8);
}

static constexpr int MAX_DEVICES = 64;
static int current_device() {
  int dev = 0;
  cudaGetDevice(&dev);
  return dev >= 0 && dev < MAX_DEVICES ? dev : 0;
}

template <int FL, bool SUM, int QMODE>
static cudaError_t launch_heavy_t(const HeavyParams& hp, int nitems, cudaStream_t s) {
  constexpr int RPT = RPT_DEFAULT, MINB = CALB2_MINB;
  using C = HeavyCfg<FL, SUM, RPT>;
  static bool configured[MAX_DEVICES] = {};  // the attribute is per device (the driver runs one plan per GPU)
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(heavy_kernel<FL, SUM, RPT, MINB, QMODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  heavy_kernel<FL, SUM, RPT, MINB, QMODE><<<nitems, C::NTHR, C::SMEM_BYTES, s>>>(hp);
  return cudaGetLastError();
}

template <int FL>
static cudaError_t launch_heavy_f(bool sum, int qmode, const HeavyParams& hp, int nitems, cudaStream_t s) {
  if (qmode == QM_INIT) return launch_heavy_t<FL, false, QM_INIT>(hp, nitems, s);
  if (qmode == QM_SINGLE)
    return sum ? launch_heavy_t<FL, true, QM_SINGLE>(hp, nitems, s) : launch_heavy_t<FL, false, QM_SINGLE>(hp, nitems, s);
  return sum ? launch_heavy_t<FL, true, QM_GENERAL>(hp, nitems, s) : launch_heavy_t<FL, false, QM_GENERAL>(hp, nitems, s);
}

template <int NTHR, int NQ, bool SINGLE, int KPM>
static cudaError_t launch_shared_t(const SharedParams& sp, int ntiles, cudaStream_t s) {
  using C = SharedCfg<NTHR, NQ, KPM>;
  static bool configured[MAX_DEVICES] = {};
  const int dev = current_device();
  if (!configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(shared_kernel<NTHR, NQ, SINGLE, KPM>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured[dev] = true;
  }
  shared_kernel<NTHR, NQ, SINGLE, KPM><<<ntiles, NTHR, C::SMEM_BYTES, s>>>(sp);
  return cudaGetLastError();
}
static constexpr int KPM_SMALL = 160, KPM_LARGE = 208, SHARED_FT = 32;
static cudaError_t launch_shared_shape(bool sum, bool single, int shape, const SharedParams& sp, int ntiles, cudaStream_t s) {
  if (shape == 0) {
    if (sum) return single ? launch_shared_t<256, 4, true, KPM_SMALL>(sp, ntiles, s) : launch_shared_t<256, 4, false, KPM_SMALL>(sp, ntiles, s);
    return single ? launch_shared_t<256, 2, true, KPM_SMALL>(sp, ntiles, s) : launch_shared_t<256, 2, false, KPM_SMALL>(sp, ntiles, s);
  }
  if (sum) return single ? launch_shared_t<512, 4, true, KPM_LARGE>(sp, ntiles, s) : launch_shared_t<512, 4, false, KPM_LARGE>(sp, ntiles, s);
  return single ? launch_shared_t<512, 2, true, KPM_LARGE>(sp, ntiles, s) : launch_shared_t<512, 2, false, KPM_LARGE>(sp, ntiles, s);
}
// groups per CTA of the four instantiations
static int shared_ms(int v, int shape) { return (shape == 0 ? 64 : 128) / (v == 0 ? 2 : 4); }

// The basis pass of one iteration: the streaming kernel over the items (groups with a private basis) and the
// shared-basis kernel over the class tiles (its large-class shape on the second stream, next to the small-class one);
// all of them write z / dcpart / partials for their own baselines and rows.
static bool tc_pass_of(const calb2_plan* pl, bool sum);
static cudaError_t launch_heavy(const calb2_plan* pl, bool sum, const HeavyParams& hp, int nitems, cudaStream_t s) {
  const int v = sum ? 1 : 0;
  // tensor-core shape: the fit's own passes; initialisation / model-only passes stay on the CUDA cores.  With the 'sum'
  // regulariser the kernel runs twice: as for the plain chi^2 (plus y and the two model sums), then once more without phase F
  // for the two extra backward rows (P w, Q w of calibration.py:1654: they depend on the gains and weights only).
  const bool use_tc = tc_pass_of(pl, sum) && !hp.init_mode && !hp.store_v;
  const int n_tc = use_tc ? (int)pl->mt_tc.size() : 0;
  const MTileDesc* small_tiles = use_tc ? pl->d_mt_rest[v][0].p : pl->d_mtiles[v][0].p;
  const int n_small = use_tc ? (int)pl->mt_rest[v][0].size() : (int)pl->mtiles[v][0].size();
  const MTileDesc* large_tiles = use_tc ? pl->d_mt_rest[v][1].p : pl->d_mtiles[v][1].p;
  const int n_large = use_tc ? (int)pl->mt_rest[v][1].size() : (int)pl->mtiles[v][1].size();
  SharedParams sp{};
  if (n_small + n_large > 0) {
    sp.A = hp.A;
    sp.cslots = pl->d_cslots.p;
    sp.cs_slot = pl->d_cs_slot.p;
    sp.bl_ant0 = hp.bl_ant0;
    sp.bl_ant1 = hp.bl_ant1;
    sp.d_r = hp.d_r;
    sp.d_i = hp.d_i;
    sp.w = hp.w;
    for (int b = 0; b < 2; ++b) {
      sp.g_r[b] = hp.g_r[b];
      sp.g_i[b] = hp.g_i[b];
    }
    sp.c_r = hp.c_r;
    sp.c_i = hp.c_i;
    sp.z = hp.z;
    sp.y = hp.y;
    sp.dcpart = hp.dcpart;
    sp.dc_plane = pl->dc_plane;
    sp.vout = hp.vout;
    sp.st = hp.st;
    sp.nfp = hp.nfp;
    sp.store_v = hp.store_v;
    sp.init_mode = hp.init_mode;
  }
  // the longest-running CTAs first: the large-class shape goes to the second stream before anything else is launched
  const bool fork_large = n_large > 0 && (n_small > 0 || nitems > 0);
  if (n_large > 0) {
    cudaStream_t sl = s;
    if (fork_large) {
      cudaError_t e = cudaEventRecord(pl->ev_fork, s);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(pl->stream2, pl->ev_fork, 0);
      if (e != cudaSuccess) return e;
      sl = pl->stream2;
    }
    sp.tiles = large_tiles;
    sp.partials = hp.partials + (size_t)(nitems + n_small) * 4;
    cudaError_t e = launch_shared_shape(sum, pl->cls_single_bl, 1, sp, n_large, sl);
    if (e == cudaSuccess && fork_large) e = cudaEventRecord(pl->ev_join, pl->stream2);
    if (e != cudaSuccess) return e;
  }
  // with the tensor-core kernel taking (nearly) every class, the few streaming items left run NEXT to it on the second stream
  // (33 us of a 640 us pass at HERA-350 when they ran in front of it)
  const bool fork_items = nitems > 0 && n_tc > 0 && !fork_large;
  if (nitems > 0) {
    cudaStream_t si = s;
    if (fork_items) {
      cudaError_t e = cudaEventRecord(pl->ev_fork, s);
      if (e == cudaSuccess) e = cudaStreamWaitEvent(pl->stream2, pl->ev_fork, 0);
      if (e != cudaSuccess) return e;
      si = pl->stream2;
    }
    const int qmode = hp.init_mode ? QM_INIT : (pl->heavy_single_bl ? QM_SINGLE : QM_GENERAL);
    cudaError_t e;
    switch (pl->FL) {
      case 16: e = launch_heavy_f<16>(sum, qmode, hp, nitems, si); break;
      case 8: e = launch_heavy_f<8>(sum, qmode, hp, nitems, si); break;
      default: e = launch_heavy_f<4>(sum, qmode, hp, nitems, si); break;
    }
    if (e == cudaSuccess && fork_items) e = cudaEventRecord(pl->ev_join, pl->stream2);
    if (e != cudaSuccess) return e;
  }
  if (n_tc > 0) {
    TcParams tp{};
    tp.At = pl->At.p;
    tp.tiles = pl->d_mt_tc.p;
    tp.cslots = pl->d_cslots.p;
    tp.bl_ant0 = hp.bl_ant0;
    tp.bl_ant1 = hp.bl_ant1;
    tp.d_r = hp.d_r;
    tp.d_i = hp.d_i;
    tp.w = hp.w;
    for (int b = 0; b < 2; ++b) {
      tp.g_r[b] = hp.g_r[b];
      tp.g_i[b] = hp.g_i[b];
    }
    tp.c_r = hp.c_r;
    tp.c_i = hp.c_i;
    tp.z = hp.z;
    tp.dcpart = hp.dcpart;
    tp.dc_plane = pl->dc_plane;
    tp.partials = hp.partials + (size_t)(nitems + n_small + n_large) * 4;
    tp.st = hp.st;
    tp.nfp = hp.nfp;
    tp.dbg = pl->tc_dbg;
    tp.prof = pl->tc_prof.p;
    tp.prof_cta = pl->tc_prof_cta;
    {
      static const int flags = getenv("CALB2_TC_FLAGS") ? atoi(getenv("CALB2_TC_FLAGS")) : 0;  // measured best at HERA-350
      tp.flags = flags;
    }
    static bool configured[MAX_DEVICES] = {};
    const int dev = current_device();
    if (!configured[dev]) {
      cudaError_t e = cudaFuncSetAttribute(shared_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(shared_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e == cudaSuccess) e = cudaFuncSetAttribute(shared_tc_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) return e;
      configured[dev] = true;
    }
    tp.y = hp.y;
    tp.mode = sum ? 1 : 0;
    if (sum) {  // the pass of the plain chi^2 plus y and the model sums, then the two extra backward rows of the regulariser
      shared_tc_kernel<1><<<n_tc, TcCfg::NTHR, pl->tc_smem_bytes, s>>>(tp);
      cudaError_t e = cudaGetLastError();
      if (e != cudaSuccess) return e;
      tp.mode = 2;
      shared_tc_kernel<2><<<n_tc, TcCfg::NTHR, pl->tc_smem_bytes, s>>>(tp);
    } else {
      shared_tc_kernel<0><<<n_tc, TcCfg::NTHR, pl->tc_smem_bytes, s>>>(tp);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  if (n_small > 0) {
    sp.tiles = small_tiles;
    sp.partials = hp.partials + (size_t)nitems * 4;
    cudaError_t e = launch_shared_shape(sum, pl->cls_single_bl, 0, sp, n_small, s);
    if (e != cudaSuccess) return e;
  }
  if (fork_large || fork_items) return cudaStreamWaitEvent(s, pl->ev_join, 0);
  return cudaSuccess;
}
// A pass writes, for every class row, the planes of ITS segmentation only; coeffs_kernel adds nplanes of them.  Before the first
// pass of another kind (CUDA-core plain / 'sum' / tensor-core) the whole buffer is cleared, so planes a class does not write are 0.
static cudaError_t prepare_dcpart(calb2_plan* pl, int mode) {
  if (pl->dc_mode == mode || pl->classes.empty()) return cudaSuccess;
  pl->dc_mode = mode;
  return cudaMemsetAsync(pl->dcpart.p, 0, pl->dcpart.bytes(), pl->stream);
}
// partial-sum slots of one pass; with the tensor-core shape in use the layout is items | small rest | large | tensor-core
static int n_partials(const calb2_plan* pl, bool sum) {
  const int v = sum ? 1 : 0;
  if (tc_pass_of(pl, sum))
    return (int)(pl->items.size() + pl->mt_rest[v][0].size() + pl->mt_rest[v][1].size() + pl->mt_tc.size());
  return (int)(pl->items.size() + pl->mtiles[v][0].size() + pl->mtiles[v][1].size());
}
static int n_partials_max(const calb2_plan* pl) {
  int n = 0;
  for (int v = 0; v < 2; ++v) n = std::max(n, (int)(pl->items.size() + pl->mtiles[v][0].size() + pl->mtiles[v][1].size()));
  for (int v = 0; v < 2; ++v) n = std::max(n, (int)(pl->items.size() + pl->mt_rest[v][0].size() + pl->mt_rest[v][1].size() + pl->mt_tc.size()));
  return n;
}

static HeavyParams heavy_params(calb2_plan* pl, const FitState* st, bool sum, int store_v, int init_mode) {
  HeavyParams hp{};
  hp.A = pl->A.p;
  hp.items = pl->d_items.p;
  hp.row_slot = pl->row_slot.p;
  hp.row_coef = pl->row_coef.p;
  hp.slot_row0 = pl->d_slot_row0.p;
  hp.slot_bl0 = pl->d_slot_bl0.p;
  hp.slot_nb = pl->d_slot_nb.p;
  hp.bl_ant0 = pl->d_bl_ant0.p;
  hp.bl_ant1 = pl->d_bl_ant1.p;
  hp.d_r = pl->d_r.p;
  hp.d_i = pl->d_i.p;
  hp.w = pl->w.p;
  for (int b = 0; b < 2; ++b) {
    hp.g_r[b] = pl->g_r[b].p;
    hp.g_i[b] = pl->g_i[b].p;
  }
  hp.c_r = pl->c_r.p;
  hp.c_i = pl->c_i.p;
  hp.z = pl->z.p;
  hp.y = sum ? pl->y.p : nullptr;
  hp.dcpart = pl->dcpart.p;
  hp.vout = pl->vout.p;
  hp.partials = pl->partials.p;
  hp.st = st;
  hp.nfp = pl->nfp;
  hp.ntiles = pl->ntiles;
  hp.store_v = store_v;
  hp.init_mode = init_mode;
  hp.fuse_update = 0;
  {
    static const int dbg = getenv("CALB2_DBG") ? atoi(getenv("CALB2_DBG")) : 0;
    hp.dbg = dbg;
    hp.dbg_out = pl->dbg_out.p;
  }
  hp.c_r_rw = pl->c_r.p;
  hp.c_i_rw = pl->c_i.p;
  hp.cm_r = pl->cm_r.p;
  hp.cu_r = pl->cu_r.p;
  hp.cm_i = pl->cm_i.p;
  hp.cu_i = pl->cu_i.p;
  return hp;
}

static GainsParams gains_params(calb2_plan* pl, const FitState* st, const FitConsts& k, int mode, bool sum, int eval) {
  GainsParams gp{};
  gp.z = pl->z.p;
  gp.y = pl->y.p;
  gp.ant_ptr = pl->ant_ptr.p;
  gp.ant_ent = pl->ant_ent.p;
  gp.ant_partner = pl->ant_partner.p;
  gp.bl_ant0 = pl->d_bl_ant0.p;
  gp.bl_ant1 = pl->d_bl_ant1.p;
  for (int b = 0; b < 2; ++b) {
    gp.g_r[b] = pl->g_r[b].p;
    gp.g_i[b] = pl->g_i[b].p;
  }
  gp.m_r = pl->gm_r.p;
  gp.u_r = pl->gu_r.p;
  gp.m_i = pl->gm_i.p;
  gp.u_i = pl->gu_i.p;
  gp.snap_r = k.use_min ? pl->gsnap_r.p : nullptr;
  gp.snap_i = k.use_min ? pl->gsnap_i.p : nullptr;
  gp.grad_r = (mode != 0) ? pl->ggrad_r.p : nullptr;
  gp.grad_i = (mode != 0) ? pl->ggrad_i.p : nullptr;
  gp.st = st;
  gp.k = k;
  gp.nfp = pl->nfp;
  gp.nf = pl->nf;
  gp.nants = pl->nants;
  gp.mode = mode;
  gp.lamb.gain_partials = pl->lamb_gain_partials.p;
  gp.lamb.coef_partials = pl->lamb_coef_partials.p;
  gp.lamb.ratio = pl->lamb_ratio.p;
  gp.sum = sum ? 1 : 0;
  gp.eval = eval;
  gp.peers.n = 0;
  gp.tail_counter = nullptr;
  return gp;
}

// tc_pass: the backward sums come from a pass that used the tensor-core kernel (plain chi^2 fit / evaluation passes)
static bool tc_pass_of(const calb2_plan* pl, bool sum) { return pl->tc_enabled && (!sum || pl->tc_sum) && !pl->mt_tc.empty(); }
static CoeffParams coeff_params(calb2_plan* pl, const FitState* st, const FitConsts& k, int mode, bool sum, bool tc_pass = false) {
  CoeffParams cp{};
  cp.dcpart = pl->dcpart.p;
  cp.coef_row0 = pl->coef_row0.p;
  cp.coef_grp = pl->coef_grp.p;
  cp.grp_nslots = pl->d_grp_nslots.p;
  cp.grp_slot0 = pl->d_grp_slot0.p;
  cp.grp_coef0 = pl->d_grp_coef0.p;
  cp.slot_row0 = pl->d_slot_row0.p;
  cp.c_r = pl->c_r.p;
  cp.c_i = pl->c_i.p;
  cp.m_r = pl->cm_r.p;
  cp.u_r = pl->cu_r.p;
  cp.m_i = pl->cm_i.p;
  cp.u_i = pl->cu_i.p;
  cp.snap_r = k.use_min ? pl->csnap_r.p : nullptr;
  cp.snap_i = k.use_min ? pl->csnap_i.p : nullptr;
  cp.grad_r = (mode == 1) ? pl->cgrad_r.p : nullptr;
  cp.grad_i = (mode == 1) ? pl->cgrad_i.p : nullptr;
  cp.st = st;
  cp.k = k;
  cp.ncoef = (int)pl->ncoef;
  cp.nq = sum ? 4 : 2;
  cp.nplanes = pl->classes.empty() ? 1 : (tc_pass ? std::max(pl->nseg[sum ? 1 : 0], pl->nseg_tc) : pl->nseg[sum ? 1 : 0]);
  cp.plane = pl->dc_plane;
  cp.first_class_row = pl->first_class_row;
  cp.mode = mode;
  cp.var_bounds = pl->d_var_bounds.p;
  cp.nvar = (int)pl->var_bounds.size() - 1;
  cp.lamb_ratio = pl->lamb_ratio.p;
  return cp;
}

static int ensure_use_min_buffers(calb2_plan* pl) {
  if (!pl->gsnap_r.p) {
    if (int r = dalloc(pl->gsnap_r, (size_t)pl->nants * pl->nfp, pl)) return r;
    if (int r = dalloc(pl->gsnap_i, (size_t)pl->nants * pl->nfp, pl)) return r;
    if (int r = dalloc(pl->csnap_r, (size_t)pl->ncoef, pl)) return r;
    if (int r = dalloc(pl->csnap_i, (size_t)pl->ncoef, pl)) return r;
  }
  return 0;
}
static int ensure_lamb_buffers(calb2_plan* pl) {
  if (pl->var_bounds.empty()) pl->var_bounds = {0ll, (long long)pl->ncoef};
  const size_t nvar = pl->var_bounds.size() - 1;
  if (!pl->var_uploaded) {
    if (int r = dalloc(pl->d_var_bounds, nvar + 1, pl)) return r;
    CU(cudaMemcpyAsync(pl->d_var_bounds.p, pl->var_bounds.data(), (nvar + 1) * sizeof(long long), cudaMemcpyHostToDevice, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    if (int r = dalloc(pl->lamb_coef_partials, nvar * LAMB_SPLIT * 4, pl)) return r;
    if (int r = dalloc(pl->lamb_ratio, 2 + 2 * nvar, pl)) return r;
    pl->var_uploaded = true;
  }
  if (!pl->lamb_gain_partials.p)
    if (int r = dalloc(pl->lamb_gain_partials, (size_t)pl->nants * ((pl->nfp + GK_CH - 1) / GK_CH) * 4, pl)) return r;
  return 0;
}
static int ensure_sum_buffers(calb2_plan* pl) {
  if (!pl->y.p) return dalloc(pl->y, (size_t)pl->nbls * pl->nfp, pl);
  return 0;
}
static int ensure_grad_buffers(calb2_plan* pl) {
  if (!pl->ggrad_r.p) {
    if (int r = dalloc(pl->ggrad_r, (size_t)pl->nants * pl->nfp, pl)) return r;
    if (int r = dalloc(pl->ggrad_i, (size_t)pl->nants * pl->nfp, pl)) return r;
  }
  if (!pl->cgrad_r.p) {
    if (int r = dalloc(pl->cgrad_r, (size_t)pl->ncoef, pl)) return r;
    if (int r = dalloc(pl->cgrad_i, (size_t)pl->ncoef, pl)) return r;
  }
  return 0;
}
static int ensure_vout(calb2_plan* pl) {
  if (!pl->vout.p) return dalloc(pl->vout, (size_t)pl->nslots * pl->nfp, pl);
  return 0;
}

// host [n][nf] -> device [n][nfp].  Arrays whose padding columns are zero (data, weights, sky model: the buffers are
// zero-initialised and nothing ever writes non-zero values into the padding) go with ONE pitched copy straight from the
// caller's memory -- a true asynchronous DMA when that memory is pinned, the driver's own staged pipeline otherwise.
// The gain tables (padding = 1) still take the staging buffer + pad kernel; they are small.
static int upload_padded(calb2_plan* pl, const float* src, float* dst, size_t nrows, float fill) {
  if (fill == 0.f) {
    CU(cudaMemcpy2DAsync(dst, (size_t)pl->nfp * sizeof(float), src, (size_t)pl->nf * sizeof(float), (size_t)pl->nf * sizeof(float),
                         nrows, cudaMemcpyHostToDevice, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    return 0;
  }
  const size_t rows_per = std::max<size_t>(1, pl->staging_floats / (size_t)pl->nf);
  for (size_t r0 = 0; r0 < nrows; r0 += rows_per) {
    const size_t n = std::min(rows_per, nrows - r0);
    memcpy(pl->h_staging, src + r0 * pl->nf, n * pl->nf * sizeof(float));
    CU(cudaMemcpyAsync(pl->staging.p, pl->h_staging, n * pl->nf * sizeof(float), cudaMemcpyHostToDevice, pl->stream));
    pad_rows_kernel<<<(unsigned)n, 128, 0, pl->stream>>>(pl->staging.p, dst + r0 * pl->nfp, pl->nf, pl->nfp, fill);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(pl->stream));
  }
  return 0;
}
static int download_unpadded(calb2_plan* pl, const float* src, float* dst, size_t nrows) {
  CU(cudaMemcpy2DAsync(dst, (size_t)pl->nf * sizeof(float), src, (size_t)pl->nfp * sizeof(float), (size_t)pl->nf * sizeof(float), nrows,
                       cudaMemcpyDeviceToHost, pl->stream));
  CU(cudaStreamSynchronize(pl->stream));
  return 0;
}

static int set_eval_state(calb2_plan* pl) {
  FitState s{};
  s.step = pl->cur_buf;
  s.stop_after = INT_MAX;
  s.upd_active = 0;
  CU(cudaMemcpyAsync(pl->state_eval.p, &s, sizeof(s), cudaMemcpyHostToDevice, pl->stream));
  return 0;
}

static int all_reduce(calb2_plan* pl, void* buf, size_t count, int dtype) {
  if (pl->nranks <= 1) return 0;
  int rc = g_nccl.AllReduce(buf, buf, count, dtype, NCCL_SUM, pl->comm, pl->stream);
  if (rc != 0) return fail(CALB2_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  return 0;
}

// One payload-free round of the peer exchange on the plan's stream: publish the next sequence value, wait (bounded) for
// every peer to publish it.  All ranks call it at the same points (fit begin, close), so it orders them.
static int peer_barrier(calb2_plan* pl) {
  const unsigned int seq = pl->xseq++;
  unsigned int* my_flag = reinterpret_cast<unsigned int*>(pl->xbuf);
  xpublish_kernel<<<1, 1, 0, pl->stream>>>(my_flag, 2u * seq + 2u);
  CU(cudaGetLastError());
  xwait_kernel<<<1, 32, 0, pl->stream>>>(pl->peers, 2u * seq + 2u, pl->state.p, 2, pl->xtimeout_ns);
  CU(cudaGetLastError());
  return 0;
}

// One optimizer iteration (calibration.py:663-668) enqueued on the plan's stream.
#ifdef CALB2_PROFILE
#define CALB2_STAGE() \
  if (pl->stage_cursor < pl->stage_ev.size()) cudaEventRecord(pl->stage_ev[pl->stage_cursor++], pl->stream);
#else
#define CALB2_STAGE()
#endif
[[maybe_unused]] static constexpr int NSTAGE = 8;  // boundaries per step: start, heavy, partials, gains-reduce, all-reduce, finalize, gains, coeffs

static int enqueue_step(calb2_plan* pl, const FitConsts& k, bool sum, bool freeze, bool want_fuse, float* hist,
                        cudaEvent_t ev0, cudaEvent_t ev1, long long* launches) {
  // coefficients can take their optimizer step in the heavy kernel's tail when nothing couples the groups
  const bool fuse = want_fuse && !sum && !freeze && pl->all_single_slot && pl->classes.empty() &&
                    k.optimizer <= CALB2_OPT_SGD && k.momentum == 0.f;
  int npartials = n_partials(pl, sum);
  const double* partials = pl->partials.p;
  if (ev0) CU(cudaEventRecord(ev0, pl->stream));
  CALB2_STAGE()
  if (freeze) {  // the model is fixed: elementwise pass over the visibilities only
    LightParams lp{};
    lp.vout = pl->vout.p;
    lp.bl_slot = pl->d_bl_slot.p;
    lp.bl_ant0 = pl->d_bl_ant0.p;
    lp.bl_ant1 = pl->d_bl_ant1.p;
    lp.d_r = pl->d_r.p;
    lp.d_i = pl->d_i.p;
    lp.w = pl->w.p;
    for (int b = 0; b < 2; ++b) {
      lp.g_r[b] = pl->g_r[b].p;
      lp.g_i[b] = pl->g_i[b].p;
    }
    lp.z = pl->z.p;
    lp.y = sum ? pl->y.p : nullptr;
    lp.partials = pl->light_partials.p;
    lp.st = pl->state.p;
    lp.nfp = pl->nfp;
    lp.nelem = pl->nbls * (long long)pl->nfp;
    lp.sum = sum ? 1 : 0;
    light_kernel<<<pl->light_blocks, 256, 0, pl->stream>>>(lp);
    CU(cudaGetLastError());
    npartials = pl->light_blocks;
    partials = pl->light_partials.p;
  } else {
    HeavyParams hp = heavy_params(pl, pl->state.p, sum, 0, 0);
    hp.fuse_update = fuse ? 1 : 0;
    hp.k = k;
    CU(launch_heavy(pl, sum, hp, (int)pl->items.size(), pl->stream));
  }
  if (ev1) CU(cudaEventRecord(ev1, pl->stream));
  CALB2_STAGE()
  FinalizeParams fp{};
  fp.st = pl->state.p;
  fp.k = k;
  fp.hist = hist;
  fp.eval_only = 0;
  fp.peers.n = 0;
  fp.xwait = 0;
  fp.xpar = 0;
  fp.timeout_ns = pl->xtimeout_ns;
  dim3 ggrid(pl->nants, (pl->nfp + GK_CH - 1) / GK_CH);
  const size_t ngrad = (size_t)2 * pl->nants * pl->nfp;
  // The coefficient update only needs finalize's scalars and the fused kernel's backward sums: it is forked onto a second
  // stream right after finalize_kernel and runs next to the gain-gradient reduce / exchange / gain update.
  const bool need_coeffs = (!freeze && !fuse) || (k.use_min && !freeze);
  bool forked = false;
  auto fork_coeffs = [&]() -> int {
    if (!need_coeffs) return 0;
    CU(cudaEventRecord(pl->ev_fork, pl->stream));
    CU(cudaStreamWaitEvent(pl->stream2, pl->ev_fork, 0));
    // mode 0: optimizer step from the stored backward sums; mode 3: only the use_min snapshot copy
    coeffs_kernel<<<(unsigned)((pl->ncoef + 255) / 256), 256, 0, pl->stream2>>>(coeff_params(pl, pl->state.p, k, fuse ? 3 : 0, sum, tc_pass_of(pl, sum)));
    CU(cudaGetLastError());
    CU(cudaEventRecord(pl->ev_join, pl->stream2));
    forked = true;
    *launches += 1;
    return 0;
  };
  if (pl->nranks > 1 && pl->peers.n > 1) {
    // ---- exchange through peer memory (NVLink): no collective library call in the loop ----
    const unsigned int seq = pl->xseq++;
    const int par = (int)(seq & 1u);
    unsigned int* my_flag = reinterpret_cast<unsigned int*>(pl->xbuf);
    double* my_scal = reinterpret_cast<double*>(pl->xbuf + XBUF_FLAG_BYTES) + par * 4;
    float* my_grad = reinterpret_cast<float*>(pl->xbuf + XBUF_FLAG_BYTES + XBUF_SCAL_BYTES) + (size_t)par * ngrad;
    fp.partials = nullptr;
    fp.nitems = 0;
    fp.peers = pl->peers;
    fp.xpar = par;
    GainsParams gred = gains_params(pl, pl->state.p, k, sum ? 1 : 4, sum, 0);
    gred.grad_r = my_grad;
    gred.grad_i = my_grad + ngrad / 2;
    GainsParams gupd = gains_params(pl, pl->state.p, k, 2, sum, 0);
    gupd.peers = pl->peers;
    gupd.xpar = par;
    if (!sum) {
      // The local gradient partial does not need finalize's alpha / beta.  One launch: gradient reduce, then its last
      // CTA reduces the per-item partial sums into the exchange buffer and publishes scalars and gradient together.
      gred.tail_counter = pl->tail_counter.p;
      gred.tail_partials = partials;
      gred.tail_npartials = npartials;
      gred.tail_scal = my_scal;
      gred.tail_flag = my_flag;
      gred.tail_value = 2u * seq + 2u;
      CALB2_STAGE()
      gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gred);
      CU(cudaGetLastError());
      CALB2_STAGE()
      fp.xwait = 2u * seq + 2u;
      CALB2_STAGE()
      finalize_kernel<<<1, 1024, 0, pl->stream>>>(fp);
      CU(cudaGetLastError());
      CALB2_STAGE()
      if (int r = fork_coeffs()) return r;
    } else {
      // 'sum' regulariser: alpha / beta need the global sums first (two publish / wait rounds per step)
      reduce_partials_kernel<<<1, 1024, 0, pl->stream>>>(partials, npartials, my_scal);
      CU(cudaGetLastError());
      xpublish_kernel<<<1, 1, 0, pl->stream>>>(my_flag, 2u * seq + 1u);
      CU(cudaGetLastError());
      fp.xwait = 2u * seq + 1u;
      finalize_kernel<<<1, 1024, 0, pl->stream>>>(fp);
      CU(cudaGetLastError());
      if (int r = fork_coeffs()) return r;
      gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gred);
      CU(cudaGetLastError());
      xpublish_kernel<<<1, 1, 0, pl->stream>>>(my_flag, 2u * seq + 2u);
      CU(cudaGetLastError());
      xwait_kernel<<<1, 32, 0, pl->stream>>>(pl->peers, 2u * seq + 2u, pl->state.p, 1, pl->xtimeout_ns);
      CU(cudaGetLastError());
      *launches += 5;
    }
    gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gupd);
    CU(cudaGetLastError());
    CALB2_STAGE()
    *launches += 1;
  } else if (pl->nranks > 1) {
    reduce_partials_kernel<<<1, 1024, 0, pl->stream>>>(partials, npartials, pl->comm_scalars.p);
    CU(cudaGetLastError());
    *launches += 1;
    CALB2_STAGE()
    fp.partials = pl->comm_scalars.p;
    fp.nitems = 1;
    if (!sum && g_nccl.GroupStart && g_nccl.GroupEnd) {
      // without the regulariser the gain-gradient reduce does not need finalize's alpha/beta: reduce first and
      // send scalars + gradient tables in ONE grouped NCCL launch
      gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gains_params(pl, pl->state.p, k, 4, sum, 0));
      CU(cudaGetLastError());
      CALB2_STAGE()
      g_nccl.GroupStart();
      int r1 = all_reduce(pl, pl->comm_scalars.p, 4, NCCL_FLOAT64);
      int r2 = all_reduce(pl, pl->ggrad_r.p, ngrad, NCCL_FLOAT32);
      g_nccl.GroupEnd();
      if (r1) return r1;
      if (r2) return r2;
      CALB2_STAGE()
      finalize_kernel<<<1, 1024, 0, pl->stream>>>(fp);
      CU(cudaGetLastError());
      CALB2_STAGE()
      if (int r = fork_coeffs()) return r;
    } else {
      if (int r = all_reduce(pl, pl->comm_scalars.p, 4, NCCL_FLOAT64)) return r;
      finalize_kernel<<<1, 1024, 0, pl->stream>>>(fp);
      CU(cudaGetLastError());
      if (int r = fork_coeffs()) return r;
      gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gains_params(pl, pl->state.p, k, 1, sum, 0));
      CU(cudaGetLastError());
      // real and imaginary gradient tables are adjacent halves of one allocation
      if (int r = all_reduce(pl, pl->ggrad_r.p, ngrad, NCCL_FLOAT32)) return r;
    }
    gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gains_params(pl, pl->state.p, k, 2, sum, 0));
    CU(cudaGetLastError());
    CALB2_STAGE()
    *launches += 1;
  } else {
    fp.partials = partials;
    fp.nitems = npartials;
    finalize_kernel<<<1, 1024, 0, pl->stream>>>(fp);
    CU(cudaGetLastError());
    if (k.optimizer == CALB2_OPT_LAMB) {
      // pass 1: moments everywhere + norm partials; trust ratio per variable; pass 2: apply.  One stream, seven launches.
      const unsigned cgrid = (unsigned)((pl->ncoef + 255) / 256);
      const int nvar = freeze ? 0 : (int)pl->var_bounds.size() - 1;
      gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gains_params(pl, pl->state.p, k, 5, sum, 0));
      CU(cudaGetLastError());
      if (!freeze) {
        coeffs_kernel<<<cgrid, 256, 0, pl->stream>>>(coeff_params(pl, pl->state.p, k, 5, sum, tc_pass_of(pl, sum)));
        CU(cudaGetLastError());
        LambNormParams np{};
        np.c_r = pl->c_r.p;
        np.c_i = pl->c_i.p;
        np.m_r = pl->cm_r.p;
        np.u_r = pl->cu_r.p;
        np.m_i = pl->cm_i.p;
        np.u_i = pl->cu_i.p;
        np.var_bounds = pl->d_var_bounds.p;
        np.partials = pl->lamb_coef_partials.p;
        np.st = pl->state.p;
        np.k = k;
        lamb_coef_norm_kernel<<<dim3((unsigned)nvar, LAMB_SPLIT), 256, 0, pl->stream>>>(np);
        CU(cudaGetLastError());
      }
      LambRatioParams rp{};
      rp.gain_partials = pl->lamb_gain_partials.p;
      rp.n_gain_partials = (int)(ggrid.x * ggrid.y);
      rp.coef_partials = pl->lamb_coef_partials.p;
      rp.ratio = pl->lamb_ratio.p;
      lamb_ratio_kernel<<<1 + nvar, 32, 0, pl->stream>>>(rp);
      CU(cudaGetLastError());
      gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gains_params(pl, pl->state.p, k, 6, sum, 0));
      CU(cudaGetLastError());
      if (!freeze) {
        coeffs_kernel<<<cgrid, 256, 0, pl->stream>>>(coeff_params(pl, pl->state.p, k, 6, sum, tc_pass_of(pl, sum)));
        CU(cudaGetLastError());
      }
      *launches += freeze ? 1 : 4;
    } else {
      if (int r = fork_coeffs()) return r;
      gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gains_params(pl, pl->state.p, k, 0, sum, 0));
      CU(cudaGetLastError());
    }
  }
  if (forked) CU(cudaStreamWaitEvent(pl->stream, pl->ev_join, 0));  // join: the next step reads the new coefficients
  CALB2_STAGE()
  *launches += 3;
  return 0;
}


// tensorize_fg_coeffs on the device (calibration.py:828-913), real and imaginary parts in one go.
static int init_coeffs_impl(calb2_plan* pl, const float* sky_r, const float* sky_i) {
  const size_t nd = (size_t)pl->nbls * pl->nfp;
  if (pl->sky_r.n < nd) {
    if (int r = dalloc(pl->sky_r, nd, pl)) return r;
    if (int r = dalloc(pl->sky_i, nd, pl)) return r;
  }
  if (int r = ensure_grad_buffers(pl)) return r;
  if (int r = upload_padded(pl, sky_r, pl->sky_r.p, (size_t)pl->nbls, 0.f)) return r;
  if (int r = upload_padded(pl, sky_i, pl->sky_i.p, (size_t)pl->nbls, 0.f)) return r;
  if (int r = set_eval_state(pl)) return r;
  // right-hand sides A^T (sky * (w != 0)) through the fused kernel's backward contraction
  HeavyParams hp = heavy_params(pl, pl->state_eval.p, false, 0, 1);
  hp.d_r = pl->sky_r.p;
  hp.d_i = pl->sky_i.p;
  CU(prepare_dcpart(pl, 1));
  CU(launch_heavy(pl, false, hp, (int)pl->items.size(), pl->stream));
  FitConsts k{};
  coeffs_kernel<<<(unsigned)((pl->ncoef + 255) / 256), 256, 0, pl->stream>>>(coeff_params(pl, pl->state_eval.p, k, 1, false));
  CU(cudaGetLastError());
  // Gram + Cholesky per group, in batches bounded by the scratch budget
  const size_t budget = (size_t)48 << 20;  // doubles (384 MiB)
  DevBuf<double> gram;
  DevBuf<GramJob> djobs;
  std::vector<GramJob> jobs;
  size_t used = 0;
  int maxn = 0;
  auto flush = [&]() -> int {
    if (jobs.empty()) return 0;
    if (gram.n < used) CU(gram.alloc(std::max(used, budget)));
    if (djobs.n < jobs.size()) CU(djobs.alloc(jobs.size() * 2));
    CU(cudaMemcpyAsync(djobs.p, jobs.data(), jobs.size() * sizeof(GramJob), cudaMemcpyHostToDevice, pl->stream));
    constexpr int FC = 32;
    const size_t smem = (size_t)maxn * (FC + 1) * sizeof(float);
    static size_t configured[MAX_DEVICES] = {};
    const int dev = current_device();
    if (smem > configured[dev]) {
      CU(cudaFuncSetAttribute(gram_kernel<FC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 48 << 10)));
      configured[dev] = smem;
    }
    gram_kernel<FC><<<(unsigned)jobs.size(), 256, smem, pl->stream>>>(pl->A.p, djobs.p, pl->slot_geom.p, gram.p, pl->nf, pl->FT);
    CU(cudaGetLastError());
    chol_solve_kernel<<<(unsigned)jobs.size(), 256, 0, pl->stream>>>(djobs.p, gram.p, pl->cgrad_r.p, pl->cgrad_i.p);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(pl->stream));
    jobs.clear();
    used = 0;
    maxn = 0;
    return 0;
  };
  for (int g = 0; g < pl->ngroups; ++g) {
    const int n = pl->grp_ncomp[g];
    if (n == 0) continue;
    const size_t need = (size_t)n * n + 2 * (size_t)n;
    if (used + need > budget && !jobs.empty())
      if (int r = flush()) return r;
    GramJob jb{};
    jb.gram_off = (long long)used;
    jb.grp = g;
    jb.n = n;
    jb.slot0 = pl->grp_slot0[g];
    jb.nslots = pl->grp_nslots[g];
    jb.coef0 = pl->grp_coef0[g];
    jobs.push_back(jb);
    used += need;
    maxn = std::max(maxn, n);
  }
  if (int r = flush()) return r;
  gram.release();
  djobs.release();
  CU(cudaMemcpyAsync(pl->c_r.p, pl->cgrad_r.p, pl->ncoef * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
  CU(cudaMemcpyAsync(pl->c_i.p, pl->cgrad_i.p, pl->ncoef * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
  CU(cudaStreamSynchronize(pl->stream));
  pl->have_coeffs = true;
  return 0;
}

}  // namespace calb2

#include "calfit_generic_host.cuh"

// ====================================================================================================
// C ABI
// ====================================================================================================
extern "C" {

const char* calb2_last_error(void) { return g_err.c_str(); }
const char* calb2_version(void) { return "calamity_b200 0.2 (sm_100a)"; }

int calb2_debug_check_guards(int64_t* nbuffers, int64_t* nviolations) {
  if (!nbuffers || !nviolations) return fail(CALB2_ERR_ARG, "null argument");
  *nbuffers = 0;
  *nviolations = 0;
  if (!guards_enabled()) return fail(CALB2_ERR_STATE, "guard zones are off: set CALB2_GUARD=1 before the first allocation");
  CU(cudaDeviceSynchronize());
  std::lock_guard<std::mutex> lk(guards().mu);
  std::vector<unsigned char> h(GUARD_BYTES);
  for (const auto& kv : guards().live) {
    cudaPointerAttributes attr{};
    if (cudaPointerGetAttributes(&attr, kv.first) != cudaSuccess) continue;
    int cur = 0;
    cudaGetDevice(&cur);
    if (attr.device != cur) continue;  // the caller checks one device at a time
    ++*nbuffers;
    for (int side = 0; side < 2; ++side) {
      const unsigned char* src = static_cast<unsigned char*>(kv.first) + (side ? GUARD_BYTES + kv.second : 0);
      CU(cudaMemcpy(h.data(), src, GUARD_BYTES, cudaMemcpyDeviceToHost));
      for (unsigned char b : h)
        if (b != GUARD_PATTERN) {
          ++*nviolations;
          break;
        }
    }
  }
  return 0;
}

int calb2_debug_tc_record(uint32_t* out16) {
  if (!out16) return fail(CALB2_ERR_ARG, "null argument");
  for (int i = 0; i < 16; ++i) out16[i] = g_tc_dbg ? g_tc_dbg[i] : 0u;
  return 0;
}

int calb2_debug_tc_profile(calb2_plan* pl, int64_t* out, int32_t n) {
  if (!pl || !out) return fail(CALB2_ERR_ARG, "null argument");
  if (!pl->tc_prof.p) return fail(CALB2_ERR_STATE, "set CALB2_TC_PROF=<cta> before creating the plan");
  CU(cudaSetDevice(pl->device));
  CU(cudaStreamSynchronize(pl->stream));
  const size_t m = std::min<size_t>((size_t)std::max(n, 0), pl->tc_prof.n);
  CU(cudaMemcpy(out, pl->tc_prof.p, m * sizeof(long long), cudaMemcpyDeviceToHost));
  return 0;
}

int calb2_device_count(int32_t* count) {
  if (!count) return fail(CALB2_ERR_ARG, "null argument");
  int n = 0;
  CU(cudaGetDeviceCount(&n));
  *count = n;
  return 0;
}

int calb2_plan_create(const calb2_plan_desc* d, calb2_plan** out) {
  if (!d || !out) return fail(CALB2_ERR_ARG, "null argument");
  if (d->nants <= 0 || d->nfreqs <= 0 || d->ngroups <= 0) return fail(CALB2_ERR_ARG, "empty problem");
  const int rpt = RPT_DEFAULT, nwarp = 8;
  if (d->dtype != CALB2_F32 && d->dtype != CALB2_F64) return fail(CALB2_ERR_ARG, "dtype must be CALB2_F32 or CALB2_F64");
  for (int g = 0; g < d->ngroups; ++g)
    if (d->group_ncomp[g] < 0 || d->group_nslots[g] <= 0) return fail(CALB2_ERR_ARG, "group %d: bad ncomp/nslots", g);
  // float64, or a group with more basis vectors than the fused kernel stages: the generic unfused path
  bool generic = d->dtype == CALB2_F64 || (getenv("CALB2_GENERIC") && atoi(getenv("CALB2_GENERIC")) != 0);

  // ---- shared-basis classes: single-slot groups whose basis block is shared by at least `min_members` groups ----
  std::vector<int> grp_cls(d->ngroups, -1);
  std::vector<std::vector<int>> cls_members;
  int shared_mode = d->shared_basis;
  if (getenv("CALB2_SHARED_BASIS")) shared_mode = atoi(getenv("CALB2_SHARED_BASIS"));
  if (d->group_class && shared_mode >= 0 && !generic) {
    int min_members = shared_mode >= 1 ? 1 : 4;
    if (getenv("CALB2_CLS_MIN")) min_members = std::max(1, atoi(getenv("CALB2_CLS_MIN")));
    std::unordered_map<int, int> index_of;  // class id -> position in `cands` (first-appearance order: deterministic)
    std::vector<std::vector<int>> cands;
    for (int g = 0; g < d->ngroups; ++g) {
      const int id = d->group_class[g];
      if (id < 0 || d->group_nslots[g] != 1 || d->group_ncomp[g] < 1 || d->group_ncomp[g] > KPM_LARGE) continue;
      auto it = index_of.find(id);
      if (it == index_of.end()) {
        it = index_of.emplace(id, (int)cands.size()).first;
        cands.emplace_back();
      }
      cands[it->second].push_back(g);
    }
    for (auto& mem : cands) {
      if ((int)mem.size() < min_members) continue;
      for (int g : mem)
        if (d->group_ncomp[g] != d->group_ncomp[mem[0]])
          return fail(CALB2_ERR_ARG, "groups %d and %d share a basis class but have %d and %d vectors", mem[0], g,
                      d->group_ncomp[mem[0]], d->group_ncomp[g]);
      for (int g : mem) grp_cls[g] = (int)cls_members.size();
      cls_members.push_back(mem);
    }
  }
  // tile width of the streaming path, from the groups that stay on it
  int fl = 8;
  {
    std::vector<int32_t> h_ncomp, h_nslots;
    for (int g = 0; g < d->ngroups; ++g)
      if (grp_cls[g] < 0) {
        h_ncomp.push_back(d->group_ncomp[g]);
        h_nslots.push_back(d->group_nslots[g]);
      }
    if (!h_ncomp.empty()) {
      calb2_plan_desc hd = *d;
      hd.ngroups = (int32_t)h_ncomp.size();
      hd.group_ncomp = h_ncomp.data();
      hd.group_nslots = h_nslots.data();
      if (int r = choose_fl(&hd, rpt, nwarp, &fl)) {
        if (r != CALB2_ERR_UNSUPPORTED || d->tile_freqs) return r;
        generic = true;  // single GPU, no classes
        fl = 8;
        std::fill(grp_cls.begin(), grp_cls.end(), -1);
        cls_members.clear();
      }
    }
  }
  CU(cudaSetDevice(d->device));
  calb2_plan* pl = new calb2_plan();
  pl->device = d->device;
  pl->dtype = d->dtype;
  pl->nants = d->nants;
  pl->nf = d->nfreqs;
  pl->ngroups = d->ngroups;
  pl->FL = fl;
  pl->G = 32 / fl;
  pl->FT = 4 * fl;
  pl->RPT = rpt;
  pl->NW = nwarp;
  pl->KMAX = nwarp * rpt * pl->G;
  {
    const int quantum = cls_members.empty() ? pl->FT : std::max(pl->FT, SHARED_FT);
    pl->nfp = ((pl->nf + quantum - 1) / quantum) * quantum;
  }
  pl->ntiles = pl->nfp / pl->FT;
  pl->ntiles_c = pl->nfp / SHARED_FT;
  pl->grp_cls = grp_cls;
  const int G = pl->G;

  // ---- flatten groups -> slots -> baselines (canonical order), then number the slots internally ----
  pl->grp_ncomp.assign(d->group_ncomp, d->group_ncomp + d->ngroups);
  pl->grp_nslots.assign(d->group_nslots, d->group_nslots + d->ngroups);
  pl->grp_slot0.resize(d->ngroups);
  pl->grp_coef0.resize(d->ngroups);
  std::vector<long long> canon_slot0(d->ngroups);
  long long ns = 0, nc = 0, nh = 0;
  for (int g = 0; g < d->ngroups; ++g) {
    canon_slot0[g] = ns;
    pl->grp_coef0[g] = (int)nc;
    if (d->group_nslots[g] != 1) pl->all_single_slot = false;
    ns += d->group_nslots[g];
    nc += d->group_ncomp[g];
    if (grp_cls[g] < 0) nh += d->group_nslots[g];
  }
  pl->nslots = ns;
  pl->ncoef = nc;
  pl->nslots_heavy = nh;
  pl->nslots_class = ns - nh;
  {
    long long next_h = 0;
    for (int g = 0; g < d->ngroups; ++g)
      if (grp_cls[g] < 0) {
        pl->grp_slot0[g] = (int)next_h;
        next_h += d->group_nslots[g];
      }
    long long next_c = nh;
    for (auto& mem : cls_members)
      for (int g : mem) pl->grp_slot0[g] = (int)next_c++;
  }
  std::vector<long long> canon_bl0(ns + 1);
  long long nb = 0;
  for (long long s = 0; s < ns; ++s) {
    canon_bl0[s] = nb;
    if (d->slot_nbls[s] <= 0) {
      delete pl;
      return fail(CALB2_ERR_ARG, "slot %lld: no baselines", s);
    }
    nb += d->slot_nbls[s];
  }
  canon_bl0[ns] = nb;
  pl->nbls = nb;
  pl->slot_nbls.resize(ns);
  pl->slot_grp.resize(ns);
  pl->slot_bl0.resize(ns + 1);
  for (int g = 0; g < d->ngroups; ++g)
    for (int s = 0; s < d->group_nslots[g]; ++s) {
      const long long cs = canon_slot0[g] + s, is = pl->grp_slot0[g] + s;
      pl->slot_grp[is] = g;
      pl->slot_nbls[is] = d->slot_nbls[cs];
      pl->slot_bl0[is] = (int)canon_bl0[cs];
      if (d->slot_nbls[cs] != 1) {
        pl->all_single_bl = false;
        (grp_cls[g] < 0 ? pl->heavy_single_bl : pl->cls_single_bl) = false;
      }
    }
  pl->slot_bl0[ns] = (int)nb;  // only meaningful when the numbering is canonical (the generic path)
  pl->bl_ant0.assign(d->bl_ant0, d->bl_ant0 + nb);
  pl->bl_ant1.assign(d->bl_ant1, d->bl_ant1 + nb);
  for (long long b = 0; b < nb; ++b)
    if (pl->bl_ant0[b] < 0 || pl->bl_ant0[b] >= d->nants || pl->bl_ant1[b] < 0 || pl->bl_ant1[b] >= d->nants) {
      delete pl;
      return fail(CALB2_ERR_ARG, "baseline %lld: antenna index out of range", b);
    }
  pl->bl_slot.resize(nb);
  for (long long s = 0; s < ns; ++s)
    for (int b = pl->slot_bl0[s]; b < pl->slot_bl0[s] + pl->slot_nbls[s]; ++b) pl->bl_slot[b] = (int)s;

  // ---- streaming path: pack its slots into items (one CTA each): rows <= KMAX, slots <= SMAX ----
  pl->slot_row0.resize(ns + 1);
  pl->slot_item.assign(ns, -1);
  std::vector<unsigned char> row_slot;
  std::vector<int> row_coef;
  ItemDesc cur{};
  cur.nrows = 0;
  cur.nslots = 0;
  long long rows = 0, a_off = 0;
  auto close_item = [&]() {
    if (cur.nslots == 0) return;
    pl->items.push_back(cur);
    a_off += (long long)cur.nrows * pl->nfp;
    cur = ItemDesc{};
  };
  for (long long s = 0; s < nh; ++s) {
    const int g = pl->slot_grp[s];
    const int ncomp = pl->grp_ncomp[g];
    const int rp = std::max(G, ((ncomp + G - 1) / G) * G);  // at least one step so the slot exists in the item
    if (cur.nslots > 0 && (cur.nrows + rp > pl->KMAX || cur.nslots >= SMAX)) close_item();
    if (cur.nslots == 0) {
      cur.a_off = a_off;
      cur.row0 = (int)rows;
      cur.slot0 = (int)s;
    }
    pl->slot_row0[s] = (int)rows;
    pl->slot_item[s] = (int)pl->items.size();
    for (int k = 0; k < rp; ++k) {
      row_slot.push_back((unsigned char)cur.nslots);
      row_coef.push_back(k < ncomp ? pl->grp_coef0[g] + k : -1);
    }
    cur.nrows += rp;
    cur.nslots += 1;
    rows += rp;
    pl->n_a_nz += (long long)ncomp * pl->nf;
  }
  close_item();
  // Launch order = longest first (rows, then slots): the grid is consumed in index order, so the kernel's tail
  // is made of the cheapest items.  Only the ORDER of the descriptors changes; rows / slots keep their places.
  if (!getenv("CALB2_NO_LPT")) {
    std::vector<int> order(pl->items.size());
    for (size_t n = 0; n < order.size(); ++n) order[n] = (int)n;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      if (pl->items[a].nrows != pl->items[b].nrows) return pl->items[a].nrows > pl->items[b].nrows;
      return pl->items[a].nslots > pl->items[b].nslots;
    });
    std::vector<ItemDesc> sorted(pl->items.size());
    std::vector<int> new_index(pl->items.size());
    for (size_t n = 0; n < order.size(); ++n) {
      sorted[n] = pl->items[order[n]];
      new_index[order[n]] = (int)n;
    }
    pl->items.swap(sorted);
    for (long long s = 0; s < nh; ++s) pl->slot_item[s] = new_index[pl->slot_item[s]];
  }
  const long long heavy_rows = rows;
  pl->first_class_row = (int)heavy_rows;
  const long long heavy_floats = heavy_rows * (long long)pl->nfp;

  // ---- shared-basis path: one block per class after the streaming tiles; backward-sum rows after the streaming rows ----
  std::vector<ClassSlot> cslots;
  std::vector<int> cs_slot;
  {
    long long off = heavy_floats;
    for (auto& mem : cls_members) {
      calb2_plan::ClsInfo ci{};
      ci.ncomp = pl->grp_ncomp[mem[0]];
      ci.kp = ((ci.ncomp + 7) / 8) * 8;
      ci.nmembers = (int)mem.size();
      ci.a_off = off;
      ci.uploaded = false;
      off += (long long)ci.kp * pl->nfp;
      ci.kpt = ((ci.ncomp + 15) / 16) * 16;
      ci.tc = ci.kpt <= TcCfg::KPMAX;
      ci.tc_off = 0;
      const int first_cs = (int)cslots.size();
      for (int g : mem) {
        const int is = pl->grp_slot0[g];
        pl->slot_row0[is] = (int)rows;
        ClassSlot cs{};
        cs.coef0 = pl->grp_coef0[g];
        cs.row0 = (int)rows;
        cs.bl0 = pl->slot_bl0[is];
        cs.nb = pl->slot_nbls[is];
        cslots.push_back(cs);
        cs_slot.push_back(is);
        rows += ci.ncomp;
        pl->n_a_nz += (long long)ci.ncomp * pl->nf;
      }
      pl->classes.push_back(ci);
      (void)first_cs;
    }
    pl->a_class_floats = off - heavy_floats;
  }
  {
    // tensor-core shape: on by default for single-baseline slots; CALB2_TC=0 keeps every class on the CUDA-core shapes
    pl->tc_enabled = pl->cls_single_bl && !(getenv("CALB2_TC") && atoi(getenv("CALB2_TC")) == 0);
    // a tensor-core CTA costs the same for 1 or 64 groups (128 accumulator rows): classes of up to 128 vectors with fewer than
    // tc_min members stay on the CUDA-core shapes (32 / 64 groups per CTA, cost by 8-group block).  Larger classes always go to
    // the tensor cores: a handful of them on the large CUDA-core shape took 435 us at HERA-350, next to a 650 us pass.
    pl->tc_sum = !(getenv("CALB2_TC_SUM") && atoi(getenv("CALB2_TC_SUM")) == 0);
    int tc_min = 16;
    if (getenv("CALB2_TC_MIN")) tc_min = std::max(1, atoi(getenv("CALB2_TC_MIN")));
    long long tc_off = 0;
    for (auto& ci : pl->classes) {
      if (!pl->tc_enabled || (ci.nmembers < tc_min && ci.kpt <= 128)) ci.tc = false;
      if (ci.tc) {  // the layout must fit the SM: 512 TMEM columns, 227 KB of shared memory (1 KB of it static)
        const TcLayout L = tc_layout(ci.kpt);
        if (L.col_qlo + 32 * L.nbuf > 512 || L.total + 1024 > 227u * 1024u) ci.tc = false;
      }
      if (ci.tc) {
        ci.tc_off = tc_off;
        tc_off += (long long)pl->ntiles_c * 4 * ci.kpt * SHARED_FT;
        pl->tc_smem_bytes = std::max(pl->tc_smem_bytes, (size_t)tc_layout(ci.kpt).total);
      }
    }
  }
  {
    // CTAs: a class is cut into tiles of MS groups; when that gives too few CTAs to balance 148 SMs (two resident CTAs
    // each for the small shape), every tile is further cut into channel segments whose backward sums go to separate
    // planes of dcpart and are added, in order, by coeffs_kernel.
    int nsm = 148;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, d->device);
    {
      // Tensor-core tiles: one CTA per SM, ~12 000 cycles of prologue + epilogue per CTA, ~3000 cycles per 32-channel tile up to
      // 160 vectors and ~6000 above (measured, profiles/round2_ncu_hera350.md section 7).  The number of channel segments is the
      // candidate with the smallest makespan of a longest-first greedy schedule over the SMs -- at HERA-350 on one GPU that is
      // 1 or 3 (1011 or 3033 CTAs), on an eighth of it 2-3 instead of the CUDA-core rule's 8.
      auto tile_cost = [](int kpt) { return kpt <= 128 ? 3000.0 : (kpt <= 160 ? 3400.0 : 6000.0); };
      double best = 1e300;
      for (int cand : {1, 2, 3, 4, 6, 8}) {
        if (cand > pl->ntiles_c) break;
        std::vector<double> costs;
        for (const auto& ci : pl->classes)
          if (ci.tc)
            for (int m0 = 0; m0 < ci.nmembers; m0 += TcCfg::MS)
              for (int sg = 0; sg < cand; ++sg) {
                const int nt = (int)((long long)pl->ntiles_c * (sg + 1) / cand) - (int)((long long)pl->ntiles_c * sg / cand);
                costs.push_back(12000.0 + nt * tile_cost(ci.kpt));
              }
        if (costs.empty()) break;
        std::sort(costs.begin(), costs.end(), std::greater<double>());
        std::priority_queue<double, std::vector<double>, std::greater<double>> sms;
        for (int i = 0; i < nsm; ++i) sms.push(0.0);
        double makespan = 0.0;
        for (double c : costs) {
          const double t = sms.top() + c;
          sms.pop();
          sms.push(t);
          makespan = std::max(makespan, t);
        }
        if (makespan < best * 0.98) {  // prefer fewer segments on a near tie
          best = makespan;
          pl->nseg_tc = cand;
        }
      }
      if (getenv("CALB2_NSEG_TC")) pl->nseg_tc = std::max(1, std::min(pl->ntiles_c, atoi(getenv("CALB2_NSEG_TC"))));
    }
    for (int v = 0; v < 2; ++v) {
      long long base = 0;
      for (const auto& ci : pl->classes) base += (ci.nmembers + shared_ms(v, ci.kp <= KPM_SMALL ? 0 : 1) - 1) / shared_ms(v, ci.kp <= KPM_SMALL ? 0 : 1);
      int nseg = 1;
      while (nseg < 8 && nseg * 2 <= pl->ntiles_c && base * nseg < (long long)nsm * 12) nseg *= 2;
      if (getenv("CALB2_NSEG")) nseg = std::max(1, std::min(pl->ntiles_c, atoi(getenv("CALB2_NSEG"))));
      pl->nseg[v] = nseg;
      int cs_cursor = 0;
      for (const auto& ci : pl->classes) {
        const int shape = ci.kp <= KPM_SMALL ? 0 : 1;
        const int MSv = shared_ms(v, shape);
        for (int m0 = 0; m0 < ci.nmembers; m0 += MSv)
          for (int sg = 0; sg < nseg; ++sg) {
            MTileDesc mt{};
            mt.a_off = ci.a_off;
            mt.kp = ci.kp;
            mt.ncomp = ci.ncomp;
            mt.nslots = std::min(MSv, ci.nmembers - m0);
            mt.cs0 = cs_cursor + m0;
            mt.j0 = (int)((long long)pl->ntiles_c * sg / nseg);
            mt.j1 = (int)((long long)pl->ntiles_c * (sg + 1) / nseg);
            mt.seg = sg;
            pl->mtiles[v][shape].push_back(mt);
            if (!ci.tc) pl->mt_rest[v][shape].push_back(mt);
          }
        if (v == 0 && ci.tc)  // the same class as 64-group tiles of the tensor-core shape
          for (int m0 = 0; m0 < ci.nmembers; m0 += TcCfg::MS)
            for (int sg = 0; sg < pl->nseg_tc; ++sg) {
              const int nseg = pl->nseg_tc;
              MTileDesc mt{};
              mt.a_off = ci.tc_off;
              mt.kp = ci.kpt;
              mt.ncomp = ci.ncomp;
              mt.nslots = std::min(TcCfg::MS, ci.nmembers - m0);
              mt.cs0 = cs_cursor + m0;
              mt.j0 = (int)((long long)pl->ntiles_c * sg / nseg);
              mt.j1 = (int)((long long)pl->ntiles_c * (sg + 1) / nseg);
              mt.seg = sg;
              pl->mt_tc.push_back(mt);
            }
        cs_cursor += ci.nmembers;
      }
      auto by_cost = [](const MTileDesc& a, const MTileDesc& b) {  // longest first: channel tiles x rows x (8-group blocks in use)
        const long long ca = (long long)(a.j1 - a.j0) * (a.kp + 40) * ((a.nslots + 7) / 8);
        const long long cb = (long long)(b.j1 - b.j0) * (b.kp + 40) * ((b.nslots + 7) / 8);
        return ca > cb;
      };
      for (int shape = 0; shape < 2; ++shape) {
        std::stable_sort(pl->mtiles[v][shape].begin(), pl->mtiles[v][shape].end(), by_cost);
        std::stable_sort(pl->mt_rest[v][shape].begin(), pl->mt_rest[v][shape].end(), by_cost);
      }
      if (v == 0) {

        std::stable_sort(pl->mt_tc.begin(), pl->mt_tc.end(), by_cost);
      }
    }
  }
  pl->slot_row0[ns] = (int)rows;
  pl->rows_total = rows;
  pl->a_floats = heavy_floats + pl->a_class_floats;
  if (nb * (long long)pl->nfp > INT_MAX) {
    delete pl;
    return fail(CALB2_ERR_UNSUPPORTED, "nbls * nfreqs = %lld exceeds the 32-bit element index of the fused kernel", nb * (long long)pl->nfp);
  }
  if (rows > INT_MAX / 4) {
    delete pl;
    return fail(CALB2_ERR_UNSUPPORTED, "too many basis rows (%lld)", rows);
  }

  // ---- coefficient -> row maps, antenna CSR ----
  std::vector<int> coef_row0(nc), coef_grp(nc);
  for (int g = 0; g < d->ngroups; ++g)
    for (int k = 0; k < pl->grp_ncomp[g]; ++k) {
      coef_row0[pl->grp_coef0[g] + k] = pl->slot_row0[pl->grp_slot0[g]] + k;
      coef_grp[pl->grp_coef0[g] + k] = g;
    }
  std::vector<int> ant_ptr(d->nants + 1, 0), ant_ent(2 * nb), ant_partner(2 * nb);
  for (long long b = 0; b < nb; ++b) {
    ant_ptr[pl->bl_ant0[b] + 1]++;
    ant_ptr[pl->bl_ant1[b] + 1]++;
  }
  for (int a = 0; a < d->nants; ++a) ant_ptr[a + 1] += ant_ptr[a];
  {
    std::vector<int> fill(ant_ptr.begin(), ant_ptr.end() - 1);
    for (long long b = 0; b < nb; ++b) {  // ascending baseline order inside every antenna's list
      ant_partner[fill[pl->bl_ant0[b]]] = pl->bl_ant1[b];
      ant_ent[fill[pl->bl_ant0[b]]++] = (int)(b << 1);
      ant_partner[fill[pl->bl_ant1[b]]] = pl->bl_ant0[b];
      ant_ent[fill[pl->bl_ant1[b]]++] = (int)(b << 1) | 1;
    }
  }

  std::vector<SlotGeom> geom(ns);
  for (long long s = 0; s < ns; ++s) {
    const int cls = grp_cls[pl->slot_grp[s]];
    if (cls < 0) {
      const ItemDesc& item = pl->items[pl->slot_item[s]];
      geom[s].a_off = item.a_off;
      geom[s].item_rows = item.nrows;
      geom[s].row_in_item = pl->slot_row0[s] - item.row0;
      geom[s].swz_ft = 0;
    } else {
      geom[s].a_off = pl->classes[cls].a_off;
      geom[s].item_rows = pl->classes[cls].kp;
      geom[s].row_in_item = 0;
      geom[s].swz_ft = SHARED_FT;
    }
    geom[s].nbls = pl->slot_nbls[s];
  }

  // ---- device allocations ----
  int rc = 0;
#define TRY(x)          \
  if (!rc) rc = (x);
  CU(cudaStreamCreateWithFlags(&pl->stream, cudaStreamNonBlocking));
  CU(cudaStreamCreateWithFlags(&pl->stream2, cudaStreamNonBlocking));
  CU(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming));
  if (generic) {
    TRY(upload(pl->d_slot_bl0, pl->slot_bl0, pl));
    TRY(upload(pl->d_bl_ant0, pl->bl_ant0, pl));
    TRY(upload(pl->d_bl_ant1, pl->bl_ant1, pl));
    TRY(upload(pl->d_bl_slot, pl->bl_slot, pl));
    TRY(upload(pl->ant_ptr, ant_ptr, pl));
    TRY(upload(pl->ant_ent, ant_ent, pl));
    TRY(upload(pl->ant_partner, ant_partner, pl));
    TRY(upload(pl->coef_grp, coef_grp, pl));
    TRY(upload(pl->d_grp_nslots, pl->grp_nslots, pl));
    TRY(upload(pl->d_grp_slot0, pl->grp_slot0, pl));
    TRY(upload(pl->d_grp_coef0, pl->grp_coef0, pl));
    TRY(upload(pl->d_grp_ncomp, pl->grp_ncomp, pl));
    if (!rc) {
      if (d->dtype == CALB2_F64) {
        auto* g = new GenericPlan<double>();
        pl->gen = g;
        rc = g->init(pl);
      } else {
        auto* g = new GenericPlan<float>();
        pl->gen = g;
        rc = g->init(pl);
      }
    }
    if (rc) {
      calb2_plan_destroy(pl);
      return rc;
    }
    *out = pl;
    return 0;
  }
  TRY(dalloc(pl->A, (size_t)pl->a_floats, pl));
  TRY(upload(pl->d_items, pl->items, pl));
  TRY(upload(pl->row_slot, row_slot, pl));
  TRY(upload(pl->row_coef, row_coef, pl));
  TRY(upload(pl->d_slot_row0, pl->slot_row0, pl));
  TRY(upload(pl->d_slot_bl0, pl->slot_bl0, pl));
  TRY(upload(pl->d_slot_nb, pl->slot_nbls, pl));
  for (int v = 0; v < 2; ++v)
    for (int shape = 0; shape < 2; ++shape) TRY(upload(pl->d_mtiles[v][shape], pl->mtiles[v][shape], pl));
  TRY(upload(pl->d_cslots, cslots, pl));
  TRY(upload(pl->d_cs_slot, cs_slot, pl));
  TRY(upload(pl->d_bl_ant0, pl->bl_ant0, pl));
  TRY(upload(pl->d_bl_ant1, pl->bl_ant1, pl));
  TRY(upload(pl->d_bl_slot, pl->bl_slot, pl));
  TRY(upload(pl->ant_ptr, ant_ptr, pl));
  TRY(upload(pl->ant_ent, ant_ent, pl));
  TRY(upload(pl->ant_partner, ant_partner, pl));
  TRY(upload(pl->coef_row0, coef_row0, pl));
  TRY(upload(pl->coef_grp, coef_grp, pl));
  TRY(upload(pl->d_grp_nslots, pl->grp_nslots, pl));
  TRY(upload(pl->d_grp_slot0, pl->grp_slot0, pl));
  TRY(upload(pl->d_grp_coef0, pl->grp_coef0, pl));
  TRY(upload(pl->d_grp_ncomp, pl->grp_ncomp, pl));
  TRY(upload(pl->slot_geom, geom, pl));
  const size_t nd = (size_t)nb * pl->nfp, ng = (size_t)d->nants * pl->nfp;
  TRY(dalloc(pl->d_r, nd, pl));
  TRY(dalloc(pl->d_i, nd, pl));
  TRY(dalloc(pl->w, nd, pl));
  TRY(dalloc(pl->z, nd, pl));
  for (int b = 0; b < 2; ++b) {
    TRY(dalloc(pl->g_r[b], ng, pl));
    TRY(dalloc(pl->g_i[b], ng, pl));
  }
  TRY(dalloc(pl->gm_r, ng, pl));
  TRY(dalloc(pl->gu_r, ng, pl));
  TRY(dalloc(pl->gm_i, ng, pl));
  TRY(dalloc(pl->gu_i, ng, pl));
  TRY(dalloc(pl->c_r, (size_t)nc, pl));
  TRY(dalloc(pl->c_i, (size_t)nc, pl));
  TRY(dalloc(pl->cm_r, (size_t)nc, pl));
  TRY(dalloc(pl->cu_r, (size_t)nc, pl));
  TRY(dalloc(pl->cm_i, (size_t)nc, pl));
  TRY(dalloc(pl->cu_i, (size_t)nc, pl));
  pl->dc_plane = rows * 4;
  TRY(dalloc(pl->dcpart, (size_t)pl->dc_plane * std::max(std::max(pl->nseg[0], pl->nseg[1]), pl->nseg_tc), pl));
  TRY(dalloc(pl->partials, (size_t)n_partials_max(pl) * 4, pl));
  TRY(upload(pl->d_mt_tc, pl->mt_tc, pl));
  for (int v = 0; v < 2; ++v)
    for (int shape = 0; shape < 2; ++shape) TRY(upload(pl->d_mt_rest[v][shape], pl->mt_rest[v][shape], pl));
  {
    long long tc_floats = 0;
    for (const auto& ci : pl->classes)
      if (ci.tc) tc_floats = std::max(tc_floats, ci.tc_off + (long long)pl->ntiles_c * 4 * ci.kpt * SHARED_FT);
    TRY(dalloc(pl->At, (size_t)tc_floats, pl));
  }
  TRY(dalloc(pl->red_d, 4096, pl));
  TRY(dalloc(pl->dbg_out, 16, pl));
  TRY(dalloc(pl->state, 1, pl));
  TRY(dalloc(pl->state_eval, 1, pl));
  TRY(dalloc(pl->comm_scalars, 4, pl));
  pl->staging_floats = (size_t)64 << 20;  // 256 MiB of floats staged per batch
  pl->staging_floats = std::max(pl->staging_floats, (size_t)pl->nf * 512);
  TRY(dalloc(pl->staging, pl->staging_floats, pl, false));
#undef TRY
  if (!rc && cudaMallocHost(&pl->h_staging, pl->staging_floats * sizeof(float)) != cudaSuccess)
    rc = fail(CALB2_ERR_CUDA, "cudaMallocHost(staging) failed");
  if (!rc && cudaMallocHost(&pl->h_state, sizeof(FitState)) != cudaSuccess)
    rc = fail(CALB2_ERR_CUDA, "cudaMallocHost(state) failed");
  if (!rc && pl->tc_enabled) {
    if (cudaHostAlloc(&pl->tc_dbg, 64, cudaHostAllocMapped) != cudaSuccess)
      rc = fail(CALB2_ERR_CUDA, "cudaHostAlloc(tc_dbg) failed");
    else {
      memset(pl->tc_dbg, 0, 64);
      g_tc_dbg = pl->tc_dbg;
    }
    if (const char* e = getenv("CALB2_TC_PROF")) {
      pl->tc_prof_cta = atoi(e);
      if (!rc) rc = dalloc(pl->tc_prof, (size_t)TC_PROF_SLOTS * TC_PROF_TILES, pl);
    }
  }
  if (rc) {
    calb2_plan_destroy(pl);
    return rc;
  }
  *out = pl;
  return 0;
}

int calb2_plan_destroy(calb2_plan* pl) {
  if (!pl) return 0;
  cudaSetDevice(pl->device);
  if (pl->stream) cudaStreamSynchronize(pl->stream);
  delete pl->gen;
  pl->gen = nullptr;
  if (pl->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(pl->comm);
  if (pl->peers_open) calb2_comm_peer_close(pl);
  if (pl->xbuf) cudaFree(pl->xbuf);
  DevBuf<float>* fb[] = {&pl->A, &pl->d_r, &pl->d_i, &pl->w, &pl->g_r[0], &pl->g_r[1], &pl->g_i[0], &pl->g_i[1],
                         &pl->gm_r, &pl->gu_r, &pl->gm_i, &pl->gu_i, &pl->gsnap_r, &pl->gsnap_i, &pl->ggrad_r,
                         &pl->ggrad_i, &pl->c_r, &pl->c_i, &pl->cm_r, &pl->cu_r, &pl->cm_i, &pl->cu_i, &pl->csnap_r,
                         &pl->csnap_i, &pl->cgrad_r, &pl->cgrad_i, &pl->dcpart, &pl->hist, &pl->scratch_f, &pl->staging};
  if (pl->ggrad_i.n == 0) pl->ggrad_i.p = nullptr;  // (NCCL exchange) a view into ggrad_r
  for (auto* b : fb) b->release();
  pl->z.release();
  pl->y.release();
  pl->vout.release();
  pl->partials.release();
  pl->red_d.release();
  pl->dbg_out.release();
  pl->comm_scalars.release();
  pl->light_partials.release();
  pl->d_items.release();
  for (int v = 0; v < 2; ++v)
    for (int shape = 0; shape < 2; ++shape) pl->d_mtiles[v][shape].release();
  pl->d_mt_tc.release();
  for (int v = 0; v < 2; ++v)
    for (int shape = 0; shape < 2; ++shape) pl->d_mt_rest[v][shape].release();
  pl->At.release();
  pl->d_cslots.release();
  pl->d_cs_slot.release();
  pl->d_slot_nb.release();
  pl->row_slot.release();
  DevBuf<int>* ib[] = {&pl->row_coef, &pl->d_slot_row0, &pl->d_slot_bl0, &pl->d_bl_ant0, &pl->d_bl_ant1, &pl->d_bl_slot,
                       &pl->ant_ptr, &pl->ant_ent, &pl->ant_partner, &pl->coef_row0, &pl->coef_grp, &pl->d_grp_nslots, &pl->d_grp_slot0,
                       &pl->d_grp_coef0, &pl->d_grp_ncomp};
  for (auto* b : ib) b->release();
  pl->state.release();
  pl->state_eval.release();
  pl->slot_geom.release();
  pl->sky_r.release();
  pl->sky_i.release();
  if (pl->h_staging) cudaFreeHost(pl->h_staging);
  if (pl->h_state) cudaFreeHost(pl->h_state);
  if (pl->tc_dbg) {
    if (g_tc_dbg == pl->tc_dbg) g_tc_dbg = nullptr;
    cudaFreeHost(pl->tc_dbg);
  }
  pl->tail_counter.release();
  if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
  if (pl->ev_join) cudaEventDestroy(pl->ev_join);
  if (pl->stream2) cudaStreamDestroy(pl->stream2);
  if (pl->stream) cudaStreamDestroy(pl->stream);
  delete pl;
  return 0;
}

int calb2_plan_get_info(const calb2_plan* pl, calb2_plan_info* info) {
  if (!pl || !info) return fail(CALB2_ERR_ARG, "null argument");
  info->n_d = pl->nbls * pl->nf;
  info->n_a_nz = pl->n_a_nz;
  info->n_a_stored = pl->a_floats;
  info->n_c_nz = pl->ncoef;
  info->nbls_total = pl->nbls;
  info->nslots_total = pl->nslots;
  info->nitems = (int64_t)pl->items.size();
  info->tile_freqs = pl->FT;
  info->rows_per_item_max = pl->KMAX;
  info->device_bytes = (int64_t)pl->device_bytes;
  info->generic = pl->gen ? 1 : 0;
  info->dtype = pl->dtype;
  info->n_classes = (int64_t)pl->classes.size();
  info->n_class_slots = pl->nslots_class;
  info->n_class_ctas = (int64_t)(pl->mtiles[0][0].size() + pl->mtiles[0][1].size());
  info->n_tc_ctas = pl->tc_enabled ? (int64_t)pl->mt_tc.size() : 0;
  info->n_tc_slots = 0;
  if (pl->tc_enabled)
    for (const auto& ci : pl->classes)
      if (ci.tc) info->n_tc_slots += ci.nmembers;
  info->n_a_class = pl->a_class_floats;
  info->n_a_class_nz = 0;
  info->class_fma = 0;
  for (const auto& ci : pl->classes) {
    info->n_a_class_nz += (int64_t)ci.ncomp * ci.nmembers * pl->nf;  // what the streaming path would read for these groups
    info->class_fma += (int64_t)4 * ci.ncomp * ci.nmembers * pl->nf; // forward + backward, real + imaginary part
  }
  if (pl->gen) {
    info->n_a_stored = pl->n_a_nz;
    info->nitems = 0;
    info->tile_freqs = 0;
    info->rows_per_item_max = 0;
    info->device_bytes += (int64_t)pl->gen->device_bytes();
  }
  return 0;
}

int calb2_plan_set_basis(calb2_plan* pl, int32_t g0, int32_t ng, const void* const* blocks_v) {
  if (!pl || !blocks_v) return fail(CALB2_ERR_ARG, "null argument");
  if (g0 < 0 || ng < 0 || g0 + ng > pl->ngroups) return fail(CALB2_ERR_ARG, "group range out of bounds");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->set_basis(g0, ng, blocks_v);
  const float* const* blocks = reinterpret_cast<const float* const*>(blocks_v);
  std::vector<RetileJob> jobs;
  std::vector<TcRetileJob> tc_jobs;
  std::unordered_map<const float*, long long> seen;
  size_t used = 0;
  DevBuf<RetileJob> djobs;
  DevBuf<TcRetileJob> d_tc_jobs;
  auto flush = [&]() -> int {
    if (jobs.empty()) return 0;
    CU(cudaMemcpyAsync(pl->staging.p, pl->h_staging, used * sizeof(float), cudaMemcpyHostToDevice, pl->stream));
    if (djobs.n < jobs.size()) CU(djobs.alloc(jobs.size() * 2));
    CU(cudaMemcpyAsync(djobs.p, jobs.data(), jobs.size() * sizeof(RetileJob), cudaMemcpyHostToDevice, pl->stream));
    retile_kernel<<<(unsigned)jobs.size(), 256, 0, pl->stream>>>(pl->staging.p, pl->A.p, djobs.p, pl->nf, pl->FT);
    CU(cudaGetLastError());
    if (!tc_jobs.empty()) {
      if (d_tc_jobs.n < tc_jobs.size()) CU(d_tc_jobs.alloc(tc_jobs.size() * 2));
      CU(cudaMemcpyAsync(d_tc_jobs.p, tc_jobs.data(), tc_jobs.size() * sizeof(TcRetileJob), cudaMemcpyHostToDevice, pl->stream));
      retile_tc_kernel<<<(unsigned)tc_jobs.size(), 256, 0, pl->stream>>>(pl->staging.p, pl->At.p, d_tc_jobs.p, pl->nf);
      CU(cudaGetLastError());
    }
    CU(cudaStreamSynchronize(pl->stream));
    jobs.clear();
    tc_jobs.clear();
    seen.clear();
    used = 0;
    return 0;
  };
  for (int gi = 0; gi < ng; ++gi) {
    const int g = g0 + gi;
    const float* blk = blocks[gi];
    const int ncomp = pl->grp_ncomp[g], nsl = pl->grp_nslots[g];
    const size_t need = (size_t)ncomp * nsl * pl->nf;
    if (need == 0) continue;
    if (!blk) return fail(CALB2_ERR_ARG, "group %d: null basis block", g);
    if (need > pl->staging_floats) return fail(CALB2_ERR_UNSUPPORTED, "group %d basis block exceeds the staging buffer", g);
    const int cls = pl->grp_cls[g];
    if (cls >= 0 && pl->classes[cls].uploaded) continue;  // shared-basis class: stored once, by its first member
    long long base;
    auto it = seen.find(blk);
    if (it != seen.end()) {
      base = it->second;
    } else {
      if (used + need > pl->staging_floats)
        if (int r = flush()) return r;
      memcpy(pl->h_staging + used, blk, need * sizeof(float));
      base = (long long)used;
      seen.emplace(blk, base);
      used += need;
    }
    if (cls >= 0) {
      RetileJob jb{};
      jb.src_off = base;
      jb.dst_off = pl->classes[cls].a_off;
      jb.ncomp = ncomp;
      jb.item_rows = pl->classes[cls].kp;
      jb.row_in_item = 0;
      jb.swz_ft = SHARED_FT;
      jobs.push_back(jb);
      if (pl->classes[cls].tc) {
        TcRetileJob tj{};
        tj.src_off = base;
        tj.dst_off = pl->classes[cls].tc_off;
        tj.ncomp = ncomp;
        tj.kpt = pl->classes[cls].kpt;
        tc_jobs.push_back(tj);
      }
      pl->classes[cls].uploaded = true;
      continue;
    }
    for (int s = 0; s < nsl; ++s) {
      const int slot = pl->grp_slot0[g] + s;
      const ItemDesc& item = pl->items[pl->slot_item[slot]];
      RetileJob jb{};
      jb.src_off = base + (long long)s * ncomp * pl->nf;
      jb.dst_off = item.a_off;
      jb.ncomp = ncomp;
      jb.item_rows = item.nrows;
      jb.row_in_item = pl->slot_row0[slot] - item.row0;
      jb.swz_ft = 0;
      jobs.push_back(jb);
    }
  }
  if (int r = flush()) return r;
  djobs.release();
  d_tc_jobs.release();
  pl->basis_groups_set += ng;
  return 0;
}

int calb2_set_integration(calb2_plan* pl, const void* data_r_v, const void* data_i_v, const void* wgts_v) {
  if (!pl || !data_r_v || !data_i_v || !wgts_v) return fail(CALB2_ERR_ARG, "null argument");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->set_integration(data_r_v, data_i_v, wgts_v);
  const float *data_r = (const float*)data_r_v, *data_i = (const float*)data_i_v, *wgts = (const float*)wgts_v;
  if (int r = upload_padded(pl, data_r, pl->d_r.p, (size_t)pl->nbls, 0.f)) return r;
  if (int r = upload_padded(pl, data_i, pl->d_i.p, (size_t)pl->nbls, 0.f)) return r;
  if (int r = upload_padded(pl, wgts, pl->w.p, (size_t)pl->nbls, 0.f)) return r;
  pl->have_data = true;
  return 0;
}

int calb2_set_gains(calb2_plan* pl, const void* g_r_v, const void* g_i_v) {
  if (!pl || !g_r_v || !g_i_v) return fail(CALB2_ERR_ARG, "null argument");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->set_gains(g_r_v, g_i_v);
  const float *g_r = (const float*)g_r_v, *g_i = (const float*)g_i_v;
  pl->cur_buf = 0;
  if (int r = upload_padded(pl, g_r, pl->g_r[0].p, (size_t)pl->nants, 1.f)) return r;
  if (int r = upload_padded(pl, g_i, pl->g_i[0].p, (size_t)pl->nants, 0.f)) return r;
  pl->have_gains = true;
  return 0;
}

int calb2_set_coeffs(calb2_plan* pl, const void* coef_r, const void* coef_i) {
  if (!pl || !coef_r || !coef_i) return fail(CALB2_ERR_ARG, "null argument");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->set_coeffs(coef_r, coef_i);
  CU(cudaMemcpy(pl->c_r.p, coef_r, pl->ncoef * sizeof(float), cudaMemcpyHostToDevice));
  CU(cudaMemcpy(pl->c_i.p, coef_i, pl->ncoef * sizeof(float), cudaMemcpyHostToDevice));
  pl->have_coeffs = true;
  return 0;
}

int calb2_get_gains(calb2_plan* pl, void* g_r_v, void* g_i_v) {
  if (!pl || !g_r_v || !g_i_v) return fail(CALB2_ERR_ARG, "null argument");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->get_gains(g_r_v, g_i_v);
  float *g_r = (float*)g_r_v, *g_i = (float*)g_i_v;
  if (int r = download_unpadded(pl, pl->g_r[pl->cur_buf].p, g_r, (size_t)pl->nants)) return r;
  return download_unpadded(pl, pl->g_i[pl->cur_buf].p, g_i, (size_t)pl->nants);
}

int calb2_get_coeffs(calb2_plan* pl, void* coef_r, void* coef_i) {
  if (!pl || !coef_r || !coef_i) return fail(CALB2_ERR_ARG, "null argument");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->get_coeffs(coef_r, coef_i);
  CU(cudaStreamSynchronize(pl->stream));
  CU(cudaMemcpy(coef_r, pl->c_r.p, pl->ncoef * sizeof(float), cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(coef_i, pl->c_i.p, pl->ncoef * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int calb2_get_weights(calb2_plan* pl, void* wgts) {
  if (!pl || !wgts) return fail(CALB2_ERR_ARG, "null argument");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->get_weights(wgts);
  return download_unpadded(pl, pl->w.p, (float*)wgts, (size_t)pl->nbls);
}

static int run_forward_store_v(calb2_plan* pl) {
  if (int r = ensure_vout(pl)) return r;
  if (int r = set_eval_state(pl)) return r;
  HeavyParams hp = heavy_params(pl, pl->state_eval.p, false, 1, 0);
  CU(launch_heavy(pl, false, hp, (int)pl->items.size(), pl->stream));
  return 0;
}

int calb2_get_model(calb2_plan* pl, void* model_r_v, void* model_i_v) {
  if (!pl || !model_r_v || !model_i_v) return fail(CALB2_ERR_ARG, "null argument");
  if (pl->gen) {
    CU(cudaSetDevice(pl->device));
    return pl->gen->get_model(model_r_v, model_i_v);
  }
  float *model_r = (float*)model_r_v, *model_i = (float*)model_i_v;
  if (!pl->have_data || !pl->have_gains || !pl->have_coeffs) return fail(CALB2_ERR_STATE, "integration, gains and coefficients must be set first");
  CU(cudaSetDevice(pl->device));
  if (int r = run_forward_store_v(pl)) return r;
  // gather per baseline through the staging buffer: real rows then imaginary rows
  const size_t rows_per = std::max<size_t>(1, pl->staging_floats / ((size_t)2 * pl->nf));
  for (size_t b0 = 0; b0 < (size_t)pl->nbls; b0 += rows_per) {
    const size_t n = std::min(rows_per, (size_t)pl->nbls - b0);
    float* sr = pl->staging.p;
    float* si = pl->staging.p + n * pl->nf;
    model_gather_kernel<<<(unsigned)n, 128, 0, pl->stream>>>(pl->vout.p, pl->d_bl_slot.p + b0, sr, si, pl->nf, pl->nfp);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(pl->h_staging, pl->staging.p, 2 * n * pl->nf * sizeof(float), cudaMemcpyDeviceToHost, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    memcpy(model_r + b0 * pl->nf, pl->h_staging, n * pl->nf * sizeof(float));
    memcpy(model_i + b0 * pl->nf, pl->h_staging + n * pl->nf, n * pl->nf * sizeof(float));
  }
  return 0;
}

static int device_sum(calb2_plan* pl, const float* x, const float* y, size_t n, double* out) {
  const int nb = 1024;
  dot_partial_kernel<<<nb, 256, 0, pl->stream>>>(x, y, n, pl->red_d.p);
  CU(cudaGetLastError());
  std::vector<double> h(nb);
  CU(cudaMemcpyAsync(h.data(), pl->red_d.p, nb * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
  CU(cudaStreamSynchronize(pl->stream));
  double t = 0.0;
  for (double v : h) t += v;
  *out = t;
  return 0;
}

int calb2_prior_sums(calb2_plan* pl, const void* sky_r_v, const void* sky_i_v, double* prior_r, double* prior_i) {
  if (!pl || !sky_r_v || !sky_i_v || !prior_r || !prior_i) return fail(CALB2_ERR_ARG, "null argument");
  if (pl->gen) {
    CU(cudaSetDevice(pl->device));
    return pl->gen->prior_sums(sky_r_v, sky_i_v, prior_r, prior_i);
  }
  const float *sky_r = (const float*)sky_r_v, *sky_i = (const float*)sky_i_v;
  if (!pl->have_data) return fail(CALB2_ERR_STATE, "set_integration first (weights)");
  CU(cudaSetDevice(pl->device));
  const size_t nd = (size_t)pl->nbls * pl->nfp;
  if (pl->scratch_f.n < nd)
    if (int r = dalloc(pl->scratch_f, nd, pl)) return r;
  double t = 0.0;
  if (int r = upload_padded(pl, sky_r, pl->scratch_f.p, (size_t)pl->nbls, 0.f)) return r;
  if (int r = device_sum(pl, pl->scratch_f.p, pl->w.p, nd, &t)) return r;
  *prior_r = (double)(float)t;
  if (int r = upload_padded(pl, sky_i, pl->scratch_f.p, (size_t)pl->nbls, 0.f)) return r;
  if (int r = device_sum(pl, pl->scratch_f.p, pl->w.p, nd, &t)) return r;
  *prior_i = (double)(float)t;
  return 0;
}

int calb2_apply_model_snr_weights(calb2_plan* pl) {
  if (!pl) return fail(CALB2_ERR_ARG, "null argument");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->apply_snr_weights();
  if (!pl->have_data || !pl->have_coeffs || !pl->have_gains) return fail(CALB2_ERR_STATE, "integration, gains and coefficients must be set first");
  if (int r = run_forward_store_v(pl)) return r;
  snr_weight_kernel<<<(unsigned)pl->nbls, 128, 0, pl->stream>>>(pl->w.p, pl->vout.p, pl->d_bl_slot.p, pl->nfp, 1.f);
  CU(cudaGetLastError());
  double t = 0.0;
  const size_t nd = (size_t)pl->nbls * pl->nfp;
  if (int r = device_sum(pl, pl->w.p, nullptr, nd, &t)) return r;
  scale_kernel<<<1024, 256, 0, pl->stream>>>(pl->w.p, nd, (float)t);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(pl->stream));
  return 0;
}

int calb2_init_coeffs(calb2_plan* pl, const void* sky_r, const void* sky_i) {
  if (!pl || !sky_r || !sky_i) return fail(CALB2_ERR_ARG, "null argument");
  CU(cudaSetDevice(pl->device));
  if (pl->gen) return pl->gen->init_coeffs(sky_r, sky_i);
  if (!pl->have_data) return fail(CALB2_ERR_STATE, "set_integration first (weights)");
  return init_coeffs_impl(pl, (const float*)sky_r, (const float*)sky_i);
}

int calb2_loss_and_grads(calb2_plan* pl, int32_t regularization, double prior_r, double prior_i, double* loss, void* dg_r_v,
                         void* dg_i_v, void* dc_r_v, void* dc_i_v) {
  if (!pl) return fail(CALB2_ERR_ARG, "null argument");
  if (pl->gen) {
    CU(cudaSetDevice(pl->device));
    return pl->gen->loss_and_grads(regularization, prior_r, prior_i, loss, dg_r_v, dg_i_v, dc_r_v, dc_i_v);
  }
  float *dg_r = (float*)dg_r_v, *dg_i = (float*)dg_i_v, *dc_r = (float*)dc_r_v, *dc_i = (float*)dc_i_v;
  if (!pl->have_data || !pl->have_gains || !pl->have_coeffs) return fail(CALB2_ERR_STATE, "integration, gains and coefficients must be set first");
  CU(cudaSetDevice(pl->device));
  const bool sum = regularization == CALB2_REG_SUM;
  if (sum)
    if (int r = ensure_sum_buffers(pl)) return r;
  if (int r = ensure_grad_buffers(pl)) return r;
  if (int r = set_eval_state(pl)) return r;
  FitConsts k{};
  k.regularization = regularization;
  k.prior_r = (float)prior_r;
  k.prior_i = (float)prior_i;
  HeavyParams hp = heavy_params(pl, pl->state_eval.p, sum, 0, 0);
  CU(prepare_dcpart(pl, tc_pass_of(pl, sum) ? (sum ? 4 : 3) : (sum ? 2 : 1)));
  CU(launch_heavy(pl, sum, hp, (int)pl->items.size(), pl->stream));
  FinalizeParams fp{};
  fp.partials = pl->partials.p;
  fp.nitems = n_partials(pl, sum);
  fp.st = pl->state_eval.p;
  fp.k = k;
  fp.eval_only = 1;
  fp.peers.n = 0;
  fp.xwait = 0;
  fp.xpar = 0;
  fp.timeout_ns = pl->xtimeout_ns;
  finalize_kernel<<<1, 1024, 0, pl->stream>>>(fp);
  CU(cudaGetLastError());
  dim3 ggrid(pl->nants, (pl->nfp + GK_CH - 1) / GK_CH);
  gains_kernel<<<ggrid, GK_THREADS, 0, pl->stream>>>(gains_params(pl, pl->state_eval.p, k, 1, sum, 1));
  CU(cudaGetLastError());
  coeffs_kernel<<<(unsigned)((pl->ncoef + 255) / 256), 256, 0, pl->stream>>>(coeff_params(pl, pl->state_eval.p, k, 1, sum, tc_pass_of(pl, sum)));
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(pl->h_state, pl->state_eval.p, sizeof(FitState), cudaMemcpyDeviceToHost, pl->stream));
  CU(cudaStreamSynchronize(pl->stream));
  if (loss) *loss = (double)pl->h_state->last_loss;
  if (dg_r)
    if (int r = download_unpadded(pl, pl->ggrad_r.p, dg_r, (size_t)pl->nants)) return r;
  if (dg_i)
    if (int r = download_unpadded(pl, pl->ggrad_i.p, dg_i, (size_t)pl->nants)) return r;
  if (dc_r) CU(cudaMemcpy(dc_r, pl->cgrad_r.p, pl->ncoef * sizeof(float), cudaMemcpyDeviceToHost));
  if (dc_i) CU(cudaMemcpy(dc_i, pl->cgrad_i.p, pl->ncoef * sizeof(float), cudaMemcpyDeviceToHost));
  return 0;
}

int calb2_plan_set_variables(calb2_plan* pl, int32_t nvars, const int64_t* coef_bounds) {
  if (!pl || nvars < 1 || !coef_bounds) return fail(CALB2_ERR_ARG, "null argument or no variables");
  const long long ncoef = (long long)pl->ncoef;
  if (coef_bounds[0] != 0 || coef_bounds[nvars] != ncoef) return fail(CALB2_ERR_ARG, "coef_bounds must run from 0 to n_c_nz = %lld", ncoef);
  for (int v = 0; v < nvars; ++v)
    if (coef_bounds[v + 1] <= coef_bounds[v]) return fail(CALB2_ERR_ARG, "coef_bounds must be strictly ascending (variable %d)", v);
  if (pl->gen) return 0;  // the generic path has no optimizer that uses them
  pl->var_bounds.assign(coef_bounds, coef_bounds + nvars + 1);
  pl->var_uploaded = false;
  return 0;
}

int calb2_fit(calb2_plan* pl, const calb2_fit_options* o, void* loss_history_v, calb2_fit_result* res) {
  if (!pl || !o || !res) return fail(CALB2_ERR_ARG, "null argument");
  if (pl->gen) {
    if (o->optimizer == CALB2_OPT_LAMB) return fail(CALB2_ERR_UNSUPPORTED, "LAMB runs on float32 plans only");
    if (o->optimizer < 0 || o->optimizer > CALB2_OPT_FTRL) return fail(CALB2_ERR_ARG, "unknown optimizer id %d", o->optimizer);
    if (o->maxsteps < 0 || o->n_profile_steps < 0) return fail(CALB2_ERR_ARG, "negative step count");
    if (o->maxsteps > 0 && !loss_history_v) return fail(CALB2_ERR_ARG, "loss_history is null");
    CU(cudaSetDevice(pl->device));
    return pl->gen->fit(o, loss_history_v, res);
  }
  float* loss_history = (float*)loss_history_v;
  if (!pl->have_data || !pl->have_gains || !pl->have_coeffs) return fail(CALB2_ERR_STATE, "integration, gains and coefficients must be set first");
  if (o->optimizer < 0 || o->optimizer > CALB2_OPT_LAMB) return fail(CALB2_ERR_ARG, "unknown optimizer id %d", o->optimizer);
  if (o->optimizer == CALB2_OPT_LAMB && pl->nranks > 1)
    return fail(CALB2_ERR_UNSUPPORTED, "LAMB's per-variable norms are not exchanged between ranks: single-GPU plans only");
  if (o->maxsteps < 0 || o->n_profile_steps < 0) return fail(CALB2_ERR_ARG, "negative step count");
  if (o->maxsteps > 0 && !loss_history) return fail(CALB2_ERR_ARG, "loss_history is null");
  CU(cudaSetDevice(pl->device));
  if (o->optimizer == CALB2_OPT_LAMB)
    if (int r = ensure_lamb_buffers(pl)) return r;
  const bool sum = o->regularization == CALB2_REG_SUM;
  const bool freeze = o->freeze_model != 0;
  if (sum)
    if (int r = ensure_sum_buffers(pl)) return r;
  if (o->use_min)
    if (int r = ensure_use_min_buffers(pl)) return r;
  if (pl->nranks > 1)
    if (int r = ensure_grad_buffers(pl)) return r;
  FitConsts k{};
  k.optimizer = o->optimizer;
  k.lr = (float)o->learning_rate;
  k.beta1 = (float)o->beta_1;
  k.beta2 = (float)o->beta_2;
  k.eps = (float)o->epsilon;
  k.rho = (float)o->rho;
  k.momentum = (float)o->momentum;
  k.init_acc = (float)o->initial_accumulator_value;
  k.l1 = (float)o->l1_regularization_strength;
  k.l2 = (float)o->l2_regularization_strength;
  k.lr_power = (float)o->learning_rate_power;
  k.nesterov = o->nesterov;
  k.weight_decay = (float)o->weight_decay;
  k.maxsteps = o->maxsteps;
  k.tol = o->tol;
  k.use_min = o->use_min;
  k.regularization = o->regularization;
  k.prior_r = (float)o->prior_r_sum;
  k.prior_i = (float)o->prior_i_sum;
  k.n_skip = o->n_profile_steps + 1;
  const long long total = (long long)k.n_skip + o->maxsteps;

  // gains must start in buffer 0 (the step parity selects the buffer)
  const size_t ng = (size_t)pl->nants * pl->nfp;
  if (pl->cur_buf == 1) {
    CU(cudaMemcpyAsync(pl->g_r[0].p, pl->g_r[1].p, ng * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
    CU(cudaMemcpyAsync(pl->g_i[0].p, pl->g_i[1].p, ng * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
    pl->cur_buf = 0;
  }
  // fresh optimizer per integration (calibration.py:571): zero slots and step counter
  DevBuf<float>* slots[] = {&pl->gm_r, &pl->gu_r, &pl->gm_i, &pl->gu_i, &pl->cm_r, &pl->cu_r, &pl->cm_i, &pl->cu_i};
  for (auto* s : slots) CU(cudaMemsetAsync(s->p, 0, s->bytes(), pl->stream));
  // accumulators that start at initial_accumulator_value: Adagrad keeps it in slot u, Ftrl in slot m
  if (k.optimizer == CALB2_OPT_ADAGRAD || k.optimizer == CALB2_OPT_FTRL) {
    DevBuf<float>* acc_u[] = {&pl->gu_r, &pl->gu_i, &pl->cu_r, &pl->cu_i};
    DevBuf<float>* acc_m[] = {&pl->gm_r, &pl->gm_i, &pl->cm_r, &pl->cm_i};
    for (auto* s : (k.optimizer == CALB2_OPT_ADAGRAD ? acc_u : acc_m)) {
      if (s->n) fill_kernel<<<256, 256, 0, pl->stream>>>(s->p, s->n, k.init_acc);
      CU(cudaGetLastError());
    }
  }
  if (pl->hist.n < (size_t)std::max(1, o->maxsteps)) {
    if (int r = dalloc(pl->hist, (size_t)std::max(1, o->maxsteps), pl)) return r;
  }
  if (freeze) {
    if (!pl->light_blocks) {
      int nsm = 148;
      cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, pl->device);
      pl->light_blocks = nsm * 8;
      if (int r = dalloc(pl->light_partials, (size_t)pl->light_blocks * 4, pl)) return r;
    }
    if (int r = run_forward_store_v(pl)) return r;  // v = sum_k c_k A_k, once: the coefficients are frozen
  }
#ifdef CALB2_PROFILE
  if (getenv("CALB2_DBG") && (atoi(getenv("CALB2_DBG")) & 1024) && pl->nranks > 1 && !sum) {
    pl->stage_ev.resize((size_t)total * NSTAGE);
    for (auto& e : pl->stage_ev) cudaEventCreate(&e);
    pl->stage_cursor = 0;
  }
#endif
  if (!freeze) CU(prepare_dcpart(pl, tc_pass_of(pl, sum) ? (sum ? 4 : 3) : (sum ? 2 : 1)));
  FitState s0{};
  s0.step = 0;
  s0.stop_after = (int)(total - 1);
  s0.min_loss = INFINITY;
  CU(cudaMemcpyAsync(pl->state.p, &s0, sizeof(s0), cudaMemcpyHostToDevice, pl->stream));
  // Peer exchange: every rank enters the fit through one publish / wait round.  The exchange buffers are double-buffered by
  // the parity of the step sequence number, so the first step of this fit may only overwrite a half once every peer has
  // finished the previous fit's last read of it -- nothing else orders the ranks between two fits.
  if (pl->nranks > 1 && pl->peers.n > 1)
    if (int r = peer_barrier(pl)) return r;

  int chunk = o->steps_per_sync > 0 ? o->steps_per_sync : 32;
  // use_graph: 1 = replay a captured graph, -1 = never, 0 = automatic: small problems are launch-latency bound (four
  // launches of a few microseconds each per iteration), so they replay a graph of `chunk` iterations
  // (basis under 256 MB AND under 4 M visibilities: with the shared-basis layout the basis alone no longer says "small")
  const bool small_problem = (size_t)pl->a_floats * sizeof(float) < ((size_t)256 << 20) && pl->nbls * (long long)pl->nfp < (4ll << 20);
  const bool want_graph = pl->nranks == 1 && (o->use_graph > 0 || (o->use_graph == 0 && small_problem && total >= 64));
  const bool time_heavy = !want_graph;
  std::vector<cudaEvent_t> evs;
  cudaEvent_t ev_begin, ev_end;
  CU(cudaEventCreate(&ev_begin));
  CU(cudaEventCreate(&ev_end));
  if (time_heavy) {
    evs.resize(2 * chunk);
    for (auto& e : evs) CU(cudaEventCreate(&e));
  }
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t gexec = nullptr;
  long long launches = 0, heavy_launches = 0;
  double heavy_ms = 0.0;
  int rc = 0;
  if (want_graph) {
    long long dummy = 0;
    CU(cudaStreamBeginCapture(pl->stream, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < chunk && !rc; ++i) rc = enqueue_step(pl, k, sum, freeze, o->fuse_tail_update != 0, pl->hist.p, nullptr, nullptr, &dummy);
    cudaError_t ce = cudaStreamEndCapture(pl->stream, &graph);
    if (rc) return rc;
    CU(ce);
    CU(cudaGraphInstantiate(&gexec, graph, 0));
  }
  CU(cudaEventRecord(ev_begin, pl->stream));
  nvtxRangePushA("calb2_fit: step loop");
  long long done = 0;
  while (done < total) {
    const int n = (int)std::min<long long>(chunk, total - done);
    if (gexec) {
      CU(cudaGraphLaunch(gexec, pl->stream));
      launches += (long long)chunk * 4;
      heavy_launches += chunk;
    } else {
      for (int i = 0; i < n; ++i) {
        // the reference traces its n_profile_steps extra steps (tf.profiler.experimental.Trace, calibration.py:681-687):
        // here they, the unrecorded warm-up step and the first recorded step get an NVTX range each (enqueue time)
        const long long stepno = done + i;
        const bool ranged = stepno <= (long long)o->n_profile_steps + 1;
        if (ranged) {
          char name[64];
          snprintf(name, sizeof(name), stepno < o->n_profile_steps ? "calb2 profile step %lld" : (stepno == o->n_profile_steps ? "calb2 warm-up step" : "calb2 step %lld"),
                   stepno < o->n_profile_steps ? stepno : stepno - o->n_profile_steps - 1);
          nvtxRangePushA(name);
        }
        const int r = enqueue_step(pl, k, sum, freeze, o->fuse_tail_update != 0, pl->hist.p, time_heavy ? evs[2 * i] : nullptr,
                                   time_heavy ? evs[2 * i + 1] : nullptr, &launches);
        if (ranged) nvtxRangePop();
        if (r) {
          nvtxRangePop();
          return r;
        }
        heavy_launches++;
      }
    }
    done += gexec ? chunk : n;
    CU(cudaMemcpyAsync(pl->h_state, pl->state.p, sizeof(FitState), cudaMemcpyDeviceToHost, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    if (pl->h_state->error) {
      nvtxRangePop();
      cudaEventDestroy(ev_begin);
      cudaEventDestroy(ev_end);
      for (auto& e : evs) cudaEventDestroy(e);
      if (gexec) cudaGraphExecDestroy(gexec);
      if (graph) cudaGraphDestroy(graph);
      return fail(CALB2_ERR_TIMEOUT, "peer exchange timed out after %.1f s waiting for rank %d at step %d (rank %d of %d): the "
                  "ranks must call calb2_fit with identical maxsteps / tol / n_profile_steps / steps_per_sync, and a rank that "
                  "failed must not leave the others waiting", pl->xtimeout_ns * 1e-9, pl->h_state->error_rank, pl->h_state->step,
                  pl->rank, pl->nranks);
    }
    const bool stopped = pl->h_state->step > pl->h_state->stop_after;
    if (time_heavy) {
        for (int i = 0; i < n; ++i) {  // kernels past the stop are no-ops with ~zero duration
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, evs[2 * i], evs[2 * i + 1]));
        heavy_ms += ms;
      }
    }
    if (stopped) break;
  }
  CU(cudaEventRecord(ev_end, pl->stream));
  CU(cudaStreamSynchronize(pl->stream));
  nvtxRangePop();
  float loop_ms = 0.f;
  CU(cudaEventElapsedTime(&loop_ms, ev_begin, ev_end));
  const FitState hs = *pl->h_state;
  pl->cur_buf = hs.step & 1;

  // calibration.py:702-710 / 722-732: pick the optimum
  if (o->use_min) {
    if (hs.any_snap) {
      CU(cudaMemcpyAsync(pl->g_r[pl->cur_buf].p, pl->gsnap_r.p, ng * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
      CU(cudaMemcpyAsync(pl->g_i[pl->cur_buf].p, pl->gsnap_i.p, ng * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
      if (!freeze) {
        CU(cudaMemcpyAsync(pl->c_r.p, pl->csnap_r.p, pl->ncoef * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
        CU(cudaMemcpyAsync(pl->c_i.p, pl->csnap_i.p, pl->ncoef * sizeof(float), cudaMemcpyDeviceToDevice, pl->stream));
      }
    }
  }
  if (hs.nrec > 0 && loss_history)
    CU(cudaMemcpyAsync(loss_history, pl->hist.p, hs.nrec * sizeof(float), cudaMemcpyDeviceToHost, pl->stream));
  CU(cudaStreamSynchronize(pl->stream));
  if (gexec) cudaGraphExecDestroy(gexec);
  if (graph) cudaGraphDestroy(graph);
  for (auto& e : evs) cudaEventDestroy(e);
  cudaEventDestroy(ev_begin);
  cudaEventDestroy(ev_end);

#ifdef CALB2_PROFILE
  if (!pl->stage_ev.empty()) {
    const char* names[NSTAGE - 1] = {"heavy", "partials", "gains-reduce", "nccl", "finalize", "gains-update", "coeffs"};
    double sum_ms[NSTAGE - 1] = {0};
    const size_t nsteps_done = pl->stage_cursor / NSTAGE;
    for (size_t st = 0; st < nsteps_done; ++st)
      for (int n = 0; n + 1 < NSTAGE; ++n) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, pl->stage_ev[st * NSTAGE + n], pl->stage_ev[st * NSTAGE + n + 1]);
        sum_ms[n] += ms;
      }
    fprintf(stderr, "[calb2 stage us/step, rank %d]", pl->rank);
    for (int n = 0; n + 1 < NSTAGE; ++n) fprintf(stderr, " %s %.1f", names[n], 1e3 * sum_ms[n] / (double)nsteps_done);
    fprintf(stderr, "\n");
    for (auto& e : pl->stage_ev) cudaEventDestroy(e);
    pl->stage_ev.clear();
  }
  if (getenv("CALB2_DBG") && (atoi(getenv("CALB2_DBG")) & 128)) {
    unsigned long long h[16];
    cudaMemcpy(h, pl->dbg_out.p, sizeof(h), cudaMemcpyDeviceToHost);
    cudaMemset(pl->dbg_out.p, 0, sizeof(h));
    const double nt = (double)h[15];
    // BAR.SYNC is defer-blocking: the wait of the F barrier lands in "Q sums", that of the Q barrier in "load q", that of
    // the B barrier in "issue"
    fprintf(stderr, "[calb2 cycles per tile, thread 0] tile wait %.0f | F %.0f | Q sums %.0f  math+stores %.0f  prefetch %.0f | load q %.0f  B %.0f | issue %.0f",
            h[0] / nt, h[1] / nt, h[2] / nt, h[3] / nt, h[4] / nt, h[5] / nt, h[6] / nt, h[7] / nt);
    fprintf(stderr, " (tiles %.0f)\n", nt);
  }
#endif
  res->nsteps_recorded = hs.nrec;
  res->nsteps_total = hs.step;
  res->final_loss = o->use_min ? hs.min_loss : hs.last_loss;
  res->loop_ms = loop_ms;
  res->heavy_ms = (float)heavy_ms;
  res->heavy_launches = time_heavy ? (int64_t)hs.step : heavy_launches;
  res->kernel_launches = launches;
  if (!std::isfinite(hs.last_loss)) g_err = "loss is not finite";
  return 0;
}

int calb2_comm_unique_id(void* id_out, const char* nccl_lib) {
  if (!id_out) return fail(CALB2_ERR_ARG, "null argument");
  if (int r = load_nccl(nccl_lib)) return r;
  int rc = g_nccl.GetUniqueId(id_out);
  if (rc != 0) return fail(CALB2_ERR_NCCL, "ncclGetUniqueId failed (%d)", rc);
  return 0;
}

int calb2_comm_init(calb2_plan* pl, const void* id, int32_t rank, int32_t nranks, const char* nccl_lib) {
  if (!pl || !id) return fail(CALB2_ERR_ARG, "null argument");
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail(CALB2_ERR_ARG, "bad rank/nranks");
  if (nranks == 1) return 0;
  if (pl->gen) return fail(CALB2_ERR_UNSUPPORTED, "the generic (float64 / oversized-group) path runs on one GPU");
  if (int r = load_nccl(nccl_lib)) return r;
  CU(cudaSetDevice(pl->device));
  IdBlob blob;
  memcpy(blob.bytes, id, sizeof(blob.bytes));
  int rc = g_nccl.CommInitRank(&pl->comm, nranks, blob, rank);
  if (rc != 0) return fail(CALB2_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "?");
  pl->rank = rank;
  pl->nranks = nranks;
  // the gain-gradient tables are all-reduced in one call: make them one allocation
  pl->ggrad_r.release();
  pl->ggrad_i.release();
  const size_t ng = (size_t)pl->nants * pl->nfp;
  CU(pl->ggrad_r.alloc(2 * ng));
  CU(cudaMemsetAsync(pl->ggrad_r.p, 0, 2 * ng * sizeof(float), pl->stream));
  pl->ggrad_i.p = pl->ggrad_r.p + ng;  // view; never released separately
  pl->ggrad_i.n = 0;
  return 0;
}

static size_t xbuf_bytes(const calb2_plan* pl) {
  return (size_t)XBUF_FLAG_BYTES + XBUF_SCAL_BYTES + (size_t)2 * 2 * pl->nants * pl->nfp * sizeof(float);
}

int calb2_comm_peer_export(calb2_plan* pl, void* ipc_handle_out) {
  if (!pl || !ipc_handle_out) return fail(CALB2_ERR_ARG, "null argument");
  if (pl->gen) return fail(CALB2_ERR_UNSUPPORTED, "the generic (float64 / oversized-group) path runs on one GPU");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is expected to be 64 bytes");
  CU(cudaSetDevice(pl->device));
  if (!pl->xbuf) {
    CU(cudaMalloc(&pl->xbuf, xbuf_bytes(pl)));
    CU(cudaMemset(pl->xbuf, 0, xbuf_bytes(pl)));
    CU(cudaDeviceSynchronize());
  }
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, pl->xbuf));
  memcpy(ipc_handle_out, &h, sizeof(h));
  return 0;
}

int calb2_comm_peer_import(calb2_plan* pl, const void* ipc_handles, int32_t rank, int32_t nranks) {
  if (!pl || !ipc_handles) return fail(CALB2_ERR_ARG, "null argument");
  if (nranks < 1 || nranks > CALB2_MAX_RANKS || rank < 0 || rank >= nranks) return fail(CALB2_ERR_ARG, "bad rank/nranks");
  if (!pl->xbuf) return fail(CALB2_ERR_STATE, "calb2_comm_peer_export first");
  CU(cudaSetDevice(pl->device));
  pl->rank = rank;
  pl->nranks = nranks;
  pl->peers.n = nranks;
  const size_t ngrad_off = (size_t)XBUF_FLAG_BYTES + XBUF_SCAL_BYTES;
  for (int r = 0; r < nranks; ++r) {
    unsigned char* base = pl->xbuf;
    if (r != rank) {
      cudaIpcMemHandle_t h;
      memcpy(&h, (const char*)ipc_handles + (size_t)r * sizeof(h), sizeof(h));
      void* ptr = nullptr;
      CU(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
      pl->xpeer[r] = ptr;
      base = (unsigned char*)ptr;
    }
    pl->peers.flag[r] = reinterpret_cast<const unsigned int*>(base);
    pl->peers.scal[r] = reinterpret_cast<const double*>(base + XBUF_FLAG_BYTES);
    pl->peers.grad[r] = reinterpret_cast<const float*>(base + ngrad_off);
  }
  pl->xseq = 0;
  pl->peers_open = true;
  if (const char* t = getenv("CALB2_PEER_TIMEOUT_MS")) pl->xtimeout_ns = (unsigned long long)std::max(1.0, atof(t)) * 1000000ull;
  if (!pl->tail_counter.p) {
    CU(pl->tail_counter.alloc(1));
    CU(cudaMemsetAsync(pl->tail_counter.p, 0, sizeof(unsigned int), pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
  }
  return 0;
}

int calb2_comm_peer_close(calb2_plan* pl) {
  if (!pl) return fail(CALB2_ERR_ARG, "null argument");
  if (!pl->peers_open) return 0;
  CU(cudaSetDevice(pl->device));
  // one last round: nobody unmaps (and the owner does not free) a buffer while a peer may still be reading it in its
  // own last step; a peer that is gone only costs the time-out
  int rc = 0;
  if (pl->peers.n > 1 && pl->state.p) {
    FitState s{};
    CU(cudaMemcpyAsync(pl->state.p, &s, sizeof(s), cudaMemcpyHostToDevice, pl->stream));
    rc = peer_barrier(pl);
    if (!rc) {
      CU(cudaMemcpyAsync(pl->h_state, pl->state.p, sizeof(FitState), cudaMemcpyDeviceToHost, pl->stream));
      CU(cudaStreamSynchronize(pl->stream));
      if (pl->h_state->error)
        rc = fail(CALB2_ERR_TIMEOUT, "peer exchange close: rank %d did not arrive within %.1f s", pl->h_state->error_rank,
                  pl->xtimeout_ns * 1e-9);
    }
  }
  for (int r = 0; r < CALB2_MAX_RANKS; ++r)
    if (pl->xpeer[r]) {
      cudaIpcCloseMemHandle(pl->xpeer[r]);
      pl->xpeer[r] = nullptr;
    }
  pl->peers_open = false;
  pl->peers.n = 0;
  pl->nranks = 1;
  return rc;
}

}  // extern "C"
