// Tensor-core shape of the shared-basis pass (sm_100a: tcgen05 + TMEM), for basis classes of <= 208 vectors, single-baseline
// slots, both regularisations ('sum': two launches, at the end of this comment).  Same work item as calfit_shared.cuh -- 64 groups of ONE class through a range of 32-channel tiles --
// with both contractions on the 5th-generation tensor cores as TF32 MMAs with a 3-term split (hi.hi + hi.lo + lo.hi; plain TF32
// would break the 1e-5 loss tolerance):
//
//   phase F   V[128 x 32]  = C[128 x kp] . A[kp x 32]        A-operand = coefficients, hi part resident in TMEM (written once per
//                                                             CTA), lo part a K-major SWIZZLE_128B operand in shared memory;
//                                                             B-operand = the tile, MN-major (SWIZZLE_128B_BASE32B: the only
//                                                             MN-major layout tf32 has), hi and lo copies
//   phase Q   the streaming kernel's arithmetic (calibration.py:1593-1609) by 8 warps on the CUDA cores: a thread owns TMEM lane
//             m = 2 group + part, reads its row of V (tcgen05.ld), pairs re/im with its neighbour lane by shuffles, and writes
//             dL/dv back to TMEM as the next A-operand (hi over V, lo next to it)
//   phase B   dC[128 x kp] += Q[128 x 32] . A[kp x 32]^T      A-operand = dL/dv hi / lo in TMEM, B-operand = the tile, K-major
//                                                             SWIZZLE_128B, hi and lo copies; the accumulator stays in TMEM for
//                                                             all tiles of the CTA
//
// The tensor core accumulates in float32 with truncation (measured: tools/tc_probe.cu, profiles/round2_tcgen05_probe.log: 26
// chained accumulations cost 6e-7 relative).  Where TMEM has the room, V goes to TWO partial accumulators (half of the k-steps
// each) plus one for the two small split terms, and the three are added in registers with round-to-nearest.
//
// TMEM (512 columns) and shared memory (227 KB) set three modes by the class's padded vector count kpt (tc_layout):
//   kpt <= 128   C hi | dC | 2 V buffers x (2 partials + small) | 2 dL/dv-lo     2 + 2 tile stages     F(j+1) overlaps Q(j)
//   kpt <= 160   C hi | dC | 2 V buffers x (1 partial  + small) | 2 dL/dv-lo     2 + 1 tile stages     F(j+1) overlaps Q(j)
//   kpt <= 208   C hi | dC | 1 V buffer  x (1 partial  + small) | 1 dL/dv-lo     1 + 1 tile stages     F, Q, B in turn
// (the MN-major pair of a tile -- phase F's operand -- and its K-major pair -- phase B's -- are separate rings: the first is
// free again once phase F retired, the second only after phase B; one copy cannot serve both, the K-major descriptor does not
// take the 32-byte-base swizzle: tools/tc_probe.cu T8).
//
// One warp issues the MMAs, one the bulk copies (an elected lane each), eight warps do phase Q; mbarriers connect them:
//   full_f[s] / full_b[s]  tile pair landed (TMA) -> MMA warp      free[s]  phase B retired (tcgen05.commit) -> K-major refill;  done -> epilogue
//   v[b]     phase F retired (commit)     -> phase-Q warps, MN-major refill      q[b]  phase Q done (8 warp arrivals) -> MMA warp issues B
// The tensor pipe executes in issue order, so B(j) reads dL/dv(j) out of its V buffer before the next phase F into that buffer
// overwrites it.  Descriptor encodings follow cute/arch/mma_sm100_desc.hpp; every operand form used here is checked by
// tools/tc_probe.cu.
//
// model_regularization = 'sum' (calibration.py:619-661, 1654): the regulariser adds two backward rows per group, P w and Q w
// (P + iQ = g_i conj(g_j)), which depend on the gains and weights only.  shared_tc_kernel<1> is the plain pass plus y = w v (read
// by the gain-gradient kernel) and the two model sums, with dcpart rows of four floats; shared_tc_kernel<2> runs WITHOUT phase F:
// phase Q writes the two rows as its dL/dv, phase B contracts them into floats 2-3 of the rows.  With no phase F in between,
// phase Q of a tile waits for phase B of the previous tile in its V buffer on a barrier of its own (b[2]).
#pragma once
#include "calfit_shared.cuh"

namespace calb2 {

struct TcCfg {
  static constexpr int MS = 64, FT = 32;
  static constexpr int KPMAX = 208;                          // rows per class tile (ncomp rounded up to 16)
  static constexpr int NEPI = 256;                           // 8 phase-Q warps: TMEM lane quadrant = warp & 3, column half = warp >> 2
  static constexpr int NTHR = NEPI + 64;                     // + the MMA warp (8) and the TMA warp (9)
  static constexpr int NPART_MAX = 2;
};

// TMEM columns and shared-memory offsets of a CTA whose class has kpt (multiple of 16) padded vectors
struct TcLayout {
  int nbuf, npart, n_mn, n_k;
  int col_dc, col_v, v_stride, v_small, col_qlo;
  uint32_t sub_bytes, half_bytes;   // one of the four copies of a tile; the MN-major pair / the K-major pair
  uint32_t off_k, off_clo, off_cs, off_ant, off_red, off_bar, off_tmem, total;
};
__host__ __device__ inline TcLayout tc_layout(int kpt) {
  TcLayout L;
  L.nbuf = kpt <= 160 ? 2 : 1;
  L.npart = kpt <= 128 ? 2 : 1;
  L.n_mn = L.nbuf;
  L.n_k = kpt <= 128 ? 2 : 1;
  L.col_dc = kpt;
  L.col_v = 2 * kpt;
  L.v_small = 32 * L.npart;
  L.v_stride = L.v_small + 32;
  L.col_qlo = L.col_v + L.nbuf * L.v_stride;      // + 32 nbuf <= 512 in every mode
  L.sub_bytes = (uint32_t)kpt * 128u;
  L.half_bytes = 2u * L.sub_bytes;
  L.off_k = (uint32_t)L.n_mn * L.half_bytes;
  L.off_clo = L.off_k + (uint32_t)L.n_k * L.half_bytes;
  L.off_cs = L.off_clo + (uint32_t)((kpt + 31) / 32) * 128u * 128u;   // C lo: [chunks of 32 vectors][128 rows][128 B]
  L.off_ant = L.off_cs + TcCfg::MS * 16;
  L.off_red = L.off_ant + TcCfg::MS * 8;
  L.off_bar = L.off_red + 8 * 16;                  // full_f[2], full_b[2], free[2], v[2], q[2], done
  L.off_tmem = L.off_bar + 14 * 8;
  L.total = L.off_tmem + 16;
  return L;
}

// ---- PTX helpers -------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout_type) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(lbo_bytes >> 4) << 16;
  d |= (uint64_t)(sbo_bytes >> 4) << 32;
  d |= (uint64_t)1 << 46;            // descriptor version 1 (Blackwell)
  d |= (uint64_t)layout_type << 61;  // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// K-major SWIZZLE_128B operand: rows of 128 bytes, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t umma_desc_k(uint32_t smem_addr) { return umma_desc(smem_addr, 16, 1024, 2); }
// MN-major tf32 operand, 32 columns, 8 k-rows = two 4-row atoms of the 32-byte-base swizzle, 512 bytes apart
__device__ __forceinline__ uint64_t umma_desc_mn32(uint32_t smem_addr) { return umma_desc(smem_addr, 512, 512, 1); }
__device__ __forceinline__ uint32_t umma_idesc_tf32(int M, int N, int b_mn) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}"
               ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accum), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// one lane of a converged warp: with elect.sync the compiler knows a single thread runs the guarded block and issues
// tcgen05.mma / commit / bulk copies directly (behind `if (lane == 0)` it wraps EVERY such instruction in an ELECT retry loop,
// ~90 cycles per MMA for the issuing thread -- measured, profiles/round2_ncu_hera350.md section 7)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld16(uint32_t addr, float* v) {
  uint32_t r[16];
  // load + wait in ONE asm statement: the registers are only defined after tcgen05.wait::ld, and nothing may be scheduled between
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
               "tcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(addr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t addr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(addr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
               "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])),
               "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])),
               "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])),
               "r"(__float_as_uint(v[15]))
               : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t addr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(addr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t addr, float* v) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
               "tcgen05.wait::ld.sync.aligned;"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(addr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// 256-bit global accesses (sm_100: LDG.256 / STG.256): a thread's 8 consecutive channels in ONE instruction -- half the L1
// wavefronts of two 128-bit accesses for the per-baseline rows, whose lanes all sit in different 128-byte lines
struct alignas(32) f32x8 { float v[8]; };
__device__ __forceinline__ f32x8 ldg256(const float* ptr) {
  f32x8 r;
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7])
               : "l"(ptr));
  return r;
}
__device__ __forceinline__ void stg256(float* ptr, float a, float b, float c, float d, float e, float f, float g, float h) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "f"(a), "f"(b), "f"(c), "f"(d), "f"(e), "f"(f),
               "f"(g), "f"(h) : "memory");
}
__device__ __forceinline__ float tf32_trunc(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ void mbar_arrive_plain(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Bounded wait: a protocol error must not hang the GPU.  mbarrier.try_wait suspends the thread in hardware (with a long
// suspend-time hint: no polling traffic, no issue slots taken); every 4th return without completion the waiter looks at the SM clock (clock64: a register read -- %globaltimer
// costs thousands of cycles per read and stretched every long wait, profiles/round2_ncu_hera350.md section 7), and after
// ~5 s records (code, tile, CTA, thread) in a mapped host buffer and traps; calb2 reports the record with the CUDA error.
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity, unsigned int* dbg, unsigned int code, int j, bool lane0_only = false) {
  // called by converged warps; lane0_only: lane 0 polls, the warp re-converges behind it
  if (!lane0_only || (threadIdx.x & 31) == 0) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    unsigned int polls = 0;
    long long t0 = 0;
    for (;;) {
      asm volatile(
          "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(addr), "r"(parity), "r"(0x989680u)  // suspend-time hint: sleep in hardware until the phase completes -- a spinning
                                                    // waiter takes issue slots from the phase-Q warps of its scheduler (measured)
          : "memory");
      if (done) break;
      if ((++polls & 3u) == 0u) {
        const long long now = clock64();
        if (t0 == 0) t0 = now;
        if (now - t0 > 10000000000ll) {
          if (dbg) {
            dbg[1] = (unsigned int)j;
            dbg[2] = blockIdx.x;
            dbg[3] = threadIdx.x;
            __threadfence_system();
            dbg[0] = code;
            __threadfence_system();
          }
          __trap();
        }
      }
    }
  }
  __syncwarp();
}

__device__ __forceinline__ void tc_mark(unsigned int* dbg, int slot, unsigned int value) {
  if (dbg && blockIdx.x == 0) {
    *reinterpret_cast<volatile unsigned int*>(dbg + slot) = value;
    __threadfence_system();
  }
}

struct TcParams {
  const float* At;      // tensor-core copies of the class tiles: per class [tile][4][kpt][32]
  const MTileDesc* tiles;  // a_off = float offset of the class in At, kp = kpt (multiple of 16)
  const ClassSlot* cslots;
  const int* bl_ant0;
  const int* bl_ant1;
  const float* d_r;
  const float* d_i;
  const float* w;
  const float* g_r[2];
  const float* g_i[2];
  const float* c_r;
  const float* c_i;
  float2* z;
  float2* y;            // 'sum' (mode 1): w v per visibility, read by the gain-gradient kernel
  float* dcpart;
  long long dc_plane;
  double* partials;
  const FitState* st;
  int nfp;
  unsigned int* dbg;    // mapped host memory: [0] = code of a wait that timed out, [1] tile, [2] CTA, [3] thread
  long long* prof;      // development aid (CALB2_TC_PROF=<cta>): clock64 stamps of that CTA, [tile][TC_PROF_SLOTS]
  int prof_cta;
  int mode;             // (= the kernel's template argument) 0: plain chi^2 (two backward rows per group, dcpart rows of 2); 1: 'sum', first launch: the same plus y and
                        // the model sums, dcpart rows of 4; 2: 'sum', second launch: no phase F, the backward rows P w / Q w
  int flags;            // development switches (CALB2_TC_FLAGS): 1 lane-0 polling, 2 L2 row prefetch, 4 clock-paced issue, 8 phase B first when ready
};
constexpr int TC_PROF_SLOTS = 24, TC_PROF_TILES = 32;  // 0-4 MMA warp, 6-10 phase-Q warp 0, 12-19 arrival of each phase-Q warp, 20-22 (tile 0) start / prologue / end
__device__ __forceinline__ void tc_stamp(const TcParams& p, int j, int slot) {
  if (p.prof && (int)blockIdx.x == p.prof_cta && j < TC_PROF_TILES && (threadIdx.x & 31) == 0) p.prof[j * TC_PROF_SLOTS + slot] = clock64();
}

template <int MODE>
__global__ void __launch_bounds__(TcCfg::NTHR, 1) shared_tc_kernel(const TcParams p) {
  using C = TcCfg;
  extern __shared__ __align__(1024) unsigned char smem[];
  const FitState* st = p.st;
  if (st->step > st->stop_after) return;  // fit already stopped (uniform across the grid)
  const MTileDesc mt = p.tiles[blockIdx.x];
  const int gsel = st->step & 1;
  const float* __restrict__ g_r = p.g_r[gsel];
  const float* __restrict__ g_i = p.g_i[gsel];
  const int kpt = mt.kp, nslots = mt.nslots;
  const TcLayout L = tc_layout(kpt);
  const bool l0 = p.flags & 1, use_prefetch = p.flags & 2, use_pace = p.flags & 4, b_first = p.flags & 8;
  constexpr int mode = MODE;        // compile-time: the plain pass keeps its instruction stream and registers
  constexpr bool no_f = MODE == 2;  // second launch of 'sum': backward contraction only

  float* Clo = reinterpret_cast<float*>(smem + L.off_clo);
  ClassSlot* s_cs = reinterpret_cast<ClassSlot*>(smem + L.off_cs);
  int2* s_ant = reinterpret_cast<int2*>(smem + L.off_ant);
  float* red = reinterpret_cast<float*>(smem + L.off_red);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L.off_bar);
  uint64_t *bar_full = bars, *bar_fullb = bars + 2, *bar_free = bars + 4, *bar_v = bars + 6, *bar_q = bars + 8, *bar_done = bars + 10,
           *bar_b = bars + 12;  // mode 2 only: phase B of the tile in V buffer b retired -> phase Q may overwrite dL/dv there
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(smem + L.off_tmem);

  const int tid = threadIdx.x, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
  if (tid == 0) tc_stamp(p, 0, 20);
  const int ntiles = mt.j1 - mt.j0;
  const uint32_t sub_bytes = L.sub_bytes, half_bytes = L.half_bytes;
  const float* Abase = p.At + mt.a_off + (size_t)mt.j0 * kpt * 4 * 32;
  const size_t tile_floats = (size_t)kpt * 4 * 32;
#ifdef CALB2_TC_SWAP_ROLES
  const bool mma_warp = warp == 9, tma_warp = warp == 8;
#else
  const bool mma_warp = warp == 8, tma_warp = warp == 9;
#endif
  const int nbuf = L.nbuf, n_mn = L.n_mn, n_k = L.n_k;
  // ring / buffer slot and barrier parity of tile j for a ring of n (1 or 2) entries
  auto slot_of = [](int j, int n) { return n == 2 ? (j & 1) : 0; };
  auto par_of = [](int j, int n) { return (uint32_t)(n == 2 ? (j >> 1) & 1 : j & 1); };

  if (mma_warp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    if (lane == 0) {
      for (int i = 0; i < 8; ++i) mbar_init(&bars[i], 1);  // full_f, full_b, free, v
      // two: a fast warp may arrive for tile j + 1 before a slow one has arrived for tile j.  One arrival per WARP: every arrival
      // wakes the waiters sleeping on the CTA's barriers, and 256 of them per tile kept the MMA / TMA warps polling (measured)
      mbar_init(&bar_q[0], C::NEPI / 32);
      mbar_init(&bar_q[1], C::NEPI / 32);
      mbar_init(bar_done, 1);
      mbar_init(&bar_b[0], 1);
      mbar_init(&bar_b[1], 1);
      mbar_fence_init();
      for (int jj = 0; jj < n_mn && jj < ntiles && !no_f; ++jj) {
        mbar_expect_tx(&bar_full[jj], half_bytes);
        bulk_g2s(smem + jj * half_bytes, Abase + jj * tile_floats, half_bytes, &bar_full[jj]);
      }
      for (int jj = 0; jj < n_k && jj < ntiles; ++jj) {
        mbar_expect_tx(&bar_fullb[jj], half_bytes);
        bulk_g2s(smem + L.off_k + jj * half_bytes, Abase + jj * tile_floats + 2 * kpt * 32, half_bytes, &bar_fullb[jj]);
      }
    }
  }
  if (tid < C::MS) {
    ClassSlot cs = {0, 0, 0, 0};
    int2 ants = make_int2(0, 0);
    if (tid < nslots) {
      cs = p.cslots[mt.cs0 + tid];
      ants = make_int2(p.bl_ant0[cs.bl0] * p.nfp, p.bl_ant1[cs.bl0] * p.nfp);
    }
    s_cs[tid] = cs;
    s_ant[tid] = ants;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *s_tmem;

  // ---- phase-Q thread identity: TMEM lane m = 2 group + part; the warp's quadrant of lanes; its half of the 32 columns
  const int quad = warp & 3, half = (warp >> 2) & 1;
  const int m = quad * 32 + lane, s = m >> 1, part = m & 1;
  const uint32_t tlane = (uint32_t)(quad * 32) << 16;
  const bool q_warp = warp < 8;
  const bool valid = q_warp && s < nslots;

  // ---- coefficients: hi part -> TMEM (the A-operand of phase F for the whole pass), lo part -> shared memory operand
  if (q_warp && !no_f) {
    // pass 1, coalesced: warp w brings rows 16 w .. 16 w + 15 (a group's coefficients are contiguous: lanes run over the vectors)
    // into the C lo buffer, already at their K-major SWIZZLE_128B positions (chunk of 32 vectors, row m, 16-byte pieces XOR-ed
    // with m & 7).  Lane-per-row loads cost 32 L1 wavefronts per instruction and made this prologue 17 000 cycles long.
    // (eight rows at a time: all their loads are issued before the first value is used)
    for (int r0 = 0; r0 < 16; r0 += 8) {
      float c[8][7];  // kpt <= 208: at most 7 chunks of 32 vectors
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int mm = warp * 16 + r0 + r, ss = mm >> 1;
        const bool ok = ss < nslots;
        const float* src = ((mm & 1) ? p.c_i : p.c_r) + s_cs[ss].coef0;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const int k = lane + 32 * q;
          c[r][q] = (ok && k < mt.ncomp) ? src[k] : 0.f;
        }
      }
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        const int mm = warp * 16 + r0 + r;
#pragma unroll
        for (int q = 0; q < 7; ++q) {
          const int k = lane + 32 * q;
          if (k < kpt) Clo[q * 128 * 32 + mm * 32 + ((((lane >> 2) ^ (mm & 7)) << 2) | (lane & 3))] = c[r][q];
        }
      }
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");  // the eight phase-Q warps
    // pass 2: thread (row m, half) splits its half of the row: hi -> TMEM (phase F's A-operand), lo stays in place
    const int k_lo = half * (kpt / 2), k_hi = k_lo + kpt / 2;  // kpt / 2 is a multiple of 8
    for (int k0 = k_lo; k0 < k_hi; k0 += 8) {
      float* row = Clo + (k0 >> 5) * 128 * 32 + m * 32;
      const int c16 = (k0 & 31) >> 2;
      float4* p0 = reinterpret_cast<float4*>(row + ((c16 ^ (m & 7)) << 2));
      float4* p1 = reinterpret_cast<float4*>(row + (((c16 + 1) ^ (m & 7)) << 2));
      const float4 a = *p0, b = *p1;
      const float c[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
      float hi[8], lo[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        hi[i] = tf32_trunc(c[i]);
        lo[i] = c[i] - hi[i];
      }
      *p0 = make_float4(lo[0], lo[1], lo[2], lo[3]);
      *p1 = make_float4(lo[4], lo[5], lo[6], lo[7]);
      tmem_st8(tmem + tlane + k0, hi);
    }
    tmem_st_wait();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // C lo was written by ordinary stores, the MMA reads it through the async proxy
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  if (tid == 0) tc_stamp(p, 0, 21);
  if (mma_warp) {
    // =====================================================================================================================
    // MMA / TMA warp: the whole warp runs the control flow and the waits, one elected lane issues
    // =====================================================================================================================
    const uint32_t idesc_f = umma_idesc_tf32(128, 32, 1);
    const uint32_t idesc_b = umma_idesc_tf32(128, kpt, 0);
    const int nks = kpt / 8;
    const int per = (nks + L.npart - 1) / L.npart;  // k-steps per partial V accumulator
    uint32_t dc_accum = 0;
    const uint32_t mn_addr = smem_u32(smem), k_addr = smem_u32(smem + L.off_k);
    const uint64_t d_clo = umma_desc_k(smem_u32(Clo));
    // phase F of tile jj into its V buffer:  V_p += C_hi . A_hi ;  V_small += C_hi . A_lo + C_lo . A_hi
    // (inside the loops only the 14-bit start-address field of a descriptor moves, in units of 16 bytes)
    // Pacing.  The tensor pipe's queue is a handful of MMAs deep; issue beyond that blocks, and a warp blocked there also holds
    // up the TMEM loads / stores of the phase-Q warps of ITS scheduler (they arrived 1500-2000 cycles after the other six, and
    // the pair moved with the MMA warp when it was put on another scheduler: profiles/round2_ncu_hera350.md section 7).
    // Commit-based pacing costs ~500 cycles per round trip, so the issuing lane paces itself on the SM clock instead: it keeps
    // an estimate of when the pipe will have finished what was issued (measured costs per MMA, tools/tc_probe.cu) and issues the
    // next three MMAs only when less than PACE_SLACK cycles of work are left -- the queue stays about four deep and never blocks.
    constexpr int PACE_SLACK = 45;  // issue the next group of three when about one MMA of work is left
    const int cost_ts32 = 22, cost_ss32 = 41, cost_b = kpt / 2 + 2;
    long long t_pipe = 0;
    auto pace = [&](int cost) {
      if (!use_pace) return;
      long long now = clock64();
      while (t_pipe - now > PACE_SLACK) now = clock64();
      t_pipe = (t_pipe > now ? t_pipe : now) + cost;
    };
    auto issue_f = [&](int jj) {
      const int sg = slot_of(jj, n_mn);
      const uint32_t vcol = tmem + L.col_v + L.v_stride * slot_of(jj, nbuf);
      tc_wait(&bar_full[sg], par_of(jj, n_mn), p.dbg, 1, jj, l0);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t sbase = mn_addr + sg * half_bytes;
        uint64_t b_hi = umma_desc_mn32(sbase), b_lo = umma_desc_mn32(sbase + sub_bytes), a_lo = d_clo;
        uint32_t a_hi = tmem, vp = vcol;
        int in_part = 0;
        for (int ks = 0; ks < nks; ++ks) {
          pace(2 * cost_ts32 + cost_ss32);
          umma_ts(vp, a_hi, b_hi, idesc_f, in_part ? 1u : 0u);
          umma_ts(vcol + L.v_small, a_hi, b_lo, idesc_f, ks ? 1u : 0u);
          umma_ss(vcol + L.v_small, a_lo, b_hi, idesc_f, 1u);
          b_hi += 64;   // 8 rows x 128 B = 1024 B
          b_lo += 64;
          a_hi += 8;
          a_lo += ((ks & 3) == 3) ? (16384u - 96u) / 16u : 2u;  // next 32 B of the row, or the next 32-vector chunk
          if (++in_part == per) {
            in_part = 0;
            vp += 32;
          }
        }
        umma_commit(&bar_v[slot_of(jj, nbuf)]);
      }
      __syncwarp();
    };
    // phase B of tile j once phase Q has written dL/dv: dC += Q_hi . A_hi^T + Q_hi . A_lo^T + Q_lo . A_hi^T
    auto issue_b = [&](int j) {
      const int sk = slot_of(j, n_k), vb = slot_of(j, nbuf);
      tc_wait(&bar_q[vb], par_of(j, nbuf), p.dbg, 2, j, l0);
      tc_wait(&bar_fullb[sk], par_of(j, n_k), p.dbg, 7, j, l0);
      tc_fence_after();
      tc_stamp(p, j, 2);
      if (elect_one()) {
        const uint32_t sbase = k_addr + sk * half_bytes;
        const uint64_t d_b_hi = umma_desc_k(sbase), d_b_lo = umma_desc_k(sbase + sub_bytes);
        const uint32_t qhi = tmem + L.col_v + L.v_stride * vb, qlo = tmem + L.col_qlo + 32 * vb;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t b_hi = d_b_hi + 2 * ks, b_lo = d_b_lo + 2 * ks;  // 32 B along the 128 B row
          pace(3 * cost_b);
          umma_ts(tmem + L.col_dc, qhi + 8 * ks, b_hi, idesc_b, ks ? 1u : dc_accum);
          umma_ts(tmem + L.col_dc, qhi + 8 * ks, b_lo, idesc_b, 1u);
          umma_ts(tmem + L.col_dc, qlo + 8 * ks, b_hi, idesc_b, 1u);
        }
        umma_commit(&bar_free[sk]);
        if (no_f) umma_commit(&bar_b[vb]);  // no phase F in between: phase Q of the next tile in this buffer waits for this instead
        // the accumulator is final after the last phase B.  (Its own barrier: with a one-entry ring the parity of free[] repeats
        // every two tiles, and a phase-Q warp can get here while phase B of the tile before the last is still in flight.)
        if (j == ntiles - 1) umma_commit(bar_done);
      }
      __syncwarp();
      dc_accum = 1u;
      tc_stamp(p, j, 3);
    };
    if (!no_f) issue_f(0);
    for (int j = 0; j < ntiles; ++j) {
      const int vb = slot_of(j, nbuf);
      bool b_issued = false;
      tc_stamp(p, j, 0);
      if (!no_f && nbuf == 2 && j + 1 < ntiles) {
        // Two V buffers: phase F of the next tile runs on the tensor cores while phase Q of this one runs on the CUDA cores.
        // (switch 8: if phase Q is already through, phase B goes first)
        uint32_t ready;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ready) : "r"(smem_u32(&bar_q[vb])), "r"(par_of(j, nbuf)) : "memory");
        if (b_first && __all_sync(0xffffffffu, ready != 0u)) {
          issue_b(j);
          b_issued = true;
        }
        issue_f(j + 1);
      }
      tc_stamp(p, j, 1);
      if (!b_issued) issue_b(j);
      // one V buffer: the next phase F goes behind this phase B in the tensor pipe (its operands landed during phase Q)
      if (!no_f && nbuf == 1 && j + 1 < ntiles) issue_f(j + 1);
      tc_stamp(p, j, 4);
    }
  } else if (tma_warp) {
    // =====================================================================================================================
    // TMA warp: refills the two tile rings as their stages retire -- off the MMA warp's path, whose issue is synchronous with the
    // tensor pipe (a bulk copy or a wait there is a bubble in the pipe: measured, profiles/round2_ncu_hera350.md section 7)
    // =====================================================================================================================
    for (int j = 0; j < ntiles; ++j) {
      if (!no_f && j + n_mn < ntiles) {  // phase F of tile j has retired: its MN-major pair takes tile j + n_mn
        tc_wait(&bar_v[slot_of(j, nbuf)], par_of(j, nbuf), p.dbg, 6, j, l0);
        if (elect_one()) {
          const int sg = slot_of(j, n_mn);
          mbar_expect_tx(&bar_full[sg], half_bytes);
          bulk_g2s(smem + sg * half_bytes, Abase + (size_t)(j + n_mn) * tile_floats, half_bytes, &bar_full[sg]);
        }
        __syncwarp();
      }
      if (j + n_k < ntiles) {   // phase B of tile j has retired: its K-major pair takes tile j + n_k
        const int sk = slot_of(j, n_k);
        tc_wait(&bar_free[sk], par_of(j, n_k), p.dbg, 3, j, l0);
        if (elect_one()) {
          mbar_expect_tx(&bar_fullb[sk], half_bytes);
          bulk_g2s(smem + L.off_k + sk * half_bytes, Abase + (size_t)(j + n_k) * tile_floats + 2 * kpt * 32, half_bytes, &bar_fullb[sk]);
        }
        __syncwarp();
      }
    }
  } else {
    // =====================================================================================================================
    // phase-Q warps
    // =====================================================================================================================
    const int npart = L.npart;
    const int cbase = 16 * half;        // this warp's columns of the tile
    const int mycol = cbase + 8 * part; // the 8 channels this thread does the arithmetic for
    const ClassSlot cs = s_cs[s];
    const int2 an = s_ant[s];
    float loss_acc = 0.f, sr_acc = 0.f, si_acc = 0.f;
    // this thread's inputs are loaded ONE TILE AHEAD (registers): their DRAM / L2 latency hides behind the previous tile's work.
    // Two register sets, the tile loop unrolled by two (no copies).
    f32x8 bufA[7], bufB[7];
    auto load_inputs = [&](int jt, f32x8 (&dst)[7]) {
      const int f0 = (mt.j0 + jt) * C::FT + mycol;
      const int o = cs.bl0 * p.nfp + f0, o0 = an.x + f0, o1 = an.y + f0;
      if (!no_f) {
        dst[0] = ldg256(p.d_r + o);
        dst[1] = ldg256(p.d_i + o);
      }
      dst[2] = ldg256(p.w + o);
      dst[3] = ldg256(g_r + o0);
      dst[4] = ldg256(g_i + o0);
      dst[5] = ldg256(g_r + o1);
      dst[6] = ldg256(g_i + o1);
    };
    // The per-baseline rows arrive 128 bytes per tile and row: every PF tiles each thread asks L2 for one line (group tid / 4, tile
    // tid % 4 of the next PF tiles) of each of the three arrays, 2 PF tiles ahead -- neighbouring lanes walk along a row, so DRAM
    // sees 512-byte pieces, and the loads below find L2.  (prefetch.global.L2 is an ordinary LSU instruction; the bulk-copy
    // prefetch went through the TMA queue and delayed the tile copies behind 192 small requests.)
    constexpr int PF = 4;
    auto prefetch_rows = [&](int jt) {
      const int ps = tid >> 2, pt = jt + (tid & 3);
      if (use_prefetch && ps < nslots && pt < ntiles) {
        const size_t o = (size_t)s_cs[ps].bl0 * p.nfp + (size_t)(mt.j0 + pt) * C::FT;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.d_r + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.d_i + o));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.w + o));
      }
    };
    prefetch_rows(0);
    prefetch_rows(PF);
    auto do_tile = [&](int j, f32x8 (&in)[7], f32x8 (&nx)[7]) {
      const int f0 = (mt.j0 + j) * C::FT + mycol;
      const int vb = slot_of(j, nbuf);
      if ((j & (PF - 1)) == 0) prefetch_rows(j + 2 * PF);
      if (valid && j + 1 < ntiles) load_inputs(j + 1, nx);
      const uint32_t vcol = tmem + tlane + L.col_v + L.v_stride * vb + cbase;
      if (tid == 0) tc_stamp(p, j, 6);
      float vr[8], vi[8];
      if (!no_f) {
        tc_wait(&bar_v[vb], par_of(j, nbuf), p.dbg, 4, j, l0);
        tc_fence_after();
        if (tid == 0) tc_stamp(p, j, 7);
        // V row of this thread, 16 columns: partial accumulators + the small terms, added with round-to-nearest
        float v[16], t[16];
        tmem_ld16(vcol, v);
        if (npart > 1) {
          tmem_ld16(vcol + 32, t);
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] += t[i];
        }
        tmem_ld16(vcol + L.v_small, t);
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] += t[i];
        if (tid == 0) tc_stamp(p, j, 8);
        // pair re / im: lane 2g holds v_r of group g, lane 2g + 1 its v_i; each takes 8 of the 16 channels
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float send = part ? v[i] : v[8 + i];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
          vr[i] = part ? recv : v[i];
          vi[i] = part ? v[8 + i] : recv;
        }
      } else if (j >= nbuf) {
        // second launch of 'sum': no phase F orders this tile behind phase B of the previous tile in this buffer
        tc_wait(&bar_b[vb], par_of(j - nbuf, nbuf), p.dbg, 9, j, l0);
        tc_fence_after();
      }
      float qr[8], qi[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) qr[i] = qi[i] = 0.f;
      if (valid) {
        const float *dr = in[0].v, *di = in[1].v, *ww = in[2].v, *gr0 = in[3].v, *gi0 = in[4].v, *gr1 = in[5].v, *gi1 = in[6].v;
        if (no_f) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // the regulariser's backward rows (calibration.py:1654): P w and Q w
            qr[i] = (gr0[i] * gr1[i] + gi0[i] * gi1[i]) * ww[i];
            qi[i] = (gr0[i] * gi1[i] - gi0[i] * gr1[i]) * ww[i];
          }
        } else {
          float2 zz[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {  // calibration.py:1593-1609, as in the other kernels
            const float P = gr0[i] * gr1[i] + gi0[i] * gi1[i];
            const float Q = gr0[i] * gi1[i] - gi0[i] * gr1[i];
            const float mr = P * vr[i] + Q * vi[i];
            const float mi = P * vi[i] - Q * vr[i];
            const float rr = dr[i] - mr, ri = di[i] - mi;
            loss_acc += (rr * rr + ri * ri) * ww[i];
            const float er = -2.f * ww[i] * rr, ei = -2.f * ww[i] * ri;
            zz[i] = make_float2(er * vr[i] + ei * vi[i], er * vi[i] - ei * vr[i]);
            qr[i] = P * er - Q * ei;
            qi[i] = Q * er + P * ei;
            if (mode == 1) {
              sr_acc += ww[i] * mr;
              si_acc += ww[i] * mi;
            }
          }
          const size_t zo = (size_t)(cs.bl0 * p.nfp + f0);
          float* zdst = reinterpret_cast<float*>(p.z + zo);
          stg256(zdst, zz[0].x, zz[0].y, zz[1].x, zz[1].y, zz[2].x, zz[2].y, zz[3].x, zz[3].y);
          stg256(zdst + 8, zz[4].x, zz[4].y, zz[5].x, zz[5].y, zz[6].x, zz[6].y, zz[7].x, zz[7].y);
          if (mode == 1) {
            float* ydst = reinterpret_cast<float*>(p.y + zo);
            stg256(ydst, ww[0] * vr[0], ww[0] * vi[0], ww[1] * vr[1], ww[1] * vi[1], ww[2] * vr[2], ww[2] * vi[2], ww[3] * vr[3],
                   ww[3] * vi[3]);
            stg256(ydst + 8, ww[4] * vr[4], ww[4] * vi[4], ww[5] * vr[5], ww[5] * vi[5], ww[6] * vr[6], ww[6] * vi[6], ww[7] * vr[7],
                   ww[7] * vi[7]);
          }
        }
      }
      if (tid == 0) tc_stamp(p, j, 9);
      // back to rows: lane 2g needs q_r of all 16 columns, lane 2g + 1 q_i
      float row[16], hi[16], lo[16];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float send = part ? qr[i] : qi[i];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        row[i] = part ? recv : qr[i];
        row[8 + i] = part ? qi[i] : recv;
      }
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        hi[i] = tf32_trunc(row[i]);
        lo[i] = row[i] - hi[i];
      }
      tmem_st16(vcol, hi);  // over partial 0 of this tile's V buffer: these lanes / columns are read by this warp only
      tmem_st16(tmem + tlane + L.col_qlo + 32 * vb + cbase, lo);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_plain(&bar_q[vb]);
      if (tid == 0) tc_stamp(p, j, 10);
      tc_stamp(p, j, 12 + warp);
    };
    if (valid) load_inputs(0, bufA);
    for (int j = 0; j < ntiles; j += 2) {
      do_tile(j, bufA, bufB);
      if (j + 1 < ntiles) do_tile(j + 1, bufB, bufA);
    }
    // ---- backward sums: wait for the last phase B, then row m of dC -> dcpart
    {
      const int last = ntiles - 1;
      tc_wait(bar_done, 0, p.dbg, 5, last, l0);
      tc_fence_after();
      // rows of dcpart are (re, im) pairs: lane 2g (re) takes the even vectors of both parts, lane 2g + 1 the odd ones
      // (plain chi^2: rows of 2 floats; 'sum': rows of 4, the first launch fills floats 0-1, the second 2-3)
      const int rs = mode ? 2 : 1;  // float2 per row
      float2* dst = reinterpret_cast<float2*>(p.dcpart + (size_t)mt.seg * p.dc_plane) + (size_t)cs.row0 * rs + (mode == 2 ? 1 : 0);
      const int k_lo = half * (kpt / 2), k_hi = k_lo + kpt / 2;
      for (int k0 = k_lo; k0 < k_hi; k0 += 8) {
        float d8[8];
        tmem_ld8(tmem + tlane + L.col_dc + k0, d8);  // .sync.aligned: the whole warp, also the lanes of missing groups
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
          // this lane keeps vector i + part: it sends the other one and receives the partner's value of its own
          const float keep = part ? d8[i + 1] : d8[i], send = part ? d8[i] : d8[i + 1];
          const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
          const int k = k0 + i + part;
          if (valid && k < mt.ncomp) dst[(size_t)k * rs] = part ? make_float2(recv, keep) : make_float2(keep, recv);
        }
      }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
      sr_acc += __shfl_xor_sync(0xffffffffu, sr_acc, off);
      si_acc += __shfl_xor_sync(0xffffffffu, si_acc, off);
    }
    if (lane == 0) {
      red[warp] = loss_acc;
      red[8 + warp] = sr_acc;
      red[16 + warp] = si_acc;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (tid == 0 && mode != 2) {  // (the second launch of 'sum' has no sums of its own)
    tc_stamp(p, 0, 22);
    double a = 0.0, b = 0.0, c = 0.0;
    for (int w8 = 0; w8 < 8; ++w8) {
      a += (double)red[w8];
      b += (double)red[8 + w8];
      c += (double)red[16 + w8];
    }
    double* dst = p.partials + (size_t)blockIdx.x * 4;
    dst[0] = a;
    dst[1] = b;
    dst[2] = c;
  }
  if (mma_warp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

}  // namespace calb2
