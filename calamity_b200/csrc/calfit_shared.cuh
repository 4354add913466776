// Shared-basis path of the fused fit kernel (sm_100a).
//
// In the reference's per-baseline DPSS layout every baseline is its own fitting group, but the basis only depends on
// the baseline's integer-nanosecond delay (modeling.py:293; operator cache modeling.py:352, 371): at HERA-350 the
// 61 075 groups share 120 distinct bases (16 036 distinct rows, 66 MB -- L2-resident), ~509 groups each.  The
// streaming kernel of calfit_kernels.cuh reads one private copy per group (25 GB per iteration).  Here every distinct
// basis ("class") is stored ONCE, and a CTA takes MS groups of one class through a range of channels, so that per staged
// [kp rows x 32 channels] tile the contraction is a small GEMM on the CUDA cores with register-tiled operands:
//
//   phase F   V[2 MS x 32]   = C[2 MS x kp] . A[kp x 32]          (calibration.py:1587-1590; rows = (part, group))
//   phase Q   per (group, channel): gains, model, weighted residual, chi^2, z, dL/dv  (calibration.py:1593-1609)
//   phase B   dC[NQ MS x kp] += Q[NQ MS x 32] . A[kp x 32]^T       (coefficient half of the tape gradient, 664-666)
//
// The dC accumulators stay in registers across all tiles of the CTA (one thread owns an 8 x TK block of
// (row, vector) pairs for the whole pass: no cross-thread reduction, deterministic), the tile comes in by one bulk
// asynchronous copy (TMA engine) per buffer with mbarrier double buffering, exactly like the streaming kernel.
// The tile is stored with the 128-byte XOR swizzle (16-byte chunk index ^ (row & 7)), which makes both the row-wise
// reads of phase F and the column-of-rows reads of phase B bank-conflict free (and is the layout a tcgen05 / TMA
// SWIZZLE_128B descriptor expects, should the contraction move to the tensor cores).
//
// DRAM traffic per iteration drops from 4 N_A_nz + 12 N_D to ~12 N_D + 8 N_D (z) + the parameters; the kernel is
// bound by the FP32 FMA pipe (8 N_A_nz flops), not by HBM.
#pragma once
#include "calfit_kernels.cuh"

namespace calb2 {

struct MTileDesc {
  long long a_off;  // float offset of the class's tile 0: [ntiles][kp][32], swizzled
  int kp;           // staged rows: ncomp rounded up to 8 (padding rows are zero)
  int ncomp;
  int nslots;       // groups of the class taken by this CTA, <= MS
  int cs0;          // first entry of the CTA in the class-slot tables
  int j0, j1;       // channel tiles [j0, j1) of this CTA (a class tile is cut into segments to get enough CTAs)
  int seg;          // segment index = plane of dcpart this CTA's backward sums go to
  int pad;
};

struct ClassSlot {
  int coef0;  // first coefficient of the group
  int row0;   // row of the group's first backward sum in dcpart
  int bl0;    // first baseline of the slot
  int nb;     // baselines of the slot (redundant baselines share the model visibility)
};

struct SharedParams {
  const float* A;
  const MTileDesc* tiles;
  const ClassSlot* cslots;
  const int* cs_slot;   // global slot index (row of vout)
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
  float2* y;
  float* dcpart;        // [segment plane][rows][NQ]
  long long dc_plane;   // floats per plane
  float2* vout;
  double* partials;     // [gridDim.x][4], already offset past the other launches of the pass
  const FitState* st;
  int nfp;
  int store_v;          // forward only, model visibilities -> vout
  int init_mode;        // backward only, dL/dv := data * (w != 0)   (coefficient initialisation, calibration.py:875-902)
};

// Two shapes of the same kernel:
//   NTHR = 256, KPM = 160: 8 warps, 32 groups per CTA (16 with the 'sum' regulariser), <= 100 KB of shared memory and 128
//                          registers -> TWO CTAs per SM, whose phases interleave (one CTA's phase Q -- global-load latency --
//                          and barriers hide behind the other's contractions).  Classes of up to 160 vectors: 85 % of the
//                          work at HERA-350.
//   NTHR = 512, KPM = 208: 16 warps, 64 groups per CTA (32 with 'sum'), 200 KB -> one CTA per SM, for the large classes
//                          whose coefficient block does not fit twice.
template <int NTHR_, int NQ_, int KPM_>
struct SharedCfg {
  static constexpr int NTHR = NTHR_;
  static constexpr int NWARP = NTHR / 32;
  static constexpr int NQ = NQ_;            // backward sums per group: 2, or 4 with the 'sum' regulariser
  static constexpr int KPM = KPM_;          // largest class (rows) the shape takes
  static constexpr int FT = 32;             // channels per tile = one 128-byte swizzle row
  static constexpr int MB = 8 * NWARP;      // backward rows (row = q * MS + group): a warp owns 8
  static constexpr int MS = MB / NQ;        // groups per CTA
  static constexpr int MF = 2 * MS;         // forward rows (part-major: row = part * MS + group)
  static constexpr int HALF = NTHR / 2;     // the two halves of the CTA split the rows k of the tile in phase F
  static constexpr int MPT = MF * 8 / HALF; // forward rows per thread (x 4 channels)
  static constexpr int TKMAX = (KPM + 31) / 32;  // vectors per thread in phase B: k = lane + 32 t
  static constexpr int KROWS = 32 * TKMAX;  // rows of a tile buffer
  static constexpr int CT_PITCH = MF + 4;   // coefficients, [k][row]: + 4 keeps the staging writes at 4-way conflicts
  static constexpr int QN = MS / NWARP;     // groups per warp in phase Q
  static constexpr int OFF_A = 0;
  static constexpr int OFF_CT = OFF_A + 2 * KROWS * FT * 4;
  static constexpr int OFF_V = OFF_CT + KPM * CT_PITCH * 4;   // V: two k-halves; dL/dv (phase Q -> B) is written IN PLACE over it
  static constexpr int OFF_CS = OFF_V + 2 * MF * FT * 4;
  static constexpr int OFF_ANT = OFF_CS + MS * 16;
  static constexpr int OFF_GSLOT = OFF_ANT + MS * 8;
  static constexpr int OFF_RED = OFF_GSLOT + MS * 4;
  static constexpr int OFF_MBAR = OFF_RED + NWARP * 4 * 4;
  static constexpr int SMEM_BYTES = OFF_MBAR + 2 * 8;
  static_assert(MPT == 2 || MPT == 4, "phase F handles 2 or 4 rows per thread");
  static_assert(NQ * MS * FT * 4 <= 2 * MF * FT * 4, "dL/dv must fit over V");
  static_assert(QN >= 1 && MS % NWARP == 0, "phase Q: whole groups per warp");
  static_assert(SMEM_BYTES <= (NTHR == 256 ? 113 : 227) * 1024, "shared memory budget");
};

__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

template <int NTHR, int NQ, bool SINGLE, int KPM>
__global__ void __launch_bounds__(NTHR, NTHR == 256 ? 2 : 1) shared_kernel(const SharedParams p) {
  using C = SharedCfg<NTHR, NQ, KPM>;
  constexpr int FT = C::FT, MPT = C::MPT, PITCH = C::CT_PITCH, MS = C::MS, TK = C::TKMAX;
  constexpr bool SUM = NQ == 4;
  extern __shared__ __align__(1024) unsigned char smem[];
  const FitState* st = p.st;
  if (st->step > st->stop_after) return;  // fit already stopped (uniform across the grid)
  const MTileDesc mt = p.tiles[blockIdx.x];
  const int gsel = st->step & 1;
  const float* __restrict__ g_r = p.g_r[gsel];
  const float* __restrict__ g_i = p.g_i[gsel];

  float* Abuf = reinterpret_cast<float*>(smem + C::OFF_A);
  float* CT = reinterpret_cast<float*>(smem + C::OFF_CT);
  float* Vs = reinterpret_cast<float*>(smem + C::OFF_V);
  float* Qs = Vs;  // in place: thread (group, channel) of phase Q reads its four V values, then writes its NQ dL/dv values
  ClassSlot* s_cs = reinterpret_cast<ClassSlot*>(smem + C::OFF_CS);
  int2* s_ant = reinterpret_cast<int2*>(smem + C::OFF_ANT);
  int* s_gslot = reinterpret_cast<int*>(smem + C::OFF_GSLOT);
  float* red = reinterpret_cast<float*>(smem + C::OFF_RED);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::OFF_MBAR);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int kp = mt.kp, nslots = mt.nslots;
  const int ntiles = mt.j1 - mt.j0;
  const uint32_t tile_bytes = (uint32_t)kp * FT * 4u;
  const float* Abase = p.A + mt.a_off + (size_t)mt.j0 * kp * FT;
  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
    mbar_expect_tx(&mbar[0], tile_bytes);
    bulk_g2s(Abuf, Abase, tile_bytes, &mbar[0]);
    if (ntiles > 1) {
      mbar_expect_tx(&mbar[1], tile_bytes);
      bulk_g2s(Abuf + C::KROWS * FT, Abase + (size_t)kp * FT, tile_bytes, &mbar[1]);
    }
  }
  {
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // rows kp .. 32 ceil(kp / 32) - 1 are never written by the bulk copies: zero them once (phase B reads whole blocks of 32)
    const int tail4 = (((kp + 31) & ~31) - kp) * FT / 4;
    for (int e = tid; e < tail4; e += NTHR) {
      reinterpret_cast<float4*>(Abuf + kp * FT)[e] = zero4;
      reinterpret_cast<float4*>(Abuf + C::KROWS * FT + kp * FT)[e] = zero4;
    }
    // V / dL/dv rows of the groups this CTA does not have stay zero for the whole pass
    for (int e = tid; e < 2 * C::MF * FT / 4; e += NTHR) reinterpret_cast<float4*>(Vs)[e] = zero4;
  }
  if (tid < MS) {
    ClassSlot cs = {0, 0, 0, 0};
    int gs = 0;
    int2 ants = make_int2(0, 0);
    if (tid < nslots) {
      cs = p.cslots[mt.cs0 + tid];
      gs = p.cs_slot[mt.cs0 + tid];
      ants = make_int2(p.bl_ant0[cs.bl0] * p.nfp, p.bl_ant1[cs.bl0] * p.nfp);  // element offsets of the two gain rows
    }
    s_cs[tid] = cs;
    s_gslot[tid] = gs;
    s_ant[tid] = ants;
  }
  __syncthreads();
  // coefficients of the CTA's groups, transposed to [k][row] so that phase F reads its rows with one LDS.128 / LDS.64
  if (!p.init_mode) {
    for (int m = warp; m < C::MF; m += C::NWARP) {
      const int part = m / MS, s = m % MS;
      const bool valid = s < nslots;
      const float* src = (part ? p.c_i : p.c_r) + s_cs[s].coef0;
      for (int k = lane; k < kp; k += 32) CT[k * PITCH + m] = (valid && k < mt.ncomp) ? src[k] : 0.f;
    }
  }
  __syncthreads();

  // ---- thread roles ----
  // phase F: the CTA's two halves split the rows k of the tile (8-row blocks, alternately); inside a half 8 chunk lanes x
  // HALF / 8 row groups of MPT rows, a warp owns 4 MPT consecutive rows = one part, 4 MPT consecutive groups.  The halves'
  // partial sums are added (fixed order) by phase Q.
  const int f_half = tid / C::HALF, f_t = tid % C::HALF, f_fg = f_t & 7, f_m0 = (f_t >> 3) * MPT;
  const bool f_active = !p.init_mode && (((f_t >> 5) * 4 * MPT) % MS) < nslots;
  // phase B: lane = vector (k = lane + 32 t), warp = 8 consecutive rows = one q, 8 consecutive groups
  const int b_m0 = warp * 8;
  const bool b_active = !p.store_v && ((warp * 8) % MS) < nslots;
  const int b_sw = lane & 7;

  float acc[8][TK];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int t = 0; t < TK; ++t) acc[i][t] = 0.f;
  float loss_acc = 0.f, sr_acc = 0.f, si_acc = 0.f;

  for (int j = 0; j < ntiles; ++j) {
    const int buf = j & 1;
    mbar_wait(&mbar[buf], (j >> 1) & 1);
    const float* Ab = Abuf + buf * C::KROWS * FT;

    // ---------------- phase F ----------------
    if (f_active) {
      float4 v[MPT];
#pragma unroll
      for (int i = 0; i < MPT; ++i) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      const float* ctp = CT + f_m0;
      for (int k0 = 8 * f_half; k0 < kp; k0 += 16) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int k = k0 + u;
          const float4 a = *reinterpret_cast<const float4*>(Ab + k * FT + ((f_fg ^ u) << 2));
          if constexpr (MPT == 4) {
            const float4 c = *reinterpret_cast<const float4*>(ctp + k * PITCH);
            axpy4(c.x, a, v[0]);
            axpy4(c.y, a, v[1]);
            axpy4(c.z, a, v[2]);
            axpy4(c.w, a, v[3]);
          } else {
            const float2 c = *reinterpret_cast<const float2*>(ctp + k * PITCH);
            axpy4(c.x, a, v[0]);
            axpy4(c.y, a, v[1]);
          }
        }
      }
      float* vdst = Vs + f_half * C::MF * FT;
#pragma unroll
      for (int i = 0; i < MPT; ++i) *reinterpret_cast<float4*>(vdst + (f_m0 + i) * FT + f_fg * 4) = v[i];
    }
    __syncthreads();

    // ---------------- phase Q: lane = channel, warp w takes groups w, w + NWARP, ... ----------------
    {
      const int fo = (mt.j0 + j) * FT + lane;
      float in[C::QN][7];
      int bl[C::QN];
#pragma unroll
      for (int n = 0; n < C::QN; ++n) {  // all loads first
        const int s = warp + C::NWARP * n;
        bl[n] = -1;
        if (s < nslots) {
          bl[n] = s_cs[s].bl0;
          const int o = bl[n] * p.nfp + fo;  // nbls * nfp < 2^31 is checked at plan creation
          in[n][0] = p.d_r[o];
          in[n][1] = p.d_i[o];
          in[n][2] = p.w[o];
          if (!p.init_mode && !p.store_v) {
            const int2 an = s_ant[s];
            const int o0 = an.x + fo, o1 = an.y + fo;
            in[n][3] = g_r[o0];
            in[n][4] = g_i[o0];
            in[n][5] = g_r[o1];
            in[n][6] = g_i[o1];
          }
        }
      }
#pragma unroll
      for (int n = 0; n < C::QN; ++n) {
        const int s = warp + C::NWARP * n;
        if (bl[n] < 0) continue;
        float qr = 0.f, qi = 0.f, pw = 0.f, qw = 0.f;
        if (p.init_mode) {  // right-hand side of the coefficient initialisation: data * (w != 0)
          const float msk0 = (fabsf(in[n][2]) <= 1e-8f) ? 0.f : 1.f;  // np.isclose(w, 0): |w| <= atol = 1e-8
          qr = in[n][0] * msk0;
          qi = in[n][1] * msk0;
          if (!SINGLE) {
            const int nb = s_cs[s].nb;
            for (int b = 1; b < nb; ++b) {
              const size_t o = (size_t)(bl[n] + b) * p.nfp + fo;
              const float msk = (fabsf(p.w[o]) <= 1e-8f) ? 0.f : 1.f;
              qr += p.d_r[o] * msk;
              qi += p.d_i[o] * msk;
            }
          }
        } else {
          // V index (2 half + part) of this thread's column; the NQ values written below go to indices 0 .. NQ - 1
          const float v_r = Vs[s * FT + lane] + Vs[(2 * MS + s) * FT + lane];
          const float v_i = Vs[(MS + s) * FT + lane] + Vs[(3 * MS + s) * FT + lane];
          if (p.store_v) {
            p.vout[(size_t)s_gslot[s] * p.nfp + fo] = make_float2(v_r, v_i);
            continue;
          }
          // one visibility: model = g_i conj(g_j) v, weighted residual, chi^2, z, dL/dv (calibration.py:1593-1609)
          auto visibility = [&](size_t o, float dr, float di, float w, float gr0, float gi0, float gr1, float gi1) {
            const float P = gr0 * gr1 + gi0 * gi1;
            const float Q = gr0 * gi1 - gi0 * gr1;
            const float mr = P * v_r + Q * v_i;
            const float mi = P * v_i - Q * v_r;
            const float rr = dr - mr, ri = di - mi;
            loss_acc += (rr * rr + ri * ri) * w;
            const float er = -2.f * w * rr, ei = -2.f * w * ri;
            p.z[o] = make_float2(er * v_r + ei * v_i, er * v_i - ei * v_r);
            qr += P * er - Q * ei;
            qi += Q * er + P * ei;
            if (SUM) {
              p.y[o] = make_float2(w * v_r, w * v_i);
              sr_acc += w * mr;
              si_acc += w * mi;
              pw += P * w;
              qw += Q * w;
            }
          };
          visibility((size_t)(bl[n] * p.nfp + fo), in[n][0], in[n][1], in[n][2], in[n][3], in[n][4], in[n][5], in[n][6]);
          if (!SINGLE) {
            const int nb = s_cs[s].nb;
            for (int b = 1; b < nb; ++b) {
              const int bb = bl[n] + b;
              const size_t o = (size_t)bb * p.nfp + fo;
              const size_t o0 = (size_t)p.bl_ant0[bb] * p.nfp + fo, o1 = (size_t)p.bl_ant1[bb] * p.nfp + fo;
              visibility(o, p.d_r[o], p.d_i[o], p.w[o], g_r[o0], g_i[o0], g_r[o1], g_i[o1]);
            }
          }
        }
        Qs[s * FT + lane] = qr;
        Qs[(MS + s) * FT + lane] = qi;
        if (SUM) {
          Qs[(2 * MS + s) * FT + lane] = pw;
          Qs[(3 * MS + s) * FT + lane] = qw;
        }
      }
      // next tile's data / weight rows towards L2 while phase B runs
      if (j + 1 < ntiles && tid < 3 * MS) {
        const int arr = tid / MS, s = tid % MS;
        if (s < nslots) {
          const float* base = arr == 0 ? p.d_r : (arr == 1 ? p.d_i : p.w);
          prefetch_l2(base + (size_t)s_cs[s].bl0 * p.nfp + (mt.j0 + j + 1) * FT);
        }
      }
    }
    __syncthreads();

    // ---------------- phase B ----------------
    if (b_active) {
#pragma unroll 2  // (fully unrolled, the body overflows the instruction cache: 'no instruction' stalls in ncu)
      for (int c4 = 0; c4 < 8; ++c4) {
        float4 q4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) q4[i] = *reinterpret_cast<const float4*>(Qs + (b_m0 + i) * FT + c4 * 4);  // warp broadcast
        const float* ap = Ab + lane * FT + ((c4 ^ b_sw) << 2);
#pragma unroll
        for (int t = 0; t < TK; ++t) {
          if (t * 32 < kp) {  // CTA-uniform: blocks of 32 vectors past the class's last row are skipped, not multiplied
            const float4 a = *reinterpret_cast<const float4*>(ap + t * 32 * FT);
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i][t] = dot4(a, q4[i], acc[i][t]);
          }
        }
      }
    }
    __syncthreads();  // every warp is done with the tile buffer and with dL/dv (which the next phase F overwrites)
    if (tid == 0 && j + 2 < ntiles) {
      mbar_expect_tx(&mbar[buf], tile_bytes);
      bulk_g2s(Abuf + buf * C::KROWS * FT, Abase + (size_t)(j + 2) * kp * FT, tile_bytes, &mbar[buf]);
    }
  }

  // ---------------- backward sums: thread-private, no reduction ----------------
  if (b_active) {
    float* plane = p.dcpart + (size_t)mt.seg * p.dc_plane;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int mb = b_m0 + i, q = mb / MS, s = mb % MS;
      if (s < nslots) {
        float* dst = plane + (size_t)s_cs[s].row0 * NQ + q;
#pragma unroll
        for (int t = 0; t < TK; ++t) {
          const int k = lane + 32 * t;
          if (k < mt.ncomp) dst[(size_t)k * NQ] = acc[i][t];
        }
      }
    }
  }

  // ---------------- per-CTA partial sums (fixed order -> deterministic) ----------------
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
    sr_acc += __shfl_xor_sync(0xffffffffu, sr_acc, off);
    si_acc += __shfl_xor_sync(0xffffffffu, si_acc, off);
  }
  if (lane == 0) {
    red[warp * 4 + 0] = loss_acc;
    red[warp * 4 + 1] = sr_acc;
    red[warp * 4 + 2] = si_acc;
  }
  __syncthreads();
  if (tid == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int w = 0; w < C::NWARP; ++w) {
      a += (double)red[w * 4 + 0];
      b += (double)red[w * 4 + 1];
      c += (double)red[w * 4 + 2];
    }
    double* dst = p.partials + (size_t)blockIdx.x * 4;
    dst[0] = a;
    dst[1] = b;
    dst[2] = c;
  }
}

}  // namespace calb2
