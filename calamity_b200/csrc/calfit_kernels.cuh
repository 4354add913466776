// Device code of the B200-native gain-and-foreground fit (sm_100a).
//
// One optimizer iteration of calibration.py:663-668 is four launches (the coefficient update is forked onto a second
// stream next to the gain update; on several GPUs the gradient exchange goes through peer memory inside these kernels):
//   heavy_kernel   streams the ragged foreground basis ONCE (cp.async.bulk -> shared memory, mbarrier
//                  double buffering) and, per staged [rows x FT channels] tile, does the forward
//                  contraction v = sum_k c_k A_k (calibration.py:1587-1590), the gain application,
//                  weighted residual and chi^2 partial sums (1593-1609 / 1643-1651), dL/dv, and the
//                  backward contraction A . dL/dv accumulated in registers across the item's tiles.
//   finalize_kernel reduces the per-CTA partial sums in a fixed order (deterministic), forms the loss
//                  (1652-1656), records it, runs the use_min / tol logic of 699-717 on the device.
//   gains_kernel   deterministic per-(antenna, channel) reduction of the gain gradient over the
//                  baselines touching the antenna (CSR order, 8 entry groups, no atomics) + optimizer step on the gains.
//   coeffs_kernel  optimizer step on the foreground coefficients (combines the regulariser terms).
// float64 plans and groups too large for the staged tile use the kernels of calfit_generic.cuh instead.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace calb2 {

// ------------------------------------------------------------------------------------------------
// shared host/device structures
// ------------------------------------------------------------------------------------------------
struct ItemDesc {
  long long a_off;  // float offset of the item's first tile in the tiled basis
  int nrows;        // staged rows (multiple of the warp-step G), <= KMAX
  int nslots;       // model-visibility slots in the item, <= SMAX
  int row0;         // global (padded) row index of the item's first row
  int slot0;        // global slot index of the item's first slot
};

struct FitState {
  int step;          // train steps executed so far (0-based index of the step in flight)
  int stop_after;    // last step allowed to execute
  int nrec;          // recorded losses
  int upd_active;    // set by finalize when this step's updates must run
  int snap;          // use_min: this step's post-update parameters are the new optimum
  int any_snap;
  float min_loss;
  float prev_loss;
  float last_loss;
  float alpha;       // 2 (S_r - P_r)   (d/dS of the 'sum' regulariser, calibration.py:1654)
  float beta;        // 2 (S_i - P_i)
  float lr_t;        // bias-corrected step size of the update in flight
  float aux[4];      // per-step optimizer scalars (Nadam: 1 - m_schedule_new, 1 - m_schedule_next, 1 - u_t, u_{t+1}; and aux2[0] below)
  float aux2[2];     // Nadam: 1 - beta_2^t; spare
  double m_schedule; // Nadam: running product of the momentum schedule (Keras `_m_cache`)
  float s_r, s_i;
  double chi2;
  int error;         // 1: a peer did not publish its partials within the exchange time-out (the fit was aborted)
  int error_rank;    // the first rank that was late
};

struct FitConsts {
  int optimizer;     // CALB2_OPT_*
  float lr, beta1, beta2, eps;
  float rho, momentum, init_acc, l1, l2, lr_power;  // RMSprop / Adadelta / SGD / Adagrad / Ftrl (Keras names)
  float weight_decay;                               // LAMB (tensorflow_addons)
  int nesterov;
  int maxsteps;
  double tol;
  int use_min;
  int regularization;
  float prior_r, prior_i;
  int n_skip;        // n_profile_steps + 1 unrecorded steps
};

// Peer-memory exchange between the ranks of one node (one process per GPU; buffers shared with cudaIpc).  Every rank
// owns ONE exchange buffer: [flag | scalars[2][4] | gain-gradient partials[2][2 nants nfp]], double-buffered by the
// parity of a step sequence number.  A rank publishes its partials of step `seq` by storing seq + 1 to its flag with
// release.sys; consumers poll all flags with acquire.sys and then read every rank's partial straight from the
// owner's memory over NVLink, adding them in rank order -- deterministic and bit-identical on all ranks.
constexpr int CALB2_MAX_RANKS = 16;
constexpr int XBUF_FLAG_BYTES = 128;
constexpr int XBUF_SCAL_BYTES = 128;  // double [2][4] + pad
struct PeerView {
  const unsigned int* flag[CALB2_MAX_RANKS];
  const double* scal[CALB2_MAX_RANKS];  // [2][4]
  const float* grad[CALB2_MAX_RANKS];   // [2][2 * nants * nfp]
  int n;                                // ranks, 0 or 1: no exchange
};

__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// the partials written by the preceding kernels of this stream become visible to the peers, then the flag moves
__global__ void xpublish_kernel(unsigned int* flag, unsigned int value) {
  __threadfence_system();
  st_release_sys(flag, value);
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// One lane per rank waits until that rank has published step `value` -- but never longer than `timeout_ns` (a peer that
// died, diverged or was given different loop parameters would otherwise hang every other GPU of the node for good).
// Returns -1 when every flag arrived, else the first late rank (identical in every thread of the CTA).
__device__ __forceinline__ int xwait_all(const PeerView& pv, unsigned int value, unsigned long long timeout_ns) {
  __shared__ int s_late;
  if (threadIdx.x == 0) s_late = 0x7fffffff;
  __syncthreads();
  if ((int)threadIdx.x < pv.n) {
    const unsigned long long t0 = global_timer_ns();
    unsigned int spins = 0;
    while ((int)(ld_acquire_sys(pv.flag[threadIdx.x]) - value) < 0) {
      if ((++spins & 1023u) == 0u && global_timer_ns() - t0 > timeout_ns) {
        atomicMin(&s_late, (int)threadIdx.x);
        break;
      }
    }
  }
  __syncthreads();
  return s_late == 0x7fffffff ? -1 : s_late;
}
// after_update: 0 = wait before this step's finalize, 1 = wait before its gain update, 2 = unconditional (fit begin / close)
__global__ void xwait_kernel(const PeerView pv, unsigned int value, FitState* st, int mode, unsigned long long timeout_ns) {
  if (mode == 0 ? (st->step > st->stop_after) : (mode == 1 ? !st->upd_active : false)) return;
  const int late = xwait_all(pv, value, timeout_ns);
  if (late >= 0 && threadIdx.x == 0) {  // abort: nothing after this point may use the peers' partials
    st->error = 1;
    st->error_rank = late;
    st->upd_active = 0;
    st->stop_after = st->step - 1;
  }
}

struct HeavyParams {
  const float* A;
  const ItemDesc* items;
  const unsigned char* row_slot;  // [rows_total] slot index local to the item
  const int* row_coef;            // [rows_total] coefficient index, -1 for alignment rows
  const int* slot_row0;           // [nslots_total + 1] global row of each slot's first row
  const int* slot_bl0;            // [nslots_total] first baseline of each slot
  const int* slot_nb;             // [nslots_total] baselines of each slot
  const int* bl_ant0;
  const int* bl_ant1;
  const float* d_r;               // [nbls][nfp]
  const float* d_i;
  const float* w;
  const float* g_r[2];            // ping-pong gains [nants][nfp]
  const float* g_i[2];
  const float* c_r;               // [ncoef]
  const float* c_i;
  float2* z;                      // [nbls][nfp]  (a, b) of the gain gradient, chi^2 part
  float2* y;                      // [nbls][nfp]  w * v (regulariser part), only when SUM
  float* dcpart;                  // [rows_total][NQ] backward contractions per staged row
  float2* vout;                   // [nslots_total][nfp] model visibilities, written when store_v
  double* partials;               // [nitems][4]: chi^2, S_r, S_i
  const FitState* st;
  int nfp;                        // channels padded to a multiple of FT
  int ntiles;
  int store_v;
  int init_mode;                  // 1: dL/dv := data * (w != 0)   (right-hand side of the lstsq initialisation)
  // fused tail update (every group single-slot, no regulariser coupling): the item's own coefficients take
  // their optimizer step right after the backward sums are complete
  int fuse_update;
  int dbg;                        // -DCALB2_PROFILE builds only: phase clocks / timing ablations (CALB2_DBG)
  unsigned long long* dbg_out;    // [16] per-phase cycle sums of thread 0
  float* c_r_rw;
  float* c_i_rw;
  float* cm_r;
  float* cu_r;
  float* cm_i;
  float* cu_i;
  FitConsts k;
};

// ------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk asynchronous copy (TMA engine, SASS UBLKCP)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t addr = smem_u32(bar), done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!done);
}

// Development instrumentation, compiled only with -DCALB2_PROFILE (tools/build_prof.sh); the product library
// carries none of it.  CALB2_TICK accumulates thread 0's clock64 per phase; CALB2_DEP makes the following tick wait
// for a value (in-order issue).  Note that BAR.SYNC is "defer blocking": the wait of a barrier shows up at the first
// dependent shared-memory access after it, not at the barrier.
#ifdef CALB2_PROFILE
#define CALB2_TICK(n)                \
  if (prof) {                        \
    const long long now = clock64(); \
    tph[n] += now - tc;              \
    tc = now;                        \
  }
#define CALB2_DEP(x) \
  if (prof && (x) == 1.2345e-30f) tph[11] += 1;
#define CALB2_PROF_BEGIN                                     \
  const bool prof = (p.dbg & 128) && threadIdx.x == 0;      \
  long long tph[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; \
  long long tc = prof ? clock64() : 0;
#define CALB2_PROF_END                                                                 \
  if (prof) {                                                                          \
    for (int n = 0; n < 12; ++n) atomicAdd(p.dbg_out + n, (unsigned long long)tph[n]); \
    atomicAdd(p.dbg_out + 15, (unsigned long long)p.ntiles);                           \
  }
#else
#define CALB2_TICK(n)
#define CALB2_DEP(x)
#define CALB2_PROF_BEGIN
#define CALB2_PROF_END
#endif

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  acc = fmaf(a.w, b.w, acc);
  return acc;
}
__device__ __forceinline__ void axpy4(float c, const float4& a, float4& v) {
  v.x = fmaf(c, a.x, v.x);
  v.y = fmaf(c, a.y, v.y);
  v.z = fmaf(c, a.z, v.z);
  v.w = fmaf(c, a.w, v.w);
}

// ------------------------------------------------------------------------------------------------
// optimizer rules (Keras OptimizerV2; see oracle/restatement.py for provenance)
//   SPARSE = the IndexedSlices form used for the gains, dense form for the coefficients
// ------------------------------------------------------------------------------------------------
enum { OPT_ADAMAX = 0, OPT_ADAM = 1, OPT_SGD = 2, OPT_RMSPROP = 3, OPT_ADAGRAD = 4, OPT_ADADELTA = 5, OPT_NADAM = 6, OPT_FTRL = 7, OPT_LAMB = 8 };

// float / double overloads so the optimizer rules below are written once for both precisions
__device__ __forceinline__ float m_sqrt(float x) { return sqrtf(x); }
__device__ __forceinline__ double m_sqrt(double x) { return sqrt(x); }
__device__ __forceinline__ float m_rsqrt(float x) { return rsqrtf(x); }
__device__ __forceinline__ double m_rsqrt(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float m_abs(float x) { return fabsf(x); }
__device__ __forceinline__ double m_abs(double x) { return fabs(x); }
__device__ __forceinline__ float m_max(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double m_max(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float m_pow(float a, float b) { return powf(a, b); }
__device__ __forceinline__ double m_pow(double a, double b) { return pow(a, b); }

// One parameter, one step.  m / u are the parameter's two optimizer slots:
//   Adamax, Adam, Nadam: first / second moment      SGD: momentum accumulator (m)        RMSprop: momentum (m), rms (u)
//   Adagrad: accumulator (u, starts at initial_accumulator_value)                         Adadelta: accum (m), accum_update (u)
//   Ftrl: accumulator (m, starts at initial_accumulator_value), linear (u)
// K / S are the hyper-parameter and state structs of the precision T (FitConsts / FitState for float32).
template <bool SPARSE, class T, class K, class S>
__device__ __forceinline__ T opt_step(const K& k, const S* st, T theta, T g, T& m, T& u, T lr_t) {
  const T one = (T)1;
  switch (k.optimizer) {
    case OPT_ADAMAX:
      m = SPARSE ? (m * k.beta1 + g * (one - k.beta1)) : (m + (g - m) * (one - k.beta1));
      u = m_max(u * k.beta2, m_abs(g));
      return theta - lr_t * (m / (u + k.eps));
    case OPT_ADAM:
      m = SPARSE ? (m * k.beta1 + g * (one - k.beta1)) : (m + (g - m) * (one - k.beta1));
      u = SPARSE ? (u * k.beta2 + (g * g) * (one - k.beta2)) : (u + (g * g - u) * (one - k.beta2));
      return theta - (m * lr_t) / (m_sqrt(u) + k.eps);
    case OPT_SGD:  // ApplyGradientDescent / ApplyKerasMomentum
      if (k.momentum == (T)0) return theta - lr_t * g;
      m = m * k.momentum - lr_t * g;
      return k.nesterov ? theta + m * k.momentum - lr_t * g : theta + m;
    case OPT_RMSPROP:  // centered = False
      u = k.rho * u + (one - k.rho) * (g * g);
      if (k.momentum > (T)0) {  // fused ApplyRMSProp: epsilon inside the square root
        m = k.momentum * m + lr_t * g * m_rsqrt(u + k.eps);
        return theta - m;
      }
      return theta - lr_t * g / (m_sqrt(u) + k.eps);
    case OPT_ADAGRAD:  // ApplyAdagradV2
      u = u + g * g;
      return theta - lr_t * g / (m_sqrt(u) + k.eps);
    case OPT_ADADELTA: {  // ApplyAdadelta
      m = m * k.rho + (g * g) * (one - k.rho);
      const T upd = m_sqrt(u + k.eps) * m_rsqrt(m + k.eps) * g;
      u = u * k.rho + (upd * upd) * (one - k.rho);
      return theta - upd * lr_t;
    }
    case OPT_NADAM: {  // Keras Nadam: aux = {1 - m_schedule_new, 1 - m_schedule_next, 1 - u_t, u_{t+1}}, aux2[0] = 1 - beta_2^t
      const T g_prime = g / st->aux[0];
      m = k.beta1 * m + (one - k.beta1) * g;
      const T m_prime = m / st->aux[1];
      u = k.beta2 * u + (one - k.beta2) * (g * g);
      const T v_prime = u / st->aux2[0];
      const T m_bar = st->aux[2] * g_prime + st->aux[3] * m_prime;
      return theta - lr_t * m_bar / (m_sqrt(v_prime) + k.eps);
    }
    default: {  // OPT_FTRL, ApplyFtrlV2 without shrinkage
      const T acc_new = m + g * g;
      const bool half = k.lr_power == (T)-0.5;
      const T p_new = half ? m_sqrt(acc_new) : m_pow(acc_new, -k.lr_power);
      const T p_old = half ? m_sqrt(m) : m_pow(m, -k.lr_power);
      u += g - (p_new - p_old) / lr_t * theta;
      const T quadratic = p_new / lr_t + (T)2 * k.l2;
      m = acc_new;
      const T sgn = u > (T)0 ? one : (u < (T)0 ? -one : (T)0);
      return m_abs(u) > k.l1 ? (sgn * k.l1 - u) / quadratic : (T)0;
    }
  }
}

// tensorflow_addons LAMB (calibration.py:15, 26; tfa.optimizers.LAMB._resource_apply_dense): Adam moments with bias correction,
// update = m_hat / (sqrt(v_hat) + eps) + weight_decay * theta, then theta -= lr * ratio * update with ONE trust ratio
// ||theta|| / ||update|| per variable (g_r, g_i, every chunk's fg_r / fg_i tensor).  Two passes: lamb_moments() updates the
// moments and returns the update (its square and theta's are summed per variable by the callers, fixed order), lamb_update()
// recomputes the update from the stored moments.  st->aux[0] = 1 - beta_1^t, st->aux[1] = 1 - beta_2^t.
template <class K, class S>
__device__ __forceinline__ float lamb_update(const K& k, const S* st, float theta, float m, float u) {
  return (m / st->aux[0]) / (sqrtf(u / st->aux[1]) + k.eps) + k.weight_decay * theta;
}
template <class K, class S>
__device__ __forceinline__ float lamb_moments(const K& k, const S* st, float theta, float g, float& m, float& u) {
  m = m * k.beta1 + g * (1.f - k.beta1);
  u = u * k.beta2 + (g * g) * (1.f - k.beta2);
  return lamb_update(k, st, theta, m, u);
}
struct LambVars {
  // partial sums (theta^2, update^2) per CTA: the gain kernel's CTAs hold g_r in [0], [1] and g_i in [2], [3]; the coefficient
  // kernel's CTAs of chunk c (its own launch) likewise for fg_r[c] / fg_i[c]
  double* gain_partials;   // [gain CTAs][4]
  double* coef_partials;   // [coefficient CTAs][4]
  float* ratio;            // [2 + 2 nchunks]: g_r, g_i, then (fg_r[c], fg_i[c]) per chunk
};

// Keras local_step = iterations + 1; powers evaluated in double and rounded once (identical in every kernel)
template <class T, class K>
__device__ __forceinline__ T bias_corrected_lr_t(const K& k, int step) {
  const double tt = (double)(step + 1);
  if (k.optimizer == OPT_ADAMAX) return k.lr / ((T)1 - (T)pow((double)k.beta1, tt));
  if (k.optimizer == OPT_ADAM) {
    const T b1p = (T)pow((double)k.beta1, tt);
    const T b2p = (T)pow((double)k.beta2, tt);
    return k.lr * m_sqrt((T)1 - b2p) / ((T)1 - b1p);
  }
  return k.lr;
}
__device__ __forceinline__ float bias_corrected_lr(const FitConsts& k, int step) { return bias_corrected_lr_t<float>(k, step); }

// per-step scalars of the Nadam momentum schedule (Keras Nadam._prepare_local, decay 0.004), kept in the fit state
template <class T, class K, class S>
__device__ __forceinline__ void nadam_schedule(const K& k, S* st, int t) {
  const double ls = (double)(t + 1);
  const T u_t = k.beta1 * ((T)1 - (T)0.5 * (T)pow(0.96, 0.004 * ls));
  const T u_t1 = k.beta1 * ((T)1 - (T)0.5 * (T)pow(0.96, 0.004 * (ls + 1.0)));
  const T ms_new = (T)(t == 0 ? 1.0 : st->m_schedule) * u_t;
  const T ms_next = ms_new * u_t1;
  st->m_schedule = (double)ms_new;
  st->aux[0] = (T)1 - ms_new;
  st->aux[1] = (T)1 - ms_next;
  st->aux[2] = (T)1 - u_t;
  st->aux[3] = u_t1;
  st->aux2[0] = (T)1 - (T)pow((double)k.beta2, ls);
}

// ------------------------------------------------------------------------------------------------
// heavy kernel configuration
//   FL   lanes along frequency (a lane owns a float4 of channels)  -> FT = 4 FL channels per tile
//   G    rows a warp reads per step (32 / FL): G consecutive rows are G*FT*4 contiguous bytes, so every
//        warp-wide LDS.128 touches one contiguous 512-byte span (bank-conflict free)
//   RPT  steps per warp; KMAX = 8 warps * RPT * G rows per item
// ------------------------------------------------------------------------------------------------
template <int FL_, bool SUM_, int RPT_>
struct HeavyCfg {
  static constexpr int FL = FL_;
  static constexpr bool SUM = SUM_;
  static constexpr int RPT = RPT_;
  static constexpr int NTHR = 256;
  static constexpr int NWARP = 8;
  static constexpr int G = 32 / FL;
  static constexpr int FT = FL * 4;
  static constexpr int NSTEP = NWARP * RPT;
  static constexpr int KMAX = NSTEP * G;
  static constexpr int TILE_FLOATS = KMAX * FT;
  static constexpr int SMAX = 8;
  static constexpr int NQ = SUM ? 4 : 2;
  static constexpr int NSEG = NWARP + SMAX;
  static constexpr int NBUF = 2;
  // shared memory carve-up (bytes)
  static constexpr int OFF_A = 0;
  static constexpr int OFF_VPART = OFF_A + NBUF * TILE_FLOATS * 4;
  static constexpr int OFF_QBUF = OFF_VPART + NSEG * 2 * FT * 4;
  // coefficients, thread-major: [warp][row group][RPT_PAD] float2, so a thread fetches the coefficients of two
  // consecutive steps with one LDS.128 (half the coefficient traffic on the shared-memory port)
  static constexpr int RPT_PAD = (RPT + 1) & ~1;
  static constexpr int OFF_CBUF = OFF_QBUF + SMAX * NQ * FT * 4;
  static constexpr int OFF_STEPSLOT = OFF_CBUF + NWARP * G * RPT_PAD * 8;
  static constexpr int OFF_SLOTSTEP0 = OFF_STEPSLOT + ((NSTEP + 15) / 16) * 16;
  static constexpr int OFF_SLOTBL0 = OFF_SLOTSTEP0 + (SMAX + 1) * 4 + 12;
  static constexpr int OFF_SLOTNB = OFF_SLOTBL0 + (SMAX + 1) * 4 + 12;
  static constexpr int OFF_RED = OFF_SLOTNB + (SMAX + 1) * 4 + 12;
  static constexpr int OFF_MBAR = OFF_RED + NWARP * 4 * 4;
  static constexpr int SMEM_BYTES = OFF_MBAR + NBUF * 8;
};

// QMODE specialises phase Q, which is a serial per-warp instruction stream on the critical path of every tile
// (measured with clock64: ~1500 of ~3800 cycles per tile before the specialisation):
//   QM_SINGLE   every slot has exactly one baseline (the per-baseline DPSS layout): straight-line code
//   QM_GENERAL  slots with several (redundant) baselines: the first from prefetched registers, the rest in a loop
//   QM_INIT     dL/dv := data * (w != 0), the right-hand side of the coefficient initialisation
enum { QM_SINGLE = 0, QM_GENERAL = 1, QM_INIT = 2 };

template <int FL, bool SUM, int RPT, int MINB, int QMODE>
__global__ void __launch_bounds__(256, MINB) heavy_kernel(const HeavyParams p) {
  using C = HeavyCfg<FL, SUM, RPT>;
  constexpr int G = C::G, FT = C::FT, NQ = C::NQ;
  extern __shared__ __align__(1024) unsigned char smem[];

  const FitState* st = p.st;
  if (st->step > st->stop_after) return;  // fit already stopped (uniform across the grid)
  const int gsel = st->step & 1;
  const float* __restrict__ g_r = p.g_r[gsel];
  const float* __restrict__ g_i = p.g_i[gsel];

  float* Abuf = reinterpret_cast<float*>(smem + C::OFF_A);
  float* vpart = reinterpret_cast<float*>(smem + C::OFF_VPART);
  float* qbuf = reinterpret_cast<float*>(smem + C::OFF_QBUF);
  float2* cbuf = reinterpret_cast<float2*>(smem + C::OFF_CBUF);
  unsigned char* step_slot = smem + C::OFF_STEPSLOT;
  int* slot_step0 = reinterpret_cast<int*>(smem + C::OFF_SLOTSTEP0);
  int* slot_bl0 = reinterpret_cast<int*>(smem + C::OFF_SLOTBL0);
  int* slot_nbs = reinterpret_cast<int*>(smem + C::OFF_SLOTNB);
  float* red = reinterpret_cast<float*>(smem + C::OFF_RED);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + C::OFF_MBAR);

  const ItemDesc it = p.items[blockIdx.x];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int usub = lane / FL, fl = lane % FL;
  const int nsteps = it.nrows / G;
  const uint32_t tile_bytes = (uint32_t)it.nrows * FT * 4u;
  const float* Abase = p.A + it.a_off;

  if (tid == 0) {
    mbar_init(&mbar[0], 1);
    mbar_init(&mbar[1], 1);
    mbar_fence_init();
    mbar_expect_tx(&mbar[0], tile_bytes);  // tiles 0 and 1 stream in while the item's metadata is gathered
    bulk_g2s(Abuf, Abase, tile_bytes, &mbar[0]);
    if (p.ntiles > 1) {
      mbar_expect_tx(&mbar[1], tile_bytes);
      bulk_g2s(Abuf + C::TILE_FLOATS, Abase + (size_t)it.nrows * FT, tile_bytes, &mbar[1]);
    }
  }
  // Rows past the item's last row are zero in both tile buffers (the bulk copies never touch them) and
  // carry zero coefficients, so every warp that owns at least one row runs all RPT steps unguarded.
  {
    const int tail4 = (C::KMAX - it.nrows) * FT / 4;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int e = tid; e < tail4; e += C::NTHR) {
      reinterpret_cast<float4*>(Abuf + it.nrows * FT)[e] = zero4;
      reinterpret_cast<float4*>(Abuf + C::TILE_FLOATS + it.nrows * FT)[e] = zero4;
    }
  }
  for (int e = tid; e < C::NWARP * G * C::RPT_PAD; e += C::NTHR) {
    // e = (warp * G + row group) * RPT_PAD + step  <->  row = ((warp * RPT + step) * G + row group)
    const int i = e % C::RPT_PAD, wu = e / C::RPT_PAD;
    const int r = ((wu / G) * RPT + i) * G + (wu % G);
    float2 c = make_float2(0.f, 0.f);
    if (i < RPT && r < it.nrows && QMODE != QM_INIT) {
      const int ci = p.row_coef[it.row0 + r];
      if (ci >= 0) c = make_float2(p.c_r[ci], p.c_i[ci]);
    }
    cbuf[e] = c;
  }
  for (int s = tid; s < nsteps; s += C::NTHR) step_slot[s] = p.row_slot[it.row0 + s * G];
  for (int s = tid; s <= it.nslots; s += C::NTHR) {
    slot_step0[s] = (p.slot_row0[it.slot0 + s] - it.row0) / G;
    if (s < it.nslots) {
      slot_bl0[s] = p.slot_bl0[it.slot0 + s];
      slot_nbs[s] = p.slot_nb[it.slot0 + s];
    }
  }
  __syncthreads();

  // warp-uniform description of this warp's RPT steps: first slot and a bit per step that starts a new slot
  const int step_base = warp * RPT;
  const bool warp_active = step_base < nsteps;
  int s_first = 0;
  unsigned chg = 0u;
  if (warp_active) {
    s_first = step_slot[step_base];
    int prev = s_first;
#pragma unroll
    for (int i = 1; i < RPT; ++i) {
      if (step_base + i < nsteps) {
        const int s = step_slot[step_base + i];
        if (s != prev) chg |= 1u << i;
        prev = s;
      }
    }
  }

  // Phase Q: this thread owns the (slot, channel) elements e = tid + m * NTHR of EVERY tile, so everything that does
  // not depend on the tile index is worked out once per item: offsets of the first baseline's data / gain rows, the
  // warps whose partial sums belong to the slot (bit mask), the partial-sum and q-buffer addresses.  The inputs of
  // the first baseline are prefetched into registers one tile ahead (issued right after Q(j) for tile j + 1), so
  // their DRAM / L2 latency is covered by phase B, the tile wait and phase F.
  constexpr int EPT = (C::SMAX * FT + C::NTHR - 1) / C::NTHR;
  float pf[EPT][7];
  int qoff[EPT], qoff0[EPT], qoff1[EPT];  // element offsets at tile 0 (data row, ant0 row, ant1 row); qoff < 0: none
  int q_vp[EPT], q_qb[EPT], q_b0[EPT], q_nb[EPT], q_vo[EPT];
  unsigned q_wmask[EPT];
#pragma unroll
  for (int m = 0; m < EPT; ++m) {
    const int e = tid + m * C::NTHR;
    qoff[m] = -1;
    qoff0[m] = qoff1[m] = 0;
    q_vp[m] = q_qb[m] = q_b0[m] = q_nb[m] = q_vo[m] = 0;
    q_wmask[m] = 0u;
    if (e < it.nslots * FT) {
      const int s = e / FT, f = e % FT;
      const int b = slot_bl0[s];
      const int w_lo = slot_step0[s] / RPT, w_hi = (slot_step0[s + 1] - 1) / RPT;
      q_wmask[m] = ((2u << w_hi) - 1u) & ~((1u << w_lo) - 1u);
      q_vp[m] = (s * 2) * FT + f;        // partial of warp w: + w * 2 * FT (real), + FT more (imaginary)
      q_qb[m] = (s * NQ) * FT + f;
      q_b0[m] = b;
      q_nb[m] = slot_nbs[s];
      q_vo[m] = (it.slot0 + s) * p.nfp + f;
      if (q_nb[m] > 0) {
        qoff[m] = b * p.nfp + f;
        qoff0[m] = p.bl_ant0[b] * p.nfp + f;
        qoff1[m] = p.bl_ant1[b] * p.nfp + f;
      }
    }
  }
  auto prefetch_q = [&](int jt) {
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      if (qoff[m] >= 0) {
        const int o = qoff[m] + jt * FT;
        pf[m][0] = p.d_r[o];
        pf[m][1] = p.d_i[o];
        pf[m][2] = p.w[o];
        if (QMODE != QM_INIT) {
          const int o0 = qoff0[m] + jt * FT, o1 = qoff1[m] + jt * FT;
          pf[m][3] = g_r[o0];
          pf[m][4] = g_i[o0];
          pf[m][5] = g_r[o1];
          pf[m][6] = g_i[o1];
        }
      }
    }
  };
  prefetch_q(0);

  float acc[RPT][NQ];
#pragma unroll
  for (int i = 0; i < RPT; ++i)
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[i][q] = 0.f;
  float loss_acc = 0.f, sr_acc = 0.f, si_acc = 0.f;
  const int row_off = (step_base * G + usub) * FT + fl * 4;  // this thread's float offset of step 0 in a tile

  CALB2_PROF_BEGIN
  for (int j = 0; j < p.ntiles; ++j) {
    const int buf = j & 1;
    mbar_wait(&mbar[buf], (j >> 1) & 1);
    CALB2_TICK(0)
    const float* Ab = Abuf + buf * C::TILE_FLOATS + row_off;
    const float4* cb4 = reinterpret_cast<const float4*>(cbuf + (warp * G + usub) * C::RPT_PAD);  // two steps per load

    // ---------------- phase F: forward contraction, partial per (warp, slot) ----------------
    if (warp_active) {
      float4 vr = make_float4(0.f, 0.f, 0.f, 0.f), vi = vr;
      auto flush = [&](int seg) {
#pragma unroll
        for (int off = FL; off < 32; off <<= 1) {
          vr.x += __shfl_xor_sync(0xffffffffu, vr.x, off);
          vr.y += __shfl_xor_sync(0xffffffffu, vr.y, off);
          vr.z += __shfl_xor_sync(0xffffffffu, vr.z, off);
          vr.w += __shfl_xor_sync(0xffffffffu, vr.w, off);
          vi.x += __shfl_xor_sync(0xffffffffu, vi.x, off);
          vi.y += __shfl_xor_sync(0xffffffffu, vi.y, off);
          vi.z += __shfl_xor_sync(0xffffffffu, vi.z, off);
          vi.w += __shfl_xor_sync(0xffffffffu, vi.w, off);
        }
        if (usub == 0) {
          *reinterpret_cast<float4*>(vpart + (seg * 2 + 0) * FT + fl * 4) = vr;
          *reinterpret_cast<float4*>(vpart + (seg * 2 + 1) * FT + fl * 4) = vi;
        }
      };
      float4 cc = make_float4(0.f, 0.f, 0.f, 0.f);
      if (chg == 0u) {  // all of this warp's rows belong to one slot: branch-free stream
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(Ab + i * G * FT);
          if ((i & 1) == 0) cc = cb4[i >> 1];
          const float2 c = (i & 1) ? make_float2(cc.z, cc.w) : make_float2(cc.x, cc.y);
          axpy4(c.x, a, vr);
          axpy4(c.y, a, vi);
        }
        flush(warp + s_first);
      } else {
        int cur = s_first;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          if (i > 0 && ((chg >> i) & 1u)) {
            flush(warp + cur);
            vr = make_float4(0.f, 0.f, 0.f, 0.f);
            vi = vr;
            ++cur;
          }
          const float4 a = *reinterpret_cast<const float4*>(Ab + i * G * FT);
          if ((i & 1) == 0) cc = cb4[i >> 1];
          const float2 c = (i & 1) ? make_float2(cc.z, cc.w) : make_float2(cc.x, cc.y);
          axpy4(c.x, a, vr);
          axpy4(c.y, a, vi);
        }
        flush(warp + cur);
      }
    }
    CALB2_TICK(1)
    __syncthreads();

    // ---------------- phase Q: gains, model, residual, chi^2, dL/dv ----------------
#pragma unroll
    for (int m = 0; m < EPT; ++m) {
      if (q_wmask[m] != 0u) {  // this thread owns an element (uniform over tiles)
        float v_r = 0.f, v_i = 0.f;
#pragma unroll
        for (int w = 0; w < C::NWARP; ++w) {  // all loads issue back to back; summation order is fixed
          if ((q_wmask[m] >> w) & 1u) {
            v_r += vpart[q_vp[m] + w * 2 * FT];
            v_i += vpart[q_vp[m] + w * 2 * FT + FT];
          }
        }
        CALB2_DEP(v_r + v_i)
        CALB2_TICK(2)
        float qr = 0.f, qi = 0.f, pw = 0.f, qw = 0.f;
        // one visibility: model = g_i conj(g_j) v, weighted residual, chi^2, z, dL/dv (calibration.py:1593-1609)
        auto visibility = [&](int o, float dr, float di, float w, float gr0, float gi0, float gr1, float gi1) {
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
        if (QMODE == QM_INIT) {  // right-hand side of the coefficient initialisation: data * (w != 0)
          if (qoff[m] >= 0) {
            const float msk0 = (fabsf(pf[m][2]) <= 1e-8f) ? 0.f : 1.f;  // np.isclose(w, 0): |w| <= atol = 1e-8
            qr = pf[m][0] * msk0;
            qi = pf[m][1] * msk0;
            for (int n = 1; n < q_nb[m]; ++n) {
              const size_t o = (size_t)(q_b0[m] + n) * p.nfp + (qoff[m] - q_b0[m] * p.nfp) + j * FT;
              const float msk = (fabsf(p.w[o]) <= 1e-8f) ? 0.f : 1.f;
              qr += p.d_r[o] * msk;
              qi += p.d_i[o] * msk;
            }
          }
        } else if (QMODE == QM_SINGLE) {
          visibility(qoff[m] + j * FT, pf[m][0], pf[m][1], pf[m][2], pf[m][3], pf[m][4], pf[m][5], pf[m][6]);
        } else if (qoff[m] >= 0) {
          visibility(qoff[m] + j * FT, pf[m][0], pf[m][1], pf[m][2], pf[m][3], pf[m][4], pf[m][5], pf[m][6]);
          const int fg = (qoff[m] - q_b0[m] * p.nfp) + j * FT;
          for (int n = 1; n < q_nb[m]; ++n) {
            const int b = q_b0[m] + n;
            const int o = b * p.nfp + fg;
            const int o0 = p.bl_ant0[b] * p.nfp + fg, o1 = p.bl_ant1[b] * p.nfp + fg;
            visibility(o, p.d_r[o], p.d_i[o], p.w[o], g_r[o0], g_i[o0], g_r[o1], g_i[o1]);
          }
        }
        qbuf[q_qb[m]] = qr;
        qbuf[q_qb[m] + FT] = qi;
        if (SUM) {
          qbuf[q_qb[m] + 2 * FT] = pw;
          qbuf[q_qb[m] + 3 * FT] = qw;
        }
        if (p.store_v) p.vout[q_vo[m] + j * FT] = make_float2(v_r, v_i);
      }
    }
    CALB2_TICK(3)
    if (j + 1 < p.ntiles) prefetch_q(j + 1);
    CALB2_TICK(4)
    __syncthreads();

    // ---------------- phase B: backward contraction, accumulated in registers ----------------
    if (warp_active) {
      float4 q0, q1, q2, q3;
      auto load_q = [&](int s) {
        q0 = *reinterpret_cast<const float4*>(qbuf + (s * NQ + 0) * FT + fl * 4);
        q1 = *reinterpret_cast<const float4*>(qbuf + (s * NQ + 1) * FT + fl * 4);
        if (SUM) {
          q2 = *reinterpret_cast<const float4*>(qbuf + (s * NQ + 2) * FT + fl * 4);
          q3 = *reinterpret_cast<const float4*>(qbuf + (s * NQ + 3) * FT + fl * 4);
        }
      };
      load_q(s_first);
      CALB2_DEP(q0.x + q1.x)
      CALB2_TICK(5)
      if (chg == 0u) {
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          const float4 a = *reinterpret_cast<const float4*>(Ab + i * G * FT);
          acc[i][0] = dot4(a, q0, acc[i][0]);
          acc[i][1] = dot4(a, q1, acc[i][1]);
          if (SUM) {
            acc[i][2] = dot4(a, q2, acc[i][2]);
            acc[i][3] = dot4(a, q3, acc[i][3]);
          }
        }
      } else {
        int cur = s_first;
#pragma unroll
        for (int i = 0; i < RPT; ++i) {
          if (i > 0 && ((chg >> i) & 1u)) load_q(++cur);
          const float4 a = *reinterpret_cast<const float4*>(Ab + i * G * FT);
          acc[i][0] = dot4(a, q0, acc[i][0]);
          acc[i][1] = dot4(a, q1, acc[i][1]);
          if (SUM) {
            acc[i][2] = dot4(a, q2, acc[i][2]);
            acc[i][3] = dot4(a, q3, acc[i][3]);
          }
        }
      }
    }
    // All warps are done with the tile buffer: refill it with tile j + 2.  (Replacing this barrier by an mbarrier
    // "buffer empty" handshake that only thread 0 waits on was measured 23 % SLOWER: 151 vs 196 it/s at HERA-350.)
    CALB2_TICK(6)
    __syncthreads();
    if (tid == 0 && j + 2 < p.ntiles) {
      mbar_expect_tx(&mbar[buf], tile_bytes);
      bulk_g2s(Abuf + buf * C::TILE_FLOATS, Abase + (size_t)(j + 2) * it.nrows * FT, tile_bytes, &mbar[buf]);
    }
    CALB2_TICK(7)
  }
  __syncthreads();  // the epilogue reuses the tile buffers
  CALB2_PROF_END

  // ---------------- item epilogue: lane reduction of the backward sums ----------------
  // The reduced sums are staged in shared memory (the tile buffers are free now) so that the write-out --
  // or, in the fused case, the optimizer step on the item's own coefficients -- is done by all 256 threads
  // with coalesced accesses instead of by one lane per row.
  float* rowdc = Abuf;  // [nrows][NQ]
#pragma unroll
  for (int i = 0; i < RPT; ++i) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float v = acc[i][q];
#pragma unroll
      for (int off = 1; off < FL; off <<= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
      acc[i][q] = v;
    }
    const int stp = step_base + i;
    if (stp < nsteps && fl == 0) {
      float* dst = rowdc + (stp * G + usub) * NQ;
      if (SUM)
        *reinterpret_cast<float4*>(dst) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      else
        *reinterpret_cast<float2*>(dst) = make_float2(acc[i][0], acc[i][1]);
    }
  }
  __syncthreads();
  if (!SUM && p.fuse_update) {
    const float lr_fused = bias_corrected_lr(p.k, st->step);
    for (int r = tid; r < it.nrows; r += C::NTHR) {
      const int ci = p.row_coef[it.row0 + r];
      if (ci >= 0) {
        const float2 g = *reinterpret_cast<const float2*>(rowdc + r * 2);
        float mr = p.cm_r[ci], ur = p.cu_r[ci], mi = p.cm_i[ci], ui = p.cu_i[ci];
        const float nr = opt_step<false, float>(p.k, st, p.c_r_rw[ci], g.x, mr, ur, lr_fused);
        const float ni = opt_step<false, float>(p.k, st, p.c_i_rw[ci], g.y, mi, ui, lr_fused);
        p.cm_r[ci] = mr;
        p.cu_r[ci] = ur;
        p.cm_i[ci] = mi;
        p.cu_i[ci] = ui;
        p.c_r_rw[ci] = nr;
        p.c_i_rw[ci] = ni;
      }
    }
  } else {
    float* dst = p.dcpart + (size_t)it.row0 * NQ;
    for (int e = tid; e < it.nrows * NQ; e += C::NTHR) dst[e] = rowdc[e];
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

// ------------------------------------------------------------------------------------------------
// finalize: deterministic reduction of the per-CTA partials + loop control (calibration.py:699-717)
// ------------------------------------------------------------------------------------------------
struct FinalizeParams {
  const double* partials;
  int nitems;
  FitState* st;
  FitConsts k;
  float* hist;
  int eval_only;  // 1: just publish loss / alpha / beta, no loop bookkeeping
  PeerView peers; // peers.n > 1: wait until every rank's flag reached `xwait`, then add their scalars in rank order
  unsigned int xwait;
  int xpar;       // parity of the step sequence number (selects the half of the double-buffered partials)
  unsigned long long timeout_ns;  // bound on the wait for the peers' flags
};

__global__ void __launch_bounds__(1024, 1) finalize_kernel(const FinalizeParams p) {
  __shared__ double sh[3][32];
  FitState* st = p.st;
  if (!p.eval_only && st->step > st->stop_after) {
    if (threadIdx.x == 0) st->upd_active = 0;
    return;
  }
  double a = 0.0, b = 0.0, c = 0.0;
  if (p.peers.n > 1) {
    const int late = xwait_all(p.peers, p.xwait, p.timeout_ns);
    if (threadIdx.x != 0) return;
    if (late >= 0) {  // a peer never published: abort the fit instead of spinning (the host reports CALB2_ERR_TIMEOUT)
      st->error = 1;
      st->error_rank = late;
      st->upd_active = 0;
      st->stop_after = st->step - 1;
      return;
    }
    const int par = p.xpar;
    for (int r = 0; r < p.peers.n; ++r) {  // rank order: identical sums on every rank
      const volatile double* sc = p.peers.scal[r] + par * 4;
      a += sc[0];
      b += sc[1];
      c += sc[2];
    }
  } else {
    for (int i = threadIdx.x; i < p.nitems; i += blockDim.x) {
      a += p.partials[(size_t)i * 4 + 0];
      b += p.partials[(size_t)i * 4 + 1];
      c += p.partials[(size_t)i * 4 + 2];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, off);
      b += __shfl_xor_sync(0xffffffffu, b, off);
      c += __shfl_xor_sync(0xffffffffu, c, off);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
      sh[0][warp] = a;
      sh[1][warp] = b;
      sh[2][warp] = c;
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    a = b = c = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += sh[0][w];
      b += sh[1][w];
      c += sh[2][w];
    }
  }
  float loss = (float)a;
  float alpha = 0.f, beta = 0.f;
  if (p.k.regularization == 1) {
    const float dr = (float)b - p.k.prior_r, di = (float)c - p.k.prior_i;
    loss = loss + dr * dr + di * di;
    alpha = 2.f * dr;
    beta = 2.f * di;
  }
  st->chi2 = a;
  st->s_r = (float)b;
  st->s_i = (float)c;
  st->alpha = alpha;
  st->beta = beta;
  st->last_loss = loss;
  if (p.eval_only) return;

  const int t = st->step;
  st->lr_t = bias_corrected_lr(p.k, t);
  if (p.k.optimizer == OPT_NADAM) nadam_schedule<float>(p.k, st, t);
  if (p.k.optimizer == OPT_LAMB) {
    st->aux[0] = 1.f - (float)pow((double)p.k.beta1, (double)(t + 1));
    st->aux[1] = 1.f - (float)pow((double)p.k.beta2, (double)(t + 1));
  }
  int snap = 0;
  const int rec = t - p.k.n_skip;
  if (rec >= 0) {
    p.hist[rec] = loss;
    st->nrec = rec + 1;
    if (p.k.use_min && loss < st->min_loss) {
      st->min_loss = loss;
      snap = 1;
      st->any_snap = 1;
    }
    if (rec >= 1 && fabs((double)(loss - st->prev_loss)) < p.k.tol) st->stop_after = t;
    if (rec + 1 >= p.k.maxsteps) st->stop_after = t;
    st->prev_loss = loss;
  }
  st->snap = snap;
  st->upd_active = 1;
  st->step = t + 1;
}

struct GainsParams {
  const float2* z;
  const float2* y;
  const int* ant_ptr;   // [nants + 1] CSR over antennas
  const int* ant_ent;   // (baseline << 1) | side, ascending baseline order
  const int* ant_partner;  // the other antenna of that baseline
  const int* bl_ant0;
  const int* bl_ant1;
  float* g_r[2];
  float* g_i[2];
  float* m_r;
  float* u_r;
  float* m_i;
  float* u_i;
  float* snap_r;
  float* snap_i;
  float* grad_r;        // optional gradient output / all-reduce buffer
  float* grad_i;
  const FitState* st;
  FitConsts k;
  int nfp;
  int nants;
  int mode;             // 0: reduce + update; 1: reduce only -> grad; 2: update only from grad; 4: as 1, before finalize;
                        // 5: LAMB pass 1 (reduce, moments, per-CTA norm partials); 6: LAMB pass 2 (apply with the trust ratios)
  int nf;               // channels that exist (the padding up to nfp takes no part in LAMB's norms)
  LambVars lamb;
  int sum;
  int eval;             // 1: stand-alone gradient evaluation (no step in flight)
  PeerView peers;       // mode 2 with peers.n > 1: gradient = sum over ranks (rank order) of their published partials
  int xpar;             // parity of the step sequence number
  // reduce-only modes, multi-GPU: the LAST CTA to finish also reduces the fused kernel's per-item partial sums into
  // this rank's exchange buffer and publishes the step (saves two launches per iteration)
  unsigned int* tail_counter;
  const double* tail_partials;
  int tail_npartials;
  double* tail_scal;
  unsigned int* tail_flag;
  unsigned int tail_value;
};

// One CTA per (antenna, 64-channel block): 32 lanes along frequency (two adjacent channels each: z / y rows are read
// as float4, gain rows as float2) x GK_EG entry groups.  Entry group g walks the antenna's baselines e0 + g,
// e0 + g + GK_EG, ... with four loads in flight; the GK_EG partial sums are then added in a fixed order through
// shared memory, so the result is deterministic and does not depend on the launch.  (The earlier one-thread-per-
// channel version walked all 2 (Nant - 1) entries serially: 250 us at HERA-350, and just as long on every rank of a
// multi-GPU run whose contiguous shard holds all baselines of a few antennas.)
constexpr int GK_EG = 8;        // entry groups per CTA
constexpr int GK_CH = 64;       // channels per CTA
constexpr int GK_THREADS = 32 * GK_EG;

__global__ void __launch_bounds__(GK_THREADS) gains_kernel(const GainsParams p) {
  __shared__ float4 sh[GK_EG][32];
  const FitState* st = p.st;
  int src;
  if (p.eval) {
    src = st->step & 1;  // stand-alone gradient evaluation at the current parameters
  } else if (p.mode == 4) {
    if (st->step > st->stop_after) return;  // reduce-only pass that runs BEFORE finalize_kernel of this step
    src = st->step & 1;
  } else {
    if (!st->upd_active) return;
    src = (st->step - 1) & 1;
  }
  // grid = (antennas, channel blocks), antennas fastest: the CTAs in flight share one 64-channel block, whose slice of z
  // (nbls x 512 B = 31 MB at HERA-350) stays in L2 between its two reads (once per end of every baseline)
  const int ant = blockIdx.x;
  const int lane = threadIdx.x & 31, eg = threadIdx.x >> 5;
  const int f = blockIdx.y * GK_CH + 2 * lane;
  const bool f_ok = f < p.nfp;  // nfp is a multiple of 16, f is even
  const float* __restrict__ gr = p.g_r[src];
  const float* __restrict__ gi = p.g_i[src];
  const size_t o = (size_t)ant * p.nfp + f;
  float2 acc_r = make_float2(0.f, 0.f), acc_i = make_float2(0.f, 0.f);
  if (p.mode != 2 && p.mode != 6) {
    if (f_ok) {
      const float alpha = st->alpha, beta = st->beta;
      const int e0 = p.ant_ptr[ant], e1 = p.ant_ptr[ant + 1];
      auto one = [&](bool side1, float zx, float zy, float yx, float yy, float pr, float pi, float& ar, float& ai) {
        if (p.sum) {
          zx += alpha * yx + beta * yy;
          zy += alpha * yy - beta * yx;
        }
        if (!side1) {  // this antenna is ant0: conj(z) * g_partner
          ar += zx * pr + zy * pi;
          ai += zx * pi - zy * pr;
        } else {       // this antenna is ant1: z * g_partner
          ar += zx * pr - zy * pi;
          ai += zx * pi + zy * pr;
        }
      };
      auto accumulate = [&](int ent, const float4& z, const float4& y, const float2& pr, const float2& pi) {
        const bool side1 = (ent & 1) != 0;
        one(side1, z.x, z.y, y.x, y.y, pr.x, pi.x, acc_r.x, acc_i.x);
        one(side1, z.z, z.w, y.z, y.w, pr.y, pi.y, acc_r.y, acc_i.y);
      };
      const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
      constexpr int U = 4;  // independent loads in flight per thread; the group's order stays ascending
      int e = e0 + eg;
      for (; e + (U - 1) * GK_EG < e1; e += U * GK_EG) {
        int ent[U], par[U];
        float4 zz[U], yy[U];
        float2 pr[U], pi[U];
#pragma unroll
        for (int k = 0; k < U; ++k) {
          ent[k] = p.ant_ent[e + k * GK_EG];
          par[k] = p.ant_partner[e + k * GK_EG];
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
          const size_t ob = (size_t)(ent[k] >> 1) * p.nfp + f;
          const size_t op = (size_t)par[k] * p.nfp + f;
          zz[k] = *reinterpret_cast<const float4*>(p.z + ob);
          yy[k] = p.sum ? *reinterpret_cast<const float4*>(p.y + ob) : zero4;
          pr[k] = *reinterpret_cast<const float2*>(gr + op);
          pi[k] = *reinterpret_cast<const float2*>(gi + op);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) accumulate(ent[k], zz[k], yy[k], pr[k], pi[k]);
      }
      for (; e < e1; e += GK_EG) {
        const int ent = p.ant_ent[e];
        const size_t ob = (size_t)(ent >> 1) * p.nfp + f;
        const size_t op = (size_t)p.ant_partner[e] * p.nfp + f;
        accumulate(ent, *reinterpret_cast<const float4*>(p.z + ob), p.sum ? *reinterpret_cast<const float4*>(p.y + ob) : zero4,
                   *reinterpret_cast<const float2*>(gr + op), *reinterpret_cast<const float2*>(gi + op));
      }
    }
    sh[eg][lane] = make_float4(acc_r.x, acc_r.y, acc_i.x, acc_i.y);
    __syncthreads();
    const bool writer = eg == 0 && f_ok;
    if (writer) {
      float4 t = sh[0][lane];
#pragma unroll
      for (int g = 1; g < GK_EG; ++g) {  // fixed order
        const float4 u = sh[g][lane];
        t.x += u.x;
        t.y += u.y;
        t.z += u.z;
        t.w += u.w;
      }
      acc_r = make_float2(t.x, t.y);
      acc_i = make_float2(t.z, t.w);
      if (p.grad_r) {
        *reinterpret_cast<float2*>(p.grad_r + o) = acc_r;
        *reinterpret_cast<float2*>(p.grad_i + o) = acc_i;
      }
    }
    if (p.mode == 1 || p.mode == 4) {
      if (p.tail_counter) {  // whole CTA: last-block-done, then partial sums + publication by that block
        __shared__ bool is_last;
        __shared__ double red[3][GK_THREADS / 32];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) is_last = atomicAdd(p.tail_counter, 1u) == gridDim.x * gridDim.y - 1u;
        __syncthreads();
        if (is_last) {
          double a = 0.0, b = 0.0, c = 0.0;
          for (int i = threadIdx.x; i < p.tail_npartials; i += GK_THREADS) {  // fixed order
            a += p.tail_partials[(size_t)i * 4 + 0];
            b += p.tail_partials[(size_t)i * 4 + 1];
            c += p.tail_partials[(size_t)i * 4 + 2];
          }
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, off);
            b += __shfl_xor_sync(0xffffffffu, b, off);
            c += __shfl_xor_sync(0xffffffffu, c, off);
          }
          if (lane == 0) {
            red[0][eg] = a;
            red[1][eg] = b;
            red[2][eg] = c;
          }
          __syncthreads();
          if (threadIdx.x == 0) {
            a = b = c = 0.0;
            for (int w = 0; w < GK_THREADS / 32; ++w) {
              a += red[0][w];
              b += red[1][w];
              c += red[2][w];
            }
            p.tail_scal[0] = a;
            p.tail_scal[1] = b;
            p.tail_scal[2] = c;
            *p.tail_counter = 0u;
            __threadfence_system();
            st_release_sys(p.tail_flag, p.tail_value);
          }
        }
      }
      return;
    }
    if (p.mode == 5 ? eg != 0 : !writer) return;  // LAMB pass 1 keeps warp 0 whole for the norm shuffles
  } else {
    if (eg != 0 || !f_ok) return;
    if (p.mode == 6) {
      // LAMB pass 2 needs no gradient
    } else if (p.peers.n > 1) {  // fused all-reduce: every rank's partial straight from its owner's memory, fixed order
      const size_t ng = (size_t)p.nants * p.nfp;
      float2 tr[CALB2_MAX_RANKS], ti[CALB2_MAX_RANKS];
#pragma unroll
      for (int r = 0; r < CALB2_MAX_RANKS; ++r) {
        if (r < p.peers.n) {
          const float* base = p.peers.grad[r] + (size_t)p.xpar * 2 * ng + o;
          tr[r] = __ldcv(reinterpret_cast<const float2*>(base));
          ti[r] = __ldcv(reinterpret_cast<const float2*>(base + ng));
        }
      }
#pragma unroll
      for (int r = 0; r < CALB2_MAX_RANKS; ++r) {
        if (r < p.peers.n) {
          acc_r.x += tr[r].x;
          acc_r.y += tr[r].y;
          acc_i.x += ti[r].x;
          acc_i.y += ti[r].y;
        }
      }
    } else {
      acc_r = *reinterpret_cast<const float2*>(p.grad_r + o);
      acc_i = *reinterpret_cast<const float2*>(p.grad_i + o);
    }
  }
  const float lr_t = st->lr_t;
  const float a_r[2] = {acc_r.x, acc_r.y}, a_i[2] = {acc_i.x, acc_i.y};
  if (p.mode == 5) {  // LAMB pass 1 (warp 0): moments, then the CTA's share of ||theta||^2 and ||update||^2, fixed order
    double sums[4] = {0.0, 0.0, 0.0, 0.0};
    if (f_ok) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const size_t oc = o + c;
        float mr = p.m_r[oc], ur = p.u_r[oc], mi = p.m_i[oc], ui = p.u_i[oc];
        const float tr = gr[oc], ti = gi[oc];
        const float upr = lamb_moments(p.k, st, tr, a_r[c], mr, ur), upi = lamb_moments(p.k, st, ti, a_i[c], mi, ui);
        p.m_r[oc] = mr;
        p.u_r[oc] = ur;
        p.m_i[oc] = mi;
        p.u_i[oc] = ui;
        if (f + c < p.nf) {
          sums[0] += (double)tr * tr;
          sums[1] += (double)upr * upr;
          sums[2] += (double)ti * ti;
          sums[3] += (double)upi * upi;
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) sums[q] += __shfl_down_sync(0xffffffffu, sums[q], off);
    if (lane == 0) {
      double* dst = p.lamb.gain_partials + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * 4;
      for (int q = 0; q < 4; ++q) dst[q] = sums[q];
    }
    return;
  }
  if (p.mode == 6) {  // LAMB pass 2
    const float rr = p.lamb.ratio[0], ri = p.lamb.ratio[1];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      const size_t oc = o + c;
      const float tr = gr[oc], ti = gi[oc];
      const float nr = tr - rr * p.k.lr * lamb_update(p.k, st, tr, p.m_r[oc], p.u_r[oc]);
      const float ni = ti - ri * p.k.lr * lamb_update(p.k, st, ti, p.m_i[oc], p.u_i[oc]);
      p.g_r[src ^ 1][oc] = nr;
      p.g_i[src ^ 1][oc] = ni;
      if (st->snap && p.snap_r) {
        p.snap_r[oc] = nr;
        p.snap_i[oc] = ni;
      }
    }
    return;
  }
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    const size_t oc = o + c;
    float mr = p.m_r[oc], ur = p.u_r[oc], mi = p.m_i[oc], ui = p.u_i[oc];
    const float nr = opt_step<true, float>(p.k, st, gr[oc], a_r[c], mr, ur, lr_t);
    const float ni = opt_step<true, float>(p.k, st, gi[oc], a_i[c], mi, ui, lr_t);
    p.m_r[oc] = mr;
    p.u_r[oc] = ur;
    p.m_i[oc] = mi;
    p.u_i[oc] = ui;
    p.g_r[src ^ 1][oc] = nr;
    p.g_i[src ^ 1][oc] = ni;
    if (st->snap && p.snap_r) {
      p.snap_r[oc] = nr;
      p.snap_i[oc] = ni;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// freeze_model (calibration.py:598-603): the foreground model never changes, so an iteration only needs the
// elementwise part -- gains, model, residual, chi^2 and z -- over the N_D visibilities; the basis is not touched.
// Grid-stride over (baseline, channel pairs); per-CTA partial sums in a fixed order.
// ------------------------------------------------------------------------------------------------
struct LightParams {
  const float2* vout;   // [nslots][nfp] model visibilities, computed once per fit
  const int* bl_slot;
  const int* bl_ant0;
  const int* bl_ant1;
  const float* d_r;
  const float* d_i;
  const float* w;
  const float* g_r[2];
  const float* g_i[2];
  float2* z;
  float2* y;
  double* partials;     // [gridDim.x][4]
  const FitState* st;
  int nfp;
  long long nelem;      // nbls * nfp
  int sum;
};

__global__ void __launch_bounds__(256) light_kernel(const LightParams p) {
  __shared__ float red[8][4];
  const FitState* st = p.st;
  if (st->step > st->stop_after) return;
  const int gsel = st->step & 1;
  const float* __restrict__ g_r = p.g_r[gsel];
  const float* __restrict__ g_i = p.g_i[gsel];
  float loss_acc = 0.f, sr_acc = 0.f, si_acc = 0.f;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < p.nelem; e += stride) {
    const int b = (int)(e / p.nfp), f = (int)(e % p.nfp);
    const float2 v = p.vout[(size_t)p.bl_slot[b] * p.nfp + f];
    const float dr = p.d_r[e], di = p.d_i[e], w = p.w[e];
    const size_t o0 = (size_t)p.bl_ant0[b] * p.nfp + f, o1 = (size_t)p.bl_ant1[b] * p.nfp + f;
    const float gr0 = g_r[o0], gi0 = g_i[o0], gr1 = g_r[o1], gi1 = g_i[o1];
    const float P = gr0 * gr1 + gi0 * gi1;
    const float Q = gr0 * gi1 - gi0 * gr1;
    const float mr = P * v.x + Q * v.y;
    const float mi = P * v.y - Q * v.x;
    const float rr = dr - mr, ri = di - mi;
    loss_acc += (rr * rr + ri * ri) * w;
    const float er = -2.f * w * rr, ei = -2.f * w * ri;
    p.z[e] = make_float2(er * v.x + ei * v.y, er * v.y - ei * v.x);
    if (p.sum) {
      p.y[e] = make_float2(w * v.x, w * v.y);
      sr_acc += w * mr;
      si_acc += w * mi;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    loss_acc += __shfl_xor_sync(0xffffffffu, loss_acc, off);
    sr_acc += __shfl_xor_sync(0xffffffffu, sr_acc, off);
    si_acc += __shfl_xor_sync(0xffffffffu, si_acc, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = loss_acc;
    red[warp][1] = sr_acc;
    red[warp][2] = si_acc;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += (double)red[w][0];
      b += (double)red[w][1];
      c += (double)red[w][2];
    }
    double* dst = p.partials + (size_t)blockIdx.x * 4;
    dst[0] = a;
    dst[1] = b;
    dst[2] = c;
  }
}

struct CoeffParams {
  const float* dcpart;
  const int* coef_row0;   // [ncoef] global row of the coefficient in its group's first slot
  const int* coef_grp;    // [ncoef]
  const int* grp_nslots;
  const int* grp_slot0;
  const int* grp_coef0;
  const int* slot_row0;
  float* c_r;
  float* c_i;
  float* m_r;
  float* u_r;
  float* m_i;
  float* u_i;
  float* snap_r;
  float* snap_i;
  float* grad_r;
  float* grad_i;
  const FitState* st;
  FitConsts k;
  int ncoef;
  int nq;
  int nplanes;    // planes of dcpart to add for rows >= first_class_row (channel segments of the shared-basis kernel);
                  // the streaming kernel's rows live in plane 0 only -- the other planes hold other layouts' data there
  int first_class_row;
  long long plane;  // floats per plane
  int mode;       // 0: update; 1: gradient only; 3: snapshot copy only (freeze_model);
                  // 5: LAMB pass 1 (moments only); 6: LAMB pass 2 (apply with the variable's trust ratio)
  const long long* var_bounds;  // [nvar + 1] coefficient ranges of the reference's per-chunk variables (LAMB)
  int nvar;
  const float* lamb_ratio;      // [2 + 2 nvar]
};

// the variable (chunk tensor of the reference's fg_r / fg_i lists) a coefficient belongs to
__device__ __forceinline__ int lamb_var_of(const long long* bounds, int nvar, long long c) {
  int lo = 0, hi = nvar - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (bounds[mid] <= c) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(256) coeffs_kernel(const CoeffParams p) {
  const FitState* st = p.st;
  if (p.mode != 1 && !st->upd_active) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.ncoef) return;
  if (p.mode == 3) {
    if (st->snap && p.snap_r) {
      p.snap_r[c] = p.c_r[c];
      p.snap_i[c] = p.c_i[c];
    }
    return;
  }
  if (p.mode == 6) {
    const int v = lamb_var_of(p.var_bounds, p.nvar, c);
    const float rr = p.lamb_ratio[2 + 2 * v], ri = p.lamb_ratio[3 + 2 * v];
    const float tr = p.c_r[c], ti = p.c_i[c];
    const float nr = tr - rr * p.k.lr * lamb_update(p.k, st, tr, p.m_r[c], p.u_r[c]);
    const float ni = ti - ri * p.k.lr * lamb_update(p.k, st, ti, p.m_i[c], p.u_i[c]);
    p.c_r[c] = nr;
    p.c_i[c] = ni;
    if (st->snap && p.snap_r) {
      p.snap_r[c] = nr;
      p.snap_i[c] = ni;
    }
    return;
  }
  const int grp = p.coef_grp[c];
  const int ns = p.grp_nslots[grp];
  float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;
  {
    const float* d = p.dcpart + (size_t)p.coef_row0[c] * p.nq;
    t0 = d[0];
    t1 = d[1];
    if (p.nq == 4) {
      t2 = d[2];
      t3 = d[3];
    }
    const int np = p.coef_row0[c] >= p.first_class_row ? p.nplanes : 1;
    for (int pn = 1; pn < np; ++pn) {  // channel segments, fixed order
      d += p.plane;
      t0 += d[0];
      t1 += d[1];
      if (p.nq == 4) {
        t2 += d[2];
        t3 += d[3];
      }
    }
  }
  if (ns > 1) {
    const int k = c - p.grp_coef0[grp];
    const int s0 = p.grp_slot0[grp];
    for (int s = 1; s < ns; ++s) {
      const float* d = p.dcpart + (size_t)(p.slot_row0[s0 + s] + k) * p.nq;
      t0 += d[0];
      t1 += d[1];
      if (p.nq == 4) {
        t2 += d[2];
        t3 += d[3];
      }
    }
  }
  float gr = t0, gi = t1;
  if (p.nq == 4) {
    const float alpha = st->alpha, beta = st->beta;
    gr = t0 + alpha * t2 - beta * t3;
    gi = t1 + alpha * t3 + beta * t2;
  }
  if (p.grad_r) {
    p.grad_r[c] = gr;
    p.grad_i[c] = gi;
  }
  if (p.mode == 1) return;
  float mr = p.m_r[c], ur = p.u_r[c], mi = p.m_i[c], ui = p.u_i[c];
  if (p.mode == 5) {
    lamb_moments(p.k, st, p.c_r[c], gr, mr, ur);
    lamb_moments(p.k, st, p.c_i[c], gi, mi, ui);
    p.m_r[c] = mr;
    p.u_r[c] = ur;
    p.m_i[c] = mi;
    p.u_i[c] = ui;
    return;
  }
  const float nr = opt_step<false, float>(p.k, st, p.c_r[c], gr, mr, ur, st->lr_t);
  const float ni = opt_step<false, float>(p.k, st, p.c_i[c], gi, mi, ui, st->lr_t);
  p.m_r[c] = mr;
  p.u_r[c] = ur;
  p.m_i[c] = mi;
  p.u_i[c] = ui;
  p.c_r[c] = nr;
  p.c_i[c] = ni;
  if (st->snap && p.snap_r) {
    p.snap_r[c] = nr;
    p.snap_i[c] = ni;
  }
}

// LAMB norms of the coefficient variables: CTA (v, s) sums slice s of variable v's coefficients, update recomputed from the
// moments pass 1 stored.  Thread-strided double sums, then a fixed-order tree: deterministic for a given launch shape.
constexpr int LAMB_SPLIT = 16;
struct LambNormParams {
  const float* c_r;
  const float* c_i;
  const float* m_r;
  const float* u_r;
  const float* m_i;
  const float* u_i;
  const long long* var_bounds;
  double* partials;  // [nvar][LAMB_SPLIT][4]
  const FitState* st;
  FitConsts k;
};
__global__ void __launch_bounds__(256) lamb_coef_norm_kernel(const LambNormParams p) {
  __shared__ double red[8][4];
  const FitState* st = p.st;
  const int v = blockIdx.x, sp = blockIdx.y;
  const long long c0 = p.var_bounds[v], c1 = p.var_bounds[v + 1];
  const long long per = (c1 - c0 + LAMB_SPLIT - 1) / LAMB_SPLIT;
  const long long a = c0 + sp * per, b = a + per < c1 ? a + per : c1;
  double sums[4] = {0.0, 0.0, 0.0, 0.0};
  if (st->upd_active)
    for (long long c = a + threadIdx.x; c < b; c += 256) {
      const float tr = p.c_r[c], ti = p.c_i[c];
      const float ur = lamb_update(p.k, st, tr, p.m_r[c], p.u_r[c]), ui = lamb_update(p.k, st, ti, p.m_i[c], p.u_i[c]);
      sums[0] += (double)tr * tr;
      sums[1] += (double)ur * ur;
      sums[2] += (double)ti * ti;
      sums[3] += (double)ui * ui;
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sums[q] += __shfl_down_sync(0xffffffffu, sums[q], off);
    if (lane == 0) red[warp][q] = sums[q];
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    p.partials[((size_t)v * LAMB_SPLIT + sp) * 4 + threadIdx.x] = t;
  }
}

// trust ratios (tfa LAMB: ||theta|| / ||update|| where both norms are positive, else 1), one warp per variable pair:
// block 0 = the gain tables (g_r, g_i), block 1 + v = coefficient variable v (fg_r[v], fg_i[v])
struct LambRatioParams {
  const double* gain_partials;
  int n_gain_partials;
  const double* coef_partials;
  float* ratio;
};
__global__ void __launch_bounds__(32) lamb_ratio_kernel(const LambRatioParams p) {
  const int lane = threadIdx.x;
  const double* src = blockIdx.x == 0 ? p.gain_partials : p.coef_partials + (size_t)(blockIdx.x - 1) * LAMB_SPLIT * 4;
  const int n = blockIdx.x == 0 ? p.n_gain_partials : LAMB_SPLIT;
  double sums[4] = {0.0, 0.0, 0.0, 0.0};
  for (int i = lane; i < n; i += 32)
    for (int q = 0; q < 4; ++q) sums[q] += src[(size_t)i * 4 + q];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) sums[q] += __shfl_down_sync(0xffffffffu, sums[q], off);
  if (lane == 0) {
    for (int h = 0; h < 2; ++h) {
      const float wn = (float)sqrt(sums[2 * h]), un = (float)sqrt(sums[2 * h + 1]);
      p.ratio[2 * blockIdx.x + h] = (wn > 0.f && un > 0.f) ? wn / un : 1.f;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// setup kernels
// ------------------------------------------------------------------------------------------------
struct RetileJob {
  long long src_off;  // float offset in the staging buffer of the slot's [ncomp][nfreqs] block
  long long dst_off;  // float offset of the item's tile 0
  int ncomp;
  int item_rows;
  int row_in_item;
  int swz_ft;         // 0: streaming-path tiles of the plan's width; 32: shared-basis block, swizzled 32-channel tiles
};

// staging [ncomp][nfreqs] row-major  ->  tiled [tile][item_rows][FT]  (shared-basis blocks: FT = 32, 16-byte chunks
// XOR-swizzled with the row index, see calfit_shared.cuh)
__global__ void retile_kernel(const float* __restrict__ staging, float* __restrict__ A, const RetileJob* jobs,
                              int nfreqs, int ft) {
  const RetileJob jb = jobs[blockIdx.x];
  const long long n = (long long)jb.ncomp * nfreqs;
  const int w = jb.swz_ft ? jb.swz_ft : ft;
  for (long long e = threadIdx.x; e < n; e += blockDim.x) {
    const int k = (int)(e / nfreqs), f = (int)(e % nfreqs);
    const int tile = f / w, row = jb.row_in_item + k;
    int fi = f % w;
    if (jb.swz_ft) fi = (((fi >> 2) ^ (row & 7)) << 2) | (fi & 3);
    A[jb.dst_off + ((long long)tile * jb.item_rows + row) * w + fi] = staging[jb.src_off + e];
  }
}

// staging [ncomp][nfreqs] -> the four tensor-core copies of a class (calfit_tc.cuh): per 32-channel tile
// [b32 hi | b32 lo | std hi | std lo], each [kpt][32]: hi = the value truncated to tf32, lo = the rest; "b32" rows carry the
// 32-byte-base swizzle of the MN-major operand (32-byte pieces ^ (row & 3)), "std" rows the 16-byte one of the K-major operand
struct TcRetileJob {
  long long src_off;
  long long dst_off;
  int ncomp;
  int kpt;
};
__global__ void retile_tc_kernel(const float* __restrict__ staging, float* __restrict__ At, const TcRetileJob* jobs, int nfreqs) {
  const TcRetileJob jb = jobs[blockIdx.x];
  const long long n = (long long)jb.ncomp * nfreqs;
  for (long long e = threadIdx.x; e < n; e += blockDim.x) {
    const int k = (int)(e / nfreqs), f = (int)(e % nfreqs);
    const int tile = f >> 5, fi = f & 31;
    const float x = staging[jb.src_off + e];
    const float hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u), lo = x - hi;
    float* t = At + jb.dst_off + (long long)tile * 4 * jb.kpt * 32 + (long long)k * 32;
    const int p32 = (((fi >> 3) ^ (k & 3)) << 3) | (fi & 7);
    const int p16 = (((fi >> 2) ^ (k & 7)) << 2) | (fi & 3);
    t[p32] = hi;
    t[(long long)jb.kpt * 32 + p32] = lo;
    t[(long long)2 * jb.kpt * 32 + p16] = hi;
    t[(long long)3 * jb.kpt * 32 + p16] = lo;
  }
}

// [n][nfreqs] host layout <-> [n][nfp] padded device layout
__global__ void pad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int nfreqs, int nfp,
                                float fill) {
  const size_t row = blockIdx.x;
  for (int f = threadIdx.x; f < nfp; f += blockDim.x)
    dst[row * nfp + f] = f < nfreqs ? src[row * nfreqs + f] : fill;
}
__global__ void unpad_rows_kernel(const float* __restrict__ src, float* __restrict__ dst, int nfreqs, int nfp) {
  const size_t row = blockIdx.x;
  for (int f = threadIdx.x; f < nfreqs; f += blockDim.x) dst[row * nfreqs + f] = src[row * nfp + f];
}
// per-baseline model: out[b][f] = vout[slot(b)][f]
__global__ void model_gather_kernel(const float2* __restrict__ vout, const int* __restrict__ bl_slot,
                                    float* __restrict__ out_r, float* __restrict__ out_i, int nfreqs, int nfp) {
  const size_t b = blockIdx.x;
  const size_t s = bl_slot[b];
  for (int f = threadIdx.x; f < nfreqs; f += blockDim.x) {
    const float2 v = vout[s * nfp + f];
    out_r[b * nfreqs + f] = v.x;
    out_i[b * nfreqs + f] = v.y;
  }
}

// Deterministic two-stage sum of x[i]*y[i] (prior sums, calibration.py:620-625) and of
// w[i] * (v_r^2 + v_i^2) (SNR weights, calibration.py:1238-1241).
__global__ void __launch_bounds__(256) dot_partial_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                          size_t n, double* __restrict__ out) {
  __shared__ double sh[8];
  double a = 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    a += (double)(y ? x[i] * y[i] : x[i]);
  for (int off = 16; off > 0; off >>= 1) a += __shfl_xor_sync(0xffffffffu, a, off);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += sh[w];
    out[blockIdx.x] = t;
  }
}

__global__ void snr_weight_kernel(float* __restrict__ w, const float2* __restrict__ vout,
                                  const int* __restrict__ bl_slot, int nfp, float scale) {
  const size_t b = blockIdx.x;
  const size_t s = bl_slot[b];
  for (int f = threadIdx.x; f < nfp; f += blockDim.x) {
    const float2 v = vout[s * nfp + f];
    const float nw = (v.x * v.x + v.y * v.y) * w[b * nfp + f];
    w[b * nfp + f] = nw * scale;
  }
}
__global__ void fill_kernel(float* __restrict__ x, size_t n, float value) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = value;
}
__global__ void scale_kernel(float* __restrict__ x, size_t n, float scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] = x[i] / scale;
}

}  // namespace calb2
