/*
 * calamity_b200 -- C ABI of the B200-native gain-and-foreground fit.
 *
 * The reference (aewallwi/calamity) is pure Python over TensorFlow and has no FFI seam of its own; the
 * seam this library is bound at is the call `fit_gains_and_foregrounds(...)` made from
 * `calibrate_and_model_tensor` (/root/reference/calamity/calibration.py:1244-1269) together with the
 * tensor builders around it.  Each entry point below names the reference lines it replaces.
 *
 * Conventions: plain C, every function returns 0 on success and a negative code on failure
 * (`calb2_last_error()` gives the message of the calling thread's last failure); no exception crosses
 * the ABI; the caller owns every host buffer, the library owns all device memory; a plan is bound to
 * one device and must be driven by one host thread at a time.  Floating point buffers are IEEE float32
 * (the reference's default `dtype=np.float32`, calibration.py:974) or float64 (`--precision 64`, calibration.py:1795),
 * as chosen once in calb2_plan_desc.dtype: every `const void*` / `void*` array argument below then points at
 * elements of that type.  All indices are int32.
 *
 * float32 plans run the fused sm_100a kernels.  float64 plans, and float32 plans with a fitting group too large
 * for the fused kernel's staged tile (more than 704 basis vectors), run a generic unfused device path with the same
 * arithmetic (single GPU only).
 *
 * Canonical orderings (they are exactly the reference's flattening of its chunked tensors):
 *   groups     : chunk-major, then group index inside the chunk        (calibration.py:165-170)
 *   slots      : per group, its redundant sub-groups in key order       (calibration.py:173)
 *   baselines  : per slot, its antenna pairs in key order               (calibration.py:175-184)
 *   coefficients: per group, component k = 0..ncomp-1                   (calibration.py:906)
 * so `data_r/data_i/wgts` are the per-chunk [ngrps, nbls, nfreqs] tensors of `tensorize_data`
 * (calibration.py:305-308) concatenated, and `coef_r/coef_i` are the non-padding entries of the
 * per-chunk [nvecs, ngrps, 1, 1] tensors of `tensorize_fg_coeffs`.
 */
#ifndef CALAMITY_B200_H
#define CALAMITY_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct calb2_plan calb2_plan;

enum {
  CALB2_OK = 0,
  CALB2_ERR_ARG = -1,      /* bad argument / inconsistent description */
  CALB2_ERR_CUDA = -2,     /* a CUDA runtime call failed */
  CALB2_ERR_UNSUPPORTED = -3,
  CALB2_ERR_STATE = -4,    /* call order violated (e.g. fit before set_integration) */
  CALB2_ERR_NCCL = -5,
  CALB2_ERR_NONFINITE = -6,
  CALB2_ERR_TIMEOUT = -7   /* peer exchange: a rank did not publish its partial sums in time (fit aborted on every rank) */
};

/* The OPTIMIZERS table of calibration.py:17-27: the eight tf.optimizers.* and tensorflow-addons' LAMB (calibration.py:15, 26;
 * float32 plans on one GPU; its trust ratio is per VARIABLE, see calb2_plan_set_variables). */
enum {
  CALB2_OPT_ADAMAX = 0, CALB2_OPT_ADAM = 1, CALB2_OPT_SGD = 2, CALB2_OPT_RMSPROP = 3, CALB2_OPT_ADAGRAD = 4,
  CALB2_OPT_ADADELTA = 5, CALB2_OPT_NADAM = 6, CALB2_OPT_FTRL = 7, CALB2_OPT_LAMB = 8
};

/* dtype of calibrate_and_model_tensor (calibration.py:974): element type of a plan's floating point buffers. */
enum { CALB2_F32 = 0, CALB2_F64 = 1 };

/* model_regularization of calibration.py:619-661: anything but "sum" is the plain chi-squared. */
enum { CALB2_REG_NONE = 0, CALB2_REG_SUM = 1 };

/* Ragged description of the foreground basis: replaces the dense zero-padded tensors built by
 * tensorize_fg_model_comps_dict (calibration.py:104-190) and the index lists of 577-594. */
typedef struct {
  int32_t device;            /* CUDA device ordinal */
  int32_t nants;             /* g_r.shape[0]  (calibration.py:575) */
  int32_t nfreqs;            /* g_r.shape[1]  (calibration.py:576) */
  int32_t ngroups;           /* fitting groups, all chunks */
  const int32_t* group_ncomp;   /* [ngroups]  basis vectors actually stored for the group (no zero padding) */
  const int32_t* group_nslots;  /* [ngroups]  redundant sub-groups sharing the group's coefficients */
  const int32_t* slot_nbls;     /* [sum nslots] baselines per slot (they share one model visibility) */
  const int32_t* bl_ant0;       /* [nbls_total] corr_inds[..][0] in canonical order */
  const int32_t* bl_ant1;       /* [nbls_total] corr_inds[..][1] */
  int32_t tile_freqs;        /* 0 = choose; else 16, 32 or 64 channels per staged tile */
  int32_t dtype;             /* CALB2_F32 or CALB2_F64: element type of every floating point buffer of this plan */
  /* Shared bases.  modeling.yield_pbl_dpss_model_comps hands the SAME ndarray to every baseline with the same
   * integer-nanosecond delay (modeling.py:293, operator cache 352/371; 120 distinct arrays for the 61 075 groups of
   * HERA-350).  group_class[g] >= 0 names the basis block of group g: groups with equal ids MUST have identical blocks
   * (same ncomp, same values); -1 or a NULL array = private basis.  Single-slot groups of a class with enough members
   * are stored once and fitted by the shared-basis kernel, everything else by the streaming kernel. */
  const int32_t* group_class;   /* [ngroups] or NULL */
  int32_t shared_basis;      /* 0 = automatic (classes of >= 4 groups), 1 = every class, -1 = never (always stream) */
} calb2_plan_desc;

/* Options of fit_gains_and_foregrounds (calibration.py:447-473). */
typedef struct {
  int32_t optimizer;         /* CALB2_OPT_*            (`optimizer`, calibration.py:460, 571) */
  int32_t maxsteps;          /* calibration.py:459, 699 */
  double learning_rate;      /* Keras names, forwarded verbatim by **opt_kwargs (calibration.py:472); real-valued options
                                are doubles and are rounded to the plan's dtype, as Keras casts them to the variable dtype */
  double beta_1;
  double beta_2;
  double epsilon;
  double tol;                /* calibration.py:458, 712 */
  int32_t use_min;           /* calibration.py:457, 702-710 */
  int32_t freeze_model;      /* calibration.py:461, 598-603 */
  int32_t regularization;    /* CALB2_REG_*            (calibration.py:470, 619) */
  double prior_r_sum;        /* sum(sky_model_r * wgts), calibration.py:620-625 */
  double prior_i_sum;
  int32_t n_profile_steps;   /* extra real steps before the warm-up step (calibration.py:681-687) */
  int32_t steps_per_sync;    /* 0 = default; iterations enqueued between host checks of the stop flag */
  int32_t use_graph;         /* 1 = replay a captured CUDA graph of steps_per_sync iterations, -1 = never, 0 = automatic
                                (graphs for small, launch-latency-bound problems on a single GPU) */
  int32_t fuse_tail_update;  /* 1 = coefficient optimizer step inside the fused kernel's tail (only when every group is
                                single-slot and regularization is NONE); measured slower than the split step, default 0 */
  /* further Keras hyper-parameters (same names as the tf.keras.optimizers constructors); ignored by optimizers
   * that do not have them */
  double rho;                         /* RMSprop, Adadelta */
  double momentum;                    /* SGD, RMSprop */
  double initial_accumulator_value;   /* Adagrad, Ftrl */
  double l1_regularization_strength;  /* Ftrl */
  double l2_regularization_strength;  /* Ftrl */
  double learning_rate_power;         /* Ftrl */
  int32_t nesterov;                   /* SGD */
  double weight_decay;                /* LAMB (tfa name: weight_decay / weight_decay_rate) */
} calb2_fit_options;

typedef struct {
  int32_t nsteps_recorded;   /* len(fit_history["loss"]) */
  int32_t nsteps_total;      /* optimizer updates applied = n_profile_steps + 1 + recorded */
  float final_loss;          /* min_loss echoed at calibration.py:734-737 */
  float loop_ms;             /* device time of the whole step loop (CUDA events on the plan's stream) */
  float heavy_ms;            /* device time spent in the fused basis-streaming kernel alone */
  int64_t heavy_launches;
  int64_t kernel_launches;   /* all kernels launched inside the loop */
} calb2_fit_result;

/* Sizes of the plan's internal layout, for roofline accounting. */
typedef struct {
  int64_t n_d;        /* nbls_total * nfreqs */
  int64_t n_a_nz;     /* sum over slots ncomp * nfreqs  (unique non-padding basis elements) */
  int64_t n_a_stored; /* basis floats actually resident (adds row/channel alignment padding) */
  int64_t n_c_nz;     /* sum over groups ncomp */
  int64_t nbls_total;
  int64_t nslots_total;
  int64_t nitems;     /* CTAs of the fused kernel */
  int32_t tile_freqs;
  int32_t rows_per_item_max;
  int64_t device_bytes;
  int32_t generic;    /* 1: the plan runs the generic unfused path (float64, or a group too large for the fused tile) */
  int32_t dtype;
  int64_t n_classes;      /* distinct bases stored once (shared-basis path) */
  int64_t n_class_slots;  /* groups fitted by the shared-basis kernel */
  int64_t n_class_ctas;   /* its CTAs (64 groups of one class each) */
  int64_t n_a_class;      /* basis floats resident for the classes (each distinct basis once) */
  int64_t n_a_class_nz;   /* basis elements those groups count for in n_a_nz (what a per-group copy would hold) */
  int64_t class_fma;      /* fused multiply-adds of the shared-basis kernel per iteration: 4 * n_a_class_nz */
  int64_t n_tc_ctas;      /* CTAs of the tensor-core shape (tcgen05, classes of <= 128 vectors); 0 = off (CALB2_TC=0) */
  int64_t n_tc_slots;     /* groups it fits */
} calb2_plan_info;

const char* calb2_last_error(void);
const char* calb2_version(void);
/* Number of CUDA devices visible to the process.  The reference's only device logic is "make one GPU visible"
 * (calibration.py:1741-1752, 1796-1800); the drop-in driver spreads the independent (polarization, time) integrations of
 * calibration.py:1160-1167 over all of them, one plan per device. */
int calb2_device_count(int32_t* count);
/* Development aid (no counterpart in the reference): with CALB2_GUARD=1 in the environment every device allocation of the
 * library carries 256 pattern bytes on either side; this call counts the live allocations on the current device and the
 * guard zones a kernel has written into.  (compute-sanitizer is not available on the GPU pool this was developed on.) */
int calb2_debug_check_guards(int64_t* nbuffers, int64_t* nviolations);
/* Development aid: the 16-word progress / time-out record of the tensor-core kernel (mapped host memory; readable while
 * another thread is blocked in a CUDA call). */
int calb2_debug_tc_record(uint32_t* out16);
/* Development aid: with CALB2_TC_PROF=<cta> set at plan creation, the SM-clock stamps of that CTA of the last tensor-core
 * launch, [32 tiles][24 slots] (slots: calfit_tc.cuh, tc_stamp). */
int calb2_debug_tc_profile(calb2_plan* plan, int64_t* out, int32_t n);

/* Once per calibrate_and_model_tensor call: mirrors calibration.py:1143-1152. */
int calb2_plan_create(const calb2_plan_desc* desc, calb2_plan** out);
int calb2_plan_destroy(calb2_plan* plan);
int calb2_plan_get_info(const calb2_plan* plan, calb2_plan_info* info);

/* Upload basis vectors for groups [group_first, group_first + ngroups): blocks[i] points at the
 * group's [nslots][ncomp][nfreqs] block (plan dtype) (row k of slot s = fg_model_comps[c][k, g, b, :] for any
 * baseline b of slot s, calibration.py:178-183).  Equal pointers are uploaded once. */
int calb2_plan_set_basis(calb2_plan* plan, int32_t group_first, int32_t ngroups, const void* const* blocks);

/* The reference keeps the foreground coefficients as one tf.Variable PER CHUNK (fg_r[c], fg_i[c], calibration.py:560-567);
 * optimizers whose rule couples the elements of a variable (LAMB's layer-wise trust ratio) need to know the chunks:
 * coef_bounds[0] = 0 < ... < coef_bounds[nvars] = n_c_nz, variable v = coefficients [coef_bounds[v], coef_bounds[v+1]).
 * Never called = one variable holding every coefficient.  The gain tables g_r and g_i are one variable each. */
int calb2_plan_set_variables(calb2_plan* plan, int32_t nvars, const int64_t* coef_bounds);

/* Per integration: mirrors tensorize_data's outputs (calibration.py:1184-1194), [nbls_total][nfreqs]. */
int calb2_set_integration(calb2_plan* plan, const void* data_r, const void* data_i, const void* wgts);
/* tensorize_gains outputs (calibration.py:1213) [nants][nfreqs]; coefficient vectors [n_c_nz]. */
int calb2_set_gains(calb2_plan* plan, const void* g_r, const void* g_i);
int calb2_set_coeffs(calb2_plan* plan, const void* coef_r, const void* coef_i);

/* tensorize_fg_coeffs (calibration.py:828-913) run on the device for both parts at once: least squares
 * of (sky * (wgts != 0)) on each group's basis.  Uses the weights given to calb2_set_integration. */
int calb2_init_coeffs(calb2_plan* plan, const void* sky_r, const void* sky_i);
/* sum(sky_model * wgts) of calibration.py:620-625, reduced on the device. */
int calb2_prior_sums(calb2_plan* plan, const void* sky_r, const void* sky_i, double* prior_r, double* prior_i);
/* use_model_snr_weights block, calibration.py:1235-1242: w <- w (v_r^2 + v_i^2) / sum. */
int calb2_apply_model_snr_weights(calb2_plan* plan);

/* The loop of fit_gains_and_foregrounds, calibration.py:571-738.  loss_history has room for
 * opts->maxsteps floats; the first result->nsteps_recorded are filled (fit_history["loss"]). */
int calb2_fit(calb2_plan* plan, const calb2_fit_options* opts, void* loss_history, calb2_fit_result* result);

/* One evaluation of loss and gradient at the current parameters without updating them (the
 * tape.gradient call of calibration.py:664-666); any output pointer may be NULL. */
int calb2_loss_and_grads(calb2_plan* plan, int32_t regularization, double prior_r_sum, double prior_i_sum,
                         double* loss, void* dg_r, void* dg_i, void* dcoef_r, void* dcoef_i);

/* g_r_opt, g_i_opt, fg_r_opt, fg_i_opt of calibration.py:738. */
int calb2_get_gains(calb2_plan* plan, void* g_r, void* g_i);
int calb2_get_coeffs(calb2_plan* plan, void* coef_r, void* coef_i);
/* Foreground model visibilities sum_k c_k A_k per baseline, [nbls_total][nfreqs]; the caller scatters
 * them into the [nants, nants, nfreqs] cube of yield_fg_model_array (calibration.py:402-444). */
int calb2_get_model(calb2_plan* plan, void* model_r, void* model_i);
int calb2_get_weights(calb2_plan* plan, void* wgts);

/* Multi-GPU: the plan holds one rank's share of the groups; the per-iteration
 * gain gradient, loss and regulariser sums are all-reduced over NCCL.  `nccl_unique_id` is the 128-byte
 * ncclUniqueId created on rank 0 and broadcast by the caller. */
int calb2_comm_init(calb2_plan* plan, const void* nccl_unique_id, int32_t rank, int32_t nranks, const char* nccl_lib);
int calb2_comm_unique_id(void* nccl_unique_id_out, const char* nccl_lib);

/* Same exchange without a collective library in the loop: every rank exports ONE device buffer through cudaIpc
 * (64-byte cudaIpcMemHandle_t), the caller gathers the handles of all ranks (rank order) and hands them to every
 * rank.  Per iteration a rank writes its partial sums / gain-gradient partial into its own buffer and raises a flag;
 * the finalize and gain-update kernels read all ranks' partials over NVLink and add them in rank order, so the
 * reduction is fused into the consumers, deterministic and bit-identical on all ranks.  One node, <= 16 ranks.
 * Every wait on a peer is bounded (20 s, environment CALB2_PEER_TIMEOUT_MS): a rank that died, or was given different
 * maxsteps / tol / n_profile_steps / steps_per_sync, makes calb2_fit return CALB2_ERR_TIMEOUT on the others instead of
 * hanging their GPUs.  calb2_fit starts with one publish / wait round, so consecutive fits cannot overrun each other's
 * buffers.  Lifetime: calb2_comm_peer_close (called by calb2_plan_destroy if the caller did not) runs one last round
 * before it unmaps the peers' buffers, so no rank frees a buffer a peer is still reading; every rank must call it. */
int calb2_comm_peer_export(calb2_plan* plan, void* ipc_handle_out);
int calb2_comm_peer_import(calb2_plan* plan, const void* ipc_handles, int32_t rank, int32_t nranks);
int calb2_comm_peer_close(calb2_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* CALAMITY_B200_H */
