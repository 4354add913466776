// Generic (unfused) device path of the gain-and-foreground fit, templated on the floating-point type.
//
// Used for (1) float64 fits (`dtype=np.float64` / `--precision 64`, calibration.py:974, 1795) and (2) float32 problems
// with a fitting group too large for the fused kernel's staged tile (more than 704 basis vectors).  Same arithmetic
// as the fused path -- model (calibration.py:1587-1605), chi^2 and regulariser (1608-1656), analytic gradient, Keras
// update rules, loop control (693-732) -- but as plain kernels that read the basis twice per iteration (forward and
// backward) from an unpadded row-major copy.  Throughput is not the point of this path; the fused kernels are.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "calfit_kernels.cuh"

namespace calb2 {

template <class T>
struct GState {  // FitState in precision T (same field names: the optimizer templates read them)
  int step, stop_after, nrec, upd_active, snap, any_snap;
  T min_loss, prev_loss, last_loss, alpha, beta, lr_t;
  T aux[4];
  T aux2[2];
  double m_schedule;
  T s_r, s_i;
  double chi2;
};

template <class T>
struct GConsts {  // FitConsts in precision T
  int optimizer;
  T lr, beta1, beta2, eps, rho, momentum, init_acc, l1, l2, lr_power;
  int nesterov;
  int maxsteps;
  double tol;
  int use_min;
  int regularization;
  T prior_r, prior_i;
  int n_skip;
};

template <class T>
struct GenParams {
  // basis, unpadded: slot s owns rows [slot_row0[s], slot_row0[s + 1]) of A[rows][nf]
  const T* A;
  const int* slot_row0;
  const int* row_slot;
  const int* slot_grp;
  const int* grp_ncomp;
  const int* grp_coef0;
  const int* grp_slot0;
  const int* grp_nslots;
  const int* coef_grp;
  const int* slot_bl0;
  const int* bl_ant0;
  const int* bl_ant1;
  const int* bl_slot;
  const int* ant_ptr;
  const int* ant_ent;
  const int* ant_partner;
  // per integration
  const T* d_r;
  const T* d_i;
  const T* w;
  T* g_r[2];
  T* g_i[2];
  T* c_r;
  T* c_i;
  // work arrays
  T* z_a;   // [nbls][nf]
  T* z_b;
  T* y_a;   // regulariser part, [nbls][nf]
  T* y_b;
  T* q;     // [nslots][4][nf]: q_r, q_i, Pw, Qw
  T* dc;    // [rows][4]
  T* vout_r;  // [nslots][nf]
  T* vout_i;
  double* partials;  // [nslots * nfb][4]
  // optimizer slots / snapshots / gradient outputs
  T* gm_r; T* gu_r; T* gm_i; T* gu_i; T* gsnap_r; T* gsnap_i; T* ggrad_r; T* ggrad_i;
  T* cm_r; T* cu_r; T* cm_i; T* cu_i; T* csnap_r; T* csnap_i; T* cgrad_r; T* cgrad_i;
  GState<T>* st;
  T* hist;
  GConsts<T> k;
  int nf, nants, nslots, nfb, ncoef, rows;
  int sum;        // 'sum' regulariser
  int mode;       // forward: 0 fit, 1 init right-hand side (q := data * (w != 0)), 2 model only (vout)
  int eval;       // 1: stand-alone evaluation, state is not advanced
  int grad_only;  // gains / coeffs: write gradients, no update
};

constexpr int GEN_THREADS = 128;

// ---- forward: v = sum_k c_k A_k, model, residual, chi^2, z, dL/dv ----
template <class T>
__global__ void __launch_bounds__(GEN_THREADS) gen_forward_kernel(const GenParams<T> p) {
  __shared__ double red[GEN_THREADS / 32][3];
  const GState<T>* st = p.st;
  if (!p.eval && st->step > st->stop_after) return;
  const int gsel = st->step & 1;
  const T* __restrict__ g_r = p.g_r[gsel];
  const T* __restrict__ g_i = p.g_i[gsel];
  const int s = blockIdx.x, f = blockIdx.y * GEN_THREADS + threadIdx.x;
  const bool ok = f < p.nf;
  const int grp = p.slot_grp[s], ncomp = p.grp_ncomp[grp], c0 = p.grp_coef0[grp];
  const size_t row0 = (size_t)p.slot_row0[s];
  T v_r = 0, v_i = 0;
  if (ok && p.mode != 1) {
    for (int k = 0; k < ncomp; ++k) {
      const T a = p.A[(row0 + k) * p.nf + f];
      v_r += p.c_r[c0 + k] * a;
      v_i += p.c_i[c0 + k] * a;
    }
  }
  double loss = 0.0, sr = 0.0, si = 0.0;
  if (ok) {
    if (p.mode == 2) {
      p.vout_r[(size_t)s * p.nf + f] = v_r;
      p.vout_i[(size_t)s * p.nf + f] = v_i;
    } else {
      T qr = 0, qi = 0, pw = 0, qw = 0;
      for (int b = p.slot_bl0[s]; b < p.slot_bl0[s + 1]; ++b) {
        const size_t o = (size_t)b * p.nf + f;
        const T dr = p.d_r[o], di = p.d_i[o], w = p.w[o];
        if (p.mode == 1) {  // np.isclose(w, 0): |w| <= 1e-8
          const T msk = (m_abs(w) <= (T)1e-8) ? (T)0 : (T)1;
          qr += dr * msk;
          qi += di * msk;
          continue;
        }
        const size_t o0 = (size_t)p.bl_ant0[b] * p.nf + f, o1 = (size_t)p.bl_ant1[b] * p.nf + f;
        const T gr0 = g_r[o0], gi0 = g_i[o0], gr1 = g_r[o1], gi1 = g_i[o1];
        const T P = gr0 * gr1 + gi0 * gi1;
        const T Q = gr0 * gi1 - gi0 * gr1;
        const T mr = P * v_r + Q * v_i;
        const T mi = P * v_i - Q * v_r;
        const T rr = dr - mr, ri = di - mi;
        loss += (double)((rr * rr + ri * ri) * w);
        const T er = (T)-2 * w * rr, ei = (T)-2 * w * ri;
        p.z_a[o] = er * v_r + ei * v_i;
        p.z_b[o] = er * v_i - ei * v_r;
        qr += P * er - Q * ei;
        qi += Q * er + P * ei;
        if (p.sum) {
          p.y_a[o] = w * v_r;
          p.y_b[o] = w * v_i;
          sr += (double)(w * mr);
          si += (double)(w * mi);
          pw += P * w;
          qw += Q * w;
        }
      }
      T* qs = p.q + (size_t)s * 4 * p.nf + f;
      qs[0] = qr;
      qs[p.nf] = qi;
      qs[2 * (size_t)p.nf] = pw;
      qs[3 * (size_t)p.nf] = qw;
    }
  }
  if (p.mode != 0) return;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    loss += __shfl_xor_sync(0xffffffffu, loss, off);
    sr += __shfl_xor_sync(0xffffffffu, sr, off);
    si += __shfl_xor_sync(0xffffffffu, si, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[warp][0] = loss;
    red[warp][1] = sr;
    red[warp][2] = si;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0, c = 0.0;
    for (int w = 0; w < GEN_THREADS / 32; ++w) {
      a += red[w][0];
      b += red[w][1];
      c += red[w][2];
    }
    double* dst = p.partials + ((size_t)s * p.nfb + blockIdx.y) * 4;
    dst[0] = a;
    dst[1] = b;
    dst[2] = c;
  }
}

// ---- backward: dc[row] = sum_f A[row][f] q[slot(row)][f], one CTA per basis row, fixed reduction tree ----
template <class T>
__global__ void __launch_bounds__(GEN_THREADS) gen_backward_kernel(const GenParams<T> p) {
  __shared__ T red[GEN_THREADS / 32][4];
  const GState<T>* st = p.st;
  if (!p.eval && st->step > st->stop_after) return;
  const int row = blockIdx.x, s = p.row_slot[row];
  const T* arow = p.A + (size_t)row * p.nf;
  const T* qs = p.q + (size_t)s * 4 * p.nf;
  const int nq = p.sum ? 4 : 2;
  T acc[4] = {0, 0, 0, 0};
  for (int f = threadIdx.x; f < p.nf; f += GEN_THREADS) {
    const T a = arow[f];
    for (int q = 0; q < nq; ++q) acc[q] += a * qs[(size_t)q * p.nf + f];
  }
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], off);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0)
    for (int q = 0; q < 4; ++q) red[warp][q] = acc[q];
  __syncthreads();
  if (threadIdx.x < 4) {
    T t = 0;
    for (int w = 0; w < GEN_THREADS / 32; ++w) t += red[w][threadIdx.x];
    p.dc[(size_t)row * 4 + threadIdx.x] = t;
  }
}

// ---- finalize: deterministic reduction of the partials + loop control (same logic as finalize_kernel) ----
template <class T>
__global__ void __launch_bounds__(1024, 1) gen_finalize_kernel(const GenParams<T> p, int npartials) {
  __shared__ double sh[3][32];
  GState<T>* st = p.st;
  if (!p.eval && st->step > st->stop_after) {
    if (threadIdx.x == 0) st->upd_active = 0;
    return;
  }
  double a = 0.0, b = 0.0, c = 0.0;
  for (int i = threadIdx.x; i < npartials; i += blockDim.x) {
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
  T loss = (T)a;
  T alpha = 0, beta = 0;
  if (p.k.regularization == 1) {
    const T dr = (T)b - p.k.prior_r, di = (T)c - p.k.prior_i;
    loss = loss + dr * dr + di * di;
    alpha = (T)2 * dr;
    beta = (T)2 * di;
  }
  st->chi2 = a;
  st->s_r = (T)b;
  st->s_i = (T)c;
  st->alpha = alpha;
  st->beta = beta;
  st->last_loss = loss;
  if (p.eval) return;

  const int t = st->step;
  st->lr_t = bias_corrected_lr_t<T>(p.k, t);
  if (p.k.optimizer == OPT_NADAM) nadam_schedule<T>(p.k, st, t);
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

// ---- gains: per (antenna, channel) gradient over the antenna's baselines (CSR order) + optimizer step ----
template <class T>
__global__ void __launch_bounds__(GEN_THREADS) gen_gains_kernel(const GenParams<T> p) {
  const GState<T>* st = p.st;
  int src;
  if (p.eval) {
    src = st->step & 1;
  } else {
    if (!st->upd_active) return;
    src = (st->step - 1) & 1;
  }
  const int ant = blockIdx.y, f = blockIdx.x * GEN_THREADS + threadIdx.x;
  if (f >= p.nf) return;
  const T* __restrict__ gr = p.g_r[src];
  const T* __restrict__ gi = p.g_i[src];
  const T alpha = st->alpha, beta = st->beta;
  T ar = 0, ai = 0;
  for (int e = p.ant_ptr[ant]; e < p.ant_ptr[ant + 1]; ++e) {
    const int ent = p.ant_ent[e];
    const size_t ob = (size_t)(ent >> 1) * p.nf + f, op = (size_t)p.ant_partner[e] * p.nf + f;
    T zx = p.z_a[ob], zy = p.z_b[ob];
    if (p.sum) {
      const T yx = p.y_a[ob], yy = p.y_b[ob];
      zx += alpha * yx + beta * yy;
      zy += alpha * yy - beta * yx;
    }
    const T pr = gr[op], pi = gi[op];
    if (!(ent & 1)) {  // this antenna is ant0: conj(z) * g_partner
      ar += zx * pr + zy * pi;
      ai += zx * pi - zy * pr;
    } else {           // this antenna is ant1: z * g_partner
      ar += zx * pr - zy * pi;
      ai += zx * pi + zy * pr;
    }
  }
  const size_t o = (size_t)ant * p.nf + f;
  if (p.grad_only) {
    p.ggrad_r[o] = ar;
    p.ggrad_i[o] = ai;
    return;
  }
  T mr = p.gm_r[o], ur = p.gu_r[o], mi = p.gm_i[o], ui = p.gu_i[o];
  const T nr = opt_step<true, T>(p.k, st, gr[o], ar, mr, ur, st->lr_t);
  const T ni = opt_step<true, T>(p.k, st, gi[o], ai, mi, ui, st->lr_t);
  p.gm_r[o] = mr;
  p.gu_r[o] = ur;
  p.gm_i[o] = mi;
  p.gu_i[o] = ui;
  p.g_r[src ^ 1][o] = nr;
  p.g_i[src ^ 1][o] = ni;
  if (st->snap && p.gsnap_r) {
    p.gsnap_r[o] = nr;
    p.gsnap_i[o] = ni;
  }
}

// ---- coefficients: sum the backward contractions of the group's slots, combine the regulariser terms, step ----
template <class T>
__global__ void __launch_bounds__(256) gen_coeffs_kernel(const GenParams<T> p, int snapshot_only) {
  const GState<T>* st = p.st;
  if (!p.grad_only && !st->upd_active) return;
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= p.ncoef) return;
  if (snapshot_only) {
    if (st->snap && p.csnap_r) {
      p.csnap_r[c] = p.c_r[c];
      p.csnap_i[c] = p.c_i[c];
    }
    return;
  }
  const int grp = p.coef_grp[c], k = c - p.grp_coef0[grp], s0 = p.grp_slot0[grp];
  T t0 = 0, t1 = 0, t2 = 0, t3 = 0;
  for (int s = 0; s < p.grp_nslots[grp]; ++s) {
    const T* d = p.dc + (size_t)(p.slot_row0[s0 + s] + k) * 4;
    t0 += d[0];
    t1 += d[1];
    t2 += d[2];
    t3 += d[3];
  }
  T gr = t0, gi = t1;
  if (p.sum) {
    const T alpha = st->alpha, beta = st->beta;
    gr = t0 + alpha * t2 - beta * t3;
    gi = t1 + alpha * t3 + beta * t2;
  }
  if (p.grad_only) {
    p.cgrad_r[c] = gr;
    p.cgrad_i[c] = gi;
    return;
  }
  T mr = p.cm_r[c], ur = p.cu_r[c], mi = p.cm_i[c], ui = p.cu_i[c];
  const T nr = opt_step<false, T>(p.k, st, p.c_r[c], gr, mr, ur, st->lr_t);
  const T ni = opt_step<false, T>(p.k, st, p.c_i[c], gi, mi, ui, st->lr_t);
  p.cm_r[c] = mr;
  p.cu_r[c] = ur;
  p.cm_i[c] = mi;
  p.cu_i[c] = ui;
  p.c_r[c] = nr;
  p.c_i[c] = ni;
  if (st->snap && p.csnap_r) {
    p.csnap_r[c] = nr;
    p.csnap_i[c] = ni;
  }
}

// ---- setup: Gram matrix of one group (float64), see gram_kernel in calfit_setup.cuh ----
template <class T>
__global__ void __launch_bounds__(256) gen_gram_kernel(const T* __restrict__ A, const GramJob* __restrict__ jobs,
                                                       const int* __restrict__ slot_row0, const int* __restrict__ slot_bl0,
                                                       double* __restrict__ gram, int nf) {
  const GramJob jb = jobs[blockIdx.x];
  const int n = jb.n;
  double* Gm = gram + jb.gram_off;
  for (int pidx = threadIdx.x; pidx < n * n; pidx += blockDim.x) {
    const int k = pidx / n, k2 = pidx % n;
    if (k2 > k) continue;
    double acc = 0.0;
    for (int s = 0; s < jb.nslots; ++s) {
      const int slot = jb.slot0 + s;
      const double wgt = (double)(slot_bl0[slot + 1] - slot_bl0[slot]);
      const T* ra = A + (size_t)(slot_row0[slot] + k) * nf;
      const T* rb = A + (size_t)(slot_row0[slot] + k2) * nf;
      double part = 0.0;
      for (int f = 0; f < nf; ++f) part += (double)ra[f] * (double)rb[f];
      acc += wgt * part;
    }
    Gm[(size_t)k * n + k2] = acc;
    Gm[(size_t)k2 * n + k] = acc;
  }
}

// Cholesky solve of the group's two right-hand sides (same algorithm as chol_solve_kernel, output type T)
template <class T>
__global__ void __launch_bounds__(256) gen_chol_solve_kernel(const GramJob* __restrict__ jobs, double* __restrict__ gram,
                                                             T* __restrict__ rhs_r, T* __restrict__ rhs_i) {
  const GramJob jb = jobs[blockIdx.x];
  const int n = jb.n;
  if (n == 0) return;
  double* L = gram + jb.gram_off;
  for (int j = 0; j < n; ++j) {
    __syncthreads();
    const double d = sqrt(L[(long long)j * n + j]);
    __syncthreads();
    if (threadIdx.x == 0) L[(long long)j * n + j] = d;
    for (int i = j + 1 + threadIdx.x; i < n; i += 256) L[(long long)i * n + j] /= d;
    __syncthreads();
    const int m = n - j - 1;
    for (int e = threadIdx.x; e < m * m; e += 256) {
      const int i = j + 1 + e / m, k = j + 1 + e % m;
      if (k <= i) L[(long long)i * n + k] -= L[(long long)i * n + j] * L[(long long)k * n + j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 2) {
    T* rhs = (threadIdx.x == 0 ? rhs_r : rhs_i) + jb.coef0;
    double* ycol = L + (long long)n * n + (long long)threadIdx.x * n;  // scratch appended after the matrix
    for (int i = 0; i < n; ++i) {
      double v = (double)rhs[i];
      for (int k = 0; k < i; ++k) v -= L[(long long)i * n + k] * ycol[k];
      ycol[i] = v / L[(long long)i * n + i];
    }
    for (int i = n - 1; i >= 0; --i) {
      double v = ycol[i];
      for (int k = i + 1; k < n; ++k) v -= L[(long long)k * n + i] * ycol[k];
      ycol[i] = v / L[(long long)i * n + i];
    }
    for (int i = 0; i < n; ++i) rhs[i] = (T)ycol[i];
  }
}

// ---- small helpers ----
template <class T>
__global__ void gen_fill_kernel(T* __restrict__ x, size_t n, T value) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = value;
}
template <class T>
__global__ void __launch_bounds__(256) gen_dot_partial_kernel(const T* __restrict__ x, const T* __restrict__ y, size_t n,
                                                              double* __restrict__ out) {
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
// w <- w * |v|^2 (calibration.py:1238-1241), v per slot
template <class T>
__global__ void gen_snr_weight_kernel(T* __restrict__ w, const T* __restrict__ v_r, const T* __restrict__ v_i,
                                      const int* __restrict__ bl_slot, int nf) {
  const size_t b = blockIdx.x, s = bl_slot[b];
  for (int f = threadIdx.x; f < nf; f += blockDim.x) {
    const T vr = v_r[s * nf + f], vi = v_i[s * nf + f];
    w[b * nf + f] = (vr * vr + vi * vi) * w[b * nf + f];
  }
}
template <class T>
__global__ void gen_div_kernel(T* __restrict__ x, size_t n, T scale) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] = x[i] / scale;
}
template <class T>
__global__ void gen_model_gather_kernel(const T* __restrict__ v_r, const T* __restrict__ v_i, const int* __restrict__ bl_slot,
                                        T* __restrict__ out_r, T* __restrict__ out_i, int nf) {
  const size_t b = blockIdx.x, s = bl_slot[b];
  for (int f = threadIdx.x; f < nf; f += blockDim.x) {
    out_r[b * nf + f] = v_r[s * nf + f];
    out_i[b * nf + f] = v_i[s * nf + f];
  }
}

}  // namespace calb2
