// Host side of the generic (unfused, precision-templated) path: see calfit_generic.cuh.  Included by calfit_api.cu after
// the definition of calb2_plan; every C-ABI entry point forwards here when the plan carries a GenericBase.
#pragma once

namespace calb2 {

struct GenericBase {
  virtual ~GenericBase() {}
  virtual int set_basis(int g0, int ng, const void* const* blocks) = 0;
  virtual int set_integration(const void* d_r, const void* d_i, const void* w) = 0;
  virtual int set_gains(const void* g_r, const void* g_i) = 0;
  virtual int set_coeffs(const void* c_r, const void* c_i) = 0;
  virtual int get_gains(void* g_r, void* g_i) = 0;
  virtual int get_coeffs(void* c_r, void* c_i) = 0;
  virtual int get_model(void* m_r, void* m_i) = 0;
  virtual int get_weights(void* w) = 0;
  virtual int init_coeffs(const void* sky_r, const void* sky_i) = 0;
  virtual int prior_sums(const void* sky_r, const void* sky_i, double* pr, double* pi) = 0;
  virtual int apply_snr_weights() = 0;
  virtual int loss_and_grads(int reg, double prior_r, double prior_i, double* loss, void* dg_r, void* dg_i, void* dc_r, void* dc_i) = 0;
  virtual int fit(const calb2_fit_options* o, void* hist, calb2_fit_result* res) = 0;
  virtual size_t device_bytes() const = 0;
};

template <class T>
struct GenericPlan : GenericBase {
  calb2_plan* pl;
  int nf, nants, nslots, ngroups, nfb;
  long long nbls, ncoef, rows;
  std::vector<int> h_slot_row0, h_row_slot;
  DevBuf<int> slot_row0, row_slot, slot_grp;
  DevBuf<T> A, d_r, d_i, w, z_a, z_b, y_a, y_b, q, dc, vout_r, vout_i, g_r[2], g_i[2], c_r, c_i, gm_r, gu_r, gm_i, gu_i, gsnap_r,
      gsnap_i, ggrad_r, ggrad_i, cm_r, cu_r, cm_i, cu_i, csnap_r, csnap_i, cgrad_r, cgrad_i, hist, sky_r, sky_i;
  DevBuf<double> partials, red_d;
  DevBuf<GState<T>> state, state_eval;
  GState<T>* h_state = nullptr;
  int cur_buf = 0;
  bool have_data = false, have_gains = false, have_coeffs = false;
  size_t bytes = 0;

  template <class U>
  int alloc(DevBuf<U>& b, size_t n, bool zero = true) {
    CU(b.alloc(n));
    bytes += b.bytes();
    if (zero && n) CU(cudaMemsetAsync(b.p, 0, b.bytes(), pl->stream));  // stream-ordered (the plan's streams are non-blocking)
    return 0;
  }

  int init(calb2_plan* plan) {
    pl = plan;
    nf = pl->nf;
    nants = pl->nants;
    nslots = (int)pl->nslots;
    ngroups = pl->ngroups;
    nbls = pl->nbls;
    ncoef = pl->ncoef;
    nfb = (nf + GEN_THREADS - 1) / GEN_THREADS;
    h_slot_row0.resize(nslots + 1);
    long long r = 0;
    for (int s = 0; s < nslots; ++s) {
      h_slot_row0[s] = (int)r;
      const int nc = pl->grp_ncomp[pl->slot_grp[s]];
      for (int k = 0; k < nc; ++k) h_row_slot.push_back(s);
      r += nc;
    }
    h_slot_row0[nslots] = (int)r;
    rows = r;
    if (rows > INT_MAX / 4) return fail(CALB2_ERR_UNSUPPORTED, "too many basis rows (%lld)", rows);
#define GTRY(x) \
  if (int rc__ = (x)) return rc__;
    GTRY(upload(slot_row0, h_slot_row0, pl));
    GTRY(upload(row_slot, h_row_slot, pl));
    GTRY(upload(slot_grp, pl->slot_grp, pl));
    const size_t nd = (size_t)nbls * nf, ng = (size_t)nants * nf;
    GTRY(alloc(A, (size_t)rows * nf));
    GTRY(alloc(d_r, nd));
    GTRY(alloc(d_i, nd));
    GTRY(alloc(w, nd));
    GTRY(alloc(z_a, nd));
    GTRY(alloc(z_b, nd));
    GTRY(alloc(q, (size_t)nslots * 4 * nf));
    GTRY(alloc(dc, (size_t)std::max<long long>(rows, 1) * 4));
    for (int b = 0; b < 2; ++b) {
      GTRY(alloc(g_r[b], ng));
      GTRY(alloc(g_i[b], ng));
    }
    GTRY(alloc(gm_r, ng));
    GTRY(alloc(gu_r, ng));
    GTRY(alloc(gm_i, ng));
    GTRY(alloc(gu_i, ng));
    GTRY(alloc(ggrad_r, ng));
    GTRY(alloc(ggrad_i, ng));
    GTRY(alloc(c_r, (size_t)ncoef));
    GTRY(alloc(c_i, (size_t)ncoef));
    GTRY(alloc(cm_r, (size_t)ncoef));
    GTRY(alloc(cu_r, (size_t)ncoef));
    GTRY(alloc(cm_i, (size_t)ncoef));
    GTRY(alloc(cu_i, (size_t)ncoef));
    GTRY(alloc(cgrad_r, (size_t)ncoef));
    GTRY(alloc(cgrad_i, (size_t)ncoef));
    GTRY(alloc(partials, (size_t)nslots * nfb * 4));
    GTRY(alloc(red_d, 1024));
    GTRY(alloc(state, 1));
    GTRY(alloc(state_eval, 1));
    if (cudaMallocHost(&h_state, sizeof(GState<T>)) != cudaSuccess) return fail(CALB2_ERR_CUDA, "cudaMallocHost(state) failed");
    return 0;
  }

  ~GenericPlan() override {
    DevBuf<T>* fb[] = {&A, &d_r, &d_i, &w, &z_a, &z_b, &y_a, &y_b, &q, &dc, &vout_r, &vout_i, &g_r[0], &g_r[1], &g_i[0], &g_i[1], &c_r,
                       &c_i, &gm_r, &gu_r, &gm_i, &gu_i, &gsnap_r, &gsnap_i, &ggrad_r, &ggrad_i, &cm_r, &cu_r, &cm_i, &cu_i, &csnap_r,
                       &csnap_i, &cgrad_r, &cgrad_i, &hist, &sky_r, &sky_i};
    for (auto* b : fb) b->release();
    slot_row0.release();
    row_slot.release();
    slot_grp.release();
    partials.release();
    red_d.release();
    state.release();
    state_eval.release();
    if (h_state) cudaFreeHost(h_state);
  }

  size_t device_bytes() const override { return bytes; }

  GenParams<T> params(GState<T>* st, bool sum, int mode, int eval) {
    GenParams<T> p{};
    p.A = A.p;
    p.slot_row0 = slot_row0.p;
    p.row_slot = row_slot.p;
    p.slot_grp = slot_grp.p;
    p.grp_ncomp = pl->d_grp_ncomp.p;
    p.grp_coef0 = pl->d_grp_coef0.p;
    p.grp_slot0 = pl->d_grp_slot0.p;
    p.grp_nslots = pl->d_grp_nslots.p;
    p.coef_grp = pl->coef_grp.p;
    p.slot_bl0 = pl->d_slot_bl0.p;
    p.bl_ant0 = pl->d_bl_ant0.p;
    p.bl_ant1 = pl->d_bl_ant1.p;
    p.bl_slot = pl->d_bl_slot.p;
    p.ant_ptr = pl->ant_ptr.p;
    p.ant_ent = pl->ant_ent.p;
    p.ant_partner = pl->ant_partner.p;
    p.d_r = d_r.p;
    p.d_i = d_i.p;
    p.w = w.p;
    for (int b = 0; b < 2; ++b) {
      p.g_r[b] = g_r[b].p;
      p.g_i[b] = g_i[b].p;
    }
    p.c_r = c_r.p;
    p.c_i = c_i.p;
    p.z_a = z_a.p;
    p.z_b = z_b.p;
    p.y_a = y_a.p;
    p.y_b = y_b.p;
    p.q = q.p;
    p.dc = dc.p;
    p.vout_r = vout_r.p;
    p.vout_i = vout_i.p;
    p.partials = partials.p;
    p.gm_r = gm_r.p; p.gu_r = gu_r.p; p.gm_i = gm_i.p; p.gu_i = gu_i.p;
    p.gsnap_r = gsnap_r.p; p.gsnap_i = gsnap_i.p; p.ggrad_r = ggrad_r.p; p.ggrad_i = ggrad_i.p;
    p.cm_r = cm_r.p; p.cu_r = cu_r.p; p.cm_i = cm_i.p; p.cu_i = cu_i.p;
    p.csnap_r = csnap_r.p; p.csnap_i = csnap_i.p; p.cgrad_r = cgrad_r.p; p.cgrad_i = cgrad_i.p;
    p.st = st;
    p.hist = hist.p;
    p.nf = nf;
    p.nants = nants;
    p.nslots = nslots;
    p.nfb = nfb;
    p.ncoef = (int)ncoef;
    p.rows = (int)rows;
    p.sum = sum ? 1 : 0;
    p.mode = mode;
    p.eval = eval;
    p.grad_only = 0;
    return p;
  }

  // ---- uploads / downloads (no padding in this path) ----
  int set_basis(int g0, int ng, const void* const* blocks) override {
    for (int gi = 0; gi < ng; ++gi) {
      const int g = g0 + gi, nc = pl->grp_ncomp[g], nsl = pl->grp_nslots[g];
      if ((size_t)nc * nsl == 0) continue;
      if (!blocks[gi]) return fail(CALB2_ERR_ARG, "group %d: null basis block", g);
      const T* blk = static_cast<const T*>(blocks[gi]);  // [nslots][ncomp][nf]
      for (int s = 0; s < nsl; ++s) {
        const int slot = pl->grp_slot0[g] + s;
        CU(cudaMemcpy(A.p + (size_t)h_slot_row0[slot] * nf, blk + (size_t)s * nc * nf, (size_t)nc * nf * sizeof(T),
                      cudaMemcpyHostToDevice));
      }
    }
    return 0;
  }
  int put(DevBuf<T>& dst, const void* src, size_t n) {
    CU(cudaMemcpy(dst.p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return 0;
  }
  int get(void* dst, const T* src, size_t n) {
    CU(cudaStreamSynchronize(pl->stream));
    CU(cudaMemcpy(dst, src, n * sizeof(T), cudaMemcpyDeviceToHost));
    return 0;
  }
  int set_integration(const void* a, const void* b, const void* c) override {
    const size_t nd = (size_t)nbls * nf;
    GTRY(put(d_r, a, nd));
    GTRY(put(d_i, b, nd));
    GTRY(put(w, c, nd));
    have_data = true;
    return 0;
  }
  int set_gains(const void* a, const void* b) override {
    cur_buf = 0;
    GTRY(put(g_r[0], a, (size_t)nants * nf));
    GTRY(put(g_i[0], b, (size_t)nants * nf));
    have_gains = true;
    return 0;
  }
  int set_coeffs(const void* a, const void* b) override {
    GTRY(put(c_r, a, (size_t)ncoef));
    GTRY(put(c_i, b, (size_t)ncoef));
    have_coeffs = true;
    return 0;
  }
  int get_gains(void* a, void* b) override {
    GTRY(get(a, g_r[cur_buf].p, (size_t)nants * nf));
    return get(b, g_i[cur_buf].p, (size_t)nants * nf);
  }
  int get_coeffs(void* a, void* b) override {
    GTRY(get(a, c_r.p, (size_t)ncoef));
    return get(b, c_i.p, (size_t)ncoef);
  }
  int get_weights(void* a) override { return get(a, w.p, (size_t)nbls * nf); }

  int set_eval_state() {
    GState<T> s{};
    s.step = cur_buf;
    s.stop_after = INT_MAX;
    CU(cudaMemcpyAsync(state_eval.p, &s, sizeof(s), cudaMemcpyHostToDevice, pl->stream));
    return 0;
  }
  int ensure_vout() {
    if (!vout_r.p) {
      GTRY(alloc(vout_r, (size_t)nslots * nf));
      GTRY(alloc(vout_i, (size_t)nslots * nf));
    }
    return 0;
  }
  int ensure_sum() {
    if (!y_a.p) {
      GTRY(alloc(y_a, (size_t)nbls * nf));
      GTRY(alloc(y_b, (size_t)nbls * nf));
    }
    return 0;
  }
  int forward_model_only() {
    GTRY(ensure_vout());
    GTRY(set_eval_state());
    GenParams<T> p = params(state_eval.p, false, 2, 1);
    gen_forward_kernel<T><<<dim3(nslots, nfb), GEN_THREADS, 0, pl->stream>>>(p);
    CU(cudaGetLastError());
    return 0;
  }
  int get_model(void* m_r, void* m_i) override {
    if (!have_coeffs) return fail(CALB2_ERR_STATE, "coefficients must be set first");
    GTRY(forward_model_only());
    DevBuf<T> out;
    CU(out.alloc((size_t)2 * nbls * nf));
    gen_model_gather_kernel<T><<<(unsigned)nbls, 128, 0, pl->stream>>>(vout_r.p, vout_i.p, pl->d_bl_slot.p, out.p, out.p + (size_t)nbls * nf, nf);
    CU(cudaGetLastError());
    int rc = get(m_r, out.p, (size_t)nbls * nf);
    if (!rc) rc = get(m_i, out.p + (size_t)nbls * nf, (size_t)nbls * nf);
    out.release();
    return rc;
  }
  int device_sum(const T* x, const T* y, size_t n, double* out) {
    const int nb = 1024;
    gen_dot_partial_kernel<T><<<nb, 256, 0, pl->stream>>>(x, y, n, red_d.p);
    CU(cudaGetLastError());
    std::vector<double> h(nb);
    CU(cudaMemcpyAsync(h.data(), red_d.p, nb * sizeof(double), cudaMemcpyDeviceToHost, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    double t = 0.0;
    for (double v : h) t += v;
    *out = t;
    return 0;
  }
  int ensure_sky() {
    const size_t nd = (size_t)nbls * nf;
    if (sky_r.n < nd) {
      GTRY(alloc(sky_r, nd));
      GTRY(alloc(sky_i, nd));
    }
    return 0;
  }
  int prior_sums(const void* s_r, const void* s_i, double* pr, double* pi) override {
    if (!have_data) return fail(CALB2_ERR_STATE, "set_integration first (weights)");
    GTRY(ensure_sky());
    const size_t nd = (size_t)nbls * nf;
    GTRY(put(sky_r, s_r, nd));
    GTRY(put(sky_i, s_i, nd));
    GTRY(device_sum(sky_r.p, w.p, nd, pr));
    return device_sum(sky_i.p, w.p, nd, pi);
  }
  int apply_snr_weights() override {
    if (!have_data || !have_coeffs) return fail(CALB2_ERR_STATE, "integration and coefficients must be set first");
    GTRY(forward_model_only());
    gen_snr_weight_kernel<T><<<(unsigned)nbls, 128, 0, pl->stream>>>(w.p, vout_r.p, vout_i.p, pl->d_bl_slot.p, nf);
    CU(cudaGetLastError());
    double t = 0.0;
    const size_t nd = (size_t)nbls * nf;
    GTRY(device_sum(w.p, nullptr, nd, &t));
    gen_div_kernel<T><<<1024, 256, 0, pl->stream>>>(w.p, nd, (T)t);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(pl->stream));
    return 0;
  }

  // tensorize_fg_coeffs (calibration.py:828-913): normal equations per group, Cholesky in float64
  int init_coeffs(const void* s_r, const void* s_i) override {
    if (!have_data) return fail(CALB2_ERR_STATE, "set_integration first (weights)");
    GTRY(ensure_sky());
    const size_t nd = (size_t)nbls * nf;
    GTRY(put(sky_r, s_r, nd));
    GTRY(put(sky_i, s_i, nd));
    GTRY(set_eval_state());
    GenParams<T> p = params(state_eval.p, false, 1, 1);
    p.d_r = sky_r.p;
    p.d_i = sky_i.p;
    gen_forward_kernel<T><<<dim3(nslots, nfb), GEN_THREADS, 0, pl->stream>>>(p);
    CU(cudaGetLastError());
    if (rows) gen_backward_kernel<T><<<(unsigned)rows, GEN_THREADS, 0, pl->stream>>>(p);
    CU(cudaGetLastError());
    p.grad_only = 1;
    gen_coeffs_kernel<T><<<(unsigned)((ncoef + 255) / 256), 256, 0, pl->stream>>>(p, 0);
    CU(cudaGetLastError());
    std::vector<GramJob> jobs;
    size_t used = 0;
    for (int g = 0; g < ngroups; ++g) {
      const int n = pl->grp_ncomp[g];
      if (n == 0) continue;
      GramJob jb{};
      jb.gram_off = (long long)used;
      jb.grp = g;
      jb.n = n;
      jb.slot0 = pl->grp_slot0[g];
      jb.nslots = pl->grp_nslots[g];
      jb.coef0 = pl->grp_coef0[g];
      jobs.push_back(jb);
      used += (size_t)n * n + 2 * (size_t)n;
    }
    if (!jobs.empty()) {
      DevBuf<double> gram;
      DevBuf<GramJob> djobs;
      CU(gram.alloc(used));
      CU(djobs.alloc(jobs.size()));
      CU(cudaMemcpyAsync(djobs.p, jobs.data(), jobs.size() * sizeof(GramJob), cudaMemcpyHostToDevice, pl->stream));
      gen_gram_kernel<T><<<(unsigned)jobs.size(), 256, 0, pl->stream>>>(A.p, djobs.p, slot_row0.p, pl->d_slot_bl0.p, gram.p, nf);
      CU(cudaGetLastError());
      gen_chol_solve_kernel<T><<<(unsigned)jobs.size(), 256, 0, pl->stream>>>(djobs.p, gram.p, cgrad_r.p, cgrad_i.p);
      CU(cudaGetLastError());
      CU(cudaStreamSynchronize(pl->stream));
      gram.release();
      djobs.release();
    }
    CU(cudaMemcpyAsync(c_r.p, cgrad_r.p, ncoef * sizeof(T), cudaMemcpyDeviceToDevice, pl->stream));
    CU(cudaMemcpyAsync(c_i.p, cgrad_i.p, ncoef * sizeof(T), cudaMemcpyDeviceToDevice, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    have_coeffs = true;
    return 0;
  }

  static GConsts<T> consts_from(const calb2_fit_options* o) {
    GConsts<T> k{};
    k.optimizer = o->optimizer;
    k.lr = (T)o->learning_rate;
    k.beta1 = (T)o->beta_1;
    k.beta2 = (T)o->beta_2;
    k.eps = (T)o->epsilon;
    k.rho = (T)o->rho;
    k.momentum = (T)o->momentum;
    k.init_acc = (T)o->initial_accumulator_value;
    k.l1 = (T)o->l1_regularization_strength;
    k.l2 = (T)o->l2_regularization_strength;
    k.lr_power = (T)o->learning_rate_power;
    k.nesterov = o->nesterov;
    k.maxsteps = o->maxsteps;
    k.tol = o->tol;
    k.use_min = o->use_min;
    k.regularization = o->regularization;
    k.prior_r = (T)o->prior_r_sum;
    k.prior_i = (T)o->prior_i_sum;
    k.n_skip = o->n_profile_steps + 1;
    return k;
  }

  int loss_and_grads(int reg, double prior_r, double prior_i, double* loss, void* dg_r, void* dg_i, void* dc_r, void* dc_i) override {
    if (!have_data || !have_gains || !have_coeffs) return fail(CALB2_ERR_STATE, "integration, gains and coefficients must be set first");
    const bool sum = reg == CALB2_REG_SUM;
    if (sum) GTRY(ensure_sum());
    GTRY(set_eval_state());
    GenParams<T> p = params(state_eval.p, sum, 0, 1);
    p.k.regularization = reg;
    p.k.prior_r = (T)prior_r;
    p.k.prior_i = (T)prior_i;
    gen_forward_kernel<T><<<dim3(nslots, nfb), GEN_THREADS, 0, pl->stream>>>(p);
    CU(cudaGetLastError());
    if (rows) gen_backward_kernel<T><<<(unsigned)rows, GEN_THREADS, 0, pl->stream>>>(p);
    CU(cudaGetLastError());
    gen_finalize_kernel<T><<<1, 1024, 0, pl->stream>>>(p, nslots * nfb);
    CU(cudaGetLastError());
    p.grad_only = 1;
    gen_gains_kernel<T><<<dim3(nfb, nants), GEN_THREADS, 0, pl->stream>>>(p);
    CU(cudaGetLastError());
    gen_coeffs_kernel<T><<<(unsigned)((ncoef + 255) / 256), 256, 0, pl->stream>>>(p, 0);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(h_state, state_eval.p, sizeof(GState<T>), cudaMemcpyDeviceToHost, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    if (loss) *loss = (double)h_state->last_loss;
    if (dg_r) GTRY(get(dg_r, ggrad_r.p, (size_t)nants * nf));
    if (dg_i) GTRY(get(dg_i, ggrad_i.p, (size_t)nants * nf));
    if (dc_r) GTRY(get(dc_r, cgrad_r.p, (size_t)ncoef));
    if (dc_i) GTRY(get(dc_i, cgrad_i.p, (size_t)ncoef));
    return 0;
  }

  int fit(const calb2_fit_options* o, void* hist_out, calb2_fit_result* res) override {
    if (!have_data || !have_gains || !have_coeffs) return fail(CALB2_ERR_STATE, "integration, gains and coefficients must be set first");
    if (pl->nranks > 1) return fail(CALB2_ERR_UNSUPPORTED, "the generic (float64 / oversized-group) path runs on one GPU");
    const bool sum = o->regularization == CALB2_REG_SUM, freeze = o->freeze_model != 0;
    if (sum) GTRY(ensure_sum());
    const size_t ng = (size_t)nants * nf;
    if (o->use_min && !gsnap_r.p) {
      GTRY(alloc(gsnap_r, ng));
      GTRY(alloc(gsnap_i, ng));
      GTRY(alloc(csnap_r, (size_t)ncoef));
      GTRY(alloc(csnap_i, (size_t)ncoef));
    }
    const GConsts<T> k = consts_from(o);
    const long long total = (long long)k.n_skip + o->maxsteps;
    if (cur_buf == 1) {
      CU(cudaMemcpyAsync(g_r[0].p, g_r[1].p, ng * sizeof(T), cudaMemcpyDeviceToDevice, pl->stream));
      CU(cudaMemcpyAsync(g_i[0].p, g_i[1].p, ng * sizeof(T), cudaMemcpyDeviceToDevice, pl->stream));
      cur_buf = 0;
    }
    DevBuf<T>* slots[] = {&gm_r, &gu_r, &gm_i, &gu_i, &cm_r, &cu_r, &cm_i, &cu_i};
    for (auto* s : slots) CU(cudaMemsetAsync(s->p, 0, s->bytes(), pl->stream));
    if (k.optimizer == CALB2_OPT_ADAGRAD || k.optimizer == CALB2_OPT_FTRL) {
      DevBuf<T>* acc_u[] = {&gu_r, &gu_i, &cu_r, &cu_i};
      DevBuf<T>* acc_m[] = {&gm_r, &gm_i, &cm_r, &cm_i};
      for (auto* s : (k.optimizer == CALB2_OPT_ADAGRAD ? acc_u : acc_m)) {
        if (s->n) gen_fill_kernel<T><<<256, 256, 0, pl->stream>>>(s->p, s->n, k.init_acc);
        CU(cudaGetLastError());
      }
    }
    if (hist.n < (size_t)std::max(1, o->maxsteps)) GTRY(alloc(hist, (size_t)std::max(1, o->maxsteps)));
    GState<T> s0{};
    s0.step = 0;
    s0.stop_after = (int)(total - 1);
    s0.min_loss = (T)INFINITY;
    CU(cudaMemcpyAsync(state.p, &s0, sizeof(s0), cudaMemcpyHostToDevice, pl->stream));
    GenParams<T> p = params(state.p, sum, 0, 0);
    p.k = k;
    if (!o->use_min) p.gsnap_r = p.gsnap_i = p.csnap_r = p.csnap_i = nullptr;
    const int chunk = o->steps_per_sync > 0 ? o->steps_per_sync : 32;
    cudaEvent_t ev_begin, ev_end;
    CU(cudaEventCreate(&ev_begin));
    CU(cudaEventCreate(&ev_end));
    CU(cudaEventRecord(ev_begin, pl->stream));
    long long done = 0, launches = 0;
    while (done < total) {
      const int n = (int)std::min<long long>(chunk, total - done);
      for (int i = 0; i < n; ++i) {
        gen_forward_kernel<T><<<dim3(nslots, nfb), GEN_THREADS, 0, pl->stream>>>(p);
        if (!freeze && rows) gen_backward_kernel<T><<<(unsigned)rows, GEN_THREADS, 0, pl->stream>>>(p);
        gen_finalize_kernel<T><<<1, 1024, 0, pl->stream>>>(p, nslots * nfb);
        gen_gains_kernel<T><<<dim3(nfb, nants), GEN_THREADS, 0, pl->stream>>>(p);
        if (!freeze) gen_coeffs_kernel<T><<<(unsigned)((ncoef + 255) / 256), 256, 0, pl->stream>>>(p, 0);
        CU(cudaGetLastError());
        launches += freeze ? 3 : 5;
      }
      done += n;
      CU(cudaMemcpyAsync(h_state, state.p, sizeof(GState<T>), cudaMemcpyDeviceToHost, pl->stream));
      CU(cudaStreamSynchronize(pl->stream));
      if (h_state->step > h_state->stop_after) break;
    }
    CU(cudaEventRecord(ev_end, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    float loop_ms = 0.f;
    CU(cudaEventElapsedTime(&loop_ms, ev_begin, ev_end));
    cudaEventDestroy(ev_begin);
    cudaEventDestroy(ev_end);
    const GState<T> hs = *h_state;
    cur_buf = hs.step & 1;
    if (o->use_min && hs.any_snap) {
      CU(cudaMemcpyAsync(g_r[cur_buf].p, gsnap_r.p, ng * sizeof(T), cudaMemcpyDeviceToDevice, pl->stream));
      CU(cudaMemcpyAsync(g_i[cur_buf].p, gsnap_i.p, ng * sizeof(T), cudaMemcpyDeviceToDevice, pl->stream));
      if (!freeze) {
        CU(cudaMemcpyAsync(c_r.p, csnap_r.p, ncoef * sizeof(T), cudaMemcpyDeviceToDevice, pl->stream));
        CU(cudaMemcpyAsync(c_i.p, csnap_i.p, ncoef * sizeof(T), cudaMemcpyDeviceToDevice, pl->stream));
      }
    }
    if (hs.nrec > 0 && hist_out) CU(cudaMemcpyAsync(hist_out, hist.p, hs.nrec * sizeof(T), cudaMemcpyDeviceToHost, pl->stream));
    CU(cudaStreamSynchronize(pl->stream));
    res->nsteps_recorded = hs.nrec;
    res->nsteps_total = hs.step;
    res->final_loss = (float)(o->use_min ? hs.min_loss : hs.last_loss);
    res->loop_ms = loop_ms;
    res->heavy_ms = 0.f;
    res->heavy_launches = hs.step;
    res->kernel_launches = launches;
    return 0;
  }
#undef GTRY
};

}  // namespace calb2
