"""NumPy restatement of the reference fit path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

All file:line citations are into /root/reference/calamity/calibration.py unless noted.
The arithmetic dtype is a parameter: float32 mimics the reference's default graph
(`dtype=np.float32`, calibration.py:974), float64 is the high-precision yardstick both
float32 implementations (this one and the CUDA one) are measured against.

Layout conventions are the reference's: per chunk c
    fg_comps[c]   [nvecs, ngrps, nbls, nfreqs]       (calibration.py:167-187)
    fg_r/fg_i[c]  [nvecs, ngrps, 1, 1]               (calibration.py:906)
    data_r/data_i/wgts[c] [ngrps, nbls, nfreqs]      (calibration.py:305-308)
    corr_inds[c][g][b] = (ant0_index, ant1_index)    (calibration.py:176-177)
    g_r/g_i       [nants, nfreqs]                    (calibration.py:397-398)
"""
import copy

import numpy as np


# --------------------------------------------------------------------------------------
# marshalling
# --------------------------------------------------------------------------------------
def chunk_by_nbls(comps_dict, use_redundancy=False, grp_size_threshold=5):
    """calibration.py:30-101.  Returns {(nbl, maxvecs): {fit_grp: array}} in first-appearance order.

    Ordering matters (it fixes group / baseline indices downstream): groups that get split
    are popped and their per-baseline replacements are appended at the END of the dict
    (calibration.py:73-81), everything else keeps its place.
    """
    work = copy.deepcopy(comps_dict)
    if not use_redundancy:
        for grp in list(work.keys()):
            sizes = np.asarray([len(rg) for rg in grp])
            if np.allclose(sizes, np.mean(sizes)) and len(sizes) < grp_size_threshold:
                vecs = work.pop(grp)
                for m in range(int(sizes[0])):
                    work[tuple((rg[m],) for rg in grp)] = vecs
    members, widest = {}, {}
    for grp, vecs in work.items():
        n = sum(len(rg) for rg in grp)
        if n not in members:
            members[n] = [grp]
            widest[n] = vecs.shape[1]
        else:
            members[n].append(grp)
            widest[n] = max(widest[n], vecs.shape[1])
    return {(n, widest[n]): {g: work[g] for g in members[n]} for n in members}


def tensorize_basis(comps_dict, ants_map, nfreqs, use_redundancy=False, dtype=np.float32, grp_size_threshold=5):
    """calibration.py:104-190: dense zero-padded [nvecs, ngrps, nbls, nfreqs] per chunk + corr_inds."""
    chunked = chunk_by_nbls(comps_dict, use_redundancy=use_redundancy, grp_size_threshold=grp_size_threshold)
    tensors, corr_inds = [], []
    for (nbls, nvecs), grp_dict in chunked.items():
        dense = np.zeros((nvecs, len(grp_dict), nbls, nfreqs))
        chunk_inds = []
        for g, (grp, vecs) in enumerate(grp_dict.items()):
            inds, b = [], 0
            for rnum, red in enumerate(grp):
                block = vecs[rnum * nfreqs : (rnum + 1) * nfreqs].T  # [ncomp, nfreqs]
                for ap in red:
                    inds.append((ants_map[ap[0]], ants_map[ap[1]]))
                    dense[: vecs.shape[1], g, b] = block
                    b += 1
            chunk_inds.append(inds)
        tensors.append(dense.astype(dtype))
        corr_inds.append(chunk_inds)
    return tensors, corr_inds


def gather_chunks(cube, corr_inds):
    """tf.gather_nd(cube, corr_inds[c]) for every chunk (calibration.py:305-308)."""
    out = []
    for chunk in corr_inds:
        idx = np.asarray(chunk, dtype=np.int64)  # [ngrps, nbls, 2]
        out.append(cube[idx[..., 0], idx[..., 1]])
    return out


def tensorize_data_cubes(data_cube, flag_cube, nsample_cube, corr_inds, data_scale_factor=1.0,
                         weights_cube=None, nsamples_in_weights=False, dtype=np.float32):
    """calibration.py:251-310 with the UVData lookups already resolved into [nants,nants,nfreqs] cubes.

    data_cube is complex and already conjugated into (i, j) orientation.  Only the (i, j)
    cells named in corr_inds are used, exactly like the reference loop.
    """
    nants, _, nfreqs = data_cube.shape
    d_r = np.zeros((nants, nants, nfreqs), dtype=dtype)
    d_i = np.zeros_like(d_r)
    w = np.zeros_like(d_r)
    wsum = 0.0
    for chunk in corr_inds:
        for grp in chunk:
            for (i, j) in grp:
                vis = data_cube[i, j] / data_scale_factor
                d_r[i, j] = vis.real.astype(dtype)
                d_i[i, j] = vis.imag.astype(dtype)
                unflagged = ~flag_cube[i, j]
                if weights_cube is None:
                    w[i, j] = unflagged
                else:
                    w[i, j] = weights_cube[i, j].astype(dtype) * unflagged
                if nsamples_in_weights:
                    w[i, j] *= nsample_cube[i, j]
                wsum += np.sum(w[i, j])
    w = (w / wsum).astype(dtype)
    return gather_chunks(d_r, corr_inds), gather_chunks(d_i, corr_inds), gather_chunks(w, corr_inds)


def init_coeffs(data, wgts, fg_comps):
    """calibration.py:828-913: unweighted least squares of (data * (w != 0)) on the leading non-zero rows."""
    out = []
    for d, w, comps in zip(data, wgts, fg_comps):
        nvecs, ngrps = comps.shape[:2]
        ndata = d.shape[1] * d.shape[2]
        mask = (~np.isclose(w, 0.0)).astype(w.dtype)
        cols = []
        for g in range(ngrps):
            rows = comps[:, g].reshape(nvecs, ndata)
            empty = np.where(np.all(np.isclose(rows, 0.0), axis=1))[0]
            nnz = int(empty.min()) if len(empty) > 0 else nvecs
            amat = rows[:nnz].T
            rhs = (d[g] * mask[g]).reshape(ndata)
            # tf.linalg.lstsq(fast=True): Cholesky solve of the normal equations, in dtype
            gram = amat.T @ amat
            sol = np.linalg.solve(gram, amat.T @ rhs) if nnz > 0 else np.zeros(0, dtype=d.dtype)
            cols.append(np.concatenate([sol, np.zeros(nvecs - nnz, dtype=sol.dtype)]).astype(d.dtype))
        out.append(np.stack(cols).T.reshape(nvecs, ngrps, 1, 1))
    return out


def model_cube(nants, nfreqs, fg_comps, fg_coeffs, corr_inds):
    """calibration.py:402-444: float64 [nants,nants,nfreqs], only cell (i,j) filled."""
    cube = np.zeros((nants, nants, nfreqs))
    for comps, coeffs, chunk in zip(fg_comps, fg_coeffs, corr_inds):
        vis = np.sum(coeffs * comps, axis=0)
        for g, grp in enumerate(chunk):
            for b, (i, j) in enumerate(grp):
                cube[i, j] = vis[g, b]
    return cube


def ant_index_arrays(corr_inds):
    """calibration.py:577-594: per chunk [ngrps, nbls] integer arrays of ant0 / ant1."""
    a0 = [np.asarray([[p[0] for p in grp] for grp in chunk], dtype=np.int64) for chunk in corr_inds]
    a1 = [np.asarray([[p[1] for p in grp] for grp in chunk], dtype=np.int64) for chunk in corr_inds]
    return a0, a1


# --------------------------------------------------------------------------------------
# model / loss / analytic gradient
# --------------------------------------------------------------------------------------
def fg_vis(fg_r, fg_i, comps):
    """calibration.py:1587-1590."""
    return np.sum(fg_r * comps, axis=0), np.sum(fg_i * comps, axis=0)


def forward_chunk(g_r, g_i, fg_r, fg_i, comps, a0, a1):
    """calibration.py:1593-1605.  Returns model_r, model_i and the intermediates the gradient needs."""
    gr0, gr1, gi0, gi1 = g_r[a0], g_r[a1], g_i[a0], g_i[a1]
    pp = gr0 * gr1 + gi0 * gi1
    qq = gr0 * gi1 - gi0 * gr1
    vr, vi = fg_vis(fg_r, fg_i, comps)
    return pp * vr + qq * vi, -qq * vr + pp * vi, (gr0, gr1, gi0, gi1, pp, qq, vr, vi)


def loss_value(g_r, g_i, fg_r, fg_i, data_r, data_i, wgts, fg_comps, corr_inds, regularization=None,
               prior_r_sum=None, prior_i_sum=None):
    """calibration.py:1612-1620 (plain) / 1623-1656 ('sum')."""
    a0s, a1s = ant_index_arrays(corr_inds)
    dt = g_r.dtype
    chi, sr, si = [], [], []
    for c in range(len(fg_comps)):
        m_r, m_i, _ = forward_chunk(g_r, g_i, fg_r[c], fg_i[c], fg_comps[c], a0s[c], a1s[c])
        chi.append(np.sum((np.square(data_r[c] - m_r) + np.square(data_i[c] - m_i)) * wgts[c], dtype=dt))
        sr.append(np.sum(m_r * wgts[c], dtype=dt))
        si.append(np.sum(m_i * wgts[c], dtype=dt))
    total = np.sum(np.stack(chi), dtype=dt)
    if regularization == "sum":
        total = total + np.square(np.sum(np.stack(sr), dtype=dt) - prior_r_sum) + np.square(
            np.sum(np.stack(si), dtype=dt) - prior_i_sum
        )
    return dt.type(total)


def sum_priors(sky_r, sky_i, wgts, dtype):
    """calibration.py:619-625."""
    dt = np.dtype(dtype)
    pr = np.sum(np.stack([np.sum(s * w, dtype=dt) for s, w in zip(sky_r, wgts)]), dtype=dt)
    pi = np.sum(np.stack([np.sum(s * w, dtype=dt) for s, w in zip(sky_i, wgts)]), dtype=dt)
    return dt.type(pr), dt.type(pi)


def loss_and_grads(g_r, g_i, fg_r, fg_i, data_r, data_i, wgts, fg_comps, corr_inds, regularization=None,
                   prior_r_sum=None, prior_i_sum=None):
    """Loss plus the closed-form reverse-mode gradient of calibration.py:664-666.

    With P = gr0 gr1 + gi0 gi1, Q = gr0 gi1 - gi0 gr1, m = (P + iQ)... (model_r = P v_r + Q v_i,
    model_i = -Q v_r + P v_i):
        e_r = -2 w (d_r - m_r) [+ 2 (S_r - P_r) w]      e_i likewise
        dL/dv_r = P e_r - Q e_i        dL/dv_i = Q e_r + P e_i
        dL/dc_r[k,g] = sum_{b,f} A[k,g,b,f] dL/dv_r[g,b,f]   (same for c_i)
        a = e_r v_r + e_i v_i ;  b = e_r v_i - e_i v_r
        d gr0 += a gr1 + b gi1 ; d gi0 += a gi1 - b gr1 ; d gr1 += a gr0 - b gi0 ; d gi1 += a gi0 + b gr0
    Returns loss, dg_r, dg_i, [dfg_r per chunk], [dfg_i per chunk].
    """
    a0s, a1s = ant_index_arrays(corr_inds)
    dt = g_r.dtype
    nchunks = len(fg_comps)
    fwd = [forward_chunk(g_r, g_i, fg_r[c], fg_i[c], fg_comps[c], a0s[c], a1s[c]) for c in range(nchunks)]
    chi = [np.sum((np.square(data_r[c] - fwd[c][0]) + np.square(data_i[c] - fwd[c][1])) * wgts[c], dtype=dt)
           for c in range(nchunks)]
    loss = np.sum(np.stack(chi), dtype=dt)
    alpha = beta = dt.type(0)
    if regularization == "sum":
        s_r = np.sum(np.stack([np.sum(fwd[c][0] * wgts[c], dtype=dt) for c in range(nchunks)]), dtype=dt)
        s_i = np.sum(np.stack([np.sum(fwd[c][1] * wgts[c], dtype=dt) for c in range(nchunks)]), dtype=dt)
        loss = loss + np.square(s_r - prior_r_sum) + np.square(s_i - prior_i_sum)
        alpha = dt.type(2) * (s_r - prior_r_sum)
        beta = dt.type(2) * (s_i - prior_i_sum)
    dg_r = np.zeros_like(g_r)
    dg_i = np.zeros_like(g_i)
    dfg_r, dfg_i = [], []
    for c in range(nchunks):
        m_r, m_i, (gr0, gr1, gi0, gi1, pp, qq, vr, vi) = fwd[c]
        w = wgts[c]
        e_r = dt.type(-2) * w * (data_r[c] - m_r) + alpha * w
        e_i = dt.type(-2) * w * (data_i[c] - m_i) + beta * w
        dv_r = pp * e_r - qq * e_i
        dv_i = qq * e_r + pp * e_i
        dfg_r.append(np.sum(fg_comps[c] * dv_r[None], axis=(2, 3), dtype=dt)[:, :, None, None])
        dfg_i.append(np.sum(fg_comps[c] * dv_i[None], axis=(2, 3), dtype=dt)[:, :, None, None])
        aa = e_r * vr + e_i * vi
        bb = e_r * vi - e_i * vr
        np.add.at(dg_r, a0s[c], aa * gr1 + bb * gi1)
        np.add.at(dg_i, a0s[c], aa * gi1 - bb * gr1)
        np.add.at(dg_r, a1s[c], aa * gr0 - bb * gi0)
        np.add.at(dg_i, a1s[c], aa * gi0 + bb * gr0)
    return dt.type(loss), dg_r, dg_i, dfg_r, dfg_i


# --------------------------------------------------------------------------------------
# Keras OptimizerV2 update rules (tensorflow >= 2.4, unpinned; arithmetic restated from the
# published ApplyAdaMax / ApplyAdam training-op kernels and the optimizer_v2 sparse paths --
# the sources are NOT in /root/reference, see DESIGN.md "parity pin status").
# --------------------------------------------------------------------------------------
KERAS_DEFAULTS = {
    "Adamax": dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7),
    "Adam": dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7),
    "SGD": dict(learning_rate=0.01, momentum=0.0, nesterov=False),
    "RMSprop": dict(learning_rate=0.001, rho=0.9, momentum=0.0, epsilon=1e-7),
    "Adagrad": dict(learning_rate=0.001, initial_accumulator_value=0.1, epsilon=1e-7),
    "Adadelta": dict(learning_rate=0.001, rho=0.95, epsilon=1e-7),
    "Nadam": dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7),
    "Ftrl": dict(learning_rate=0.001, learning_rate_power=-0.5, initial_accumulator_value=0.1,
                 l1_regularization_strength=0.0, l2_regularization_strength=0.0),
    # tensorflow_addons.optimizers.LAMB (calibration.py:15, 26); exclude_from_* lists are not restated
    "LAMB": dict(learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-6, weight_decay=0.0),
}


class KerasOptimizer:
    """State: one step counter shared by all variables, two slots per variable (created lazily).

    `sparse` marks variables whose gradient arrives as IndexedSlices (the gains, because they are
    read through tf.gather, calibration.py:1594-1597); Keras then uses the *_sparse update, whose
    algebra is the dense rule written as m*b1 + g*(1-b1) instead of m + (g-m)*(1-b1).
    """

    def __init__(self, name, **kwargs):
        if name not in KERAS_DEFAULTS:
            raise KeyError(name)
        unknown = set(kwargs) - set(KERAS_DEFAULTS[name])
        if unknown:
            raise TypeError(f"unexpected optimizer kwargs {sorted(unknown)}")
        self.name = name
        self.hp = dict(KERAS_DEFAULTS[name], **kwargs)
        self.iterations = 0
        self.slots = {}
        self.m_schedule = 1.0  # Nadam `_m_cache`

    def _other_rules(self, p, g, m, u, t, dt, nadam):
        """RMSprop / Adagrad / Adadelta / Nadam / Ftrl (tf.keras.optimizers, OptimizerV2; the raw-op formulas
        ApplyRMSProp, ApplyAdagradV2, ApplyAdadelta, ApplyFtrlV2 and Nadam._resource_apply_dense).  For these the
        IndexedSlices (`*_sparse`) variants compute the same expressions, so there is one form."""
        hp = self.hp
        lr = dt(hp["learning_rate"])
        one = dt(1)
        if self.name == "RMSprop":
            rho, mom, eps = dt(hp["rho"]), dt(hp["momentum"]), dt(hp["epsilon"])
            u[...] = rho * u + (one - rho) * (g * g)
            if mom > 0:  # fused op: epsilon inside the square root
                m[...] = mom * m + lr * g / np.sqrt(u + eps)
                p -= m
            else:
                p -= lr * g / (np.sqrt(u) + eps)
        elif self.name == "Adagrad":
            u += g * g
            p -= lr * g / (np.sqrt(u) + dt(hp["epsilon"]))
        elif self.name == "Adadelta":
            rho, eps = dt(hp["rho"]), dt(hp["epsilon"])
            m[...] = m * rho + (g * g) * (one - rho)
            upd = np.sqrt(u + eps) / np.sqrt(m + eps) * g
            u[...] = u * rho + (upd * upd) * (one - rho)
            p -= upd * lr
        elif self.name == "Nadam":
            b1, b2, eps = dt(hp["beta_1"]), dt(hp["beta_2"]), dt(hp["epsilon"])
            u_t, u_t1, ms_new, ms_next = nadam
            g_prime = g / (one - ms_new)
            m[...] = b1 * m + (one - b1) * g
            m_prime = m / (one - ms_next)
            u[...] = b2 * u + (one - b2) * (g * g)
            v_prime = u / (one - dt(np.power(b2, dt(t))))
            m_bar = (one - u_t) * g_prime + u_t1 * m_prime
            p -= lr * m_bar / (np.sqrt(v_prime) + eps)
        elif self.name == "Ftrl":
            lp, l1, l2 = dt(hp["learning_rate_power"]), dt(hp["l1_regularization_strength"]), dt(hp["l2_regularization_strength"])
            acc_new = m + g * g
            pw = (lambda x: np.sqrt(x)) if lp == dt(-0.5) else (lambda x: np.power(x, -lp))
            u += g - (pw(acc_new) - pw(m)) / lr * p
            quadratic = pw(acc_new) / lr + dt(2) * l2
            m[...] = acc_new
            p[...] = np.where(np.abs(u) > l1, (np.sign(u) * l1 - u) / quadratic, dt(0))
        else:
            raise KeyError(self.name)

    def apply(self, params, grads, sparse_flags):
        self.iterations += 1
        t = self.iterations
        nadam = None
        if self.name == "Nadam":  # Nadam._prepare_local: one schedule update per step, shared by all variables
            dt0 = params[0].dtype.type
            b1 = dt0(self.hp["beta_1"])
            u_t = b1 * (dt0(1) - dt0(0.5) * dt0(np.power(0.96, 0.004 * t)))
            u_t1 = b1 * (dt0(1) - dt0(0.5) * dt0(np.power(0.96, 0.004 * (t + 1))))
            ms_new = dt0(self.m_schedule) * u_t
            self.m_schedule = float(ms_new)
            nadam = (u_t, u_t1, ms_new, ms_new * u_t1)
        for n, (p, g, sp) in enumerate(zip(params, grads, sparse_flags)):
            dt = p.dtype.type
            if n not in self.slots:
                self.slots[n] = (np.zeros_like(p), np.zeros_like(p))
                if self.name == "Adagrad":
                    self.slots[n][1][...] = dt(self.hp["initial_accumulator_value"])
                if self.name == "Ftrl":
                    self.slots[n][0][...] = dt(self.hp["initial_accumulator_value"])
            m, u = self.slots[n]
            if self.name == "SGD":
                lr0, mom = dt(self.hp["learning_rate"]), dt(self.hp["momentum"])
                if mom == 0:
                    p -= lr0 * g
                else:  # ApplyKerasMomentum
                    m[...] = m * mom - lr0 * g
                    p += (m * mom - lr0 * g) if self.hp["nesterov"] else m
                continue
            if self.name == "LAMB":
                # tfa.optimizers.LAMB._resource_apply_dense (the _sparse variant computes the same expressions): Adam moments
                # with bias correction, then a PER-VARIABLE trust ratio ||w|| / ||update|| (1 when either norm is 0)
                lr, b1, b2, eps, wd = (dt(self.hp[k]) for k in ("learning_rate", "beta_1", "beta_2", "epsilon", "weight_decay"))
                m[...] = m * b1 + g * (dt(1) - b1)
                u[...] = u * b2 + (g * g) * (dt(1) - b2)
                m_hat = m / (dt(1) - dt(np.power(b1, dt(t))))
                u_hat = u / (dt(1) - dt(np.power(b2, dt(t))))
                upd = m_hat / (np.sqrt(u_hat) + eps)
                if wd != 0:
                    upd = upd + wd * p
                w_norm = dt(np.sqrt(np.sum(np.square(p, dtype=np.float64))))
                g_norm = dt(np.sqrt(np.sum(np.square(upd, dtype=np.float64))))
                ratio = (w_norm / g_norm) if (w_norm > 0 and g_norm > 0) else dt(1)
                p -= ratio * lr * upd
                continue
            if self.name not in ("Adamax", "Adam"):
                self._other_rules(p, g, m, u, t, dt, nadam)
                continue
            lr, b1, b2, eps = (dt(self.hp[k]) for k in ("learning_rate", "beta_1", "beta_2", "epsilon"))
            b1p = dt(np.power(b1, dt(t)))
            if self.name == "Adamax":
                if sp:
                    m[...] = m * b1 + g * (dt(1) - b1)
                else:
                    m[...] = m + (g - m) * (dt(1) - b1)
                u[...] = np.maximum(u * b2, np.abs(g))
                if sp:
                    p += (-(lr / (dt(1) - b1p))) * (m / (u + eps))
                else:
                    p -= (lr / (dt(1) - b1p)) * (m / (u + eps))
            else:  # Adam
                b2p = dt(np.power(b2, dt(t)))
                lr_t = lr * np.sqrt(dt(1) - b2p) / (dt(1) - b1p)
                if sp:
                    m[...] = m * b1 + g * (dt(1) - b1)
                    u[...] = u * b2 + (g * g) * (dt(1) - b2)
                else:
                    m[...] = m + (g - m) * (dt(1) - b1)
                    u[...] = u + (g * g - u) * (dt(1) - b2)
                p -= (m * lr_t) / (np.sqrt(u) + eps)


# --------------------------------------------------------------------------------------
# fit loop
# --------------------------------------------------------------------------------------
def fit(g_r, g_i, fg_r, fg_i, data_r, data_i, wgts, fg_comps, corr_inds, use_min=False, tol=1e-14,
        maxsteps=10000, optimizer="Adamax", freeze_model=False, n_profile_steps=0, sky_model_r=None,
        sky_model_i=None, model_regularization=None, grad_fn=loss_and_grads, **opt_kwargs):
    """calibration.py:447-738 (loop semantics Q1-Q5 of SURVEY.md section 7).

    Returns g_r, g_i, fg_r, fg_i, {'loss': [...]} with every array a fresh copy in the input dtype.
    """
    opt = KerasOptimizer(optimizer, **opt_kwargs)  # calibration.py:571 (KeyError on unknown name)
    dt = np.dtype(g_r.dtype)
    g_r, g_i = g_r.copy(), g_i.copy()
    fg_r = [x.copy() for x in fg_r]
    fg_i = [x.copy() for x in fg_i]
    pr = pi = None
    if model_regularization == "sum":
        pr, pi = sum_priors(sky_model_r, sky_model_i, wgts, dt)
    reg = "sum" if model_regularization == "sum" else None
    nchunks = len(fg_comps)

    def train_step():  # calibration.py:663-668: returns the PRE-update loss
        loss, dgr, dgi, dfr, dfi = grad_fn(g_r, g_i, fg_r, fg_i, data_r, data_i, wgts, fg_comps, corr_inds,
                                           regularization=reg, prior_r_sum=pr, prior_i_sum=pi)
        if freeze_model:
            opt.apply([g_r, g_i], [dgr, dgi], [True, True])
        else:
            opt.apply([g_r, g_i] + fg_r + fg_i, [dgr, dgi] + dfr + dfi, [True, True] + [False] * (2 * nchunks))
        return loss

    for _ in range(n_profile_steps):  # calibration.py:681-687 -- real, state-advancing steps
        train_step()
    train_step()  # calibration.py:693 unrecorded warm-up
    history = []
    min_loss = 9e99
    best = None
    for step in range(maxsteps):  # calibration.py:699-717
        loss = train_step()
        history.append(dt.type(loss))
        if use_min and history[-1] < min_loss:
            min_loss = history[-1]
            best = (g_r.copy(), g_i.copy(), [x.copy() for x in fg_r], [x.copy() for x in fg_i])
        if step >= 1 and np.abs(history[-1] - history[-2]) < tol:
            break
    if not use_min:
        if len(history) == 0:
            raise IndexError("list index out of range")  # calibration.py:723 with maxsteps == 0
        best = (g_r, g_i, fg_r, fg_i)
    elif freeze_model:
        raise UnboundLocalError("fg_r_opt")  # calibration.py:738 (quirk Q4)
    return best[0], best[1], best[2], best[3], {"loss": history}
