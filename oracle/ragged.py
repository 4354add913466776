"""Ragged, basis-batched NumPy restatement of the fit path (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Same arithmetic as ``oracle/restatement.py`` (model / loss / analytic gradient of calibration.py:1587-1656 and
664-666, Keras rules, loop semantics of 447-738) but on the library's flat description -- groups -> slots ->
baselines, no zero padding -- and with the contraction batched over the groups that share one basis block, so
the float64 yardstick can be evaluated at the benchmarked sizes (HERA-128 / HERA-350, config-5 joint groups)
where the dense `[nvecs, ngrps, nbls, nfreqs]` tensors of the reference (51 GB at HERA-350) cannot be built.
It is pinned against the dense restatement on the small configurations by tests/test_oracle_ragged_cpu.py.

Parameters are flat: g_r, g_i [nants, nfreqs]; c_r, c_i [ncoef] (canonical coefficient order, calibration.py:906);
data_r, data_i, wgts [nbls, nfreqs] (canonical baseline order, calibration.py:305-308).
"""
import numpy as np

from . import restatement as R


class RaggedProblem:
    """Index structure of a RaggedLayout, grouped by shared basis block (object identity of layout.blocks)."""

    def __init__(self, layout, dtype=np.float64):
        self.nants, self.nfreqs = layout.nants, layout.nfreqs
        self.dtype = np.dtype(dtype)
        self.ncoef, self.nbls = layout.ncoef, layout.nbls
        self.bl_ant0 = np.asarray(layout.bl_ant0, dtype=np.int64)
        self.bl_ant1 = np.asarray(layout.bl_ant1, dtype=np.int64)
        slot0 = np.concatenate([[0], np.cumsum(layout.group_nslots)]).astype(np.int64)
        self.nslots = int(slot0[-1])
        self.bl_slot = np.repeat(np.arange(self.nslots), layout.slot_nbls)
        self._single_bl = bool(np.all(np.asarray(layout.slot_nbls) == 1))
        # segment sums over the baselines of an antenna / of a slot: stable sort + reduceat (np.add.at is far too slow at
        # 61 075 x 1024); the order inside a segment is the canonical baseline order
        self._seg = {}
        for name, key, n in (("a0", self.bl_ant0, self.nants), ("a1", self.bl_ant1, self.nants),
                             ("slot", self.bl_slot, self.nslots)):
            order = np.argsort(key, kind="stable")
            present, start = np.unique(key[order], return_index=True)
            self._seg[name] = (order, present, start, n)
        # classes: groups sharing one [nslots, ncomp, nfreqs] block
        by_block = {}
        for g, blk in enumerate(layout.blocks):
            by_block.setdefault(id(blk), []).append(g)
        self.classes = []
        for members in by_block.values():
            blk = np.asarray(layout.blocks[members[0]], dtype=self.dtype)  # [nslots, ncomp, nfreqs]
            ns, nc, _ = blk.shape
            members = np.asarray(members, dtype=np.int64)
            coef_idx = layout.group_coef0[members][:, None] + np.arange(nc)[None, :]   # [nmem, ncomp]
            slot_idx = slot0[members][:, None] + np.arange(ns)[None, :]                # [nmem, nslots]
            self.classes.append((blk, coef_idx, slot_idx))

    def _segsum(self, name, x):
        order, present, start, n = self._seg[name]
        out = np.zeros((n,) + x.shape[1:], dtype=x.dtype)
        if len(order):
            out[present] = np.add.reduceat(x[order], start, axis=0)
        return out

    # -- forward: model visibility per slot, v = sum_k c_k A_k (calibration.py:1587-1590)
    def slot_vis(self, c_r, c_i):
        v_r = np.zeros((self.nslots, self.nfreqs), dtype=self.dtype)
        v_i = np.zeros_like(v_r)
        for blk, coef_idx, slot_idx in self.classes:
            for s in range(blk.shape[0]):
                v_r[slot_idx[:, s]] = c_r[coef_idx] @ blk[s]
                v_i[slot_idx[:, s]] = c_i[coef_idx] @ blk[s]
        return v_r, v_i

    def loss_and_grads(self, g_r, g_i, c_r, c_i, data_r, data_i, wgts, regularization=None, prior_r_sum=None,
                       prior_i_sum=None):
        """Formulas of restatement.loss_and_grads (calibration.py:1593-1656 + tape gradient)."""
        dt = self.dtype.type
        a0, a1 = self.bl_ant0, self.bl_ant1
        sv_r, sv_i = self.slot_vis(c_r, c_i)
        vr, vi = sv_r[self.bl_slot], sv_i[self.bl_slot]
        gr0, gr1, gi0, gi1 = g_r[a0], g_r[a1], g_i[a0], g_i[a1]
        pp = gr0 * gr1 + gi0 * gi1
        qq = gr0 * gi1 - gi0 * gr1
        m_r = pp * vr + qq * vi
        m_i = -qq * vr + pp * vi
        loss = np.sum((np.square(data_r - m_r) + np.square(data_i - m_i)) * wgts, dtype=self.dtype)
        alpha = beta = dt(0)
        if regularization == "sum":
            s_r = np.sum(m_r * wgts, dtype=self.dtype)
            s_i = np.sum(m_i * wgts, dtype=self.dtype)
            loss = loss + np.square(s_r - prior_r_sum) + np.square(s_i - prior_i_sum)
            alpha = dt(2) * (s_r - prior_r_sum)
            beta = dt(2) * (s_i - prior_i_sum)
        e_r = dt(-2) * wgts * (data_r - m_r) + alpha * wgts
        e_i = dt(-2) * wgts * (data_i - m_i) + beta * wgts
        dv_r = pp * e_r - qq * e_i
        dv_i = qq * e_r + pp * e_i
        # redundant baselines of a slot share the model visibility: sum dL/dv over them
        q_r, q_i = (dv_r, dv_i) if self._single_bl else (self._segsum("slot", dv_r), self._segsum("slot", dv_i))
        dc_r = np.zeros(self.ncoef, dtype=self.dtype)
        dc_i = np.zeros_like(dc_r)
        for blk, coef_idx, slot_idx in self.classes:
            for s in range(blk.shape[0]):
                dc_r[coef_idx] += q_r[slot_idx[:, s]] @ blk[s].T
                dc_i[coef_idx] += q_i[slot_idx[:, s]] @ blk[s].T
        aa = e_r * vr + e_i * vi
        bb = e_r * vi - e_i * vr
        dg_r = self._segsum("a0", aa * gr1 + bb * gi1) + self._segsum("a1", aa * gr0 - bb * gi0)
        dg_i = self._segsum("a0", aa * gi1 - bb * gr1) + self._segsum("a1", aa * gi0 + bb * gr0)
        return dt(loss), dg_r, dg_i, dc_r, dc_i

    def fit(self, g_r, g_i, c_r, c_i, data_r, data_i, wgts, use_min=False, tol=1e-14, maxsteps=10000,
            optimizer="Adamax", freeze_model=False, n_profile_steps=0, model_regularization=None,
            prior_r_sum=None, prior_i_sum=None, coef_var_bounds=None, **opt_kwargs):
        """Loop of restatement.fit (calibration.py:447-738) on flat parameters.  The Keras rules are elementwise,
        so applying them to the flat coefficient vector equals applying them to the per-chunk tensors; LAMB's trust
        ratio is per variable, so for it `coef_var_bounds` ([nchunks + 1] coefficient offsets) splits the flat vectors
        into the reference's per-chunk variables fg_r[c] / fg_i[c] (calibration.py:560-567), in the reference's order
        g_r, g_i, fg_r[0..], fg_i[0..]."""
        opt = R.KerasOptimizer(optimizer, **opt_kwargs)
        vb = [0, None] if coef_var_bounds is None else [int(b) for b in coef_var_bounds]
        nv = len(vb) - 1
        dt = self.dtype
        g_r, g_i, c_r, c_i = (np.array(x, dtype=dt) for x in (g_r, g_i, c_r, c_i))
        data_r, data_i, wgts = (np.asarray(x, dtype=dt) for x in (data_r, data_i, wgts))
        reg = "sum" if model_regularization == "sum" else None

        def train_step():
            loss, dgr, dgi, dcr, dci = self.loss_and_grads(g_r, g_i, c_r, c_i, data_r, data_i, wgts, regularization=reg,
                                                           prior_r_sum=prior_r_sum, prior_i_sum=prior_i_sum)
            if freeze_model:
                opt.apply([g_r, g_i], [dgr, dgi], [True, True])
            else:
                opt.apply([g_r, g_i] + [c_r[vb[v] : vb[v + 1]] for v in range(nv)] + [c_i[vb[v] : vb[v + 1]] for v in range(nv)],
                          [dgr, dgi] + [dcr[vb[v] : vb[v + 1]] for v in range(nv)] + [dci[vb[v] : vb[v + 1]] for v in range(nv)],
                          [True, True] + [False] * (2 * nv))
            return loss

        for _ in range(n_profile_steps):
            train_step()
        train_step()
        history, min_loss, best = [], 9e99, None
        for step in range(maxsteps):
            loss = train_step()
            history.append(dt.type(loss))
            if use_min and history[-1] < min_loss:
                min_loss = history[-1]
                best = (g_r.copy(), g_i.copy(), c_r.copy(), c_i.copy())
            if step >= 1 and np.abs(history[-1] - history[-2]) < tol:
                break
        if not use_min:
            best = (g_r, g_i, c_r, c_i)
        return best[0], best[1], best[2], best[3], {"loss": history}
