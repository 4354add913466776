"""Deterministic synthetic HERA-layout inputs.

Antennas are points of a 14.6 m hexagonal lattice sorted by (radius, angle); every pair i < j is its own
fitting group `(((i, j),),)` exactly as `modeling.yield_pbl_dpss_model_comps` lays the DPSS case out
(/root/reference/calamity/modeling.py:353-356); the basis of a group is the DPSS set for the baseline's
horizon delay (modeling.py:293-300).  Baselines sharing a delay share one ndarray object, so even the
HERA-350 dict costs a few hundred MB on the host.
"""
import numpy as np

from .modeling import dpss_basis


def hex_antenna_positions(nants, spacing=14.6):
    """First `nants` points of a hexagonal lattice ordered by (radius, angle). Returns [nants, 3] ENU metres."""
    n = int(np.ceil(np.sqrt(nants))) + 2
    pts = []
    for a in range(-n, n + 1):
        for b in range(-n, n + 1):
            x = spacing * (a + 0.5 * b)
            y = spacing * (np.sqrt(3.0) / 2.0) * b
            pts.append((round(np.hypot(x, y), 6), round(np.arctan2(y, x), 9), x, y))
    pts.sort()
    return np.asarray([[p[2], p[3], 0.0] for p in pts[:nants]])


def dpss_comps_dict(antpos, freqs, horizon=1.0, min_dly=0.0, offset=0.0, eigenval_cutoff=1e-10, include_autos=False):
    """{(((i, j),),): ndarray[nfreqs, ncomp]} for all i < j (i <= j with autos), i-major order."""
    cache = {}
    comps = {}
    nants = len(antpos)
    for i in range(nants):
        for j in range(i if include_autos else i + 1, nants):
            bllen = float(np.linalg.norm(antpos[i] - antpos[j]))
            dly_ns = int(np.ceil(max(min_dly, bllen / 0.3 * horizon + offset)))
            if dly_ns not in cache:
                cache[dly_ns] = dpss_basis(freqs, dly_ns / 1e9, eigenval_cutoff)
            comps[(((i, j),),)] = cache[dly_ns]
    return comps


class SyntheticProblem:
    """One integration in the library's canonical flat layout (baselines i-major, i < j)."""

    def __init__(self, nants, nfreqs, seed, f0=100e6, bandwidth=None, df=None, gain_scatter=0.1, noise=1e-4,
                 flag_fraction=0.0, init_gain_scatter=0.0, coeff_error=0.0):
        rng = np.random.default_rng(seed)
        self.nants, self.nfreqs = nants, nfreqs
        if df is None:
            df = (100e6 if bandwidth is None else bandwidth) / 1024.0
        self.freqs = f0 + df * np.arange(nfreqs)
        self.antpos = hex_antenna_positions(nants)
        self.comps_dict = dpss_comps_dict(self.antpos, self.freqs)
        keys = list(self.comps_dict.keys())
        self.ant0 = np.asarray([k[0][0][0] for k in keys], dtype=np.int32)
        self.ant1 = np.asarray([k[0][0][1] for k in keys], dtype=np.int32)
        nbls = len(keys)
        self.nbls = nbls
        g_true = 1.0 + gain_scatter * (rng.standard_normal((nants, nfreqs)) + 1j * rng.standard_normal((nants, nfreqs)))
        self.g_true = g_true
        # true coefficients and noiseless foreground visibilities, batched over baselines sharing a basis
        self.ncomp = np.asarray([self.comps_dict[k].shape[1] for k in keys], dtype=np.int32)
        self.coef0 = np.concatenate([[0], np.cumsum(self.ncomp)]).astype(np.int64)
        c_true = np.zeros(int(self.coef0[-1]), dtype=np.complex128)
        vis = np.zeros((nbls, nfreqs), dtype=np.complex128)
        by_basis = {}
        for b, k in enumerate(keys):
            by_basis.setdefault(id(self.comps_dict[k]), []).append(b)
        for members in by_basis.values():
            basis = self.comps_dict[keys[members[0]]]  # [nfreqs, ncomp]
            nc = basis.shape[1]
            decay = np.exp(-3.0 * np.arange(nc) / max(nc, 1))
            c = (rng.standard_normal((len(members), nc)) + 1j * rng.standard_normal((len(members), nc))) * decay
            vis[members] = c @ basis.T
            for m, b in enumerate(members):
                c_true[self.coef0[b] : self.coef0[b + 1]] = c[m]
        self.c_true = c_true
        self.vis_true = vis
        data = g_true[self.ant0] * np.conj(g_true[self.ant1]) * vis
        data = data + noise * (rng.standard_normal(data.shape) + 1j * rng.standard_normal(data.shape))
        flags = rng.random(data.shape) < flag_fraction if flag_fraction > 0 else np.zeros(data.shape, dtype=bool)
        self.flags = flags
        rms = np.sqrt(np.mean(np.abs(data[~flags]) ** 2.0))  # calibration.py:1178-1182
        self.rms = rms
        data = data / rms
        self.data_r = np.ascontiguousarray(data.real, dtype=np.float32)
        self.data_i = np.ascontiguousarray(data.imag, dtype=np.float32)
        w = (~flags).astype(np.float64)
        self.wgts = np.ascontiguousarray(w / w.sum(), dtype=np.float32)  # calibration.py:300-303
        g0 = np.ones((nants, nfreqs), dtype=np.complex128)
        if init_gain_scatter > 0:
            g0 = g0 + init_gain_scatter * (rng.standard_normal(g0.shape) + 1j * rng.standard_normal(g0.shape))
        self.g0_r = np.ascontiguousarray(g0.real, dtype=np.float32)
        self.g0_i = np.ascontiguousarray(g0.imag, dtype=np.float32)
        # initial coefficients: projection of the (scaled) data on the basis, as tensorize_fg_coeffs would give
        # when the sky model is the data itself (calibration.py:1131-1136 with unity gains)
        c0 = np.zeros_like(c_true)
        for members in by_basis.values():
            basis = self.comps_dict[keys[members[0]]]
            proj = (data[members] * (~flags[members])) @ basis
            for m, b in enumerate(members):
                c0[self.coef0[b] : self.coef0[b + 1]] = proj[m]
        if coeff_error > 0:
            c0 = c0 * (1.0 + coeff_error * rng.standard_normal(c0.shape))
        self.c0_r = np.ascontiguousarray(c0.real, dtype=np.float32)
        self.c0_i = np.ascontiguousarray(c0.imag, dtype=np.float32)

    def select_baselines(self, sel):
        """The same antennas, gains and realisation restricted to baselines `sel` (indices in canonical order):
        a cheap way to get benchmark-shaped groups (350 antennas, 1024 channels, the longest baselines' 204-vector
        bases) at a size a CPU oracle can follow.  Weights are renormalised to sum 1 (calibration.py:300-303)."""
        import copy

        sel = np.asarray(sel, dtype=np.int64)
        sub = copy.copy(self)
        keys = list(self.comps_dict.keys())
        sub.comps_dict = {keys[b]: self.comps_dict[keys[b]] for b in sel}
        sub.ant0, sub.ant1, sub.ncomp = self.ant0[sel], self.ant1[sel], self.ncomp[sel]
        sub.nbls = len(sel)
        sub.coef0 = np.concatenate([[0], np.cumsum(sub.ncomp)]).astype(np.int64)
        cidx = (np.concatenate([np.arange(self.coef0[b], self.coef0[b + 1]) for b in sel]) if len(sel)
                else np.zeros(0, np.int64))
        sub.c_true, sub.c0_r, sub.c0_i = self.c_true[cidx], self.c0_r[cidx], self.c0_i[cidx]
        sub.vis_true, sub.flags = self.vis_true[sel], self.flags[sel]
        sub.data_r, sub.data_i = self.data_r[sel], self.data_i[sel]
        w = self.wgts[sel].astype(np.float64)
        sub.wgts = np.ascontiguousarray(w / w.sum(), dtype=np.float32)
        return sub

    def layout(self, dtype=np.float32):
        from .layout import RaggedLayout

        ants_map = {a: a for a in range(self.nants)}
        chunked = {(1, int(self.ncomp.max())): self.comps_dict}
        return RaggedLayout.from_chunked_dict(chunked, ants_map, self.nfreqs, nants=self.nants, dtype=dtype)


# configs of BASELINE.json, in order
CONFIGS = {
    "test6": dict(nants=6, nfreqs=200, df=100e3),
    "hera37": dict(nants=37, nfreqs=384),
    "hera128": dict(nants=128, nfreqs=1024),
    "hera350": dict(nants=350, nfreqs=1024),
    # the one configuration the reference publishes a number for (examples/Calamity_Tutorial.ipynb:1178): 15 antennas,
    # 105 baselines x 200 channels
    "tutorial15": dict(nants=15, nfreqs=200, df=100e3),
}


def make(config, seed_offset=0, **overrides):
    kw = dict(CONFIGS[config])
    kw.update(overrides)
    seed = 20260101 + list(CONFIGS).index(config) + seed_offset
    return SyntheticProblem(seed=seed, **kw)
