"""Synthetic stand-ins for the reference's uvh5 fixtures (which need h5py/pyuvdata to read; neither is
installed).  File-name parameters of calamity/data/*.uvh5 are re-created: 6 antennas on an E-W line at
2 m x [0, 1, 4, 10, 12, 17], 200 channels from 100 MHz in 100 kHz steps, one time, pol 'xx'; and a
"3 antennas x 2 copies" redundant array."""
import copy

import numpy as np

from calamity_b200 import cal_utils, modeling
from calamity_b200.uvstandins import MiniUVData, MiniUVFlag

T0 = 2458000.123


def _point_source_vis(antpos, antpairs, freqs, rng, nsrc=40, max_zenith_sine=1.0):
    """Foregrounds: flat-ish spectrum point sources anywhere above the horizon -> delay-limited visibilities."""
    l = rng.uniform(-max_zenith_sine, max_zenith_sine, nsrc)
    flux = rng.uniform(0.5, 2.0, nsrc) * 10.0
    alpha = rng.uniform(-1.0, -0.5, nsrc)
    vis = np.zeros((len(antpairs), len(freqs)), dtype=np.complex128)
    for n, (a, b) in enumerate(antpairs):
        bl = antpos[b][0] - antpos[a][0]
        phase = -2j * np.pi * bl * np.outer(l, freqs) / 3e8
        vis[n] = np.sum(flux[:, None] * (freqs[None, :] / freqs[0]) ** alpha[:, None] * np.exp(phase), axis=0)
    return vis


def line_array(seed=0, ntimes=1, with_autos=False):
    rng = np.random.default_rng(1000 + seed)
    antpos = {i: np.array([2.0 * x, 0.0, 0.0]) for i, x in enumerate([0, 1, 4, 10, 12, 17])}
    freqs = 100e6 + 100e3 * np.arange(200)
    aps = [(i, j) for i in range(6) for j in range(i if with_autos else i + 1, 6)]
    times = T0 + 2.0 * np.arange(ntimes)
    uvd = MiniUVData(antpos, freqs, times, aps)
    vis = _point_source_vis(antpos, aps, freqs, rng)
    for t in range(ntimes):
        uvd.data_array[t * len(aps) : (t + 1) * len(aps), 0, :, 0] = vis
    return uvd


def redundant_array(seed=0):
    """Two copies of a 3-antenna line: antennas (0,1,2) and (3,4,5) with identical spacings -> redundant pairs."""
    rng = np.random.default_rng(2000 + seed)
    xs = [0.0, 2.0, 8.0, 40.0, 42.0, 48.0]
    antpos = {i: np.array([x, 0.0, 0.0]) for i, x in enumerate(xs)}
    freqs = 100e6 + 100e3 * np.arange(200)
    aps = [(i, j) for i in range(6) for j in range(i + 1, 6)]
    uvd = MiniUVData(antpos, freqs, [T0], aps)
    uvd.data_array[:, 0, :, 0] = _point_source_vis(antpos, aps, freqs, rng)
    return uvd


def project_on_dpss(uvd, comps):
    """test_calibration.py:144-156: replace each baseline by its projection on its DPSS vectors."""
    out = copy.deepcopy(uvd)
    for ap in out.get_antpairs():
        rows = out.antpair2ind(ap)
        key = ((ap,),) if ((ap,),) in comps else ((ap[::-1],),)
        basis = comps[key]
        out.data_array[rows, 0, :, 0] = (basis @ (out.data_array[rows, 0, :, 0] @ basis).T).T
    return out


def add_noise_like_eor(uvd, level_db=-50.0, seed=5):
    rng = np.random.default_rng(seed)
    out = copy.deepcopy(uvd)
    rms = np.sqrt(np.mean(np.abs(uvd.data_array) ** 2))
    amp = rms * 10 ** (level_db / 20.0)
    out.data_array = out.data_array + amp * (rng.standard_normal(out.data_array.shape) + 1j * rng.standard_normal(out.data_array.shape)) / np.sqrt(2)
    return out


def randomized_gains(uvd, seed=7, scatter=1e-2):
    rng = np.random.default_rng(seed)
    g = cal_utils.blank_uvcal_from_uvdata(uvd)
    g.gain_array = g.gain_array + scatter * rng.standard_normal(g.gain_array.shape) + 1j * scatter * rng.standard_normal(g.gain_array.shape)
    return g


def unit_weights(uvd):
    uvf = MiniUVFlag(uvd, mode="flag")
    uvf.weights_array = np.ones_like(uvf.flag_array).astype(float)
    return uvf


def dpss_vectors(uvd):
    return modeling.yield_pbl_dpss_model_comps(uvd, offset=2.0 / 0.3, min_dly=2.0 / 0.3)


def rms(x):
    return np.sqrt(np.mean(np.abs(x) ** 2.0))
