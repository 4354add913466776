"""Analytic sinc covariance between (baseline, frequency) samples and its leading eigenvectors
(mirror of /root/reference/calamity/simple_cov.py:7-182).  Input producer for mixed-mode fits: it runs once
per fitting group, outside the fit loop, in NumPy.  `use_tensorflow` is accepted for signature
compatibility and ignored (there is no TensorFlow here).
"""
import datetime

import numpy as np

from .utils import echo


def simple_cov_matrix(
    blvecs,
    freqs,
    ant_dly=0.0,
    horizon=1.0,
    offset=0.0,
    min_dly=0.0,
    dtype=np.float64,
    use_tensorflow=False,
    verbose=False,
):
    """(Nbls*Nfreqs) x (Nbls*Nfreqs) covariance sinc(2 max(min_dly |dnu|, horizon |du| + offset |dnu|)) *
    sinc(2 ant_dly |dnu|) with du in wavelengths-per-GHz... i.e. |b nu / c| differences, dnu in GHz and
    delays in ns; rows ordered baseline-major then frequency."""
    uvw = np.asarray(blvecs, dtype=dtype)
    nu = np.asarray(freqs, dtype=dtype)
    nbls, nf = len(uvw), len(nu)
    # coordinates of every (baseline, channel) sample in units of cycles per Hz * Hz = wavelengths
    coords = (uvw[:, None, :] * (nu[None, :, None] / 3e8)).reshape(nbls * nf, 3)
    sep2 = np.zeros((nbls * nf, nbls * nf), dtype=dtype)
    for axis in range(3):
        col = coords[:, axis]
        sep2 += np.abs(col[:, None] - col[None, :]) ** 2.0
    sep = np.sqrt(sep2) * horizon
    del sep2
    nu_all = np.tile(nu, nbls)
    dnu = np.abs(nu_all[:, None] - nu_all[None, :]) / 1e9
    sep += dnu * offset
    cov = np.sinc(2 * np.maximum(min_dly * dnu, sep))
    del sep
    cov = cov * np.sinc(2 * dnu * ant_dly)
    return cov


def yield_simple_multi_baseline_model_comps(
    blvecs,
    freqs,
    ant_dly=0.0,
    horizon=1.0,
    offset=0.0,
    min_dly=0.0,
    dtype=np.float64,
    verbose=False,
    use_tensorflow=False,
    eigenval_cutoff=1e-10,
):
    """Eigenvectors of the covariance whose eigenvalue is at least `eigenval_cutoff` of the largest, strongest
    first: (Nbls*Nfreqs) x Ncomponents."""
    cov = simple_cov_matrix(blvecs, freqs, ant_dly=ant_dly, horizon=horizon, offset=offset, min_dly=min_dly,
                            dtype=dtype, verbose=verbose)
    echo(f"{datetime.datetime.now()} Deriving modeling components with eigenvalue decomposition...\n", verbose=verbose)
    evals, evecs = np.linalg.eigh(cov)
    keep = evals / evals[-1] >= eigenval_cutoff
    echo(f"{datetime.datetime.now()} Finished spectral decomposition. Using {np.count_nonzero(keep)} of {len(keep)} "
         f"eigenvectors to model foregrounds...\n", verbose=verbose)
    return evecs[:, keep][:, ::-1]
