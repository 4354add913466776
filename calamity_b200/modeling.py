"""Producers of `fg_model_comps_dict` (the INPUT of the fit path).

Out of the hot path: this module only has to emit the dict format the reference's
`modeling.yield_*_comps` emit (/root/reference/calamity/modeling.py:356, 365-374, 449-473) so that the
drop-in wrappers work without `hera_filters` / `pyuvdata`, neither of which is installed in this image.

`dpss_basis` restates `hera_filters.dspec.dpss_operator(freqs, [0.], [dly], eigenval_cutoff=[cut])[0].real`
(third-party, unpinned git HEAD in the reference's setup.py:59-66; its source is not in /root/reference):
DPSS tapers of time-bandwidth product nf*df*dly, keeping the leading tapers whose concentration
eigenvalue is >= the cutoff.  hera_filters keeps `nterms = max(where(eigenvalue >= cutoff))` tapers,
i.e. it uses the INDEX of the last qualifying eigenvalue as the count; that convention is kept here.
"""
import datetime

import numpy as np
from scipy.signal import windows

from .utils import echo, PBARS


def dpss_basis(freqs, dly, eigenval_cutoff=1e-10):
    """[nfreqs, ncomp] float64 DPSS vectors for a delay half-width `dly` (seconds)."""
    freqs = np.asarray(freqs)
    nf = len(freqs)
    df = np.abs(freqs[1] - freqs[0])
    nw = nf * df * dly
    kmax = int(min(nf, max(8, np.ceil(2.0 * nw) + 40)))
    while True:
        vecs, ratios = windows.dpss(nf, nw, kmax, return_ratios=True)
        good = np.where(ratios >= eigenval_cutoff)[0]
        if len(good) == 0:
            return np.zeros((nf, 0))
        if good.max() < kmax - 1 or kmax == nf:
            break
        kmax = min(nf, kmax * 2)  # the cut was not reached inside the computed window
    nterms = int(good.max())
    return np.ascontiguousarray(vecs[:nterms].T)


def yield_dpss_model_comps_bl_grp(
    length,
    freqs,
    horizon=1.0,
    min_dly=0.0,
    offset=0.0,
    operator_cache=None,
    eigenval_cutoff=1e-10,
):
    """Per-baseline DPSS modeling vectors, Nfreqs x Ncomponents (modeling.py:255-301)."""
    if operator_cache is None:
        operator_cache = {}
    dly_ns = np.ceil(max(min_dly, length / 0.3 * horizon + offset))
    key = (len(freqs), float(freqs[0]), float(freqs[-1]), float(dly_ns), float(eigenval_cutoff))
    if key not in operator_cache:
        operator_cache[key] = dpss_basis(freqs, dly_ns / 1e9, eigenval_cutoff)
    return operator_cache[key]


def _antenna_enu(uvdata):
    """{antenna number: ENU position}; accepts pyuvdata's get_ENU_antpos or a plain `antpos` dict."""
    if hasattr(uvdata, "get_ENU_antpos"):
        pos, nums = uvdata.get_ENU_antpos(pick_data_ants=True)
        return {int(n): np.asarray(p, dtype=float) for n, p in zip(nums, pos)}
    return {int(n): np.asarray(p, dtype=float) for n, p in uvdata.antpos.items()}


def get_redundant_grps_data(uvdata, remove_redundancy=False, tol=1.0, include_autos=False):
    """Redundant baseline groups of the data, conjugation-consistent (modeling.py:10-81).

    The reference delegates to pyuvdata's `get_redundancies(use_antpos=True, include_conjugates=True)`;
    when the object offers that method it is used, otherwise the groups are found directly from the
    antenna positions with the same convention (baselines oriented to u > 0, or u == 0 and v > 0).
    Returns antpairs (empty set, as in the reference), red_grps, vec_bin_centers, lengths.
    """
    ap_data = set(uvdata.get_antpairs())
    if hasattr(uvdata, "get_redundancies") and hasattr(uvdata, "baseline_to_antnums"):
        red_grps, centers, lengths, _ = uvdata.get_redundancies(
            use_antpos=True, include_conjugates=True, include_autos=include_autos, tol=tol
        )
        red_grps = [[uvdata.baseline_to_antnums(bl) for bl in grp] for grp in red_grps]
    else:
        red_grps, centers, lengths = _redundancies_from_positions(_antenna_enu(uvdata), tol, include_autos)
    kept = [[ap for ap in grp if ap in ap_data or ap[::-1] in ap_data] for grp in red_grps]
    lengths = [ln for ln, grp in zip(lengths, kept) if len(grp) > 0]
    centers = [c for c, grp in zip(centers, kept) if len(grp) > 0]
    red_grps = [grp for grp in kept if len(grp) > 0]
    if remove_redundancy:
        flat = [([ap], c, ln) for grp, c, ln in zip(red_grps, centers, lengths) for ap in grp]
        red_grps = [f[0] for f in flat]
        centers = [f[1] for f in flat]
        lengths = [f[2] for f in flat]
    return set(), red_grps, centers, lengths


def _redundancies_from_positions(antpos, tol, include_autos):
    ants = sorted(antpos)
    vecs, pairs = [], []
    for a, i in enumerate(ants):
        for j in ants[a if include_autos else a + 1 :]:
            vec = antpos[j] - antpos[i]
            pair = (i, j)
            # conjugate so that the baseline points to u > 0 (or v > 0 on the meridian)
            if vec[0] < -tol or (abs(vec[0]) <= tol and vec[1] < -tol) or (
                abs(vec[0]) <= tol and abs(vec[1]) <= tol and vec[2] < -tol
            ):
                vec, pair = -vec, (j, i)
            vecs.append(vec)
            pairs.append(pair)
    groups, centers = [], []
    for vec, pair in zip(vecs, pairs):
        for n, c in enumerate(centers):
            if np.linalg.norm(vec - c) <= tol:
                groups[n].append(pair)
                break
        else:
            groups.append([pair])
            centers.append(np.array(vec))
    centers = [np.mean([antpos[j] - antpos[i] for (i, j) in grp], axis=0) for grp in groups]
    lengths = [float(np.linalg.norm(c)) for c in centers]
    order = np.argsort(lengths, kind="stable")
    return [groups[n] for n in order], [centers[n] for n in order], [lengths[n] for n in order]


def yield_pbl_dpss_model_comps(
    uvdata,
    horizon=1.0,
    min_dly=0.0,
    offset=0.0,
    include_autos=False,
    use_redundancy=False,
    red_tol=1.0,
    eigenval_cutoff=1e-10,
    notebook_progressbar=False,
    verbose=False,
):
    """{(tuple(red_grp),): ndarray[Nfreqs, ncomp]} per-baseline DPSS components (modeling.py:304-374)."""
    cache = {}
    _, red_grps, centers, _ = get_redundant_grps_data(
        uvdata, remove_redundancy=not use_redundancy, tol=red_tol, include_autos=include_autos
    )
    freqs = uvdata.freq_array[0]
    echo(f"{datetime.datetime.now()} Computing DPSS modeling vectors...\n", verbose=verbose)
    out = {}
    for n in PBARS[notebook_progressbar](range(len(red_grps))):
        out[(tuple(red_grps[n]),)] = yield_dpss_model_comps_bl_grp(
            freqs=freqs,
            length=np.linalg.norm(centers[n]),
            offset=offset,
            horizon=horizon,
            min_dly=min_dly,
            operator_cache=cache,
            eigenval_cutoff=eigenval_cutoff,
        )
    return out
