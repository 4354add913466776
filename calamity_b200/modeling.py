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
    """Fallback for objects without pyuvdata's `get_redundancies`: same grouping via the stand-in's finder."""
    from .uvstandins import MiniUVData

    probe = MiniUVData.__new__(MiniUVData)
    probe.antpos = antpos
    bl_groups, centers, lengths, _ = probe.get_redundancies(tol=tol, use_antpos=True, include_conjugates=True,
                                                            include_autos=include_autos)
    return [[MiniUVData.baseline_to_antnums(bl) for bl in grp] for grp in bl_groups], centers, lengths


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


def get_uv_overlapping_grps_conjugated(
    uvdata,
    red_tol=1.0,
    include_autos=False,
    red_tol_freq=0.5,
    n_angle_bins=200,
    notebook_progressbar=False,
    require_exact_angle_match=True,
    angle_match_tol=1e-3,
):
    """Fitting groups = sets of redundant groups whose uv tracks come within `red_tol_freq` wavelengths of each
    other somewhere in the band (modeling.py:84-252).

    Returns fitting_grps (list of lists of redundant-group tuples), fitting_vec_centers, connections, grp_labels.
    The traversal order (angle bins, (angle, length) sort, set iteration over connections) follows the reference
    step for step because it defines the group ordering the golden vector of test_modeling.py:24-32 pins.
    """
    _, red_grps, centers, _ = get_redundant_grps_data(uvdata, include_autos=include_autos, tol=red_tol,
                                                      remove_redundancy=False)
    freqs = np.asarray(uvdata.freq_array[0])
    fmin, fmax = uvdata.freq_array.min(), uvdata.freq_array.max()
    center_of = {}
    connections = {}
    bins = {n: [] for n in range(n_angle_bins)}
    dangle = np.pi / n_angle_bins
    for gnum, (grp, vec) in enumerate(zip(red_grps, centers)):
        center_of[tuple(grp)] = vec
        if np.abs(vec[0]) > 0.0:
            which = int(np.min([np.round((np.arctan(vec[1] / vec[0]) + np.pi / 2) / dangle), n_angle_bins - 2]))
        else:
            which = n_angle_bins - 1
        bins[which].append(gnum)

    def track_overlap(lo0, hi0, lo1, hi1):
        return (lo0 > lo1 and lo0 < hi1) or (lo1 > lo0 and lo1 < hi0)

    for which in PBARS[notebook_progressbar](range(n_angle_bins)):
        members = bins[which]
        for a, g0 in enumerate(members):
            vec0 = centers[g0]
            key0 = tuple(red_grps[g0])
            if key0 not in connections:
                connections[key0] = set({})
                center_of[key0] = vec0
            for g1 in members[a + 1 :]:
                vec1 = centers[g1]
                len0, len1 = np.linalg.norm(vec0), np.linalg.norm(vec1)
                if not track_overlap(fmin * len0 / 3e8, fmax * len0 / 3e8, fmin * len1 / 3e8, fmax * len1 / 3e8):
                    continue
                if require_exact_angle_match and not (
                    np.abs(np.arctan(vec0[1] / vec0[0]) - np.arctan(vec1[1] / vec1[0])) <= angle_match_tol
                ):
                    continue
                u0, v0 = vec0[0] * freqs / 3e8, vec0[1] * freqs / 3e8
                u1, v1 = vec1[0] * freqs / 3e8, vec1[1] * freqs / 3e8
                du_m, dv_m = u0[None, :] - u1[:, None], v0[None, :] - v1[:, None]
                du_p, dv_p = u0[None, :] + u1[:, None], v0[None, :] + v1[:, None]
                if np.any(np.sqrt(np.abs(du_m) ** 2.0 + dv_m ** 2.0) <= red_tol_freq):
                    key1 = tuple(red_grps[g1])
                elif np.any(np.sqrt(np.abs(du_p) ** 2.0 + dv_p ** 2.0) <= red_tol_freq):
                    # the second group overlaps once conjugated: flip it for everything that follows
                    red_grps[g1] = [ap[::-1] for ap in red_grps[g1]]
                    centers[g1] = [-c for c in centers[g1]]
                    key1 = tuple(red_grps[g1])
                else:
                    continue
                connections[key0].add(key1)
                if key1 not in connections:
                    connections[key1] = set({})
                    center_of[key1] = vec1
                connections[key1].add(key0)

    keys = [k for k in center_of if k in connections]
    lengths = [np.linalg.norm(center_of[k]) for k in keys]
    angles = [np.arccos(center_of[k][0] / ln) for k, ln in zip(keys, lengths)]
    order = sorted(range(len(keys)), key=lambda n: (angles[n], lengths[n]))
    fitting = {}
    grp_labels = {}
    for n in PBARS[notebook_progressbar](order):
        grp = keys[n]
        if grp not in grp_labels:
            fitting[grp] = [grp]
            grp_labels[grp] = grp
            parent = grp
        else:
            parent = grp_labels[grp]
        for other in connections[grp]:
            if other not in grp_labels:
                fitting[parent].append(other)
                grp_labels[other] = parent
    fitting_grps = list(fitting.values())
    fitting_vec_centers = [[center_of[grp] for grp in fit] for fit in fitting_grps]
    return fitting_grps, fitting_vec_centers, connections, grp_labels


def yield_mixed_comps(
    fitting_grps,
    fitting_blvecs,
    freqs,
    eigenval_cutoff=1e-10,
    ant_dly=0.0,
    horizon=1.0,
    offset=0.0,
    min_dly=0.0,
    verbose=False,
    dtype=np.float64,
    notebook_progressbar=False,
    use_tensorflow=False,
    grp_size_threshold=5,
):
    """DPSS vectors for fitting groups of at most `grp_size_threshold` redundant groups (one dict entry per
    redundant group, delay offset = ant_dly, as upstream), joint covariance eigenvectors keyed by the whole
    fitting group otherwise (modeling.py:377-474)."""
    from . import simple_cov

    cache = {}
    out = {}
    for n in PBARS[notebook_progressbar](range(len(fitting_grps))):
        fit = tuple(fitting_grps[n]) if isinstance(fitting_grps[n], list) else fitting_grps[n]
        vecs = fitting_blvecs[n]
        if len(fit) <= grp_size_threshold:
            for red, ln in zip(fit, np.linalg.norm(vecs, axis=1)):
                out[(red,)] = yield_dpss_model_comps_bl_grp(freqs=freqs, length=ln, offset=ant_dly, horizon=horizon,
                                                            min_dly=min_dly, operator_cache=cache,
                                                            eigenval_cutoff=eigenval_cutoff)
        else:
            out[fit] = simple_cov.yield_simple_multi_baseline_model_comps(
                blvecs=vecs, ant_dly=ant_dly, offset=offset, min_dly=min_dly, horizon=horizon, dtype=dtype, freqs=freqs,
                eigenval_cutoff=eigenval_cutoff, use_tensorflow=use_tensorflow, verbose=verbose)
    return out
