"""Drop-in replacement for `calamity.calibration` whose fit loop runs on a B200.

Same names, signatures, defaults and return values as /root/reference/calamity/calibration.py; the
TensorFlow tensors, GradientTape gradient and Keras optimizer step are replaced by calls into the
C-ABI CUDA library (include/calamity_b200.h) through `fitter.FitPlan`.  There is no CPU path for the
arithmetic of the fit: without the built library and a CUDA device the fit functions raise.

"Tensors" returned by the tensorize_* helpers are `DeviceLikeArray`s: NumPy arrays with the `.numpy()`
method the reference's callers use on tf.Tensor objects (so `.dtype == np.float64`, `.shape`, slicing
and `.numpy()` all behave as in the reference's tests).
"""
import argparse
import copy
import datetime
import json
import os
import time as _time

import numpy as np

from . import cal_utils
from . import modeling
from . import utils
from .fitter import FitPlan, REFERENCE_OPTIMIZERS
from .layout import RaggedLayout
from .utils import echo
from .utils import PBARS

try:  # pragma: no cover - pyuvdata is not installed in the build image
    from pyuvdata import UVData, UVCal, UVFlag
    from pyuvdata import utils as uvutils

    _polstr2num = uvutils.polstr2num
except Exception:
    from .uvstandins import MiniUVData as UVData, MiniUVCal as UVCal, MiniUVFlag as UVFlag
    from .uvstandins import polstr2num as _polstr2num

# names accepted by the reference's OPTIMIZERS table (calibration.py:17-27)
OPTIMIZERS = {name: name for name in REFERENCE_OPTIMIZERS}


class DeviceLikeArray(np.ndarray):
    """ndarray with the small tf.Tensor surface the reference's callers rely on."""

    def numpy(self):
        return np.asarray(self)

    def value(self):
        return self


def _as_tensor(arr, dtype=None):
    return np.ascontiguousarray(arr, dtype=dtype).view(DeviceLikeArray)


def _device_index():
    return int(os.environ.get("CALAMITY_B200_DEVICE", "0"))


def _device_indices(nunits, sequential):
    """CUDA devices the driver spreads its independent (polarization, time) integrations over (calibration.py:1160-1167:
    the units share the basis and nothing else, unless init_guesses_from_previous_time_step chains them, calibration.py:1210).
    CALAMITY_B200_DEVICES = "all" (default) | "0,2,3"; CALAMITY_B200_DEVICE (or the file driver's gpu_index) pins one
    device, like the reference's `gpu_index` (calibration.py:1741-1752).  Never more devices than units."""
    if sequential or nunits < 2 or "CALAMITY_B200_DEVICE" in os.environ:
        return [_device_index()]
    spec = os.environ.get("CALAMITY_B200_DEVICES", "all").strip()
    if spec == "all":
        import ctypes

        from . import _native as nat

        n = ctypes.c_int32(0)
        nat.check(nat.load().calb2_device_count(ctypes.byref(n)))
        devs = list(range(max(1, n.value)))
    else:
        devs = [int(x) for x in spec.split(",") if x.strip() != ""]
    return devs[: max(1, min(len(devs), nunits))] or [0]


# ----------------------------------------------------------------------------------------------------
# marshalling (calibration.py:30-444)
# ----------------------------------------------------------------------------------------------------
def chunk_fg_comp_dict_by_nbls(fg_model_comps_dict, use_redundancy=False, grp_size_threshold=5):
    """Group fitting groups into chunks keyed (baselines per group, widest basis) -- calibration.py:30-101.

    Without `use_redundancy`, a fitting group whose redundant sub-groups all have the same length and that
    has fewer than `grp_size_threshold` of them is replaced by one group per member baseline; the
    replacements share the original component array and go to the END of the ordering, which is what fixes
    the group / baseline indices of everything downstream.
    """
    ordered = dict(copy.deepcopy(fg_model_comps_dict))
    if not use_redundancy:
        for grp in list(ordered):
            counts = np.asarray([len(red) for red in grp])
            if len(counts) < grp_size_threshold and np.allclose(counts, np.mean(counts)):
                shared = ordered.pop(grp)
                for member in range(int(counts[0])):
                    ordered[tuple((red[member],) for red in grp)] = shared
    by_size, width = {}, {}
    for grp, comps in ordered.items():
        nbl = int(sum(len(red) for red in grp))
        by_size.setdefault(nbl, []).append(grp)
        width[nbl] = max(width.get(nbl, 0), comps.shape[1])
    return {(nbl, width[nbl]): {grp: ordered[grp] for grp in grps} for nbl, grps in by_size.items()}


def _layout_from_dict(fg_model_comps_dict, ants_map, nfreqs, use_redundancy, grp_size_threshold, nants=None,
                      dtype=np.float32):
    chunked = chunk_fg_comp_dict_by_nbls(fg_model_comps_dict, use_redundancy=use_redundancy,
                                         grp_size_threshold=grp_size_threshold)
    return RaggedLayout.from_chunked_dict(chunked, ants_map, nfreqs, nants=nants, dtype=dtype)


def _fit_dtype(dtype):
    """np.float32 -> fused sm_100a kernels, np.float64 -> the generic device path (calibration.py:974, 1795)."""
    dt = np.dtype(dtype)
    if dt not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError(f"dtype must be np.float32 or np.float64, not {dtype!r}")
    return dt


def tensorize_fg_model_comps_dict(
    fg_model_comps_dict,
    ants_map,
    nfreqs,
    use_redundancy=False,
    dtype=np.float32,
    notebook_progressbar=False,
    verbose=False,
    grp_size_threshold=5,
):
    """Dense zero-padded (nvecs, ngrps, nbls, nfreqs) tensors, one per chunk, and
    corr_inds[chunk][group][baseline] = (i, j) -- calibration.py:104-190.

    Kept for API compatibility: `calibrate_and_model_tensor` itself never builds these dense tensors, it
    hands the ragged description straight to the device.
    """
    echo(f"{datetime.datetime.now()} Computing foreground components matrices...\n", verbose=verbose)
    lay = _layout_from_dict(fg_model_comps_dict, ants_map, nfreqs, use_redundancy, grp_size_threshold)
    # the reference fills float64 then converts: identical to rounding each entry once
    dense = []
    chunked = chunk_fg_comp_dict_by_nbls(fg_model_comps_dict, use_redundancy=use_redundancy,
                                         grp_size_threshold=grp_size_threshold)
    for (nbls, nvecs), grp_dict in chunked.items():
        block = np.zeros((nvecs, len(grp_dict), nbls, nfreqs))
        for g, (grp, comps) in enumerate(grp_dict.items()):
            b = 0
            for rnum, red in enumerate(grp):
                rows = comps[rnum * nfreqs : (rnum + 1) * nfreqs].T
                for _ in red:
                    block[: comps.shape[1], g, b] = rows
                    b += 1
        dense.append(_as_tensor(block, dtype=dtype))
    return dense, lay.corr_inds()


def _resolve_baseline(uvdata, ants_map_inv, i, j, polarization, time):
    """Row of `uvdata` holding antenna-index pair (i, j) at `time`, whether it is stored conjugated, and
    the polarization index -- the lookups of calibration.py:260-272."""
    key = (ants_map_inv[i], ants_map_inv[j], polarization)
    fwd, rev, pol_ind = uvdata._key2inds(key)
    if len(fwd) > 0:
        rows, conj, pind = fwd, False, pol_ind[0]
    else:
        rows, conj, pind = rev, True, pol_ind[1]
    pind = int(np.asarray(pind).ravel()[0])
    rows = np.asarray(rows)
    hit = np.where(np.isclose(uvdata.time_array[rows], time, rtol=0.0, atol=1e-7))[0][0]  # IndexError if absent
    return int(rows[hit]), conj, pind


def _rows_at_time(obj, time):
    """{(ant1, ant2): row} of the baseline-time rows of a UVData / UVFlag-like object at `time`."""
    rows = np.nonzero(np.isclose(np.asarray(obj.time_array), time, rtol=0.0, atol=1e-7))[0]
    if hasattr(obj, "ant_1_array"):
        a1, a2 = np.asarray(obj.ant_1_array)[rows], np.asarray(obj.ant_2_array)[rows]
        return {(int(a), int(b)): int(r) for a, b, r in zip(a1.tolist(), a2.tolist(), rows.tolist())}
    out = {}  # objects that only expose antpair2ind (UVFlag)
    at_time = set(rows.tolist())
    for ap in obj.get_antpairs():
        for r in np.asarray(obj.antpair2ind(*ap)).tolist():
            if r in at_time:
                out[(int(ap[0]), int(ap[1]))] = int(r)
    return out


def _resolve_baselines(uvdata, ants_map_inv, bl_pairs, polarization, time):
    """Vectorised form of the per-baseline lookups of calibration.py:260-272: for every antenna-index pair the row of
    `uvdata` at `time`, whether the baseline is stored the other way round, and the polarization index to read.
    A baseline counts as stored "forward" when its (ant1, ant2) ordering exists at ANY time (that is what
    `_key2inds` looks at); a baseline absent at `time` raises IndexError, as upstream."""
    stored = set((int(a), int(b)) for a, b in uvdata.get_antpairs())
    at_time = _rows_at_time(uvdata, time)
    rows = np.empty(len(bl_pairs), dtype=np.int64)
    conj = np.zeros(len(bl_pairs), dtype=bool)
    first_fwd = first_rev = None
    for n, (i, j) in enumerate(bl_pairs):
        ap = (int(ants_map_inv[i]), int(ants_map_inv[j]))
        if ap in stored:
            if first_fwd is None:
                first_fwd = ap
        elif ap[::-1] in stored:
            conj[n] = True
            ap = ap[::-1]
            if first_rev is None:
                first_rev = ap[::-1]
        else:
            raise KeyError(f"antenna pair {ap} not found in data")
        if ap not in at_time:
            raise IndexError("index 0 is out of bounds for axis 0 with size 0")  # np.where(...)[0][0] upstream
        rows[n] = at_time[ap]
    pind_f = pind_r = 0
    if first_fwd is not None:
        pind_f = int(np.asarray(uvdata._key2inds(first_fwd + (polarization,))[2][0]).ravel()[0])
    if first_rev is not None:
        pind_r = int(np.asarray(uvdata._key2inds(first_rev + (polarization,))[2][1]).ravel()[0])
    return rows, conj, np.where(conj, pind_r, pind_f)


def _tensorize_data_flat(uvdata, bl_pairs, ants_map, polarization, time, data_scale_factor, weights,
                         nsamples_in_weights, dtype):
    """Per-baseline rows [nbls, nfreqs] of scaled data and normalised weights in `bl_pairs` order (the arithmetic of
    calibration.py:257-303, vectorised over baselines: the reference's Python loop is O(Nbls) per integration)."""
    inv = {v: k for k, v in ants_map.items()}
    rows, conj, pind = _resolve_baselines(uvdata, inv, bl_pairs, polarization, time)
    if len(pind) and np.all(pind == pind[0]):
        pind = int(pind[0])  # one polarization index for all baselines (always, for xx / yy): plain row gathers
    vis = uvdata.data_array[rows, 0, :, pind] / data_scale_factor  # fancy indexing: a copy, [nbls, nfreqs]
    d_r = np.ascontiguousarray(vis.real, dtype=dtype)
    d_i = np.ascontiguousarray(np.where(conj[:, None], -vis.imag, vis.imag), dtype=dtype)
    unflagged = ~uvdata.flag_array[rows, 0, :, pind]
    if weights is None:
        w = unflagged.astype(dtype)
    else:
        wat = _rows_at_time(weights, time)
        wrows = np.empty(len(bl_pairs), dtype=np.int64)
        for n, (i, j) in enumerate(bl_pairs):
            ap = (int(inv[i]), int(inv[j]))
            if ap not in wat and ap[::-1] in wat:
                ap = ap[::-1]
            if ap not in wat:
                raise IndexError("index 0 is out of bounds for axis 0 with size 0")
            wrows[n] = wat[ap]
        wpol = np.where(weights.polarization_array == _polstr2num(polarization, x_orientation=weights.x_orientation))[0][0]
        w = weights.weights_array[wrows, 0, :, wpol].astype(dtype) * unflagged
    if nsamples_in_weights:
        w = (w * uvdata.nsample_array[rows, 0, :, pind]).astype(dtype)
    wsum = 0.0
    for s in np.sum(w, axis=1).tolist():  # same accumulation order as the reference's running python-float sum
        wsum += s
    w = (w / wsum).astype(dtype)
    return d_r, d_i, np.ascontiguousarray(w)


def tensorize_data(
    uvdata,
    corr_inds,
    ants_map,
    polarization,
    time,
    data_scale_factor=1.0,
    weights=None,
    nsamples_in_weights=False,
    dtype=np.float32,
):
    """Data / weights of one (polarization, time) as per-chunk (ngrps, nbls, nfreqs) tensors --
    calibration.py:193-310.  Weights are normalised to unit sum over everything in corr_inds."""
    pairs = [pair for chunk in corr_inds for grp in chunk for pair in grp]
    d_r, d_i, w = _tensorize_data_flat(uvdata, pairs, ants_map, polarization, time, data_scale_factor, weights,
                                       nsamples_in_weights, dtype)
    out_r, out_i, out_w, pos = [], [], [], 0
    for chunk in corr_inds:
        ngrps = len(chunk)
        nbls = len(chunk[0]) if ngrps else 0
        n = ngrps * nbls
        shape = (ngrps, nbls, uvdata.Nfreqs)
        out_r.append(_as_tensor(d_r[pos : pos + n].reshape(shape), dtype=dtype))
        out_i.append(_as_tensor(d_i[pos : pos + n].reshape(shape), dtype=dtype))
        out_w.append(_as_tensor(w[pos : pos + n].reshape(shape), dtype=dtype))
        pos += n
    return out_r, out_i, out_w


def renormalize(uvdata_reference_model, uvdata_deconv, gains, polarization, time, additional_flags=None):
    """Fix the overall amplitude degeneracy after a 'post_hoc' fit (calibration.py:313-366): scale the model
    by rms |reference / model| over unflagged samples and the gains by that factor ** -1/2.  In place."""
    pnum = np.where(
        uvdata_deconv.polarization_array == _polstr2num(polarization, x_orientation=uvdata_deconv.x_orientation)
    )[0][0]
    rows = np.isclose(uvdata_deconv.time_array, time, atol=1e-7, rtol=0.0)
    good = ~uvdata_deconv.flag_array[rows, :, :, pnum] & ~uvdata_reference_model.flag_array[rows, :, :, pnum]
    if additional_flags is not None:
        good = good & ~additional_flags[rows, :, :, pnum]
    with np.errstate(divide="ignore", invalid="ignore"):
        ratio = uvdata_reference_model.data_array[rows, :, :, pnum][good] / uvdata_deconv.data_array[rows, :, :, pnum][good]
    ratio[~np.isfinite(ratio)] = np.nan
    scale = np.sqrt(np.nanmean(np.abs(ratio) ** 2.0))  # amplitude only; the phase term is unused upstream too
    uvdata_deconv.data_array[rows, :, :, pnum] *= scale
    jnum = np.where(gains.jones_array == _polstr2num(polarization, x_orientation=uvdata_deconv.x_orientation))[0][0]
    tnum = np.where(np.isclose(gains.time_array, time, atol=1e-7, rtol=0.0))[0][0]
    gains.gain_array[:, :, :, tnum, jnum] *= scale ** -0.5


def tensorize_gains(uvcal, polarization, time, dtype=np.float32):
    """Real and imaginary gain tables (Nants, Nfreqs) of one Jones term and time -- calibration.py:369-399."""
    jnum = np.where(uvcal.jones_array == _polstr2num(polarization, x_orientation=uvcal.x_orientation))[0][0]
    tnum = np.where(np.isclose(uvcal.time_array, time, atol=1e-7, rtol=0.0))[0][0]
    g = uvcal.gain_array[:, 0, :, tnum, jnum].squeeze()
    return _as_tensor(g.real, dtype=dtype), _as_tensor(g.imag, dtype=dtype)


def _contraction_dtype(arrays):
    """float64 when every tensor involved is float64 (the reference then contracts in float64), else float32."""
    return np.dtype(np.float64) if all(np.asarray(a).dtype == np.float64 for a in arrays) else np.dtype(np.float32)


def _nants_from(corr_inds, g_r=None):
    if g_r is not None:
        return int(np.shape(g_r)[0])
    return 1 + max(max(p) for chunk in corr_inds for grp in chunk for p in grp)


def yield_fg_model_array(
    nants,
    nfreqs,
    fg_model_comps,
    fg_coeffs,
    corr_inds,
):
    """(nants, nants, nfreqs) float64 cube of sum_k coeff_k * comp_k, cell (i, j) only -- calibration.py:402-444.
    The contraction runs on the device."""
    comps = [np.asarray(c) for c in fg_model_comps]
    coeffs = [np.asarray(c) for c in fg_coeffs]
    # float64 tensors are contracted in float64, as the reference does (the generic device path), everything else in float32
    fdt = _contraction_dtype(comps + coeffs)
    lay = RaggedLayout.from_dense(comps, corr_inds, nants, dtype=fdt)
    flat = lay.flatten_coeffs(coeffs)
    with FitPlan(lay, device=_device_index()) as plan:
        zeros = np.zeros((lay.nbls, lay.nfreqs), dtype=fdt)
        plan.set_integration(zeros, zeros, zeros)
        plan.set_gains(np.ones((nants, nfreqs), dtype=fdt), np.zeros((nants, nfreqs), dtype=fdt))
        plan.set_coeffs(flat, np.zeros_like(flat))
        vis, _ = plan.get_model()
    cube = np.zeros((nants, nants, nfreqs))
    cube[lay.bl_ant0, lay.bl_ant1] = vis
    return cube


# ----------------------------------------------------------------------------------------------------
# the fit (calibration.py:447-738)
# ----------------------------------------------------------------------------------------------------
def _profile_dump(profile_log_dir, payload):
    os.makedirs(profile_log_dir, exist_ok=True)
    name = os.path.join(profile_log_dir, f"calamity_b200_profile_{int(_time.time() * 1e3)}.json")
    with open(name, "w") as f:
        json.dump(payload, f, indent=1)


def _run_fit(plan, lay, g_r, g_i, fg_r, fg_i, use_min, tol, maxsteps, optimizer, freeze_model, verbose,
             n_profile_steps, profile_log_dir, model_regularization, priors, dtype, opt_kwargs, graph_mode=False):
    """Shared tail of fit_gains_and_foregrounds / calibrate_and_model_tensor: run the loop on a loaded plan
    and translate the results back to the reference's shapes.  `graph_mode=True` (the reference's "pre-compile the
    computational graph", calibration.py:670-679) forces CUDA-graph replay of the step loop; otherwise the library
    decides (graphs for small, launch-bound problems)."""
    if optimizer not in OPTIMIZERS:
        raise KeyError(optimizer)  # calibration.py:571
    dtype = _fit_dtype(dtype)
    echo(f"{datetime.datetime.now()} Performing gradient descent on {np.prod(np.shape(g_r))} complex gain parameters...",
         verbose=verbose)
    if not freeze_model:
        echo(f"Performing gradient descent on total of {lay.ncoef} complex foreground parameters", verbose=verbose)
    if n_profile_steps > 0:
        echo(f"{datetime.datetime.now()} Profiling with {n_profile_steps}. And writing output to {profile_log_dir}...")
    pr, pi = priors
    hist, res = plan.fit(optimizer=optimizer, maxsteps=maxsteps, tol=tol, use_min=use_min, freeze_model=freeze_model,
                         model_regularization=model_regularization, prior_r_sum=pr, prior_i_sum=pi,
                         n_profile_steps=n_profile_steps, use_graph=True if graph_mode else None, **opt_kwargs)
    if n_profile_steps > 0:
        _profile_dump(profile_log_dir, dict(res, n_profile_steps=n_profile_steps, note="CUDA-event timings of the step "
                                            "loop; the profiled steps are real optimizer steps, as in the reference"))
    if not use_min and len(hist) == 0:
        raise IndexError("list index out of range")  # calibration.py:723 when maxsteps == 0
    if use_min and freeze_model:
        raise UnboundLocalError("local variable 'fg_r_opt' referenced before assignment")  # calibration.py:738
    out_gr, out_gi = plan.get_gains()
    c_r, c_i = plan.get_coeffs()
    fit_history = {"loss": [dtype.type(x) for x in hist]}
    echo(f"{datetime.datetime.now()} Finished Gradient Descent. MSE of {res['final_loss']:.2e}...\n", verbose=verbose)
    return out_gr, out_gi, c_r, c_i, fit_history


def fit_gains_and_foregrounds(
    g_r,
    g_i,
    fg_r,
    fg_i,
    data_r,
    data_i,
    wgts,
    fg_comps,
    corr_inds,
    use_min=False,
    tol=1e-14,
    maxsteps=10000,
    optimizer="Adamax",
    freeze_model=False,
    verbose=False,
    notebook_progressbar=False,
    dtype=np.float32,
    graph_mode=False,
    n_profile_steps=0,
    profile_log_dir="./logdir",
    sky_model_r=None,
    sky_model_i=None,
    model_regularization=None,
    graph_args_dict=None,
    **opt_kwargs,
):
    """Gradient-descent fit of gains and foreground coefficients -- calibration.py:447-738.

    Same contract as the reference: one unrecorded warm-up step, then up to `maxsteps` recorded steps whose
    PRE-update loss goes into fit_history['loss']; stop when two consecutive recorded losses differ by less
    than `tol`; `use_min` returns the post-update parameters of the step with the smallest recorded loss.
    `graph_mode=True` replays the step loop as a CUDA graph (the analogue of the reference's tf.function mode);
    `graph_args_dict` holds tf.function options and has no effect here.
    """
    echo(f"Using {str(dtype)} precision.")
    echo(f"{datetime.datetime.now()} Provided the following opt_kwargs")
    for k in opt_kwargs:
        echo(f"{k}: {opt_kwargs[k]}")
    if optimizer not in OPTIMIZERS:
        raise KeyError(optimizer)
    nants = int(np.shape(g_r)[0])
    lay = RaggedLayout.from_dense([np.asarray(c) for c in fg_comps], corr_inds, nants, dtype=_fit_dtype(dtype))
    with FitPlan(lay, device=_device_index()) as plan:
        w_flat = lay.flatten_data(wgts)
        plan.set_integration(lay.flatten_data(data_r), lay.flatten_data(data_i), w_flat)
        plan.set_gains(np.asarray(g_r), np.asarray(g_i))
        plan.set_coeffs(lay.flatten_coeffs(fg_r), lay.flatten_coeffs(fg_i))
        priors = (0.0, 0.0)
        if model_regularization == "sum":
            priors = plan.prior_sums(lay.flatten_data(sky_model_r), lay.flatten_data(sky_model_i))
        out_gr, out_gi, c_r, c_i, fit_history = _run_fit(
            plan, lay, g_r, g_i, fg_r, fg_i, use_min, tol, maxsteps, optimizer, freeze_model, verbose, n_profile_steps,
            profile_log_dir, model_regularization, priors, dtype, opt_kwargs, graph_mode=graph_mode)
    if freeze_model:
        fg_r_opt, fg_i_opt = fg_r, fg_i  # calibration.py:730-732: handed back untouched
    else:
        fg_r_opt = [_as_tensor(t) for t in lay.unflatten_coeffs(c_r, template=fg_r, dtype=lay.dtype)]
        fg_i_opt = [_as_tensor(t) for t in lay.unflatten_coeffs(c_i, template=fg_i, dtype=lay.dtype)]
    return _as_tensor(out_gr), _as_tensor(out_gi), fg_r_opt, fg_i_opt, fit_history


def insert_model_into_uvdata_tensor(
    uvdata,
    time,
    polarization,
    ants_map,
    red_grps,
    model_r,
    model_i,
    scale_factor=1.0,
):
    """Write cube cells (i, j) back into the rows of `uvdata` at `time` (conjugating baselines stored the other
    way round) times `scale_factor` -- calibration.py:741-795.  In place."""
    stored = set((int(a), int(b)) for a, b in uvdata.get_antpairs())
    pnum = np.where(uvdata.polarization_array == _polstr2num(polarization, x_orientation=uvdata.x_orientation))[0][0]
    at_time = _rows_at_time(uvdata, time)
    rows, ii, jj, sign = [], [], [], []
    for red in red_grps:
        for ap in red:
            key = (int(ap[0]), int(ap[1]))
            fwd = key in stored
            if not fwd:
                key = key[::-1]
            if key not in at_time:
                raise IndexError("index 0 is out of bounds for axis 0 with size 0")
            rows.append(at_time[key])
            ii.append(ants_map[ap[0]])
            jj.append(ants_map[ap[1]])
            sign.append(1.0 if fwd else -1.0)
    if rows:
        ii, jj = np.asarray(ii), np.asarray(jj)
        vis = np.asarray(model_r)[ii, jj] + 1j * np.asarray(sign)[:, None] * np.asarray(model_i)[ii, jj]
        uvdata.data_array[np.asarray(rows), 0, :, pnum] = vis * scale_factor


def insert_gains_into_uvcal(uvcal, time, polarization, gains_re, gains_im):
    """Write (Nants, Nfreqs) gain tables into `uvcal` at (time, polarization) -- calibration.py:798-825."""
    jnum = np.where(uvcal.jones_array == _polstr2num(polarization, x_orientation=uvcal.x_orientation))[0][0]
    tnum = np.where(np.isclose(uvcal.time_array, time, atol=1e-7, rtol=0.0))[0][0]
    uvcal.gain_array[:, 0, :, tnum, jnum] = np.asarray(gains_re)[: uvcal.Nants_data] + 1j * np.asarray(gains_im)[: uvcal.Nants_data]


def tensorize_fg_coeffs(
    data,
    wgts,
    fg_model_comps,
    notebook_progressbar=False,
    verbose=False,
):
    """Least-squares starting coefficients, one (nvecs, ngrps, 1, 1) tensor per chunk -- calibration.py:828-913:
    per group, unweighted least squares of data * (wgts != 0) on the basis vectors that precede the first
    all-zero row, zero-padded back to nvecs.  Solved on the device (normal equations + Cholesky)."""
    echo(f"{datetime.datetime.now()} Computing initial foreground coefficient guesses using linear-leastsq...\n",
         verbose=verbose)
    comps = [np.asarray(c) for c in fg_model_comps]
    nants = 1  # antenna indices are irrelevant for the projection: give every baseline the pair (0, 0)
    corr = [[[(0, 0)] * c.shape[2] for _ in range(c.shape[1])] for c in comps]
    # the reference cuts each group's basis at its FIRST all-zero row (calibration.py:886-892)
    trimmed = []
    for c in comps:
        c = np.array(c, copy=True)
        nvecs, ngrps = c.shape[:2]
        for g in range(ngrps):
            empty = np.where(np.all(np.isclose(c[:, g].reshape(nvecs, -1), 0.0), axis=1))[0]
            if len(empty) > 0:
                c[int(empty.min()) :, g] = 0.0
        trimmed.append(c)
    dtype = np.asarray(data[0]).dtype
    lay = RaggedLayout.from_dense(trimmed, corr, nants, dtype=np.float64 if dtype == np.float64 else np.float32)
    with FitPlan(lay, device=_device_index()) as plan:
        flat = lay.flatten_data(data)
        plan.set_integration(flat, flat, lay.flatten_data(wgts))
        plan.init_coeffs(flat, np.zeros_like(flat))
        c_r, _ = plan.get_coeffs()
    echo(f"{datetime.datetime.now()} Finished initial foreground coefficient guesses...\n", verbose=verbose)
    return [_as_tensor(t, dtype=dtype) for t in lay.unflatten_coeffs(c_r, dtype=dtype)]


def get_auto_weights(uvdata, delay_extent=25.0):
    """Inverse-variance weights from DPSS-smoothed autocorrelations (calibration.py:916-960).  Host-side helper,
    outside the fit path."""
    comps = modeling.yield_dpss_model_comps_bl_grp(0.0, uvdata.freq_array[0], offset=delay_extent)
    out = UVFlag(uvdata, mode="flag")
    out.weights_array = np.zeros(uvdata.data_array.shape)
    smooth = {}
    keys = uvdata.get_antpairpols()
    for key in keys:
        if key[0] != key[1]:
            continue
        fits = []
        for spec, ok in zip(uvdata.get_data(key), ~uvdata.get_flags(key)):
            sol, *_ = np.linalg.lstsq(comps[ok], spec[ok].real, rcond=None)
            fits.append(comps @ sol)
        smooth[key] = np.atleast_2d(np.asarray(fits))
    for key in keys:
        wgt = 1.0 / (smooth[key[0], key[0], key[-1]] * smooth[key[1], key[1], key[-1]])
        wgt = wgt * ~uvdata.get_flags(key)
        rows = out.antpair2ind(*key[:2])
        pnum = np.where(out.polarization_array == _polstr2num(key[-1], x_orientation=out.x_orientation))[0][0]
        out.weights_array[rows, 0, :, pnum] = wgt
    return out


# ----------------------------------------------------------------------------------------------------
# driver (calibration.py:963-1331)
# ----------------------------------------------------------------------------------------------------
def calibrate_and_model_tensor(
    uvdata,
    fg_model_comps_dict,
    gains=None,
    freeze_model=False,
    optimizer="Adamax",
    tol=1e-14,
    maxsteps=10000,
    include_autos=False,
    verbose=False,
    sky_model=None,
    dtype=np.float32,
    use_min=False,
    use_redundancy=False,
    notebook_progressbar=False,
    correct_resid=False,
    correct_model=True,
    weights=None,
    nsamples_in_weights=True,
    graph_mode=False,
    grp_size_threshold=5,
    n_profile_steps=0,
    profile_log_dir="./logdir",
    model_regularization="sum",
    init_guesses_from_previous_time_step=False,
    skip_threshold=0.5,
    use_model_snr_weights=False,
    **opt_kwargs,
):
    """Simultaneous gain calibration and foreground modelling of every (polarization, time) in `uvdata` --
    calibration.py:963-1331.  Returns (model, resid, gains, fit_history) exactly as the reference does;
    fit_history[polnum][time_index]['loss'] is a list of np.float32 (np.float64 when dtype=np.float64).

    The foreground basis is uploaded to the device once per call (the reference tensorises it once per call
    too, calibration.py:1143-1152); each integration then only moves its data, weights and gains.
    """
    antpairs_data = uvdata.get_antpairs()
    if not include_autos:
        antpairs_data = set(ap for ap in antpairs_data if ap[0] != ap[1])
    uvdata = uvdata.select(inplace=False, bls=[ap for ap in antpairs_data])
    resid = copy.deepcopy(uvdata)
    model = copy.deepcopy(uvdata)
    model.data_array[:] = 0.0
    model.flag_array[:] = False
    red_grps = [red for fit_grp in fg_model_comps_dict.keys() for red in fit_grp]
    if gains is None:
        echo(f"{datetime.datetime.now()} Gains are None. Initializing gains starting with unity...\n", verbose=verbose)
        gains = cal_utils.blank_uvcal_from_uvdata(uvdata)
    if sky_model is None and model_regularization is not None:
        echo(f"{datetime.datetime.now()} Sky model is None. Initializing from data...\n", verbose=verbose)
        sky_model = cal_utils.apply_gains(uvdata, gains)
    else:
        sky_model = sky_model.select(inplace=False, bls=[ap for ap in antpairs_data])  # AttributeError on None (quirk Q7)
    if optimizer not in OPTIMIZERS:
        raise KeyError(optimizer)

    fit_history = {}
    ants_map = {ant: i for i, ant in enumerate(gains.ant_array)}
    echo(f"{datetime.datetime.now()} Computing foreground components matrices...\n", verbose=verbose)
    fdt = _fit_dtype(dtype)
    lay = _layout_from_dict(fg_model_comps_dict, ants_map, sky_model.Nfreqs, use_redundancy, grp_size_threshold,
                            nants=max(len(ants_map), uvdata.Nants_data), dtype=fdt)
    bl_pairs = list(zip(lay.bl_ant0.tolist(), lay.bl_ant1.tolist()))
    del fg_model_comps_dict
    pols = list(uvdata.get_pols())
    times = list(np.unique(uvdata.time_array))
    devices = _device_indices(len(pols) * len(times), init_guesses_from_previous_time_step)
    plans = [FitPlan(lay, device=d) for d in devices]  # the basis is replicated: one upload per device, once per call
    echo(f"{datetime.datetime.now()}Finished Converting Foreground Modeling Components to Tensors...\n", verbose=verbose)
    state = {}  # (plan index) -> carried (g_r, g_i, first_time) for init_guesses_from_previous_time_step

    def process(plan, pidx, polnum, pol, time_index, time):
        """One (polarization, time) integration on `plan` (calibration.py:1167-1320).  Units write disjoint slices of
        model / gains / resid, so several of them may run on different devices at once."""
        echo(f"{datetime.datetime.now()} Working on pol {pol}, {polnum + 1} of {uvdata.Npols}, time {time_index + 1} of "
             f"{uvdata.Ntimes}...\n", verbose=verbose)
        bltsel = np.isclose(uvdata.time_array, time, atol=1e-7, rtol=0.0)
        unflagged = ~uvdata.flag_array[bltsel, 0, :, polnum]
        frac_unflagged = np.count_nonzero(unflagged) / (uvdata.Nbls * uvdata.Nfreqs)
        history = None
        if frac_unflagged >= skip_threshold:
            rmsdata = np.sqrt(np.mean(np.abs(uvdata.data_array[bltsel, 0, :, polnum][unflagged]) ** 2.0))
            echo(f"{datetime.datetime.now()} Tensorizing data...\n", verbose=verbose)
            d_r, d_i, w = _tensorize_data_flat(uvdata, bl_pairs, ants_map, pol, time, rmsdata, weights,
                                               nsamples_in_weights, fdt)
            plan.set_integration(d_r, d_i, w)
            s_r = s_i = None
            if sky_model is not None:
                echo(f"{datetime.datetime.now()} Tensorizing sky model...\n", verbose=verbose)
                s_r, s_i, _ = _tensorize_data_flat(sky_model, bl_pairs, ants_map, pol, time, rmsdata, weights,
                                                   False, fdt)
            carried = state.get((pidx, polnum))
            if carried is None or not init_guesses_from_previous_time_step:
                echo(f"{datetime.datetime.now()} Tensorizing Gains...\n", verbose=verbose)
                g_r, g_i = tensorize_gains(gains, dtype=fdt, time=time, polarization=pol)
                g_r, g_i = _pad_gain_rows(g_r, lay.nants), _pad_gain_rows(g_i, lay.nants, fill=0.0)
                plan.set_gains(g_r, g_i)
                echo(f"{datetime.datetime.now()} Tensorizing Foreground coeffs...\n", verbose=verbose)
                plan.init_coeffs(s_r, s_i)  # TypeError on a None sky model, like tensorize_fg_coeffs(None, ...)
                if use_model_snr_weights:
                    plan.apply_model_snr_weights()
            else:
                g_r, g_i = carried
            priors = (0.0, 0.0)
            if model_regularization == "sum":
                priors = plan.prior_sums(s_r, s_i)
            out = _run_fit(plan, lay, g_r, g_i, None, None, use_min, tol, maxsteps, optimizer, freeze_model,
                           verbose, n_profile_steps, profile_log_dir, model_regularization, priors, dtype,
                           opt_kwargs, graph_mode=graph_mode)
            g_r, g_i, _, _, history = out
            state[(pidx, polnum)] = (g_r, g_i)
            vis_r, vis_i = plan.get_model()
            cube_r = np.zeros((lay.nants, lay.nants, lay.nfreqs))
            cube_i = np.zeros_like(cube_r)
            cube_r[lay.bl_ant0, lay.bl_ant1] = vis_r
            cube_i[lay.bl_ant0, lay.bl_ant1] = vis_i
            insert_model_into_uvdata_tensor(uvdata=model, time=time, polarization=pol, ants_map=ants_map,
                                            red_grps=red_grps, model_r=cube_r, model_i=cube_i,
                                            scale_factor=rmsdata)
            insert_gains_into_uvcal(uvcal=gains, time=time, polarization=pol, gains_re=g_r, gains_im=g_i)
        else:
            echo(f"{datetime.datetime.now()}: Only {frac_unflagged * 100}-percent of data unflagged. Skipping...\n",
                 verbose=verbose)
            flag_poltime(resid, time=time, polarization=pol)
            flag_poltime(gains, time=time, polarization=pol)
            flag_poltime(model, time=time, polarization=pol)
        if not freeze_model and model_regularization == "post_hoc" and np.any(~model.flag_array[bltsel]):
            renormalize(uvdata_reference_model=sky_model, uvdata_deconv=model, gains=gains, polarization=pol,
                        time=time, additional_flags=uvdata.flag_array)
        return history

    units = [(polnum, pol, ti, t) for polnum, pol in enumerate(pols) for ti, t in enumerate(times)]
    results = {}
    try:
        if len(plans) == 1:
            for (polnum, pol, ti, t) in units:
                results[(polnum, ti)] = process(plans[0], 0, polnum, pol, ti, t)
        else:
            # one host thread per device (ctypes releases the GIL during native calls; a plan is driven by one thread at a
            # time, include/calamity_b200.h), units dealt round-robin, no communication between devices
            from concurrent.futures import ThreadPoolExecutor

            def worker(k):
                return {(u[0], u[2]): process(plans[k], k, *u) for u in units[k :: len(plans)]}

            with ThreadPoolExecutor(max_workers=len(plans)) as pool:
                for part in pool.map(worker, range(len(plans))):
                    results.update(part)
    finally:
        for plan in plans:
            plan.close()
    for polnum, _ in enumerate(pols):
        # skipped units leave no entry (the reference overwrites its "skipped!" marker the same way, quirk Q10)
        fit_history[polnum] = {ti: results[(polnum, ti)] for ti, _ in enumerate(times) if results.get((polnum, ti)) is not None}

    model_with_gains = cal_utils.apply_gains(model, gains, inverse=True)
    if not correct_model:
        model = model_with_gains
    resid.data_array -= model_with_gains.data_array
    resid.data_array[model_with_gains.flag_array] = 0.0
    resid.data_array[uvdata.flag_array] = 0.0
    if correct_resid:
        resid = cal_utils.apply_gains(resid, gains)
    return model, resid, gains, fit_history


def _pad_gain_rows(g, nants, fill=1.0):
    g = np.asarray(g)
    if g.dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        g = g.astype(np.float32)
    if g.ndim == 1:
        g = g[None, :]
    if g.shape[0] >= nants:
        return g
    pad = np.full((nants - g.shape[0], g.shape[1]), fill, dtype=g.dtype)
    return np.concatenate([g, pad], axis=0)


def _is_uvdata_like(obj):
    return all(hasattr(obj, a) for a in ("data_array", "flag_array", "time_array", "polarization_array"))


def _is_uvcal_like(obj):
    return all(hasattr(obj, a) for a in ("gain_array", "flag_array", "time_array", "jones_array"))


def flag_poltime(data_object, time, polarization):
    """Flag one (time, polarization) of a UVData (data -> 0) or UVCal (gains -> 1) -- calibration.py:1334-1350."""
    if isinstance(data_object, UVData) or (_is_uvdata_like(data_object) and not _is_uvcal_like(data_object)):
        rows = np.isclose(data_object.time_array, time, atol=1e-7, rtol=0.0)
        pnum = np.where(
            data_object.polarization_array == _polstr2num(polarization, x_orientation=data_object.x_orientation)
        )[0][0]
        data_object.flag_array[rows, :, :, pnum] = True
        data_object.data_array[rows, :, :, pnum] = 0.0
    elif isinstance(data_object, UVCal) or _is_uvcal_like(data_object):
        jnum = np.where(data_object.jones_array == _polstr2num(polarization, x_orientation=data_object.x_orientation))[0][0]
        tnum = np.where(np.isclose(data_object.time_array, time, atol=1e-7, rtol=0.0))[0][0]
        data_object.gain_array[:, 0, :, tnum, jnum] = 1.0
        data_object.flag_array[:, 0, :, tnum, jnum] = True
    else:
        raise ValueError("only supports data_object that is UVCal or UVData.")


def calibrate_and_model_mixed(
    uvdata,
    horizon=1.0,
    min_dly=0.0,
    offset=0.0,
    ant_dly=0.0,
    include_autos=False,
    verbose=False,
    red_tol=1.0,
    red_tol_freq=0.5,
    n_angle_bins=200,
    notebook_progressbar=False,
    use_redundancy=False,
    use_tensorflow_to_derive_modeling_comps=False,
    eigenval_cutoff=1e-10,
    dtype_matinv=np.float64,
    require_exact_angle_match=True,
    angle_match_tol=1e-3,
    grp_size_threshold=5,
    model_comps_dict=None,
    save_dict_to=None,
    **fitting_kwargs,
):
    """Fit with DPSS vectors for isolated baselines and joint covariance eigenvectors for baselines that overlap
    in the uv plane -- calibration.py:1353-1500.  `use_tensorflow_to_derive_modeling_comps` selects a
    TensorFlow eigensolver upstream and is ignored here (the NumPy path is used)."""
    fitting_grps, blvecs, _, _ = modeling.get_uv_overlapping_grps_conjugated(
        uvdata, red_tol=red_tol, include_autos=include_autos, red_tol_freq=red_tol_freq, n_angle_bins=n_angle_bins,
        notebook_progressbar=notebook_progressbar, require_exact_angle_match=require_exact_angle_match,
        angle_match_tol=angle_match_tol)
    if model_comps_dict is None:
        model_comps_dict = modeling.yield_mixed_comps(
            fitting_grps, blvecs, uvdata.freq_array[0], eigenval_cutoff=eigenval_cutoff,
            use_tensorflow=use_tensorflow_to_derive_modeling_comps, ant_dly=ant_dly, horizon=horizon, offset=offset,
            min_dly=min_dly, verbose=verbose, dtype=dtype_matinv, notebook_progressbar=notebook_progressbar,
            grp_size_threshold=grp_size_threshold)
    if save_dict_to is not None:
        np.save(save_dict_to, model_comps_dict)
    return calibrate_and_model_tensor(uvdata=uvdata, fg_model_comps_dict=model_comps_dict, include_autos=include_autos,
                                      verbose=verbose, notebook_progressbar=notebook_progressbar,
                                      use_redundancy=use_redundancy, **fitting_kwargs)


def calibrate_and_model_dpss(
    uvdata,
    horizon=1.0,
    min_dly=0.0,
    offset=0.0,
    include_autos=False,
    verbose=False,
    red_tol=1.0,
    notebook_progressbar=False,
    fg_model_comps_dict=None,
    **fitting_kwargs,
):
    """Fit with per-baseline DPSS foreground vectors -- calibration.py:1503-1584.  As upstream, the basis is
    always rebuilt from the data's baselines (the `fg_model_comps_dict` argument is accepted and unused) and
    `use_redundancy` only reaches the chunking (quirk Q6)."""
    comps = modeling.yield_pbl_dpss_model_comps(uvdata, horizon=horizon, min_dly=min_dly, offset=offset,
                                                include_autos=include_autos, red_tol=red_tol,
                                                notebook_progressbar=notebook_progressbar, verbose=verbose)
    return calibrate_and_model_tensor(uvdata=uvdata, fg_model_comps_dict=comps, include_autos=include_autos,
                                      verbose=verbose, notebook_progressbar=notebook_progressbar, **fitting_kwargs)


def fg_model(fg_r, fg_i, fg_comps):
    """sum_k coeff_k * comp_k for one chunk (calibration.py:1587-1590), evaluated on the device."""
    comps = np.asarray(fg_comps)
    nvecs, ngrps, nbls, nfreqs = comps.shape
    corr = [[[(0, 0)] * nbls for _ in range(ngrps)]]
    fdt = _contraction_dtype([comps, np.asarray(fg_r), np.asarray(fg_i)])
    lay = RaggedLayout.from_dense([comps], corr, 1, dtype=fdt)
    with FitPlan(lay, device=_device_index()) as plan:
        zeros = np.zeros((lay.nbls, nfreqs), dtype=fdt)
        plan.set_integration(zeros, zeros, zeros)
        plan.set_gains(np.ones((1, nfreqs), dtype=fdt), np.zeros((1, nfreqs), dtype=fdt))
        plan.set_coeffs(lay.flatten_coeffs([fg_r]), lay.flatten_coeffs([fg_i]))
        v_r, v_i = plan.get_model()
    return _as_tensor(v_r.reshape(ngrps, nbls, nfreqs)), _as_tensor(v_i.reshape(ngrps, nbls, nfreqs))


def read_calibrate_and_model_dpss(
    input_data_files,
    input_model_files=None,
    input_gain_files=None,
    resid_outfilename=None,
    gain_outfilename=None,
    model_outfilename=None,
    fitted_info_outfilename=None,
    x_orientation="east",
    clobber=False,
    bllen_min=0.0,
    bllen_max=np.inf,
    bl_ew_min=0.0,
    ex_ants=None,
    select_ants=None,
    gpu_index=None,
    gpu_memory_limit=None,
    precision=32,
    use_autocorrs_in_weights=False,
    **calibration_kwargs,
):
    """File-level driver -- calibration.py:1659-1817.  File I/O needs pyuvdata; already-loaded objects can be
    passed instead of paths.  `gpu_index` selects the CUDA device; `gpu_memory_limit` is a TensorFlow
    allocator knob with no equivalent here (the library allocates exactly what the plan needs)."""
    if gpu_index is not None:
        os.environ["CALAMITY_B200_DEVICE"] = str(int(gpu_index))

    def _load(cls, files, reader):
        if isinstance(files, str):
            files = [files]
        if isinstance(files, list):
            obj = cls()
            getattr(obj, reader)(files)
            return obj
        return files

    uvd = _load(UVData, input_data_files, "read")
    weights = get_auto_weights(uvd) if use_autocorrs_in_weights else None
    utils.select_baselines(uvd, bllen_min=bllen_min, bllen_max=bllen_max, bl_ew_min=bl_ew_min, ex_ants=ex_ants,
                           select_ants=select_ants)
    uvd_model = _load(UVData, input_model_files, "read") if input_model_files is not None else None
    if uvd_model is not None:
        utils.select_baselines(uvd, bllen_min=bllen_min, bllen_max=bllen_max, bl_ew_min=bl_ew_min)
    uvc = _load(UVCal, input_gain_files, "read_calfits") if input_gain_files is not None else None
    dtype = {32: np.float32, 64: np.float64}[precision]
    model_fit, resid_fit, gains_fit, fit_info = calibrate_and_model_dpss(
        uvdata=uvd, sky_model=uvd_model, gains=uvc, dtype=dtype, weights=weights, **calibration_kwargs)
    if resid_outfilename is not None:
        resid_fit.write_uvh5(resid_outfilename, clobber=clobber)
    if gain_outfilename is not None:
        gains_fit.x_orientation = x_orientation
        gains_fit.write_calfits(gain_outfilename, clobber=clobber)
    if model_outfilename is not None:
        model_fit.write_uvh5(model_outfilename, clobber=clobber)
    fit_info["calibration_kwargs"] = calibration_kwargs
    fit_info["calibration_kwargs"]["dtype"] = dtype
    return model_fit, resid_fit, gains_fit, fit_info


# ----------------------------------------------------------------------------------------------------
# command line (calibration.py:1820-1942): same flags and defaults
# ----------------------------------------------------------------------------------------------------
def input_output_parser():
    ap = argparse.ArgumentParser()
    sp = ap.add_argument_group("Input and Output Arguments.")
    sp.add_argument("--input_data_files", type=str, nargs="+", help="paths to data files to calibrate.", required=True)
    sp.add_argument("--input_model_files", type=str, nargs="+", help="paths to model files to set overal amplitude and phase.")
    sp.add_argument("--input_gain_files", type=str, nargs="+", help="paths to gains to use as a staring point.")
    sp.add_argument("--resid_outfilename", type=str, default=None, help="postfix for resid output file.")
    sp.add_argument("--model_outfilename", type=str, default=None, help="postfix for foreground model file.")
    sp.add_argument("--gain_outfilename", type=str, default=None, help="path for writing fitted gains.")
    sp.add_argument("--clobber", action="store_true", default="False", help="Overwrite existing outputs.")
    sp.add_argument("--x_orientation", default="east", type=str, help="x_orientation of feeds to set in output gains.")
    sp.add_argument("--bllen_min", default=0.0, type=float, help="minimum baseline length to include in calibration and outputs.")
    sp.add_argument("--bllen_max", default=np.inf, type=float, help="maximum baseline length to include in calbration and outputs.")
    sp.add_argument("--bl_ew_min", default=0.0, type=float, help="minimum EW baseline component to include in calibration and outputs.")
    sp.add_argument("--ex_ants", default=None, type=int, nargs="+", help="Antennas to exclude from calibration and modeling.")
    sp.add_argument("--select_ants", default=None, type=int, nargs="+", help="Antennas to select exclusively for calibration and modeling.")
    sp.add_argument("--gpu_index", default=None, type=int, help="Index of GPU to run on (if on a multi-GPU machine).")
    sp.add_argument("--gpu_memory_limit", default=None, type=int, help="Limit GPU memory use to this many GBytes.")
    sp.add_argument("--precision", default=32, type=int, help="Number of bits to keep track of.")
    return ap


def fitting_argparser():
    ap = input_output_parser()
    sp = ap.add_argument_group("General Fitting Arguments.")
    sp.add_argument("--tol", type=float, default=1e-14, help="Stop gradient descent after cost function converges to within this value.")
    sp.add_argument("--optimizer", type=str, default="Adamax", help="First order optimizer to use for gradient descent.")
    sp.add_argument("--maxsteps", type=int, default=10000, help="Max number of steps to iterate during optimization.")
    sp.add_argument("--verbose", default=False, action="store_true", help="lots of text ouputs.")
    sp.add_argument("--use_min", default=False, action="store_true",
                    help="Use params for mimimum cost function derived. Otherwise, use the params last visited by the descent. Avoids momentum overshoot.")
    sp.add_argument("--use_redundancy", default=False, action="store_true", help="Model redundant visibilities with the same set of foreground parameters.")
    sp.add_argument("--correct_model", default=True, action="store_true", help="Remove gain effects from foreground model.")
    sp.add_argument("--correct_resid", default=False, action="store_true", help="Apply fitted gains to the fitted residuals.")
    sp.add_argument("--graph_mode", default=False, action="store_true", help="Pre-compile computational graph before running gradient descent. Not reccomended for GPUs.")
    sp.add_argument("--init_guesses_from_previous_time_step", default=False, action="store_true",
                    help="initialize gain and foreground guesses from previous time step when calibrating multiple times.")
    sp.add_argument("--learning_rate", type=float, default=1e-2, help="gradient descent learning rate.")
    sp.add_argument("--red_tol", type=float, default=1.0, help="Tolerance for determining redundancy between baselines [meters].")
    sp.add_argument("--skip_threshold", type=float, default=0.5, help="Skip and flag time/polarization if more then this fractionf of data is flagged.")
    sp.add_argument("--model_regularization", type=str, default="post_hoc")
    sp.add_argument("--nsamples_in_weights", default=False, action="store_true", help="Weight contributions to MSE by nsamples.")
    sp.add_argument("--use_model_snr_weights", default=False, action="store_true", help="If True, weight contributions to MSE as proportional to SNR.")
    sp.add_argument("--use_autocorrs_in_weights", default=False, action="store_true", help="If True, use autocorrelations to derive relative SNR weights.")
    return ap


def dpss_fit_argparser():
    ap = fitting_argparser()
    sp = ap.add_argument_group("DPSS Specific Fitting Arguments.")
    sp.add_argument("--horizon", default=1.0, type=float, help="Fraction of horizon delay to model with DPSS modes.")
    sp.add_argument("--min_dly", default=0.0, type=float, help="Minimum delay [ns] to model with DPSS modes.")
    sp.add_argument("--offset", default=0.0, type=float, help="Offset from horizon delay [ns] to model with DPSS modes.")
    return ap
