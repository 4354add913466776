"""CPU: host-side logic of the drop-in layer -- chunking / tensor layout / index construction (bit-exact
against the oracle), the modeling golden vector of the reference, argparse defaults, sharding."""
import os
import re
import sys

import numpy as np
import pytest

from calamity_b200 import calibration, modeling, simple_cov
from calamity_b200.layout import RaggedLayout
from calamity_b200.sharding import make_shard, partition_groups
from calamity_b200.uvstandins import MiniUVData
from oracle import restatement as R
from tests import fixtures_uv as fx

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol(native_built):
    header = open(os.path.join(ROOT, "include", "calamity_b200.h")).read()
    declared = set(re.findall(r"\b(calb2_[a-z0-9_]+)\s*\(", header))
    from calamity_b200 import _native

    assert declared == set(_native.EXPORTED_SYMBOLS), declared ^ set(_native.EXPORTED_SYMBOLS)
    for sym in declared:
        assert hasattr(native_built, sym)
    assert b"sm_100a" in native_built.calb2_version()


def test_get_uv_overlapping_grps_conjugated_golden():
    """The one bit-exact vector in the reference's tests: calamity/tests/test_modeling.py:20-32."""
    uvd = fx.line_array()
    grps, centers, connections, labels = modeling.get_uv_overlapping_grps_conjugated(uvdata=uvd, red_tol_freq=0.5, n_angle_bins=200)
    assert grps == [
        [((0, 1),)],
        [((3, 4),)],
        [((1, 2),)],
        [((0, 2),)],
        [((4, 5),)],
        [((2, 3),), ((3, 5),), ((2, 4),), ((1, 3),), ((0, 3),), ((1, 4),), ((0, 4),), ((2, 5),)],
        [((1, 5),), ((0, 5),)],
    ]


@pytest.mark.parametrize("horizon, offset, min_dly, ant_dly", [(1.0, 20.0, 0.0, 0.0), (0.8, 123.0, 200.0, 0.0), (1.0, 0.0, 0.0, 2 / 0.3)])
def test_simple_cov_closed_form(horizon, offset, min_dly, ant_dly):
    """calamity/tests/test_simple_cov.py:21-45."""
    freqs = 100e6 + 100e3 * np.arange(200)
    blvecs = np.array([[2.0, 0.0, 0.0]])
    fg0, fg1 = np.meshgrid(freqs, freqs)
    bldly = np.max([np.linalg.norm(blvecs[0]) * horizon / 0.3 + offset, min_dly])
    want = np.sinc(2 * bldly * (fg0 - fg1) / 1e9)
    if ant_dly > 0:
        want *= np.sinc(2 * (fg0 - fg1) / 1e9 * ant_dly)
    got = simple_cov.simple_cov_matrix(blvecs, freqs, ant_dly=ant_dly, horizon=horizon, offset=offset, min_dly=min_dly)
    assert np.allclose(got, want)


def test_dpss_fit_argparser_defaults():
    """calamity/tests/test_calibration.py:758-765."""
    argv = sys.argv
    sys.argv = [argv[0], "--input_data_files", "input.uvh5"]
    try:
        args = calibration.dpss_fit_argparser().parse_args()
    finally:
        sys.argv = argv
    assert args.learning_rate == 1e-2 and args.tol == 1e-14 and args.maxsteps == 10000
    assert args.input_data_files == ["input.uvh5"]
    assert args.model_regularization == "post_hoc" and args.optimizer == "Adamax" and args.precision == 32


def _mixed_dict(nfreqs=24, seed=0):
    rng = np.random.default_rng(seed)
    d = {}
    d[(((0, 1), (1, 2)), ((0, 2),))] = rng.standard_normal((2 * nfreqs, 5))        # unequal sub-group sizes: never split
    d[(((2, 3), (3, 4)), ((1, 3), (2, 4)))] = rng.standard_normal((2 * nfreqs, 7))  # equal sizes: split unless use_redundancy
    d[(((0, 4),),)] = rng.standard_normal((nfreqs, 3))
    d[(((0, 3),), ((1, 4),), ((0, 5),), ((1, 5),), ((2, 5),), ((3, 5),))] = rng.standard_normal((6 * nfreqs, 9))  # too many to split
    d[(((4, 5),),)] = rng.standard_normal((nfreqs, 4))
    return d, nfreqs


@pytest.mark.parametrize("use_redundancy", [False, True])
@pytest.mark.parametrize("threshold", [1, 3, 5, 7])
def test_chunking_and_tensor_layout_match_oracle_bit_exact(use_redundancy, threshold):
    d, nf = _mixed_dict()
    ants_map = {a: a for a in range(6)}
    mine = calibration.chunk_fg_comp_dict_by_nbls(d, use_redundancy=use_redundancy, grp_size_threshold=threshold)
    want = R.chunk_by_nbls(d, use_redundancy=use_redundancy, grp_size_threshold=threshold)
    assert list(mine.keys()) == list(want.keys())
    for k in want:
        assert list(mine[k].keys()) == list(want[k].keys())
    dense, corr = calibration.tensorize_fg_model_comps_dict(d, ants_map, nf, use_redundancy=use_redundancy,
                                                            dtype=np.float64, grp_size_threshold=threshold)
    odense, ocorr = R.tensorize_basis(d, ants_map, nf, use_redundancy=use_redundancy, dtype=np.float64,
                                      grp_size_threshold=threshold)
    assert corr == ocorr
    assert len(dense) == len(odense)
    for a, b in zip(dense, odense):
        assert a.dtype == np.float64 and np.array_equal(a.numpy(), b)
    # the ragged layout carries the same information as the dense tensors
    lay = calibration._layout_from_dict(d, ants_map, nf, use_redundancy, threshold)
    for a, b in zip(lay.dense_chunks(np.float32), odense):
        assert np.array_equal(a, b.astype(np.float32))
    assert lay.corr_inds() == ocorr
    assert sum(t.shape[1] * t.shape[2] * t.shape[3] for t in dense) == lay.nbls * nf
    # dense -> ragged -> dense is lossless, and flat <-> chunked vectors round-trip
    lay2 = RaggedLayout.from_dense([b.astype(np.float32) for b in odense], ocorr, 6)
    for a, b in zip(lay2.dense_chunks(np.float32), odense):
        assert np.array_equal(a, b.astype(np.float32))
    rng = np.random.default_rng(1)
    flat = rng.standard_normal(lay.ncoef).astype(np.float32)
    assert np.array_equal(lay.flatten_coeffs(lay.unflatten_coeffs(flat)), flat)
    data = rng.standard_normal((lay.nbls, nf)).astype(np.float32)
    assert np.array_equal(lay.flatten_data(lay.unflatten_data(data)), data)


def test_chunk_dpss_dict_is_one_chunk():
    """calamity/tests/test_calibration.py:274-278."""
    comps = fx.dpss_vectors(fx.line_array())
    chunked = calibration.chunk_fg_comp_dict_by_nbls(comps)
    maxvecs = np.max([comps[k].shape[1] for k in comps])
    assert len(chunked) == 1 and list(chunked.keys())[0] == (1, maxvecs)


def test_tensorize_fg_model_comps_dpss_rows_and_padding():
    """calamity/tests/test_calibration.py:244-271."""
    uvd = fx.line_array()
    comps = fx.dpss_vectors(uvd)
    gains = fx.cal_utils.blank_uvcal_from_uvdata(uvd)
    ants_map = {ant: i for i, ant in enumerate(gains.ant_array)}
    tensors, corr = calibration.tensorize_fg_model_comps_dict(comps, ants_map, dtype=np.float64, nfreqs=uvd.Nfreqs)
    seen = 0
    for c in range(len(corr)):
        for g in range(len(corr[c])):
            for b, bl in enumerate(corr[c][g]):
                rows = tensors[c][:, g, b].numpy().squeeze()
                want = comps[((bl,),)].T
                assert np.allclose(want, rows[: want.shape[0]])
                assert np.allclose(0.0, rows[want.shape[0] :])
                seen += 1
    assert seen == len(comps)


def test_tensorize_gains_layout():
    """calamity/tests/test_calibration.py:233-241."""
    uvd = fx.line_array()
    gains = fx.cal_utils.blank_uvcal_from_uvdata(uvd)
    for i, ant in enumerate(gains.ant_array):
        gains.gain_array[i] *= ant + 1.0
    g_r, g_i = calibration.tensorize_gains(gains, polarization="xx", time=gains.time_array[0], dtype=np.float64)
    assert g_r.dtype == np.float64 and g_i.dtype == np.float64
    for ant in gains.ant_array:
        assert np.allclose(g_r.numpy()[ant], ant + 1) and np.allclose(g_i.numpy()[ant], 0.0)


def test_tensorize_data_matches_oracle_and_handles_conjugates():
    uvd = fx.line_array()
    rng = np.random.default_rng(3)
    uvd.flag_array[:] = rng.random(uvd.flag_array.shape) < 0.1
    uvd.nsample_array[:] = rng.integers(1, 4, uvd.nsample_array.shape)
    gains = fx.cal_utils.blank_uvcal_from_uvdata(uvd)
    ants_map = {ant: i for i, ant in enumerate(gains.ant_array)}
    # ask for some baselines in the orientation the data does NOT store: they must come back conjugated
    comps = {}
    for n, ap in enumerate(uvd.get_antpairs()):
        key = ap if n % 2 == 0 else ap[::-1]
        comps[((key,),)] = np.ones((uvd.Nfreqs, 2))
    _, corr = calibration.tensorize_fg_model_comps_dict(comps, ants_map, uvd.Nfreqs)
    before = uvd.data_array.copy()
    d_r, d_i, w = calibration.tensorize_data(uvd, corr, ants_map, "xx", uvd.time_array[0], data_scale_factor=3.0,
                                             nsamples_in_weights=True, dtype=np.float64)
    assert np.array_equal(before, uvd.data_array)  # the caller's data are not rescaled in place
    n = uvd.Nants_data
    cube = np.zeros((n, n, uvd.Nfreqs), dtype=np.complex128)
    flg = np.ones((n, n, uvd.Nfreqs), dtype=bool)
    ns = np.zeros((n, n, uvd.Nfreqs))
    for ap in uvd.get_antpairs():
        row = uvd.antpair2ind(ap)[0]
        i, j = ants_map[ap[0]], ants_map[ap[1]]
        cube[i, j] = uvd.data_array[row, 0, :, 0]
        cube[j, i] = np.conj(uvd.data_array[row, 0, :, 0])
        flg[i, j] = flg[j, i] = uvd.flag_array[row, 0, :, 0]
        ns[i, j] = ns[j, i] = uvd.nsample_array[row, 0, :, 0]
    o_r, o_i, o_w = R.tensorize_data_cubes(cube, flg, ns, corr, data_scale_factor=3.0, nsamples_in_weights=True, dtype=np.float64)
    for a, b in zip(d_r + d_i + w, o_r + o_i + o_w):
        assert np.array_equal(a.numpy(), b)
    assert abs(sum(float(x.numpy().sum()) for x in w) - 1.0) < 1e-12
    with pytest.raises(IndexError):
        calibration.tensorize_data(uvd, corr, ants_map, "xx", uvd.time_array[0] + 1.0)


def test_flag_poltime_and_renormalize():
    """calamity/tests/test_calibration.py:222-230, 599-607."""
    uvd = fx.line_array(ntimes=2)
    ref = fx.copy.deepcopy(uvd)
    t0 = np.unique(uvd.time_array)[0]
    calibration.flag_poltime(uvd, time=t0, polarization="xx")
    assert np.all(uvd.flag_array[: uvd.Nbls]) and not np.any(uvd.flag_array[uvd.Nbls : 2 * uvd.Nbls])
    assert np.allclose(uvd.data_array[: uvd.Nbls], 0.0)
    assert np.allclose(uvd.data_array[uvd.Nbls :], ref.data_array[uvd.Nbls :])
    gains = fx.cal_utils.blank_uvcal_from_uvdata(ref)
    gains.gain_array *= 3.0
    calibration.flag_poltime(gains, time=t0, polarization="xx")
    assert np.allclose(gains.gain_array[:, 0, :, 0, 0], 1.0) and np.all(gains.flag_array[:, 0, :, 0, 0])
    assert np.allclose(gains.gain_array[:, 0, :, 1, 0], 3.0)
    with pytest.raises(ValueError):
        calibration.flag_poltime(data_object="blarghle", time=0, polarization="xx")
    sky = fx.line_array()
    sky_ref = fx.copy.deepcopy(sky)
    g = fx.cal_utils.blank_uvcal_from_uvdata(sky)
    g.gain_array *= (51.0 + 23j) ** -0.5
    sky.data_array *= 51.0 + 23j
    calibration.renormalize(sky_ref, sky, g, polarization="xx", time=sky.time_array[0])
    assert np.allclose(np.abs(g.gain_array), 1.0)
    assert np.allclose(np.abs(sky_ref.data_array), np.abs(sky.data_array))


def test_insert_gains_and_apply_gains_roundtrip():
    uvd = fx.line_array()
    gains = fx.randomized_gains(uvd, scatter=0.1)
    cal = fx.cal_utils.apply_gains(uvd, gains)
    back = fx.cal_utils.apply_gains(cal, gains, inverse=True)
    assert np.allclose(back.data_array, uvd.data_array)
    ap = uvd.get_antpairs()[3]
    row = uvd.antpair2ind(ap)[0]
    i0 = np.where(gains.ant_array == ap[0])[0][0]
    i1 = np.where(gains.ant_array == ap[1])[0][0]
    want = uvd.data_array[row, 0, :, 0] / (gains.gain_array[i0, 0, :, 0, 0] * np.conj(gains.gain_array[i1, 0, :, 0, 0]))
    assert np.allclose(cal.data_array[row, 0, :, 0], want)
    g2 = fx.cal_utils.blank_uvcal_from_uvdata(uvd)
    rng = np.random.default_rng(0)
    gr, gi = rng.standard_normal((6, 200)), rng.standard_normal((6, 200))
    calibration.insert_gains_into_uvcal(g2, uvd.time_array[0], "xx", gr, gi)
    assert np.allclose(g2.gain_array[:, 0, :, 0, 0], gr + 1j * gi)


def test_partition_balances_bytes_not_counts():
    w = np.concatenate([np.full(100, 5), np.full(20, 200)])
    for n in (2, 4, 8):
        ranges = partition_groups(w, n)
        assert ranges[0][0] == 0 and ranges[-1][1] == len(w)
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        loads = np.array([w[a:b].sum() for a, b in ranges])
        assert loads.max() <= w.sum() / n + w.max()


def test_unknown_optimizer_is_a_keyerror(native_built):
    from calamity_b200.fitter import FitPlan

    with pytest.raises(KeyError):
        FitPlan.fit(None, optimizer="NotAnOptimizer")


def test_tensorize_data_with_uvflag_weights_multitime():
    """UVFlag weights (calibration.py:287-296) on a two-time data set: the weights row of the requested time is used,
    multiplied by the unflagged mask, and everything is normalised to unit sum (300-303)."""
    uvd = fx.line_array(ntimes=2)
    rng = np.random.default_rng(5)
    uvd.flag_array[:] = rng.random(uvd.flag_array.shape) < 0.2
    weights = fx.unit_weights(uvd)
    weights.weights_array[:] = rng.uniform(0.5, 2.0, weights.weights_array.shape)
    gains = fx.cal_utils.blank_uvcal_from_uvdata(uvd)
    ants_map = {ant: i for i, ant in enumerate(gains.ant_array)}
    comps = {((ap,),): np.ones((uvd.Nfreqs, 2)) for ap in uvd.get_antpairs()}
    _, corr = calibration.tensorize_fg_model_comps_dict(comps, ants_map, uvd.Nfreqs)
    t1 = np.unique(uvd.time_array)[1]
    d_r, d_i, w = calibration.tensorize_data(uvd, corr, ants_map, "xx", t1, weights=weights, dtype=np.float64)
    want_w, want_d = [], []
    for grp in corr[0]:
        (i, j), = grp
        rows = uvd.antpair2ind(int(gains.ant_array[i]), int(gains.ant_array[j]))
        row = rows[np.isclose(uvd.time_array[rows], t1, rtol=0.0, atol=1e-7)][0]
        wrow = weights.antpair2ind(int(gains.ant_array[i]), int(gains.ant_array[j]))
        wrow = wrow[np.isclose(weights.time_array[wrow], t1, rtol=0.0, atol=1e-7)][0]
        want_w.append(weights.weights_array[wrow, 0, :, 0] * ~uvd.flag_array[row, 0, :, 0])
        want_d.append(uvd.data_array[row, 0, :, 0])
    want_w = np.asarray(want_w)
    want_w = want_w / want_w.sum()
    assert np.allclose(w[0].numpy()[:, 0], want_w, rtol=1e-13)
    assert np.array_equal(d_r[0].numpy()[:, 0], np.asarray(want_d).real)
    assert np.array_equal(d_i[0].numpy()[:, 0], np.asarray(want_d).imag)


def test_insert_model_into_uvdata_tensor_conjugates_and_times():
    """calibration.py:741-795: cube cell (i, j) goes to the row of the requested time only, conjugated when the data
    store the baseline the other way round, times scale_factor."""
    uvd = fx.line_array(ntimes=2)
    gains = fx.cal_utils.blank_uvcal_from_uvdata(uvd)
    ants_map = {ant: i for i, ant in enumerate(gains.ant_array)}
    n, nf = uvd.Nants_data, uvd.Nfreqs
    rng = np.random.default_rng(9)
    m_r, m_i = rng.standard_normal((n, n, nf)), rng.standard_normal((n, n, nf))
    aps = uvd.get_antpairs()
    red_grps = [[ap if k % 2 == 0 else ap[::-1]] for k, ap in enumerate(aps)]  # half of them asked for reversed
    before = uvd.data_array.copy()
    t1 = np.unique(uvd.time_array)[1]
    calibration.insert_model_into_uvdata_tensor(uvd, t1, "xx", ants_map, red_grps, m_r, m_i, scale_factor=2.5)
    for k, ap in enumerate(aps):
        rows = uvd.antpair2ind(ap)
        r0 = rows[np.isclose(uvd.time_array[rows], np.unique(uvd.time_array)[0], rtol=0.0, atol=1e-7)][0]
        r1 = rows[np.isclose(uvd.time_array[rows], t1, rtol=0.0, atol=1e-7)][0]
        assert np.array_equal(uvd.data_array[r0], before[r0])  # the other time is untouched
        i, j = ants_map[ap[0]], ants_map[ap[1]]
        if k % 2 == 0:
            want = (m_r[i, j] + 1j * m_i[i, j]) * 2.5
        else:  # asked as (j, i): stored the other way round -> conjugate of cell (j, i)
            want = (m_r[j, i] - 1j * m_i[j, i]) * 2.5
        assert np.allclose(uvd.data_array[r1, 0, :, 0], want, rtol=1e-15)
    with pytest.raises(IndexError):
        calibration.insert_model_into_uvdata_tensor(uvd, t1 + 7.0, "xx", ants_map, red_grps, m_r, m_i)


def test_header_is_plain_c_and_links_from_c(native_built, tmp_path):
    """The boundary is a C ABI: include/calamity_b200.h must compile as C99 (no C++ in the signatures) and a C program
    must link against the shared library and call into it without a GPU (calb2_version)."""
    import ctypes as C
    import shutil
    import subprocess

    from calamity_b200 import _native

    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = ROOT
    src = tmp_path / "use_cabi.c"
    src.write_text(
        '#include <stdio.h>\n#include <string.h>\n#include "calamity_b200.h"\n'
        "int main(void) {\n"
        "  calb2_plan_desc d; calb2_fit_options o; calb2_fit_result r; calb2_plan_info info;\n"
        "  memset(&d, 0, sizeof d); memset(&o, 0, sizeof o); memset(&r, 0, sizeof r); memset(&info, 0, sizeof info);\n"
        "  o.optimizer = CALB2_OPT_ADAMAX; o.regularization = CALB2_REG_SUM; d.dtype = CALB2_F64;\n"
                '  printf("%s %d %d %d %d %d\\n", calb2_version(), (int)sizeof(calb2_plan_desc), (int)sizeof(calb2_fit_result),\n'
        "         (int)sizeof(calb2_plan_info), (int)sizeof(calb2_fit_options), calb2_plan_create(0, 0));\n"
        "  return 0;\n}\n")
    exe = tmp_path / "use_cabi"
    libdir = os.path.dirname(_native.lib_path())
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"), str(src),
                    "-o", str(exe), "-L", libdir, "-lcalamity_b200", f"-Wl,-rpath,{libdir}"], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()
    assert "sm_100a" in " ".join(out)
    # the ctypes mirrors have the sizes the C compiler gives the structs; a null description is refused
    sizes = [int(x) for x in out[-5:-1]]
    assert sizes == [C.sizeof(_native.PlanDesc), C.sizeof(_native.FitResult), C.sizeof(_native.PlanInfo),
                     C.sizeof(_native.FitOptions)], sizes
    assert int(out[-1]) == -1


def test_product_package_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under calamity_b200/ (the product path) may import, open or execute
    anything under oracle/ -- and there is no CPU fallback to route to: the native loader raises when the library is
    missing."""
    import ast

    pkg = os.path.join(ROOT, "calamity_b200")
    for dirpath, _, files in os.walk(pkg):
        for name in files:
            if not name.endswith(".py"):
                continue
            path = os.path.join(dirpath, name)
            tree = ast.parse(open(path).read(), filename=path)
            for node in ast.walk(tree):
                mods = []
                if isinstance(node, ast.Import):
                    mods = [a.name for a in node.names]
                elif isinstance(node, ast.ImportFrom):
                    mods = [node.module or ""]
                assert not any(m == "oracle" or m.startswith("oracle.") for m in mods), (path, mods)
            assert "oracle/" not in open(path).read().replace("oracle/restatement.py for provenance", ""), path
    from calamity_b200 import _native

    saved, saved_lib = _native._LIB_PATH, _native._lib
    try:
        _native._LIB_PATH, _native._lib = os.path.join(pkg, "_lib", "does_not_exist.so"), None
        with pytest.raises(_native.NativeError):
            _native.load()
    finally:
        _native._LIB_PATH, _native._lib = saved, saved_lib


def test_driver_device_selection(monkeypatch):
    """calibrate_and_model_tensor spreads independent integrations over devices (calibration.py:1160-1167) but never when
    they are chained by init_guesses_from_previous_time_step (1210), never over more devices than units, and a pinned
    device (the file driver's gpu_index, calibration.py:1741-1752) wins."""
    from calamity_b200 import calibration as C

    monkeypatch.delenv("CALAMITY_B200_DEVICE", raising=False)
    monkeypatch.setenv("CALAMITY_B200_DEVICES", "0,2,5")
    assert C._device_indices(60, False) == [0, 2, 5]
    assert C._device_indices(2, False) == [0, 2]
    assert C._device_indices(1, False) == [0]
    assert C._device_indices(60, True) == [0]
    monkeypatch.setenv("CALAMITY_B200_DEVICE", "3")
    assert C._device_indices(60, False) == [3]


def test_get_auto_weights_inverse_variance_of_smoothed_autos():
    """calibration.py:916-960: weights of baseline (i, j) = 1 / (smooth auto_i x smooth auto_j) on unflagged samples, 0 on
    flagged ones, the autos smoothed by a least-squares fit of their real part to the DPSS modes of a `delay_extent` ns
    window over the unflagged channels."""
    from calamity_b200 import calibration as C
    from calamity_b200 import modeling
    from tests import fixtures_uv as fx

    uvd = fx.line_array(ntimes=2, with_autos=True)
    rng = np.random.default_rng(4)
    freqs = uvd.freq_array[0]
    comps = modeling.yield_dpss_model_comps_bl_grp(0.0, freqs, offset=25.0)
    truth = {}
    for a in range(6):  # smooth, positive autocorrelations inside the DPSS span + a little noise
        coeff = np.zeros(comps.shape[1])
        coeff[:3] = [40.0 + 5.0 * a, 3.0, -2.0]
        truth[a] = comps @ coeff
        rows = uvd.antpair2ind(a, a)
        uvd.data_array[rows, 0, :, 0] = truth[a][None, :] + 1e-3 * rng.standard_normal((len(rows), len(freqs)))
    uvd.flag_array[:] = rng.random(uvd.flag_array.shape) < 0.1
    out = C.get_auto_weights(uvd, delay_extent=25.0)
    assert out.weights_array.shape == uvd.data_array.shape
    for (i, j) in [(0, 1), (2, 5), (3, 3)]:
        rows = uvd.antpair2ind(i, j)
        want = 1.0 / (truth[i] * truth[j])
        got = out.weights_array[rows, 0, :, 0]
        flg = uvd.flag_array[rows, 0, :, 0]
        assert np.all(got[flg] == 0.0)
        assert np.allclose(got[~flg], np.broadcast_to(want, got.shape)[~flg], rtol=2e-3)


def test_chunk_variables_for_lamb_cover_the_coefficients_in_chunk_order():
    """LAMB's trust ratio is per tf.Variable; the reference holds one coefficient variable per chunk (calibration.py:560-567).
    `chunk_coef_bounds` must cut the flat coefficient vector exactly at the chunk boundaries of `flatten_coeffs`."""
    from calamity_b200 import fitter
    from oracle.restatement import KERAS_DEFAULTS
    from tests.helpers import mixed_problem

    p = mixed_problem(nants=24, nfreqs=96, seed=3, n_dpss_bls=120, joint=((3, 9, 20), (2, 7, 12)))
    lay = p.lay
    b = lay.chunk_coef_bounds()
    assert b[0] == 0 and b[-1] == lay.ncoef and np.all(np.diff(b) > 0) and len(b) - 1 == len(lay.chunks)
    # a coefficient tensor per chunk filled with the chunk's index lands exactly in that chunk's range
    chunks = [np.full((ch["nvecs"], ch["ngrps"], 1, 1), float(c)) for c, ch in enumerate(lay.chunks)]
    flat = lay.flatten_coeffs(chunks)
    for c in range(len(lay.chunks)):
        assert np.all(flat[b[c] : b[c + 1]] == c)
    # the product's LAMB defaults are the oracle's (tensorflow_addons.optimizers.LAMB)
    assert fitter.KERAS_DEFAULTS["LAMB"] == KERAS_DEFAULTS["LAMB"] and fitter.OPTIMIZER_IDS["LAMB"] == 8
