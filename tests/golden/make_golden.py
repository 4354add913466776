#!/usr/bin/env python
"""Generate tests/golden/reference_golden.npz by EXECUTING THE REFERENCE'S OWN CODE
(/root/reference/calamity/*.py, unmodified) under oracle/tf_shim (fake tensorflow / pyuvdata / hera_filters).

Run in the build container only (it needs /root/reference):   python tests/golden/make_golden.py
Inputs are regenerated from seeds by tests/golden_inputs.py, so only the reference's OUTPUTS are stored.
See oracle/tf_shim/README.md for what this does and does not pin.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import tf_shim  # noqa: E402

tf_shim.activate()
import tensorflow as tf  # noqa: E402  (the shim)
from calamity import calibration as ref  # noqa: E402
from calamity import cal_utils as ref_cu  # noqa: E402
from calamity import modeling as ref_mod  # noqa: E402

from tests import golden_inputs as gi  # noqa: E402

out = {}


def put(name, value):
    out[name] = np.asarray(value)


# ---- G1: chunking + dense tensor layout + corr_inds (calibration.py:30-190) -----------------------------------
d, nf = gi.mixed_dict()
ants_map = {a: a for a in range(6)}
for use_red in (False, True):
    for thr in (1, 3, 5):
        tag = f"g1_r{int(use_red)}_t{thr}"
        chunked = ref.chunk_fg_comp_dict_by_nbls(d, use_redundancy=use_red, grp_size_threshold=thr)
        put(tag + "_keys", np.asarray(list(chunked.keys())))
        tensors, corr = ref.tensorize_fg_model_comps_dict(d, ants_map, nf, use_redundancy=use_red, dtype=np.float64,
                                                          grp_size_threshold=thr)
        put(tag + "_nchunks", len(tensors))
        for c, (t, ci) in enumerate(zip(tensors, corr)):
            put(f"{tag}_tensor{c}", t.numpy())
            put(f"{tag}_corr{c}", np.asarray(ci))

# ---- G2: loss + tape gradient at fixed parameters (calibration.py:1587-1656, 664-666) -------------------------
for dtype, dn in ((np.float32, "f32"), (np.float64, "f64")):
    t = gi.reference_problem(dtype)
    conv = lambda xs: [tf.convert_to_tensor(x, dtype=dtype) for x in xs]
    g_r, g_i = tf.Variable(tf.convert_to_tensor(t["g_r"], dtype=dtype)), tf.Variable(tf.convert_to_tensor(t["g_i"], dtype=dtype))
    fg_r = [tf.Variable(x) for x in conv(t["fg_r"])]
    fg_i = [tf.Variable(x) for x in conv(t["fg_i"])]
    a0 = [[[p[0] for p in grp] for grp in chunk] for chunk in t["corr_inds"]]
    a1 = [[[p[1] for p in grp] for grp in chunk] for chunk in t["corr_inds"]]
    common = dict(g_r=g_r, g_i=g_i, fg_r=fg_r, fg_i=fg_i, fg_comps=conv(t["fg_comps"]), nchunks=len(fg_r),
                  data_r=conv(t["data_r"]), data_i=conv(t["data_i"]), wgts=conv(t["wgts"]), ant0_inds=a0, ant1_inds=a1,
                  dtype=dtype)
    for reg in ("none", "sum"):
        with tf.GradientTape() as tape:
            if reg == "sum":
                loss = ref.mse_chunked_sum_regularized(prior_r_sum=tf.constant(gi.PRIOR_R, dtype), prior_i_sum=tf.constant(gi.PRIOR_I, dtype), **common)
            else:
                loss = ref.mse_chunked(**common)
        grads = tape.gradient(loss, [g_r, g_i] + fg_r + fg_i)
        put(f"g2_{dn}_{reg}_loss", loss.numpy())
        put(f"g2_{dn}_{reg}_dg_r", grads[0].numpy())
        put(f"g2_{dn}_{reg}_dg_i", grads[1].numpy())
        put(f"g2_{dn}_{reg}_dfg_r", grads[2].numpy())
        put(f"g2_{dn}_{reg}_dfg_i", grads[3].numpy())

# ---- G3: fit_gains_and_foregrounds trajectories (calibration.py:447-738) ---------------------------------------
for name, kw in gi.FIT_CASES.items():
    t = gi.reference_problem(np.float32)
    conv = lambda xs: [tf.convert_to_tensor(x, dtype=np.float32) for x in xs]
    res = ref.fit_gains_and_foregrounds(
        g_r=tf.convert_to_tensor(t["g_r"]), g_i=tf.convert_to_tensor(t["g_i"]), fg_r=conv(t["fg_r"]), fg_i=conv(t["fg_i"]),
        data_r=conv(t["data_r"]), data_i=conv(t["data_i"]), wgts=conv(t["wgts"]), fg_comps=conv(t["fg_comps"]),
        corr_inds=t["corr_inds"], sky_model_r=conv(t["data_r"]), sky_model_i=conv(t["data_i"]), **kw)
    put(f"g3_{name}_loss", np.asarray(res[4]["loss"], dtype=np.float32))
    put(f"g3_{name}_g_r", res[0].numpy())
    put(f"g3_{name}_g_i", res[1].numpy())
    put(f"g3_{name}_fg_r", res[2][0].numpy())
    put(f"g3_{name}_fg_i", res[3][0].numpy())

# ---- G4: tensorize_fg_coeffs + yield_fg_model_array on redundant / multi-slot groups (828-913, 402-444) --------
for use_red in (False, True):
    tensors, corr = ref.tensorize_fg_model_comps_dict(d, ants_map, nf, use_redundancy=use_red, dtype=np.float64)
    sky = gi.random_chunk_data(corr, nf, seed=21)
    w = gi.random_chunk_weights(corr, nf, seed=22)
    coeffs = ref.tensorize_fg_coeffs([tf.convert_to_tensor(x) for x in sky], [tf.convert_to_tensor(x) for x in w], tensors)
    for c, t_ in enumerate(coeffs):
        put(f"g4_r{int(use_red)}_coeffs{c}", t_.numpy())
    put(f"g4_r{int(use_red)}_cube", ref.yield_fg_model_array(6, nf, tensors, coeffs, corr))

# ---- G5: the whole driver on the 6-antenna fixture (calibration.py:963-1331) -----------------------------------
for name, kw in gi.DRIVER_CASES.items():
    uvd, gains = gi.driver_inputs()
    model, resid, gains_out, hist = ref.calibrate_and_model_dpss(uvdata=uvd, gains=gains, **kw)
    put(f"g5_{name}_model", model.data_array)
    put(f"g5_{name}_resid", resid.data_array)
    put(f"g5_{name}_gains", gains_out.gain_array)
    put(f"g5_{name}_loss", np.asarray(hist[0][0]["loss"], dtype=np.float32))

path = os.path.join(ROOT, "tests", "golden", "reference_golden.npz")
np.savez_compressed(path, **out)
print(f"wrote {path}: {len(out)} arrays, {os.path.getsize(path) / 1e3:.0f} kB")
