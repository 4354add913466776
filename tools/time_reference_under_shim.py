#!/usr/bin/env python
"""Build container only (needs /root/reference): time the reference's OWN fit loop -- the unmodified
`calamity.calibration.fit_gains_and_foregrounds` (calibration.py:447-738) -- executed under oracle/tf_shim (torch-CPU
standing in for TensorFlow), on the configurations it can hold: the 6-antenna test case, the tutorial's 15-antenna
105-baseline x 200-channel case (examples/Calamity_Tutorial.ipynb:1178: 61.77 it/s on a Tesla P100) and HERA-37 x 384.
Prints a markdown table (profiles/round2_cpu_reference_under_shim.md).  bench.py cannot do this on the GPU box:
/root/reference does not travel.  kind = "ref-under-shim": the reference's code and loop, the shim's arithmetic.
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import tf_shim  # noqa: E402

tf_shim.activate()
import tensorflow as tf  # noqa: E402  (the shim)
import torch  # noqa: E402
from calamity import calibration as ref  # noqa: E402

from calamity_b200 import synth  # noqa: E402
from tests.helpers import reference_tensors  # noqa: E402


def time_case(name, steps, threads, reg):
    prob = synth.make(name, init_gain_scatter=0.02, coeff_error=0.05)
    t = reference_tensors(prob, np.float32)
    conv = lambda xs: [tf.convert_to_tensor(x, dtype=np.float32) for x in xs]
    torch.set_num_threads(threads)
    kw = dict(optimizer="Adamax", maxsteps=steps, tol=0.0, learning_rate=1e-2, model_regularization=reg)
    args = dict(g_r=tf.convert_to_tensor(t["g_r"]), g_i=tf.convert_to_tensor(t["g_i"]), fg_r=conv(t["fg_r"]),
                fg_i=conv(t["fg_i"]), data_r=conv(t["data_r"]), data_i=conv(t["data_i"]), wgts=conv(t["wgts"]),
                fg_comps=conv(t["fg_comps"]), corr_inds=t["corr_inds"], sky_model_r=conv(t["data_r"]),
                sky_model_i=conv(t["data_i"]))
    t0 = time.perf_counter()
    res = ref.fit_gains_and_foregrounds(**args, **kw)
    dt = time.perf_counter() - t0
    loss = res[4]["loss"]
    return (steps + 1) / dt, float(loss[0]), float(loss[-1]), prob


if __name__ == "__main__":
    cores = os.cpu_count() or 1
    print("| workload | baselines x channels | model_regularization | threads | steps | it/s (reference loop under tf_shim) | loss first -> last |")
    print("|---|---|---|---|---|---|---|")
    for name, steps in (("test6", 200), ("tutorial15", 200), ("hera37", 60)):
        for reg in ("sum", "post_hoc"):
            best = None
            for th in sorted({1, cores}):
                rate, l0, l1, prob = time_case(name, steps, th, reg)
                if best is None or rate > best[0]:
                    best = (rate, th, l0, l1)
            print(f"| {name} | {prob.nbls} x {prob.nfreqs} | {reg} | {best[1]} | {steps} (+1 warm-up) | {best[0]:.1f} | {best[2]:.3e} -> {best[3]:.3e} |", flush=True)
