"""Run with CALB2_GUARD=1 (tests/test_gpu_guards.py does): every device allocation of the library then carries guard zones;
after a tour of the entry points on both basis paths -- ragged sizes that do not divide the tile shapes, redundant slots,
joint groups, float64 -- no kernel may have written outside its buffers.  compute-sanitizer is closed on this GPU pool
(profiles/round2_sanitizer_closed.log); this is the library's own bounds check."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
assert os.environ.get("CALB2_GUARD") == "1"

from calamity_b200 import _native as nat  # noqa: E402
from calamity_b200 import synth  # noqa: E402
from calamity_b200.fitter import FitPlan  # noqa: E402
from calamity_b200.layout import RaggedLayout  # noqa: E402
from tests.helpers import flat_from_synth, mixed_problem  # noqa: E402


def check(tag):
    nb, nv = C.c_int64(), C.c_int64()
    nat.check(nat.load().calb2_debug_check_guards(C.byref(nb), C.byref(nv)))
    print(f"  {tag}: {nb.value} live device buffers, {nv.value} guard zones written", flush=True)
    assert nb.value > 20 and nv.value == 0, tag


def tour(p, tag, **plan_kw):
    plan = FitPlan(p.lay, device=0, **plan_kw)
    plan.set_integration(p.data_r, p.data_i, p.wgts)
    plan.set_gains(p.g0_r, p.g0_i)
    plan.init_coeffs(p.data_r, p.data_i)
    plan.set_coeffs(p.c0_r, p.c0_i)
    pr, pi = plan.prior_sums(p.data_r, p.data_i)
    for reg in (None, "sum"):
        plan.loss_and_grads(model_regularization=reg, prior_r_sum=float(pr), prior_i_sum=float(pi))
        plan.fit(optimizer="Adamax", maxsteps=12, tol=0.0, learning_rate=1e-2, model_regularization=reg, prior_r_sum=float(pr),
                 prior_i_sum=float(pi), use_min=True)
    plan.get_model()
    plan.apply_model_snr_weights()
    plan.fit(optimizer="Adam", maxsteps=6, tol=0.0, learning_rate=1e-2, freeze_model=True)
    plan.fit(optimizer="Adamax", maxsteps=70, tol=0.0, learning_rate=1e-2, use_graph=True, steps_per_sync=16)
    check(tag)
    plan.close()


if __name__ == "__main__":
    p37 = flat_from_synth(synth.make("hera37", init_gain_scatter=0.02, coeff_error=0.05))
    tour(p37, "hera37 streaming", shared_basis=-1)
    tour(p37, "hera37 shared-basis (every class, ragged tiles)", shared_basis=1)
    p6 = flat_from_synth(synth.make("test6", init_gain_scatter=0.02, coeff_error=0.05, flag_fraction=0.1))
    tour(p6, "test6 (200 channels: padded last tile) shared", shared_basis=1)
    tour(p6, "test6 tile 16", shared_basis=-1, tile_freqs=16)
    pm = mixed_problem(nants=40, nfreqs=200, seed=9, n_dpss_bls=150, joint=((5, 9, 70), (3, 6, 181)))
    tour(pm, "mixed: joint groups stream, DPSS classes share (hybrid plan)")
    # the large-class shape (kp > 160) and a class that fills several tiles
    sub = synth.make("hera128", init_gain_scatter=0.02, coeff_error=0.05)
    sub = sub.select_baselines(np.argsort(-sub.ncomp, kind="stable")[:700])
    tour(flat_from_synth(sub), "hera128: 700 longest baselines x 1024 channels")
    # float64: the generic path
    lay64 = synth.make("test6").layout()
    lay64.dtype = np.dtype(np.float64)
    lay64.blocks = [np.ascontiguousarray(b, dtype=np.float64) for b in lay64.blocks]
    p64 = flat_from_synth(synth.make("test6", init_gain_scatter=0.02, coeff_error=0.05))
    p64.lay = lay64
    for name in ("data_r", "data_i", "wgts", "g0_r", "g0_i", "c0_r", "c0_i"):
        setattr(p64, name, getattr(p64, name).astype(np.float64))
    plan = FitPlan(lay64, device=0)
    plan.set_integration(p64.data_r, p64.data_i, p64.wgts)
    plan.set_gains(p64.g0_r, p64.g0_i)
    plan.set_coeffs(p64.c0_r, p64.c0_i)
    plan.fit(optimizer="Adamax", maxsteps=10, tol=0.0, learning_rate=1e-2)
    plan.close()
    print("PASS", flush=True)
