"""Development aid: the tensor-core pass under CUDA-graph replay (small problems replay a captured graph of iterations) against the
same fit launched kernel by kernel, both regularisations."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from calamity_b200 import synth  # noqa: E402
from calamity_b200.fitter import FitPlan  # noqa: E402

prob = synth.make("hera37", init_gain_scatter=0.02, coeff_error=0.05)
for reg in (None, "sum"):
    outs = []
    for graph in (False, True):
        plan = FitPlan(prob.layout(), device=0)
        plan.set_integration(prob.data_r, prob.data_i, prob.wgts)
        plan.set_gains(prob.g0_r, prob.g0_i)
        plan.set_coeffs(prob.c0_r, prob.c0_i)
        pr, pi = plan.prior_sums(prob.data_r, prob.data_i)
        hist, res = plan.fit(optimizer="Adamax", maxsteps=100, tol=0.0, learning_rate=1e-2, model_regularization=reg,
                             prior_r_sum=0.9 * pr, prior_i_sum=1.1 * pi, use_graph=graph)
        outs.append((hist, plan.get_gains()[0], plan.info["n_tc_ctas"], plan.info["nitems"], res["kernel_launches"]))
        plan.close()
    same = np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    print(f"reg={reg}: tc ctas {outs[0][2]}, streaming items {outs[0][3]}, launches {outs[0][4]} / {outs[1][4]}, "
          f"graph replay bit-identical to direct launches: {same}; loss {outs[0][0][0]:.6e} -> {outs[0][0][-1]:.6e}")
    assert same
print("graph + tensor-core pass OK")
