"""Development aid: the tensor-core shape (CALB2_TC=1, default) against the CUDA-core shapes (CALB2_TC=0) and the float64 oracle."""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(name, nsel):
    from calamity_b200 import synth
    from calamity_b200.fitter import FitPlan
    from oracle.ragged import RaggedProblem
    from tests.helpers import flat_from_synth, rel_err

    prob = synth.make(name, init_gain_scatter=0.05, coeff_error=0.1)
    no_oracle = nsel < 0  # timing / stamps of the full problem only
    if nsel > 0:
        prob = prob.select_baselines(np.arange(0, prob.nbls, max(1, prob.nbls // nsel)))
    p = flat_from_synth(prob)
    F = np.float64
    if not no_oracle:
        rp = RaggedProblem(p.lay)
        args = [np.asarray(x, dtype=F) for x in (p.g0_r, p.g0_i, p.c0_r, p.c0_i, p.data_r, p.data_i, p.wgts)]
        ol, ogr, ogi, ocr, oci = rp.loss_and_grads(*args)
    print("creating plan", flush=True)
    plan = FitPlan(p.lay, device=0, shared_basis=1)
    print("plan created", {k: plan.info[k] for k in ("n_tc_ctas", "n_tc_slots", "n_class_ctas", "nitems")}, flush=True)
    plan.set_integration(p.data_r, p.data_i, p.wgts)
    plan.set_gains(p.g0_r, p.g0_i)
    plan.set_coeffs(p.c0_r, p.c0_i)
    info = plan.info
    print("inputs set; loss_and_grads ...", flush=True)
    import ctypes
    import threading

    from calamity_b200 import _native as nat

    def watchdog():  # the native call blocks in a CUDA sync if the kernel hangs: print the kernel's progress record
        import time

        for _ in range(3):
            time.sleep(8)
            rec = (ctypes.c_uint32 * 16)()
            nat.load().calb2_debug_tc_record(rec)
            print("   [watchdog] tc record:", list(rec), flush=True)

    threading.Thread(target=watchdog, daemon=True).start()
    loss, dgr, dgi, dcr, dci = plan.loss_and_grads()
    if no_oracle:
        ol, ogr, ogi, ocr, oci = loss, dgr, dgi, dcr, dci
    print(f"{name}: tc ctas {info['n_tc_ctas']} tc slots {info['n_tc_slots']}/{info['nslots_total']}  loss {float(loss):.7e} oracle {float(ol):.7e} "
          f"rel {abs(float(loss) - float(ol)) / abs(float(ol)):.2e}  grads g {rel_err(dgr, ogr):.2e} {rel_err(dgi, ogi):.2e} "
          f"c {rel_err(dcr, ocr):.2e} {rel_err(dci, oci):.2e}", flush=True)
    if os.environ.get("CALB2_TC_PROF"):
        NS = 24
        buf = (ctypes.c_int64 * (32 * NS))()
        nat.check(nat.load().calb2_debug_tc_profile(plan._handle, buf, 32 * NS))
        a = np.array(buf, dtype=np.int64).reshape(32, NS)
        t0 = a[0, 20]
        rel = lambda xs: [int(x - t0) if x else -1 for x in xs]
        print(f"   CTA {os.environ['CALB2_TC_PROF']}: start 0, prologue done {rel([a[0, 21]])[0]}, end {rel([a[0, 22]])[0]} (cycles)")
        print("   mma[before F(j+1), after issue, after wait q, after B issue, after refill] | q warp 0 [before wait v, after, V loaded, "
              "computed, arrived] | arrival of the 8 phase-Q warps")
        for j in range(32):
            if a[j, 0] == 0:
                break
            print(f"   tile {j:2d}: mma {rel(a[j, 0:5])}  q {rel(a[j, 6:11])}  arrivals {rel(a[j, 12:20])}")
    hist, res = plan.fit(optimizer="Adamax", maxsteps=30, tol=0.0, learning_rate=1e-2)
    if no_oracle:
        print(f"   30 steps: loop {res['loop_ms']:.2f} ms, basis pass {res['heavy_ms']:.2f} ms", flush=True)
        plan.close()
        return
    o = rp.fit(p.g0_r, p.g0_i, p.c0_r, p.c0_i, p.data_r, p.data_i, p.wgts, optimizer="Adamax", maxsteps=30, tol=0.0, learning_rate=1e-2)
    ref = np.asarray(o[4]["loss"], dtype=F)
    print(f"   30-step trajectory: max rel loss err {np.max(np.abs(hist - ref) / ref):.2e}; gains {rel_err(plan.get_gains()[0], o[0]):.2e} "
          f"coeffs {rel_err(plan.get_coeffs()[0], o[2]):.2e}; loop {res['loop_ms']:.2f} ms, basis pass {res['heavy_ms']:.2f} ms", flush=True)
    plan.close()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    else:
        for tc in ("0", "1"):
            for name, nsel in (("hera37", 0), ("hera128", 1500)):
                env = dict(os.environ, CALB2_TC=tc)
                print(f"--- CALB2_TC={tc}", flush=True)
                r = subprocess.run([sys.executable, __file__, name, str(nsel)], env=env, timeout=60)
                if r.returncode != 0:
                    print("   FAILED rc", r.returncode, flush=True)
