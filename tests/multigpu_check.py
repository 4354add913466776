"""Run under torchrun on >= 2 GPUs: a sharded fit (NCCL all-reduce of the gain gradient inside the native loop)
must reproduce the single-GPU loss history.  Prints PASS/FAIL; used by tests/test_gpu_multigpu.py."""
import faulthandler
import os
import sys

import numpy as np

faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def log(*a):
    print(f"[rank {os.environ.get('RANK', '?')}]", *a, file=sys.stderr, flush=True)


def main():
    import torch
    import torch.distributed as dist

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from calamity_b200 import synth
    from calamity_b200.fitter import FitPlan, comm_init_peer, nccl_unique_id
    from calamity_b200.sharding import make_shard

    workload = sys.argv[1] if len(sys.argv) > 1 else "hera37"
    reg = sys.argv[2] if len(sys.argv) > 2 else "none"
    comm = sys.argv[3] if len(sys.argv) > 3 else "peer"
    scenario = sys.argv[4] if len(sys.argv) > 4 else "parity"
    prob = synth.make(workload, init_gain_scatter=0.02, coeff_error=0.05)
    full = prob.layout()
    shard = make_shard(full, rank, world)
    log("shard groups", len(shard.groups), "of", full.ngroups)
    plan = FitPlan(shard.layout, device=local)
    if comm == "peer":  # NVLink peer-memory exchange fused into the update kernels
        comm_init_peer(plan, rank, world)
    else:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        log("got id")
        plan.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    log("comm ok", comm)
    plan.set_integration(*(shard.take_baselines(x) for x in (prob.data_r, prob.data_i, prob.wgts)))
    plan.set_gains(prob.g0_r, prob.g0_i)
    plan.set_coeffs(shard.take_coeffs(prob.c0_r), shard.take_coeffs(prob.c0_i))
    pr = float(np.sum(prob.data_r.astype(np.float64) * prob.wgts))
    pi = float(np.sum(prob.data_i.astype(np.float64) * prob.wgts))
    kw = dict(optimizer="Adamax", maxsteps=40, tol=0.0, learning_rate=1e-2, model_regularization=None if reg == "none" else "sum",
              prior_r_sum=pr, prior_i_sum=pi)
    if scenario == "timeout":
        # rank 1 asks for fewer steps than rank 0: rank 0 must come back with CALB2_ERR_TIMEOUT (-7), not hang its GPU
        from calamity_b200._native import NativeError

        kw["maxsteps"] = 40 if rank == 0 else 8
        t0 = __import__("time").perf_counter()
        try:
            plan.fit(**kw)
            outcome = "returned"
        except NativeError as e:
            outcome = "timeout" if "(-7)" in str(e) else f"other error: {e}"
        dt = __import__("time").perf_counter() - t0
        log("fit", outcome, f"{dt:.1f} s")
        plan.close()
        want = "timeout" if rank == 0 else "returned"
        ok = outcome == want and dt < 30.0
        print(f"{'PASS' if ok else 'FAIL'} rank {rank} scenario=timeout outcome={outcome} ({dt:.1f} s)", flush=True)
        res = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(res, op=dist.ReduceOp.MIN)
        dist.destroy_process_group()
        sys.exit(0 if int(res.item()) == 1 else 1)
    hist, res = plan.fit(**kw)
    g_r, g_i = plan.get_gains()
    torch.cuda.synchronize()
    plan.close()  # runs the exchange's last publish / wait round: no rank unmaps a buffer a peer still reads
    log("sharded fit done", hist[:2], hist[-1])
    ok = True
    if rank == 0:
        single = FitPlan(full, device=local)
        single.set_integration(prob.data_r, prob.data_i, prob.wgts)
        single.set_gains(prob.g0_r, prob.g0_i)
        single.set_coeffs(prob.c0_r, prob.c0_i)
        h1, _ = single.fit(**kw)
        g1, _ = single.get_gains()
        single.close()
        err = float(np.max(np.abs(hist.astype(np.float64) - h1) / h1))
        gerr = float(np.max(np.abs(g_r - g1)) / np.max(np.abs(g1)))
        ok = err < 1e-5 and gerr < 1e-4
        print(f"{'PASS' if ok else 'FAIL'} world={world} {workload} reg={reg} comm={comm} loss_rel_err={err:.2e} gain_rel_err={gerr:.2e} "
              f"launches={res['kernel_launches']}", flush=True)
    # every rank must hold identical gains (replicated state)
    t = torch.from_numpy(g_r.copy()).cuda()
    t0 = t.clone()
    dist.broadcast(t0, 0)
    same = bool(torch.equal(t, t0))
    if not same:
        print(f"FAIL rank {rank}: gains differ from rank 0", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if (ok and same) else 1)


if __name__ == "__main__":
    main()
