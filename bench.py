#!/usr/bin/env python
"""Benchmark of the gain-and-foreground fit loop (BASELINE.json: fit iterations/sec + HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hera350] [--reg post_hoc|sum]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU port of the reference's TensorFlow graph

A "step" is one optimizer iteration (forward model, chi^2, gradient, Adamax update; calibration.py:663-668) of one
(time, polarisation) integration.  W warm-up steps are run untimed, then exactly K steps are timed with CUDA
events on the library's stream, bracketed by a barrier + device synchronize, max over ranks.  Inputs are larger
than L2 for hera128/hera350 (2.3 / 26 GB streamed per step vs 126 MB of L2), so no explicit L2 flush is needed.
At N > 1 the baseline groups of the ONE integration are sharded across ranks ("strong" scaling); per iteration the
gain gradient and three scalars are exchanged through NVLink peer memory and reduced inside the update kernels
(`--comm nccl`: NCCL all-reduce instead).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# The contract is ONE JSON line on stdout.  Native libraries (NCCL prints "NCCL version ..." when NCCL_DEBUG is
# VERSION or higher) write to file descriptor 1 directly, so fd 1 is pointed at stderr for the whole run and the
# result line goes out through a private duplicate of the original stdout.
sys.stdout.flush()
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


METRIC = "fit_iterations_per_sec"
UNIT = "it/s"


def measured_hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for row in self.rows:
            parts = [p.strip() for p in row.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(np.max(smax)) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: torch port of the reference's dense TensorFlow graph on a bounded sample of the workload
# ---------------------------------------------------------------------------------------------------
def cpu_port_rate(prob, reg, max_sample_bls, steps, warmup, threads):
    """iterations/sec of the torch-CPU port, extrapolated linearly in the padded basis size when the workload
    is sub-sampled.  Returns (it/s for the FULL workload, description of the sample)."""
    import torch

    from oracle import torch_port as T

    torch.set_num_threads(threads)
    nbls = prob.nbls
    stride = max(1, int(np.ceil(nbls / max_sample_bls)))
    sel = np.arange(0, nbls, stride)
    nvecs = int(prob.ncomp.max())
    nf = prob.nfreqs
    keys = list(prob.comps_dict.keys())
    comps = np.zeros((nvecs, len(sel), 1, nf), dtype=np.float32)
    fg_r = np.zeros((nvecs, len(sel), 1, 1), dtype=np.float32)
    fg_i = np.zeros_like(fg_r)
    for n, b in enumerate(sel):
        basis = prob.comps_dict[keys[b]]
        comps[: basis.shape[1], n, 0] = basis.T
        fg_r[: basis.shape[1], n, 0, 0] = prob.c0_r[prob.coef0[b] : prob.coef0[b + 1]]
        fg_i[: basis.shape[1], n, 0, 0] = prob.c0_i[prob.coef0[b] : prob.coef0[b + 1]]
    corr = [[[(int(prob.ant0[b]), int(prob.ant1[b]))] for b in sel]]
    d_r = [prob.data_r[sel][:, None, :]]
    d_i = [prob.data_i[sel][:, None, :]]
    w = [prob.wgts[sel][:, None, :]]
    tp = T.TorchProblem(prob.g0_r, prob.g0_i, [fg_r], [fg_i], d_r, d_i, w, [comps], corr, optimizer="Adamax",
                        learning_rate=1e-2, model_regularization="sum" if reg == "sum" else None,
                        sky_model_r=d_r, sky_model_i=d_i)
    for _ in range(warmup):
        tp.train_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        tp.train_step()
    dt = (time.perf_counter() - t0) / steps
    frac = len(sel) / nbls
    rate_full = (1.0 / dt) * frac
    sample = (f"{len(sel)} of {nbls} baselines (stride {stride}), dense padded basis [{nvecs},{len(sel)},1,{nf}] f32, "
              f"{steps} timed steps after {warmup} warm-up, {dt * 1e3:.1f} ms/step on the sample; "
              + ("extrapolated linearly in basis size to the full workload" if stride > 1 else "full workload"))
    return rate_full, sample, dt


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    from calamity_b200 import synth

    threads = os.cpu_count() or 1
    prob = synth.make(args.workload)
    steps = max(1, min(args.steps, args.cpu_steps))
    warmup = max(1, min(args.warmup, 2))
    rate, sample, dt = cpu_port_rate(prob, args.reg, args.cpu_sample_bls, steps, warmup, threads)
    lay_sizes = prob.layout().sizes()
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 / rate, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, prob, lay_sizes, world),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample,
                         "note": "torch-CPU op-for-op port of the reference's TensorFlow graph (TensorFlow is not "
                                 "installable in this image); never the reference's own TensorFlow build"},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, prob, sizes, world):
    return {
        "workload": f"synthetic {args.workload}: {prob.nants} antennas, {prob.nbls} baselines x {prob.nfreqs} channels, "
                    f"per-baseline DPSS fit, single integration",
        "optimizer": "Adamax lr=1e-2", "model_regularization": args.reg, "n_d": sizes["n_d"], "n_a_nz": sizes["n_a_nz"],
        "n_c_nz": sizes["n_c_nz"], "b_iter_bytes": sizes["b_iter"],
        "parallelism": "1 GPU" if world == 1 else (
            f"baseline groups sharded over {world} GPUs; per iteration the gain gradient and 3 scalars are exchanged "
            + ("through NVLink peer memory, reduced inside the update kernels (no collective call)"
               if getattr(args, "comm", "peer") == "peer" else "with NCCL all-reduce")),
        "l2": "inputs larger than L2 (no flush)" if sizes["b_iter"] > 4 * 126e6 else "working set near L2 size; "
              "L2-resident workload, HBM fraction not meaningful",
    }


# ---------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hera350", choices=["test6", "hera37", "hera128", "hera350"])
    ap.add_argument("--reg", default="post_hoc", choices=["post_hoc", "sum"])
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--fuse", type=int, default=0)
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gain-gradient exchange fused into the update kernel over NVLink peer memory, or NCCL")
    ap.add_argument("--shared-basis", type=int, default=0, choices=[-1, 0, 1],
                    help="0: groups that share a basis block take the shared-basis kernel (default); -1: stream a private "
                         "copy per group (round-1 path)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    # CPU arm: ~10-20 s of host work -- 2048 of the 61 075 baselines (dense padded basis 1.7 GB), 20 steps
    ap.add_argument("--cpu-sample-bls", type=int, default=2048)
    ap.add_argument("--cpu-steps", type=int, default=20)
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, rank, world)
        return
    if args.warmup < 1 or args.steps < 1:
        raise SystemExit("--steps and --warmup must be >= 1")

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    from calamity_b200 import synth
    from calamity_b200.fitter import FitPlan, comm_init_peer, nccl_unique_id
    from calamity_b200.sharding import make_shard

    t_setup = time.perf_counter()
    prob = synth.make(args.workload)
    full = prob.layout()
    sizes = full.sizes()
    shard = make_shard(full, rank, world)
    plan = FitPlan(shard.layout, device=local_rank, tile_freqs=args.tile, shared_basis=args.shared_basis)
    if world > 1 and args.comm == "peer":
        comm_init_peer(plan, rank, world)
    elif world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        plan.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)
    def pinned(a):  # page-locked host copies: the e2e leg's host->device copies are plain DMA transfers
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    d_r, d_i, w = (pinned(shard.take_baselines(x)) for x in (prob.data_r, prob.data_i, prob.wgts))
    c_r, c_i = pinned(shard.take_coeffs(prob.c0_r)), pinned(shard.take_coeffs(prob.c0_i))
    g0_r, g0_i = pinned(prob.g0_r), pinned(prob.g0_i)
    reg = "sum" if args.reg == "sum" else None
    pr = pi = 0.0
    if reg == "sum":  # priors from the data itself (sky_model=None path, calibration.py:1131-1136)
        pr = float(np.sum(prob.data_r.astype(np.float64) * prob.wgts))
        pi = float(np.sum(prob.data_i.astype(np.float64) * prob.wgts))
    fit_kw = dict(optimizer="Adamax", tol=0.0, learning_rate=1e-2, model_regularization=reg, prior_r_sum=pr,
                  prior_i_sum=pi, use_graph=bool(args.graph), fuse_tail_update=bool(args.fuse), steps_per_sync=max(args.steps, args.warmup) + 1)

    def load_inputs():
        plan.set_integration(d_r, d_i, w)
        plan.set_gains(g0_r, g0_i)
        plan.set_coeffs(c_r, c_i)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    load_inputs()
    setup_s = time.perf_counter() - t_setup

    # ---- warm-up: W untimed steps (the reference's own unrecorded step + W-1 recorded ones)
    plan.fit(maxsteps=args.warmup - 1, **fit_kw)
    load_inputs()  # the timed run starts from the same parameters as a fresh fit would
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed: exactly K steps, inputs resident in HBM
    t0 = time.perf_counter()
    hist, res = plan.fit(maxsteps=args.steps - 1, **fit_kw)
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    assert res["nsteps_total"] == args.steps, res
    loop_ms, heavy_ms = float(res["loop_ms"]), float(res["heavy_ms"])

    # ---- end to end through the public handle with HOST buffers: H2D inputs, K steps, D2H results
    barrier()
    t0 = time.perf_counter()
    load_inputs()
    hist2, res2 = plan.fit(maxsteps=args.steps - 1, **fit_kw)
    g_r, g_i = plan.get_gains()
    co_r, co_i = plan.get_coeffs()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    h2d = 3 * d_r.nbytes + 2 * g0_r.nbytes + 2 * c_r.nbytes
    d2h = 2 * g_r.nbytes + 2 * co_r.nbytes + hist2.nbytes

    times = np.array([loop_ms, heavy_ms, e2e_ms, wall_ms], dtype=np.float64)
    if dist is not None:
        tt = torch.tensor(times, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times = tt.cpu().numpy()
        by = torch.tensor([float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        dist.all_reduce(by)
        h2d, d2h = (float(x) for x in by.cpu().numpy())
    loop_ms, heavy_ms, e2e_ms, wall_ms = (float(x) for x in times)

    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        info = plan.info
        # algorithmic bytes of ONE launch of the fused kernel on this rank: non-padding basis rows once, data_r /
        # data_i / weights once, coefficients once (DESIGN.md "Roofline accounting"); whole-iteration figure is B_iter.
        sh = shard.layout.sizes()
        heavy_bytes = 4 * sh["n_a_nz"] + 12 * sh["n_d"] + 8 * sh["n_c_nz"]
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "heavy_traffic.json")) as f:
                tr = json.load(f).get(args.workload)
            if tr and world == 1 and args.reg != "sum":
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except Exception:
            pass
        heavy_avg_ms = heavy_ms / args.steps
        achieved = heavy_bytes / (heavy_avg_ms * 1e-3) / 1e9 if heavy_avg_ms > 0 else None
        iter_gbs = sizes["b_iter"] / (loop_ms / args.steps * 1e-3) / 1e9 / world
        line = {
            "metric": METRIC, "value": args.steps / (loop_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": loop_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, prob, sizes, world),
            "clocks": clocks,
            "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, "ms_total": e2e_ms,
                    "what": "set_integration + set_gains + set_coeffs from pinned host buffers, K iterations, "
                            "get_gains + get_coeffs + loss history back to the host"},
            "gpu_launches": int(res["kernel_launches"]),
            "roofline": {
                "bound": "hbm", "kernel": f"heavy_kernel<FL={info['tile_freqs'] // 4},SUM={int(reg == 'sum')}>",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": peak_src, "traffic": traffic, "algorithmic_bytes_per_launch": heavy_bytes,
                "avg_launch_ms": heavy_avg_ms, "kernel_share_of_step": heavy_ms / loop_ms if loop_ms > 0 else None,
                "iteration": {"b_iter_bytes": sizes["b_iter"], "per_gpu_gbs": iter_gbs, "frac": iter_gbs / peak},
            },
            "wall_ms_timed_call": wall_ms, "setup_s": setup_s,
            "loss_first_last": [float(hist[0]), float(hist[-1])] if len(hist) else None,
            "plan": {k: int(v) for k, v in info.items()},
        }
        if not args.no_cpu_baseline and world == 1:
            threads = os.cpu_count() or 1
            rate, sample, _ = cpu_port_rate(prob, args.reg, args.cpu_sample_bls, args.cpu_steps, 2, threads)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample}
        else:
            line["cpu_baseline"] = None
        emit(line)
    plan.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
