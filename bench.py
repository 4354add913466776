#!/usr/bin/env python
"""Benchmark of the gain-and-foreground fit loop (BASELINE.json: fit iterations/sec + HBM GB/s vs peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload hera350] [--reg post_hoc|sum]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...        # the CPU port of the reference's TensorFlow graph

A "step" is one optimizer iteration (forward model, chi^2, gradient, Adamax update; calibration.py:663-668) of one
(time, polarisation) integration.  W warm-up steps are run untimed, then exactly K steps are timed with CUDA
events on the library's stream, bracketed by a barrier + device synchronize, max over ranks.

Workloads (BASELINE.json configs):
  hera350 (default, config 4)  ONE integration, 61 075 baselines x 1024 channels.  At N > 1 its baseline groups are
                               sharded across the ranks ("strong" scaling); per iteration the gain gradient and three
                               scalars are exchanged through NVLink peer memory and reduced inside the update kernels
                               (`--comm nccl`: NCCL all-reduce instead).  At N > 1 rank 0 also refits the SAME problem
                               unsharded and the line carries `parity_vs_single_gpu`.
  hera128x60 (config 3)        independent integrations of HERA-128 x 1024 sharing one basis, 8 per GPU (64 >= 60 at 8
                               GPUs), no data-path collective ("weak" scaling); value = integration-iterations / s.
  hera128, hera37, test6, tutorial15   single integrations (tutorial15 = the 105-baseline x 200-channel fit of
                               examples/Calamity_Tutorial.ipynb:1178, the only number the reference publishes).
Inputs are larger than L2 for hera128/hera350 on the streaming path; on the shared-basis path the distinct bases
(66 MB at HERA-350) are L2-resident by design and the per-baseline arrays (1.5 GB per step) are not.
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
PUBLISHED = {"tutorial15": 61.77}  # examples/Calamity_Tutorial.ipynb:1178, Tesla P100, 'sum' regulariser


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
    """Times `steps` train steps of the torch-CPU port (oracle/torch_port.py: the reference's dense zero-padded graph,
    autograd, Keras-style update) on every `stride`-th baseline of the workload.  Returns a dict with the MEASURED rate on
    the sample and, separately, the rate extrapolated linearly in the padded basis size to the full workload."""
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
    # thread count: the fastest of {1, half, all} host threads on two probe steps (tiny tensors run 30x slower on 8 threads
    # than on one; the big samples want them all) -- the CPU arm gets its best configuration
    best = None
    for th in sorted({1, max(1, threads // 2), threads}):
        torch.set_num_threads(th)
        tp.train_step()
        t0 = time.perf_counter()
        tp.train_step()
        dt1 = time.perf_counter() - t0
        if best is None or dt1 < best[0]:
            best = (dt1, th)
    used = best[1]
    torch.set_num_threads(used)
    for _ in range(warmup):
        tp.train_step()
    t0 = time.perf_counter()
    for _ in range(steps):
        tp.train_step()
    dt = (time.perf_counter() - t0) / steps
    frac = len(sel) / nbls
    sample = (f"{len(sel)} of {nbls} baselines (stride {stride}), dense padded basis [{nvecs},{len(sel)},1,{nf}] f32, "
              f"{steps} timed steps after {warmup} warm-up, {dt * 1e3:.1f} ms/step on the sample; "
              + ("value = sample rate x sample fraction (linear in basis size)" if stride > 1 else "full workload"))
    return {"value": (1.0 / dt) * frac, "sample_rate": 1.0 / dt, "sample_fraction": frac, "extrapolated": stride > 1,
            "sample": sample, "steps": steps, "warmup": warmup, "threads": used}


def run_reference_arm(args, rank, world):
    """`--impl reference`: TensorFlow cannot be installed in this image, so the arm times the torch-CPU op-for-op port of
    the reference's graph (kind "port") with all host threads, on a bounded sample of the same workload."""
    if rank != 0:
        return
    from calamity_b200 import synth

    threads = os.cpu_count() or 1
    name = "hera128" if args.workload == "hera128x60" else args.workload
    prob = synth.make(name)
    steps = max(1, min(args.steps, args.cpu_steps))
    warmup = max(1, min(args.warmup, 3))
    r = cpu_port_rate(prob, args.reg, args.cpu_sample_bls, steps, warmup, threads)
    lay_sizes = prob.layout().sizes()
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 / r["value"], "higher_is_better": True,
        "scaling": "weak" if args.workload == "hera128x60" else "strong",
        "vs_baseline": (r["value"] / PUBLISHED[name]) if name in PUBLISHED and args.reg == "sum" else None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args, prob, lay_sizes, world),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"],
                         "sample_rate": r["sample_rate"], "sample_fraction": r["sample_fraction"],
                         "extrapolated": r["extrapolated"],
                         "note": "torch-CPU op-for-op port of the reference's TensorFlow graph (TensorFlow is not "
                                 "installable in this image); never the reference's own TensorFlow build.  The "
                                 "reference's own fit loop executed under oracle/tf_shim is timed in the build "
                                 "container only (profiles/round2_cpu_reference_under_shim.md)"},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def workload_config(args, prob, sizes, world):
    shared = getattr(args, "shared_basis", 0) >= 0
    return {
        "workload": f"synthetic {args.workload}: {prob.nants} antennas, {prob.nbls} baselines x {prob.nfreqs} channels, "
                    f"per-baseline DPSS fit, " + ("8 independent integrations per GPU sharing one basis"
                                                  if args.workload == "hera128x60" else "single integration"),
        "optimizer": "Adamax lr=1e-2", "model_regularization": args.reg, "n_d": sizes["n_d"], "n_a_nz": sizes["n_a_nz"],
        "n_c_nz": sizes["n_c_nz"], "b_iter_bytes": sizes["b_iter"],
        "basis_path": "shared (each distinct basis stored once, calfit_shared.cuh)" if shared else "streaming (private copy per group)",
        "parallelism": "1 GPU" if world == 1 else (
            f"{world} GPUs, independent integrations, no collective" if args.workload == "hera128x60" else
            f"baseline groups sharded over {world} GPUs; per iteration the gain gradient and 3 scalars are exchanged "
            + ("through NVLink peer memory, reduced inside the update kernels (no collective call)"
               if getattr(args, "comm", "peer") == "peer" else "with NCCL all-reduce")),
        "l2": ("per-baseline arrays (>= 1.4 GB per step at hera350) larger than L2, no flush; the distinct bases are "
               "L2-resident by design" if shared else "inputs larger than L2 (no flush)") if sizes["b_iter"] > 4 * 126e6
              else "working set near L2 size; L2-resident workload, HBM fraction not meaningful",
    }


def pinned(a):
    """Page-locked host copy: the e2e leg's host->device copies are plain DMA transfers."""
    import torch

    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def measured_bf16_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        return float(m["bf16_tflops"]), "measured dense bf16, MEASURED_PEAKS.json (burst)"
    except Exception:
        return 2250.0, "nominal dense bf16 (MEASURED_PEAKS.json absent)"


def fma_peak_tflops(clocks):
    import torch

    props = torch.cuda.get_device_properties(torch.cuda.current_device())
    mhz = (clocks or {}).get("sm_max_mhz") or 1965.0
    return props.multi_processor_count * 128 * 2 * mhz * 1e6 / 1e12, f"{props.multi_processor_count} SMs x 128 FP32 lanes x 2 x {mhz:.0f} MHz (nominal; not in MEASURED_PEAKS.json)"


# ---------------------------------------------------------------------------------------------------
def bench_single_integration(args, rank, world, local_rank, dist):
    import torch

    from calamity_b200 import synth
    from calamity_b200.fitter import FitPlan, comm_init_peer, nccl_unique_id
    from calamity_b200.sharding import make_shard

    t_setup = time.perf_counter()
    prob = synth.make(args.workload)
    fdt = np.float64 if args.precision == 64 else np.float32
    full = prob.layout(dtype=fdt)
    if args.precision == 64:
        if world > 1:
            raise SystemExit("--precision 64 runs the generic path on one GPU")
        for name in ("data_r", "data_i", "wgts", "g0_r", "g0_i", "c0_r", "c0_i"):
            setattr(prob, name, getattr(prob, name).astype(np.float64))
    sizes = full.sizes()
    synth_s = time.perf_counter() - t_setup
    shard_mode = "class" if args.shared_basis >= 0 else "cyclic"
    shard = make_shard(full, rank, world, mode=shard_mode)
    t_plan = time.perf_counter()
    plan = FitPlan(shard.layout, device=local_rank, tile_freqs=args.tile, shared_basis=args.shared_basis)
    plan_s = time.perf_counter() - t_plan
    if world > 1 and args.comm == "peer":
        comm_init_peer(plan, rank, world)
    elif world > 1:
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt.copy_(torch.frombuffer(bytearray(nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(idt, 0)
        plan.comm_init(bytes(idt.cpu().numpy().tobytes()), rank, world)

    d_r, d_i, w = (pinned(shard.take_baselines(x)) for x in (prob.data_r, prob.data_i, prob.wgts))
    c_r, c_i = pinned(shard.take_coeffs(prob.c0_r)), pinned(shard.take_coeffs(prob.c0_i))
    g0_r, g0_i = pinned(prob.g0_r), pinned(prob.g0_i)
    # priors from the data itself (sky_model=None path, calibration.py:1131-1136)
    pr = float(np.sum(prob.data_r.astype(np.float64) * prob.wgts))
    pi = float(np.sum(prob.data_i.astype(np.float64) * prob.wgts))

    def fit_kw(reg):
        return dict(optimizer="Adamax", tol=0.0, learning_rate=1e-2, model_regularization="sum" if reg == "sum" else None,
                    prior_r_sum=pr, prior_i_sum=pi, use_graph=bool(args.graph), fuse_tail_update=bool(args.fuse),
                    steps_per_sync=max(args.steps, args.warmup) + 1)

    def load_inputs(p=plan, arrs=None):
        a = arrs or (d_r, d_i, w, g0_r, g0_i, c_r, c_i)
        p.set_integration(a[0], a[1], a[2])
        p.set_gains(a[3], a[4])
        p.set_coeffs(a[5], a[6])

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    load_inputs()
    setup_s = time.perf_counter() - t_setup

    def timed_fit(reg, p=plan, arrs=None):
        """W untimed steps (the reference's own unrecorded step + W-1 recorded ones), reload, K timed steps."""
        p.fit(maxsteps=args.warmup - 1, **fit_kw(reg))
        load_inputs(p, arrs)  # the timed run starts from the same parameters as a fresh fit would
        barrier()
        t0 = time.perf_counter()
        hist, res = p.fit(maxsteps=args.steps - 1, **fit_kw(reg))
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        assert res["nsteps_total"] == args.steps, res
        return hist, res, wall

    sampler = ClockSampler(local_rank)
    sampler.start()
    # ---- timed: exactly K steps, inputs resident in HBM
    hist, res, wall_ms = timed_fit(args.reg)
    g_main = plan.get_gains()
    loop_ms, heavy_ms = float(res["loop_ms"]), float(res["heavy_ms"])

    # ---- end to end through the public handle with HOST buffers: H2D inputs, K steps, D2H results
    barrier()
    t0 = time.perf_counter()
    load_inputs()
    hist2, res2 = plan.fit(maxsteps=args.steps - 1, **fit_kw(args.reg))
    g_r, g_i = plan.get_gains()
    co_r, co_i = plan.get_coeffs()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    h2d = 3 * d_r.nbytes + 2 * g0_r.nbytes + 2 * c_r.nbytes
    d2h = 2 * g_r.nbytes + 2 * co_r.nbytes + hist2.nbytes

    # ---- the other regularisation (the API default is 'sum', the CLI default 'post_hoc'), same plan, same K
    other = "sum" if args.reg != "sum" else "post_hoc"
    hist_o, res_o, _ = timed_fit(other)
    g_other = plan.get_gains()

    times = np.array([loop_ms, heavy_ms, e2e_ms, wall_ms, float(res_o["loop_ms"]), float(res_o["heavy_ms"])], dtype=np.float64)
    if dist is not None:
        tt = torch.tensor(times, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times = tt.cpu().numpy()
        by = torch.tensor([float(h2d), float(d2h)], device="cuda", dtype=torch.float64)
        dist.all_reduce(by)
        h2d, d2h = (float(x) for x in by.cpu().numpy())
    loop_ms, heavy_ms, e2e_ms, wall_ms, loop_o_ms, heavy_o_ms = (float(x) for x in times)

    # ---- N > 1: rank 0 refits the same problem unsharded; every sharded number above is checked against it
    parity = None
    if world > 1:
        if rank == 0:
            single = FitPlan(full, device=local_rank, shared_basis=args.shared_basis)
            arrs = (prob.data_r, prob.data_i, prob.wgts, prob.g0_r, prob.g0_i, prob.c0_r, prob.c0_i)
            parity = {}
            for reg, h_sh, g_sh in ((args.reg, hist, g_main), (other, hist_o, g_other)):
                load_inputs(single, arrs)
                h1, _ = single.fit(maxsteps=args.steps - 1, **fit_kw(reg))
                g1 = single.get_gains()
                lerr = float(np.max(np.abs(h_sh.astype(np.float64) - h1) / np.abs(h1)))
                gerr = float(max(np.max(np.abs(g_sh[0] - g1[0])), np.max(np.abs(g_sh[1] - g1[1]))) / np.max(np.abs(g1[0])))
                parity[reg] = {"loss_rel_err": lerr, "gain_rel_err": gerr, "steps": int(len(h1)) + 1}
            parity["tolerance"] = {"loss_rel_err": 1e-5, "gain_rel_err": 1e-4}
            parity["ok"] = all(v["loss_rel_err"] <= 1e-5 and v["gain_rel_err"] <= 1e-4 for k, v in parity.items()
                               if isinstance(v, dict) and "loss_rel_err" in v)
            single.close()
        barrier()

    line = None
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        info = plan.info
        sh = shard.layout.sizes()
        # algorithmic bytes of ONE basis pass on this rank: non-padding basis rows once per group,
        # data_r / data_i / weights once, coefficients once; whole-iteration figure is B_iter
        esz = 8 if args.precision == 64 else 4
        heavy_bytes = esz * sh["n_a_nz"] + 3 * esz * sh["n_d"] + 2 * esz * sh["n_c_nz"]
        shared_path = info["n_class_slots"] > 0
        tc_path = shared_path and info.get("n_tc_ctas", 0) > 0  # 'sum' too: two launches of the tensor-core kernel per pass
        traffic = None
        try:  # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch, from the committed ncu capture
            with open(os.path.join(ROOT, "profiles", "heavy_traffic.json")) as f:
                tr = json.load(f).get(args.workload + ("_tc" if tc_path else "_shared" if shared_path else ""))
            if tr and world == 1 and args.reg != "sum":
                traffic = tr["dram_bytes_read"] + tr["dram_bytes_write"]
        except Exception:
            pass
        heavy_avg_ms = heavy_ms / args.steps
        achieved = heavy_bytes / (heavy_avg_ms * 1e-3) / 1e9 if heavy_avg_ms > 0 else None
        iter_gbs = sizes["b_iter"] / (loop_ms / args.steps * 1e-3) / 1e9 / world
        kernel = (("shared_tc_kernel<0> (tcgen05.mma kind::tf32, 3-term split; basis pass = this kernel + the streaming kernel's few items)"
                   if args.reg != "sum" else "shared_tc_kernel<1> + <2> (tcgen05.mma kind::tf32, 3-term split; the second launch is the "
                   "regulariser's backward rows)")
                  if tc_path else f"shared_kernel<256|512,NQ={4 if args.reg == 'sum' else 2}> (two shapes, one basis pass)" if shared_path
                  else ("generic forward + backward kernels (float64, two passes over the basis)" if info["generic"]
                        else f"heavy_kernel<FL={info['tile_freqs'] // 4},SUM={int(args.reg == 'sum')}>"))
        line = {
            "metric": METRIC, "value": args.steps / (loop_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": loop_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "dtype": "f64" if args.precision == 64 else "f32", "data": "synthetic",
            "vs_baseline": (args.steps / (loop_ms * 1e-3) / PUBLISHED[args.workload])
            if args.workload in PUBLISHED and args.reg == "sum" else None,
            "config": workload_config(args, prob, sizes, world),
            "clocks": clocks,
            "e2e": {"value": args.steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d / args.steps,
                    "d2h_bytes_per_step": d2h / args.steps, "ms_total": e2e_ms,
                    "what": "set_integration + set_gains + set_coeffs from pinned host buffers, K iterations, "
                            "get_gains + get_coeffs + loss history back to the host"},
            "gpu_launches": int(res["kernel_launches"]),
            "roofline": {
                "bound": "hbm", "kernel": kernel,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": (achieved / peak) if achieved else None,
                "peak_source": peak_src, "traffic": traffic, "algorithmic_bytes_per_launch": heavy_bytes,
                "avg_launch_ms": heavy_avg_ms, "kernel_share_of_step": heavy_ms / loop_ms if loop_ms > 0 else None,
                "iteration": {"b_iter_bytes": sizes["b_iter"], "per_gpu_gbs": iter_gbs, "frac": iter_gbs / peak},
                "dram": ({"bytes_per_launch": traffic, "gbs": traffic / (heavy_avg_ms * 1e-3) / 1e9,
                          "frac": traffic / (heavy_avg_ms * 1e-3) / 1e9 / peak,
                          "what": "the bytes the kernel really moves (ncu, per launch) over the live launch duration"}
                         if traffic and heavy_avg_ms > 0 else None),
                "note": ("ALGORITHMIC bytes are SURVEY.md section 8(d)'s: every group's own basis rows counted once per "
                         "iteration.  The shared-basis kernels do not move them: each DISTINCT basis is kept once "
                         "(L2-resident) and only the per-baseline arrays come from DRAM (`traffic`, `dram`), so achieved / "
                         "peak exceeds 1; what bounds the kernel is under `roofline_fma` (tensor pipe / FP32 FMA pipe).  The "
                         "HBM-streaming kernel of round 1 is re-measured under `streaming`.") if shared_path else None,
            },
            "wall_ms_timed_call": wall_ms, "setup_s": setup_s,
            "setup_breakdown_s": {"synthesis": synth_s, "plan_create_and_basis_upload": plan_s,
                                  "shard_pin_first_load": setup_s - synth_s - plan_s},
            "loss_first_last": [float(hist[0]), float(hist[-1])] if len(hist) else None,
            "plan": {k: int(v) for k, v in info.items()},
        }
        if shared_path:
            tf = 2.0 * info["class_fma"] / (heavy_avg_ms * 1e-3) / 1e12
            if tc_path:
                bf16, bsrc = measured_bf16_peak()
                fpeak, fsrc = bf16 / 2.0, f"tf32 dense = half of the bf16 figure ({bsrc})"
                line["roofline_fma"] = {
                    "bound": "tensor", "kernel": kernel, "achieved": tf, "peak": fpeak, "unit": "TFLOP/s", "frac": tf / fpeak,
                    "peak_source": fsrc, "flops_per_launch": 2 * info["class_fma"], "executed_over_algorithmic": 3.0,
                    "what": "ALGORITHMIC flops (8 per group, basis row and channel: forward + backward contraction, real + "
                            "imaginary part; SURVEY.md section 8d: 8 N_A_nz) over the basis-pass duration.  The tensor cores "
                            "execute three TF32 MMAs per product (hi.hi + hi.lo + lo.hi) on 128-row tiles padded to 16 "
                            "vectors, so the pipe does >= 3x this; the pass is bound by the CUDA-core phase between the two "
                            "contractions and the MMA issue path, see profiles/round2_ncu_hera350.md section 7"}
            else:
                fpeak, fsrc = fma_peak_tflops(clocks)
                line["roofline_fma"] = {"bound": "fp32_fma", "kernel": kernel, "achieved": tf, "peak": fpeak, "unit": "TFLOP/s",
                                        "frac": tf / fpeak, "peak_source": fsrc, "flops_per_launch": 2 * info["class_fma"],
                                        "what": "8 flops per (group, basis row, channel): forward + backward contraction, real + "
                                                "imaginary part (SURVEY.md section 8d: 8 N_A_nz), over the basis-pass duration"}
        line["other_regularization"] = {
            "model_regularization": other, "value": args.steps / (loop_o_ms * 1e-3), "unit": UNIT,
            "ms_per_step": loop_o_ms / args.steps, "basis_pass_ms": heavy_o_ms / args.steps,
            "loss_first_last": [float(hist_o[0]), float(hist_o[-1])] if len(hist_o) else None,
            "what": "same plan, same K steps, the other value of model_regularization ('sum' is the API default, "
                    "calibration.py:986; 'post_hoc' the CLI default, calibration.py:1915)"}
        if parity is not None:
            line["parity_vs_single_gpu"] = parity
    plan.close()

    # ---- the round-1 HBM-streaming path re-measured in the same run (1 GPU, large workloads): private basis copy per group
    if rank == 0 and world == 1 and args.shared_basis >= 0 and not args.no_streaming and line["plan"]["n_class_slots"] > 0 \
            and args.workload in ("hera128", "hera350"):
        t0 = time.perf_counter()
        sp = FitPlan(full, device=local_rank, tile_freqs=args.tile, shared_basis=-1)
        up_s = time.perf_counter() - t0
        load_inputs(sp)
        k2 = min(args.steps, 50)
        sp.fit(maxsteps=max(0, min(args.warmup, 5) - 1), **fit_kw(args.reg))
        load_inputs(sp)
        torch.cuda.synchronize()
        h_s, r_s = sp.fit(maxsteps=k2 - 1, **fit_kw(args.reg))
        sinfo = sp.info
        sp.close()
        peak, _ = measured_hbm_peak()
        hb = 4 * sizes["n_a_nz"] + 12 * sizes["n_d"] + 8 * sizes["n_c_nz"]
        hms = float(r_s["heavy_ms"]) / k2
        line["streaming"] = {
            "value": k2 / (float(r_s["loop_ms"]) * 1e-3), "unit": UNIT, "steps": k2, "ms_per_step": float(r_s["loop_ms"]) / k2,
            "roofline": {"bound": "hbm", "kernel": f"heavy_kernel<FL={sinfo['tile_freqs'] // 4},SUM={int(args.reg == 'sum')}>",
                         "achieved": hb / (hms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                         "frac": hb / (hms * 1e-3) / 1e9 / peak, "avg_launch_ms": hms, "algorithmic_bytes_per_launch": hb},
            "plan_create_and_basis_upload_s": up_s, "device_bytes": int(sinfo["device_bytes"]),
            "loss_first_last": [float(h_s[0]), float(h_s[-1])] if len(h_s) else None,
            "what": "shared_basis=-1: every group streams its own basis copy from HBM once per iteration (the round-1 path)"}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1 and args.precision == 32:
            threads = os.cpu_count() or 1
            r = cpu_port_rate(prob, args.reg, args.cpu_sample_bls, args.cpu_steps, 2, threads)
            line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["threads"], "kind": "port", "sample": r["sample"],
                                    "sample_rate": r["sample_rate"], "sample_fraction": r["sample_fraction"],
                                    "extrapolated": r["extrapolated"]}
        else:
            line["cpu_baseline"] = None
        emit(line)


def bench_integrations(args, rank, world, local_rank, dist):
    """Config 3: independent (time, pol) integrations of HERA-128 x 1024 sharing one basis, sharded over the GPUs with no
    data-path collective (calibration.py:1160-1167: the units are independent).  Every rank runs `--integrations-per-gpu`
    integrations through its own plan: H2D of the integration, K iterations, gains back -- one after the other."""
    import torch

    from calamity_b200 import synth
    from calamity_b200.fitter import FitPlan

    t_setup = time.perf_counter()
    nint = args.integrations_per_gpu
    base = synth.make("hera128")
    full = base.layout()
    sizes = full.sizes()
    plan = FitPlan(full, device=local_rank, tile_freqs=args.tile, shared_basis=args.shared_basis)
    # integrations = noise / gain realisations sharing the basis: per-rank seeds
    rng = np.random.default_rng(777 + rank)
    units = []
    for n in range(nint):
        scale = 1.0 + 0.05 * rng.standard_normal((base.nbls, 1)).astype(np.float32)
        units.append(tuple(pinned(x) for x in (base.data_r * scale, base.data_i * scale, base.wgts, base.g0_r, base.g0_i,
                                               base.c0_r, base.c0_i)))
    reg = "sum" if args.reg == "sum" else None
    pr = float(np.sum(base.data_r.astype(np.float64) * base.wgts))
    pi = float(np.sum(base.data_i.astype(np.float64) * base.wgts))
    kw = dict(optimizer="Adamax", tol=0.0, learning_rate=1e-2, model_regularization=reg, prior_r_sum=pr, prior_i_sum=pi,
              steps_per_sync=max(args.steps, args.warmup) + 1, use_graph=bool(args.graph))

    def run_unit(u, steps):
        plan.set_integration(u[0], u[1], u[2])
        plan.set_gains(u[3], u[4])
        plan.set_coeffs(u[5], u[6])
        hist, res = plan.fit(maxsteps=steps - 1, **kw)
        g = plan.get_gains()
        return hist, res, g

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
            torch.cuda.synchronize()

    setup_s = time.perf_counter() - t_setup
    run_unit(units[0], max(args.warmup, 3))
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t0 = time.perf_counter()
    loop_ms = heavy_ms = 0.0
    launches = 0
    last = None
    for u in units:
        hist, res, g = run_unit(u, args.steps)
        loop_ms += float(res["loop_ms"])
        heavy_ms += float(res["heavy_ms"])
        launches += int(res["kernel_launches"])
        last = hist
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3
    clocks = sampler.stop()
    times = np.array([loop_ms, heavy_ms, e2e_ms], dtype=np.float64)
    if dist is not None:
        tt = torch.tensor(times, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        times = tt.cpu().numpy()
    loop_ms, heavy_ms, e2e_ms = (float(x) for x in times)
    if rank == 0:
        peak, peak_src = measured_hbm_peak()
        info = plan.info
        total_steps = world * nint * args.steps
        h2d = sum(a.nbytes for a in units[0]) * nint * world
        d2h = (2 * units[0][3].nbytes + 4 * args.steps) * nint * world
        heavy_bytes = 4 * sizes["n_a_nz"] + 12 * sizes["n_d"] + 8 * sizes["n_c_nz"]
        hms = heavy_ms / (nint * args.steps)
        achieved = heavy_bytes / (hms * 1e-3) / 1e9 if hms > 0 else 0.0
        line = {
            "metric": METRIC, "value": total_steps / (loop_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": loop_ms / (nint * args.steps), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args, base, sizes, world), integrations_per_gpu=nint,
                           integrations_total=nint * world,
                           unit_of_work="integration-iterations per second, summed over all GPUs; a step = one "
                                        "iteration of one integration"),
            "clocks": clocks,
            "e2e": {"value": total_steps / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d / total_steps,
                    "d2h_bytes_per_step": d2h / total_steps, "ms_total": e2e_ms,
                    "what": "per integration: set_integration + set_gains + set_coeffs from pinned host buffers, K "
                            "iterations, gains + loss history back"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "kernel": "shared_kernel<MS=64,NQ=2>" if info["n_class_slots"] else "heavy_kernel",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                         "traffic": None, "algorithmic_bytes_per_launch": heavy_bytes, "avg_launch_ms": hms,
                         "kernel_share_of_step": heavy_ms / loop_ms},
            "setup_s": setup_s, "loss_first_last": [float(last[0]), float(last[-1])],
            "plan": {k: int(v) for k, v in info.items()}, "cpu_baseline": None,
        }
        if info["n_class_slots"]:
            fpeak, fsrc = fma_peak_tflops(clocks)
            tf = 2.0 * info["class_fma"] / (hms * 1e-3) / 1e12 if hms > 0 else 0.0
            line["roofline_fma"] = {"bound": "fp32_fma", "achieved": tf, "peak": fpeak, "unit": "TFLOP/s", "frac": tf / fpeak,
                                    "peak_source": fsrc}
        emit(line)
    plan.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="hera350", choices=["test6", "tutorial15", "hera37", "hera128", "hera350", "hera128x60"])
    ap.add_argument("--reg", default="post_hoc", choices=["post_hoc", "sum"])
    ap.add_argument("--tile", type=int, default=0)
    ap.add_argument("--graph", type=int, default=0)
    ap.add_argument("--fuse", type=int, default=0)
    ap.add_argument("--comm", default="peer", choices=["peer", "nccl"],
                    help="N > 1: gain-gradient exchange fused into the update kernel over NVLink peer memory, or NCCL")
    ap.add_argument("--shared-basis", type=int, default=0, choices=[-1, 0, 1],
                    help="0: groups that share a basis block take the shared-basis kernel (default); -1: stream a private "
                         "copy per group (round-1 path)")
    ap.add_argument("--no-streaming", action="store_true", help="skip the re-measurement of the streaming path")
    ap.add_argument("--precision", type=int, default=32, choices=[32, 64],
                    help="64: the reference's --precision 64 / dtype=np.float64 (generic unfused device path, single GPU)")
    ap.add_argument("--integrations-per-gpu", type=int, default=8)
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
    if args.workload == "hera128x60":
        bench_integrations(args, rank, world, local_rank, dist)
    else:
        bench_single_integration(args, rank, world, local_rank, dist)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
