#!/usr/bin/env python
"""bench.py -- configs/sec of the fused hot path (g-curve + collision/distance + cost).

    python bench.py --gpus N --steps K --warmup W            # our arm (libctrb200, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: the oracle port on all host threads

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): 3-tube STATIC model,
tube_lengths [33,66,99] at delta_x = 1 (100 arc-length nodes on the longest tube), against
colon_straight.STL (11 492 triangles) posed mid-lumen.  A "step" is one pass of the hot path over one
batch: per-configuration equilibrium solve (MINPACK hybrd, 30 unknowns) + section curves, chain-vs-mesh
distance, goal distance, own cost, and the best-node reduction (two 8-byte NCCL all-reduces when N > 1).
The kinematic-model variant of the same chain (g-curve positions never leave the SM) is
`--workload c3_kinematic_*` and is always reported under "extra" together with the configs[1] g-curve
kernel, the eta / follow-the-leader kernel, the other mesh and one 10^7-configuration static step.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for what each key means.
"""
import argparse
import json
import os
import pathlib
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = pathlib.Path(__file__).resolve().parent
for p in (ROOT, ROOT / "ctr-design-and-path-plan_b200", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

BASES, DOF = ["linear"] * 3, [2, 3, 3]
Q = [0.02, 0.001, 0.05, 0.0001, 0.002, 0.03, 0.0002, 0.0005]
LENGTHS, DX, RAD = [33, 66, 99], 1, [1.2, 0.9, 0.6]
NPTS = [34, 67, 100]
GOAL_WEIGHT = 3.0
ENV_FILE = ROOT / "tests" / "fixtures" / "env_c3_colon_straight.json"
ENV_NEW_FILE = ROOT / "tests" / "fixtures" / "env_colon_straight_new.json"
RIN = [1.0, 0.7, 0.4]
WORKLOADS = {  # name -> (model, exposure lo, hi, default configs per step per GPU); SURVEY 8d "shallow" / "uniform"
    "c3_static_colon_straight_shallow": ("static", 1, 8, 10_000_000),
    "c3_static_colon_straight_uniform": ("static", 1, 33, 1 << 18),
    "c3_kinematic_colon_straight_shallow": ("kinematic", 1, 8, 10_000_000),
    "c3_kinematic_colon_straight_uniform": ("kinematic", 1, 33, 10_000_000),
}
BYTES_PER_CONFIG = 3 * (4 + 8) + (8 + 8 + 8 + 1)  # idx + theta in, obstacle_min + goal_dist + cost + bit out
# DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of each hot kernel at bench scale, from the
# committed `ncu --set full` summary of this very command line; used only when workload and launch size match
NCU_BENCH_SCALE = ROOT / "profiles" / "r2_ncu_bench_scale.json"


def workload_description(kind):
    """model / environment strings shared by both arms' `config` (the reference arm runs the same workload on the CPU)"""
    model = (f"3-tube {kind} (linear bases, q_dof [2,3,3]), tube_lengths [33,66,99], delta_x 1"
             + (", last_strain_base degree 2, 30 unknowns, MINPACK hybrd xtol 1.49e-8" if kind == "static" else ""))
    env = "colon_straight.STL (11492 triangles) translated to a mid-lumen pose + Sphere(5) goal"
    return model, env


def make_inputs(B, seed, lo, hi):
    rng = np.random.default_rng(seed)
    e = np.stack([rng.integers(lo, min(hi, NPTS[t] - 1) + 1, B) for t in range(3)], 1)
    idx = (np.asarray(NPTS)[None, :] - 1 - e).astype(np.int32)
    theta = rng.uniform(0, 2 * np.pi, (B, 3))
    return idx, theta


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(workload, B):
    """{kernel: DRAM bytes of one launch} from the committed bench-scale capture, or None when it does not apply"""
    if not NCU_BENCH_SCALE.exists():
        return None
    d = json.loads(NCU_BENCH_SCALE.read_text())
    if d.get("workload") != workload or int(d.get("configs_per_step", -1)) != int(B):
        return None
    return d.get("dram_bytes_per_launch")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.t.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def oracle_objects(kind="kinematic", env_file=ENV_FILE):
    from oracle import oracle as orc
    if kind == "static":
        om = orc.make_static(BASES, DOF, Q, LENGTHS, DX, RAD, RIN, 2)
    else:
        om = orc.make_model(BASES, DOF, Q, LENGTHS, DX)
    oenv = orc.Env.from_json(env_file)
    return orc, om, oenv


def time_oracle(om, oenv, idx, theta, threads):
    """the reference path on the CPU: Static.solve_g's literal hybrd (or Kinematic.solve_g) -> check_collision -> cost"""
    from oracle import oracle as orc
    t0 = time.perf_counter()
    if isinstance(om, orc.OrcStatic):
        obs = orc.static_eval_batch(oenv, om, idx, theta, RAD, GOAL_WEIGHT, nthreads=threads)[0]
    else:
        obs = oenv.eval_batch(om, idx, theta, RAD, GOAL_WEIGHT, nthreads=threads)[0]
    return time.perf_counter() - t0, obs


def run_reference(args, rank):
    """CPU arm: the reference's own path is pure Python over python-fcl (not installable offline), so the
    timed implementation is the C oracle port of it, on every host thread this process may use."""
    if rank != 0:
        return
    kind, lo, hi, _ = WORKLOADS[args.workload]
    threads = len(os.sched_getaffinity(0))
    _, om, oenv = oracle_objects(kind)
    if args.ref_sample <= 0:
        args.ref_sample = 16384 if kind == "static" else 65536
    idx, theta = make_inputs(args.ref_sample, 1003, lo, hi)
    for _ in range(args.warmup):
        time_oracle(om, oenv, idx, theta, threads)
    times = []
    for _ in range(args.steps):
        dt, obs = time_oracle(om, oenv, idx, theta, threads)
        times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = args.ref_sample / (ms * 1e-3)
    sample = f"{args.ref_sample} configs/step of {args.workload} (seed 1003), collide fraction {(obs < 0).mean():.3f}"
    model_s, env_s = workload_description(kind)
    line = {"impl": "reference", "metric": "configs/sec (g-curve + collision + cost)", "value": value,
            "unit": "configs/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": {"workload": args.workload, "configs_per_step": args.ref_sample,
                                            "model": model_s, "environment": env_s, "exposure_per_tube": [lo, hi],
                                            "collide_fraction": float((obs < 0).mean()), "host_threads": threads,
                                            "parallelism": f"{threads} host threads (pthreads), bounded sample of the workload"},
            "cpu_baseline": {"value": value, "unit": "configs/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "configs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def static_flop_model(idx_h, nfev_mean):
    """the structured marching count of the reference's residual -- 450 FLOP per section-1 node, 200 per
    section-2 node, 40 per section-3 node, per residual evaluation -- plus 260 FLOP per output node of integrate_g."""
    e = np.asarray(NPTS)[None, :] - 1 - idx_h
    nodes = e + 1                                     # nodes of sections 1, 2, 3 (three tubes)
    per_eval = float((450.0 * nodes[:, 0] + 200.0 * nodes[:, 1] + 40.0 * nodes[:, 2]).mean())
    out_nodes = float(nodes.sum(1).mean())
    return nfev_mean * per_eval + 260.0 * out_nodes, per_eval, out_nodes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=0, help="configurations per step and per GPU (0 = workload default)")
    ap.add_argument("--workload", default="c3_static_colon_straight_shallow", choices=list(WORKLOADS))
    ap.add_argument("--ref-sample", type=int, default=0, help="configs per step of the CPU arm (0 = workload default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary workloads / kernels report")
    ap.add_argument("--timed-only", action="store_true", help="profiling runs: warm-up + timed steps only (no anatomy / e2e / extras)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    import ctypes as C
    from ctrdapp_b200 import _lib as L
    from ctrdapp_b200 import dist as cdist
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.model.kinematic import Kinematic

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libctrb200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # one non-default stream for everything: the library's kernels, torch's events and the NCCL collectives order on it
    # (the legacy default stream's handle is 0, which ctrb_ctx_set_stream reads as "the ctx's own stream")
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx = L.Context(local_rank)
    ctx.set_stream(stream.cuda_stream)
    from ctrdapp_b200.model.static import Static
    qs = [np.array(Q[0:2]), np.array(Q[2:5]), np.array(Q[5:8])]
    mk = Kinematic(3, qs, DOF, LENGTHS, DX, BASES, ctx=ctx)
    ms = Static(3, qs, DOF, LENGTHS, DX, BASES, {"outer": RAD, "inner": RIN}, 2, ctx=ctx)
    cc = CollisionChecker(ENV_FILE, ctx=ctx)
    kind, lo, hi, default_B = WORKLOADS[args.workload]
    m = ms if kind == "static" else mk
    B = args.batch or default_B
    idx_h, theta_h = make_inputs(B, 1003 + rank, lo, hi)
    host_threads = len(os.sched_getaffinity(0))

    # ---- device-resident arm: inputs already in HBM
    idx_d = torch.from_numpy(idx_h).to(dev)
    theta_d = torch.from_numpy(theta_h).to(dev)
    obs_d = torch.empty(B, dtype=torch.float64, device=dev)
    goal_d = torch.empty(B, dtype=torch.float64, device=dev)
    cost_d = torch.empty(B, dtype=torch.float64, device=dev)
    coll_d = torch.empty(B, dtype=torch.uint8, device=dev)
    rad = np.ascontiguousarray(RAD, np.float64)
    cd = L.CostDesc(L.COST_TYPES["square_obstacle_avg_plus_weighted_goal"], GOAL_WEIGHT, 1, 0.0)
    lib = L.lib()

    info_d = torch.empty(B, dtype=torch.int32, device=dev)
    nfev_d = torch.empty(B, dtype=torch.int32, device=dev)
    key_d = torch.empty(2, dtype=torch.int64, device=dev)

    def eval_any(model, n, idx_p, theta_p, outs, mem, checker=None):
        """one pass of the hot path over n configurations (device or host buffers)"""
        o, g, c, cl, inf, nf = outs
        env = (checker or cc).handle
        if getattr(model, "num_unknowns", None):
            L.check(lib.ctrb_static_eval_batch(ctx.handle, model.handle, env, C.byref(cd), n, L.ptr(idx_p),
                                               L.ptr(theta_p), L.ptr(rad), L.ptr(o), L.ptr(g), L.ptr(c), L.ptr(cl),
                                               L.ptr(inf), L.ptr(nf), mem))
        else:
            L.check(lib.ctrb_eval_batch(ctx.handle, model.handle, env, C.byref(cd), n, L.ptr(idx_p), L.ptr(theta_p),
                                        L.ptr(rad), L.ptr(o), L.ptr(g), L.ptr(c), L.ptr(cl), mem))

    dev_outs = (obs_d, goal_d, cost_d, coll_d, info_d, nfev_d)

    def step():
        """the hot path over this rank's batch + the best node of the whole job; nothing here waits for the device"""
        eval_any(m, B, idx_d, theta_d, dev_outs, L.MEM_DEVICE)
        L.check(lib.ctrb_argmin_cost_device(ctx.handle, B, L.ptr(cost_d), rank * B, L.ptr(key_d)))
        cdist.reduce_best(key_d)        # N > 1: two 8-byte all-reduce(min) over NCCL, on the same stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 1)):
        step()
    barrier()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    best = cdist.decode_best(key_d)
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = world * B / (ms_per_step * 1e-3)

    if args.timed_only:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms_per_step, "timed_only": True}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- step anatomy: the library's CUDA events around each hot kernel, on its stream (three more steps, each
    # read back before the next; the timed region above never waits for the device)
    names = ["static_solve", "static_fast", "static_qr", "static_curves", "chain_eval"] if kind == "static" else ["chain_eval"]
    anatomy = {n: [] for n in names}
    fallbacks = []
    for _ in range(3):
        step()
        for n in names:
            anatomy[n].append(ctx.last_kernel_ms(n))
        if kind == "static":
            fallbacks.append(ctx.static_fallbacks())
    kms = {n: float(np.mean(v)) for n, v in anatomy.items()}
    dom = "static_solve" if kind == "static" else "chain_eval"
    k_ms = kms[dom]
    per_rank = None
    if world > 1:  # every rank's own step time, dominant-kernel time and SM clock: shows which GPU sets the maximum
        mine = torch.tensor([ms_total / args.steps, k_ms, float(clocks.get("sm_mhz") or 0.0)], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"ms_per_step": [round(float(a[0]), 2) for a in allr], "dominant_kernel_ms": [round(float(a[1]), 2) for a in allr],
                    "sm_mhz": [float(a[2]) for a in allr]}
        # the per-configuration costs of the whole job on every rank (what a planner sharded over ranks asks for):
        # one all_gather of B doubles per rank over NCCL, checked against the reduced best node
        torch.cuda.synchronize()
        tg0 = time.perf_counter()
        allc = cdist.gather_costs(cost_d, device=dev)
        torch.cuda.synchronize()
        gather_ms = 1e3 * (time.perf_counter() - tg0)
        fin = torch.where(torch.isfinite(allc), allc, torch.full_like(allc, float("inf")))
        gmin, gidx = torch.min(fin, 0)
        per_rank["cost_gather"] = {"bytes_per_rank": int(B * 8), "ms": gather_ms,
                                   "matches_reduced_best": bool(abs(float(gmin) - best[0]) == 0.0 and int(gidx) == best[1])}
    work = ctx.work_counters()
    collide_frac = float((obs_d < 0).double().mean().item())

    # ---- e2e arm: HOST (pinned) buffers through the C ABI, copies inside the timed region, the same K steps
    pin = lambda a: torch.from_numpy(a).pin_memory().numpy()  # noqa: E731
    idx_p, theta_p = pin(idx_h), pin(theta_h)
    obs_p, goal_p, cost_p = (pin(np.empty(B)) for _ in range(3))
    coll_p = pin(np.empty(B, np.uint8))
    info_p, nfev_p = pin(np.empty(B, np.int32)), pin(np.empty(B, np.int32))

    def step_host():
        eval_any(m, B, idx_p, theta_p, (obs_p, goal_p, cost_p, coll_p, info_p, nfev_p), L.MEM_HOST)
        return float(np.fmin.reduce(cost_p))  # best node of the step, read on the host (NaN = colliding, ignored)

    step_host()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_host()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * B / float(te.item())
    assert np.array_equal(obs_p < 0, (obs_d < 0).cpu().numpy())  # host-buffer and device-buffer paths agree

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel sequence, timed live with CUDA events on its stream
    peak, peak_src = measured_peaks()
    f32_peak, f64_peak = ctx.measure_fma_peak(32), ctx.measure_fma_peak(64)
    per = 1.0 / max(work["configs"], 1)
    traffic_tab = ncu_traffic(args.workload, B)
    if kind == "static":
        nfev_mean = float(nfev_d.double().mean().item())
        conv_frac = float((info_d == 1).double().mean().item())
        flop_cfg, flop_eval, out_nodes = static_flop_model(idx_h, nfev_mean)
        achieved = flop_cfg * B / (k_ms * 1e-3) / 1e12
        # algorithmic HBM bytes of the solve: inputs, info/nfev, the unknowns it hands to the curves kernel (written and read
        # back once) and the node positions the collision kernel reads (24 B / node)
        bpc = 3 * 12 + 8 + 2 * 8 * ms.num_unknowns + 24.0 * out_nodes
        fb = float(np.mean(fallbacks)) / B
        traffic = None
        if traffic_tab:
            traffic = float(sum(traffic_tab.get(k) or 0.0 for k in ("static_fast_kernel", "static_solve_kernel", "static_curves_pos_kernel")))
        roofline = {"bound": "fp64_fma", "achieved": achieved, "peak": f64_peak, "unit": "TFLOP/s", "frac": achieved / f64_peak,
                    "traffic": traffic,
                    "traffic_source": (str(NCU_BENCH_SCALE.relative_to(ROOT)) + ": dram__bytes_read.sum + dram__bytes_write.sum of one launch "
                                       "of each of the three kernels at this launch size") if traffic is not None else None,
                    "kernel": "static_fast_kernel + static_solve_kernel + static_curves_pos_kernel (one launch sequence = Static.solve_g)",
                    "kernel_ms": k_ms, "fast_kernel_ms": kms["static_fast"], "qr_fallback_kernel_ms": kms["static_qr"],
                    "curves_kernel_ms": kms["static_curves"], "qr_fallback_fraction": fb,
                    "flop_per_config": flop_cfg, "flop_per_residual_evaluation": flop_eval,
                    "residual_evaluations_per_config": nfev_mean, "output_nodes_per_config": out_nodes,
                    "flop_model": "SURVEY 8(d) marching count of the reference's residual: (450 x section-1 nodes + 200 x section-2 nodes "
                                  "+ 40 x section-3 nodes) per evaluation x evaluations the kernel reports + 260 per output node",
                    "peak_source": "ctrb_measure_fma_peak(64) in this run: fma_peak_kernel (profiles/README.md) -- MEASURED_PEAKS.json "
                                   "holds no fp64 figure",
                    "hbm": {"achieved_gbs": bpc * B / (k_ms * 1e-3) / 1e9, "peak_gbs": peak, "frac": bpc * B / (k_ms * 1e-3) / 1e9 / peak,
                            "algorithmic_bytes_per_config": bpc, "peak_source": peak_src},
                    "note": "a per-configuration nonlinear solve, one warp each: bound by fp64 CUDA-core issue and latency (dependent "
                            "chains of 27-element operations), not by HBM -- the hbm figure is given for completeness"}
        # the kernel's own work model (closed-form residual: 12 node integrands instead of a march)
        n_jac = max(1.0, nfev_mean / 45.0)
        flop_exec = nfev_mean * 2600.0 + n_jac * 2 * 30 ** 3 + max(nfev_mean - 30 * n_jac, 1.0) * 9 * 30 ** 2 + out_nodes * 260.0
        compute = {"residual_evaluations_per_config": nfev_mean, "converged_fraction": conv_frac,
                   "executed_fp64_flop_per_config_model": flop_exec, "executed_fp64_tflops": flop_exec * B / (k_ms * 1e-3) / 1e12,
                   "fp64_fma_peak_tflops_measured": f64_peak, "fp32_fma_peak_tflops_measured": f32_peak,
                   "collision_kernel_ms": kms["chain_eval"], "collision_share_of_step": kms["chain_eval"] / ms_per_step,
                   "solve_share_of_step": k_ms / ms_per_step,
                   "bvh_nodes_per_config": work["bvh_nodes"] * per, "exact_tests_per_config": work["exact_tests"] * per,
                   "executed_flop_model": "2.6 kFLOP/residual evaluation (closed forms) + hybrd dense algebra (n = 30) + 260/output node"}
    else:
        achieved = BYTES_PER_CONFIG * B / (k_ms * 1e-3) / 1e9
        flop32 = (work["bvh_nodes"] * 30.0 + work["prefilter_tests"] * 60.0) * per
        flop64 = work["exact_tests"] * 350.0 * per
        t32 = flop32 * B / (k_ms * 1e-3) / 1e12
        roofline = {"bound": "fp32_fma", "achieved": t32, "peak": f32_peak, "unit": "TFLOP/s", "frac": t32 / f32_peak,
                    "traffic": float(traffic_tab["chain_eval_group_kernel<true>"]) if traffic_tab and "chain_eval_group_kernel<true>" in traffic_tab else None,
                    "kernel": "chain_eval_group_kernel<true>", "kernel_ms": k_ms,
                    "flop_model": "SURVEY 8(d): 30 FLOP per BVH node visited + 60 per point-triangle prefilter (fp32), from the kernel's work counters",
                    "hbm": {"achieved_gbs": achieved, "peak_gbs": peak, "frac": achieved / peak,
                            "algorithmic_bytes_per_config": BYTES_PER_CONFIG, "peak_source": peak_src},
                    "note": "data-dependent tree walk bound by L1 hit latency (DESIGN.md 11); neither the fp32 pipe nor HBM is near its limit"}
        compute = {"bvh_nodes_per_config": work["bvh_nodes"] * per, "prefilter_tests_per_config": work["prefilter_tests"] * per,
                   "exact_tests_per_config": work["exact_tests"] * per, "fp32_flop_per_config": flop32,
                   "fp64_flop_per_config": flop64, "fp32_tflops_achieved": t32,
                   "fp64_tflops_achieved": flop64 * B / (k_ms * 1e-3) / 1e12, "fp32_fma_peak_tflops_measured": f32_peak,
                   "fp64_fma_peak_tflops_measured": f64_peak}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import static_parity as sp
        orc, om, oenv = oracle_objects(kind)
        n0 = 2048 if kind == "static" else 8192
        dt, _ = time_oracle(om, oenv, idx_h[:n0], theta_h[:n0], host_threads)
        n = int(min(B, max(n0, 15.0 / max(dt / n0, 1e-9))))
        dt, o_obs = time_oracle(om, oenv, idx_h[:n], theta_h[:n], host_threads)
        agree = float(np.mean((o_obs < 0) == (obs_p[:n] < 0)))
        note = f"collision-bit agreement with the GPU on these configurations: {agree:.4f}"
        if kind == "static":
            # where the disagreements come from (tests/test_gpu_headline.py asserts this breakdown): per-bucket comparison of
            # the GPU's unknowns with the oracle's under the product's one-step rule, and of the oracle with itself
            n2 = min(n, 32768)
            res = ms.solve_g_batch(idx_h[:n2], theta_h[:n2], want_g=False)
            o1 = orc.static_batch(om, idx_h[:n2], theta_h[:n2], env=oenv, rad=RAD, goal_weight=GOAL_WEIGHT, one_step_rule=1)
            x0 = orc.static_init_guess(om)
            x0p = x0[None, :] * (1.0 + 1e-13 * np.random.default_rng(7).standard_normal((n2, x0.size)))
            o1p = orc.static_batch(om, idx_h[:n2], theta_h[:n2], one_step_rule=1, x0=x0p)
            short = sp.one_step(idx_h[:n2], NPTS)
            bg = sp.buckets(res["solution"], res["info"], o1["solution"], o1["info"])
            bo = sp.buckets(o1p["solution"], o1p["info"], o1["solution"], o1["info"])
            parts = []
            for nm, sel in (("well-posed", ~short), ("one-step sections", short)):
                if sel.any():
                    rg, ro = sp.rates(bg, sel), sp.rates(bo, sel)
                    parts.append(f"{nm} ({int(sel.sum())}): same root {rg['A']:.4f}, other root {rg['B']:.4f}, flag differs {rg['F']:.4f}, "
                                 f"neither converged {rg['N']:.4f} [oracle vs itself under a 1e-13 perturbation: {ro['A']:.4f} / {ro['B']:.4f} / "
                                 f"{ro['F']:.4f} / {ro['N']:.4f}]")
            a = bg["A"]
            bit_a = float(np.mean((o1["obs"][a] < 0) == (obs_p[:n2][a] < 0))) if a.any() else float("nan")
            note += (f"; the timed oracle is the reference's literal hybrd, whose treatment of one-step sections is decided by rounding. "
                     f"Against the oracle under the product's one-step rule, first {n2} configurations: " + "; ".join(parts)
                     + f"; collision bit where the root is the same: {bit_a:.5f}")
        cpu = {"value": n / dt, "unit": "configs/s", "cores": host_threads, "kind": "port",
               "sample": f"first {n} configs of the timed batch ({dt:.1f} s, oracle C port of the reference path, pthreads; {note})"}

    extra = {}
    if not args.no_extra:
        extra = run_extra(args, ctx, dev, ms, mk, cc, eval_any, dev_outs, B, peak, lib, C, L, torch, host_threads)

    line = {"metric": "configs/sec (g-curve + collision + cost)", "value": value, "unit": "configs/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": args.workload, "configs_per_step_per_gpu": B,
                       "model": workload_description(kind)[0], "environment": workload_description(kind)[1],
                       "exposure_per_tube": [lo, hi], "collide_fraction": collide_frac, "host_threads": host_threads,
                       "l2": f"per step {BYTES_PER_CONFIG * B / 1e6:.0f} MB of inputs+outputs"
                             + (f", {8 * 30 * B / 1e6:.0f} MB of unknowns and ~{24 * 17 * B / 1e6:.0f} MB of node positions" if kind == "static" else "")
                             + " stream through HBM (L2 is 126 MB): inputs larger than L2, no explicit flush",
                       "parallelism": f"dp{world} (configs sharded, best node by two 8-byte all-reduces over NCCL)" if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "configs/s", "h2d_bytes_per_step": B * 3 * 12, "d2h_bytes_per_step": B * 33 if kind == "static" else B * 25,
                    "steps": args.steps,
                    "path": ("ctrb_static_eval_batch" if kind == "static" else "ctrb_eval_batch (chunked copy/compute pipeline)")
                            + "(CTRB_MEM_HOST) on pinned numpy buffers, best node read on the host"},
            "gpu_launches": int(launches), "roofline": roofline, "compute_roofline": compute,
            "kernel_ms": kms, "best_node": {"cost": best[0], "index": best[1]}, "bvh": cc.stats()}
    if per_rank:
        line["per_rank"] = per_rank
    if cpu:
        line["cpu_baseline"] = cpu
    if extra:
        line["extra"] = extra
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_extra(args, ctx, dev, ms, mk, cc, eval_any, dev_outs, B, peak, lib, C, L, torch, host_threads):
    """secondary figures (not bench lines): the other exposure distribution and model, the other mesh, one 10^7-configuration
    static step, BASELINE configs[0] and configs[1], and the eta / follow-the-leader kernel"""
    from ctrdapp_b200.collision import CollisionChecker
    from ctrdapp_b200.model.kinematic import Kinematic
    from ctrdapp_b200.model.static import Static
    extra = {}

    def timed_eval(model, checker, Bx, l2, h2, reps=2, seed=1003):
        ih, th = make_inputs(Bx, seed, l2, h2)
        idd, thd = torch.from_numpy(ih).to(dev), torch.from_numpy(th).to(dev)
        outs = tuple(torch.empty(Bx, dtype=t.dtype, device=dev) for t in dev_outs)
        for _ in range(reps):
            ctx.mark(0)
            eval_any(model, Bx, idd, thd, outs, L.MEM_DEVICE, checker)
            ctx.mark(1)
            ms_ = ctx.elapsed_ms()
        r = {"configs_per_s": Bx / (ms_ * 1e-3), "configs": Bx, "ms": ms_,
             "collide_fraction": float((outs[0] < 0).double().mean().item())}
        if getattr(model, "num_unknowns", None):
            r["converged_fraction"] = float((outs[4] == 1).double().mean().item())
            r["residual_evaluations_per_config"] = float(outs[5].double().mean().item())
        return r, ih, th

    for name, (k2, l2, h2, _) in WORKLOADS.items():
        if name == args.workload:
            continue
        r, ih, th = timed_eval(ms if k2 == "static" else mk, cc, min(B, 1 << 18) if k2 == "static" else 2_000_000, l2, h2)
        if k2 == "static" and not args.no_cpu_baseline:
            # the oracle on a sample of the same batch: its converged fraction and evaluation count beside the GPU's
            orc, om, _ = oracle_objects("static")
            o = orc.static_batch(om, ih[:4096], th[:4096], one_step_rule=1)
            o0 = orc.static_batch(om, ih[:4096], th[:4096], one_step_rule=0)
            r["oracle_sample"] = {"configs": 4096, "converged_fraction": float((o["info"] == 1).mean()),
                                  "residual_evaluations_per_config": float(o["nfev"].mean()),
                                  "literal_hybrd_converged_fraction": float((o0["info"] == 1).mean()),
                                  "literal_hybrd_residual_evaluations_per_config": float(o0["nfev"].mean())}
        extra[name] = r
    # SURVEY 8(d) "also report": the author's re-posed mesh colon_straight_new.STL (free-space regime: few configurations collide)
    cc_new = CollisionChecker(ENV_NEW_FILE, ctx=ctx)
    for tag, l2, h2, Bx in (("shallow", 1, 8, 1 << 18), ("uniform", 1, 33, 1 << 17)):
        extra[f"c3_static_colon_straight_new_{tag}"] = timed_eval(ms, cc_new, Bx, l2, h2)[0]
        extra[f"c3_kinematic_colon_straight_new_{tag}"] = timed_eval(mk, cc_new, 2_000_000, l2, h2)[0]
    # the headline at smaller steps (round 1 and the first half of round 2 stepped 2^20 configurations)
    if args.workload != "c3_static_colon_straight_shallow" or B != 1 << 20:
        extra["c3_static_colon_straight_shallow_2^20"] = timed_eval(ms, cc, 1 << 20, 1, 8)[0]
    if args.workload != "c3_static_colon_straight_shallow" or B != 1 << 22:
        extra["c3_static_colon_straight_shallow_2^22"] = timed_eval(ms, cc, 1 << 22, 1, 8)[0]
    # BASELINE configs[0] (C1): configuration/config.yaml as a STATIC model, 1024 random configurations against the
    # translated box of SURVEY 8d's C1b variant (against in_simple.STL itself every configuration collides at the base)
    q1 = [np.array([0.02, 0.001]), np.array([0.05, 0.0001, 0.002])]
    ms1 = Static(2, q1, [2, 3], [120, 120], 1, ["linear", "linear"], {"outer": [0.9, 0.6], "inner": [0.7, 0.4]}, 2, ctx=ctx)
    cc1 = CollisionChecker(ROOT / "tests" / "fixtures" / "env_c1b_in_simple.json", ctx=ctx)
    cd = L.CostDesc(L.COST_TYPES["square_obstacle_avg_plus_weighted_goal"], GOAL_WEIGHT, 1, 0.0)
    rng1 = np.random.default_rng(1001)
    i1 = torch.from_numpy(rng1.integers(0, 120, (1024, 2)).astype(np.int32)).to(dev)
    t1 = torch.from_numpy(rng1.uniform(0, 2 * np.pi, (1024, 2))).to(dev)
    o1 = tuple(torch.empty(1024, dtype=t.dtype, device=dev) for t in dev_outs)
    rad1 = np.ascontiguousarray([0.9, 0.6])
    for rep in range(3):
        ctx.mark(0)
        L.check(lib.ctrb_static_eval_batch(ctx.handle, ms1.handle, cc1.handle, C.byref(cd), 1024, L.ptr(i1), L.ptr(t1),
                                           L.ptr(rad1), L.ptr(o1[0]), L.ptr(o1[1]), L.ptr(o1[2]), L.ptr(o1[3]),
                                           L.ptr(o1[4]), L.ptr(o1[5]), L.MEM_DEVICE))
        ctx.mark(1)
        ms_c1 = ctx.elapsed_ms()
    extra["c1_static_two_tube_1024"] = {"configs_per_s": 1024 / (ms_c1 * 1e-3), "ms": ms_c1, "configs": 1024,
                                        "collide_fraction": float((o1[0] < 0).double().mean().item()),
                                        "converged_fraction": float((o1[4] == 1).double().mean().item()),
                                        "note": "one launch sequence over 1024 configurations: latency, not throughput"}
    # BASELINE configs[1] (C2): the g-curve kernel, 10^6 configurations, fp64 and fp32
    m2 = Kinematic(2, [np.array([0.02, 0.001]), np.array([0.05, 0.0001, 0.002])], [2, 3], [120, 120], 1,
                   ["linear", "linear"], ctx=ctx)
    rng = np.random.default_rng(1002)
    B2 = 1_000_000
    i2 = torch.from_numpy(rng.integers(0, 121, (B2, 2)).astype(np.int32)).to(dev)
    t2 = torch.from_numpy(rng.uniform(0, 2 * np.pi, (B2, 2))).to(dev)
    off = torch.empty(B2 * 2 + 1, dtype=torch.int64, device=dev)
    L.check(lib.ctrb_g_node_offsets(ctx.handle, m2.handle, B2, L.ptr(i2), 0, L.ptr(off), L.MEM_DEVICE))
    total = int(off[-1].item())
    for prec, dt_ in ((64, torch.float64), (32, torch.float32)):
        g = torch.empty((total, 4, 4), dtype=dt_, device=dev)
        for rep in range(3):
            ctx.mark(0)
            L.check(lib.ctrb_solve_g_batch(ctx.handle, m2.handle, B2, L.ptr(i2), L.ptr(t2), 0, prec, L.ptr(off),
                                           L.ptr(g), L.MEM_DEVICE))
            ctx.mark(1)
            ms = ctx.elapsed_ms()
        gbs = g.numel() * g.element_size() / (ms * 1e-3) / 1e9
        extra[f"c2_gcurve_fp{prec}"] = {"configs_per_s": B2 / (ms * 1e-3), "ms": ms, "nodes": total,
                                        "hbm_write_gbs": gbs, "frac_of_hbm_peak": gbs / peak}
        del g
    # eta / follow-the-leader (Kinematic.solve_eta, kinematic.py:132-198): RRT* calls it up to 2 x 15 times per iteration.
    # Parents = the C2 curves just written (resident in HBM); every candidate moves each tube by one node and 0.05 rad.
    # Per parent node the kernel reads the 3 x 4 block of g (96 of the node's 128 bytes) and writes a 6-vector (48 B).
    prev_idx = i2.clamp(max=119)                      # a parent node exists for every tube (the reference raises IndexError otherwise)
    L.check(lib.ctrb_g_node_offsets(ctx.handle, m2.handle, B2, L.ptr(prev_idx), 0, L.ptr(off), L.MEM_DEVICE))
    total = int(off[-1].item())
    g = torch.empty((total, 4, 4), dtype=torch.float64, device=dev)
    L.check(lib.ctrb_solve_g_batch(ctx.handle, m2.handle, B2, L.ptr(prev_idx), L.ptr(t2), 0, 64, L.ptr(off), L.ptr(g), L.MEM_DEVICE))
    new_idx = (prev_idx - 1).clamp(min=0).contiguous()
    dth = torch.full((B2, 2), 0.05, dtype=torch.float64, device=dev)
    eta = torch.empty((B2, 2, 6), dtype=torch.float64, device=dev)
    ftl = torch.empty((total, 6), dtype=torch.float64, device=dev)
    mag = torch.empty((B2, 4), dtype=torch.float64, device=dev)
    for rep in range(3):
        L.check(lib.ctrb_eta_ftl_batch(ctx.handle, m2.handle, B2, L.ptr(prev_idx), L.ptr(new_idx), None, L.ptr(dth), L.ptr(g),
                                       L.ptr(off), L.ptr(eta), L.ptr(ftl), L.ptr(mag), L.MEM_DEVICE))
        ms = ctx.last_kernel_ms("eta_ftl")
    assert ctx.device_status() == 0
    alg = total * (96 + 48) + B2 * (2 * (4 + 4 + 8) + 2 * 48 + 32)
    extra["c2_eta_ftl_fp64"] = {"configs_per_s": B2 / (ms * 1e-3), "ms": ms, "parent_nodes": total,
                                "algorithmic_bytes": alg, "hbm_gbs": alg / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak,
                                "bytes_model": "96 B read + 48 B written per parent node (SURVEY 8d) + per-candidate inputs and eta",
                                "moved_bytes_upper_bound": total * (128 + 48),
                                "note": "the 4x4 nodes are 128 B apart, so the sectors holding the unused last row are fetched too"}
    return extra


if __name__ == "__main__":
    main()
