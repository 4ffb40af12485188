#!/usr/bin/env python
"""bench.py -- routed segment-timesteps/sec of the B200 routing path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (config.workload): BASELINE.json configs[2] -- the synthetic CONUS-scale forest the metric is quoted on:
2,729,077 segments in 14,713 basins (largest ~50 %), NHD-like confluences (SURVEY.md 8d in-degree mix), MC-only,
288 x 300 s steps, dependent upstream flows (assume_short_ts = False, the reference default).  One bench "step" = ONE
routing call = all 288 timesteps of all segments (786 M segment-timesteps).  Synthetic parameters/forcing, cold start; see
troute_b200/synth.py.  --workload tree is BASELINE configs[1] (binary tree, 1,048,576 segments); --levelpools N adds
reservoirs (configs[4]).

  set-up    (not timed, once per network as in production where the handle is cached across calls): flatten + upload of the
            network, one calibration call that records the secant trip count of every segment, rebuild with the segments of
            every wavefront level ordered by it (config.within_level_order; --no-trip-order skips it).
  value     segment-timesteps/s with forcing and state already resident in HBM: K x (flow-state reset + routing kernels +
            result pass) timed with CUDA events on the launching stream.  Working set (flow state 9.4 GB + result 9.4 GB)
            is far larger than the 126 MB L2, so no explicit flush between iterations.
  e2e       the same metric through the C-ABI call a T-Route maintainer would bind (trt_route): pinned HOST qlat / q0 in,
            pinned HOST [n, 3*nsteps] result out, copies inside the timed region (trt_route overlaps the result copies
            with the kernels by time chunks).
  roofline  dominant kernel (the dataflow kernel over the wide levels): algorithmic bytes (68 B per segment-timestep,
            SURVEY.md 8d / DESIGN.md) x segment-timesteps of that launch / its CUDA-event duration (events recorded by the
            engine on its stream around that launch), against MEASURED_PEAKS.json hbm_gbs; traffic = ncu dram bytes of
            that kernel (profiles/traffic.json).
  cpu_baseline  the oracle's platform-libm build (C restatement of the Fortran; the reference cannot be compiled here: no
            Fortran compiler) on a bounded sample of the same workload, one core, rank 0, N = 1 only.

--impl reference times that CPU restatement with every host thread, decomposed the way the reference's parallel modes do
(orders of sub-networks, jobs of an order in parallel), on a bounded sample.

N > 1: torchrun, one rank per GPU; the network is sharded by sub-basin (troute_b200/partition.py), total work fixed ->
"scaling": "strong"; cut-edge flows cross GPUs as peer-memory stores inside the kernels (no collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "t-route_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np

BYTES_PER_SEGSTEP = 68.0       # 12 write + 8 own state + 32 params + 4 qlat + 8 gather + 4 index
DT = 300.0
QTS = 12


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="conus", choices=["conus", "tree", "diffusive"],
                    help="conus / tree: Muskingum-Cunge routing (BASELINE configs[2] / configs[1]); diffusive: a batch of "
                         "diffusive-wave mainstem domains (configs[3] kernel, see run_diffusive)")
    ap.add_argument("--domains", type=int, default=296, help="diffusive: independent tailwater domains per call")
    ap.add_argument("--mainstem", type=int, default=24, help="diffusive: mainstem reaches per domain")
    ap.add_argument("--style", default="nhd", choices=["nhd", "hack"],
                    help="basin generator of the conus workload: nhd = NHD-like confluences (SURVEY.md 8d in-degree mix), "
                         "hack = main stems with dozens of tributaries per node (gather stress case)")
    ap.add_argument("--segments", type=int, default=0, help="override the segment count (testing only)")
    ap.add_argument("--nsteps", type=int, default=288, help="routing timesteps per call")
    ap.add_argument("--short-ts", type=int, default=0)
    ap.add_argument("--mode", type=int, default=4)
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--levelpools", type=int, default=0,
                    help="replace this many in-line segments by level-pool reservoirs (BASELINE config 5)")
    ap.add_argument("--deep-lanes", type=int, default=0,
                    help="segments per GPU that march (deepest levels); 0 = 8192 on 1 GPU, 4096 on 2, 2048 on 4+ (one lane per "
                         "warp: the main stem is the critical path once the wide levels are spread over many GPUs)")
    ap.add_argument("--no-trip-order", action="store_true",
                    help="skip the calibration call that orders the segments of a level by their secant trip counts")
    ap.add_argument("--trip-buckets", type=int, default=0,
                    help="time slices of the calibration call's trip counts (0 = network.TRIP_BUCKETS, 1 = totals only)")
    ap.add_argument("--sharded-trip-order", action="store_true",
                    help="N > 1: calibrate and re-order every shard like the single-GPU run does (opt-in, unmeasured)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-seconds", type=float, default=20.0)
    return ap.parse_args()


def _cached(key, make):
    """Synthetic topologies take ~10-30 s of host time to generate; keep them in /tmp between bench invocations."""
    d = os.path.join("/tmp", "trt_synth_cache")
    f = os.path.join(d, key + ".npy")
    try:
        if os.path.exists(f):
            return np.load(f)
    except Exception:
        pass
    a = make()
    try:
        os.makedirs(d, exist_ok=True)
        tmp = f + f".{os.getpid()}.tmp.npy"
        np.save(tmp, a)
        os.replace(tmp, f)
    except Exception:
        pass
    return a


def deep_lanes_for(args, world):
    return args.deep_lanes or {1: 8192, 2: 4096}.get(world, 2048)


def build_workload(args):
    from troute_b200 import synth
    if args.workload == "conus":
        n = args.segments or 2_729_077
        basins = max(1, int(round(14_713 * n / 2_729_077)))
        down = _cached(f"conus_{args.style}_{n}_{basins}_16",
                       lambda: synth.conus_like(n_total=n, n_basins=basins, seed=16, style=args.style))
        name = (f"synthetic CONUS-scale forest ({args.style}-style basins), {n} segments / {basins} basins, MC-only, "
                f"{args.nsteps} x 300 s")
    else:
        n = args.segments or 1_048_576
        down = synth.binary_tree(n)
        name = f"synthetic balanced binary tree, {n} segments, MC-only, {args.nsteps} x 300 s"
    params = synth.channel_params(down, dt=DT, seed=16)
    qlat = synth.lateral_inflow(n, args.nsteps, QTS, seed=16)
    q0 = np.zeros((n, 3), dtype=np.float32)
    up_ptr, up_rows = synth.upstream_csr(down)
    kind = np.zeros(n, dtype=np.uint8)
    lp_rows, wbody = np.zeros(0, np.int64), np.zeros((0, 11))
    if args.levelpools:
        rng = np.random.default_rng(23)
        cand = np.nonzero(np.diff(up_ptr) > 0)[0]
        lp_rows = np.sort(rng.choice(cand, size=min(args.levelpools, cand.size), replace=False)).astype(np.int64)
        kind[lp_rows] = 1
        wbody = synth.levelpool_params(lp_rows.size, seed=16)
        name += f", {lp_rows.size} level-pool reservoirs"
    return dict(name=name, n=n, down=down, params=params, cols=synth.PARAM_COLS, qlat=qlat, q0=q0, up_ptr=up_ptr,
                up_rows=up_rows, kind=kind, lp_rows=lp_rows, wbody=wbody)


# ---------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ---------------------------------------------------------------------------------------------------
class Clocks:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the oracle (C restatement of the reference's Fortran + Cython loop), platform-libm arithmetic
# ---------------------------------------------------------------------------------------------------
def cpu_arm(wl, args, threads, budget_s):
    """Time the CPU path on a bounded sample: the FULL network for `ts` timesteps (ts chosen from a probe so that
    the run takes about budget_s).  Returns (seg-steps/s, description, cores)."""
    from oracle import oracle as o
    from troute_b200 import hostgraph
    o.build()
    n = wl["n"]
    scols = np.asarray(o.column_mapper(wl["cols"]), dtype=np.int32)
    # the reach decomposition and the by-subnetwork job list are set-up (the reference builds them once per run), kept
    # across the steps of this process
    plan = wl.setdefault("_cpu_plan", {})
    if "reaches" not in plan:
        plan["reaches"] = hostgraph.segment_reaches_level_order(wl["down"], wl["up_ptr"], wl["up_rows"])
    reaches = plan["reaches"]
    jobs = None
    if threads > 1:
        if "jobs" not in plan:
            plan["jobs"] = hostgraph.subnetwork_jobs(wl["down"], wl["up_ptr"], wl["up_rows"], reaches["order"], target_size=10000)
        jobs = plan["jobs"]

    def run(ts):
        nq = max(1, int(np.ceil(ts / QTS)))
        t0 = time.perf_counter()
        o.route_network_flat(ts, DT, QTS, n, reaches["reach_ptr"], reaches["reach_rows"], reaches["reach_type"],
                             reaches["reach_up_ptr"], reaches["reach_up_rows"], wl["params"], scols, wl["q0"],
                             wl["qlat"][:, :nq], assume_short_ts=bool(args.short_ts), pow_mode=o.POW_LIBM,
                             jobs=jobs, nthreads=threads)
        return time.perf_counter() - t0

    probe_ts = 2
    tp = run(probe_ts)
    rate = n * probe_ts / tp
    ts = int(max(2, min(args.nsteps, budget_s * rate / n)))
    tt = run(ts)
    value = n * ts / tt
    kind = "by-subnetwork-jit orders/jobs over OpenMP threads" if jobs is not None else "serial reference loop order"
    sample = f"all {n} segments x first {ts} of {args.nsteps} timesteps ({tt:.1f} s), {kind}"
    return value, sample, threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    wl = build_workload(args)
    threads = os.cpu_count() or 1
    vals = []
    sample = ""
    budget = max(5.0, 90.0 / max(1, args.steps + args.warmup))
    for i in range(args.warmup + args.steps):
        v, sample, cores = cpu_arm(wl, args, threads, budget)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "routed segment-timesteps/sec", "value": value, "unit": "segment-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * wl["n"] * args.nsteps / value, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["name"], "assume_short_ts": bool(args.short_ts), "qts_subdivisions": QTS,
                   "note": "ms_per_step extrapolates the sampled rate to one full routing call"},
        "cpu_baseline": {"value": value, "unit": "segment-timesteps/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "segment-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from troute_b200 import _lib
    _lib.lib()   # fail loudly if the CUDA extension is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the routing path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: some boxes export NCCL_DEBUG=VERSION, which prints a banner on stdout
        os.environ["NCCL_DEBUG"] = os.environ.get("TRT_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = build_workload(args)
    T = args.nsteps
    if world > 1:
        from troute_b200 import multigpu
        runner = multigpu.ShardedRouter(wl, world, rank, local_rank, T, QTS, bool(args.short_ts), mode=args.mode,
                                        deep_lanes=deep_lanes_for(args, world))
    else:
        from troute_b200 import multigpu
        runner = multigpu.SingleRouter(wl, local_rank, T, QTS, bool(args.short_ts), mode=args.mode)
        runner.net.set_option("deep_lanes", deep_lanes_for(args, world))
        if args.levelpools:
            runner.net.set_levelpools(wl["lp_rows"], wl["wbody"], routing_period=DT)

    for kv in args.opt:
        k, v = kv.split("=")
        (runner.set_option if hasattr(runner, "set_option") else runner.net.set_option)(k, int(v))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    clocks = Clocks(local_rank)

    # ---- device-resident throughput ("value") ----
    runner.upload()
    engine_opts = {"deep_lanes": deep_lanes_for(args, world)}
    engine_opts.update({kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt})
    if world == 1 and not args.no_trip_order and args.mode in (2, 4):
        runner.reorder_by_trip_history(engine_opts, args.trip_buckets or None)   # set-up, not timed: once per network in production
    if world > 1 and args.sharded_trip_order and not args.no_trip_order and args.mode in (2, 4):
        runner.reorder_by_trip_history(args.trip_buckets or None)
    for _ in range(args.warmup):
        runner.run_resident()
    barrier()
    if rank == 0:
        clocks.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms, launches = [], 0
    ev0.record(runner.stream)
    for _ in range(args.steps):
        runner.run_resident()
        # stats of this call are read after the final synchronize (events stay valid per handle until the next run);
        # the per-step kernel time is collected by the runner itself
    ev1.record(runner.stream)
    barrier()
    total_ms = max_over_ranks(ev0.elapsed_time(ev1))
    clk = clocks.stop() if rank == 0 else None
    stats = runner.collect_stats()
    kern_ms = stats["kernel_ms_per_call"]            # mean duration of the wavefront kernel on this rank
    lane_steps_rank = stats["lane_steps"]
    launches = stats["launches_per_call"] * args.steps
    total_units = sum_over_ranks(float(lane_steps_rank))     # whole job, per routing call
    value = total_units * args.steps / (total_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    if not args.no_e2e:
        runner.alloc_host()
        for _ in range(min(args.warmup, 2)):
            runner.run_e2e()
        barrier()
        t_e2e = []
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(runner.stream)
        w0 = time.perf_counter()
        for _ in range(args.steps):
            runner.run_e2e()
        e1.record(runner.stream)
        barrier()
        wall_ms = (time.perf_counter() - w0) * 1e3
        # trt_route is synchronous (it returns when the result is in host memory), so wall clock and the event pair
        # bracket the same region; report the larger, max over ranks
        e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall_ms))
        e2e = {"value": total_units * args.steps / (e2e_ms * 1e-3), "unit": "segment-timesteps/s",
               "h2d_bytes_per_step": int(sum_over_ranks(float(runner.h2d_bytes))),
               "d2h_bytes_per_step": int(sum_over_ranks(float(runner.d2h_bytes))),
               "ms_per_step": e2e_ms / args.steps}
        launches += stats["launches_per_call_e2e"] * args.steps

    # ---- roofline of the dominant kernel ----
    # mode 4: the dataflow kernel over the wide shallow levels (~99 % of the segment-timesteps, ~80 % of the time); its
    # duration is the CUDA-event interval the engine records on its stream around that launch.  Other modes: the one
    # routing kernel.
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak = float(json.load(open(peaks_path))["hbm_gbs"]); peak_src = "MEASURED_PEAKS.json hbm_gbs (burst copy), of measured"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md), of fallback"
    if args.mode == 4 and stats["wide_lane_steps"] > 0:
        dom_name, dom_ms, dom_units = "trt::dataflow_kernel", stats["wide_ms"], stats["wide_lane_steps"]
    else:
        dom_name, dom_ms, dom_units = stats["kernel_name"], kern_ms, lane_steps_rank
    achieved = BYTES_PER_SEGSTEP * dom_units / (dom_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1 and args.workload == "conus" and not args.segments and args.style == "nhd":
        try:
            traffic = json.load(open(tpath)).get(dom_name)
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "kernel": dom_name, "kernel_ms": dom_ms,
                "algorithmic_bytes_per_launch": BYTES_PER_SEGSTEP * dom_units, "peak_source": peak_src,
                "all_routing_kernels_ms": kern_ms, "marching_kernel_ms": stats["march_ms"],
                "first_marching_level": stats["first_marching_level"],
                "note": "rank 0; 68 B per segment-timestep (SURVEY.md 8d).  The solve is instruction-issue bound "
                        "(~2,050 thread-instructions per segment-timestep at ~53 % lane efficiency, ncu: 69-74 % of issue "
                        "slots busy), not HBM bound; see DESIGN.md section 5"}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, sample, cores = cpu_arm(wl, args, 1, args.cpu_sample_seconds)
        cpu = {"value": v, "unit": "segment-timesteps/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "routed segment-timesteps/sec", "value": value, "unit": "segment-timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["name"], "assume_short_ts": bool(args.short_ts), "qts_subdivisions": QTS,
                       "levels": stats["levels"], "stages_per_call": stats["stages"],
                       "schedule": {0: "launch per stage", 1: "persistent cooperative wavefront (grid.sync per stage)",
                                    2: "dataflow wavefront (ordered unit queue, lanes poll their inputs)",
                                    3: "marching lanes (one lane per segment, all timesteps)",
                                    4: "dataflow wavefront over the wide shallow levels, marching lanes over the deep "
                                       "levels",
                                    5: "time-blocked marching lanes over the wide shallow levels (stage = level + "
                                       "block), marching lanes over the deep levels, one persistent kernel"}[args.mode],
                       "l2": "inputs larger than L2 (38 GB working set), no flush", "sharding": stats["sharding"],
                       "within_level_order": getattr(runner, "order_source", "secant trip counts of one calibration call")
                       if getattr(runner, "reordered", False) else "caller row order"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clk,
        }
        print(json.dumps(line), flush=True)
    runner.close()
    if world > 1:
        dist.destroy_process_group()


# ---------------------------------------------------------------------------------------------------
# diffusive-wave mainstem domains (BASELINE configs[3]): one CTA per domain, whole time loop on the device
# ---------------------------------------------------------------------------------------------------
def run_diffusive(args, rank, world):
    """`--workload diffusive`: `--domains` synthetic tailwater domains of `--mainstem` reaches (plus tributaries and a
    second arm), `--nsteps` x 300 s.  A step = one compute_diffusive_batch call (host dicts in, host arrays out: that is the
    only API, so `value` uses the device time of the time-loop kernel and `e2e` the wall time of the call).  Unit =
    mainstem segment-timesteps (segments x output rows) per second, the unit of the Muskingum-Cunge line."""
    if rank != 0:
        return
    from troute_b200 import synth_diffusive as sd
    doms = [sd.diffusive_domain(n_mainstem=args.mainstem, nodes=(5, 12), n_branch=max(0, args.mainstem // 6), nsteps=args.nsteps,
                                seed=1000 + k) for k in range(args.domains if args.impl == "ours" else min(args.domains, 8))]
    segs = sum(int(sum(d["frnw_g"][j, 0] - 1 for j in d["mainstem"])) for d in doms)
    units = float(segs * args.nsteps)
    name = (f"{len(doms)} synthetic diffusive-wave domains x {args.mainstem} mainstem reaches "
            f"({segs} mainstem segments), {args.nsteps} x 300 s, normal-depth tailwater")
    if args.impl == "reference":
        from oracle import diffusive as od
        od.build()
        vals = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            for d in doms:
                od.compute_diffusive(d, od.POW_LIBM)
            if i >= args.warmup:
                vals.append(units / (time.perf_counter() - t0))
        v = float(np.mean(vals))
        print(json.dumps({"impl": "reference", "metric": "routed segment-timesteps/sec", "value": v,
                          "unit": "segment-timesteps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": 1e3 * units / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64", "data": "synthetic", "config": {"workload": name},
                          "cpu_baseline": {"value": v, "unit": "segment-timesteps/s", "cores": 1, "kind": "port",
                                           "sample": f"{len(doms)} domains, serial loop over domains as compute.py:1764"},
                          "e2e": {"value": v, "unit": "segment-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "gpu_launches": 0}), flush=True)
        return
    import torch
    from troute_b200 import _lib
    from troute_b200.routing.fast_reach import diffusive
    _lib.lib()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the diffusive path has no CPU fallback")
    clocks = Clocks(0)
    for _ in range(args.warmup):
        diffusive.compute_diffusive_batch(doms)
    clocks.start()
    loop_ms, table_ms, wall = [], [], []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        out = diffusive.compute_diffusive_batch(doms)
        wall.append((time.perf_counter() - t0) * 1e3)
        a, b, _n = diffusive.last_run()
        table_ms.append(a); loop_ms.append(b)
    clk = clocks.stop()
    h2d = sum(sum(np.asarray(v).nbytes for v in d.values() if isinstance(v, np.ndarray)) for d in doms)
    d2h = sum(sum(o.nbytes for o in res) for res in out)
    dev_ms = float(np.mean(loop_ms) + np.mean(table_ms))
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import diffusive as od
        od.build()
        t0 = time.perf_counter(); k = 0
        while k < len(doms) and time.perf_counter() - t0 < args.cpu_sample_seconds:
            od.compute_diffusive(doms[k], od.POW_LIBM); k += 1
        cs = sum(int(sum(d["frnw_g"][j, 0] - 1 for j in d["mainstem"])) for d in doms[:k])
        cpu = {"value": cs * args.nsteps / (time.perf_counter() - t0), "unit": "segment-timesteps/s", "cores": 1,
               "kind": "port", "sample": f"first {k} of {len(doms)} domains, serial (compute.py:1764 loops over domains)"}
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak = float(json.load(open(peaks_path))["hbm_gbs"]) if os.path.exists(peaks_path) else 6650.0
    # algorithmic bytes of the time-loop kernel per mainstem node and OUTPUT step: 24 B of results + one pass over the 33
    # state arrays (264 B); the adaptive step makes the real count a small multiple.  The kernel is a dependency chain
    # of Newton solves per domain (latency-bound), so this fraction is reported because the contract asks for it.
    nodes = sum(int(sum(d["frnw_g"][j, 0] for j in d["mainstem"])) for d in doms)
    abytes = 288.0 * nodes * args.nsteps
    ach = abytes / (float(np.mean(loop_ms)) * 1e-3) / 1e9
    print(json.dumps({
        "metric": "routed segment-timesteps/sec", "value": units / (dev_ms * 1e-3), "unit": "segment-timesteps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": name, "l2": "tables (32 KB per node) are rebuilt by every call; inputs re-uploaded",
                   "schedule": "one CTA per domain, whole time loop in one launch"},
        "e2e": {"value": units / (float(np.mean(wall)) * 1e-3), "unit": "segment-timesteps/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": float(np.mean(wall))},
        "gpu_launches": 4 * args.steps,
        "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None,
                     "kernel": "time_loop_kernel", "kernel_ms": float(np.mean(loop_ms)), "table_kernels_ms": float(np.mean(table_ms)),
                     "algorithmic_bytes_per_launch": abytes,
                     "note": "latency-bound: per domain one dependency chain of Newton solves per time step; see DESIGN.md"},
        "cpu_baseline": cpu, "clocks": clk}), flush=True)


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 and args.gpus > 1 and args.impl == "ours":
        # convenience: re-launch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29513", os.path.abspath(__file__)] + sys.argv[1:]
        os.execv(sys.executable, cmd)
    if args.workload == "diffusive":
        run_diffusive(args, rank, world)
    elif args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
