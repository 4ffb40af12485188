#!/usr/bin/env python
"""bench.py -- routed segment-timesteps/sec of the B200 routing path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload conus|conus-lp7d|tree|diffusive]

Workload (config.workload): BASELINE.json configs[2] -- the synthetic CONUS-scale forest the metric is quoted on:
2,729,077 segments in 14,713 basins (largest ~50 %), NHD-like confluences (SURVEY.md 8d in-degree mix), MC-only,
288 x 300 s steps, dependent upstream flows (assume_short_ts = False, the reference default).  One bench "step" = ONE
routing call = all 288 timesteps of all segments (786 M segment-timesteps).  Synthetic parameters/forcing, cold start; see
troute_b200/synth.py.  --workload tree is BASELINE configs[1] (binary tree, 1,048,576 segments); --workload conus-lp7d is
configs[4]: the same forest with 5,000 level-pool reservoirs, a 7-day hindcast routed as 7 windows of 288 steps whose state
is handed from window to window ON THE DEVICE (trt_continue), one "step" = all 7 windows (5.5 G segment-timesteps).

  set-up    (not timed, once per network as in production where the handle is cached across calls): flatten + upload of the
            network, one calibration call ON A DIFFERENT STORM that records the secant trip count of every segment, rebuild
            with the segments of every wavefront level ordered by it (config.within_level_order; --no-trip-order skips it).
  value     segment-timesteps/s with forcing and state already resident in HBM: K x (flow-state reset + routing kernels +
            result pass) timed with CUDA events on the launching stream.  Working set (flow state 6.3 GB + result 9.4 GB +
            tile records 0.17 GB per stage) is far larger than the 126 MB L2, so no explicit flush between iterations.
            `value_incl_h2d` is the same region with qlat / q0 coming from pinned host memory (SURVEY.md 8d's definition);
            `value_uncalibrated` the same as `value` before the trip-count ordering (caller row order).
  e2e       the same metric through the C-ABI call a T-Route maintainer would bind (trt_route / trt_continue +
            trt_run_download): pinned HOST qlat / q0 in, pinned HOST [n, 3*nsteps] result out, copies inside the timed region
            (trt_route overlaps the result copies with the kernels by time chunks).
  roofline  dominant kernel (the dataflow kernel over the wide levels): algorithmic bytes (68 B per segment-timestep,
            SURVEY.md 8d / DESIGN.md) x segment-timesteps of that launch / its CUDA-event duration (events recorded by the
            engine on its stream around that launch), against MEASURED_PEAKS.json hbm_gbs; traffic = ncu dram bytes of
            that kernel (profiles/traffic.json, stamped with the commit it was captured on).
  verify    proof that THIS run's numbers are right: (i) `hash` = 64-bit checksum of the result bits in global row order
            (sum over rows of hash(row, bits): the per-rank checksums of a sharded run add up, so N = 1/2/4/8 print the same
            value); (ii) whole independent basins of the timed network (>= 50,000 segments, every timestep) routed again
            by the CPU oracle: `mismatches` = values whose bits differ from the oracle's bit-specified-pow build (must be 0),
            `frac_within_1e-5_of_libm` = agreement with the oracle's platform-libm build (what a gfortran build of the
            reference computes here).  The oracle is the checker, never the thing timed.
  cpu_baseline  the oracle's platform-libm build (C restatement of the Fortran; neither this container nor the GPU box has a
            Fortran compiler) on a bounded sample: the SAME generator at reduced scale, ALL timesteps, one core, rank 0, N = 1.

--impl reference times that CPU restatement with every host thread, decomposed the way the reference's parallel modes do
(orders of sub-networks, jobs of an order in parallel).  Each step routes ALL timesteps of a scaled-down network of the
same generator (sized so that K + W steps take about three minutes); ms_per_step is the measured time of such a step.

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
STORM = ((8.0, 2.0, 8.0),)                              # the timed forcing: 1 + 2 exp(-(hour - 8)^2 / 8)  
STORM_7D = ((8.0, 2.0, 8.0), (40.0, 1.5, 10.0), (70.0, 3.0, 12.0), (110.0, 1.0, 8.0), (150.0, 2.5, 16.0))
STORM_CALIBRATION = ((14.0, 3.0, 16.0),)                # the storm the within-level order is calibrated on (another one)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="conus", choices=["conus", "conus-lp7d", "tree", "diffusive"],
                    help="conus / tree: Muskingum-Cunge routing (BASELINE configs[2] / configs[1]); conus-lp7d: configs[4], the "
                         "conus forest + 5,000 level pools, 7 windows of 288 steps with the state carried on the device; "
                         "diffusive: a batch of diffusive-wave mainstem domains (configs[3] kernel, see run_diffusive)")
    ap.add_argument("--domains", type=int, default=296, help="diffusive: independent tailwater domains per call")
    ap.add_argument("--mainstem", type=int, default=24, help="diffusive: mainstem reaches per domain")
    ap.add_argument("--style", default="nhd", choices=["nhd", "hack"],
                    help="basin generator of the conus workload: nhd = NHD-like confluences (SURVEY.md 8d in-degree mix), "
                         "hack = main stems with dozens of tributaries per node (gather stress case)")
    ap.add_argument("--segments", type=int, default=0, help="override the segment count (testing only)")
    ap.add_argument("--nsteps", type=int, default=288, help="routing timesteps per call (per window)")
    ap.add_argument("--windows", type=int, default=0, help="routing windows per step (0 = 1, or 7 for conus-lp7d)")
    ap.add_argument("--short-ts", type=int, default=0)
    ap.add_argument("--mode", type=int, default=4)
    ap.add_argument("--opt", action="append", default=[], help="engine option key=value (repeatable)")
    ap.add_argument("--levelpools", type=int, default=-1,
                    help="replace this many in-line segments by level-pool reservoirs (default 0; 5000 for conus-lp7d)")
    ap.add_argument("--deep-lanes", type=int, default=0,
                    help="segments per GPU that march (deepest levels); 0 = 8192 on 1 GPU, 4096 on 2, 2048 on 4+ (one lane per "
                         "warp: the main stem is the critical path once the wide levels are spread over many GPUs)")
    ap.add_argument("--no-trip-order", action="store_true",
                    help="skip the calibration call that orders the segments of a level by their secant trip counts")
    ap.add_argument("--calibrate-on", default="other-storm", choices=["other-storm", "same-storm"],
                    help="forcing of the calibration call: a different storm than the timed one (default), or the timed one")
    ap.add_argument("--trip-buckets", type=int, default=0,
                    help="time slices of the calibration call's trip counts (0 = network.TRIP_BUCKETS, 1 = totals only)")
    ap.add_argument("--sharded-trip-order", action="store_true",
                    help="N > 1: calibrate and re-order every shard too (default: caller row order -- measured, an order "
                         "calibrated on another storm is worth nothing, profiles/r02_v4_final, r02_multi_gpu)")
    ap.add_argument("--no-verify", action="store_true", help="skip the verify object (result hash + oracle comparison)")
    ap.add_argument("--verify-segments", type=int, default=50000, help="oracle comparison: at least this many segments")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-seconds", type=float, default=20.0)
    args = ap.parse_args()
    if args.workload == "conus-lp7d":
        args.windows = args.windows or 7
        args.levelpools = 5000 if args.levelpools < 0 else args.levelpools
    args.windows = args.windows or 1
    args.levelpools = max(0, args.levelpools)
    return args


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


def build_workload(args, scale=1.0):
    """The network and forcing of the run.  `scale` < 1: the same generator at reduced size (CPU arms: a bounded sample that
    still routes ALL timesteps -- sampling in time instead would shorten the jobs of the reference's parallel decomposition
    and understate it)."""
    from troute_b200 import synth
    windows = args.windows
    total_steps = args.nsteps * windows
    if args.workload in ("conus", "conus-lp7d"):
        n_full = args.segments or 2_729_077
        n = max(2000, int(round(n_full * scale)))
        basins = max(1, int(round(14_713 * n / 2_729_077)))
        down = _cached(f"conus_{args.style}_{n}_{basins}_16",
                       lambda: synth.conus_like(n_total=n, n_basins=basins, seed=16, style=args.style))
        name = (f"synthetic CONUS-scale forest ({args.style}-style basins), {n} segments / {basins} basins, MC-only, "
                f"{args.nsteps} x 300 s")
    else:
        n = max(1023, int(round((args.segments or 1_048_576) * scale)))
        down = synth.binary_tree(n)
        name = f"synthetic balanced binary tree, {n} segments, MC-only, {args.nsteps} x 300 s"
    params = synth.channel_params(down, dt=DT, seed=16)
    storms = STORM_7D if (windows > 1 or total_steps > 288) else STORM      # a week of weather for a week of steps
    qlat = synth.lateral_inflow(n, total_steps, QTS, seed=16, storms=storms)
    q0 = np.zeros((n, 3), dtype=np.float32)
    up_ptr, up_rows = synth.upstream_csr(down)
    kind = np.zeros(n, dtype=np.uint8)
    lp_rows, wbody = np.zeros(0, np.int64), np.zeros((0, 11))
    n_lp = int(round(args.levelpools * (n / (args.segments or 2_729_077)))) if scale != 1.0 else args.levelpools
    if n_lp:
        rng = np.random.default_rng(23)
        cand = np.nonzero(np.diff(up_ptr) > 0)[0]
        lp_rows = np.sort(rng.choice(cand, size=min(n_lp, cand.size), replace=False)).astype(np.int64)
        kind[lp_rows] = 1
        wbody = synth.levelpool_params(lp_rows.size, seed=16)
        name = name.replace("MC-only", f"MC + {lp_rows.size} level-pool reservoirs")
    if windows > 1:
        name += f" per window, {windows} windows ({total_steps} steps) with device-resident state between windows"
    return dict(name=name, n=n, down=down, params=params, cols=synth.PARAM_COLS, qlat=qlat, q0=q0, up_ptr=up_ptr,
                up_rows=up_rows, kind=kind, lp_rows=lp_rows, wbody=wbody, total_steps=total_steps, windows=windows)


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
def _oracle_plan(wl, threads):
    """Reach decomposition (every segment a one-segment reach, level order; level pools typed as such) and, for a parallel
    run, the by-subnetwork job list -- set-up the reference builds once per run, kept across the steps of this process."""
    from troute_b200 import hostgraph
    plan = wl.setdefault("_cpu_plan", {})
    if "reaches" not in plan:
        r = hostgraph.segment_reaches_level_order(wl["down"], wl["up_ptr"], wl["up_rows"])
        order = r["order"]
        r["reach_type"] = wl["kind"][order].astype(np.int32)
        wb = np.full(wl["n"], -1, dtype=np.int32)
        wb[wl["lp_rows"]] = np.arange(len(wl["lp_rows"]), dtype=np.int32)
        r["reach_wbody"] = wb[order]
        plan["reaches"] = r
    if threads > 1 and "jobs" not in plan:
        plan["jobs"] = hostgraph.subnetwork_jobs(wl["down"], wl["up_ptr"], wl["up_rows"], plan["reaches"]["order"], target_size=10000)
    return plan["reaches"], (plan["jobs"] if threads > 1 else None)


def oracle_route(wl, nsteps, short_ts, pow_mode, threads=1):
    """All `nsteps` steps of workload `wl` through the oracle; returns flowveldepth [n, nsteps + 1, 3]."""
    from oracle import oracle as o
    o.build()
    scols = np.asarray(o.column_mapper(wl["cols"]), dtype=np.int32)
    reaches, jobs = _oracle_plan(wl, threads)
    nq = max(1, int(np.ceil(nsteps / QTS)))
    fvd, _, _ = o.route_network_flat(nsteps, DT, QTS, wl["n"], reaches["reach_ptr"], reaches["reach_rows"], reaches["reach_type"],
                                     reaches["reach_up_ptr"], reaches["reach_up_rows"], wl["params"], scols, wl["q0"],
                                     wl["qlat"][:, :nq], assume_short_ts=bool(short_ts), reach_wbody=reaches["reach_wbody"],
                                     wbody_cols=wl["wbody"], pow_mode=pow_mode, jobs=jobs, nthreads=threads)
    return fvd


def cpu_sample_workload(args, threads, budget_s):
    """The bounded sample of the CPU arms: the SAME generator at a scale chosen so that routing ALL timesteps takes about
    budget_s (rate probed on a small network first)."""
    from oracle import oracle as o
    probe = build_workload(args, scale=min(1.0, 60_000 / (args.segments or 2_729_077)))
    _oracle_plan(probe, threads)
    t0 = time.perf_counter()
    oracle_route(probe, min(48, probe["total_steps"]), args.short_ts, o.POW_LIBM, threads)
    rate = probe["n"] * min(48, probe["total_steps"]) / (time.perf_counter() - t0)
    full_units = (args.segments or (2_729_077 if args.workload != "tree" else 1_048_576)) * probe["total_steps"]
    scale = min(1.0, budget_s * rate / full_units)
    wl = build_workload(args, scale=scale)
    _oracle_plan(wl, threads)                          # set-up, not timed
    return wl


def cpu_step(wl, args, threads):
    """One timed CPU step: every timestep of the sample network.  Returns (seg-steps/s, seconds)."""
    from oracle import oracle as o
    t0 = time.perf_counter()
    oracle_route(wl, wl["total_steps"], args.short_ts, o.POW_LIBM, threads)
    dt = time.perf_counter() - t0
    return wl["n"] * wl["total_steps"] / dt, dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nrun = max(1, args.steps + args.warmup)
    wl = cpu_sample_workload(args, threads, budget_s=max(2.0, 170.0 / nrun))
    vals, secs = [], []
    for i in range(args.warmup + args.steps):
        v, dt = cpu_step(wl, args, threads)
        if i >= args.warmup:
            vals.append(v); secs.append(dt)
    value = float(np.mean(vals))
    full = build_workload_name(args)
    sample = (f"same generator at reduced scale: {wl['n']} segments x all {wl['total_steps']} timesteps per step "
              f"({np.mean(secs):.1f} s), by-subnetwork-jit orders/jobs over OpenMP threads")
    line = {
        "impl": "reference", "metric": "routed segment-timesteps/sec", "value": value, "unit": "segment-timesteps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * float(np.mean(secs)), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": full, "assume_short_ts": bool(args.short_ts), "qts_subdivisions": QTS,
                   "note": "each step routes every timestep of a scaled-down network of the same generator; ms_per_step is "
                           "the measured time of such a step, value its rate"},
        "cpu_baseline": {"value": value, "unit": "segment-timesteps/s", "cores": threads, "kind": "port",
                         "sample": sample},
        "e2e": {"value": value, "unit": "segment-timesteps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def build_workload_name(args):
    """config.workload of the full-size run, without building it (the reference arm reports the same string)"""
    n = args.segments or (1_048_576 if args.workload == "tree" else 2_729_077)
    if args.workload == "tree":
        name = f"synthetic balanced binary tree, {n} segments, MC-only, {args.nsteps} x 300 s"
    else:
        basins = max(1, int(round(14_713 * n / 2_729_077)))
        name = (f"synthetic CONUS-scale forest ({args.style}-style basins), {n} segments / {basins} basins, MC-only, "
                f"{args.nsteps} x 300 s")
    if args.levelpools:
        name = name.replace("MC-only", f"MC + {args.levelpools} level-pool reservoirs")
    if args.windows > 1:
        name += (f" per window, {args.windows} windows ({args.nsteps * args.windows} steps) with device-resident state "
                 f"between windows")
    return name


# ---------------------------------------------------------------------------------------------------
# verification of the timed run (the oracle is the checker, never the thing timed)
# ---------------------------------------------------------------------------------------------------
def _mix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
        return x ^ (x >> np.uint64(31))


def verification_basins(wl, min_segments):
    """>= 3 whole independent basins (none of them the giant one) holding at least `min_segments` segments together:
    (global rows ascending, sub-workload routed on its own by the oracle)."""
    from troute_b200 import synth
    root = synth.basin_of(wl["down"])
    ids, sizes = np.unique(root, return_counts=True)
    giant = int(np.argmax(sizes))                         # the basin that holds half the network is not a sample
    order = np.argsort(np.abs(sizes - 1.2 * min_segments / 3.0), kind="stable")   # basins of about a third of the target
    chosen, total = [], 0
    for i in order:
        if i == giant:
            continue
        chosen.append(ids[i]); total += int(sizes[i])
        if len(chosen) >= 3 and total >= min_segments:
            break
    rows = np.nonzero(np.isin(root, np.asarray(chosen)))[0].astype(np.int64)
    remap = -np.ones(wl["n"], dtype=np.int64); remap[rows] = np.arange(rows.size)
    down = np.where(wl["down"][rows] >= 0, remap[np.maximum(wl["down"][rows], 0)], -1)
    up_ptr, up_rows = synth.upstream_csr(down)
    lp_mask = np.isin(wl["lp_rows"], rows)
    sub = dict(n=int(rows.size), down=down, params=wl["params"][rows], cols=wl["cols"], qlat=wl["qlat"][rows], q0=wl["q0"][rows],
               up_ptr=up_ptr, up_rows=up_rows, kind=wl["kind"][rows], lp_rows=remap[wl["lp_rows"][lp_mask]],
               wbody=np.asarray(wl["wbody"])[lp_mask], total_steps=wl["total_steps"], windows=wl["windows"])
    return rows, sub, len(chosen)


class Verifier:
    """Collects, window by window, (i) the checksum of this rank's result rows and (ii) the result rows of the verification
    basins; finish() compares the latter with the oracle on rank 0."""

    def __init__(self, args, wl, runner, rank, world, dist):
        self.args, self.wl, self.runner, self.rank, self.world, self.dist = args, wl, runner, rank, world, dist
        self.rows, self.sub, self.n_basins = verification_basins(wl, args.verify_segments)
        self.win_hashes = []                             # per window: checksum of this rank's own rows (they ADD over ranks)
        self.parts = []                                  # per window: (global rows owned, their [*, 3T] results)

    def window_done(self, w):
        self.win_hashes.append(int(self.runner.window_hash()))
        self.parts.append(self.runner.download_global_rows(self.rows))

    @staticmethod
    def combine(per_rank):
        """per_rank[r][w] = checksum of rank r's rows of window w -> one 64-bit value: the rank sums of a window add up to
        the checksum of the whole window (whatever the sharding); windows are then mixed with their index and summed."""
        total = 0
        for w in range(len(per_rank[0])):
            hw = sum(int(h[w]) for h in per_rank) % (1 << 64)
            total = (total + int(_mix64(np.asarray([hw ^ int(_mix64(np.asarray([w], dtype=np.uint64))[0])], dtype=np.uint64))[0])) % (1 << 64)
        return total

    def finish(self):
        from oracle import oracle as o
        every = [self.win_hashes]
        gathered = [self.parts]
        if self.world > 1:
            every = [None] * self.world
            self.dist.all_gather_object(every, self.win_hashes)
            gathered = [None] * self.world if self.rank == 0 else None
            self.dist.gather_object(self.parts, gathered, dst=0)
        if self.rank != 0:
            return None
        total = self.combine(every)
        T, W = self.args.nsteps, self.wl["windows"]
        got = np.full((self.rows.size, 3 * T * W), np.nan, dtype=np.float32)
        for parts in gathered:
            for w, (grows, vals) in enumerate(parts):
                got[np.searchsorted(self.rows, grows), w * 3 * T:(w + 1) * 3 * T] = vals
        t0 = time.perf_counter()
        threads = os.cpu_count() or 1
        det = oracle_route(self.sub, T * W, self.args.short_ts, o.POW_DET, threads)[:, 1:, :].reshape(self.rows.size, -1)
        libm = oracle_route(self.sub, T * W, self.args.short_ts, o.POW_LIBM, threads)[:, 1:, :].reshape(self.rows.size, -1)
        mism = int(((got.view(np.int32) != det.view(np.int32)) & ~(np.isnan(got) & np.isnan(det))).sum())
        q_got, q_ref = got[:, 0::3].astype(np.float64), libm[:, 0::3].astype(np.float64)
        rel = np.abs(q_got - q_ref) / np.maximum(np.abs(q_ref), 1e-30)
        close = (rel <= 1e-5) | (np.abs(q_got - q_ref) <= 1e-9)
        return {"hash": f"{total:016x}",
                "hash_of": "result bits of every window in global row order (sum over rows of hash(row, bits); the per-rank "
                           "sums of a sharded run add up)",
                "oracle_basins": self.n_basins, "oracle_segments": int(self.rows.size), "oracle_steps": T * W,
                "values_compared": int(got.size), "mismatches": mism,
                "mismatches_vs": "oracle, bit-specified pow (trt_powf_det), bit for bit, q / v / d of every step",
                "frac_within_1e-5_of_libm": float(close.mean()),
                "libm_note": "streamflow vs the oracle's platform-libm build (what a gfortran build of the reference "
                             "computes on this host); the remainder are secant-termination flips of a 1-ulp powf difference",
                "oracle_seconds": round(time.perf_counter() - t0, 1)}


# ---------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------
def bind_to_gpu_numa_node(local_rank):
    """Run this rank on the CPUs next to its GPU (NVML's ideal CPU affinity): the pinned result buffer is then first-touched on
    that NUMA node and the D2H DMA of 8 ranks does not cross the socket interconnect.  Best effort; returns a note."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, ((os.cpu_count() or 64) + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        allowed = os.sched_getaffinity(0)
        use = sorted(cpus & allowed)
        if use and len(use) < len(allowed):
            os.sched_setaffinity(0, use)
            return f"rank bound to {len(use)} of {len(allowed)} CPUs next to its GPU"
        return "single NUMA domain (no binding needed)"
    except Exception as e:                                # noqa: BLE001 -- placement is an optimisation
        return f"not bound ({type(e).__name__})"


def run_ours(args, rank, world, local_rank):
    numa_note = bind_to_gpu_numa_node(local_rank)
    import torch
    import torch.distributed as dist
    from troute_b200 import _lib, synth
    _lib.lib()   # fail loudly if the CUDA extension is missing
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the routing path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # keep stdout to the one JSON line: some boxes export NCCL_DEBUG=VERSION, which prints a banner on stdout
        os.environ["NCCL_DEBUG"] = os.environ.get("TRT_NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    wl = build_workload(args)
    T, W = args.nsteps, args.windows
    from troute_b200 import multigpu
    if world > 1:
        runner = multigpu.ShardedRouter(wl, world, rank, local_rank, T, QTS, bool(args.short_ts), mode=args.mode,
                                        deep_lanes=deep_lanes_for(args, world), windows=W)
    else:
        runner = multigpu.SingleRouter(wl, local_rank, T, QTS, bool(args.short_ts), mode=args.mode, windows=W)
        runner.set_option("deep_lanes", deep_lanes_for(args, world))
    for kv in args.opt:
        k, v = kv.split("=")
        runner.set_option(k, int(v))

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

    def timed(fn, steps, warmup):
        """K steps of fn bracketed by a barrier + synchronize on both sides, CUDA events on the launching stream, max over
        ranks; returns ms for the K steps."""
        for _ in range(warmup):
            fn()
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        ev0.record(runner.stream)
        for _ in range(steps):
            fn()
        ev1.record(runner.stream)
        barrier()
        wall_ms = (time.perf_counter() - w0) * 1e3
        return max_over_ranks(ev0.elapsed_time(ev1)), max_over_ranks(wall_ms)

    clocks = Clocks(local_rank)
    runner.upload()                                          # first touch: allocations, peer wiring

    # ---- the same region before the trip-count ordering (caller row order) ----
    reorder = (not args.no_trip_order and args.mode in (2, 4) and (world == 1 or args.sharded_trip_order))
    uncal = None
    if reorder:
        ms, _ = timed(runner.run_resident, max(1, min(args.steps, 3)), 2)
        lane_steps0 = sum_over_ranks(float(runner.collect_stats()["lane_steps"]))
        uncal = lane_steps0 * max(1, min(args.steps, 3)) / (ms * 1e-3)
        # set-up, not timed: once per network in production.  The calibration call routes a DIFFERENT storm than the timed one.
        cal = None
        if args.calibrate_on == "other-storm":
            cal = synth.lateral_inflow(wl["n"], T, QTS, seed=17, storms=STORM_CALIBRATION)
        runner.reorder_by_trip_history(args.trip_buckets or None, cal)

    # ---- device-resident throughput ("value") ----
    for _ in range(args.warmup):
        runner.run_resident()
    barrier()
    if rank == 0:
        clocks.start()
    total_ms, _ = timed(runner.run_resident, args.steps, 0)
    clk = clocks.stop() if rank == 0 else None
    stats = runner.collect_stats()
    kern_ms = stats["kernel_ms_per_call"]            # duration of the routing kernels of the last window on this rank
    lane_steps_rank = stats["lane_steps"]
    launches = stats["launches_per_call"] * args.steps
    total_units = sum_over_ranks(float(lane_steps_rank))     # whole job, per bench step (all windows)
    value = total_units * args.steps / (total_ms * 1e-3)

    # ---- the same with the forcing coming from pinned host memory (SURVEY.md 8d counts the qlat / q0 upload) ----
    runner.alloc_host(result=not args.no_e2e)
    h2d_ms, _ = timed(runner.run_resident_incl_h2d, args.steps, 1)
    value_incl_h2d = total_units * args.steps / (h2d_ms * 1e-3)

    # ---- end to end through the C ABI with host buffers ----
    e2e = None
    verifier = None if args.no_verify else Verifier(args, wl, runner, rank, world, dist)
    if not args.no_e2e:
        ev_ms, wall_ms = timed(runner.run_e2e, args.steps, min(args.warmup, 2))
        # the call is synchronous (it returns when the result is in host memory), so wall clock and the event pair bracket
        # the same region; report the larger, max over ranks
        e2e_ms = max(ev_ms, wall_ms)
        e2e = {"value": total_units * args.steps / (e2e_ms * 1e-3), "unit": "segment-timesteps/s",
               "h2d_bytes_per_step": int(sum_over_ranks(float(runner.h2d_bytes))),
               "d2h_bytes_per_step": int(sum_over_ranks(float(runner.d2h_bytes))),
               "ms_per_step": e2e_ms / args.steps}
        launches += stats["launches_per_call_e2e"] * args.steps

    # ---- verify: one more pass, untimed, stopped after every window to checksum / fetch the device-resident result ----
    verify = None
    if verifier is not None:
        runner.run_checked(verifier.window_done)
        barrier()
        verify = verifier.finish()

    # ---- roofline of the dominant kernel ----
    # mode 4: the dataflow kernel over the wide shallow levels (~99 % of the segment-timesteps, ~75 % of the time); its
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
        dom_name, dom_ms, dom_units = stats["kernel_name"], kern_ms, lane_steps_rank / W
    achieved = BYTES_PER_SEGSTEP * dom_units / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath) and world == 1 and args.workload == "conus" and not args.segments and args.style == "nhd":
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(dom_name)
            traffic_src = f"{tj.get('source')}; captured on commit {tj.get('commit')}"
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "kernel": dom_name, "kernel_ms": dom_ms,
                "algorithmic_bytes_per_launch": BYTES_PER_SEGSTEP * dom_units, "peak_source": peak_src,
                "all_routing_kernels_ms": kern_ms, "marching_kernel_ms": stats["march_ms"],
                "first_marching_level": stats["first_marching_level"],
                "note": "rank 0, one launch = one window; 68 B per segment-timestep (SURVEY.md 8d).  The solve is "
                        "instruction-issue bound (~2,000 thread-instructions per segment-timestep), not HBM bound: ncu "
                        "(profiles/) shows DRAM traffic at 2.0x the algorithmic bytes and 13 % of DRAM throughput; see "
                        "DESIGN.md section 5"}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        swl = cpu_sample_workload(args, 1, args.cpu_sample_seconds)
        v, dt = cpu_step(swl, args, 1)
        cpu = {"value": v, "unit": "segment-timesteps/s", "cores": 1, "kind": "port",
               "sample": f"same generator at reduced scale: {swl['n']} segments x all {swl['total_steps']} timesteps "
                         f"({dt:.1f} s), serial reference loop order"}

    if rank == 0:
        line = {
            "metric": "routed segment-timesteps/sec", "value": value, "unit": "segment-timesteps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "value_incl_h2d": value_incl_h2d, "value_uncalibrated": uncal,
            "config": {"workload": wl["name"], "assume_short_ts": bool(args.short_ts), "qts_subdivisions": QTS,
                       "levels": stats["levels"], "stages_per_call": stats["stages"],
                       "schedule": {0: "launch per stage", 1: "persistent cooperative wavefront (grid.sync per stage)",
                                    2: "dataflow wavefront (ordered unit queue, lanes poll their inputs)",
                                    3: "marching lanes (one lane per segment, all timesteps)",
                                    4: "dataflow wavefront over the wide shallow levels, marching lanes over the deep "
                                       "levels"}[args.mode],
                       "l2": "inputs larger than L2 (16 GB working set per window), no flush", "sharding": stats["sharding"],
                       "host_placement": numa_note,
                       "value_definition": "forcing resident in HBM; value_incl_h2d: the same region with qlat / q0 copied from "
                                           "pinned host memory inside it (SURVEY.md 8d); value_uncalibrated: caller row order",
                       "within_level_order": (getattr(runner, "order_source", "") + (
                           "; calibrated on a different storm than the timed one" if args.calibrate_on == "other-storm"
                           else "; calibrated on the timed storm")) if getattr(runner, "reordered", False) else "caller row order"},
            "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "verify": verify,
            "clocks": clk,
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
