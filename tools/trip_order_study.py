"""Offline study (CPU only, test infrastructure: uses the oracle): how full are the warps of the dataflow kernel for a
given within-level order of the segments?

The dataflow kernel (csrc/routing_kernels.cu) gives 32 consecutive positions of one wavefront level to a warp; the warp
runs until its slowest lane has finished its secant solve, so its cost is 32 x (max trips + 1) lane-trips (the first trip
evaluates two cross-sections, later trips one) while the useful work is the sum of (trips + 1) over the lanes with flow.
Any within-level order gives the same bits (tests/test_gpu_parity.py), so the order is a pure performance knob:
`trt_network_create_ordered(order_key)`.  This script computes the secant trip count of every (segment, timestep) with
the oracle (pinned-pow build: the counts the device sees) and evaluates candidate keys.

    python tools/trip_order_study.py [n_segments] [nsteps]

Prints one line per candidate: lane efficiency = useful / cost, the implied active lanes out of 32, and the share of
warp-steps whose lanes disagree about being above bankfull depth (such a warp executes both branches of the celerity).
"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)


def trip_matrix(o, case, fvd):
    """trips[s, t-1] of every segment-step, recomputed from the routed flows (reference loop semantics, mc_reach.pyx:496-505)"""
    n, T, qts = case["n"], case["nsteps"], case["qts"]
    cols = list(case["cols"])
    P = case["params"]
    col = {c: P[:, cols.index(c)] for c in cols}
    q = np.concatenate([case["q0"][:, 0:1], fvd[:, 0::3]], axis=1)          # q[s, t], t = 0..T
    d = np.concatenate([case["q0"][:, 2:3], fvd[:, 2::3]], axis=1)
    up_ptr, up_rows = case["up_ptr"], case["up_rows"]
    seg_of_edge = np.repeat(np.arange(n), np.diff(up_ptr))
    trips = np.zeros((n, T), dtype=np.int16)
    in15 = np.zeros((n, 15), dtype=np.float32)
    in15[:, 0] = col["dt"]
    for j, c in enumerate(("dx", "bw", "tw", "twcc", "n", "ncc", "cs", "s0")):
        in15[:, 5 + j] = col[c]
    for t in range(1, T + 1):
        # float32 sums in CSR order, as the device and the oracle do
        quc = np.zeros(n, dtype=np.float32)
        qup = np.zeros(n, dtype=np.float32)
        # in-degree is small: accumulate edge by edge rank to keep the summation order
        rank = np.arange(up_rows.size) - np.repeat(up_ptr[:-1], np.diff(up_ptr))
        for r in range(int(rank.max()) + 1 if rank.size else 0):
            m = rank == r
            quc[seg_of_edge[m]] += q[up_rows[m], t]
            qup[seg_of_edge[m]] += q[up_rows[m], t - 1]
        in15[:, 1] = qup
        in15[:, 2] = quc
        in15[:, 3] = q[:, t - 1]
        in15[:, 4] = case["qlat"][:, (t - 1) // qts]
        in15[:, 14] = d[:, t - 1]
        _, it = o.mc_segment_batch(in15, pow_mode=o.POW_DET)
        trips[:, t - 1] = it
    return trips


def efficiency(trips, level, key, min_width=64):
    """(useful, cost) in lane-trips over all levels at least min_width wide, segments of a level sorted by `key`"""
    order = np.lexsort((key, level))
    lv = level[order]
    tr = trips[order]
    starts = np.flatnonzero(np.r_[True, lv[1:] != lv[:-1]])
    ends = np.r_[starts[1:], lv.size]
    useful = cost = 0
    for a, b in zip(starts, ends):
        if b - a < min_width:
            continue
        x = tr[a:b]
        pad = (-x.shape[0]) % 32
        if pad:
            x = np.concatenate([x, np.zeros((pad, x.shape[1]), dtype=x.dtype)])
        g = x.reshape(-1, 32, x.shape[1]).astype(np.int32)
        flow = g > 0
        useful += int((g + 1)[flow].sum())
        mx = g.max(axis=1)
        cost += int(((mx + 1) * (mx > 0)).sum()) * 32
    return useful, cost


def overbank_steps(case, fvd):
    """[n, T] bool: the segment ended the step above bankfull depth in a compound channel (what McResult.over counts)"""
    cols = list(case["cols"])
    P = np.asarray(case["params"], dtype=np.float32)
    bw, tw, cs, twcc, ncc = (P[:, cols.index(c)] for c in ("bw", "tw", "cs", "twcc", "ncc"))
    one, two = np.float32(1.0), np.float32(2.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(cs == 0, one, one / cs).astype(np.float32)
        bfd = np.where(bw > tw, bw / np.float32(0.00001), np.where(bw == tw, bw / (two * z), (tw - bw) / (two * z)))
    return (fvd[:, 2::3] > bfd.astype(np.float32)[:, None]) & ((twcc > 0) & (ncc > 0))[:, None]


def mixed_fraction(over, level, key, min_width=64):
    """share of warp-steps whose 32 lanes disagree about being over bank (such a warp executes both celerity branches)"""
    order = np.lexsort((key, level))
    lv = level[order]
    starts = np.flatnonzero(np.r_[True, lv[1:] != lv[:-1]])
    ends = np.r_[starts[1:], lv.size]
    mixed = total = 0
    for a, b in zip(starts, ends):
        m = (b - a) // 32 * 32
        if m < min_width:
            continue
        ov = over[order[a:a + m]].reshape(-1, 32, over.shape[1])
        mixed += int((ov.any(axis=1) & ~ov.all(axis=1)).sum())
        total += ov.shape[0] * over.shape[1]
    return mixed / max(total, 1)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 288
    import helpers as H
    from oracle import oracle as o
    from troute_b200 import synth, hostgraph
    o.build()
    down = synth.conus_like(n_total=n, n_basins=max(4, n // 185), seed=16, style="nhd")
    case = H.make_case(down, nsteps=T)
    t0 = time.time()
    fvd, _, extras = H.oracle_route(o, case, False)
    print(f"routed {n} x {T} in {time.time() - t0:.1f} s; trip histogram {extras['iter_hist'][:8].tolist()}", flush=True)
    t0 = time.time()
    trips = trip_matrix(o, case, fvd)
    print(f"trip matrix in {time.time() - t0:.1f} s; mean trips of lanes with flow {trips[trips > 0].mean():.3f}", flush=True)
    level = hostgraph.levels(down, case["up_ptr"]).astype(np.int64)
    tot = trips.sum(axis=1).astype(np.int64)

    over = overbank_steps(case, fvd)
    print(f"lane-steps above bankfull depth: {100 * over.mean():.1f} %", flush=True)

    def report(name, key):
        u, c = efficiency(trips, level, key)
        print(f"{name:58s} efficiency {u / c:.4f}  ({32 * u / c:.2f} of 32 lanes), "
              f"{100 * mixed_fraction(over, level, key):.1f} % of the warp-steps mixed in/over bank", flush=True)

    report("caller row order", np.arange(n))
    report("sum of trips over the call (what the engine does today)", tot)
    for B in (2, 3, 4, 6, 8):
        # lexicographic key over coarse time buckets, most significant = the bucket with the largest spread
        edges = np.linspace(0, T, B + 1).astype(int)
        bs = np.stack([trips[:, a:b].sum(axis=1) for a, b in zip(edges[:-1], edges[1:])], axis=1).astype(np.int64)
        sig = np.argsort(-bs.std(axis=0))
        key = np.zeros(n, dtype=np.int64)
        for j in sig:
            key = key * (int(bs.max()) + 1) + bs[:, j]
        report(f"lexicographic, {B} time buckets (largest spread first)", key)
    from troute_b200.network import order_key_from_trips, TRIP_BUCKETS
    bk = (np.arange(T) * TRIP_BUCKETS) // T
    table = np.stack([trips[:, bk == k].sum(axis=1) for k in range(TRIP_BUCKETS)], axis=0)
    report("engine key: time-resolved trips (order_key_from_trips)", order_key_from_trips(table, T))
    report("engine key: over-bank class first, then trips", order_key_from_trips(table, T, overbank=over.sum(axis=1)))
    # 1-D embedding of the whole trip series: first principal component of the centred trip matrix
    x = trips.astype(np.float32)
    x -= x.mean(axis=0, keepdims=True)
    cov = (x.T @ x) / n
    w, v = np.linalg.eigh(cov.astype(np.float64))
    for k in (1, 2):
        pc = x @ v[:, -k].astype(np.float32)
        report(f"principal component {k} of the trip series", np.argsort(np.argsort(pc)))
    pc1 = x @ v[:, -1].astype(np.float32)
    pc2 = x @ v[:, -2].astype(np.float32)
    q1 = np.digitize(pc1, np.quantile(pc1, np.linspace(0, 1, 65)[1:-1]))
    report("64 quantile bins of PC1, then PC2 inside a bin", q1.astype(np.int64) * (1 << 40) + np.argsort(np.argsort(pc2)))
    # bound: a different order at every step (not realisable with a static layout)
    u = c = 0
    order = np.lexsort((tot, level))
    lv = level[order]
    starts = np.flatnonzero(np.r_[True, lv[1:] != lv[:-1]])
    ends = np.r_[starts[1:], lv.size]
    for a, b in zip(starts, ends):
        if b - a < 64:
            continue
        x = np.sort(trips[order[a:b]].astype(np.int32), axis=0)
        pad = (-x.shape[0]) % 32
        if pad:
            x = np.concatenate([np.zeros((pad, x.shape[1]), dtype=x.dtype), x])
        g = x.reshape(-1, 32, x.shape[1])
        u += int((g + 1)[g > 0].sum())
        mx = g.max(axis=1)
        c += int(((mx + 1) * (mx > 0)).sum()) * 32
    print(f"{'bound: re-sorted at every step':58s} efficiency {u / c:.4f}  ({32 * u / c:.2f} of 32 lanes)")


if __name__ == "__main__":
    main()
