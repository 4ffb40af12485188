"""Within-level ordering of the segments by their secant trip history (troute_b200.network.order_key_from_trips,
trt_trip_counts_bucketed): host logic and the warp-occupancy model of tools/trip_order_study.py on the CPU; the device
counters against the oracle's per-step trip counts on the GPU."""
import os
import sys

import numpy as np
import pytest

import helpers as H

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def _bucketed(trips, B):
    T = trips.shape[1]
    b = (np.arange(T) * B) // T
    return np.stack([trips[:, b == k].sum(axis=1) for k in range(B)], axis=0).astype(np.int32)


def _overbank_steps(case, fvd):
    """steps every segment ended above bankfull depth in a compound channel, with the float32 expressions of mc_channel
    (MCsingleSegStime_f2py_NOLOOP.f90:49-61) -- what the device counts (McResult.over)"""
    cols = list(case["cols"])
    P = np.asarray(case["params"], dtype=np.float32)
    bw, tw, cs, twcc, ncc = (P[:, cols.index(c)] for c in ("bw", "tw", "cs", "twcc", "ncc"))
    one, two = np.float32(1.0), np.float32(2.0)
    with np.errstate(divide="ignore", invalid="ignore"):
        z = np.where(cs == 0, one, one / cs).astype(np.float32)
        bfd = np.where(bw > tw, bw / np.float32(0.00001), np.where(bw == tw, bw / (two * z), (tw - bw) / (two * z))).astype(np.float32)
    depth = fvd[:, 2::3]
    return ((depth > bfd[:, None]) & ((twcc > 0) & (ncc > 0))[:, None])


def _mixed_fraction(over, level, key):
    """share of warp-steps (32 consecutive segments of a level in `key` order) whose lanes disagree about being over bank"""
    order = np.lexsort((key, level))
    lv = level[order]
    starts = np.flatnonzero(np.r_[True, lv[1:] != lv[:-1]])
    ends = np.r_[starts[1:], lv.size]
    mixed = total = 0
    for a, b in zip(starts, ends):
        m = (b - a) // 32 * 32
        if m < 64:
            continue
        ov = over[order[a:a + m]].reshape(-1, 32, over.shape[1])
        mixed += int((ov.any(axis=1) & ~ov.all(axis=1)).sum())
        total += ov.shape[0] * over.shape[1]
    return mixed / total


def test_order_key_is_a_ranking_and_total_breaks_ties():
    from troute_b200.network import order_key_from_trips
    total = np.array([5, 1, 3, 3], dtype=np.int32)
    assert order_key_from_trips(total, expensive_first=False).tolist() == total.tolist()   # 1-D: the totals are the key
    assert order_key_from_trips(total, expensive_first=True).tolist() == (-total).tolist() # ... expensive segments first
    # two slices of 4 steps; rows 0 and 1 are slow early, rows 2 and 3 late; the second slice spreads more
    t = np.array([[12, 12, 8, 8], [8, 9, 16, 20]], dtype=np.int32)
    key = order_key_from_trips(t, nsteps=8, expensive_first=False)
    assert sorted(key.tolist()) == [0, 1, 2, 3]
    assert key[0] < key[1] < key[2] < key[3]                                 # slice 2 first (8 < 9 < 16 < 20)
    assert order_key_from_trips(t, nsteps=8, expensive_first=True).tolist() == (3 - key).tolist()   # the same ranking, reversed
    same = np.array([[4, 4, 4], [4, 4, 4]], dtype=np.int32)
    assert order_key_from_trips(same, nsteps=2, expensive_first=False).tolist() == [0, 1, 2]        # stable for equal rows
    # quantisation: means 2.0 and 2.1 trips per step fall into one class, the total then decides
    t = np.array([[20, 21, 20], [30, 20, 20]], dtype=np.int32)
    key = order_key_from_trips(t, nsteps=20, expensive_first=False)
    assert key[1] < key[0] and key[2] < key[0]


def test_time_resolved_key_fills_the_warps_better_than_the_total(oracle):
    """The cost model of the dataflow kernel (32 lanes wait for the slowest) on the trip counts the oracle produces for an
    NHD-like network: caller order < total-trips order < time-resolved order."""
    import trip_order_study as S
    from troute_b200 import synth, hostgraph
    from troute_b200.network import order_key_from_trips, TRIP_BUCKETS
    n, T = 12000, 96
    down = synth.conus_like(n_total=n, n_basins=60, seed=16, style="nhd")
    case = H.make_case(down, nsteps=T)
    fvd, _, extras = H.oracle_route(oracle, case, False)
    trips = S.trip_matrix(oracle, case, fvd)
    hist = extras["iter_hist"]
    # the tool sees the oracle's trips: histogram slots 0..7 are exact counts, the rest hold 8-15, 16-31, ... (troute_oracle.c)
    assert np.bincount(trips.ravel().astype(np.int64), minlength=8)[:8].tolist() == hist[:8].tolist()
    assert int((trips >= 8).sum()) == int(hist[8:].sum())
    level = hostgraph.levels(down, case["up_ptr"]).astype(np.int64)

    def lanes(key):
        u, c = S.efficiency(trips, level, key)
        return 32.0 * u / c

    rows = lanes(np.arange(n))
    total = lanes(order_key_from_trips(trips.sum(axis=1)))
    key_t = order_key_from_trips(_bucketed(trips, TRIP_BUCKETS), nsteps=T)
    resolved = lanes(key_t)
    assert rows < total < resolved, (rows, total, resolved)
    assert resolved > 1.03 * total, (total, resolved)
    # over-bank steps as the primary key: far fewer warps execute both branches of the celerity, the trips stay grouped
    over = _overbank_steps(case, fvd)
    assert 0.05 < over.mean() < 0.95                       # the storm pulse floods part of the network
    key_o = order_key_from_trips(_bucketed(trips, TRIP_BUCKETS), nsteps=T, overbank=over.sum(axis=1))
    assert sorted(key_o.tolist()) == list(range(n))
    assert _mixed_fraction(over, level, key_o) < 0.75 * _mixed_fraction(over, level, key_t)
    assert lanes(key_o) > 0.99 * resolved


@pytest.mark.gpu
@pytest.mark.parametrize("mode,chunks", [(2, 1), (4, 3)])
def test_device_trip_counters_equal_the_oracle_per_time_slice(oracle, mode, chunks):
    """trt_trip_counts_bucketed == the oracle's trip count of every (segment, step) summed per slice, exactly (same
    arithmetic); the totals are their sum; a network rebuilt in that order gives the same bits."""
    import __graft_entry__ as g
    g.build()
    import trip_order_study as S
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    n, T, B = 20000, 50, 8
    case = H.make_case(synth.conus_like(n_total=n, n_basins=30, seed=9, style="nhd"), nsteps=T, warm=True)
    ref, _, _ = H.oracle_route(oracle, case, False)
    want = _bucketed(S.trip_matrix(oracle, case, ref), B)
    net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
    net.set_option("mode", mode); net.set_option("route_chunks", chunks); net.set_option("deep_lanes", 0 if mode == 2 else 2000)
    net.collect_trips(B)
    out, _ = net.route_call(T, 12, case["qlat"], case["q0"])
    got = net.trip_counts(B)
    tot = net.trip_counts()
    over = net.overbank_counts()
    key = net.trip_order_key(B)
    levels = net.levels()
    first_marching = net.last_run_stats()["first_marching_level"] if mode == 4 else levels.max() + 1
    net.close()
    H.assert_bit_equal(out, ref, "collecting run")
    assert np.array_equal(tot, got.sum(axis=0))
    wide = levels < first_marching                      # rows routed by the marching kernel report 0
    assert wide.sum() > 0.5 * n
    assert np.array_equal(got[:, wide], want[:, wide])
    assert (got[:, ~wide] == 0).all()
    want_over = _overbank_steps(case, ref).sum(axis=1)
    assert np.array_equal(over[wide], want_over[wide]) and (over[~wide] == 0).all()
    net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"], order_key=key)
    net.set_option("mode", mode)
    out, _ = net.route(T, 12, case["qlat"], case["q0"])
    net.close()
    H.assert_bit_equal(out, ref, "network ordered by the time-resolved key")
