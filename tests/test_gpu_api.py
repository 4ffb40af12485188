"""The reference-facing entry points on the GPU vs the oracle called with the SAME reference-style arguments:
compute_network_structured (mc_reach.pyx:164), compute_nhd_routing_v02 (compute.py:507), reach.compute_reach_kernel /
compute_reach (reach.pyx:66, :119)."""
import json
import os
from datetime import datetime

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _reference_style_case(n=4000, seed=7, n_lp=12, nsteps=36, qts=12, single=False):
    """A Hack-law basin expressed the way the reference's callers express it: arbitrary segment ids, reaches as
    lists of ids (multi-segment), reverse-connection dict, sorted data_idx, parameter table with named columns."""
    from troute_b200 import synth
    rng = np.random.default_rng(seed)
    down = synth.hack_tree(n, seed=seed)
    ids = np.sort(rng.choice(10 ** 7, size=n, replace=False)).astype(np.int64)
    up_ptr, up_rows = synth.upstream_csr(down)
    # reservoirs: in-line segments with at least one upstream neighbour; they must be singleton reaches
    cand = np.nonzero(np.diff(up_ptr) > 0)[0]
    lp_rows = np.sort(rng.choice(cand, size=n_lp, replace=False))
    is_lp = np.zeros(n, dtype=bool); is_lp[lp_rows] = True
    # reaches: chains broken at junctions and at reservoirs
    indeg = np.diff(up_ptr)
    level = synth.levels_from_down(down)
    def starts_reach(i):
        if single or indeg[i] != 1 or is_lp[i]:
            return True
        u = up_rows[up_ptr[i]]
        return bool(is_lp[u])
    heads = [i for i in range(n) if starts_reach(i)]
    reaches = []
    for h in sorted(heads, key=lambda i: (int(level[i]), i)):
        r = [h]
        cur = h
        while not is_lp[h]:
            d = int(down[cur])
            if d < 0 or starts_reach(d):
                break
            r.append(d); cur = d
        reaches.append(r)
    reaches_wTypes = [([int(ids[s]) for s in r], 1 if is_lp[r[0]] else 0) for r in reaches]
    upstream_connections = {int(ids[i]): [int(ids[u]) for u in up_rows[up_ptr[i]:up_ptr[i + 1]]] for i in range(n)}
    params = synth.channel_params(down, seed=seed)
    params[lp_rows] = np.nan                                  # lake rows have no channel parameters (compute.py:1455)
    qlat = synth.lateral_inflow(n, nsteps, qts, seed=seed)
    q0 = np.stack([rng.uniform(0, 2, n), rng.uniform(0, 2, n), rng.uniform(0, 1, n)], axis=1).astype(np.float32)
    lake_numbers = [int(ids[r]) for r in lp_rows]
    wbody = synth.levelpool_params(n_lp, seed=seed)
    return dict(n=n, ids=ids, down=down, reaches_wTypes=reaches_wTypes, upstream_connections=upstream_connections,
                params=params, cols=np.array(synth.PARAM_COLS, dtype=object), qlat=qlat, q0=q0,
                lake_numbers=lake_numbers, wbody=wbody, nsteps=nsteps, qts=qts, lp_rows=lp_rows)


def _call(fn, c, upstream_results=None, assume_short_ts=False, gages=None, **kw):
    e_f = np.zeros(0, np.float32); e_i = np.zeros(0, np.int32); e_f2 = np.zeros((0, 0), np.float32)
    g = gages or dict(usgs_values=e_f2, usgs_positions=e_i, usgs_positions_reach=e_i, usgs_positions_gage=e_i,
                      lastobs_values_init=e_f, time_since_lastobs_init=e_f, da_decay_coefficient=0.0)
    return fn(
        c["nsteps"], 300.0, c["qts"], c["reaches_wTypes"], c["upstream_connections"], c["ids"], c["cols"], c["params"],
        c["q0"], c["qlat"], c["lake_numbers"], c["wbody"], {}, np.ones((len(c["lake_numbers"]), 1), np.int32), False,
        "2021-08-23_13:00:00", g["usgs_values"], g["usgs_positions"], g["usgs_positions_reach"], g["usgs_positions_gage"],
        g["lastobs_values_init"], g["time_since_lastobs_init"], g["da_decay_coefficient"],
        e_f2, e_i, e_f, e_f, e_f, e_f, e_f,
        e_f2, e_i, e_f, e_f, e_f, e_f, e_f,
        e_f2, e_i, e_i, [], e_i, e_i, e_f, e_i, e_i,
        e_i, e_i, e_f, e_i, e_f, e_i, e_i, e_f2,
        upstream_results or {}, assume_short_ts, False, **kw)


@pytest.mark.parametrize("short_ts", [False, True])
def test_compute_network_structured_matches_oracle(oracle, short_ts):
    from troute_b200.routing.fast_reach.mc_reach import compute_network_structured, clear_network_cache
    c = _reference_style_case()
    ref = _call(oracle.compute_network_structured, c, assume_short_ts=short_ts)
    got = _call(compute_network_structured, c, assume_short_ts=short_ts)
    assert len(got) == 10 and got[2] == 0
    assert np.array_equal(got[0], ref[0])
    H.assert_bit_equal(got[1], ref[1], "flowveldepth")
    H.assert_bit_equal(got[6][c["lp_rows"]], ref[6][c["lp_rows"]], "reservoir inflow")
    assert got[8].shape == (0, c["nsteps"] + 1)
    # second call hits the cached device network and must give the same answer
    again = _call(compute_network_structured, c, assume_short_ts=short_ts)
    H.assert_bit_equal(again[1], got[1], "cached network")
    clear_network_cache()


def test_mirror_results_come_from_a_pool_of_pinned_blocks(oracle):
    """mc_reach.RESULT_POOL_BYTES / network.PinnedPool: results live in page-locked memory (DMA copies, no 9.4 GB
    allocation per CONUS call); a block is re-used only after every reference to its array is gone, so nothing a caller
    holds is overwritten; beyond the limit the mirror falls back to pageable arrays.  Same bits in every case."""
    from troute_b200.routing.fast_reach import mc_reach
    c = _reference_style_case()
    ref = _call(oracle.compute_network_structured, c)
    c2 = dict(c); c2["qlat"] = c["qlat"] * np.float32(1.5)
    ref2 = _call(oracle.compute_network_structured, c2)
    mc_reach.clear_network_cache()
    try:
        a = _call(mc_reach.compute_network_structured, c)
        H.assert_bit_equal(a[1], ref[1], "flowveldepth (pinned block)")
        H.assert_bit_equal(a[6], ref[6], "upstream_array")
        b = _call(mc_reach.compute_network_structured, c2)
        assert not np.shares_memory(a[1], b[1])                   # `a` is still held: another block
        H.assert_bit_equal(a[1], ref[1], "first result untouched by the second call")
        H.assert_bit_equal(b[1], ref2[1], "second result")
        addr_a = a[1].ctypes.data
        pinned = mc_reach._RESULT_POOL.pinned_bytes
        assert pinned >= 2 * a[1].nbytes
        del a
        import gc; gc.collect()
        d = _call(mc_reach.compute_network_structured, c)
        assert d[1].ctypes.data == addr_a and mc_reach._RESULT_POOL.pinned_bytes == pinned      # the released block, re-used
        H.assert_bit_equal(d[1], ref[1], "re-used block")
        H.assert_bit_equal(b[1], ref2[1], "held result still untouched")
        mc_reach.RESULT_POOL_BYTES = 0                            # pool off: pageable arrays
        e = _call(mc_reach.compute_network_structured, c)
        H.assert_bit_equal(e[1], ref[1], "pageable result")
    finally:
        mc_reach.RESULT_POOL_BYTES = 32 << 30
        mc_reach.clear_network_cache()


def _gage_inputs(c, n_gages, seed, obs_steps):
    """Gages at the last segment of randomly chosen reaches (T-Route breaks reaches at gages), the four observation
    regimes of simple_da: observations (with gaps), window ended -> decay from the last one, no observation at all,
    last observation carried in from before the call."""
    rng = np.random.default_rng(seed)
    reach_idx = np.sort(rng.choice(len(c["reaches_wTypes"]), size=n_gages, replace=False))
    last_ids = np.asarray([c["reaches_wTypes"][r][0][-1] for r in reach_idx], dtype=np.int64)
    order = np.argsort(last_ids)                               # gage arrays follow the sorted usgs index
    reach_idx, last_ids = reach_idx[order], last_ids[order]
    pos = np.searchsorted(c["ids"], last_ids).astype(np.int32)
    usgs = rng.uniform(0.5, 30.0, size=(n_gages, obs_steps)).astype(np.float32)
    usgs[rng.random(usgs.shape) < 0.25] = np.nan               # gaps
    usgs[: n_gages // 5] = np.nan                              # gages that never report
    lastobs = rng.uniform(0.5, 30.0, n_gages).astype(np.float32)
    since = -rng.uniform(0.0, 7200.0, n_gages).astype(np.float32)   # seconds before the start of the call
    none = rng.random(n_gages) < 0.3
    lastobs[none] = np.nan; since[none] = np.nan
    # the reference lists one (reach, gage) pair per gage, in reach order (compute.py:124-140)
    by_reach = np.argsort(reach_idx, kind="stable")
    return dict(usgs_values=usgs, usgs_positions=pos, usgs_positions_reach=reach_idx[by_reach].astype(np.int32),
                usgs_positions_gage=by_reach.astype(np.int32), lastobs_values_init=lastobs,
                time_since_lastobs_init=since, da_decay_coefficient=120.0)


@pytest.mark.parametrize("short_ts", [False, True])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_streamflow_nudging_matches_oracle(oracle, short_ts, mode):
    """simple_da on the device (mc_reach.pyx:380-411, :761-796; simple_da.pyx:21-128): replaced flows, nudge series and
    final last-observation state equal the oracle's, bit for bit, in every schedule; observation window shorter than
    the run so that the persistence/decay branch is exercised."""
    from troute_b200.routing.fast_reach import mc_reach
    c = _reference_style_case(n=6000, seed=11, n_lp=10, nsteps=48)
    gages = _gage_inputs(c, n_gages=150, seed=3, obs_steps=30)
    ref = _call(oracle.compute_network_structured, c, assume_short_ts=short_ts, gages=gages)
    nogage = _call(oracle.compute_network_structured, c, assume_short_ts=short_ts)
    assert not np.array_equal(ref[1], nogage[1])               # nudging changes the answer
    mc_reach.DEFAULT_OPTIONS = {"mode": mode}
    try:
        got = _call(mc_reach.compute_network_structured, c, assume_short_ts=short_ts, gages=gages)
    finally:
        mc_reach.DEFAULT_OPTIONS = {}
        mc_reach.clear_network_cache()
    H.assert_bit_equal(got[1], ref[1], "flowveldepth with nudging")
    H.assert_bit_equal(got[8], ref[8], "nudge")
    assert np.array_equal(got[3][0], ref[3][0])
    H.assert_bit_equal(got[3][1], ref[3][1], "lastobs_times")
    H.assert_bit_equal(got[3][2], ref[3][2], "lastobs_values")


def test_nudging_requires_gage_at_reach_end(oracle):
    from troute_b200.routing.fast_reach import mc_reach
    c = _reference_style_case(n=2000, seed=5, n_lp=0, nsteps=12)
    gages = _gage_inputs(c, n_gages=20, seed=3, obs_steps=12)
    r = next(i for i in gages["usgs_positions_reach"] if len(c["reaches_wTypes"][i][0]) > 1)
    k = list(gages["usgs_positions_reach"]).index(r)
    g = gages["usgs_positions_gage"][k]
    gages["usgs_positions"][g] = np.searchsorted(c["ids"], c["reaches_wTypes"][r][0][0])   # first, not last, segment
    with pytest.raises(NotImplementedError):
        _call(mc_reach.compute_network_structured, c, gages=gages)
    mc_reach.clear_network_cache()


def test_cached_network_is_reordered_after_first_call(oracle):
    """Large networks are rebuilt after their first call with each wavefront level ordered by the trip counts that call
    collected (mc_reach.REORDER_MIN_ROWS); results of the first and of later calls are the oracle's, bit for bit."""
    from troute_b200.routing.fast_reach import mc_reach
    c = _reference_style_case(n=5000, seed=21, n_lp=8, nsteps=24)
    gages = _gage_inputs(c, n_gages=40, seed=5, obs_steps=24)
    ref = _call(oracle.compute_network_structured, c, gages=gages)
    old = mc_reach.REORDER_MIN_ROWS
    mc_reach.REORDER_MIN_ROWS = 1000
    try:
        first = _call(mc_reach.compute_network_structured, c, gages=gages)
        entry = next(iter(mc_reach._NET_CACHE.values()))
        assert entry["ordered"] and "flat" not in entry
        second = _call(mc_reach.compute_network_structured, c, gages=gages)
    finally:
        mc_reach.REORDER_MIN_ROWS = old
        mc_reach.clear_network_cache()
    for got in (first, second):
        H.assert_bit_equal(got[1], ref[1], "flowveldepth")
        H.assert_bit_equal(got[8], ref[8], "nudge")
        H.assert_bit_equal(got[6][c["lp_rows"]], ref[6][c["lp_rows"]], "reservoir inflow")


def test_upstream_results_injection(oracle):
    """by-subnetwork hand-off (compute.py:882-900 -> mc_reach.pyx:458-469): route the part of the basin below a cut
    with the cut segment's series prescribed; rows of prescribed segments are masked out of the result."""
    from troute_b200.routing.fast_reach.mc_reach import compute_network_structured, clear_network_cache
    from troute_b200 import synth
    c = _reference_style_case(n_lp=0, single=True)
    full = _call(compute_network_structured, c)
    sizes = synth.subtree_sizes(c["down"])
    cand = np.nonzero((sizes > 300) & (sizes < 1500))[0]
    cut = int(cand[0])
    up_ptr, up_rows = synth.upstream_csr(c["down"])
    above = np.zeros(c["n"], dtype=bool)
    stack = up_rows[up_ptr[cut]:up_ptr[cut + 1]].tolist()
    while stack:
        r = stack.pop(); above[r] = True
        stack.extend(up_rows[up_ptr[r]:up_ptr[r + 1]].tolist())
    keep = ~above                                              # the cut row stays as an off-network upstream row
    ids_keep = set(c["ids"][keep].tolist())
    cut_id = int(c["ids"][cut])
    sub = dict(c)
    sub["ids"] = c["ids"][keep]; sub["params"] = c["params"][keep]; sub["qlat"] = c["qlat"][keep]; sub["q0"] = c["q0"][keep]
    sub["reaches_wTypes"] = [(r, t) for r, t in c["reaches_wTypes"] if r[0] in ids_keep and r[0] != cut_id]
    ups = {k: list(v) for k, v in c["upstream_connections"].items() if k in ids_keep}
    ups[cut_id] = []
    sub["upstream_connections"] = ups
    pos = int(np.searchsorted(sub["ids"], cut_id))
    series = full[1][np.searchsorted(full[0], cut_id)]
    ur = {cut_id: {"results": series, "position_index": pos}}
    got = _call(compute_network_structured, sub, upstream_results=ur)
    ref = _call(oracle.compute_network_structured, sub, upstream_results=ur)
    assert cut_id not in got[0].tolist()
    assert np.array_equal(got[0], ref[0])
    H.assert_bit_equal(got[1], ref[1], "cut network vs oracle")
    sel = np.searchsorted(full[0], got[0])
    H.assert_bit_equal(got[1], full[1][sel], "cut network vs uncut run")
    clear_network_cache()


def test_compute_nhd_routing_v02_frames(oracle):
    """The pandas-level entry point: same frames in, one result tuple out; values equal the oracle called through the
    reference's serial-mode slicing (compute.py:1397-1577) on each tail-water."""
    import pandas as pd
    from troute_b200 import synth
    from troute_b200.routing.compute import compute_nhd_routing_v02
    from troute_b200.routing.fast_reach.mc_reach import clear_network_cache
    rng = np.random.default_rng(5)
    down = synth.conus_like(n_total=6000, n_basins=12, seed=5)
    n = down.size
    ids = np.sort(rng.choice(10 ** 7, size=n, replace=False)).astype(np.int64)
    reaches, ups = synth.reaches_from_down(down)
    # group reaches by tail-water (outlets carry the highest level of their basin: walk downstream-first)
    tw_of = np.full(n, -1, dtype=np.int64)
    for i in np.argsort(-synth.levels_from_down(down), kind="stable"):
        tw_of[i] = i if down[i] < 0 else tw_of[down[i]]
    reaches_bytw = {}
    for r in reaches:
        reaches_bytw.setdefault(int(ids[tw_of[r[0]]]), []).append([int(ids[s]) for s in r])
    rconn = {int(ids[k]): [int(ids[u]) for u in v] for k, v in ups.items()}
    connections = {int(ids[i]): ([int(ids[down[i]])] if down[i] >= 0 else []) for i in range(n)}
    independent_networks = {tw: {s: rconn[s] for r in rl for s in r} for tw, rl in reaches_bytw.items()}
    params = synth.channel_params(down, seed=5)
    param_df = pd.DataFrame(params[:, 1:], index=ids, columns=synth.PARAM_COLS[1:])     # dt is added by the callee
    nsteps, qts = 24, 12
    qlats = pd.DataFrame(synth.lateral_inflow(n, nsteps, qts, seed=5), index=ids)
    q0 = pd.DataFrame(np.zeros((n, 3), np.float32), index=ids, columns=["qu0", "qd0", "h0"])
    empty = pd.DataFrame()
    results, sl = compute_nhd_routing_v02(
        connections, rconn, {}, reaches_bytw, "V02-structured", "by-subnetwork-jit-clustered", 10000, 4,
        datetime(2021, 8, 23, 13), 300.0, nsteps, qts, independent_networks, param_df, q0, qlats, empty, empty,
        empty, empty, empty, empty, empty, empty, empty, empty, empty, {}, False, False, empty, {}, empty, False,
        [None, None])
    assert len(results) == 1 and sl == [None, None]
    got_ids, got_fvd = results[0][0], results[0][1]
    assert np.array_equal(np.sort(got_ids), ids)
    # oracle on the same flat network
    case = H.make_case(down, nsteps=nsteps, qts=qts, seed=5)
    ref, _, _ = H.oracle_route(oracle, case, False)
    H.assert_bit_equal(got_fvd[np.argsort(got_ids)], ref, "compute_nhd_routing_v02")
    clear_network_cache()


def test_reach_entry_points(oracle):
    from troute_b200.routing.fast_reach import reach
    k = json.load(open(os.path.join(GOLD, "mc_demo_kat.json")))
    c, s = k["channel"], k["single"]
    r = reach.compute_reach_kernel(c["dt"], s["qup"], s["quc"], s["qdp"], c["ql"], c["dx"], c["bw"], c["tw"], c["twcc"],
                                   c["n"], c["ncc"], c["cs"], c["s0"], s["velp"], s["depthp"])
    assert set(r) == {"qdc", "velc", "depthc", "ck", "cn", "X"}
    assert r["depthc"] == np.float32(s["expected"]["depthc"])
    o = oracle.mc_segment(c["dt"], s["qup"], s["quc"], s["qdp"], c["ql"], c["dx"], c["bw"], c["tw"], c["twcc"],
                          c["n"], c["ncc"], c["cs"], c["s0"], s["velp"], s["depthp"], pow_mode=oracle.POW_DET)
    for key in ("qdc", "velc", "depthc", "ck", "cn", "X"):
        assert np.float32(r[key]).view(np.int32) == np.float32(o[key]).view(np.int32), key
    # compute_reach over a 4-segment reach == chaining the oracle segment by segment (reach.pyx:166-206)
    rng = np.random.default_rng(0)
    prev = rng.uniform(0.1, 2.0, (4, 3)).astype(np.float32)
    par = np.tile(np.array([[0.5, 300, 1800, 20, 30, 90, 0.05, 0.1, 0.6, 0.002]], np.float32), (4, 1))
    out = np.zeros((4, 3), np.float32)
    reach.compute_reach(np.array([1.0, 1.2], np.float32), prev, par, out)
    qup, quc = np.float32(1.0), np.float32(1.2)
    for i in range(4):
        p = par[i]
        e = oracle.mc_segment(p[1], qup, quc, prev[i, 0], p[0], p[2], p[3], p[4], p[5], p[6], p[7], p[8], p[9],
                              prev[i, 1], prev[i, 2], pow_mode=oracle.POW_DET)
        assert out[i, 0] == e["qdc"] and out[i, 1] == e["velc"] and out[i, 2] == e["depthc"]
        quc = e["qdc"]; qup = prev[i, 0]
    with pytest.raises(ValueError):
        reach.compute_reach(np.array([1.0, 1.2], np.float32), prev, par[:3], out)
    with pytest.raises(IndexError):
        reach.compute_reach(np.array([1.0], np.float32), prev, par, out)
