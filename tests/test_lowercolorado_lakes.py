"""The LowerColorado_TX NextGen hydrofabric with its 19 level-pool reservoirs broken out
(waterbody_parameters.break_network_at_waterbodies: True in test/LowerColorado_TX_v4/test_AnA_V4_HYFeature.yaml): lake
table and flowpath -> lake map read by troute_b200.hyfeatures, every lake collapsed to one node and the reaches cut at
lakes and junctions by the REFERENCE's own graph code (fixture tests/golden/lowercolorado_v4_lakes.npz, made by
tests/golden/make_golden.py), routed for 24 h with the reference's channel forcing through compute_nhd_routing_v02 with
DataFrames, the way nwm_route calls it.
CPU: the fixture is consistent, and the product's DataFrame glue (lake rows appended to the parameter table, reach types,
waterbody table slicing) with the oracle standing in for the device call equals the oracle called directly.  GPU: the
same call on the device, bit for bit."""
import os
from datetime import datetime

import numpy as np
import pandas as pd
import pytest

import helpers as H
import test_lowercolorado as LC

GOLD = LC.GOLD
NTS, QTS, DT = 96, 12, 300.0                 # 8 h are enough for the reservoirs to fill above their orifices and spill
WB_COLS = ["LkArea", "LkMxE", "OrificeA", "OrificeC", "OrificeE", "WeirC", "WeirE", "WeirL", "ifd", "qd0", "h0"]


def _load():
    mc = LC._load()
    z = np.load(os.path.join(GOLD, "lowercolorado_v4_lakes.npz"))
    nodes = z["nodes"]
    index = set(nodes.tolist())
    connections = {int(k): ([int(d)] if int(d) in index else []) for k, d in zip(nodes, z["downstream"])}
    rconn = {int(k): [] for k in nodes}
    for k, v in connections.items():
        for d in v:
            rconn[d].append(k)
    starts = np.concatenate([[0], np.cumsum(z["reach_len"])])
    reaches = [z["reach_ids"][a:b].tolist() for a, b in zip(starts[:-1], starts[1:])]
    reaches_bytw = {}
    for r, tw in zip(reaches, z["reach_tw"].tolist()):
        reaches_bytw.setdefault(int(tw), []).append(r)
    wbody_conn = dict(zip(z["wbody_seg"].tolist(), z["wbody_lake"].tolist()))
    wb = pd.DataFrame(z["lake_table"], index=pd.Index(z["lake_ids"], name="lake_id"), columns=[str(c) for c in z["lake_cols"]])
    wb["qd0"] = 0.0
    wb["h0"] = wb["OrificeE"] + 0.8 * (wb["WeirE"] - wb["OrificeE"])     # start 80 % of the way from the orifice to the weir
    return dict(mc=mc, nodes=nodes, connections=connections, rconn=rconn, reaches=reaches, reaches_bytw=reaches_bytw,
                wbody_conn=wbody_conn, waterbodies_df=wb, lake_ids=z["lake_ids"],
                link_lake=dict(zip(z["link_lake_lake"].tolist(), z["link_lake_seg"].tolist())))


def _frames(c):
    mc = c["mc"]
    ids = mc["ids"]
    param_df = pd.DataFrame(mc["params"][:, 1:], index=ids, columns=mc["cols"][1:])          # every flowpath; dt is added by the callee
    seg_index = np.concatenate([ids, c["lake_ids"]])                                            # AbstractNetwork.segment_index
    qlats = pd.DataFrame(mc["qlat"], index=ids).reindex(seg_index).fillna(0.0).astype("float32")
    q0 = pd.DataFrame(np.zeros((seg_index.shape[0], 3), np.float32), index=seg_index, columns=["qu0", "qd0", "h0"])
    return param_df, qlats, q0


def _route(c, short_ts):
    from troute_b200.routing.compute import compute_nhd_routing_v02
    param_df, qlats, q0 = _frames(c)
    e = pd.DataFrame()
    indep = {tw: c["rconn"] for tw in c["reaches_bytw"]}
    results, _ = compute_nhd_routing_v02(
        c["connections"], c["rconn"], c["wbody_conn"], c["reaches_bytw"], "V02-structured", "by-subnetwork-jit-clustered",
        10000, 36, datetime(2023, 4, 2), DT, NTS, QTS, indep, param_df, q0, qlats, e, e,
        e, e, e, e, e, e, e, e, e, {}, short_ts, False, c["waterbodies_df"][WB_COLS + ["id"]], {}, e, False, [None, None])
    (r,) = results
    order = np.argsort(r[0])
    return r[0][order], r[1][order], r[6][order]


def _oracle_direct(oracle, c, short_ts):
    """the oracle called with hand-built arguments: nodes sorted, lake rows NaN, lakes as one-node reaches of type 1"""
    mc = c["mc"]
    nodes = c["nodes"]
    lakes = set(c["lake_ids"].tolist())
    row = {int(k): i for i, k in enumerate(mc["ids"])}
    ncol = len(mc["cols"])
    params = np.full((nodes.shape[0], ncol), np.nan, dtype=np.float32)
    qlat = np.zeros((nodes.shape[0], mc["qlat"].shape[1]), dtype=np.float32)
    for i, k in enumerate(nodes.tolist()):
        if k not in lakes:
            params[i] = mc["params"][row[k]]
            qlat[i] = mc["qlat"][row[k]]
    wb = c["waterbodies_df"]
    lake_numbers = sorted(lakes)
    e_f = np.zeros(0, np.float32); e_i = np.zeros(0, np.int32); e_f2 = np.zeros((0, 0), np.float32)
    out = oracle.compute_network_structured(
        NTS, DT, QTS, [(r, 1 if set(r) & lakes else 0) for r in c["reaches"]], c["rconn"], nodes,
        np.array(mc["cols"], dtype=object), params, np.zeros((nodes.shape[0], 3), np.float32), qlat,
        lake_numbers, wb.loc[lake_numbers, WB_COLS].values, {}, np.ones((len(lake_numbers), 1), np.int32), False,
        "2023-04-02_00:00:00", e_f2, e_i, e_i, e_i, e_f, e_f, 0.0,
        e_f2, e_i, e_f, e_f, e_f, e_f, e_f, e_f2, e_i, e_f, e_f, e_f, e_f, e_f,
        e_f2, e_i, e_i, [], e_i, e_i, e_f, e_i, e_i, e_i, e_i, e_f, e_i, e_f, e_i, e_i, e_f2,
        {}, short_ts, False)
    return out[0], out[1], out[6]


def test_fixture_collapses_every_lake_into_one_node():
    c = _load()
    mc = c["mc"]
    lakes = set(c["lake_ids"].tolist())
    in_lake = set(c["wbody_conn"])
    assert len(lakes) == 19 and len(in_lake) == 248 and set(c["wbody_conn"].values()) == lakes
    # Three lakes drain straight into another lake.  The reference's replace_waterbodies_connections leaves the entry
    # flowpath of the downstream lake as their outlet although it is no longer part of the graph, and then routes it as a
    # one-segment tail-water (reservoir_shore, nhd_network.py:611-618); the fixture keeps that behaviour.
    phantom = set(c["nodes"].tolist()) & in_lake
    assert len(phantom) == 3 and all(c["connections"][p] == [] for p in phantom)
    assert set(c["nodes"].tolist()) == (set(mc["ids"].tolist()) - in_lake) | lakes | phantom        # 6971 - 248 + 19 + 3
    assert sorted(s for r in c["reaches"] for s in r) == sorted(c["nodes"].tolist())
    lake_reaches = [r for r in c["reaches"] if set(r) & lakes]
    assert len(lake_reaches) == 19 and all(len(r) == 1 for r in lake_reaches)                       # a lake is a reach of its own
    for lake in lakes:
        (outlet,) = c["connections"][lake]                                                          # one outlet ...
        assert c["wbody_conn"].get(outlet) != lake and (outlet not in in_lake or outlet in phantom) # ... outside the lake
        assert not (set(c["rconn"][lake]) & in_lake)                                                # fed from outside, if at all
        assert c["link_lake"][lake] in in_lake                                                      # outlet flowpath lies in the lake
    wb = c["waterbodies_df"]
    assert (wb["LkArea"] > 0).all() and (wb["WeirE"] > wb["OrificeE"]).all() and (wb["LkMxE"] > wb["WeirE"]).all()


@pytest.mark.parametrize("short_ts", [True, False])
def test_dataframe_glue_with_lakes_equals_the_oracle(oracle, monkeypatch, short_ts):
    from troute_b200.routing import compute

    def stand_in(*a, **k):
        k.pop("device", None)
        return oracle.compute_network_structured(*a, **k)
    monkeypatch.setitem(compute._compute_func_map, "V02-structured", stand_in)
    c = _load()
    ids, fvd, inflow = _route(c, short_ts)
    ref_ids, ref_fvd, ref_inflow = _oracle_direct(oracle, c, short_ts)
    assert np.array_equal(ids, ref_ids)
    H.assert_bit_equal(fvd, ref_fvd, "flowveldepth with reservoirs")
    lake_rows = np.searchsorted(ids, c["lake_ids"])
    H.assert_bit_equal(inflow[lake_rows], ref_inflow[lake_rows], "reservoir inflow")
    q = fvd[:, 0::3]
    d = fvd[:, 2::3]
    assert np.isfinite(fvd).all() and (q[lake_rows, -1] > 0).all()                  # every reservoir releases water
    wb = c["waterbodies_df"].loc[ids[lake_rows]]
    assert (d[lake_rows, -1] > wb["OrificeE"].values).all() and (d[lake_rows, -1] < wb["LkMxE"].values).all()
    # a reservoir attenuates: its peak outflow stays below its peak inflow
    assert (q[lake_rows].max(axis=1) <= inflow[lake_rows].max(axis=1) + 1e-3).sum() >= 15


@pytest.mark.gpu
@pytest.mark.parametrize("short_ts", [True, False])
def test_gpu_routes_lowercolorado_with_its_reservoirs(oracle, short_ts):
    from troute_b200.routing.fast_reach.mc_reach import clear_network_cache
    c = _load()
    try:
        ids, fvd, inflow = _route(c, short_ts)
    finally:
        clear_network_cache()
    ref_ids, ref_fvd, ref_inflow = _oracle_direct(oracle, c, short_ts)
    assert np.array_equal(ids, ref_ids)
    H.assert_bit_equal(fvd, ref_fvd, "LowerColorado with reservoirs: flowveldepth")
    lake_rows = np.searchsorted(ids, c["lake_ids"])
    H.assert_bit_equal(inflow[lake_rows], ref_inflow[lake_rows], "reservoir inflow")


def test_windows_with_reservoirs_equal_one_call(oracle, monkeypatch):
    """The driver loop on the real network with its reservoirs: four 24-step windows through nwm_routing.route_windows
    (nwm_route per window, then new_q0 and update_waterbody_water_elevation: a reservoir starts the next window from its last
    outflow and water elevation) == one 96-step call.  Oracle as the stand-in for the device call."""
    from troute_b200 import nwm_routing
    from troute_b200.routing import compute

    def stand_in(*a, **k):
        k.pop("device", None)
        return oracle.compute_network_structured(*a, **k)
    monkeypatch.setitem(compute._compute_func_map, "V02-structured", stand_in)
    c = _load()
    _, full, _ = _route(c, True)
    param_df, qlats, q0 = _frames(c)
    wb = c["waterbodies_df"][WB_COLS + ["id"]].copy()
    e = pd.DataFrame()
    indep = {tw: c["rconn"] for tw in c["reaches_bytw"]}
    W, n_w = 4, NTS // 4                                            # 24 steps = 2 forcing hours per window (qlat is hourly)

    def route_window(w, q0_df, wb_df, lo_df):
        ql = qlats.iloc[:, w * (n_w // QTS):(w + 1) * (n_w // QTS)]
        return nwm_routing.nwm_route(
            c["connections"], c["rconn"], c["wbody_conn"], c["reaches_bytw"], "by-subnetwork-jit-clustered", "V02-structured",
            10000, 36, datetime(2023, 4, 2), DT, n_w, QTS, indep, param_df, q0_df, ql, e, e, e, e, e, e, e, e, e, e, e, {},
            True, False, wb_df, {}, e, False, None, e, None, None, [None, None], e, e)

    per_window, q0_end, wb_end, _ = nwm_routing.route_windows(route_window, range(W), q0, wb, e, DT, n_w)
    pieces = []
    for results in per_window:
        (r,) = results
        pieces.append(r[1][np.argsort(r[0])])
    H.assert_bit_equal(np.concatenate(pieces, axis=1), full, "4 windows with reservoirs vs one call")
    ids = np.sort(per_window[-1][0][0])
    lake_rows = np.searchsorted(ids, c["lake_ids"])
    assert np.array_equal(wb_end.loc[c["lake_ids"], "h0"].to_numpy(np.float32), full[lake_rows, -1])      # last water elevation
    assert np.array_equal(wb_end.loc[c["lake_ids"], "qd0"].to_numpy(np.float32), full[lake_rows, -3])     # last outflow


REF_COMPUTE = "/root/reference/src/troute-routing/troute/routing/compute.py"


def _reference_compute_nhd_routing_v02(kernel):
    """the reference's compute_nhd_routing_v02 and its module-level helpers, compiled out of compute.py (whose imports need
    the Cython kernels) with `kernel` registered as the compute function"""
    import ast
    import copy
    import logging
    import time
    from collections import defaultdict
    from datetime import timedelta
    from functools import partial
    from itertools import chain
    tree = ast.parse(open(REF_COMPUTE).read())
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name != "compute_diffusive_routing"]
    ns = {"pd": pd, "np": np, "time": time, "LOG": logging.getLogger("reference"), "chain": chain, "defaultdict": defaultdict,
          "partial": partial, "datetime": datetime, "timedelta": timedelta, "copy": copy, "os": os,
          "_compute_func_map": defaultdict(lambda: kernel, {"V02-structured": kernel})}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "reference compute.py", "exec"), ns)
    return ns["compute_nhd_routing_v02"]


@pytest.mark.skipif(not os.path.exists(REF_COMPUTE), reason="pins compute_nhd_routing_v02 against the reference tree, present only in the build container")
def _gage_frames(c, n_gages=25, obs_steps=30, seed=3):
    """usgs_df (one row per gage segment, one column per routing step, gaps) and lastobs_df as the DA object hands them to
    nwm_route; gages sit on the last segment of channel reaches (T-Route breaks reaches at gages)"""
    rng = np.random.default_rng(seed)
    lakes = set(c["lake_ids"].tolist())
    cand = [r[-1] for r in c["reaches"] if not (set(r) & lakes)]
    segs = np.sort(rng.choice(np.asarray(cand, dtype=np.int64), size=n_gages, replace=False))
    obs = rng.uniform(0.5, 30.0, size=(n_gages, obs_steps))
    obs[rng.random(obs.shape) < 0.25] = np.nan
    t0 = datetime(2023, 4, 2)
    usgs_df = pd.DataFrame(obs, index=segs, columns=pd.date_range(t0, periods=obs_steps, freq="300s"))
    since = -300.0 * rng.integers(0, 12, n_gages).astype(np.float64)
    lastobs_df = pd.DataFrame({"time_since_lastobs": since, "lastobs_discharge": rng.uniform(0.5, 30.0, n_gages)}, index=segs)
    lastobs_df.iloc[::6] = np.nan                                            # gages without a previous observation
    return usgs_df, lastobs_df


@pytest.mark.parametrize("gages", [False, True])
@pytest.mark.parametrize("short_ts", [True, False])
def test_compute_nhd_routing_v02_equals_the_reference_function(oracle, monkeypatch, short_ts, gages):
    """The drop-in boundary B1 against the function it replaces: the reference's own compute_nhd_routing_v02
    (compute.py:507-1738, `serial` branch: one kernel call per tail-water, frames sliced per network) and the mirror (one
    call for all tail-waters), BOTH with the oracle as the compute kernel, on the real hydrofabric with its 19 reservoirs and
    4 tail-waters: same segment ids, bit-identical flow / velocity / depth, same reservoir inflow series."""
    from troute_b200.routing import compute

    def kernel(*a, **k):
        k.pop("device", None)
        return oracle.compute_network_structured(*a, **k)
    monkeypatch.setitem(compute._compute_func_map, "V02-structured", kernel)
    c = _load()
    param_df, qlats, q0 = _frames(c)
    e = pd.DataFrame()
    indep = {tw: c["rconn"] for tw in c["reaches_bytw"]}
    wb = c["waterbodies_df"][WB_COLS + ["id"]]
    ref_fn = _reference_compute_nhd_routing_v02(kernel)
    usgs_df, lastobs_df = _gage_frames(c) if gages else (e, e)
    da = {"da_decay_coefficient": 120.0} if gages else {}
    args = (c["connections"], c["rconn"], c["wbody_conn"], c["reaches_bytw"], "V02-structured", "serial", 10000, 1,
            datetime(2023, 4, 2), DT, NTS, QTS, indep, param_df, q0, qlats, usgs_df, lastobs_df, e, e, e, e, e, e, e, e, e, da,
            short_ts, False, wb, {}, e, False, [None, None])
    ref, _ = ref_fn(*args)
    assert len(ref) == len(c["reaches_bytw"]) == 4                                   # one tuple per tail-water
    ids = np.concatenate([r[0] for r in ref])
    order = np.argsort(ids)
    if gages:
        # streamflow nudging: the mirror prepares the gage arrays for ONE network, the reference per tail-water
        (r,) = compute.compute_nhd_routing_v02(*(args[:5] + ("by-subnetwork-jit-clustered",) + args[6:]))[0]
        o2 = np.argsort(r[0])
        got_ids, got_fvd, got_inflow = r[0][o2], r[1][o2], r[6][o2]
        plain = _route(c, short_ts)[1]
        assert not np.array_equal(got_fvd, plain)                                    # the gages change the answer
        # last-observation state per gage, whatever the grouping into tuples
        ref_lo = {int(g): (t, v) for rr in ref for g, t, v in zip(rr[3][0], rr[3][1], rr[3][2])}
        got_lo = {int(g): (t, v) for g, t, v in zip(r[3][0], r[3][1], r[3][2])}
        assert set(ref_lo) == set(got_lo) == set(usgs_df.index.tolist())
        for g in ref_lo:
            assert np.array_equal(np.float32(ref_lo[g]), np.float32(got_lo[g]), equal_nan=True), g
    else:
        got_ids, got_fvd, got_inflow = _route(c, short_ts)
    assert np.array_equal(ids[order], got_ids)
    H.assert_bit_equal(got_fvd, np.concatenate([r[1] for r in ref])[order], "mirror vs the reference function: flowveldepth")
    lake_rows = np.searchsorted(got_ids, c["lake_ids"])
    H.assert_bit_equal(got_inflow[lake_rows], np.concatenate([r[6] for r in ref])[order][lake_rows], "reservoir inflow")
