"""GPU debugging aid: the sharded-nudging case of tests/test_gpu_parity.py in variants (with / without gages, schedules),
one line per variant with the engine's abort diagnostics.  python tools/dbg_sharded_nudging.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np
import helpers as H
from oracle import oracle
from troute_b200 import synth, partition, hostgraph, multigpu
from troute_b200.network import RoutingNetwork

oracle.build()
T = 30
down = synth.conus_like(n_total=30000, n_basins=20, seed=14, style="nhd")
case = H.make_case(down, nsteps=T, warm=True)
n = case["n"]
rng = np.random.default_rng(4)
G = 150
grow = np.sort(rng.choice(n, size=G, replace=False)).astype(np.int32)
usgs = rng.uniform(0.2, 20.0, size=(G, 18)).astype(np.float32)
usgs[rng.random(usgs.shape) < 0.3] = np.nan
lastobs = rng.uniform(0.2, 20.0, G).astype(np.float32)
since = -rng.uniform(0.0, 3600.0, G).astype(np.float32)
lastobs[::7] = np.nan; since[::7] = np.nan
level = hostgraph.levels(down, case["up_ptr"])
inv_order = np.empty(n, dtype=np.int64)
inv_order[np.argsort(level, kind="stable")] = np.arange(n)
g_oracle = dict(usgs_values=usgs, usgs_positions=grow, usgs_positions_reach=inv_order[grow].astype(np.int32),
                usgs_positions_gage=np.arange(G, dtype=np.int32), lastobs_values_init=lastobs,
                time_since_lastobs_init=since, da_decay_coefficient=120.0)
ref_g, _, extras = H.oracle_route(oracle, case, False, gages=g_oracle)
ref_0, _, _ = H.oracle_route(oracle, case, False)
q0_eff = case["q0"].copy()
obs0 = ~np.isnan(usgs[:, 0])
q0_eff[grow[obs0], 0] = usgs[obs0, 0]


def run(P, with_gages, mode, deep_lanes, grid=74):
    shard, plans, stats = partition.plan_shards(down, case["up_ptr"], case["up_rows"], case["kind"], P, pieces_per_shard=6, level=level)
    deep = multigpu.global_deep_level(level, shard, P, deep_lanes)
    nets, gsel = [], []
    for p in plans:
        net = RoutingNetwork(p.up_ptr, p.up_rows, p.kind, case["params"][p.rows], case["cols"], levels=p.levels)
        net.set_option("grid_blocks", grid); net.set_option("mode", mode)
        if mode == 4:
            net.set_option("deep_level", deep)
        net.set_imports(p.imports)
        sel = np.nonzero(np.isin(grow, p.rows[p.own]))[0]
        if with_gages:
            loc = np.searchsorted(p.rows, grow[sel]).astype(np.int32)
            nloc = p.rows.size
            net.set_gages(dict(usgs_values=usgs[sel], usgs_positions=loc, usgs_positions_reach=loc,
                               usgs_positions_gage=np.arange(sel.size, dtype=np.int32), lastobs_values_init=lastobs[sel],
                               time_since_lastobs_init=since[sel], da_decay_coefficient=120.0,
                               reach_len=np.ones(nloc, dtype=np.int64), seg_rows=np.arange(nloc)), T, routing_period=300.0)
        net.upload(T, 12, case["qlat"][p.rows], (q0_eff if with_gages else case["q0"])[p.rows])
        nets.append(net); gsel.append(sel)
    pos = [net.positions() for net in nets]
    loc_of = [dict(zip(q.rows.tolist(), range(q.rows.size))) for q in plans]
    n_exp_gage = 0
    for p, net in zip(plans, nets):
        rows, dst, glob = p.exports
        n_exp_gage += int(np.isin(glob, grow).sum())
        for d in sorted(set(dst.tolist())):
            net.set_peer_ptr(d, nets[d].state_ptr(), plans[d].rows.size)
        peer_pos = [pos[int(d)][loc_of[int(d)][int(g)]] for d, g in zip(dst, glob)]
        net.set_exports(rows, dst.astype(np.int32), np.asarray(peer_pos, dtype=np.int64))
    for net in nets:
        net.prepare()
    for net in nets:
        net.run_async(False)
    msg = []
    for i, net in enumerate(nets):
        try:
            net.sync()
        except Exception as e:
            msg.append(f"shard {i}: {e}")
    ok = None
    if not msg:
        ref = ref_g if with_gages else ref_0
        bad = 0
        for p, net in zip(plans, nets):
            out, _ = net.download()
            bad += int((out[p.own].view(np.int32) != ref[p.rows[p.own]].view(np.int32)).sum())
        ok = bad
    for net in nets:
        net.close()
    print(f"P={P} gages={with_gages} mode={mode} deep_lanes={deep_lanes} deep_level={deep} cut={stats['n_cut_edges']} "
          f"exported_gages={n_exp_gage} -> " + (f"mismatches={ok}" if ok is not None else " | ".join(msg)), flush=True)


for P in (2, 3):
    for gages in (False, True):
        for mode, dl in ((4, 2000), (2, 0), (3, 0), (4, 500)):
            try:
                run(P, gages, mode, dl)
            except Exception as e:
                print(f"P={P} gages={gages} mode={mode}: EXC {type(e).__name__}: {e}", flush=True)
