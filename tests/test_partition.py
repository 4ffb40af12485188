"""Sub-basin sharding: structural invariants of the plan, a toy linear routing recurrence that must give the same
answer sharded and unsharded (exercises the import / export wiring stage by stage), and the same thing across two
real processes over torch.distributed (gloo, world_size 2).  CPU only."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from troute_b200 import hostgraph, partition, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net(n=20000, seed=3):
    down = synth.conus_like(n_total=n, n_basins=40, seed=seed)
    up_ptr, up_rows = synth.upstream_csr(down)
    return down, up_ptr, up_rows


@pytest.mark.parametrize("P", [1, 2, 4, 8])
def test_plan_invariants(P):
    down, up_ptr, up_rows = _net()
    n = down.size
    kind = np.zeros(n, dtype=np.uint8)
    level = hostgraph.levels(down, up_ptr)
    shard, plans, stats = partition.plan_shards(down, up_ptr, up_rows, kind, P, level=level)
    assert len(plans) == P and shard.min() >= 0 and shard.max() < P
    # every segment is routed by exactly one shard
    owned = np.concatenate([p.rows[p.own] for p in plans])
    assert np.array_equal(np.sort(owned), np.arange(n))
    assert stats["imbalance"] < 1.15
    n_imports = 0
    for p in plans:
        # local CSR reproduces the global upstream lists (order included) of the rows this shard routes
        for i in np.nonzero(p.own)[0][:: max(1, p.rows.size // 500)]:
            g = p.rows[i]
            assert np.array_equal(p.rows[p.up_rows[p.up_ptr[i]:p.up_ptr[i + 1]]], up_rows[up_ptr[g]:up_ptr[g + 1]])
        # imports: boundary rows without upstream, owned by somebody else, and levels are the global ones
        assert (p.kind[p.imports] == 2).all() and (~p.own[p.imports]).all()
        assert (np.diff(p.up_ptr)[p.imports] == 0).all()
        assert np.array_equal(p.levels, level[p.rows])
        assert (shard[p.rows[p.imports]] == p.import_src).all()
        n_imports += p.imports.size
        # exports: own rows whose downstream segment lives on the destination shard
        rows, dst, glob = p.exports
        assert np.array_equal(p.rows[rows], glob) and p.own[rows].all()
        assert (shard[down[glob]] == dst).all() and (dst != p.rank).all()
    assert n_imports == stats["n_cut_edges"] == sum(p.exports[0].size for p in plans)
    if P == 1:
        assert stats["n_cut_edges"] == 0
    else:
        assert stats["n_cut_edges"] < 40 * P


def _toy_global(down, up_ptr, up_rows, level, qlat, T):
    """q[s, t] = 0.5 q[s, t-1] + 0.25 * sum_u (q[u, t] + q[u, t-1]) + qlat[s]   (same data dependences as the MC loop)"""
    n = down.size
    q = np.zeros((T + 1, n))
    order = np.argsort(level, kind="stable")
    for t in range(1, T + 1):
        for s in order:
            u = up_rows[up_ptr[s]:up_ptr[s + 1]]
            q[t, s] = 0.5 * q[t - 1, s] + 0.25 * (q[t, u].sum() + q[t - 1, u].sum()) + qlat[s]
    return q


def _toy_shard_stage(plan, q_loc, qlat, k, T, level_order):
    """Route, on one shard, every (row, t) with level + t == k; returns the exported values {(global id, t): q}."""
    out = {}
    rows, dst, glob = plan.exports
    exp_of_row = {int(r): int(g) for r, g in zip(rows, glob)}
    for i in level_order:
        if not plan.own[i]:
            continue
        t = k - int(plan.levels[i])
        if t < 1 or t > T:
            continue
        u = plan.up_rows[plan.up_ptr[i]:plan.up_ptr[i + 1]]
        q_loc[t, i] = 0.5 * q_loc[t - 1, i] + 0.25 * (q_loc[t, u].sum() + q_loc[t - 1, u].sum()) + qlat[plan.rows[i]]
        if i in exp_of_row:
            out[(exp_of_row[i], t)] = q_loc[t, i]
    return out


@pytest.mark.parametrize("P", [2, 5])
def test_toy_recurrence_sharded_equals_unsharded(P):
    down, up_ptr, up_rows = _net(n=3000, seed=8)
    n, T = down.size, 6
    level = hostgraph.levels(down, up_ptr)
    qlat = np.random.default_rng(0).uniform(0.1, 1.0, n)
    ref = _toy_global(down, up_ptr, up_rows, level, qlat, T)
    shard, plans, _ = partition.plan_shards(down, up_ptr, up_rows, np.zeros(n, np.uint8), P, pieces_per_shard=6, level=level)
    q = [np.full((T + 1, p.rows.size), np.nan) for p in plans]
    for p, ql in zip(plans, q):
        ql[0] = 0.0
    orders = [np.argsort(p.levels, kind="stable") for p in plans]
    loc = [dict(zip(p.rows.tolist(), range(p.rows.size))) for p in plans]
    for k in range(1, int(level.max()) + 1 + T):
        sent = {}
        for p, ql, od in zip(plans, q, orders):
            sent.update(_toy_shard_stage(p, ql, qlat, k, T, od))
        # "peer stores": the value lands in the import row of the shard that owns the downstream segment
        for (g, t), v in sent.items():
            d = int(shard[down[g]])
            q[d][t, loc[d][g]] = v
    for p, ql in zip(plans, q):
        assert np.array_equal(ql[:, p.own], ref[:, p.rows[p.own]])


_WORKER = r"""
import os, sys
sys.path[:0] = [{root!r}, os.path.join({root!r}, "t-route_b200"), os.path.join({root!r}, "tests")]
import numpy as np
import torch.distributed as dist
from troute_b200 import hostgraph, partition, synth
import test_partition as TP
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
down, up_ptr, up_rows = TP._net(n=3000, seed=8)
n, T = down.size, 5
level = hostgraph.levels(down, up_ptr)
qlat = np.random.default_rng(0).uniform(0.1, 1.0, n)
shard, plans, stats = partition.plan_shards(down, up_ptr, up_rows, np.zeros(n, np.uint8), world, pieces_per_shard=6, level=level)
p = plans[rank]
q = np.full((T + 1, p.rows.size), np.nan); q[0] = 0.0
od = np.argsort(p.levels, kind="stable")
loc = dict(zip(p.rows.tolist(), range(p.rows.size)))
for k in range(1, int(level.max()) + 1 + T):
    sent = TP._toy_shard_stage(p, q, qlat, k, T, od)
    everyone = [None] * world
    dist.all_gather_object(everyone, sent)          # stands in for the peer-memory stores of the CUDA kernel
    for r, msgs in enumerate(everyone):
        for (g, t), v in msgs.items():
            if int(shard[down[g]]) == rank:
                q[t, loc[g]] = v
if rank == 0:
    ref = TP._toy_global(down, up_ptr, up_rows, level, qlat, T)
else:
    ref = None
obj = [ref]
dist.broadcast_object_list(obj, src=0)
ok = np.array_equal(q[:, p.own], obj[0][:, p.rows[p.own]])
res = [None] * world
dist.all_gather_object(res, bool(ok))
dist.destroy_process_group()
assert all(res), res
print("OK", rank, stats["n_cut_edges"])
"""


def test_two_process_gloo_sharded_exchange(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port),
                   OMP_NUM_THREADS="1")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
        assert "OK" in o


def test_global_deep_level_bounds_every_shard():
    """multigpu.global_deep_level: the smallest split level such that no shard marches more than `deep_lanes` segments;
    every shard of a run uses it (a dataflow kernel must not wait for a value another shard produces while marching)."""
    from troute_b200 import multigpu, synth, hostgraph, partition
    down = synth.conus_like(n_total=30000, n_basins=20, seed=2, style="nhd")
    up_ptr, up_rows = synth.upstream_csr(down)
    level = hostgraph.levels(down, up_ptr)
    for P in (1, 2, 4):
        shard, plans, _ = partition.plan_shards(down, up_ptr, up_rows, np.zeros(down.size, np.uint8), P, pieces_per_shard=8,
                                                level=level)
        for cap in (100, 1000, 10 ** 6):
            L = multigpu.global_deep_level(level, shard, P, cap)
            per_shard = [int(((level >= L) & (shard == r)).sum()) for r in range(P)]
            assert max(per_shard) <= cap
            if L > 0:                                   # minimal: one level lower overflows some shard
                assert max(int(((level >= L - 1) & (shard == r)).sum()) for r in range(P)) > cap
        assert multigpu.global_deep_level(level, shard, P, 10 ** 6) == 0
