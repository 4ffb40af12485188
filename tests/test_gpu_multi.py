"""Two real ranks on two GPUs (torchrun-style, NCCL rendezvous + CUDA IPC peer memory).  Skipped on a 1-GPU box; run with
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r"""
import os, sys
sys.path[:0] = [{root!r}, os.path.join({root!r}, "t-route_b200"), os.path.join({root!r}, "tests")]
import numpy as np, torch, torch.distributed as dist
import helpers as H
from oracle import oracle as o
from troute_b200 import synth, multigpu
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
down = synth.conus_like(n_total=60000, n_basins=40, seed=12)
case = H.make_case(down, nsteps=36, warm=True, n_lp=40)
wl = dict(n=case["n"], down=down, params=case["params"], cols=case["cols"], qlat=case["qlat"], q0=case["q0"],
          up_ptr=case["up_ptr"], up_rows=case["up_rows"], kind=case["kind"], lp_rows=case["lp_rows"], wbody=case["wbody"])
# streamflow nudging (simple_da): 150 gages, gaps in the observations, carried-in last observations, some unknown
from troute_b200 import hostgraph
rng = np.random.default_rng(4)
G, n = 150, case["n"]
grow = np.sort(rng.choice(np.nonzero(case["kind"] == 0)[0], size=G, replace=False)).astype(np.int32)
usgs = rng.uniform(0.2, 20.0, size=(G, 18)).astype(np.float32)
usgs[rng.random(usgs.shape) < 0.3] = np.nan
lastobs = rng.uniform(0.2, 20.0, G).astype(np.float32)
since = -rng.uniform(0.0, 3600.0, G).astype(np.float32)
lastobs[::7] = np.nan; since[::7] = np.nan
level = hostgraph.levels(down, case["up_ptr"])
inv_order = np.empty(n, dtype=np.int64)
inv_order[np.argsort(level, kind="stable")] = np.arange(n)      # H.oracle_route lists one-segment reaches in this order
gages = dict(usgs_values=usgs, usgs_positions=grow, usgs_positions_reach=inv_order[grow].astype(np.int32),
             usgs_positions_gage=np.arange(G, dtype=np.int32), lastobs_values_init=lastobs,
             time_since_lastobs_init=since, da_decay_coefficient=120.0)
for short, nudging in ((False, False), (True, False), (False, True)):
    w = dict(wl, gages=gages) if nudging else wl
    r = multigpu.ShardedRouter(w, world, rank, rank, 36, 12, short, pieces_per_shard=6)
    r.upload(); r.alloc_host()
    for rep in range(2):
        r.run_e2e()
    rows, out = r.host_result()
    ref, _, extras = H.oracle_route(o, case, short, gages=gages if nudging else None)
    ok = bool(np.array_equal(out.view(np.int32), ref[rows].view(np.int32)))
    n_here = 0
    if nudging:
        plain, _, _ = H.oracle_route(o, case, short)
        assert not np.array_equal(ref, plain)                   # the gages do change the flows
        sel, nudge, lt, lv = r.gage_results()
        n_here = int(sel.size)
        same = lambda a, b: bool(np.array_equal(np.asarray(a, np.float32).view(np.int32), np.asarray(b, np.float32).view(np.int32)))
        ok = ok and same(nudge, extras["nudge"][sel]) and same(lt, extras["lastobs_times"][sel]) and same(lv, extras["lastobs_values"][sel])
    res = [None] * world
    dist.all_gather_object(res, (ok, n_here))
    assert all(x[0] for x in res), (short, nudging, res)
    if nudging:
        assert sum(x[1] for x in res) == G and min(x[1] for x in res) > 0, res   # every gage assimilated, on both GPUs
    if rank == 0:
        print("OK short_ts=%s nudging=%s cut_edges=%d" % (short, nudging, r.plan_stats["n_cut_edges"]), flush=True)
    r.close()

# consecutive windows with the state handed over ON THE DEVICES (trt_continue per shard; BASELINE configs[4] in small): two
# windows of 24 steps == the oracle's one call over 48 steps, every row of every shard (level pools and cut edges included),
# and the per-window result checksums of the ranks add up to the checksum of the unsharded table
case2 = H.make_case(down, nsteps=48, warm=True, n_lp=40)
# ... with reservoirs sitting right ABOVE cut edges: the importing shard must start from the reservoir's initial outflow
# (qd0 of the waterbody table), not from q0 (a bug of round 1 found by tools/gpu_verify_windows.py)
from troute_b200 import partition
from troute_b200._lib import TRT_KIND_LEVELPOOL
_, plans0, _ = partition.plan_shards(down, case2["up_ptr"], case2["up_rows"], case2["kind"], world, pieces_per_shard=6,
                                     level=hostgraph.levels(down, case2["up_ptr"]))
exp = np.unique(np.concatenate([p.exports[2] for p in plans0]))
cand = exp[(np.diff(case2["up_ptr"])[exp] > 0) & (case2["kind"][exp] == 0)]
assert cand.size > 0
case2["lp_rows"] = np.sort(np.concatenate([case2["lp_rows"], cand])).astype(np.int64)
case2["kind"][cand] = TRT_KIND_LEVELPOOL
case2["wbody"] = synth.levelpool_params(case2["lp_rows"].size, seed=16)
wl2 = dict(n=case2["n"], down=down, params=case2["params"], cols=case2["cols"], qlat=case2["qlat"], q0=case2["q0"],
           up_ptr=case2["up_ptr"], up_rows=case2["up_rows"], kind=case2["kind"], lp_rows=case2["lp_rows"], wbody=case2["wbody"])
ref2, _, _ = H.oracle_route(o, case2, False)
for overlap in (0, 1):
    r = multigpu.ShardedRouter(wl2, world, rank, rank, 24, 12, False, pieces_per_shard=6, windows=2)
    r.set_option("overlap_march", overlap)
    lp_on_cut = [None] * world
    dist.all_gather_object(lp_on_cut, int((case2["kind"][r.plan.exports[2]] == TRT_KIND_LEVELPOOL).sum()))
    assert sum(lp_on_cut) > 0, lp_on_cut
    r.upload()
    oks, hashes = [], []
    def on_window(w):
        out, _ = r.net.download()
        own = r.plan.own
        want = ref2[r.plan.rows[own], 3 * 24 * w:3 * 24 * (w + 1)]
        oks.append(bool(np.array_equal(out[own].view(np.int32), want.view(np.int32))))
        hashes.append(int(r.window_hash()))
    r.run_checked(on_window)
    res = [None] * world
    dist.all_gather_object(res, (oks, hashes))
    assert all(all(x[0]) for x in res), (overlap, [x[0] for x in res])
    for w in range(2):
        total = sum(x[1][w] for x in res) % (1 << 64)
        assert total == H.result_hash(ref2[:, 3 * 24 * w:3 * 24 * (w + 1)]), (overlap, w)
    if rank == 0:
        print("OK windows=2 overlap_march=%d cut_edges=%d" % (overlap, r.plan_stats["n_cut_edges"]), flush=True)
    r.close()
dist.destroy_process_group()
"""


def test_two_gpu_sharded_routing_matches_oracle(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER.format(root=ROOT))
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=900)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "OK short_ts=True" in outs[0] and "OK short_ts=False nudging=True" in outs[0], outs[0]
    assert "OK windows=2 overlap_march=0" in outs[0] and "OK windows=2 overlap_march=1" in outs[0], outs[0]
    print(outs[0])
