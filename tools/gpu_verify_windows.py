"""Every row of a multi-window (trt_continue) run of the config-5 workload in small against the oracle, on one GPU and -- under
torchrun -- on N GPUs: which rows differ (kind, level, window), per-window checksums.
    python tools/gpu_verify_windows.py [segments=400000] [windows=3] [nsteps=96]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/gpu_verify_windows.py ..."""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import bench
import helpers as H

segs = int(sys.argv[1]) if len(sys.argv) > 1 else 400000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = int(sys.argv[3]) if len(sys.argv) > 3 else 96
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
args = types.SimpleNamespace(workload="conus-lp7d", segments=segs, nsteps=T, windows=W, levelpools=5000, style="nhd", short_ts=0)
wl = bench.build_workload(args)
import torch
from troute_b200 import multigpu, hostgraph
torch.cuda.set_device(rank)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
    r = multigpu.ShardedRouter(wl, world, rank, rank, T, 12, False, windows=W, deep_lanes=4096)
else:
    r = multigpu.SingleRouter(wl, 0, T, 12, False, windows=W)
r.upload()
outs, hashes = [], []
def on_window(w):
    o, _ = r.net.download()
    outs.append(o.copy()); hashes.append(int(r.window_hash()))
r.run_checked(on_window)
if world > 1:
    rows, own = r.plan.rows, r.plan.own
else:
    rows, own = np.arange(wl["n"]), np.ones(wl["n"], bool)
ref = bench.oracle_route(wl, T * W, 0, __import__("oracle.oracle", fromlist=["x"]).POW_DET, os.cpu_count() or 1)[:, 1:, :].reshape(wl["n"], -1)
level = hostgraph.levels(wl["down"], wl["up_ptr"])
tot_bad = 0
for w in range(W):
    want = ref[rows[own], 3 * T * w:3 * T * (w + 1)]
    got = outs[w][own]
    bad = ((got.view(np.int32) != want.view(np.int32)) & ~(np.isnan(got) & np.isnan(want))).any(axis=1)
    tot_bad += int(bad.sum())
    gr = rows[own][bad]
    print(f"rank {rank} window {w}: {int(bad.sum())} of {int(own.sum())} rows differ from the oracle; hash {hashes[w]:016x}; "
          f"oracle hash of these rows {H.result_hash(want, rows[own]):016x}", flush=True)
    if bad.any():
        k = wl["kind"][gr]; lv = level[gr]
        print("   kinds", np.bincount(k, minlength=3).tolist(), "levels min/median/max", int(lv.min()), int(np.median(lv)), int(lv.max()),
              "first rows", gr[:8].tolist(), flush=True)
        i = int(np.nonzero(bad)[0][0]); j = int(np.nonzero(got[i].view(np.int32) != want[i].view(np.int32))[0][0])
        print("   first diff: row", int(rows[own][i]), "column", j, "(step", j // 3 + 1, "qvd"[j % 3] + ")", got[i, j], want[i, j], flush=True)
        # roots of the error: differing rows none of whose upstream rows differ
        badset = set(gr.tolist())
        up_ptr, up_rows = wl["up_ptr"], wl["up_rows"]
        shard_of = getattr(r, "shard", None)
        for g in gr.tolist():
            ups = up_rows[up_ptr[g]:up_ptr[g + 1]].tolist()
            if any(u in badset for u in ups):
                continue
            li = int(np.nonzero(rows[own] == g)[0][0])
            cols = np.nonzero(got[li].view(np.int32) != want[li].view(np.int32))[0]
            print(f"   ROOT row {g}: kind {int(wl['kind'][g])} level {int(level[g])} upstream {ups} kinds {[int(wl['kind'][u]) for u in ups]} "
                  f"levels {[int(level[u]) for u in ups]} owners {[int(shard_of[u]) for u in ups] if shard_of is not None else None} "
                  f"first differing column {int(cols[0])} (step {int(cols[0]) // 3 + 1} {'qvd'[int(cols[0]) % 3]}) got {got[li, cols[0]]!r} want {want[li, cols[0]]!r}; "
                  f"differing columns {cols.size}; deep_level {getattr(r, 'deep_level', None)}", flush=True)
print(f"rank {rank}: total differing rows {tot_bad}", flush=True)
r.close()
if world > 1:
    dist.destroy_process_group()
