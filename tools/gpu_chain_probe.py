"""Narrow-phase latency probe: a deep, thin network (Hack-law basin, default 60k segments) is almost entirely the
latency-bound tail of the wavefront.  Prints us/stage for a list of engine options so that scheduling changes can be
compared in seconds of GPU time.   python tools/gpu_chain_probe.py [n_segments] [kind: hack|chain]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from troute_b200 import synth
from troute_b200.network import RoutingNetwork

T = 288
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60000
kind = sys.argv[2] if len(sys.argv) > 2 else "hack"
if kind == "chain":
    down = synth.chain(n)
elif kind == "comb":
    # chain of n nodes, one headwater tributary per chain node, plus FILL isolated headwaters that blow up the footprint;
    # ids are shuffled so that the tributaries are scattered among the fillers in position order
    FILL = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    rng = np.random.default_rng(1)
    heads = rng.permutation(n + FILL)            # ids n .. 2n+FILL-1 are level-0 nodes, in random order
    down = np.full(2 * n + FILL, -1, dtype=np.int64)
    down[1:n] = np.arange(n - 1)                 # wait: chain i+1 drains into i
    down[n + heads[:n]] = np.arange(n)           # tributary of chain node i
    n = down.size
else:
    down = synth.hack_tree(n, seed=3)
params = synth.channel_params(down, seed=16)
qlat = synth.lateral_inflow(n, T, 12, seed=16)
q0 = np.zeros((n, 3), np.float32)
up_ptr, up_rows = synth.upstream_csr(down)
net = RoutingNetwork(up_ptr, up_rows, np.zeros(n, np.uint8), params, synth.PARAM_COLS)
net.upload(T, 12, qlat, q0)
print(f"{kind} n={n} levels={net.num_levels} T={T}", flush=True)


def trial(label, **opts):
    for k, v in opts.items():
        net.set_option(k, v)
    net.run(False)
    best = 1e30
    for _ in range(3):
        net.run(False)
        st = net.last_run_stats()
        best = min(best, st["kernel_ms"])
    print(f"{label:44s} wide={st['wide_ms']:8.2f} march={st['march_ms']:8.2f} kernel_ms={best:9.3f}  us/stage={1e3 * best / st['stages']:8.2f}  Mlane-steps/s={st['lane_steps'] / best / 1e3:9.1f}",
          flush=True)


if __name__ == "__main__":
    trials = [
        ("mode2 default", dict(mode=2, grid_blocks=0, gate=0)),
        ("mode3 march G=1", dict(mode=3, march_group=1)),
        ("mode3 march G=2", dict(mode=3, march_group=2)),
        ("mode3 march G=4", dict(mode=3, march_group=4)),
        ("mode3 march G=8", dict(mode=3, march_group=8)),
        ("mode3 march G=16", dict(mode=3, march_group=16)),
        ("mode3 march G=32", dict(mode=3, march_group=32)),
        ("mode4 deep_lanes=8k G=4", dict(mode=4, deep_lanes=8192, march_group=4)),
        ("mode4 deep_lanes=16k G=4", dict(mode=4, deep_lanes=16384, march_group=4)),
        ("mode4 deep_lanes=16k G=8", dict(mode=4, deep_lanes=16384, march_group=8)),
        ("mode4 deep_lanes=32k G=8", dict(mode=4, deep_lanes=32768, march_group=8)),
        ("mode4 deep_lanes=64k G=16", dict(mode=4, deep_lanes=65536, march_group=16)),
    ]
    extra = os.environ.get("TRT_PROBE_EXTRA")
    for label, opts in trials:
        try:
            trial(label, **opts)
        except Exception as e:  # noqa: BLE001
            print(label, "FAILED", e, flush=True)
