"""Per-stage latency probe: W parallel chains of length L routed for T steps (stage width ~ W*min(T, .)), per mode/grid."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from troute_b200 import synth
from troute_b200.network import RoutingNetwork

def comb(W, L):
    # W chains of L segments: segment (w, l) has id w*L + l and drains into (w, l+1)
    ids = np.arange(W * L, dtype=np.int64)
    down = ids + 1
    down[(ids % L) == L - 1] = -1
    return down

for W, L, T in ((1, 2000, 1), (32, 2000, 1), (600, 2000, 1), (600, 1000, 8), (2, 2000, 288)):
    down = comb(W, L)
    n = down.size
    params = synth.channel_params(down, seed=16)
    qlat = synth.lateral_inflow(n, max(T, 12), 12, seed=16)
    q0 = np.stack([np.full(n, 1.0), np.full(n, 1.0), np.full(n, 0.3)], axis=1).astype(np.float32)
    up_ptr, up_rows = synth.upstream_csr(down)
    for kindname, kind in (("mc", np.zeros(n, np.uint8)), ("boundary(no work)", np.full(n, 2, np.uint8))):
        if kindname != "mc" and W != 600:
            continue
        up_p, up_r = (up_ptr, up_rows) if kindname == "mc" else (np.zeros(n + 1, np.int64), np.zeros(0, np.int64))
        lv = None if kindname == "mc" else (np.arange(n) % L).astype(np.int32)
        net = RoutingNetwork(up_p, up_r, kind, params, synth.PARAM_COLS, levels=lv)
        net.upload(T, 12, qlat, q0)
        for spec in ("mode=0", "mode=1", "mode=1,grid_blocks=148", "mode=2", "mode=2,grid_blocks=148", "mode=2,grid_blocks=16"):
            opts = dict(kv.split("=") for kv in spec.split(","))
            net.set_option("grid_blocks", 0)
            for k, v in opts.items():
                net.set_option(k, int(v))
            ms = []
            for rep in range(3):
                net.run(False)
                ms.append(net.last_run_stats()["kernel_ms"])
            st = net.last_run_stats()
            print(f"W={W:4d} L={L} T={T:3d} {kindname:18s} {spec:26s} kernel_ms={min(ms):8.2f} stages={st['stages']} us/stage={1e3*min(ms)/st['stages']:7.2f}", flush=True)
        net.close()
