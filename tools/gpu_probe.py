"""Timing probe (not a benchmark): one workload, several engine option sets, kernel_ms of the wavefront kernel each.
usage: python tools/gpu_probe.py [conus|tree] [nsteps] 'mode=2,gate=0,gate_min=12' 'mode=1' ..."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from troute_b200 import synth
from troute_b200.network import RoutingNetwork

kind = sys.argv[1] if len(sys.argv) > 1 else "conus"
T = int(sys.argv[2]) if len(sys.argv) > 2 else 288
sets = sys.argv[3:] or ["mode=2"]
t0 = time.time()
down = synth.conus_like() if kind == "conus" else synth.binary_tree(1_048_576)
n = down.size
params = synth.channel_params(down, seed=16)
qlat = synth.lateral_inflow(n, T, 12, seed=16)
q0 = np.zeros((n, 3), np.float32)
up_ptr, up_rows = synth.upstream_csr(down)
net = RoutingNetwork(up_ptr, up_rows, np.zeros(n, np.uint8), params, synth.PARAM_COLS)
print(f"{kind}: n={n} levels={net.num_levels} setup {time.time()-t0:.1f}s", flush=True)
net.upload(T, 12, qlat, q0)
ref = None
for spec in sets:
    opts = dict(kv.split("=") for kv in spec.split(","))
    short = int(opts.pop("short", 0))
    for k, v in opts.items():
        net.set_option(k, int(v))
    ms = []
    for rep in range(3):
        net.run(bool(short))
        ms.append(net.last_run_stats()["kernel_ms"])
    st = net.last_run_stats()
    print(f"{spec:50s} kernel_ms={min(ms):9.2f} (runs {', '.join('%.1f' % m for m in ms)}) seg-steps/s={st['lane_steps']/(min(ms)*1e-3):.3e}", flush=True)
