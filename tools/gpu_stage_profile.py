"""Where does the time go, stage by stage (mode 0 + events)?  CONUS workload."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from troute_b200 import synth
from troute_b200.network import RoutingNetwork
T = 288
NSEG = int(sys.argv[2]) if len(sys.argv) > 2 else 2_729_077
down = synth.conus_like(n_total=NSEG, n_basins=max(1, int(14713 * NSEG / 2_729_077)), style=os.environ.get('TRT_STYLE', 'nhd')); n = down.size
params = synth.channel_params(down, seed=16)
qlat = synth.lateral_inflow(n, T, 12, seed=16)
q0 = np.zeros((n, 3), np.float32)
up_ptr, up_rows = synth.upstream_csr(down)
net = RoutingNetwork(up_ptr, up_rows, np.zeros(n, np.uint8), params, synth.PARAM_COLS)
net.upload(T, 12, qlat, q0)
MODE = int(sys.argv[1]) if len(sys.argv) > 1 else 2
net.set_option("mode", MODE); net.set_option("profile_stages", 1)
for kv in sys.argv[3:]:
    k, v = kv.split("="); net.set_option(k, int(v))
net.run(False); net.run(False)
ms, w = net.stage_profile()
print("mode", MODE, "(mode 2: time between consecutive stage completions)")
print("total kernel_ms", net.last_run_stats()["kernel_ms"], "sum stage ms", ms.sum())
print("n", n, "levels", net.num_levels, "stages", ms.size - 1)
for lo, hi in ((1, 50), (50, 150), (150, 289), (289, 400), (400, 600), (600, 1000), (1000, 2000), (2000, 3000), (3000, 4571)):
    if lo >= ms.size:
        break
    sl = slice(lo, min(hi, ms.size))
    print(f"stages {lo:5d}-{hi:5d}: ms={ms[sl].sum():8.2f} lanes={w[sl].sum():12d} mean us/stage={1e3*ms[sl].mean():8.1f} mean width={w[sl].mean():10.0f} ns/lane={1e6*ms[sl].sum()/max(1,w[sl].sum()):.3f}")
top = np.argsort(-ms)[:25]
print("slowest stages:", [(int(k), round(float(ms[k]), 3), int(w[k])) for k in top])
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.save(os.path.join(ROOT, "gpurun_out", f"stage_profile_ms_mode{MODE}.npy"), ms)
