"""Mode-4 phase split and per-segment marching profile on the CONUS workload.
   python tools/gpu_march_profile.py [n_segments] [key=value ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from troute_b200 import synth
from troute_b200.network import RoutingNetwork

T = 288
NSEG = int(sys.argv[1]) if len(sys.argv) > 1 else 2_729_077
down = synth.conus_like(n_total=NSEG, n_basins=max(1, int(14713 * NSEG / 2_729_077)), style=os.environ.get('TRT_STYLE', 'nhd')); n = down.size
params = synth.channel_params(down, seed=16)
qlat = synth.lateral_inflow(n, T, 12, seed=16)
q0 = np.zeros((n, 3), np.float32)
up_ptr, up_rows = synth.upstream_csr(down)
net = RoutingNetwork(up_ptr, up_rows, np.zeros(n, np.uint8), params, synth.PARAM_COLS)
net.upload(T, 12, qlat, q0)
net.set_option("mode", 4)
for kv in sys.argv[2:]:
    k, v = kv.split("="); net.set_option(k, int(v))
net.set_option("march_profile", 1)
net.run(False); net.run(False)
st = net.last_run_stats()
print({k: st[k] for k in ("kernel_ms", "wide_ms", "march_ms", "first_marching_level")}, "levels", net.num_levels, flush=True)
prof = net.march_profile().astype(np.float64)
lev = net.levels()
m = prof[:, 1] > 0
print("marching segments:", int(m.sum()))
first, last, wait_cyc, fails = prof[m, 0] * 1e-3, prof[m, 1] * 1e-3, prof[m, 2], prof[m, 3]
L = lev[m]
order = np.argsort(L)
# the critical chain: per level, the time the LAST segment of that level finished its first / last step
for name, arr in (("first step done (us)", first), ("last step done (us)", last)):
    by_level = np.zeros(lev.max() + 1); np.maximum.at(by_level, L, arr)
    ls = np.unique(L)
    picks = ls[np.linspace(0, ls.size - 1, 12).astype(int)]
    print(name, [(int(l), round(float(by_level[l]), 1)) for l in picks])
    if ls.size > 10:
        slope = np.polyfit(ls[ls.size // 2:], by_level[ls[ls.size // 2:]], 1)[0]
        print("   slope over the deeper half of the levels: %.2f us per level" % slope)
dur = last - first
print("per-segment time from first to last step (us): median %.0f  p90 %.0f  max %.0f  -> us per step median %.2f" %
      (np.median(dur), np.percentile(dur, 90), dur.max(), np.median(dur) / (T - 1)))
print("solve time per step (inputs arrived -> flow published), us: median %.2f p90 %.2f ; share of the segment's wall time %.2f ; failed polls per step median %.1f" %
      (np.median(wait_cyc) * 1e-3 / T, np.percentile(wait_cyc, 90) * 1e-3 / T, np.median(wait_cyc * 1e-3 / np.maximum(1.0, last)), np.median(fails) / T))
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
rows = np.nonzero(m)[0]
np.savez(os.path.join(ROOT, "gpurun_out", "march_profile.npz"), rows=rows, level=lev[rows], prof=prof[rows])
