"""First-light check on a GPU box: parity of powf / single segment / small networks vs the oracle,
then one timed run of the config-2 tree.  Prints a short report; not a test, not a benchmark."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
from oracle import oracle as o
from troute_b200 import synth
from troute_b200.network import RoutingNetwork, mc_segment_batch, powf_batch
import helpers as H

rng = np.random.default_rng(0)
n = 1_000_000
x = np.exp(rng.uniform(np.log(1e-6), np.log(1e5), n)).astype(np.float32)
y = np.array([2/3, 5/3, 0.5, 1.5], dtype=np.float32)[rng.integers(0, 4, n)]
g = powf_batch(x, y); c = o.powf(x, y, o.POW_DET)
print("powf det GPU==CPU:", np.array_equal(g.view(np.int32), c.view(np.int32)), flush=True)

case = H.make_case(synth.binary_tree(4095), nsteps=48, warm=True)
for short in (True, False):
    ref, _, ex = H.oracle_route(o, case, short)
    for mode in (0, 1, 2):
        t = time.time(); out, _, st = H.engine_route(case, short, mode=mode); dt = time.time() - t
        same = np.array_equal(out.view(np.int32), ref.view(np.int32))
        print(f"tree4095 short={short} mode={mode}: bit-equal={same} maxrel={np.max(np.abs(out-ref)/(np.abs(ref)+1e-30)):.3g} stats={st} wall={dt:.3f}", flush=True)
    print("iter hist", ex["iter_hist"], flush=True)

case = H.make_case(synth.hack_tree(60000, seed=3, hack_c=1.2), nsteps=60, n_lp=50, warm=False)
for short in (True, False):
    ref, upref, _ = H.oracle_route(o, case, short)
    out, up, st = H.engine_route(case, short, mode=1)
    print(f"hack60000+lp short={short}: bit-equal={np.array_equal(out.view(np.int32), ref.view(np.int32))} nan={np.isnan(out).sum()} "
          f"upstream-equal={np.array_equal(up[case['lp_rows']], upref[case['lp_rows']])} stats={st}", flush=True)

# timing: config 2
N, T = 1_048_576, 288
down = synth.binary_tree(N)
case = H.make_case(down, nsteps=T)
net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
for mode in (2, 1, 0):
    net.set_option("mode", mode)
    for short in (False, True):
        net.upload(T, 12, case["qlat"], case["q0"])
        for rep in range(3):
            t = time.time(); net.run(short); w = time.time() - t
        st = net.last_run_stats()
        print(f"config2 mode={mode} short={short}: kernel_ms={st['kernel_ms']:.2f} wall_ms={w*1e3:.2f} "
              f"seg-steps/s={N*T/(st['kernel_ms']*1e-3):.3e} stages={st['stages']}", flush=True)
fvd, _ = net.download()
print("finite:", np.isfinite(fvd).all(), "q outlet last:", fvd[0, -3])
