"""Small routing calls (level pools, every polling schedule, time-chunked route) for compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/gpu_sanitize.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import helpers as H
from oracle import oracle as o
from troute_b200 import synth
from troute_b200.network import RoutingNetwork

case = H.make_case(synth.conus_like(n_total=3000, n_basins=6, seed=2, style="nhd"), nsteps=16, n_lp=6, warm=True)
ref, upref, _ = H.oracle_route(o, case, False)
for mode, opts in ((2, {}), (3, {"march_group": 4}), (4, {"deep_lanes": 600})):
    net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
    net.set_levelpools(case["lp_rows"], case["wbody"])
    net.set_option("mode", mode)
    for k, v in opts.items():
        net.set_option(k, v)
    net.set_option("route_chunks", 3)
    out, up = net.route_call(16, 12, case["qlat"], case["q0"], want_upstream=True)
    net.close()
    assert np.array_equal(out.view(np.int32), ref.view(np.int32)), mode
    print("mode", mode, "ok", flush=True)
