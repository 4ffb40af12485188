"""BASELINE configs[3] on the real hydrofabric, timed: the coastal diffusive domain of LowerColorado_TX_v4 (787 mainstem
segments in 640 reaches, 7 Muskingum-Cunge tributaries; fixtures under tests/golden) for a full day (288 x 300 s) through
trt_c_diffnw on the device, beside the CPU oracle (platform libm) on one core.  Prints one JSON line.
    python tools/gpu_lc_hybrid_timing.py [nts=288] [surveyed]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "t-route_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import test_lowercolorado as LC
import test_lowercolorado_hybrid as LH
from oracle import oracle as o, diffusive as od
from troute_b200.routing.fast_reach import diffusive

nts = int(sys.argv[1]) if len(sys.argv) > 1 else 288
LH.NTS = min(nts, LC.NTS)
o.build(); od.build()
c, dnd, results, q0, qlats, _, _ = LH.hybrid_inputs(o)
ins = LH.pack(dnd, results, q0, qlats)
segs = int(sum(ins["frnw_g"][j, 0] - 1 for j in range(int(ins["nrch_g"]))))
for _ in range(2):
    out = diffusive.compute_diffusive(ins)
wall = []
for _ in range(3):
    t0 = time.perf_counter(); out = diffusive.compute_diffusive(ins); wall.append(time.perf_counter() - t0)
table_ms, loop_ms, _ = diffusive.last_run()
t0 = time.perf_counter(); ref = od.compute_diffusive(ins, od.POW_LIBM); cpu_s = time.perf_counter() - t0
det = od.compute_diffusive(ins, od.POW_DET)
same = bool(np.array_equal(out[0], det[0], equal_nan=True) and np.array_equal(out[2], det[2], equal_nan=True))
q_ref, q_got = np.asarray(ref[0]), np.asarray(out[0])
rel = np.abs(q_got - q_ref) / np.maximum(np.abs(q_ref), 1e-30)
print(json.dumps({"workload": f"LowerColorado_TX_v4 coastal diffusive domain (tailwater {LH.TW}), {int(ins['nrch_g'])} reaches, "
                              f"{segs} segments incl. tributary junction nodes, {LH.NTS} x 300 s",
                  "device_table_ms": table_ms, "device_time_loop_ms": loop_ms, "call_wall_ms": 1e3 * float(np.mean(wall)),
                  "cpu_oracle_s": cpu_s, "speedup_vs_cpu_oracle_device": cpu_s / ((table_ms + loop_ms) * 1e-3),
                  "speedup_vs_cpu_oracle_call": cpu_s / float(np.mean(wall)),
                  "bit_equal_to_pinned_oracle": same, "frac_within_1e-5_of_libm_oracle": float((rel[np.isfinite(rel)] <= 1e-5).mean())}))
