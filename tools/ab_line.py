"""One line per A/B bench run for gpurun_out/box.txt:  python tools/ab_line.py gpurun_out/ab_NAME.json"""
import json
import sys

try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d["roofline"]
    print(f"value={d['value']:.4g} ms_per_step={d['ms_per_step']:.2f} dataflow_ms={r['kernel_ms']:.2f} "
          f"march_ms={r.get('marching_kernel_ms', 0.0):.2f} order={d['config'].get('within_level_order', '?')!r}")
except Exception as e:                                  # noqa: BLE001 -- a failed run must still leave its line
    print("unreadable:", type(e).__name__, e)
