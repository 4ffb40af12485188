"""Static SASS instruction mix of one kernel of libtroute_b200.so:  python tools/sass_mix.py dataflow_kernel"""
import collections, re, subprocess, sys
lib = "t-route_b200/troute_b200/lib/libtroute_b200.so"
pat = sys.argv[1] if len(sys.argv) > 1 else "dataflow_kernel"
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
for sec in txt.split("Function : ")[1:]:
    name = sec.split("\n", 1)[0]
    if pat not in name:
        continue
    ops = collections.Counter()
    for m in re.finditer(r"^\s*/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", sec, re.M):
        ops[m.group(1)] += 1
    print(name, "total", sum(ops.values()))
    print("  ", ", ".join(f"{o} {c}" for o, c in ops.most_common(24)))
