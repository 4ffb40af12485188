"""Executed warp-instructions per CUDA source line of one kernel, from an `ncu --set full --import-source on` report of a
library built with -lineinfo (run where ncu is installed):

    python tools/ncu_source_lines.py gpurun_out/prof_wavefront.ncu-rep dataflow_kernel [top]

Reads `ncu -i REP --page source --print-source cuda,sass --csv`; prints the share of every source file and the `top` lines
with their active lanes per instruction."""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          f"regex:{kernel}"], capture_output=True, text=True, check=True).stdout
    cur, cols, agg = None, None, []
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            cols = r
        elif r[0].isdigit() and cols:
            def g(name):
                try:
                    return float(r[cols.index(name)])
                except (ValueError, IndexError):
                    return 0.0
            agg.append((cur, int(r[0]), " ".join(r[1].split())[:100], g("Instructions Executed"),
                        g("Thread Instructions Executed")))
    tot = sum(a[3] for a in agg)
    files = collections.Counter()
    for a in agg:
        files[a[0]] += a[3]
    print(f"{kernel}: {tot:.4g} warp-instructions; by file: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in files.most_common()))
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        print(f"{a[0]:22s} {a[1]:4d} {100 * a[3] / tot:5.2f} %  {a[4] / a[3] if a[3] else 0:4.1f} lanes  {a[2]}")


if __name__ == "__main__":
    main()
