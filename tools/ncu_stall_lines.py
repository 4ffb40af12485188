"""Warp-stall samples per CUDA source line of one kernel (ncu --set full --import-source on, library built with -lineinfo):

    python tools/ncu_stall_lines.py REP.ncu-rep kernel_regex [stall_column=stall_long_sb] [top=30]

Reads `ncu -i REP --page source --print-source cuda,sass --csv`; prints the kernel's total samples per stall reason and the
`top` source lines for the chosen reason."""
import collections, csv, io, subprocess, sys


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    col = sys.argv[3] if len(sys.argv) > 3 else "stall_long_sb"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv", "--kernel-name",
                          f"regex:{kernel}"], capture_output=True, text=True, check=True).stdout
    cur, cols, rows = None, None, []
    for r in csv.reader(io.StringIO(txt)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif r[0] == "Line No":
            cols = r
        elif r[0].isdigit() and cols:
            rows.append((cur, int(r[0]), " ".join(r[1].split())[:90], {c: float(v) for c, v in zip(cols, r) if c.startswith("stall_") and "Not Issued" not in c and v.replace(".", "").isdigit()}))
    tot = collections.Counter()
    for _, _, _, d in rows:
        tot.update(d)
    allsum = sum(tot.values())
    print(f"{kernel}: {allsum:.0f} stall samples: " + ", ".join(f"{k[6:]} {100 * v / allsum:.1f}%" for k, v in tot.most_common(10)))
    for f, ln, src, d in sorted(rows, key=lambda a: -a[3].get(col, 0))[:top]:
        print(f"{f:22s} {ln:4d} {100 * d.get(col, 0) / max(tot[col], 1):5.1f} % of {col[6:]}  {src}")


if __name__ == "__main__":
    main()
