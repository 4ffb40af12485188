"""Dynamic instruction mix of one kernel from an `ncu --set full --import-source on` report (run where ncu is installed):

    python tools/ncu_sass_mix.py gpurun_out/prof_wavefront.ncu-rep dataflow_kernel

Reads the SASS page (`ncu -i REP --page source --csv`): executed warp-instructions and thread-instructions per SASS
instruction -> share of every opcode class, active lanes per class, histogram of warp-instructions by active lanes, top
warp-stall reasons.  Complements tools/sass_mix.py (static counts from cuobjdump)."""
import collections
import csv
import io
import subprocess
import sys

CATS = {
    "control": {"BRA", "BSSY", "BSYNC", "EXIT", "CALL", "RET", "WARPSYNC", "BREAK", "BMOV", "NOP", "YIELD", "NANOSLEEP"},
    "fp32": {"FFMA", "FADD", "FMUL", "FMNMX", "FSEL", "FSETP", "FCHK", "MUFU", "FSET", "F2I", "I2F", "F2F", "I2FP", "F2FP"},
    "fp64": {"DFMA", "DMUL", "DADD", "DSETP", "DMNMX"},
    "memory": {"LDG", "STG", "LDS", "STS", "LDC", "LDCU", "ATOM", "ATOMG", "RED", "LD", "ST", "LDL", "STL", "ULDC", "MEMBAR",
               "CCTL", "ERRBAR"},
}


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kernel}"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    head = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    idx = {k: i for i, k in enumerate(rows[head])}
    data = rows[head + 1:]

    def f(r, k):
        try:
            return float(r[idx[k]])
        except (ValueError, IndexError):
            return 0.0

    warp = collections.Counter()
    thread = collections.Counter()
    lanes_hist = collections.Counter()
    for r in data:
        s = r[idx["Source"]].split()
        if not s:
            continue
        op = (s[1] if s[0].startswith("@") else s[0]).split(".")[0]
        w, t = f(r, "Instructions Executed"), f(r, "Thread Instructions Executed")
        warp[op] += w
        thread[op] += t
        if w > 0:
            lanes_hist[min(31, int(t / w)) // 4 * 4] += w
    tot_w, tot_t = sum(warp.values()), sum(thread.values())
    print(f"{kernel}: {tot_w:.4g} warp-instructions, {tot_t / tot_w:.2f} active lanes per instruction")
    share = collections.Counter()
    for op, c in warp.items():
        share[next((k for k, v in CATS.items() if op in v), "integer / move")] += c
    print("classes : " + ", ".join(f"{k} {100 * v / tot_w:.1f} %" for k, v in share.most_common()))
    print("opcodes : " + ", ".join(f"{o} {100 * c / tot_w:.1f} % ({thread[o] / c:.1f} lanes)" for o, c in warp.most_common(24)))
    print("lanes   : " + ", ".join(f"{k}-{k + 3}: {100 * v / tot_w:.1f} %" for k, v in sorted(lanes_hist.items())))
    stalls = {k: sum(f(r, k) for r in data) for k in idx if k.startswith("stall_") and "Not Issued" not in k}
    s = sum(stalls.values()) or 1.0
    print("stalls  : " + ", ".join(f"{k[6:]} {100 * v / s:.1f} %" for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:8]))


if __name__ == "__main__":
    main()
