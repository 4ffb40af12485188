"""PCIe characteristics of the box the e2e number is measured on: D2H / H2D rate of pinned copies -- contiguous, as the
strided 2-D copy trt_run_download issues per time chunk ([n_rows] x (3 * Tc * 4) bytes out of a 3 * T * 4 pitch), from a
kernel writing straight into mapped host memory -- with the process pinned to each NUMA node in turn.
    python tools/gpu_pcie_probe.py            (gpurun; prints one line per case)"""
import os, sys, time
import torch

def numa_nodes():
    base = "/sys/devices/system/node"
    out = {}
    if os.path.isdir(base):
        for d in sorted(os.listdir(base)):
            if d.startswith("node") and d[4:].isdigit():
                cpus = open(os.path.join(base, d, "cpulist")).read().strip()
                out[int(d[4:])] = cpus
    return out

def parse_cpulist(s):
    cpus = []
    for part in s.split(","):
        if not part:
            continue
        if "-" in part:
            a, b = part.split("-"); cpus += list(range(int(a), int(b) + 1))
        else:
            cpus.append(int(part))
    return cpus

def rate(fn, nbytes, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return nbytes / best / 1e6   # GB/s

def main():
    dev = torch.device("cuda", 0)
    print("nodes:", numa_nodes(), "| affinity:", len(os.sched_getaffinity(0)), "cpus", flush=True)
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(0)
        n = (os.cpu_count() + 63) // 64
        print("gpu0 cpu affinity mask:", [hex(x) for x in pynvml.nvmlDeviceGetCpuAffinity(h, n)], flush=True)
        try:
            print("gpu0 numa node:", pynvml.nvmlDeviceGetNumaNodeId(h))
        except Exception as e:
            print("gpu0 numa node: n/a", type(e).__name__)
        print("pcie gen/width:", pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h), pynvml.nvmlDeviceGetCurrPcieLinkWidth(h))
    except Exception as e:
        print("pynvml:", type(e).__name__, e)
    os.system("nvidia-smi topo -m 2>&1 | head -12")
    n_rows, T = 2_729_077 // 4, 288          # a quarter of the result: 2.36 GB
    d = torch.empty((n_rows, 3 * T), dtype=torch.float32, device=dev).normal_()
    all_cpus = sorted(os.sched_getaffinity(0))
    nodes = numa_nodes() or {0: ",".join(map(str, all_cpus))}
    for node, cpul in nodes.items():
        cpus = [c for c in parse_cpulist(cpul) if c in all_cpus]
        if not cpus:
            continue
        os.sched_setaffinity(0, cpus)
        h = torch.empty((n_rows, 3 * T), dtype=torch.float32, pin_memory=True)
        h.zero_()                                   # first touch on this node
        nb = d.numel() * 4
        print(f"node {node}: D2H contiguous {rate(lambda: h.copy_(d, non_blocking=True), nb):6.1f} GB/s | "
              f"H2D contiguous {rate(lambda: d.copy_(h, non_blocking=True), nb):6.1f} GB/s", flush=True)
        for chunks in (4, 12, 36):
            Tc = T // chunks
            def f():
                for c in range(chunks):
                    h[:, 3 * c * Tc:3 * (c + 1) * Tc].copy_(d[:, 3 * c * Tc:3 * (c + 1) * Tc], non_blocking=True)
            print(f"node {node}: D2H as {chunks} strided chunks (width {3 * Tc * 4} B of pitch {3 * T * 4}) {rate(f, nb):6.1f} GB/s", flush=True)
        # two streams at once (two copy engines)
        s2 = torch.cuda.Stream()
        def g():
            half = n_rows // 2
            h[:half].copy_(d[:half], non_blocking=True)
            with torch.cuda.stream(s2):
                h[half:].copy_(d[half:], non_blocking=True)
            torch.cuda.current_stream().wait_stream(s2)
        print(f"node {node}: D2H contiguous on two streams {rate(g, nb):6.1f} GB/s", flush=True)
        del h
    os.sched_setaffinity(0, all_cpus)

if __name__ == "__main__":
    main()
