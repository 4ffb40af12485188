"""cudaMemcpy2DAsync D2H rates (the copy trt_run_download issues per time chunk), via ctypes on libcudart:
both sides strided / device side contiguous / split over several streams.   python tools/gpu_pcie_probe2.py"""
import ctypes as C, os, sys
import torch

rt = C.CDLL("libcudart.so.12")
rt.cudaMemcpy2DAsync.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
D2H = 2

def rate(fn, nbytes, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return nbytes / best / 1e6

def main():
    n, T = 2_729_077 // 2, 288
    pitch = 3 * T * 4
    d = torch.empty((n, 3 * T), dtype=torch.float32, device="cuda").normal_()
    h = torch.empty((n, 3 * T), dtype=torch.float32, pin_memory=True); h.zero_()
    streams = [torch.cuda.Stream() for _ in range(4)]
    cur = torch.cuda.current_stream()
    for chunks in (1, 2, 4, 8):
        Tc = T // chunks
        w = 3 * Tc * 4
        dc = torch.empty((n, 3 * Tc), dtype=torch.float32, device="cuda").normal_()
        for ns in (1, 2, 4):
            def both_strided():
                for c in range(chunks):
                    for k in range(ns):
                        r0, r1 = n * k // ns, n * (k + 1) // ns
                        st = streams[k] if ns > 1 else cur
                        if ns > 1: st.wait_stream(cur)
                        rt.cudaMemcpy2DAsync(h.data_ptr() + r0 * pitch + c * w, pitch, d.data_ptr() + r0 * pitch + c * w, pitch, w, r1 - r0, D2H, st.cuda_stream)
                if ns > 1:
                    for k in range(ns): cur.wait_stream(streams[k])
            def src_contig():
                for c in range(chunks):
                    for k in range(ns):
                        r0, r1 = n * k // ns, n * (k + 1) // ns
                        st = streams[k] if ns > 1 else cur
                        if ns > 1: st.wait_stream(cur)
                        rt.cudaMemcpy2DAsync(h.data_ptr() + r0 * pitch + c * w, pitch, dc.data_ptr() + r0 * w, w, w, r1 - r0, D2H, st.cuda_stream)
                if ns > 1:
                    for k in range(ns): cur.wait_stream(streams[k])
            nb = n * pitch
            print(f"chunks {chunks} (width {w:5d} B) streams {ns}: both strided {rate(both_strided, nb):6.1f} GB/s | device side contiguous {rate(src_contig, nb):6.1f} GB/s", flush=True)

main()
