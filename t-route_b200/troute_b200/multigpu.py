"""Routers used by bench.py and by the host interface: one RoutingNetwork per process / GPU.

SingleRouter   the whole network on one device.
ShardedRouter  one sub-basin shard per rank (see partition.py); confluence flows cross shards through peer memory.

torch is used for what it is good at here: pinned host memory, a CUDA stream the caller can time with events,
and torch.distributed for rendezvous.  All arithmetic is in libtroute_b200.so.
"""
import numpy as np

from .network import RoutingNetwork


class SingleRouter:
    kernel_name = "trt::persistent_kernel"

    def __init__(self, wl, device, nsteps, qts, short_ts, mode=1):
        import torch
        self.torch = torch
        self.wl = wl
        self.n = wl["n"]
        self.T = nsteps
        self.qts = qts
        self.short_ts = short_ts
        self.mode = mode
        self.device = device
        self.net = RoutingNetwork(wl["up_ptr"], wl["up_rows"], wl["kind"], wl["params"], wl["cols"], device=device)
        self.net.set_option("mode", mode)
        self.tstream = torch.cuda.Stream(device=device)
        self.stream = self.tstream
        self.net.set_option("stream", self.tstream.cuda_stream)
        self.nq = wl["qlat"].shape[1]
        self.h2d_bytes = wl["qlat"].nbytes + wl["q0"].nbytes
        self.d2h_bytes = self.n * 3 * nsteps * 4
        self._host = None
        self._kernel_ms = []

    def upload(self):
        self.net.upload(self.T, self.qts, self.wl["qlat"], self.wl["q0"])

    def run_resident(self):
        self.net.run_async(self.short_ts)

    def alloc_host(self):
        torch = self.torch
        qlat = torch.from_numpy(np.ascontiguousarray(self.wl["qlat"])).pin_memory()
        q0 = torch.from_numpy(np.ascontiguousarray(self.wl["q0"])).pin_memory()
        out = torch.empty((self.n, 3 * self.T), dtype=torch.float32, pin_memory=True)
        self._host = (qlat, q0, out)

    def run_e2e(self):
        qlat, q0, out = self._host
        self.net.route_ptr(self.T, self.qts, self.short_ts, qlat.data_ptr(), self.nq, q0.data_ptr(), out.data_ptr())

    def host_result(self):
        return self._host[2].numpy()

    def collect_stats(self):
        self.net.sync()
        st = self.net.last_run_stats()
        launches = st["launches"]
        return {"kernel_ms_per_call": st["kernel_ms"], "lane_steps": st["lane_steps"], "launches_per_call": launches,
                "launches_per_call_e2e": launches + 2 + (1 if self.n else 0), "stages": st["stages"],
                "levels": self.net.num_levels, "kernel_name": self.kernel_name if self.mode == 1 else "trt::stage_kernel",
                "sharding": "single GPU"}

    def close(self):
        self.net.close()
