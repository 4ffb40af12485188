"""Routers used by bench.py and by the host interface: one RoutingNetwork per process / GPU.

SingleRouter   the whole network on one device.
ShardedRouter  one sub-basin shard per rank (see partition.py); confluence flows cross shards through peer memory.

torch is used for what it is good at here: pinned host memory, a CUDA stream the caller can time with events,
and torch.distributed for rendezvous.  All arithmetic is in libtroute_b200.so.
"""
import numpy as np

from .network import RoutingNetwork


def global_deep_level(level, shard, n_shards, deep_lanes):
    """Smallest level L such that no shard holds more than `deep_lanes` segments of level >= L."""
    L = 0
    for r in range(n_shards):
        lv = np.sort(level[shard == r])[::-1]
        if lv.size > deep_lanes:
            L = max(L, int(lv[deep_lanes]) + 1)
    return L


class SingleRouter:
    kernel_names = {0: "trt::stage_kernel", 1: "trt::persistent_kernel", 2: "trt::dataflow_kernel",
                    3: "trt::march_kernel", 4: "trt::dataflow_kernel + trt::march_kernel",
                    5: "trt::march_kernel"}

    def __init__(self, wl, device, nsteps, qts, short_ts, mode=4):
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
        self.reordered = False
        self.nq = wl["qlat"].shape[1]
        self.h2d_bytes = wl["qlat"].nbytes + wl["q0"].nbytes
        self.d2h_bytes = self.n * 3 * nsteps * 4
        self._host = None
        self._kernel_ms = []

    def upload(self):
        self.net.upload(self.T, self.qts, self.wl["qlat"], self.wl["q0"])

    def reorder_by_trip_history(self, options=None, buckets=None):
        """One calibration call that records how many secant trips every segment needed, then the network is rebuilt
        with the segments of every wavefront level ordered by that history (network.order_key_from_trips: the trips of
        16 time slices of the call): a segment's trip count repeats from step to step (p = 0.82) and moves with the storm
        pulse, so the 32 lanes of a warp -- which run in lockstep -- now mostly need the same number of trips.
        What an operational deployment would do once per network (the handle is cached across calls); the results do
        not depend on the order."""
        from ._lib import TrouteB200Error
        from .network import TRIP_BUCKETS
        buckets = TRIP_BUCKETS if buckets is None else int(buckets)
        self.net.collect_trips(buckets)
        self.net.run(self.short_ts)
        try:
            if buckets <= 1:
                raise TrouteB200Error("totals requested")
            key = self.net.trip_order_key(buckets)
            self.order_source = f"secant trip counts of {buckets} time slices of one calibration call"
        except TrouteB200Error:                      # the time-resolved table is an optimisation: the totals still order
            key = self.net.trip_counts()
            self.order_source = "secant trip counts of one calibration call"
        wl = self.wl
        self.net.close()
        self.net = RoutingNetwork(wl["up_ptr"], wl["up_rows"], wl["kind"], wl["params"], wl["cols"], device=self.device,
                                  order_key=key)
        self.net.set_option("mode", self.mode)
        self.net.set_option("stream", self.tstream.cuda_stream)
        for k, v in (options or {}).items():
            self.net.set_option(k, v)
        if len(wl.get("lp_rows", ())):
            self.net.set_levelpools(wl["lp_rows"], wl["wbody"], routing_period=wl.get("dt", 300.0))
        self.upload()
        self.reordered = True

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
        lev = self.net.levels()
        wide_rows = int((lev < st["first_marching_level"]).sum()) if self.mode == 4 else (self.n if self.mode < 3 else 0)
        return {"kernel_ms_per_call": st["kernel_ms"], "lane_steps": st["lane_steps"], "launches_per_call": launches,
                "wide_ms": st["wide_ms"], "march_ms": st["march_ms"], "wide_lane_steps": wide_rows * self.T,
                "first_marching_level": st["first_marching_level"],
                "launches_per_call_e2e": launches + 2, "stages": st["stages"],
                "levels": self.net.num_levels, "kernel_name": self.kernel_names[self.mode],
                "sharding": "single GPU"}

    def close(self):
        self.net.close()


class ShardedRouter:
    """One sub-basin shard per rank.  Cut-edge flows travel through CUDA-IPC-mapped peer memory inside the routing
    kernel (no collective on the data path); torch.distributed is used for rendezvous (IPC handles, import positions)
    and for the barrier between resetting the flow state and launching."""
    kernel_names = SingleRouter.kernel_names

    def __init__(self, wl, world, rank, device, nsteps, qts, short_ts, mode=4, pieces_per_shard=16, deep_lanes=8192):
        import torch
        import torch.distributed as dist
        from . import hostgraph, partition
        if mode not in (2, 3, 4, 5):
            raise ValueError("sharded routing needs a polling schedule (mode 2, 3 or 4)")
        self.torch, self.dist = torch, dist
        self.wl, self.world, self.rank, self.device = wl, world, rank, device
        self.T, self.qts, self.short_ts, self.mode = nsteps, qts, short_ts, mode
        level = hostgraph.levels(wl["down"], wl["up_ptr"])
        self.shard, plans, self.plan_stats = partition.plan_shards(wl["down"], wl["up_ptr"], wl["up_rows"], wl["kind"],
                                                                   world, pieces_per_shard=pieces_per_shard, level=level)
        self.plan = plan = plans[rank]
        self.n = int(plan.rows.size)
        self.n_own = int(plan.own.sum())
        # one split level for ALL shards: a dataflow (wide) kernel may wait only for values that other shards produce
        # in THEIR dataflow kernels; the deepest levels (at most deep_lanes segments on any shard) march
        self.deep_level = global_deep_level(level, self.shard, world, deep_lanes) if mode >= 4 else None
        self.tstream = torch.cuda.Stream(device=device)
        self.stream = self.tstream
        self.options = {}
        self.reordered = False
        self._build_net(None)
        self.qlat = np.ascontiguousarray(wl["qlat"][plan.rows])
        self.q0 = np.ascontiguousarray(wl["q0"][plan.rows])
        self.nq = self.qlat.shape[1]
        self.h2d_bytes = self.qlat.nbytes + self.q0.nbytes
        self.d2h_bytes = self.n * 3 * nsteps * 4
        self._host = None
        self._wired = False

    def _build_net(self, order_key):
        """The shard's device network (again, in a new within-level order, for reorder_by_trip_history)."""
        wl, plan = self.wl, self.plan
        self.net = RoutingNetwork(plan.up_ptr, plan.up_rows, plan.kind, wl["params"][plan.rows], wl["cols"],
                                  device=self.device, levels=plan.levels, order_key=order_key)
        self.net.set_option("mode", self.mode)
        lp_local = np.nonzero(plan.kind == 1)[0]
        if lp_local.size:
            # level pools of this shard: the rows of wl["wbody"] that belong to its reservoirs
            lp_index = {int(r): i for i, r in enumerate(np.asarray(wl["lp_rows"]).tolist())}
            rows = np.asarray([lp_index[int(g)] for g in plan.rows[lp_local]], dtype=np.int64)
            self.net.set_levelpools(lp_local, np.asarray(wl["wbody"])[rows], routing_period=wl.get("dt", 300.0))
        if self.deep_level is not None:
            self.net.set_option("deep_level", self.deep_level)
        self.net.set_option("stream", self.tstream.cuda_stream)
        for k, v in self.options.items():
            self.net.set_option(k, v)
        self.net.set_imports(plan.imports)
        self._wired = False

    def set_option(self, key, value):
        """Engine option that survives a rebuild of the shard's network."""
        self.options[key] = int(value)
        self.net.set_option(key, int(value))

    def reorder_by_trip_history(self, buckets=None):
        """SingleRouter.reorder_by_trip_history for a sharded run: every rank records the trips of its own segments in one
        calibration call, rebuilds its network in that order and the shards are wired again (positions and flow arrays are
        new).  Opt-in (bench.py --sharded-trip-order) until it has been measured."""
        from .network import TRIP_BUCKETS
        buckets = TRIP_BUCKETS if buckets is None else int(buckets)
        self.net.collect_trips(buckets)
        self.run_resident()
        self.net.sync()
        self.dist.barrier()
        key = self.net.trip_order_key(buckets) if buckets > 1 else self.net.trip_counts()
        self.order_source = (f"secant trip counts of {buckets} time slices of one calibration call" if buckets > 1
                             else "secant trip counts of one calibration call")
        self.dist.barrier()                         # nobody closes a flow array a peer may still be writing to
        self.net.close_peers()
        self.net.close()
        self._build_net(key)
        self.upload()
        self.reordered = True

    def _wire(self):
        """After the first upload (the flow array exists): exchange IPC handles and import positions, open peers."""
        dist, plan = self.dist, self.plan
        pos = self.net.positions()
        mine = {"handle": self.net.ipc_handle(), "n": self.n,
                "import_pos": dict(zip(plan.rows[plan.imports].tolist(), pos[plan.imports].tolist()))}
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        exp_rows, exp_shard, exp_global = plan.exports
        peers = sorted(set(exp_shard.tolist()))
        if len(peers) > 16:
            raise ValueError("more than 16 destination shards")
        for p in peers:
            self.net.open_peer(p, everyone[p]["handle"], everyone[p]["n"])
        peer_pos = np.asarray([everyone[int(s)]["import_pos"][int(g)] for s, g in zip(exp_shard, exp_global)],
                              dtype=np.int64)
        self.net.set_exports(exp_rows, exp_shard.astype(np.int32), peer_pos)
        self._wired = True
        dist.barrier()

    def upload(self):
        self.net.upload(self.T, self.qts, self.qlat, self.q0)
        if not self._wired:
            self._wire()

    def run_resident(self):
        self.net.prepare()             # flow state back to "not yet written"; complete before any peer may write
        self.dist.barrier()
        self.net.run_async(self.short_ts)

    def alloc_host(self):
        torch = self.torch
        self._host = (torch.from_numpy(self.qlat).pin_memory(), torch.from_numpy(self.q0).pin_memory(),
                      torch.empty((self.n, 3 * self.T), dtype=torch.float32, pin_memory=True))

    def run_e2e(self):
        qlat, q0, out = self._host
        self.net.upload_ptr(self.T, self.qts, qlat.data_ptr(), self.nq, q0.data_ptr())
        self.net.prepare()
        self.dist.barrier()
        self.net.run_download_ptr(self.short_ts, out.data_ptr())

    def host_result(self):
        """(global rows of this shard's own segments, their [n_own, 3T] results)"""
        return self.plan.rows[self.plan.own], self._host[2].numpy()[self.plan.own]

    def collect_stats(self):
        self.net.sync()
        st = self.net.last_run_stats()
        lev = self.net.levels()
        kinds = self.plan.kind
        wide_rows = int(((lev < st["first_marching_level"]) & (kinds != 2)).sum()) if self.mode == 4 else (
            self.n_own if self.mode < 3 else 0)
        return {"kernel_ms_per_call": st["kernel_ms"], "lane_steps": st["lane_steps"], "launches_per_call": st["launches"],
                "wide_ms": st["wide_ms"], "march_ms": st["march_ms"], "wide_lane_steps": wide_rows * self.T,
                "first_marching_level": st["first_marching_level"],
                "launches_per_call_e2e": st["launches"] + 2, "stages": st["stages"], "levels": self.net.num_levels,
                "kernel_name": self.kernel_names[self.mode],
                "sharding": f"{self.world} sub-basin shards, {self.plan_stats['n_cut_edges']} cut edges, "
                            f"imbalance {self.plan_stats['imbalance']:.3f}, peer-memory stores (no collective)"}

    def close(self):
        self.net.sync()
        self.dist.barrier()
        self.net.close_peers()
        self.net.close()
