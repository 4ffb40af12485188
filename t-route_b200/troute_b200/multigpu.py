"""Routers used by bench.py and by the host interface: one RoutingNetwork per process / GPU.

SingleRouter   the whole network on one device.
ShardedRouter  one sub-basin shard per rank (see partition.py); confluence flows cross shards through peer memory.

Both route a call as `windows` consecutive windows of `nsteps` steps (BASELINE configs[4]: a 7-day hindcast = 7 windows of
288 steps): window 0 starts from q0, every later one from the state the previous window left ON THE DEVICE (trt_continue)
-- the device-resident form of the reference's window loop (nwm_routing/__main__.py:258-266, troute_model.py:250-262).

torch is used for what it is good at here: pinned host memory, device buffers for resident forcing, a CUDA stream the
caller can time with events, and torch.distributed for rendezvous.  All arithmetic is in libtroute_b200.so.
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


def _window_cols(nsteps, qts, w):
    """qlat columns of window w (every window has nsteps steps; nsteps must be a multiple of qts for w > 0)"""
    per = -(-nsteps // qts)
    return slice(w * per, (w + 1) * per)


def _q0_with_step0_observations(wl):
    """Initial state [n, 3] every router starts from.  Two things make q[s, 0] differ from the q0 the caller hands in, and both
    are applied by the network that OWNS the segment on the device -- but a shard that only IMPORTS the segment (it feeds
    one of its own across a cut edge) reads q[u, 0] from its import row, which is initialised from the q0 it was given:
      * a gage with an observation at step 0 replaces the initial flow of its segment (mc_reach.pyx:403-411;
        reset_gages_kernel);
      * a level pool starts from the outflow of the waterbody table, `qd0` (mc_reach.pyx:298, :359-361;
        init_levelpool_kernel), not from q0.
    So the replaced values go into the q0 of every router (found by tools/gpu_verify_windows.py: 164 rows downstream of the
    cut edges below reservoirs were off at the first steps of a sharded run)."""
    g = wl.get("gages")
    q0 = wl["q0"]
    lp_rows = np.asarray(wl.get("lp_rows", ()), dtype=np.int64)
    have_gages = bool(g) and len(g["usgs_positions"]) > 0
    if not have_gages and lp_rows.size == 0:
        return q0
    q0 = np.array(q0, dtype=np.float32, copy=True)
    if lp_rows.size:
        q0[lp_rows, 0] = np.asarray(wl["wbody"], dtype=np.float64)[:, 9].astype(np.float32)
    if have_gages:
        usgs = np.asarray(g["usgs_values"], dtype=np.float32).reshape(len(g["usgs_positions"]), -1)
        if usgs.shape[1] > 0:
            obs0 = ~np.isnan(usgs[:, 0])
            q0[np.asarray(g["usgs_positions"])[obs0], 0] = usgs[obs0, 0]
    return q0


class _RouterBase:
    def _set_gages(self, local_rows_of_global):
        """Streamflow nudging (simple_da, mc_reach.pyx:380-411, :761-796) for the gages of wl["gages"] (the reference's
        arguments with GLOBAL rows in usgs_positions; every gage ends its reach) that sit on rows this router owns.
        `local_rows_of_global(rows)` -> local row or -1.  Remembers which gages are here (self.gage_sel)."""
        g = self.wl.get("gages")
        self.gage_sel = np.zeros(0, dtype=np.int64)
        if not g or len(g["usgs_positions"]) == 0:
            return
        grow = np.asarray(g["usgs_positions"], dtype=np.int64)
        loc_all = np.asarray(local_rows_of_global(grow), dtype=np.int64)
        sel = np.nonzero(loc_all >= 0)[0]
        self.gage_sel = sel
        loc = loc_all[sel].astype(np.int32)
        G = len(grow)
        usgs = np.asarray(g["usgs_values"], dtype=np.float32).reshape(G, -1)
        self.net.set_gages(dict(usgs_values=usgs[sel], usgs_positions=loc, usgs_positions_reach=loc,
                                usgs_positions_gage=np.arange(sel.size, dtype=np.int32),
                                lastobs_values_init=np.asarray(g["lastobs_values_init"], dtype=np.float32)[sel],
                                time_since_lastobs_init=np.asarray(g["time_since_lastobs_init"], dtype=np.float32)[sel],
                                da_decay_coefficient=g["da_decay_coefficient"],
                                reach_len=np.ones(self.n, dtype=np.int64), seg_rows=np.arange(self.n)),
                           self.T, routing_period=self.wl.get("dt", 300.0))

    def gage_results(self):
        """(indices into wl["gages"] of the gages routed here, nudge [*, T + 1], lastobs_times, lastobs_values) of the last window"""
        if self.gage_sel.size == 0:
            return self.gage_sel, np.zeros((0, self.T + 1), np.float32), np.zeros(0, np.float32), np.zeros(0, np.float32)
        nudge, lt, lv = self.net.download_gages()
        return self.gage_sel, nudge, lt, lv

    kernel_names = {0: "trt::stage_kernel", 1: "trt::persistent_kernel", 2: "trt::dataflow_kernel",
                    3: "trt::march_kernel", 4: "trt::dataflow_kernel + trt::march_kernel"}

    # ---- forcing ------------------------------------------------------------------------------------------------
    def _make_resident(self):
        """Device copies of the forcing of every window (the `value` leg: inputs resident in HBM before the timed region)."""
        torch = self.torch
        self._q0_dev = torch.from_numpy(np.ascontiguousarray(self.q0)).to(f"cuda:{self.device}")
        self._qlat_dev = [torch.from_numpy(np.ascontiguousarray(self.qlat[:, _window_cols(self.T, self.qts, w)])).to(f"cuda:{self.device}")
                          for w in range(self.windows)]
        torch.cuda.synchronize(self.device)

    def upload(self):
        """window 0 of the call from host arrays (set-up path: calibration, first touch of the flow array)"""
        self.net.upload(self.T, self.qts, np.ascontiguousarray(self.qlat[:, _window_cols(self.T, self.qts, 0)]), self.q0)

    def _start_window(self, w, resident):
        if resident:
            if self._qlat_dev is None:
                self._make_resident()
            ql = self._qlat_dev[w]
            if w == 0:
                self.net.upload_ptr(self.T, self.qts, ql.data_ptr(), ql.shape[1], self._q0_dev.data_ptr())
            else:
                self.net.continue_ptr(self.T, self.qts, ql.data_ptr(), ql.shape[1])
        else:
            qlat, q0, _ = self._host
            ql = qlat[w]
            if w == 0:
                self.net.upload_ptr(self.T, self.qts, ql.data_ptr(), ql.shape[1], q0.data_ptr())
            else:
                self.net.continue_ptr(self.T, self.qts, ql.data_ptr(), ql.shape[1])

    def alloc_host(self, result=True):
        """Pinned host copies of the forcing and (result=True) the pinned [n, 3T] result buffer of the end-to-end leg."""
        torch = self.torch
        qlat = [torch.from_numpy(np.ascontiguousarray(self.qlat[:, _window_cols(self.T, self.qts, w)])).pin_memory()
                for w in range(self.windows)]
        q0 = torch.from_numpy(np.ascontiguousarray(self.q0)).pin_memory()
        out = torch.empty((self.n, 3 * self.T), dtype=torch.float32, pin_memory=True) if result else None
        self._host = (qlat, q0, out)
        self.h2d_bytes = sum(int(q.numel()) * 4 for q in qlat) + int(q0.numel()) * 4
        self.d2h_bytes = self.n * 3 * self.T * 4 * self.windows

    def host_result(self):
        return self._host[2].numpy()

    # ---- verification -------------------------------------------------------------------------------------------
    def window_hash(self):
        """64-bit checksum of the device-resident result of the last window over this router's OWN rows, labelled with
        their global row numbers: the checksums of all ranks add up (mod 2^64) to the checksum of the unsharded result."""
        return self.net.result_hash(rows=self.own_local_rows, ids=self.own_global_rows)

    def download_global_rows(self, global_rows):
        """(global rows among `global_rows` that this router owns, their result rows of the last window)"""
        loc = self.local_of_global(np.asarray(global_rows, dtype=np.int64))
        mine = loc >= 0
        return np.asarray(global_rows, dtype=np.int64)[mine], self.net.download_rows(loc[mine])


class SingleRouter(_RouterBase):
    def __init__(self, wl, device, nsteps, qts, short_ts, mode=4, windows=1):
        import torch
        self.torch = torch
        self.wl = wl
        self.n = wl["n"]
        self.T = nsteps
        self.qts = qts
        self.short_ts = short_ts
        self.mode = mode
        self.device = device
        self.windows = int(windows)
        self.qlat, self.q0 = wl["qlat"], _q0_with_step0_observations(wl)
        self.tstream = torch.cuda.Stream(device=device)
        self.stream = self.tstream
        self.options = {}
        self.reordered = False
        self._build_net(None)
        self.nq = wl["qlat"].shape[1]
        self.h2d_bytes = wl["qlat"].nbytes + wl["q0"].nbytes
        self.d2h_bytes = self.n * 3 * nsteps * 4 * self.windows
        self._host = None
        self._qlat_dev = None
        self.own_local_rows = None                     # every row, labelled by its row number
        self.own_global_rows = None

    def _build_net(self, order_key):
        wl = self.wl
        self.net = RoutingNetwork(wl["up_ptr"], wl["up_rows"], wl["kind"], wl["params"], wl["cols"], device=self.device,
                                  order_key=order_key)
        self.net.set_option("mode", self.mode)
        self.net.set_option("stream", self.tstream.cuda_stream)
        for k, v in self.options.items():
            self.net.set_option(k, v)
        if len(wl.get("lp_rows", ())):
            self.net.set_levelpools(wl["lp_rows"], wl["wbody"], routing_period=wl.get("dt", 300.0))
        self._set_gages(lambda rows: rows)

    def set_option(self, key, value):
        self.options[key] = int(value)
        self.net.set_option(key, int(value))

    def local_of_global(self, rows):
        return rows

    def reorder_by_trip_history(self, buckets=None, calibration_qlat=None):
        """One calibration call that records how many secant trips every segment needed, then the network is rebuilt
        with the segments of every wavefront level ordered by that history (network.order_key_from_trips: the trips of
        16 time slices of the call): a segment's trip count repeats from step to step (p = 0.82) and moves with the storm
        pulse, so the 32 lanes of a warp -- which run in lockstep -- now mostly need the same number of trips.
        What an operational deployment would do once per network (the handle is cached across calls); the results do
        not depend on the order.  `calibration_qlat`: forcing of the calibration call ([n, columns]; None = window 0 of
        the router's own forcing) -- bench.py calibrates on a DIFFERENT storm than the one it times."""
        from ._lib import TrouteB200Error
        from .network import TRIP_BUCKETS
        buckets = TRIP_BUCKETS if buckets is None else int(buckets)
        self.net.collect_trips(buckets)
        if calibration_qlat is None:
            self.upload()
        else:
            self.net.upload(self.T, self.qts, np.ascontiguousarray(calibration_qlat), self.q0)
        self.net.run(self.short_ts)
        try:
            if buckets <= 1:
                raise TrouteB200Error("totals requested")
            key = self.net.trip_order_key(buckets)
            self.order_source = f"secant trip counts of {buckets} time slices of one calibration call"
        except TrouteB200Error:                      # the time-resolved table is an optimisation: the totals still order
            key = self.net.trip_counts()
            self.order_source = "secant trip counts of one calibration call"
        self.net.close()
        self._build_net(key)
        self.upload()
        self.reordered = True

    def run_resident(self):
        for w in range(self.windows):
            self._start_window(w, True)
            self.net.run_async(self.short_ts)

    def run_resident_incl_h2d(self):
        """the resident step with the forcing coming from pinned HOST memory (SURVEY.md 8d counts the qlat / q0 upload)"""
        for w in range(self.windows):
            self._start_window(w, False)
            self.net.run_async(self.short_ts)

    def run_e2e(self):
        _, _, out = self._host
        for w in range(self.windows):
            self._start_window(w, False)
            self.net.run_download_ptr(self.short_ts, out.data_ptr())

    def run_checked(self, on_window):
        """run_resident with a stop after every window: on_window(w) sees the complete device-resident result of window w"""
        for w in range(self.windows):
            self._start_window(w, True)
            self.net.run_async(self.short_ts)
            self.net.sync()
            on_window(w)

    def collect_stats(self):
        self.net.sync()
        st = self.net.last_run_stats()
        launches = st["launches"]
        lev = self.net.levels()
        wide_rows = int(((lev < st["first_marching_level"]) & (self.wl["kind"] != 2)).sum()) if self.mode == 4 else (
            self.n if self.mode < 3 else 0)
        return {"kernel_ms_per_call": st["kernel_ms"], "lane_steps": st["lane_steps"] * self.windows,
                "launches_per_call": launches * self.windows,
                "wide_ms": st["wide_ms"], "march_ms": st["march_ms"], "wide_lane_steps": wide_rows * self.T,
                "first_marching_level": st["first_marching_level"],
                "launches_per_call_e2e": launches * self.windows, "stages": st["stages"],
                "levels": self.net.num_levels, "kernel_name": self.kernel_names[self.mode],
                "sharding": "single GPU"}

    def close(self):
        self.net.close()


class ShardedRouter(_RouterBase):
    """One sub-basin shard per rank.  Cut-edge flows travel through CUDA-IPC-mapped peer memory inside the routing
    kernel (no collective on the data path); torch.distributed is used for rendezvous (IPC handles, import positions)
    and for the barrier between resetting the flow state and launching."""

    def __init__(self, wl, world, rank, device, nsteps, qts, short_ts, mode=4, pieces_per_shard=16, deep_lanes=8192, windows=1):
        import torch
        import torch.distributed as dist
        from . import hostgraph, partition
        if mode not in (2, 3, 4):
            raise ValueError("sharded routing needs a polling schedule (mode 2, 3 or 4)")
        self.torch, self.dist = torch, dist
        self.wl, self.world, self.rank, self.device = wl, world, rank, device
        self.T, self.qts, self.short_ts, self.mode = nsteps, qts, short_ts, mode
        self.windows = int(windows)
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
        self.q0 = np.ascontiguousarray(_q0_with_step0_observations(wl)[plan.rows])
        self.nq = self.qlat.shape[1]
        self.h2d_bytes = self.qlat.nbytes + self.q0.nbytes
        self.d2h_bytes = self.n * 3 * nsteps * 4 * self.windows
        self._host = None
        self._qlat_dev = None
        self._wired = False
        self.own_local_rows = np.nonzero(plan.own)[0].astype(np.int64)
        self.own_global_rows = plan.rows[plan.own].astype(np.int64)

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
        self.net.set_option("host_shards", max(1, min(64, self.world)))   # the GPUs of one host share its copy bandwidth
        for k, v in self.options.items():
            self.net.set_option(k, v)
        self.net.set_imports(plan.imports)
        self._set_gages(self.local_of_global)          # every shard assimilates the gages on the segments it owns
        self._wired = False

    def set_option(self, key, value):
        """Engine option that survives a rebuild of the shard's network."""
        self.options[key] = int(value)
        self.net.set_option(key, int(value))

    def local_of_global(self, rows):
        """local row of every global row this shard OWNS, -1 elsewhere"""
        plan = self.plan
        i = np.searchsorted(plan.rows, rows)
        i = np.minimum(i, plan.rows.size - 1)
        hit = (plan.rows[i] == rows) & plan.own[i]
        return np.where(hit, i, -1).astype(np.int64)

    def reorder_by_trip_history(self, buckets=None, calibration_qlat=None):
        """SingleRouter.reorder_by_trip_history for a sharded run: every rank records the trips of its own segments in one
        calibration call, rebuilds its network in that order and the shards are wired again (positions and flow arrays are
        new)."""
        from .network import TRIP_BUCKETS
        buckets = TRIP_BUCKETS if buckets is None else int(buckets)
        self.net.collect_trips(buckets)
        if calibration_qlat is None:
            self.upload()
        else:
            self.net.upload(self.T, self.qts, np.ascontiguousarray(calibration_qlat[self.plan.rows]), self.q0)
            if not self._wired:
                self._wire()
        self.net.prepare()
        self.dist.barrier()
        self.net.run_async(self.short_ts)
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
        self.net.upload(self.T, self.qts, np.ascontiguousarray(self.qlat[:, _window_cols(self.T, self.qts, 0)]), self.q0)
        if not self._wired:
            self._wire()

    def _run_windows(self, resident, download, on_window=None):
        for w in range(self.windows):
            self._start_window(w, resident)
            self.net.prepare()             # flow state back to "not yet written"; complete before any peer may write
            self.dist.barrier()
            if download:
                self.net.run_download_ptr(self.short_ts, self._host[2].data_ptr())
            else:
                self.net.run_async(self.short_ts)
            if on_window is not None:
                self.net.sync()
                on_window(w)
            if w + 1 < self.windows:
                # a peer may still be writing its last values into this shard's flow array: the next window's reset must
                # not start before every shard has finished this one
                self.net.sync()
                self.dist.barrier()

    def run_resident(self):
        self._run_windows(True, False)

    def run_resident_incl_h2d(self):
        self._run_windows(False, False)

    def run_e2e(self):
        self._run_windows(False, True)

    def run_checked(self, on_window):
        self._run_windows(True, False, on_window)

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
        return {"kernel_ms_per_call": st["kernel_ms"], "lane_steps": st["lane_steps"] * self.windows,
                "launches_per_call": st["launches"] * self.windows,
                "wide_ms": st["wide_ms"], "march_ms": st["march_ms"], "wide_lane_steps": wide_rows * self.T,
                "first_marching_level": st["first_marching_level"],
                "launches_per_call_e2e": st["launches"] * self.windows, "stages": st["stages"], "levels": self.net.num_levels,
                "kernel_name": self.kernel_names[self.mode],
                "sharding": f"{self.world} sub-basin shards, {self.plan_stats['n_cut_edges']} cut edges, "
                            f"imbalance {self.plan_stats['imbalance']:.3f}, peer-memory stores (no collective)"}

    def close(self):
        self.net.sync()
        self.dist.barrier()
        self.net.close_peers()
        self.net.close()
