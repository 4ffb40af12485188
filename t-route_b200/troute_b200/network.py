"""RoutingNetwork -- flat-array host interface of the B200 routing engine.

A `RoutingNetwork` is the device-resident twin of what compute_network_structured builds per call out
of Python objects (/root/reference/src/troute-routing/troute/routing/fast_reach/mc_reach.pyx:287-378,
:472-481): the reach / segment structs, their upstream id arrays and parameter rows.  It is built
once from CSR arrays in the caller's row order (rows = positions in the sorted segment index
`data_idx`) and then routes any number of forcing chunks.

Only numpy and ctypes are used here; all arithmetic happens in libtroute_b200.so.
"""
import ctypes as C
import os

import threading
import weakref

import numpy as np

from . import _lib
from ._lib import TRT_KIND_BOUNDARY, TRT_KIND_LEVELPOOL, TRT_KIND_MC, as_c, check, ptr

# order of the nine per-segment parameters the kernel consumes (column_mapper, mc_reach.pyx:150-162)
PARAM_COLUMNS = ["dt", "dx", "bw", "tw", "twcc", "n", "ncc", "cs", "s0"]


def column_mapper(src_cols):
    """Map source column labels to the order the kernel expects (mc_reach.pyx:150-162)."""
    index = {label: i for i, label in enumerate(src_cols)}
    return [index[label] for label in PARAM_COLUMNS]


# time slices the trip-count collection of a calibration call is resolved into (trt_trip_counts_bucketed)
TRIP_BUCKETS = 16


OVERBANK_BINS = 16


def _expensive_first():
    # A/B switch (TRT_TRIP_ORDER=asc restores the ascending order of round 1); read at call time
    return os.environ.get("TRT_TRIP_ORDER", "desc") != "asc"


def order_key_from_trips(trips, nsteps=None, quantum=0.5, overbank=None, expensive_first=None):
    """Within-level sort key (one int32 per row) from the secant trip counts of a calibration call.

    The 32 lanes of a dataflow warp run until the slowest of them has finished its secant solve, so a warp is full when its
    segments need the same number of trips AT THE SAME TIME.  The total over the call (a 1-D `trips`) already groups
    segments that are slow on average; a [buckets, n_rows] table additionally tells WHEN a segment is slow (the rising limb
    of the storm pulse, the recession, ...).  Key = lexicographic order of the mean trips per time slice, quantised to
    `quantum` trips, the slice with the largest spread between segments first, ties broken by the total.  Measured with
    the cost model of tools/trip_order_study.py on an NHD-like network of 100,000 segments x 288 steps: 27.9 of 32 lanes
    busy, against 25.9 for the total alone and 22.4 in caller row order.  Any order gives the same results
    (tests/test_gpu_parity.py::test_within_level_order_and_trip_counts).

    `overbank` (steps every row ended above bankfull depth, trt_overbank_counts) becomes the primary key, in OVERBANK_BINS
    classes: a flooded compound channel takes the other branch of the celerity (one more power) and a warp with both kinds
    of lanes executes both branches.  Same model, 60,000 x 288: 25 % of the warp-steps mixed instead of 55 %, busy lanes
    28.4 (28.0 without)."""
    expensive_first = _expensive_first() if expensive_first is None else expensive_first
    trips = np.asarray(trips)
    if trips.ndim == 1:
        return np.ascontiguousarray(-trips if expensive_first else trips, dtype=np.int32)
    B, n = trips.shape
    if nsteps:
        # steps of the call that fall into each slice: step t (1-based) -> (t - 1) * B // nsteps
        lens = np.bincount((np.arange(int(nsteps)) * B) // int(nsteps), minlength=B).astype(np.float64)
    else:
        lens = np.ones(B)
    q = np.rint(trips / np.maximum(lens, 1.0)[:, None] / float(quantum)).astype(np.int64)
    primary_first = np.argsort(-q.std(axis=1), kind="stable")
    keys = [trips.sum(axis=0, dtype=np.int64)] + [q[j] for j in primary_first[::-1]]      # lexsort: last key is primary
    if overbank is not None:
        ob = np.asarray(overbank, dtype=np.int64)
        span = int(nsteps) + 1 if nsteps else int(ob.max()) + 1
        keys.append(np.minimum(ob * OVERBANK_BINS // span, OVERBANK_BINS - 1))
    order = np.lexsort(tuple(keys))
    key = np.empty(n, dtype=np.int32)
    key[order] = np.arange(n, dtype=np.int32)
    # Units of a stage are claimed in position order: with the EXPENSIVE tiles first (longest processing time first) a
    # stage ends with its cheap tiles, so its completion -- which the run-ahead gate of later stages waits for -- is not
    # held up by a tile of 6-trip lanes that was claimed last.
    return (n - 1 - key).astype(np.int32) if expensive_first else key


class RoutingNetwork:
    """Handle on a levelled, device-resident river network.

    Parameters
    ----------
    up_ptr, up_rows : CSR of upstream rows (reference summation order, mc_reach.pyx:499-502)
    kind            : [n_rows] uint8, TRT_KIND_MC / TRT_KIND_LEVELPOOL / TRT_KIND_BOUNDARY
    data_values     : [n_rows, ncols] float32 parameter table
    data_cols       : column labels of data_values (must contain PARAM_COLUMNS)
    device          : CUDA device ordinal
    """

    def __init__(self, up_ptr, up_rows, kind, data_values, data_cols, device=0, levels=None, order_key=None):
        L = _lib.lib()
        self._L = L
        self._h = C.c_void_p()
        up_ptr = as_c(up_ptr, np.int64)
        up_rows = as_c(up_rows, np.int64)
        kind = as_c(kind, np.uint8)
        data_values = as_c(data_values, np.float32)
        n_rows = int(kind.shape[0])
        if up_ptr.shape[0] != n_rows + 1:
            raise ValueError(f"up_ptr must have n_rows+1 = {n_rows + 1} entries, got {up_ptr.shape[0]}")
        if data_values.ndim != 2 or data_values.shape[0] != n_rows or data_values.shape[1] != len(data_cols):
            raise ValueError("data_values shape mismatch")            # mc_reach.pyx:249-250
        if n_rows and int(up_ptr[-1]) != up_rows.shape[0]:
            raise ValueError("up_ptr[-1] must equal len(up_rows)")
        scols = np.asarray(column_mapper(list(data_cols)), dtype=np.int32)
        self.n_rows = n_rows
        self.device = int(device)
        self.kind = kind
        self._keep = (up_ptr, up_rows, kind, data_values, scols)
        lv = None
        if levels is not None:
            lv = as_c(levels, np.int32)
            if lv.shape[0] != n_rows:
                raise ValueError("levels must have one entry per row")
        ok = None
        if order_key is not None:
            # within-level order of the segments in the engine (any order gives the same results), e.g. trip_counts()
            ok = as_c(order_key, np.int32)
            if ok.shape[0] != n_rows:
                raise ValueError("order_key must have one entry per row")
        check(L.trt_network_create_ordered(self.device, n_rows, ptr(up_ptr, C.c_int64), ptr(up_rows, C.c_int64),
                                           ptr(kind, C.c_uint8), ptr(data_values, C.c_float), int(data_values.shape[1]),
                                           ptr(scols, C.c_int32), ptr(lv, C.c_int32) if lv is not None else None,
                                           ptr(ok, C.c_int32) if ok is not None else None, C.byref(self._h)))
        self._keep = None
        self.nsteps = 0
        self._lp_rows = np.zeros(0, dtype=np.int64)
        # TRT_OPTIONS="key=value,...": engine options applied to every network of the process -- lets a whole test or
        # bench run go through an alternative schedule knob (e.g. TRT_OPTIONS=march_group=4 pytest -m gpu)
        for kv in filter(None, os.environ.get("TRT_OPTIONS", "").split(",")):
            k, _, v = kv.partition("=")
            self.set_option(k.strip(), int(v))

    # -- lifetime -----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            self._L.trt_network_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- topology -----------------------------------------------------------------------------
    @property
    def num_levels(self):
        out = C.c_int32()
        check(self._L.trt_network_num_levels(self._h, C.byref(out)))
        return out.value

    def levels(self):
        out = np.empty(self.n_rows, dtype=np.int32)
        check(self._L.trt_network_get_levels(self._h, ptr(out, C.c_int32)))
        return out

    def positions(self):
        out = np.empty(self.n_rows, dtype=np.int32)
        check(self._L.trt_network_get_positions(self._h, ptr(out, C.c_int32)))
        return out

    # -- configuration ------------------------------------------------------------------------
    def set_levelpools(self, lp_rows, wbody_cols, routing_period=300.0):
        """wbody_cols: [n_lp, 11] float64 rows (LkArea, LkMxE, OrificeA, OrificeC, OrificeE, WeirC, WeirE,
        WeirL, ifd, qd0, h0) -- compute.py:1416-1430."""
        lp_rows = as_c(lp_rows, np.int64)
        wbody_cols = as_c(wbody_cols, np.float64).reshape(-1, 11) if len(lp_rows) else np.zeros((0, 11))
        if wbody_cols.shape[0] != lp_rows.shape[0]:
            raise ValueError("one wbody_cols row per level-pool row is required")
        check(self._L.trt_network_set_levelpools(self._h, int(lp_rows.shape[0]), ptr(lp_rows, C.c_int64),
                                                 ptr(wbody_cols, C.c_double), float(routing_period)))
        self._lp_rows = lp_rows

    def set_option(self, key, value):
        check(self._L.trt_set_option(self._h, key.encode(), int(value)))

    def set_gages(self, gages, nsteps, routing_period=300.0):
        """Streamflow nudging set-up (mc_reach.pyx:380-411).  `gages` holds the reference's arguments (usgs_values,
        usgs_positions, usgs_positions_reach, usgs_positions_gage, lastobs_values_init, time_since_lastobs_init,
        da_decay_coefficient) plus the reach structure (seg_rows, reach_len); None switches nudging off.
        Returns (nudge, lastobs_times, lastobs_values) placeholders for a gage-free call."""
        self._n_gages = 0
        if gages is None or len(gages["usgs_positions"]) == 0:
            check(self._L.trt_network_set_gages(self._h, 0, None, None, None, 0, None, None, 0.0, float(routing_period)))
            return (np.zeros((0, nsteps + 1), dtype=np.float32), np.zeros(0, dtype=np.float32),
                    np.zeros(0, dtype=np.float32))
        pos = as_c(gages["usgs_positions"], np.int64)
        G = int(pos.shape[0])
        usgs = as_c(gages["usgs_values"], np.float32)
        usgs = usgs.reshape(G, -1) if usgs.size else np.zeros((G, 0), dtype=np.float32)
        # reach_has_gage[usgs_positions_reach[i]] = usgs_positions_gage[i]: the last gage listed for a reach wins (:398)
        reach_gage = {}
        for r, g in zip(np.asarray(gages["usgs_positions_reach"]).tolist(), np.asarray(gages["usgs_positions_gage"]).tolist()):
            reach_gage[int(r)] = int(g)
        active = np.zeros(G, dtype=np.uint8)
        starts = np.concatenate([[0], np.cumsum(gages["reach_len"])])
        for r, g in reach_gage.items():
            if g < 0:
                continue
            if not (0 <= g < G and 0 <= r < len(starts) - 1):
                raise ValueError(f"gage {g} / reach {r} out of range")
            last = int(gages["seg_rows"][starts[r + 1] - 1])
            if int(pos[g]) != last:
                raise NotImplementedError(
                    f"gage at row {int(pos[g])} is not the last segment of reach {r}: reaches must be broken at gages "
                    f"(nhd_network.split_at_gages_waterbodies_and_junctions)")
            active[g] = 1
        lv = as_c(gages["lastobs_values_init"], np.float32)
        lt = as_c(gages["time_since_lastobs_init"], np.float32)
        if lv.shape[0] != G or lt.shape[0] != G:
            raise ValueError("one last-observation value and time per gage is required")
        check(self._L.trt_network_set_gages(self._h, G, ptr(pos, C.c_int64), ptr(active, C.c_uint8),
                                            ptr(usgs, C.c_float) if usgs.size else None, int(usgs.shape[1]),
                                            ptr(lv, C.c_float), ptr(lt, C.c_float),
                                            float(gages["da_decay_coefficient"]), float(routing_period)))
        self._n_gages = G
        return None

    def download_gages(self):
        """(nudge [n_gages, nsteps + 1], lastobs_times, lastobs_values) of the last run."""
        G = self._n_gages
        nudge = np.zeros((G, self.nsteps + 1), dtype=np.float32)
        t = np.zeros(G, dtype=np.float32)
        v = np.zeros(G, dtype=np.float32)
        if G:
            check(self._L.trt_download_gages(self._h, ptr(nudge, C.c_float), ptr(t, C.c_float), ptr(v, C.c_float)))
        return nudge, t, v

    # -- routing ------------------------------------------------------------------------------
    def _check_forcing(self, nsteps, qts_subdivisions, qlat, q0):
        if qlat.ndim != 2 or qlat.shape[0] != self.n_rows:
            raise ValueError(
                f"Number of rows in Qlat is incorrect: expected ({self.n_rows}), got ({qlat.shape[0]})")  # :243-244
        if q0.shape != (self.n_rows, 3):
            raise ValueError(f"initial_conditions must be ({self.n_rows}, 3), got {q0.shape}")

    def upload(self, nsteps, qts_subdivisions, qlat, q0, bnd_rows=None, bnd_fvd=None):
        """Host -> device: qlat [n_rows, nqcols], q0 [n_rows, 3], optional prescribed rows."""
        qlat = as_c(qlat, np.float32)
        q0 = as_c(q0, np.float32)
        self._check_forcing(nsteps, qts_subdivisions, qlat, q0)
        n_bnd = 0
        rows_p = None
        fvd_p = None
        if bnd_rows is not None and len(bnd_rows):
            bnd_rows = as_c(bnd_rows, np.int64)
            bnd_fvd = as_c(bnd_fvd, np.float32)
            if bnd_fvd.shape != (bnd_rows.shape[0], 3 * nsteps):
                raise ValueError("bnd_fvd must be [n_bnd, 3*nsteps]")
            n_bnd = int(bnd_rows.shape[0])
            rows_p = ptr(bnd_rows, C.c_int64)
            fvd_p = bnd_fvd.ctypes.data
        check(self._L.trt_upload_forcing(self._h, int(nsteps), int(qts_subdivisions), qlat.ctypes.data,
                                         int(qlat.shape[1]), q0.ctypes.data, n_bnd, rows_p, fvd_p))
        self.nsteps = int(nsteps)

    def upload_ptr(self, nsteps, qts_subdivisions, qlat_ptr, nqcols, q0_ptr):
        """Same as upload() for raw host addresses (e.g. pinned torch tensors' data_ptr())."""
        check(self._L.trt_upload_forcing(self._h, int(nsteps), int(qts_subdivisions), qlat_ptr, int(nqcols), q0_ptr,
                                         0, None, None))
        self.nsteps = int(nsteps)

    def continue_window(self, nsteps, qts_subdivisions, qlat, usgs_values=None, bnd_rows=None, bnd_fvd=None):
        """The next routing window, started from the device-resident state of the finished one (trt_continue): no q0, no
        reservoir table, no last-observation table cross PCIe.  `usgs_values` [n_gages, columns]: the observation table of
        the new window (None keeps the current one)."""
        qlat = as_c(qlat, np.float32)
        if qlat.ndim != 2 or qlat.shape[0] != self.n_rows:
            raise ValueError(f"Number of rows in Qlat is incorrect: expected ({self.n_rows}), got ({qlat.shape[0]})")
        if usgs_values is not None and getattr(self, "_n_gages", 0):
            usgs = as_c(usgs_values, np.float32).reshape(self._n_gages, -1)
            check(self._L.trt_network_update_gage_observations(self._h, usgs.ctypes.data, int(usgs.shape[1])))
        n_bnd, rows_p, fvd_p = 0, None, None
        if bnd_rows is not None and len(bnd_rows):
            bnd_rows = as_c(bnd_rows, np.int64); bnd_fvd = as_c(bnd_fvd, np.float32)
            n_bnd, rows_p, fvd_p = int(bnd_rows.shape[0]), ptr(bnd_rows, C.c_int64), bnd_fvd.ctypes.data
        check(self._L.trt_continue(self._h, int(nsteps), int(qts_subdivisions), qlat.ctypes.data, int(qlat.shape[1]),
                                   n_bnd, rows_p, fvd_p))
        self.nsteps = int(nsteps)

    def continue_ptr(self, nsteps, qts_subdivisions, qlat_ptr, nqcols):
        """continue_window for a raw host address (pinned buffers)."""
        check(self._L.trt_continue(self._h, int(nsteps), int(qts_subdivisions), qlat_ptr, int(nqcols), 0, None, None))
        self.nsteps = int(nsteps)

    def download_rows(self, rows):
        """Result rows `rows` of the last run -> [len(rows), 3 * nsteps] (trt_download_rows)."""
        rows = as_c(rows, np.int64)
        out = np.empty((rows.shape[0], 3 * self.nsteps), dtype=np.float32)
        check(self._L.trt_download_rows(self._h, int(rows.shape[0]), ptr(rows, C.c_int64), out.ctypes.data))
        return out

    def result_hash(self, rows=None, ids=None):
        """64-bit checksum of the device-resident result (trt_result_hash): sum over rows of hash(id, row bits)."""
        out = C.c_uint64()
        r = as_c(rows, np.int64) if rows is not None else None
        i = as_c(ids, np.int64) if ids is not None else None
        nsel = int(r.shape[0]) if r is not None else (int(i.shape[0]) if i is not None else 0)
        check(self._L.trt_result_hash(self._h, nsel, ptr(r, C.c_int64) if r is not None else None,
                                      ptr(i, C.c_int64) if i is not None else None, C.byref(out)))
        return int(out.value)

    def run(self, assume_short_ts=False):
        check(self._L.trt_run(self._h, 1 if assume_short_ts else 0))

    def run_async(self, assume_short_ts=False):
        check(self._L.trt_run_async(self._h, 1 if assume_short_ts else 0))

    def sync(self):
        check(self._L.trt_sync(self._h))

    def download(self, want_upstream=False, out=None):
        """Device -> host: ([n_rows, 3*nsteps] float32, [n_rows, nsteps] float32 | None)."""
        fvd = out if out is not None else np.empty((self.n_rows, 3 * self.nsteps), dtype=np.float32)
        up = np.empty((self.n_rows, self.nsteps), dtype=np.float32) if want_upstream else None
        check(self._L.trt_download_results(self._h, fvd.ctypes.data, up.ctypes.data if up is not None else None))
        return fvd, up

    def download_levelpool_inflow(self, n_lp):
        """Reservoir inflow of the level pools alone, [n_lp, nsteps] in the order of set_levelpools (trt_download_levelpool_inflow)."""
        out = np.zeros((int(n_lp), self.nsteps), dtype=np.float32)
        if out.size:
            check(self._L.trt_download_levelpool_inflow(self._h, out.ctypes.data))
        return out

    def download_last_step(self):
        """(q, v, d) of the last timestep of the last run, [n_rows, 3] in caller row order (trt_download_last_step)."""
        out = np.empty((self.n_rows, 3), dtype=np.float32)
        check(self._L.trt_download_last_step(self._h, out.ctypes.data))
        return out

    def run_download_ptr(self, assume_short_ts, fvd_ptr):
        """Time-chunked run with the result copies overlapped (trt_run_download) into a raw host address."""
        check(self._L.trt_run_download(self._h, 1 if assume_short_ts else 0, fvd_ptr, None))

    def download_ptr(self, fvd_ptr):
        check(self._L.trt_download_results(self._h, fvd_ptr, None))

    def route(self, nsteps, qts_subdivisions, qlat, q0, assume_short_ts=False, bnd_rows=None, bnd_fvd=None,
              want_upstream=False):
        """upload + run + download.  Returns (fvd [n_rows, 3*nsteps], upstream [n_rows, nsteps] | None)."""
        self.upload(nsteps, qts_subdivisions, qlat, q0, bnd_rows, bnd_fvd)
        self.run(assume_short_ts)
        return self.download(want_upstream)

    def route_call(self, nsteps, qts_subdivisions, qlat, q0, assume_short_ts=False, want_upstream=False, bnd_rows=None,
                   bnd_fvd=None, out=None):
        """The single-call C entry point trt_route on numpy arrays (time-chunked, copies overlapped with compute).
        `out`: a caller-owned [n_rows, 3*nsteps] float32 buffer for the result (e.g. pinned memory kept across calls)."""
        qlat = as_c(qlat, np.float32)
        q0 = as_c(q0, np.float32)
        self._check_forcing(nsteps, qts_subdivisions, qlat, q0)
        fvd = out if out is not None else np.empty((self.n_rows, 3 * int(nsteps)), dtype=np.float32)
        if fvd.shape != (self.n_rows, 3 * int(nsteps)) or fvd.dtype != np.float32 or not fvd.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float32 array of shape (n_rows, 3 * nsteps)")
        up = np.empty((self.n_rows, int(nsteps)), dtype=np.float32) if want_upstream else None
        n_bnd, rows_p, fvd_p = 0, None, None
        if bnd_rows is not None and len(bnd_rows):
            bnd_rows = as_c(bnd_rows, np.int64)
            bnd_fvd = as_c(bnd_fvd, np.float32)
            if bnd_fvd.shape != (bnd_rows.shape[0], 3 * int(nsteps)):
                raise ValueError("bnd_fvd must have shape (len(bnd_rows), 3 * nsteps)")
            n_bnd, rows_p, fvd_p = int(bnd_rows.shape[0]), ptr(bnd_rows, C.c_int64), bnd_fvd.ctypes.data
        check(self._L.trt_route(self._h, int(nsteps), int(qts_subdivisions), 1 if assume_short_ts else 0,
                                qlat.ctypes.data, int(qlat.shape[1]), q0.ctypes.data, n_bnd, rows_p, fvd_p, fvd.ctypes.data,
                                up.ctypes.data if up is not None else None))
        self.nsteps = int(nsteps)
        return fvd, up

    def route_ptr(self, nsteps, qts_subdivisions, assume_short_ts, qlat_ptr, nqcols, q0_ptr, fvd_ptr):
        """The single-call C entry point trt_route on raw host addresses (end-to-end timing path)."""
        check(self._L.trt_route(self._h, int(nsteps), int(qts_subdivisions), 1 if assume_short_ts else 0, qlat_ptr,
                                int(nqcols), q0_ptr, 0, None, None, fvd_ptr, None))
        self.nsteps = int(nsteps)

    # -- sharding (peer memory) ---------------------------------------------------------------
    def prepare(self):
        check(self._L.trt_prepare(self._h))

    def state_ptr(self):
        out = C.c_void_p()
        check(self._L.trt_network_state_ptr(self._h, C.byref(out)))
        return out.value

    def ipc_handle(self):
        """64-byte CUDA IPC handle of the flow array (to be opened by the peer process)."""
        buf = (C.c_uint8 * 64)()
        check(self._L.trt_ipc_get_handle(C.c_void_p(self.state_ptr()), buf))
        return bytes(buf)

    def open_peer(self, peer, handle_bytes, peer_n_rows):
        buf = (C.c_uint8 * 64).from_buffer_copy(handle_bytes)
        out = C.c_void_p()
        check(self._L.trt_ipc_open_handle(self.device, buf, C.byref(out)))
        check(self._L.trt_network_set_peer(self._h, int(peer), out, int(peer_n_rows)))
        self._peer_ptrs = getattr(self, "_peer_ptrs", []) + [out.value]
        return out.value

    def set_peer_ptr(self, peer, device_ptr, peer_n_rows):
        """Same-process peer (tests / single process driving several GPUs): pass the other handle's state_ptr()."""
        check(self._L.trt_network_set_peer(self._h, int(peer), C.c_void_p(device_ptr), int(peer_n_rows)))

    def close_peers(self):
        for p in getattr(self, "_peer_ptrs", []):
            self._L.trt_ipc_close_handle(C.c_void_p(p))
        self._peer_ptrs = []

    def set_exports(self, rows, peer, peer_pos):
        rows = as_c(rows, np.int64); peer = as_c(peer, np.int32); peer_pos = as_c(peer_pos, np.int64)
        check(self._L.trt_network_set_exports(self._h, int(rows.shape[0]), ptr(rows, C.c_int64), ptr(peer, C.c_int32),
                                              ptr(peer_pos, C.c_int64)))

    def set_imports(self, rows):
        rows = as_c(rows, np.int64)
        check(self._L.trt_network_set_imports(self._h, int(rows.shape[0]), ptr(rows, C.c_int64)))

    # -- device-side access -------------------------------------------------------------------
    def export_flow_series(self, rows, dst_device_ptr):
        rows = as_c(rows, np.int64)
        check(self._L.trt_export_flow_series(self._h, int(rows.shape[0]), ptr(rows, C.c_int64), dst_device_ptr))

    def import_boundary_flow(self, rows, src_device_ptr):
        rows = as_c(rows, np.int64)
        check(self._L.trt_import_boundary_flow(self._h, int(rows.shape[0]), ptr(rows, C.c_int64), src_device_ptr))

    def device_results_ptr(self):
        out = C.c_void_p()
        check(self._L.trt_device_results(self._h, C.byref(out)))
        return out.value

    def stage_profile(self):
        """(stage_ms, stage_width) of the last mode-0 run with option profile_stages=1; index = stage k."""
        cnt = C.c_int64()
        check(self._L.trt_stage_profile(self._h, 0, None, None, C.byref(cnt)))
        ms = np.zeros(cnt.value, dtype=np.float32)
        w = np.zeros(cnt.value, dtype=np.int64)
        check(self._L.trt_stage_profile(self._h, cnt.value, ptr(ms, C.c_float), ptr(w, C.c_int64), C.byref(cnt)))
        return ms, w

    def last_run_stats(self):
        ms = C.c_double()
        launches = C.c_int64()
        stages = C.c_int64()
        lane_steps = C.c_int64()
        check(self._L.trt_last_run_stats(self._h, C.byref(ms), C.byref(launches), C.byref(stages), C.byref(lane_steps)))
        wide, march, lvl = C.c_double(), C.c_double(), C.c_int32()
        check(self._L.trt_last_run_phases(self._h, C.byref(wide), C.byref(march), C.byref(lvl)))
        return {"kernel_ms": ms.value, "launches": launches.value, "stages": stages.value,
                "lane_steps": lane_steps.value, "wide_ms": wide.value, "march_ms": march.value,
                "first_marching_level": lvl.value}

    def trip_counts(self, buckets=None):
        """Secant trips of every row summed over the last run (option collect_trips = 1 must have been set before it).
        With `buckets` (== option trip_buckets of that run): [buckets, n_rows], the sums of equal time slices of the call."""
        if buckets is None:
            out = np.zeros(self.n_rows, dtype=np.int32)
            check(self._L.trt_trip_counts(self._h, ptr(out, C.c_int32)))
            return out
        out = np.zeros((int(buckets), self.n_rows), dtype=np.int32)
        check(self._L.trt_trip_counts_bucketed(self._h, int(buckets), ptr(out, C.c_int32)))
        return out

    def collect_trips(self, buckets=None):
        """Switch the trip-count collection on for the following runs (time-resolved into TRIP_BUCKETS slices by default)."""
        self.set_option("trip_buckets", TRIP_BUCKETS if buckets is None else int(buckets))
        self.set_option("collect_trips", 1)

    def trip_order_key(self, buckets=None):
        """order_key for a rebuilt network (RoutingNetwork(order_key=...)) from the trips the last run collected."""
        b = TRIP_BUCKETS if buckets is None else int(buckets)
        return order_key_from_trips(self.trip_counts(b), self.nsteps, overbank=self.overbank_counts())

    def overbank_counts(self):
        """Steps every row ended above its bankfull depth in the last collecting run (trt_overbank_counts)."""
        out = np.zeros(self.n_rows, dtype=np.int32)
        check(self._L.trt_overbank_counts(self._h, ptr(out, C.c_int32)))
        return out

    def march_profile(self):
        """[n_rows, 4] uint64 of the last run with option march_profile=1 (see trt_march_profile)."""
        rows = C.c_int64()
        check(self._L.trt_march_profile(self._h, 0, None, C.byref(rows)))
        out = np.zeros((rows.value, 4), dtype=np.uint64)
        if rows.value:
            check(self._L.trt_march_profile(self._h, rows.value, out.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(rows)))
        return out


# -------------------------------------------------------------------------------------------------
# known-answer entry points (GPU twins of the reference's python-callable kernels)
# -------------------------------------------------------------------------------------------------
def mc_segment_batch(in15, device=0, want_iters=False):
    """in15 [count, 15] rows (dt,qup,quc,qdp,ql,dx,bw,tw,twcc,n,ncc,cs,s0,velp,depthp) -> [count, 6]
    rows (qdc, velc, depthc, ck, cn, X)."""
    L = _lib.lib()
    in15 = as_c(in15, np.float32).reshape(-1, 15)
    out = np.empty((in15.shape[0], 6), dtype=np.float32)
    iters = np.empty(in15.shape[0], dtype=np.int32)
    check(L.trt_mc_segment_batch(int(device), int(in15.shape[0]), ptr(in15, C.c_float), ptr(out, C.c_float),
                                 ptr(iters, C.c_int32)))
    return (out, iters) if want_iters else out


def levelpool_series(wbody_row, inflow, lateral_inflow=0.0, routing_period=300.0, device=0):
    L = _lib.lib()
    wbody_row = as_c(wbody_row, np.float64).reshape(11)
    inflow = as_c(inflow, np.float32)
    q = np.empty_like(inflow)
    h = np.empty_like(inflow)
    check(L.trt_levelpool_series(int(device), ptr(wbody_row, C.c_double), int(inflow.shape[0]), ptr(inflow, C.c_float),
                                 float(lateral_inflow), float(routing_period), ptr(q, C.c_float), ptr(h, C.c_float)))
    return q, h


def powf_batch(x, y, device=0):
    L = _lib.lib()
    x = as_c(x, np.float32)
    y = as_c(y, np.float32)
    out = np.empty_like(x)
    check(L.trt_powf_batch(int(device), int(x.shape[0]), ptr(x, C.c_float), ptr(y, C.c_float), ptr(out, C.c_float)))
    return out


class _PinnedBlock:
    """owner of one trt_host_alloc block"""
    def __init__(self, nbytes):
        self._L = _lib.lib()
        self.ptr = C.c_void_p()
        check(self._L.trt_host_alloc(C.byref(self.ptr), int(nbytes)))
        self.nbytes = int(nbytes)

    def free(self):
        if self.ptr:
            self._L.trt_host_free(self.ptr)
            self.ptr = C.c_void_p()

    def __del__(self):
        try:
            self.free()
        except Exception:          # noqa: BLE001 -- interpreter shutdown
            pass


class PinnedPool:
    """numpy arrays in page-locked host memory (trt_host_alloc) that go BACK to the pool when the last reference to them --
    the array, any view or slice of it -- is dropped.  Pinning gigabytes takes seconds; a routing loop that consumes the result
    of one window before it routes the next (T-Route's does) keeps re-using one block, and a caller that holds on to old
    results simply makes the pool allocate another block (up to `limit_bytes`; beyond that `take` returns None and the caller
    falls back to pageable memory).  No array handed out is ever overwritten behind its owner's back."""
    _block_type = None           # set below: _PinnedBlock (tests substitute a host-memory stand-in)

    def __init__(self):
        self._free = {}          # nbytes -> [blocks]
        self._total = 0
        self._lock = threading.Lock()

    def take(self, shape, dtype, limit_bytes):
        dtype = np.dtype(dtype)
        count = int(np.prod(shape))
        nbytes = max(count * dtype.itemsize, 1)
        evicted = []
        with self._lock:
            blocks = self._free.get(nbytes)
            block = blocks.pop() if blocks else None
            if block is None:
                # over the limit: idle blocks of OTHER sizes go first (a caller whose result shape changed)
                for size in sorted(self._free, reverse=True):
                    while self._total + nbytes > limit_bytes and self._free[size]:
                        evicted.append(self._free[size].pop())
                        self._total -= size
                if self._total + nbytes > limit_bytes:
                    block = False
                else:
                    self._total += nbytes
        for b in evicted:
            b.free()
        if block is False:
            return None
        if block is None:
            try:
                block = self._block_type(nbytes)
            except Exception:      # noqa: BLE001 -- no pinned memory to be had: the caller uses pageable memory
                with self._lock:
                    self._total -= nbytes
                return None
        buf = (C.c_uint8 * nbytes).from_address(block.ptr.value)
        weakref.finalize(buf, self._give_back, block)
        return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

    def _give_back(self, block):
        with self._lock:
            self._free.setdefault(block.nbytes, []).append(block)

    def clear(self):
        """free the blocks nobody holds"""
        with self._lock:
            blocks = [b for bs in self._free.values() for b in bs]
            self._free = {}
            self._total -= sum(b.nbytes for b in blocks)
        for b in blocks:
            b.free()

    @property
    def pinned_bytes(self):
        return self._total


PinnedPool._block_type = _PinnedBlock


def pinned_empty(shape, dtype=np.float32):
    """numpy array in page-locked host memory of its own (freed when the array is collected)"""
    arr = PinnedPool().take(shape, dtype, 1 << 62)
    if arr is None:
        raise MemoryError("trt_host_alloc failed")
    return arr


def fdiv_batch(a, d, device=0):
    """McDivFast element by element on the device: (quotients, inside-the-window flags)"""
    L = _lib.lib()
    a = as_c(a, np.float32)
    d = as_c(d, np.float32)
    out = np.empty_like(a)
    inside = np.zeros(a.shape[0], dtype=np.uint8)
    check(L.trt_fdiv_batch(int(device), int(a.shape[0]), ptr(a, C.c_float), ptr(d, C.c_float), ptr(out, C.c_float),
                           ptr(inside, C.c_uint8)))
    return out, inside.astype(bool)


__all__ = ["RoutingNetwork", "PARAM_COLUMNS", "column_mapper", "mc_segment_batch", "levelpool_series", "powf_batch", "fdiv_batch", "pinned_empty", "PinnedPool",
           "TRT_KIND_MC", "TRT_KIND_LEVELPOOL", "TRT_KIND_BOUNDARY"]
