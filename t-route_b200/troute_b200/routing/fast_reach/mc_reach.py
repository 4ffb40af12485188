"""GPU twin of troute.routing.fast_reach.mc_reach
(/root/reference/src/troute-routing/troute/routing/fast_reach/mc_reach.pyx).

`compute_network_structured` keeps the reference's 51 positional + 5 keyword arguments (:164-224) and returns
the reference's 10-tuple (:811-845), so it can be registered in troute.routing.compute._compute_func_map
(compute.py:21-26) as a drop-in backend.  What happens underneath is different: instead of
building one Python/C object per reach and segment on every call (:287-378) and looping time-outer /
reach-inner on one core (:492-800), the reach lists are flattened ONCE into CSR arrays, a device-resident
`RoutingNetwork` is built and cached, and every call only moves forcing in and results out.

Scope: Muskingum-Cunge reaches, plain level-pool reservoirs, prescribed upstream boundary
series (`upstream_results`).  Hybrid (USGS/USACE persistence), RFC-forecast and Great-Lakes reservoir data
assimilation are host-side, observation-file driven Python in the reference (reservoir_*_da.py) and are out of
scope: non-empty inputs for them raise NotImplementedError rather than being silently ignored.
"""
import hashlib
import zlib
from collections import OrderedDict
from itertools import chain, repeat
from operator import itemgetter

import numpy as np

from ..._lib import TrouteB200Error
from ...network import RoutingNetwork, TRT_KIND_BOUNDARY, TRT_KIND_LEVELPOOL, TRT_KIND_MC


def binary_find(arr, els):
    """mc_reach.pyx:36-66: positions of `els` in the sorted array `arr`; ValueError if an element is absent."""
    arr = np.asarray(arr)
    els = np.asarray(list(els) if not isinstance(els, np.ndarray) else els, dtype=np.int64)
    if els.size == 0:
        return np.zeros(0, dtype=np.int64)
    if arr.size == 0:
        raise ValueError(f"element {els[0]} not found in {arr}")
    idx = np.searchsorted(arr, els)
    safe = np.minimum(idx, arr.shape[0] - 1)
    bad = (idx >= arr.shape[0]) | (arr[safe] != els)
    if bad.any():
        raise ValueError(f"element {els[bad][0]} not found in {arr}")
    return idx.astype(np.int64)


def column_mapper(src_cols):
    """mc_reach.pyx:150-162."""
    index = {label: i for i, label in enumerate(src_cols)}
    return [index[label] for label in ["dt", "dx", "bw", "tw", "twcc", "n", "ncc", "cs", "s0"]]


# -------------------------------------------------------------------------------------------------
# flattening: (reaches_wTypes, upstream_connections, data_idx) -> CSR over rows
# -------------------------------------------------------------------------------------------------
def flatten_network(reaches_wTypes, upstream_connections, data_idx):
    """Rows = positions in the sorted index data_idx.  Inside a reach a segment's only upstream row is its
    predecessor (compute_reach_kernel, mc_reach.pyx:133-138); the head of a reach collects
    upstream_connections[reach[0]] in list order (:288-289, summed in that order :499-502).  Rows that belong to no
    reach (off-network upstream rows of the bmi / by-subnetwork modes) become TRT_KIND_BOUNDARY.

    Returns (up_ptr, up_rows, kind, seg_rows, reach_len, reach_type)."""
    n = int(data_idx.shape[0])
    nreach = len(reaches_wTypes)
    segs_of = list(map(itemgetter(0), reaches_wTypes))                         # iteration at C speed: 2.1 M reaches for CONUS
    reach_len = np.fromiter(map(len, segs_of), dtype=np.int64, count=nreach)
    reach_type = np.fromiter(map(itemgetter(1), reaches_wTypes), dtype=np.int64, count=nreach)
    total = int(reach_len.sum())
    seg_ids = np.fromiter(chain.from_iterable(segs_of), dtype=np.int64, count=total)
    seg_rows = binary_find(data_idx, seg_ids)                                  # ValueError on unknown ids (:65)
    if np.unique(seg_rows).size != total:
        raise ValueError("a segment appears in more than one reach")
    starts = np.zeros(nreach + 1, dtype=np.int64)
    np.cumsum(reach_len, out=starts[1:])
    head_rows = seg_rows[starts[:-1]] if nreach else np.zeros(0, np.int64)

    head_ups = list(map(upstream_connections.get, seg_ids[starts[:-1]].tolist(), repeat(())))
    head_cnt = np.fromiter(map(len, head_ups), dtype=np.int64, count=nreach)
    head_up_ids = np.fromiter(chain.from_iterable(head_ups), dtype=np.int64, count=int(head_cnt.sum()))
    head_up_rows = binary_find(data_idx, head_up_ids)

    indeg = np.zeros(n, dtype=np.int64)
    is_head = np.zeros(total, dtype=bool)
    is_head[starts[:-1]] = True
    indeg[seg_rows[~is_head]] = 1
    indeg[head_rows] = head_cnt
    up_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(indeg, out=up_ptr[1:])
    up_rows = np.empty(int(up_ptr[-1]), dtype=np.int64)
    # interior segments: predecessor in the reach
    inner = np.nonzero(~is_head)[0]
    up_rows[up_ptr[seg_rows[inner]]] = seg_rows[inner - 1]
    # heads: the upstream list, in order
    if head_up_rows.size:
        offs = np.zeros(nreach + 1, dtype=np.int64)
        np.cumsum(head_cnt, out=offs[1:])
        dst = np.repeat(up_ptr[head_rows], head_cnt) + (np.arange(head_up_rows.size) - np.repeat(offs[:-1], head_cnt))
        up_rows[dst] = head_up_rows

    kind = np.full(n, TRT_KIND_BOUNDARY, dtype=np.uint8)
    seg_type = np.repeat(reach_type, reach_len)
    kind[seg_rows[seg_type == 0]] = TRT_KIND_MC
    kind[seg_rows[seg_type == 1]] = TRT_KIND_LEVELPOOL
    if ((reach_type == 1) & (reach_len != 1)).any():
        raise ValueError("a reservoir reach must hold exactly one segment")      # "singleton list reaches" (:296)
    return up_ptr, up_rows, kind, seg_rows, reach_len, reach_type


# -------------------------------------------------------------------------------------------------
# device network cache (the analogue of the reference's subnetwork_list caching, compute.py:556,652-656)
# -------------------------------------------------------------------------------------------------
_NET_CACHE = OrderedDict()
_NET_CACHE_MAX = 4
DEFAULT_OPTIONS = {}          # engine options (trt_set_option) applied to every network built here, e.g. {"mode": 2}
# Networks of at least this many rows record the secant trip count of every segment during their FIRST call and are then
# rebuilt with the segments of each wavefront level ordered by it (RoutingNetwork(order_key=...)): the lanes of a warp run
# in lockstep and a segment's trip count repeats from step to step, so later calls run ~8 % faster.  None switches it off.
REORDER_MIN_ROWS = 200_000


def _fingerprint(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device):
    """Cache key of a device network: everything flatten_network reads -- the parameter table, the segment ids of every
    reach in order, the reach types and the upstream list of every reach head (the confluence wiring) -- so that two calls
    with the same parameters but different connectivity never share a device topology.  Iterates at C speed (map /
    itemgetter / fromiter): ~0.35 s per million segments."""
    h = hashlib.blake2b(digest_size=16)
    h.update(np.ascontiguousarray(data_idx).view(np.uint8))
    h.update(np.ascontiguousarray(data_values).view(np.uint8))
    h.update(repr([str(c) for c in data_cols]).encode())
    nreach = len(reaches_wTypes)
    h.update(repr((nreach, device)).encode())
    if nreach:
        segs_of = list(map(itemgetter(0), reaches_wTypes))
        lens = np.fromiter(map(len, segs_of), dtype=np.int64, count=nreach)
        types = np.fromiter(map(itemgetter(1), reaches_wTypes), dtype=np.int64, count=nreach)
        segs = np.fromiter(chain.from_iterable(segs_of), dtype=np.int64, count=int(lens.sum()))
        ups = list(map(upstream_connections.get, map(itemgetter(0), segs_of), repeat(())))
        up_cnt = np.fromiter(map(len, ups), dtype=np.int64, count=nreach)
        up_ids = np.fromiter(chain.from_iterable(ups), dtype=np.int64, count=int(up_cnt.sum()))
        for a in (lens, types, segs, up_cnt, up_ids):
            h.update(a.view(np.uint8))
    return h.hexdigest()


# The full fingerprint walks 2.1 M reaches of a CONUS network: seconds per call, for a routing call of 0.2 s.  A caller that
# passes THE SAME topology objects again -- what T-Route's loop does: the reach lists of a subnetwork are built once and kept
# (compute.py:556, :652-656) -- is recognised by a quick key instead: identity and length of the two topology containers, a
# sample of reaches spread over the list with their upstream lists, a CRC of the id and parameter arrays.  In-place edits of
# a cached topology between calls are therefore NOT seen unless they touch a sampled reach or change a length; callers that
# do such edits set VERIFY_TOPOLOGY_EVERY_CALL (or call clear_network_cache()).
VERIFY_TOPOLOGY_EVERY_CALL = False
_QUICK_SAMPLES = 256
_QUICK = {}          # quick key -> full fingerprint


def _quick_key(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device):
    nreach = len(reaches_wTypes)
    step = max(1, nreach // _QUICK_SAMPLES)
    sample = tuple((tuple(r), t, tuple(upstream_connections.get(r[0], ()))) for r, t in reaches_wTypes[::step])
    return (id(reaches_wTypes), nreach, id(upstream_connections), len(upstream_connections), hash(sample),
            zlib.crc32(np.ascontiguousarray(data_idx).view(np.uint8)), zlib.crc32(np.ascontiguousarray(data_values).view(np.uint8)),
            tuple(str(c) for c in data_cols), device)


def _network_key(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device):
    if VERIFY_TOPOLOGY_EVERY_CALL or not isinstance(reaches_wTypes, list):
        return _fingerprint(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device)
    qk = _quick_key(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device)
    full = _QUICK.get(qk)
    if full is None or full not in _NET_CACHE:
        full = _fingerprint(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device)
        if len(_QUICK) > 64:
            _QUICK.clear()
        _QUICK[qk] = full
    return full


def clear_network_cache():
    _QUICK.clear()
    if _RESULT_POOL is not None:
        _RESULT_POOL.clear()
    while _NET_CACHE:
        _, entry = _NET_CACHE.popitem()
        entry["net"].close()


def _get_network(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device):
    key = _network_key(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device)
    entry = _NET_CACHE.get(key)
    if entry is not None:
        _NET_CACHE.move_to_end(key)
        return entry
    up_ptr, up_rows, kind, seg_rows, reach_len, reach_type = flatten_network(reaches_wTypes, upstream_connections,
                                                                            data_idx)
    # level-pool rows carry NaN channel parameters in param_df_sub (compute.py:1458-1460); the engine ignores them
    vals = np.nan_to_num(np.asarray(data_values, dtype=np.float32), nan=0.0)
    cols = [str(c) for c in data_cols]
    net = RoutingNetwork(up_ptr, up_rows, kind, vals, cols, device=device)
    for k, v in DEFAULT_OPTIONS.items():
        net.set_option(k, v)
    entry = dict(net=net, kind=kind, seg_rows=seg_rows, reach_len=reach_len, reach_type=reach_type, ordered=True)
    if REORDER_MIN_ROWS is not None and kind.shape[0] >= REORDER_MIN_ROWS:
        net.collect_trips()
        entry.update(ordered=False, flat=(up_ptr, up_rows, vals, cols, device))
    _NET_CACHE[key] = entry
    while len(_NET_CACHE) > _NET_CACHE_MAX:
        _, old = _NET_CACHE.popitem(last=False)
        old["net"].close()
    return entry


def _take_rows(a, mask):
    """a[mask] (:811-813 `flowveldepth[fill_index_mask]`) without the copy when every row is returned -- 9.4 GB for a CONUS day"""
    return a if mask.all() else a[mask]


# The result of a CONUS day is 9.4 GB.  Into a fresh pageable array it costs the page faults of the allocation and a staged
# copy (2.7 s per call measured); into page-locked memory the copies are DMA at PCIe rate (0.33 s per call).  Results are
# therefore handed out from a pool of pinned blocks that take an array BACK only when its last reference is gone
# (network.PinnedPool): a loop that consumes one window before it routes the next keeps re-using one block, a caller that
# keeps old results makes the pool pin another one, and beyond RESULT_POOL_BYTES the mirror falls back to pageable arrays.
# Nothing a caller holds is ever overwritten.  0 switches the pool off.
RESULT_POOL_BYTES = 32 << 30
_RESULT_POOL = None


def _result_buffer(n, nsteps):
    global _RESULT_POOL
    if RESULT_POOL_BYTES <= 0:
        return None
    if _RESULT_POOL is None:
        from ...network import PinnedPool
        _RESULT_POOL = PinnedPool()
    return _RESULT_POOL.take((int(n), 3 * int(nsteps)), np.float32, RESULT_POOL_BYTES)


def _empty(a):
    return a is None or np.asarray(a).size == 0


def compute_network_structured(
    nsteps, dt, qts_subdivisions, reaches_wTypes, upstream_connections, data_idx, data_cols, data_values,
    initial_conditions, qlat_values, lake_numbers_col, wbody_cols, data_assimilation_parameters, reservoir_types,
    reservoir_type_specified, model_start_time, usgs_values, usgs_positions, usgs_positions_reach,
    usgs_positions_gage, lastobs_values_init, time_since_lastobs_init, da_decay_coefficient,
    reservoir_usgs_obs, reservoir_usgs_wbody_idx, reservoir_usgs_time, reservoir_usgs_update_time,
    reservoir_usgs_prev_persisted_flow, reservoir_usgs_persistence_update_time, reservoir_usgs_persistence_index,
    reservoir_usace_obs, reservoir_usace_wbody_idx, reservoir_usace_time, reservoir_usace_update_time,
    reservoir_usace_prev_persisted_flow, reservoir_usace_persistence_update_time, reservoir_usace_persistence_index,
    reservoir_rfc_obs, reservoir_rfc_wbody_idx, reservoir_rfc_totalCounts, reservoir_rfc_file,
    reservoir_rfc_use_forecast, reservoir_rfc_timeseries_idx, reservoir_rfc_update_time, reservoir_rfc_da_timestep,
    reservoir_rfc_persist_days, great_lakes_idx, great_lakes_times, great_lakes_discharge, great_lakes_param_idx,
    great_lakes_param_prev_assim_flow, great_lakes_param_prev_assim_times, great_lakes_param_update_times,
    great_lakes_climatology, upstream_results={}, assume_short_ts=False, return_courant=False, da_check_gage=-1,
    from_files=True, device=0,
):
    """Route a (sub)network for `nsteps` timesteps on the GPU.  Arguments and return value: mc_reach.pyx:164-224,
    :811-845.  `device` (extra keyword) selects the CUDA device."""
    data_idx = np.ascontiguousarray(data_idx, dtype=np.int64)
    data_values = np.asarray(data_values, dtype=np.float32)
    initial_conditions = np.asarray(initial_conditions, dtype=np.float32)
    qlat_values = np.asarray(qlat_values, dtype=np.float32)
    n = int(data_idx.shape[0])

    # shape checks, same messages (:243-250)
    if qlat_values.shape[0] != n:
        raise ValueError(f"Number of rows in Qlat is incorrect: expected ({n}), got ({qlat_values.shape[0]})")
    if qlat_values.shape[1] < nsteps / qts_subdivisions:
        raise ValueError(
            f"Number of columns (timesteps) in Qlat is incorrect: expected at most ({n}), got "
            f"({qlat_values.shape[1]}). The number of columns in Qlat must be equal to or less than the number of "
            f"routing timesteps")
    if data_values.shape[0] != n or data_values.shape[1] != len(data_cols):
        raise ValueError("data_values shape mismatch")

    # out-of-scope reservoir data assimilation must not be silently dropped
    if not (_empty(reservoir_usgs_wbody_idx) and _empty(reservoir_usace_wbody_idx) and _empty(reservoir_rfc_wbody_idx)
            and _empty(great_lakes_idx) and _empty(great_lakes_param_idx)):
        raise NotImplementedError("hybrid / RFC / Great-Lakes reservoir data assimilation is outside the GPU path")
    lake_numbers_col = list(lake_numbers_col)
    wbody = np.asarray(wbody_cols, dtype=np.float64).reshape(-1, 11) if len(lake_numbers_col) else np.zeros((0, 11))
    if reservoir_type_specified and len(lake_numbers_col):
        rt = np.asarray(reservoir_types).reshape(len(lake_numbers_col), -1)[:, 0]
        if ((rt == 4) | (rt == 5)).any() and from_files:
            raise NotImplementedError("RFC forecast reservoirs (types 4, 5) are outside the GPU path")

    entry = _get_network(reaches_wTypes, upstream_connections, data_idx, data_cols, data_values, device)
    net, kind = entry["net"], entry["kind"]

    # level pools: every call passes the table again (the elevation state lives in column h0)  (:291-305)
    lp_rows = np.nonzero(kind == TRT_KIND_LEVELPOOL)[0]
    if lp_rows.size:
        lake_arr = np.asarray(lake_numbers_col, dtype=np.int64)
        if lake_arr.size and (np.diff(lake_arr) < 0).any():
            order = np.argsort(lake_arr, kind="stable")
            wb_idx = order[binary_find(lake_arr[order], data_idx[lp_rows])]
        else:
            wb_idx = binary_find(lake_arr, data_idx[lp_rows])                  # wbody_index (:294)
        net.set_levelpools(lp_rows, wbody[wb_idx], routing_period=dt)
    else:
        net.set_levelpools(np.zeros(0, np.int64), np.zeros((0, 11)))

    # prescribed rows: upstream_results (:458-469) + rows that belong to no reach (they stay zero, :253)
    q0 = initial_conditions
    bnd_rows = np.nonzero(kind == TRT_KIND_BOUNDARY)[0]
    fill_index_mask = np.ones(n, dtype=bool)
    bnd_fvd = None
    if bnd_rows.size:
        q0 = np.array(initial_conditions, dtype=np.float32, copy=True)
        q0[bnd_rows] = 0.0
        bnd_fvd = np.zeros((bnd_rows.size, 3 * nsteps), dtype=np.float32)
        slot = {int(r): i for i, r in enumerate(bnd_rows)}
        lake_set = set(lake_numbers_col)
        for upstream_tw_id, tmp in upstream_results.items():
            fill_index = int(tmp["position_index"])
            if fill_index not in slot:
                raise ValueError(f"upstream_results row {fill_index} is also routed by a reach")
            fill_index_mask[fill_index] = False
            bnd_fvd[slot[fill_index]] = np.asarray(tmp["results"], dtype=np.float32).reshape(-1)[: 3 * nsteps]
            if int(data_idx[fill_index]) in lake_set:
                res_idx = lake_numbers_col.index(int(data_idx[fill_index]))
                q0[fill_index, 0] = wbody[res_idx, 9]
            else:
                q0[fill_index, 0] = initial_conditions[fill_index, 0]
                q0[fill_index, 2] = initial_conditions[fill_index, 2]
    elif upstream_results:
        raise ValueError("upstream_results rows must not be part of a reach")

    usgs_positions = np.asarray(usgs_positions, dtype=np.int32)
    gages = None
    if usgs_positions.size:
        gages = dict(usgs_values=np.asarray(usgs_values, dtype=np.float32),
                     usgs_positions=usgs_positions,
                     usgs_positions_reach=np.asarray(usgs_positions_reach, dtype=np.int32),
                     usgs_positions_gage=np.asarray(usgs_positions_gage, dtype=np.int32),
                     lastobs_values_init=np.asarray(lastobs_values_init, dtype=np.float32),
                     time_since_lastobs_init=np.asarray(time_since_lastobs_init, dtype=np.float32),
                     da_decay_coefficient=float(da_decay_coefficient),
                     reach_len=entry["reach_len"], seg_rows=entry["seg_rows"])
    placeholders = net.set_gages(gages, nsteps, routing_period=dt)
    if gages is None:
        nudge, lastobs_times, lastobs_values = placeholders

    # One C call (trt_route): the result columns of a time chunk travel home while the next chunk is routed.  The reservoir
    # inflows come back compact ([n_lp, nsteps]) and are scattered into the reference's upstream_array (zero everywhere else,
    # :807-813) on the host: the table of zeros never crosses PCIe.
    fvd, _ = net.route_call(nsteps, qts_subdivisions, qlat_values, q0, assume_short_ts=bool(assume_short_ts),
                            bnd_rows=bnd_rows if bnd_rows.size else None, bnd_fvd=bnd_fvd, out=_result_buffer(n, nsteps))
    upstream = np.zeros((n, int(nsteps)), dtype=np.float32)
    if lp_rows.size:
        upstream[lp_rows] = net.download_levelpool_inflow(lp_rows.size)
    if gages is not None:
        nudge, lastobs_times, lastobs_values = net.download_gages()
    if not entry["ordered"]:
        # first call on this network: rebuild it with every level ordered by the trip counts just collected
        up_ptr_f, up_rows_f, vals_f, cols_f, dev_f = entry.pop("flat")
        entry["ordered"] = True
        try:
            order_key = net.trip_order_key()
        except TrouteB200Error:            # the re-ordering is an optimisation: without the counters the network stays as it is
            order_key = None
            net.set_option("collect_trips", 0)
        if order_key is not None:
            ordered = RoutingNetwork(up_ptr_f, up_rows_f, kind, vals_f, cols_f, device=dev_f, order_key=order_key)
            for k, v in DEFAULT_OPTIONS.items():
                ordered.set_option(k, v)
            net.close()
            entry["net"] = ordered

    empty_f = np.zeros(0, dtype=np.float32)
    empty_i = np.zeros(0, dtype=np.int32)
    shift = np.float32(nsteps * dt)                                            # (timestep-1)*dt  (:822-836)
    return (
        _take_rows(data_idx.astype(np.intp), fill_index_mask),
        _take_rows(fvd, fill_index_mask),
        0,
        (np.asarray([data_idx[p] for p in usgs_positions]), lastobs_times, lastobs_values),
        (np.asarray(reservoir_usgs_wbody_idx, dtype=np.int32).reshape(-1), empty_f - shift, empty_f, empty_f,
         empty_f - shift),
        (np.asarray(reservoir_usace_wbody_idx, dtype=np.int32).reshape(-1), empty_f - shift, empty_f, empty_f,
         empty_f - shift),
        _take_rows(upstream, fill_index_mask),
        (np.asarray(reservoir_rfc_wbody_idx, dtype=np.int32).reshape(-1), empty_f - shift, empty_i),
        nudge,
        (empty_i, empty_f, empty_i, empty_i),
    )
