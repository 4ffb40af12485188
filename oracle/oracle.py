"""TEST INFRASTRUCTURE -- ctypes front end of the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and the cpu_baseline / --impl reference legs of bench.py may
import this module.  The product package (t-route_b200/troute_b200) never does.

`compute_network_structured` below takes the reference's own arguments
(/root/reference/src/troute-routing/troute/routing/fast_reach/mc_reach.pyx:164-224) and returns the
reference's 10-tuple (:811-845), so that parity tests read like calls into the reference.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")
POW_LIBM = 0   # x**y = platform powf (what a gfortran build of the reference computes)
POW_DET = 1    # x**y = trt_powf_det (include/trt_detmath.h), the bit-specified powf of the CUDA path

_lib = None
_f32p = C.POINTER(C.c_float)
_i64p = C.POINTER(C.c_int64)
_i32p = C.POINTER(C.c_int32)
_f64p = C.POINTER(C.c_double)


def build(force=False):
    """Compile liboracle.so with oracle/Makefile (gcc -O2 -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("troute_oracle.c", "mc_wrfhydro.c", "mc_kernel.inc")]
    stale = os.path.exists(LIB_PATH) and any(os.path.getmtime(f) > os.path.getmtime(LIB_PATH) for f in srcs)
    if force or stale or not os.path.exists(LIB_PATH):
        flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
        args = ["make", "-C", _HERE] + (["-B"] if force else [])
        if " fma" not in flags:
            args.append("FMA=")
        subprocess.run(args, check=True, stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_mc_segment.restype = C.c_int
        L.oracle_mc_segment.argtypes = [C.c_int] + [C.c_float] * 15 + [_f32p]
        L.oracle_mc_segment_batch.restype = None
        L.oracle_mc_segment_batch.argtypes = [C.c_int, C.c_long, _f32p, _f32p, _i32p]
        L.oracle_powf_det_array.argtypes = [C.c_long, _f32p, _f32p, _f32p]
        L.oracle_powf_libm_array.argtypes = [C.c_long, _f32p, _f32p, _f32p]
        L.oracle_levelpool_series.restype = None
        L.oracle_levelpool_series.argtypes = [C.c_int, _f64p, C.c_long, _f32p, C.c_float, C.c_float, _f32p, _f32p, _f32p]
        L.oracle_simple_da_with_decay.restype = C.c_float
        L.oracle_simple_da_with_decay.argtypes = [C.c_float] * 4
        L.oracle_max_threads.restype = C.c_int
        L.oracle_wrfhydro_mc_batch.restype = None
        L.oracle_wrfhydro_mc_batch.argtypes = [C.c_int, C.c_longlong, _f32p, _f32p, _i32p, _i32p]
        L.oracle_route_network.restype = C.c_int
        L.oracle_route_network.argtypes = [
            C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,            # pow_mode nsteps dt qts short_ts
            C.c_int64,                                               # n_rows
            C.c_int64, _i64p, _i64p, _i32p,                          # reaches
            _i64p, _i64p,                                            # upstream CSR per reach
            _i32p, _f64p,                                            # reach_wbody, wbody_cols
            _f32p, C.c_int, _i32p,                                   # data_values ncols scols
            _f32p,                                                   # initial_conditions
            _f32p, C.c_int,                                          # qlat nqcols
            C.c_int32, C.c_int32, _f32p, _i32p, _i32p, _i32p,        # gages
            _f32p, _f32p, C.c_double,                                # lastobs init, decay
            _f32p, _f32p, _f32p,                                     # lastobs out, nudge out
            C.c_int64, _i64p, _i64p, _i64p, C.c_int,                 # job decomposition, nthreads
            _f32p, _f32p, _i64p,                                     # flowveldepth, upstream_array, iter_hist
        ]
        _lib = L
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


# -------------------------------------------------------------------------------------------------
def mc_segment(dt, qup, quc, qdp, ql, dx, bw, tw, twcc, n, ncc, cs, s0, velp, depthp, pow_mode=POW_LIBM):
    """reach.compute_reach_kernel (reach.pyx:66-103): dict with qdc, velc, depthc, ck, cn, X."""
    out = (C.c_float * 6)()
    iters = lib().oracle_mc_segment(pow_mode, *[np.float32(v) for v in
                                                 (dt, qup, quc, qdp, ql, dx, bw, tw, twcc, n, ncc, cs, s0, velp, depthp)], out)
    keys = ("qdc", "velc", "depthc", "ck", "cn", "X")
    rv = {k: np.float32(out[i]) for i, k in enumerate(keys)}
    rv["iters"] = iters
    return rv


def mc_segment_batch(in15, pow_mode=POW_DET):
    in15 = np.ascontiguousarray(in15, dtype=np.float32).reshape(-1, 15)
    out = np.empty((in15.shape[0], 6), dtype=np.float32)
    iters = np.empty(in15.shape[0], dtype=np.int32)
    lib().oracle_mc_segment_batch(pow_mode, in15.shape[0], _p(in15, C.c_float), _p(out, C.c_float), _p(iters, C.c_int32))
    return out, iters


WRF_RETRY, WRF_NO_FLOODPLAIN, WRF_ZERO_CELERITY, WRF_ZERO_PERIMETER, WRF_ONLY_QUC = 1, 2, 4, 8, 16


def wrfhydro_mc_batch(in15, pow_mode=POW_DET):
    """The WRF-Hydro original (MUSKINGCUNGE.f90, oracle/mc_wrfhydro.c) on [count, 15] rows in the argument order of
    mc_segment_batch -> ([count, 3] (qdc, velc, depthc), flags [count] (WRF_* bits: where the two Fortran sources differ by
    design), secant trips [count])."""
    in15 = np.ascontiguousarray(in15, dtype=np.float32).reshape(-1, 15)
    out = np.zeros((in15.shape[0], 3), dtype=np.float32)
    flags = np.zeros(in15.shape[0], dtype=np.int32)
    iters = np.zeros(in15.shape[0], dtype=np.int32)
    lib().oracle_wrfhydro_mc_batch(pow_mode, in15.shape[0], _p(in15, C.c_float), _p(out, C.c_float), _p(flags, C.c_int32),
                                   _p(iters, C.c_int32))
    return out, flags, iters


def powf(x, y, pow_mode=POW_DET):
    x = np.ascontiguousarray(x, dtype=np.float32)
    y = np.ascontiguousarray(y, dtype=np.float32)
    out = np.empty_like(x)
    fn = lib().oracle_powf_det_array if pow_mode == POW_DET else lib().oracle_powf_libm_array
    fn(x.shape[0], _p(x, C.c_float), _p(y, C.c_float), _p(out, C.c_float))
    return out


def levelpool_series(wbody_row, inflow, lateral_inflow=0.0, routing_period=300.0, pow_mode=POW_LIBM):
    wbody_row = np.ascontiguousarray(wbody_row, dtype=np.float64).reshape(11)
    inflow = np.ascontiguousarray(inflow, dtype=np.float32)
    out2 = np.empty(2, dtype=np.float32)
    q = np.empty_like(inflow)
    h = np.empty_like(inflow)
    lib().oracle_levelpool_series(pow_mode, _p(wbody_row, C.c_double), inflow.shape[0], _p(inflow, C.c_float),
                                  lateral_inflow, routing_period, _p(out2, C.c_float), _p(q, C.c_float), _p(h, C.c_float))
    return q, h


def simple_da_with_decay(last_valid_obs, model_val, minutes_since_last_valid, decay_coeff):
    return float(lib().oracle_simple_da_with_decay(last_valid_obs, model_val, minutes_since_last_valid, decay_coeff))


def max_threads():
    return int(lib().oracle_max_threads())


# -------------------------------------------------------------------------------------------------
def binary_find(arr, els):
    """mc_reach.pyx:36-66 -- positions of `els` in sorted `arr`; ValueError when one is absent."""
    arr = np.asarray(arr)
    els = np.asarray(list(els), dtype=arr.dtype if arr.size else np.int64)
    if els.size == 0:
        return np.zeros(0, dtype=np.int64)
    idx = np.searchsorted(arr, els)
    bad = (idx >= arr.shape[0]) | (arr[np.minimum(idx, arr.shape[0] - 1)] != els)
    if bad.any():
        raise ValueError(f"element {els[bad][0]} not found in {arr}")
    return idx.astype(np.int64)


def column_mapper(src_cols):
    index = {label: i for i, label in enumerate(src_cols)}
    return [index[label] for label in ["dt", "dx", "bw", "tw", "twcc", "n", "ncc", "cs", "s0"]]


def route_network_flat(nsteps, dt, qts_subdivisions, n_rows, reach_ptr, reach_rows, reach_type, reach_up_ptr,
                       reach_up_rows, data_values, scols, initial_conditions, qlat, assume_short_ts=False,
                       reach_wbody=None, wbody_cols=None, pow_mode=POW_DET, flowveldepth=None, gages=None,
                       jobs=None, nthreads=0, want_hist=False):
    """Flat-array call into oracle_route_network.  Returns (flowveldepth [n_rows, nsteps+1, 3],
    upstream_array [n_rows, nsteps+1], extras dict)."""
    L = lib()
    n_reaches = len(reach_ptr) - 1
    reach_ptr = np.ascontiguousarray(reach_ptr, dtype=np.int64)
    reach_rows = np.ascontiguousarray(reach_rows, dtype=np.int64)
    reach_type = np.ascontiguousarray(reach_type, dtype=np.int32)
    reach_up_ptr = np.ascontiguousarray(reach_up_ptr, dtype=np.int64)
    reach_up_rows = np.ascontiguousarray(reach_up_rows, dtype=np.int64)
    data_values = np.ascontiguousarray(data_values, dtype=np.float32)
    scols = np.ascontiguousarray(scols, dtype=np.int32)
    initial_conditions = np.ascontiguousarray(initial_conditions, dtype=np.float32)
    qlat = np.ascontiguousarray(qlat, dtype=np.float32)
    if reach_wbody is None:
        reach_wbody = np.full(max(n_reaches, 1), -1, dtype=np.int32)
    reach_wbody = np.ascontiguousarray(reach_wbody, dtype=np.int32)
    if wbody_cols is None or len(wbody_cols) == 0:
        wbody_cols = np.zeros((1, 11), dtype=np.float64)
    wbody_cols = np.ascontiguousarray(wbody_cols, dtype=np.float64)
    if flowveldepth is None:
        flowveldepth = np.zeros((n_rows, nsteps + 1, 3), dtype=np.float32)      # mc_reach.pyx:253
    upstream_array = np.zeros((n_rows, nsteps + 1), dtype=np.float32)
    hist = np.zeros(16, dtype=np.int64) if want_hist else None

    g = gages or {}
    n_gages = int(len(g.get("usgs_positions", [])))
    usgs_values = np.ascontiguousarray(g.get("usgs_values", np.zeros((0, 0))), dtype=np.float32)
    gage_max = int(usgs_values.shape[1]) if usgs_values.ndim == 2 else 0
    usgs_positions = np.ascontiguousarray(g.get("usgs_positions", []), dtype=np.int32)
    usgs_positions_reach = np.ascontiguousarray(g.get("usgs_positions_reach", []), dtype=np.int32)
    usgs_positions_gage = np.ascontiguousarray(g.get("usgs_positions_gage", []), dtype=np.int32)
    lastobs_values_init = np.ascontiguousarray(g.get("lastobs_values_init", []), dtype=np.float32)
    time_since_lastobs_init = np.ascontiguousarray(g.get("time_since_lastobs_init", []), dtype=np.float32)
    decay = float(g.get("da_decay_coefficient", 0.0))
    lastobs_times = np.full(max(n_gages, 1), np.nan, dtype=np.float32)
    lastobs_values = np.full(max(n_gages, 1), np.nan, dtype=np.float32)
    nudge = np.zeros((max(n_gages, 1), nsteps + 1), dtype=np.float32)

    if jobs is not None:
        order_ptr = np.ascontiguousarray(jobs["order_ptr"], dtype=np.int64)
        job_ptr = np.ascontiguousarray(jobs["job_ptr"], dtype=np.int64)
        job_reaches = np.ascontiguousarray(jobs["job_reaches"], dtype=np.int64)
        n_orders = len(order_ptr) - 1
    else:
        order_ptr = job_ptr = job_reaches = None
        n_orders = 0

    rc = L.oracle_route_network(
        pow_mode, nsteps, dt, qts_subdivisions, 1 if assume_short_ts else 0, n_rows,
        n_reaches, _p(reach_ptr, C.c_int64), _p(reach_rows, C.c_int64), _p(reach_type, C.c_int32),
        _p(reach_up_ptr, C.c_int64), _p(reach_up_rows, C.c_int64),
        _p(reach_wbody, C.c_int32), _p(wbody_cols, C.c_double),
        _p(data_values, C.c_float), data_values.shape[1], _p(scols, C.c_int32),
        _p(initial_conditions, C.c_float), _p(qlat, C.c_float), qlat.shape[1],
        n_gages, gage_max, _p(usgs_values, C.c_float), _p(usgs_positions, C.c_int32),
        _p(usgs_positions_reach, C.c_int32), _p(usgs_positions_gage, C.c_int32),
        _p(lastobs_values_init, C.c_float), _p(time_since_lastobs_init, C.c_float), decay,
        _p(lastobs_times, C.c_float), _p(lastobs_values, C.c_float), _p(nudge, C.c_float),
        n_orders, _p(order_ptr, C.c_int64) if order_ptr is not None else None,
        _p(job_ptr, C.c_int64) if job_ptr is not None else None,
        _p(job_reaches, C.c_int64) if job_reaches is not None else None, int(nthreads),
        _p(flowveldepth, C.c_float), _p(upstream_array, C.c_float), _p(hist, C.c_int64) if hist is not None else None)
    if rc == -2:
        raise ValueError("Number of columns (timesteps) in Qlat is incorrect")   # mc_reach.pyx:246-247
    if rc != 0:
        raise RuntimeError(f"oracle_route_network failed: {rc}")
    extras = {"lastobs_times": lastobs_times[:n_gages], "lastobs_values": lastobs_values[:n_gages],
              "nudge": nudge[:n_gages], "iter_hist": hist}
    return flowveldepth, upstream_array, extras


def flatten_reaches(reaches_wTypes, upstream_connections, data_idx, lake_numbers_col=()):
    """The object set-up of mc_reach.pyx:287-378 as flat arrays (rows = binary_find positions)."""
    reach_ptr = [0]
    reach_rows = []
    reach_type = []
    reach_up_ptr = [0]
    reach_up_rows = []
    reach_wbody = []
    lake_numbers_col = list(lake_numbers_col)
    for reach, rtype in reaches_wTypes:
        upstream_reach = upstream_connections.get(reach[0], ())
        reach_up_rows.extend(binary_find(data_idx, upstream_reach).tolist())      # :288-289
        reach_up_ptr.append(len(reach_up_rows))
        reach_rows.extend(binary_find(data_idx, reach).tolist())                  # :293 / :359
        reach_ptr.append(len(reach_rows))
        reach_type.append(int(rtype))
        if rtype == 1:
            reach_wbody.append(int(binary_find(np.asarray(lake_numbers_col), reach)[0]))   # :294
        else:
            reach_wbody.append(-1)
    return (np.asarray(reach_ptr, np.int64), np.asarray(reach_rows, np.int64), np.asarray(reach_type, np.int32),
            np.asarray(reach_up_ptr, np.int64), np.asarray(reach_up_rows, np.int64), np.asarray(reach_wbody, np.int32))


def compute_network_structured(
    nsteps, dt, qts_subdivisions, reaches_wTypes, upstream_connections, data_idx, data_cols, data_values,
    initial_conditions, qlat_values, lake_numbers_col, wbody_cols, data_assimilation_parameters, reservoir_types,
    reservoir_type_specified, model_start_time, usgs_values, usgs_positions, usgs_positions_reach,
    usgs_positions_gage, lastobs_values_init, time_since_lastobs_init, da_decay_coefficient,
    reservoir_usgs_obs=None, reservoir_usgs_wbody_idx=None, reservoir_usgs_time=None, reservoir_usgs_update_time=None,
    reservoir_usgs_prev_persisted_flow=None, reservoir_usgs_persistence_update_time=None,
    reservoir_usgs_persistence_index=None, reservoir_usace_obs=None, reservoir_usace_wbody_idx=None,
    reservoir_usace_time=None, reservoir_usace_update_time=None, reservoir_usace_prev_persisted_flow=None,
    reservoir_usace_persistence_update_time=None, reservoir_usace_persistence_index=None, reservoir_rfc_obs=None,
    reservoir_rfc_wbody_idx=None, reservoir_rfc_totalCounts=None, reservoir_rfc_file=None,
    reservoir_rfc_use_forecast=None, reservoir_rfc_timeseries_idx=None, reservoir_rfc_update_time=None,
    reservoir_rfc_da_timestep=None, reservoir_rfc_persist_days=None, great_lakes_idx=None, great_lakes_times=None,
    great_lakes_discharge=None, great_lakes_param_idx=None, great_lakes_param_prev_assim_flow=None,
    great_lakes_param_prev_assim_times=None, great_lakes_param_update_times=None, great_lakes_climatology=None,
    upstream_results={}, assume_short_ts=False, return_courant=False, da_check_gage=-1, from_files=True,
    pow_mode=POW_DET,
):
    """Oracle with the reference's signature (mc_reach.pyx:164-224); hybrid / RFC / Great-Lakes DA
    arguments are accepted and ignored (they must be empty: those reservoir types are out of scope)."""
    data_idx = np.asarray(data_idx, dtype=np.int64)
    data_values = np.ascontiguousarray(data_values, dtype=np.float32)
    initial_conditions = np.ascontiguousarray(initial_conditions, dtype=np.float32)
    qlat_values = np.ascontiguousarray(qlat_values, dtype=np.float32)
    n_rows = data_idx.shape[0]
    if qlat_values.shape[0] != n_rows:                                            # :243-244
        raise ValueError(f"Number of rows in Qlat is incorrect: expected ({n_rows}), got ({qlat_values.shape[0]})")
    if qlat_values.shape[1] < nsteps / qts_subdivisions:                          # :246-247
        raise ValueError("Number of columns (timesteps) in Qlat is incorrect")
    if data_values.shape[0] != n_rows or data_values.shape[1] != len(data_cols):  # :249-250
        raise ValueError("data_values shape mismatch")
    scols = np.asarray(column_mapper(list(data_cols)), dtype=np.int32)
    (reach_ptr, reach_rows, reach_type, reach_up_ptr, reach_up_rows, reach_wbody) = flatten_reaches(
        reaches_wTypes, upstream_connections, data_idx, lake_numbers_col)

    fvd = np.zeros((n_rows, nsteps + 1, 3), dtype=np.float32)
    fill_index_mask = np.ones(n_rows, dtype=bool)
    lake_set = set(lake_numbers_col)
    wb = np.asarray(wbody_cols, dtype=np.float64).reshape(-1, 11) if len(lake_numbers_col) else np.zeros((0, 11))
    for upstream_tw_id, tmp in upstream_results.items():                          # :458-469
        fill_index = tmp["position_index"]
        fill_index_mask[fill_index] = False
        res = np.asarray(tmp["results"], dtype=np.float32)
        fvd[fill_index, 1:, :] = res.reshape(nsteps, 3)
        if int(data_idx[fill_index]) in lake_set:
            res_idx = int(binary_find(np.asarray(list(lake_numbers_col)), [int(data_idx[fill_index])])[0])
            fvd[fill_index, 0, 0] = wb[res_idx, 9]
        else:
            fvd[fill_index, 0, 0] = initial_conditions[fill_index, 0]
            fvd[fill_index, 0, 2] = initial_conditions[fill_index, 2]

    gages = None
    usgs_positions = np.asarray(usgs_positions, dtype=np.int32)
    if usgs_positions.shape[0]:
        gages = dict(usgs_values=usgs_values, usgs_positions=usgs_positions, usgs_positions_reach=usgs_positions_reach,
                     usgs_positions_gage=usgs_positions_gage, lastobs_values_init=lastobs_values_init,
                     time_since_lastobs_init=time_since_lastobs_init, da_decay_coefficient=da_decay_coefficient)

    fvd, upstream_array, extras = route_network_flat(
        nsteps, dt, qts_subdivisions, n_rows, reach_ptr, reach_rows, reach_type, reach_up_ptr, reach_up_rows,
        data_values, scols, initial_conditions, qlat_values, assume_short_ts=assume_short_ts, reach_wbody=reach_wbody,
        wbody_cols=wb, pow_mode=pow_mode, flowveldepth=fvd, gages=gages)

    output = fvd[:, 1:, :]
    output_upstream = upstream_array[:, 1:]
    empty_f = np.zeros(0, dtype=np.float32)
    empty_i = np.zeros(0, dtype=np.int32)
    return (
        data_idx.astype(np.intp)[fill_index_mask],
        output.reshape(n_rows, -1)[fill_index_mask],
        0,
        (np.asarray([data_idx[p] for p in usgs_positions]), extras["lastobs_times"], extras["lastobs_values"]),
        (empty_i, empty_f, empty_f, empty_f, empty_f),
        (empty_i, empty_f, empty_f, empty_f, empty_f),
        output_upstream.reshape(n_rows, -1)[fill_index_mask],
        (empty_i, empty_f, empty_i),
        extras["nudge"],
        (empty_i, empty_f, empty_i, empty_i),
    )
