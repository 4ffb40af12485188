"""Mirror of nwm_routing.__main__.nwm_route (/root/reference/src/troute-nwm/src/nwm_routing/__main__.py:1122-1300), the
function the CLI loop, the BMI model (troute_model.py:232) and the tests call to route one time window: Muskingum-Cunge
(+ level pools, nudging) on the network, then -- when diffusive_network_data is given -- the diffusive wave on every
mainstem domain, fed by the MC results.  Same 41 positional arguments and keywords, same return value
(results list, subnetwork_list); the run-log writers (compute_log_mc / compute_log_diff, firstRun) are not mirrored.

Also here: the state hand-off between consecutive windows that the CLI loop (__main__.py:258-266) and the BMI model
(troute_model.py:250-262) perform after every nwm_route call -- new_q0 / update_waterbody_water_elevation
(AbstractNetwork.py:177-198), new_lastobs (DataAssimilation.py:1506-1551) -- and `route_windows`, that loop without its file
I/O.  The device network stays resident across the windows (the handle cache of compute_network_structured); only the
forcing of a window and its result cross PCIe.
"""
import logging
import time

import numpy as np
import pandas as pd

from .routing.compute import compute_diffusive_routing, compute_nhd_routing_v02

LOG = logging.getLogger("")


def nwm_route(
    downstream_connections, upstream_connections, waterbodies_in_connections, reaches_bytw, parallel_compute_method,
    compute_kernel, subnetwork_target_size, cpu_pool, t0, dt, nts, qts_subdivisions, independent_networks, param_df, q0, qlats,
    usgs_df, lastobs_df, reservoir_usgs_df, reservoir_usgs_param_df, reservoir_usace_df, reservoir_usace_param_df,
    reservoir_rfc_df, reservoir_rfc_param_df, great_lakes_df, great_lakes_param_df, great_lakes_climatology_df,
    da_parameter_dict, assume_short_ts, return_courant, waterbodies_df, data_assimilation_parameters, waterbody_types_df,
    waterbody_type_specified, diffusive_network_data, topobathy_df, refactored_diffusive_domain, refactored_reaches,
    subnetwork_list, coastal_boundary_depth_df, unrefactored_topobathy_df, firstRun=False, logFileName="troute_run_log.txt",
    flowveldepth_interorder={}, from_files=False,
):
    start = time.time()
    results, subnetwork_list = compute_nhd_routing_v02(
        downstream_connections, upstream_connections, waterbodies_in_connections, reaches_bytw, compute_kernel,
        parallel_compute_method, subnetwork_target_size, cpu_pool, t0, dt, nts, qts_subdivisions, independent_networks,
        param_df, q0, qlats, usgs_df, lastobs_df, reservoir_usgs_df, reservoir_usgs_param_df, reservoir_usace_df,
        reservoir_usace_param_df, reservoir_rfc_df, reservoir_rfc_param_df, great_lakes_df, great_lakes_param_df,
        great_lakes_climatology_df, da_parameter_dict, assume_short_ts, return_courant, waterbodies_df,
        data_assimilation_parameters, waterbody_types_df, waterbody_type_specified, subnetwork_list, flowveldepth_interorder,
        from_files=from_files)
    LOG.debug("MC computation complete in %s seconds.", time.time() - start)
    if diffusive_network_data:
        t1 = time.time()
        results.extend(compute_diffusive_routing(
            results, diffusive_network_data, cpu_pool, t0, dt, nts, q0, qlats, qts_subdivisions, usgs_df, lastobs_df,
            da_parameter_dict, waterbodies_df, topobathy_df, refactored_diffusive_domain, refactored_reaches,
            coastal_boundary_depth_df, unrefactored_topobathy_df))
        LOG.debug("Diffusive computation complete in %s seconds.", time.time() - t1)
    LOG.debug("ordered reach computation complete in %s seconds.", time.time() - start)
    return results, subnetwork_list


def new_q0(run_results):
    """Initial conditions of the next window: qu0 = qd0 = last flow, h0 = last depth of every segment
    (what AbstractNetwork.new_q0 derives, AbstractNetwork.py:177-191).  One frame over all (sub)networks, rows in result
    order."""
    ids = np.concatenate([np.asarray(r[0]) for r in run_results]) if run_results else np.zeros(0, dtype=np.int64)
    last = [np.asarray(r[1]) for r in run_results]
    flow = np.concatenate([a[:, -3] for a in last]) if last else np.zeros(0, dtype=np.float32)
    depth = np.concatenate([a[:, -1] for a in last]) if last else np.zeros(0, dtype=np.float32)
    return pd.DataFrame({"qu0": flow, "qd0": flow, "h0": depth}, index=ids)


def update_waterbody_water_elevation(waterbodies_df, q0):
    """Reservoirs start the next window from their last outflow and water elevation: the `qd0` / `h0` columns of the
    waterbody table take the rows of q0 with the same (lake) id, in place (AbstractNetwork.py:193-198)."""
    shared = waterbodies_df.index.intersection(q0.index)
    for col in ("qd0", "h0"):
        if col in waterbodies_df.columns:
            waterbodies_df.loc[shared, col] = q0.loc[shared, col].to_numpy(dtype=waterbodies_df[col].dtype)
    return waterbodies_df


def new_lastobs(run_results, time_increment):
    """Last-observation table of the next window from element [3] of every result tuple -- (gage segment ids, seconds from
    the start of the window to the last assimilated observation, its value) -- with the times re-based to the start of
    the next window (DataAssimilation.new_lastobs, DataAssimilation.py:1506-1551)."""
    ids = np.concatenate([np.asarray(rr[3][0]) for rr in run_results]) if run_results else np.zeros(0, dtype=np.int64)
    # the arithmetic stays in the precision of the kernel's arrays (float32), as in the reference
    when = np.concatenate([np.asarray(rr[3][1]) for rr in run_results]) if run_results else np.zeros(0, dtype=np.float32)
    value = np.concatenate([np.asarray(rr[3][2]) for rr in run_results]) if run_results else np.zeros(0, dtype=np.float32)
    return pd.DataFrame({"time_since_lastobs": when - when.dtype.type(time_increment), "lastobs_discharge": value}, index=ids)


def route_windows(route_window, windows, q0, waterbodies_df, lastobs_df, dt, nts):
    """The loop of nwm_routing.__main__.main_v04 (:150-330) / troute_model.run without file I/O.

    route_window(window, q0, waterbodies_df, lastobs_df) -> (results, subnetwork_list) routes one window (normally a
    closure over nwm_route with that window's qlats / usgs_df).  After every window the next initial state is derived from
    its results exactly as the reference does.  Returns (list of per-window results, q0, waterbodies_df, lastobs_df)."""
    out = []
    for window in windows:
        results, _ = route_window(window, q0, waterbodies_df, lastobs_df)
        out.append(results)
        q0 = new_q0(results)
        if waterbodies_df is not None and not waterbodies_df.empty:
            update_waterbody_water_elevation(waterbodies_df, q0)
        if lastobs_df is not None and not lastobs_df.empty:
            lastobs_df = new_lastobs(results, dt * nts)
    return out, q0, waterbodies_df, lastobs_df
