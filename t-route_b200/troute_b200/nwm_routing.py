"""Mirror of nwm_routing.__main__.nwm_route (/root/reference/src/troute-nwm/src/nwm_routing/__main__.py:1122-1300), the
function the CLI loop, the BMI model (troute_model.py:232) and the tests call to route one time window: Muskingum-Cunge
(+ level pools, nudging) on the network, then -- when diffusive_network_data is given -- the diffusive wave on every
mainstem domain, fed by the MC results.  Same 41 positional arguments and keywords, same return value
(results list, subnetwork_list); the run-log writers (compute_log_mc / compute_log_diff, firstRun) are not mirrored.
"""
import logging
import time

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
