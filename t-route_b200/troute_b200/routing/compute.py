"""GPU twin of troute.routing.compute (/root/reference/src/troute-routing/troute/routing/compute.py).

`compute_nhd_routing_v02` keeps the reference's 38-parameter signature (:507-545) and its
`(results, subnetwork_list)` return (:1738), so `nwm_route` (src/troute-nwm/src/nwm_routing/__main__.py:1215-1253)
and the BMI driver (src/troute_model.py:228-271) can call it unchanged.

The reference spends this function decomposing the network for CPU parallelism (serial / by-network /
by-subnetwork-jit / -clustered / bmi, :553-1736): slicing pandas frames per (sub)network, pickling them to joblib-loky
workers and handing tail-water series from one order of sub-networks to the next.  On the GPU the parallelism is
inside the kernel (topological wavefront over ALL segments of ALL independent networks at once), so every
`parallel_compute_method` maps to ONE call of compute_network_structured on the union of the tail-waters.  The
caller only concatenates result tuples (nwm_routing/output.py:213-216, AbstractNetwork.new_q0 :177-191), so
returning a single tuple is within the contract ("row order inside a tuple is free").
"""
import time
import zlib
from collections import OrderedDict, defaultdict
from itertools import chain, repeat
import logging

import numpy as np
import pandas as pd

from .fast_reach.mc_reach import compute_network_structured

LOG = logging.getLogger("")

# compute.py:21-26 -- a new backend registers under a new key; the default stays the structured kernel
_compute_func_map = defaultdict(
    lambda: compute_network_structured,
    {
        "V02-structured": compute_network_structured,
        "V02-structured-b200": compute_network_structured,
    },
)

_WB_COLS = ["LkArea", "LkMxE", "OrificeA", "OrificeC", "OrificeE", "WeirC", "WeirE", "WeirL", "ifd", "qd0", "h0"]
_PARAM_COLS = ["dt", "bw", "tw", "twcc", "dx", "n", "ncc", "cs", "s0", "alt"]


def _build_reach_type_list(reach_list, wbodies_segs):
    """compute.py:41-47: a reach that holds a waterbody segment is of type 1.  `isdisjoint` looks the few segments of a reach
    up in the set; the reference's `set(reach) & wbodies_segs` ITERATES the smaller operand, and a set that was built large and
    emptied (symmetric_difference) keeps its table: 29 s per 100,000 segments."""
    if not wbodies_segs:
        return list(zip(reach_list, repeat(0, len(reach_list))))
    wb = set(wbodies_segs)
    return [(reaches, 0 if wb.isdisjoint(reaches) else 1) for reaches in reach_list]


_TOPO_CACHE = OrderedDict()          # see compute_nhd_routing_v02


def _frame_crc(df):
    """CRC of a frame's index and values (a few ms per million rows): cached sub-frames are reused only for identical tables"""
    v = df.values
    if v.dtype == object:
        return hash((df.shape, tuple(df.columns)))
    return zlib.crc32(np.ascontiguousarray(v).view(np.uint8), zlib.crc32(np.ascontiguousarray(df.index.values).view(np.uint8)))


def _finite_f32(v):
    """float32 view of a forcing table with NaN -> 0 and +-inf -> the largest finite value (what np.nan_to_num does: rows that
    the re-indexing added -- lakes, off-network tail-waters -- carry NaN), in one pass when there is nothing to replace"""
    v = np.asarray(v, dtype=np.float32)
    return v if np.isfinite(v).all() else np.nan_to_num(v)


def _nonempty(df):
    return df is not None and hasattr(df, "empty") and not df.empty


def compute_nhd_routing_v02(
    connections,
    rconn,
    wbody_conn,
    reaches_bytw,
    compute_func_name,
    parallel_compute_method,
    subnetwork_target_size,
    cpu_pool,
    t0,
    dt,
    nts,
    qts_subdivisions,
    independent_networks,
    param_df,
    q0,
    qlats,
    usgs_df,
    lastobs_df,
    reservoir_usgs_df,
    reservoir_usgs_param_df,
    reservoir_usace_df,
    reservoir_usace_param_df,
    reservoir_rfc_df,
    reservoir_rfc_param_df,
    great_lakes_df,
    great_lakes_param_df,
    great_lakes_climatology_df,
    da_parameter_dict,
    assume_short_ts,
    return_courant,
    waterbodies_df,
    data_assimilation_parameters,
    waterbody_types_df,
    waterbody_type_specified,
    subnetwork_list,
    flowveldepth_interorder={},
    from_files=True,
    device=0,
):
    da_decay_coefficient = da_parameter_dict.get("da_decay_coefficient", 0) if da_parameter_dict else 0
    param_df["dt"] = dt                                                        # :547-549
    param_df = param_df.astype("float32")
    start_time = time.time()
    compute_func = _compute_func_map[compute_func_name]

    for name, df in (("reservoir_usgs_df", reservoir_usgs_df), ("reservoir_usace_df", reservoir_usace_df),
                     ("reservoir_rfc_df", reservoir_rfc_df), ("great_lakes_df", great_lakes_df)):
        if _nonempty(df):
            raise NotImplementedError(f"{name}: hybrid / RFC / Great-Lakes reservoir DA is outside the GPU path")

    # union of all tail-waters: one device call routes every independent network
    offnetwork_upstreams = set(flowveldepth_interorder.keys()) if parallel_compute_method == "bmi" else set()
    # Everything below that depends on the topology and the parameter tables alone -- reach lists, reach types, the sorted
    # parameter sub-frame, the lake rows -- is built on the first call and kept in a small module-level cache, the analogue of
    # the per-network structures the reference's by-subnetwork methods keep in `subnetwork_list` between the calls of a run
    # (compute.py:556, :652-656, :823-827; `subnetwork_list` itself is returned untouched, as the reference's serial branch
    # does).  An entry is reused only for the same topology objects and bit-identical parameter / waterbody tables (CRC); a
    # CONUS call otherwise spends seconds here.
    topo_key = (id(reaches_bytw), len(reaches_bytw), id(rconn) if rconn is not None else id(independent_networks),
                tuple(sorted(offnetwork_upstreams)), param_df.shape, _frame_crc(param_df),
                _frame_crc(waterbodies_df) if _nonempty(waterbodies_df) else 0,
                _frame_crc(waterbody_types_df) if _nonempty(waterbody_types_df) else 0)
    cached = _TOPO_CACHE.get(topo_key)
    if cached is not None:
        _TOPO_CACHE.move_to_end(topo_key)
        reach_list, reaches_list_with_type = cached["reach_list"], cached["reaches_list_with_type"]
        lake_segs, waterbodies_df_sub, waterbody_types_df_sub = cached["lake_segs"], cached["wb_sub"], cached["wbt_sub"]
        param_df_sub, routed_index, upstream_connections = cached["param_df_sub"], cached["routed_index"], cached["rconn"]
    else:
        reach_list = list(chain.from_iterable(reaches_bytw.values()))
        segs = list(chain.from_iterable(reach_list))
        segs_all = segs + list(offnetwork_upstreams)
        common_segs = param_df.index.intersection(segs_all)
        wbodies_segs = set(segs_all).symmetric_difference(common_segs)             # :1406-1408

        waterbody_types_df_sub = pd.DataFrame()
        if _nonempty(waterbodies_df):
            lake_segs = list(waterbodies_df.index.intersection(segs_all))       # bmi: off-network tail-waters included (:1585-1600)
            waterbodies_df_sub = waterbodies_df.loc[lake_segs, _WB_COLS]
            if _nonempty(waterbody_types_df):
                waterbody_types_df_sub = waterbody_types_df.loc[lake_segs, ["reservoir_type"]]
        else:
            lake_segs = []
            waterbodies_df_sub = pd.DataFrame()

        param_df_sub = param_df.loc[common_segs, _PARAM_COLS].sort_index()         # :1443-1446
        reaches_list_with_type = _build_reach_type_list(reach_list, wbodies_segs)
        routed_index = param_df_sub.index.difference(list(offnetwork_upstreams))
        param_df_sub = param_df_sub.reindex(param_df_sub.index.tolist() + lake_segs).sort_index()   # :1455-1457
        # the connections of every tail-water, merged (independent_networks[tw] is rconn restricted to that basin)
        upstream_connections = rconn if rconn is not None else {
            k: v for net in independent_networks.values() for k, v in net.items()}
        _TOPO_CACHE[topo_key] = dict(reach_list=reach_list, reaches_list_with_type=reaches_list_with_type, lake_segs=lake_segs,
                                     wb_sub=waterbodies_df_sub, wbt_sub=waterbody_types_df_sub, param_df_sub=param_df_sub,
                                     routed_index=routed_index, rconn=upstream_connections)
        while len(_TOPO_CACHE) > 4:
            _TOPO_CACHE.popitem(last=False)
    # The reference slices the forcing with .loc (:1451-1452, :1640-1641): a routed segment without a qlat / q0 row is a
    # KeyError there, not a segment routed with zeros.  Off-network tail-waters carry no lateral inflow.
    for name, df in (("qlats", qlats), ("q0", q0)):
        want = routed_index if name == "qlats" else param_df_sub.index
        if df.index is want or df.index.equals(want):
            continue
        missing = want.difference(df.index)
        if len(missing):
            raise KeyError(f"{list(missing[:10])} not in index of {name} ({len(missing)} routed segments without a row)")
    # forcing / state rows follow the parameter index; lake and off-network rows carry no lateral inflow
    qlat_sub = qlats if qlats.index.equals(param_df_sub.index) else qlats.reindex(param_df_sub.index)
    q0_sub = q0 if q0.index.equals(param_df_sub.index) else q0.reindex(param_df_sub.index)

    upstream_results = {}
    for us_subn_tw in offnetwork_upstreams:                                    # :1652-1658
        pos = param_df_sub.index.get_loc(us_subn_tw)
        flowveldepth_interorder[us_subn_tw]["position_index"] = pos
        upstream_results[us_subn_tw] = flowveldepth_interorder[us_subn_tw]

    if _nonempty(usgs_df) or _nonempty(lastobs_df):
        from ._da import prep_da_dataframes, prep_da_positions_byreach
        usgs_df_sub, lastobs_df_sub, da_positions_list_byseg = prep_da_dataframes(
            usgs_df, lastobs_df, param_df_sub.index, offnetwork_upstreams)
        da_positions_list_byreach, da_positions_list_bygage = prep_da_positions_byreach(reach_list, lastobs_df_sub.index)
        usgs_values = usgs_df_sub.values.astype("float32")
        lastobs_discharge = lastobs_df_sub.get(
            "lastobs_discharge", pd.Series(index=lastobs_df_sub.index, name="Null", dtype="float32")).values.astype("float32")
        time_since_lastobs = lastobs_df_sub.get(
            "time_since_lastobs", pd.Series(index=lastobs_df_sub.index, name="Null", dtype="float32")).values.astype("float32")
    else:
        usgs_values = np.zeros((0, 0), dtype=np.float32)
        da_positions_list_byseg, da_positions_list_byreach, da_positions_list_bygage = [], [], []
        lastobs_discharge = np.zeros(0, dtype=np.float32)
        time_since_lastobs = np.zeros(0, dtype=np.float32)

    e_f = np.zeros(0, dtype=np.float32)
    e_i = np.zeros(0, dtype=np.int32)
    e_f2 = np.zeros((0, 0), dtype=np.float32)
    result = compute_func(
        nts, dt, qts_subdivisions, reaches_list_with_type, upstream_connections,
        param_df_sub.index.values.astype("int64"), param_df_sub.columns.values, param_df_sub.values,
        _finite_f32(q0_sub.values), _finite_f32(qlat_sub.values),
        lake_segs, waterbodies_df_sub.values, data_assimilation_parameters,
        waterbody_types_df_sub.values.astype("int32"), waterbody_type_specified,
        t0.strftime("%Y-%m-%d_%H:%M:%S") if hasattr(t0, "strftime") else str(t0),
        usgs_values, np.array(da_positions_list_byseg, dtype="int32"), np.array(da_positions_list_byreach, dtype="int32"),
        np.array(da_positions_list_bygage, dtype="int32"), lastobs_discharge, time_since_lastobs, da_decay_coefficient,
        e_f2, e_i, e_f, e_f, e_f, e_f, e_f,          # USGS hybrid reservoir DA
        e_f2, e_i, e_f, e_f, e_f, e_f, e_f,          # USACE hybrid reservoir DA
        e_f2, e_i, e_i, [], e_i, e_i, e_f, e_i, e_i,  # RFC reservoir DA
        e_i, e_i, e_f, e_i, e_f, e_i, e_i, e_f2,     # Great Lakes DA
        upstream_results, assume_short_ts, return_courant, from_files=from_files, device=device,
    )
    LOG.debug("B200 routing of %d segments x %d steps complete in %s seconds.", len(param_df_sub), nts,
              time.time() - start_time)
    return [result], subnetwork_list


def compute_diffusive_routing(
    results, diffusive_network_data, cpu_pool, t0, dt, nts, q0, qlats, qts_subdivisions, usgs_df, lastobs_df,
    da_parameter_dict, waterbodies_df, topobathy, refactored_diffusive_domain, refactored_reaches,
    coastal_boundary_depth_df, unrefactored_topobathy,
):
    """Mirror of troute.routing.compute.compute_diffusive_routing (compute.py:1740-1884): diffusive-wave routing of the
    mainstem of every tailwater domain, fed at its junctions by the Muskingum-Cunge flows already in `results`.

    Same arguments, same list of 10-tuples (ids, [q, NaN, depth] x nts per segment, 0, placeholders ...).  The reference
    loops over the tailwaters and calls the Fortran solver once per domain (":1764 TODO by-network parallel loop"); here all
    domains are packed first and routed by ONE trt_diffnw_batch call, one CTA per domain."""
    from . import diffusive_utils as diff_utils
    from .fast_reach import diffusive
    if refactored_diffusive_domain:
        raise NotImplementedError("refactored hydrofabric (see diffusive_utils.diffusive_input_data_v02)")
    packed = []
    for tw in diffusive_network_data:
        net = diffusive_network_data[tw]
        # junction inflows: the flow series of the tributary segments out of the MC results (:1764-1781)
        segs, flows = [], []
        for r in results:
            x = np.isin(r[0], net["tributary_segments"])
            if x.any():
                segs.append(np.asarray(r[0])[x]); flows.append(np.asarray(r[1])[x, ::3])
        junction_inflows = pd.DataFrame(data=np.concatenate(flows) if flows else None,
                                        index=np.concatenate(segs) if segs else None)
        topobathy_bytw = pd.DataFrame()
        if topobathy is not None and not topobathy.empty:
            topobathy_bytw = topobathy.loc[net["mainstem_segs"]]                                  # :1786-1796
        diffusive_usgs_df = usgs_df if "diffusive_streamflow_nudging" in (da_parameter_dict or {}) else pd.DataFrame()
        coastal_bytw = pd.DataFrame()
        if coastal_boundary_depth_df is not None and not coastal_boundary_depth_df.empty and tw in coastal_boundary_depth_df.index:
            coastal_bytw = coastal_boundary_depth_df.loc[tw].to_frame().T                         # :1816-1819
        diffusive_qlats = qlats.copy()
        diffusive_qlats.columns = range(diffusive_qlats.shape[1])                                 # :1823-1824
        packed.append(diff_utils.diffusive_input_data_v02(
            tw, net["connections"], net["rconn"], net["reaches"], net["mainstem_segs"], net["tributary_segments"], None,
            net["param_df"], diffusive_qlats, q0, junction_inflows, qts_subdivisions, t0, nts, dt, waterbodies_df,
            topobathy_bytw, diffusive_usgs_df, None, None, coastal_bytw, pd.DataFrame()))
    outputs = diffusive.compute_diffusive_batch(packed)
    results_diffusive = []
    e = np.asarray([])
    for tw, ins, (out_q, out_elv, out_depth) in zip(diffusive_network_data, packed, outputs):
        rch_list, dat_all = diff_utils.unpack_output(ins["pynw"], ins["ordered_reaches"], out_q, out_depth)
        x = np.isin(rch_list, diffusive_network_data[tw]["tributary_segments"])                   # MC segments: keep MC's answer
        results_diffusive.append((
            rch_list[~x], dat_all[~x, 3:], 0, (e, e, e), (e, e, e, e, e), (e, e, e, e, e), np.zeros(dat_all[~x, 3::3].shape),
            (e, e, e), np.empty(shape=(0, nts + 1), dtype="float32"), (e, e, e, e)))
    return results_diffusive
