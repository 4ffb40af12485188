"""Results -> `troute_output_<t0>` files: the stream output of the reference (nhd_io.write_flowveldepth, nhd_io.py:2363-2461,
called from nwm_routing/output.py:281-305 with the flowveldepth table assembled from the results tuples).

Same arguments, same file names, same row / column selection:

* `flowveldepth`  DataFrame indexed by segment id, columns (timestep, 'q' | 'v' | 'd') in the interleaved order of
  results[1] (mc_reach.pyx:807-813);
* `nudge` / `usgs_positions_id`  results[8] / results[3][0] concatenated over the sub-networks;
* every `stream_output_internal_frequency // (dt // 60)`-th timestep is written, `stream_output_timediff` hours per file
  (-1: one file);
* `.csv` / `.pkl`: the long table of write_flowveldepth_csv_pkl (:2053-2087) -- the files are byte-identical to the
  reference's (tests/test_output.py runs the reference's own functions, compiled out of nhd_io.py, beside these);
* `.nc`: the variables, dimensions and attributes of write_flowveldepth_netcdf (:2089-2235).  The reference writes NetCDF-4
  through the netCDF4 package, which this environment does not have; this writer produces a NetCDF-3 (64-bit offset)
  file through scipy.io -- same names, types, shapes and attributes, readable by netCDF4 / xarray / scipy alike (the
  `type` variable is a (feature_id, type_strlen) character array, the classic-format spelling of the reference's
  fixed-length string).

Kept quirk: the reference looks the (featureID, Type) tuples of the output index up in a nudge table indexed by plain segment
ids (:2398-2401), so the `nudge` column of every file holds the fill value -9999 on every row; same here.

Masks: none (every segment, the reference's behaviour without `mask_output`) or a `wb:` list of segment ids (9999 = all).
`nex:` masks aggregate flowpaths to their nexus (updated_flowveldepth :2276-2361) and need the nexus crosswalk of the
hydrofabric; they are refused, not approximated.
"""
import os
from datetime import timedelta

import numpy as np
import pandas as pd


def _read_mask(stream_output_mask):
    if not stream_output_mask:
        return {}
    import yaml
    with open(stream_output_mask, "r") as f:
        return yaml.safe_load(f) or {}


def updated_flowveldepth(flowveldepth, mask_list):
    """Index (featureID, Type='wb'), rows restricted to the `wb` ids of the mask (nhd_io.py:2276-2290, :2356-2361)."""
    if mask_list and mask_list.get("nex"):
        raise NotImplementedError("nexus masks (mask_output: nex) need the nexus crosswalk of the hydrofabric")
    fvd = flowveldepth.copy(deep=True)
    fvd.index.name = "featureID"
    fvd["Type"] = "wb"
    fvd.set_index("Type", append=True, inplace=True)
    seg_id = []
    if mask_list and mask_list.get("wb"):
        seg_id = [9999] if 9999 in mask_list["wb"] else list(mask_list["wb"])
    if seg_id:
        ids = fvd.index.get_level_values("featureID")
        keep = ids.isin(ids if 9999 in seg_id else seg_id)
        sel = fvd[keep]
        if not sel.empty:
            fvd = pd.concat([sel, pd.DataFrame()])
    return fvd


def write_flowveldepth_csv_pkl(stream_output_directory, file_name, flow, velocity, depth, nudge_df, timestamps, t0):
    """nhd_io.py:2053-2087"""
    formatted = [str(timedelta(seconds=t)) for t in timestamps]
    parts = []
    for i in range(len(formatted)):
        parts.append(pd.DataFrame({"t0": str(t0), "time": formatted[i], "flow": flow.iloc[:, i], "velocity": velocity.iloc[:, i],
                                   "depth": depth.iloc[:, i], "nudge": nudge_df.iloc[:, i]}, index=flow.index))
    df = pd.concat(parts)
    df["current_time"] = pd.to_datetime(df["t0"]) + pd.to_timedelta(df["time"])
    df = df[["current_time", "flow", "velocity", "depth", "nudge"]]
    path = os.path.join(str(stream_output_directory), file_name)
    ext = file_name.split(".")[-1]
    if ext == "csv":
        df.to_csv(path, index=True)
    elif ext == "pkl":
        df.to_pickle(path)
    return df


def write_flowveldepth_netcdf(stream_output_directory, file_name, flow, velocity, depth, nudge_df, timestamps, t0):
    """The variables of nhd_io.py:2089-2235 in a NetCDF-3 (64-bit offset) file."""
    from scipy.io import netcdf_file
    path = os.path.join(str(stream_output_directory), file_name)
    types = [str(x) for x in flow.index.get_level_values("Type")]
    strlen = max(len(t) for t in types)
    with netcdf_file(path, "w", version=2) as nc:
        nc.createDimension("feature_id", len(flow))
        nc.createDimension("time", len(timestamps))
        nc.createDimension("type_strlen", strlen)
        v = nc.createVariable("time", "f8", ("time",))
        v[:] = np.asarray(timestamps, dtype=np.float64)
        v.long_name = "valid output time"; v.standard_name = "time"
        v.units = f'seconds since {t0.strftime("%Y-%m-%d %H:%M:%S")}'
        v.missing_value = -9999.0; v._FillValue = np.float64(-9999.0)
        # classic format has no 64-bit integers: ids above 2^31 would need the NetCDF-4 writer
        ids = np.asarray(flow.index.get_level_values("featureID"), dtype=np.int64)
        if ids.size and (ids.max() > np.iinfo(np.int32).max or ids.min() < np.iinfo(np.int32).min):
            raise NotImplementedError("feature ids beyond 32 bits need a NetCDF-4 writer (netCDF4 is not installed)")
        v = nc.createVariable("feature_id", "i4", ("feature_id",))
        v[:] = ids.astype(np.int32)
        v.long_name = "Segment ID"
        v = nc.createVariable("type", "c", ("feature_id", "type_strlen"))
        v[:] = np.asarray([list(t.ljust(strlen)) for t in types], dtype="S1").reshape(len(types), strlen)
        v.long_name = "Type"
        for name, frame, long_name, units in (("flow", flow, "Flow", "m3 s-1"), ("velocity", velocity, "Velocity", "m/s"),
                                              ("depth", depth, "Depth", "m"),
                                              ("nudge", nudge_df, "Streamflow Nudge Value", "m3 s-1")):
            v = nc.createVariable(name, "f4", ("feature_id", "time"))
            v[:] = frame.to_numpy(dtype=np.float32)
            v.long_name = long_name; v.units = units
            v.missing_value = np.float32(-9999.0); v._FillValue = np.float32(-9999.0)
        nc.TITLE = "OUTPUT FROM T-ROUTE"
        nc.file_reference_time = t0.strftime("%Y-%m-%d_%H:%M:%S")
        nc.code_version = ""
    return path


def write_flowveldepth(stream_output_directory, stream_output_mask, flowveldepth, nudge, usgs_positions_id, t0, dt,
                       stream_output_timediff, stream_output_type, stream_output_internal_frequency=5, cpu_pool=1,
                       poi_crosswalk=None, nexus_dict=None):
    """nhd_io.write_flowveldepth (:2363-2461); returns the list of files written."""
    mask_list = _read_mask(stream_output_mask)
    flowveldepth = updated_flowveldepth(flowveldepth, mask_list)
    n_timesteps = flowveldepth.shape[1] // 3
    ts = stream_output_internal_frequency // (dt // 60)
    ind = [i for i in range(int(ts) - 1, n_timesteps, int(ts))]
    timestamps_sec = [(i + 1) * dt for i in ind]
    flow = flowveldepth.iloc[:, 0::3].iloc[:, ind]
    velocity = flowveldepth.iloc[:, 1::3].iloc[:, ind]
    depth = flowveldepth.iloc[:, 2::3].iloc[:, ind]
    nudge = np.asarray(nudge)
    if nudge.size and np.all(nudge[:, 0] == 0):                      # the t = 0 column of results[8]
        nudge = nudge[:, 1:]
    nudge_df = pd.DataFrame(data=nudge, index=usgs_positions_id).iloc[:, ind]
    empty_ids = list(set(flowveldepth.index).difference(set(nudge_df.index)))
    empty_df = pd.DataFrame(index=empty_ids, columns=nudge_df.columns).fillna(-9999.0)
    nudge_df = pd.concat([nudge_df, empty_df]).loc[flowveldepth.index]
    writer = write_flowveldepth_netcdf if stream_output_type == ".nc" else write_flowveldepth_csv_pkl
    file_name_time = t0
    written = []
    if stream_output_timediff > 0:
        ts_per_file = int(stream_output_timediff * 60 // stream_output_internal_frequency)
        num_files = int(flowveldepth.shape[1] // 3 * dt // (stream_output_timediff * 60 * 60)) or 1
        for _ in range(num_files):
            name = "troute_output_" + file_name_time.strftime("%Y%m%d%H%M") + stream_output_type
            writer(stream_output_directory, name, flow.iloc[:, 0:ts_per_file], velocity.iloc[:, 0:ts_per_file],
                   depth.iloc[:, 0:ts_per_file], nudge_df.iloc[:, 0:ts_per_file], timestamps_sec[0:ts_per_file], t0)
            written.append(os.path.join(str(stream_output_directory), name))
            flow = flow.iloc[:, ts_per_file:]; velocity = velocity.iloc[:, ts_per_file:]; depth = depth.iloc[:, ts_per_file:]
            nudge_df = nudge_df.iloc[:, ts_per_file:]; timestamps_sec = timestamps_sec[ts_per_file:]
            file_name_time = file_name_time + timedelta(hours=stream_output_timediff)
    elif stream_output_timediff == -1:
        name = "troute_output_" + file_name_time.strftime("%Y%m%d%H%M") + stream_output_type
        writer(stream_output_directory, name, flow, velocity, depth, nudge_df, timestamps_sec, t0)
        written.append(os.path.join(str(stream_output_directory), name))
    return written


def flowveldepth_frame(results, nts):
    """The table nwm_routing/output.py:205-218 assembles from the results tuples: ids of results[0], values of results[1],
    columns (timestep, 'q' | 'v' | 'd')."""
    cols = pd.MultiIndex.from_product([range(int(nts)), ["q", "v", "d"]]).to_flat_index()
    return pd.concat([pd.DataFrame(r[1], index=r[0], columns=cols) for r in results])


# -------------------------------------------------------------------------------------------------
# lite restart files (checkpoint / resume of the window loop)
# -------------------------------------------------------------------------------------------------
def write_lite_restart(q0, waterbodies_df, t0, restart_parameters):
    """nhd_io.write_lite_restart (nhd_io.py:1458-1505), called after every routing loop (nwm_routing/__main__.py:269-277): the
    channel state `q0` and, when the domain has waterbodies, their `qd0` / `h0`, each with a `time` column holding `t0`, pickled
    as `channel_restart_<YYYYmmddHHMM>` / `waterbody_restart_<YYYYmmddHHMM>` under `lite_restart_output_directory`.  Same file
    names and the same frames as the reference writes (tests/test_output.py); returns the paths written (the reference
    returns nothing)."""
    import pathlib
    output_directory = restart_parameters.get("lite_restart_output_directory", None) if restart_parameters else None
    if not output_directory:
        return []
    output_path = pathlib.Path(output_directory)
    t0_str = t0.strftime("%Y%m%d%H%M")
    written = []
    q0_out = q0.copy()
    q0_out["time"] = t0
    path = output_path / ("channel_restart_" + t0_str)
    q0_out.to_pickle(path)
    written.append(str(path))
    if waterbodies_df is not None and not waterbodies_df.empty:
        wbody_initial_states = waterbodies_df.loc[:, ["qd0", "h0"]].copy()
        wbody_initial_states["time"] = t0
        path = output_path / ("waterbody_restart_" + t0_str)
        wbody_initial_states.to_pickle(path)
        written.append(str(path))
    return written


def read_lite_restart(file):
    """nhd_io.read_lite_restart (nhd_io.py:1433-1455): (restart states without the `time` column, restart datetime)"""
    import pathlib
    df = pd.read_pickle(pathlib.Path(file))
    t0 = df["time"].iloc[0].to_pydatetime()
    return df.drop(columns="time"), t0
