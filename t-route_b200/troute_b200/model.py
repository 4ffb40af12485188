"""`troute_model`-shaped window driver with the model state resident on the device.

The reference's BMI model (src/troute_model.py:138-345 `troute_model.run`, driven by bmi_troute.update / update_until,
src/bmi_troute.py:318) advances the network one coupling window at a time: it turns the window's lateral inflows into a
DataFrame, calls nwm_route, and then derives the next initial state from the results ON THE HOST -- new_q0 (qu0 = qd0 = last
flow, h0 = last depth), update_waterbody_water_elevation, the last-observation table of the gages
(AbstractNetwork.py:177-198, DataAssimilation.update_after_compute) -- so the whole state of the model crosses the
host / device boundary twice per window in a GPU port of that loop.

`DeviceResidentModel.run(values, until)` keeps that state where it is computed: the first window uploads q0, every later
window starts from the last column of the flow state, the reservoir elevations and the gages' last observations the
previous window left in HBM (trt_continue); per window only the window's lateral inflows (and gage observations) go in and
only what the caller asks for comes out -- the last timestep of every segment by default (the BMI output variables of
:318-330, 12 bytes per segment), the full [n, 3 * nts] table on request (`fvd_results`, :300).  `pcie` counts the bytes,
so "zero state bytes after the first window" is a test (tests/test_gpu_model.py), not a claim.

The `values` keys read and written are the reference's (troute_model.py:155-158, :296-330).  Out of scope here, as in
nwm_routing.py: the hydrofabric / configuration parsing of troute_model.__init__ and preprocess_static_vars (the caller
hands over the flattened network), hybrid (diffusive) domains, reservoir DA.
"""
import numpy as np

from .network import RoutingNetwork


class DeviceResidentModel:
    def __init__(self, segment_index, up_ptr, up_rows, kind, data_values, data_cols, q0, time_step=300.0, qts_subdivisions=12,
                 lp_rows=None, wbody_cols=None, gages=None, assume_short_ts=False, device=0, options=None):
        """segment_index: ids of the rows (sorted, as network.segment_index); up_ptr / up_rows / kind / data_values /
        data_cols: the flattened network (RoutingNetwork); q0 [n, 3] (qu0, qd0, h0); lp_rows / wbody_cols: level pools
        (rows, the 11 waterbody columns of compute.py:1416-1430); gages: the reference's nudging arguments with row positions
        (RoutingNetwork.set_gages; `usgs_values` then comes per window through values['usgs_values'])."""
        self.segment_index = np.asarray(segment_index, dtype=np.int64)
        self.n = int(self.segment_index.shape[0])
        self._row_of_id = {int(s): i for i, s in enumerate(self.segment_index.tolist())}
        self.time_step = float(time_step)
        self.qts = int(qts_subdivisions)
        self.short_ts = bool(assume_short_ts)
        self.q0 = np.ascontiguousarray(q0, dtype=np.float32)
        self.net = RoutingNetwork(up_ptr, up_rows, kind, data_values, data_cols, device=device)
        self.lp_rows = np.asarray(lp_rows if lp_rows is not None else [], dtype=np.int64)
        if self.lp_rows.size:
            self.net.set_levelpools(self.lp_rows, wbody_cols, routing_period=self.time_step)
        for k, v in (options or {}).items():
            self.net.set_option(k, int(v))
        self._gages = gages
        self.time = 0.0
        self.windows = 0
        # bytes over PCIe: `state_h2d` = initial conditions going in, `state_d2h` = state coming back only to be sent in
        # again (never, here); `forcing_h2d` = lateral inflows / observations of the windows; `results_d2h` = what the caller read
        self.pcie = {"state_h2d": 0, "state_d2h": 0, "forcing_h2d": 0, "results_d2h": 0}

    # ---- troute_model.run ----------------------------------------------------------------------------------------
    def _qlat_table(self, values):
        """[n, columns] lateral inflow of the window in row order: rows of ids the network does not have are dropped,
        segments without a value get zeros (troute_model.py:155-166)."""
        src = np.asarray(values["land_surface_water_source__volume_flow_rate"], dtype=np.float32)
        ids = np.asarray(values["land_surface_water_source__id"]).astype(np.int64)
        src = src.reshape(ids.shape[0], -1)
        out = np.zeros((self.n, src.shape[1]), dtype=np.float32)
        rows = np.asarray([self._row_of_id.get(int(i), -1) for i in ids.tolist()], dtype=np.int64)
        ok = rows >= 0
        out[rows[ok]] = src[ok]
        return out

    def run(self, values, until=300, full_output=False):
        """Advance the model by `until` seconds (a multiple of the time step).  Reads the window's lateral inflows from
        `values` (and `usgs_values` [n_gages, nts + 1] when the model has gages), writes the BMI output variables of the
        last timestep and -- with full_output -- `fvd_results` / `fvd_index`."""
        nts = int(until / self.time_step)
        if nts < 1:
            raise ValueError("until must cover at least one time step")
        qlat = self._qlat_table(values)
        usgs = values.get("usgs_values")
        if self.windows == 0:
            if self._gages is not None:
                g = dict(self._gages)
                g["usgs_values"] = np.asarray(usgs if usgs is not None else np.full((len(g["usgs_positions"]), 0), np.nan), dtype=np.float32)
                g.setdefault("reach_len", np.ones(self.n, dtype=np.int64))
                g.setdefault("seg_rows", np.arange(self.n))
                g.setdefault("usgs_positions_reach", g["usgs_positions"])
                g.setdefault("usgs_positions_gage", np.arange(len(g["usgs_positions"]), dtype=np.int32))
                self.net.set_gages(g, nts, routing_period=self.time_step)
                self.pcie["forcing_h2d"] += int(g["usgs_values"].nbytes)
            self.net.upload(nts, self.qts, qlat, self.q0)
            self.pcie["state_h2d"] += int(self.q0.nbytes)
        else:
            # the state the previous window left on the device is the initial state of this one: nothing to send
            self.net.continue_window(nts, self.qts, qlat, usgs_values=None if usgs is None else np.asarray(usgs, dtype=np.float32))
            if usgs is not None:
                self.pcie["forcing_h2d"] += int(np.asarray(usgs, dtype=np.float32).nbytes)
        self.pcie["forcing_h2d"] += int(qlat.nbytes)
        self.net.run(self.short_ts)
        last = self.net.download_last_step()
        self.pcie["results_d2h"] += int(last.nbytes)
        # final-timestep outputs (troute_model.py:318-330, _retrieve_last_output)
        values["channel_exit_water_x-section__volume_flow_rate"] = last[:, 0].copy()
        values["channel_water_flow__speed"] = last[:, 1].copy()
        values["channel_water__mean_depth"] = last[:, 2].copy()
        if self.lp_rows.size:
            values["lake_water~outgoing__volume_flow_rate"] = last[self.lp_rows, 0].copy()
            values["lake_surface__elevation"] = last[self.lp_rows, 2].copy()
        # the state the reference hands back to its caller every window (values['q0'], :304): derived from the SAME last
        # column, for callers that checkpoint it; it is never sent back in
        values["q0"] = np.stack([last[:, 0], last[:, 0], last[:, 2]], axis=1).flatten()
        values["q0_index"] = self.segment_index
        if full_output:
            fvd, _ = self.net.download()
            self.pcie["results_d2h"] += int(fvd.nbytes)
            values["fvd_results"] = fvd.flatten()
            values["fvd_index"] = self.segment_index
        if self._gages is not None:
            nudge, lt, lv = self.net.download_gages()
            values["nudging"] = nudge[:, 1:].flatten()
            values["nudging_ids"] = self.segment_index[np.asarray(self._gages["usgs_positions"], dtype=np.int64)]
            values["lastobs_df"] = np.stack([lt, lv], axis=1).flatten()
        self.time += self.time_step * nts
        self.windows += 1
        return values

    def close(self):
        self.net.close()
