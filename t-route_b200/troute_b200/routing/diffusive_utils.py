"""Input packer / output unpacker of the diffusive-wave solver: the host-side mirror of
troute.routing.diffusive_utils_v02 (/root/reference/src/troute-routing/troute/routing/diffusive_utils_v02.py).

    diffusive_input_data_v02(tw, connections, rconn, reach_list, mainstem_seg_list, trib_seg_list, ...) -> diff_ins
    unpack_output(pynw, ordered_reaches, out_q, out_elv) -> (segment ids, [n_seg, 3 * nts] float32)

Same arguments, same dict keys, same array contents as the reference (:659-1153, :1156-1215): tests/test_diffusive_packer.py
compares every array with the output of the reference's own function on recorded random domains
(tests/golden/diffusive_inputs.npz, made by tests/golden/make_golden_diffusive.py).  The reference fills the arrays with
nested Python loops over DataFrame.loc; here every per-node table is gathered with one indexed numpy operation.

Node configuration (what the Fortran expects): a reach of n segments has n + 1 nodes; the packer appends a "fake" segment
id (last id with a '2' appended, :823-826) for the bottom node, which takes its channel geometry from the last real segment
and its bed elevation from the head of the downstream reach (:10-52).  Reaches are numbered from the outermost junction
order down to the tailwater reach (:887-893), which is the order the solver sweeps them in.

Not mirrored (raise NotImplementedError): the refactored hydrofabric (`refactored_diffusive_domain`), whose crosswalk the
reference itself no longer fills (:1036-1041).  Gage data (`usgs_df`) is packed into `usgs_da_g` / `usgs_da_reach_g` the way
the reference does (:512-574) although the Fortran solver ignores both (its DA branch is commented out,
diffusive.f90:1283-1306): compute_diffusive_routing always forwards `usgs_df` (compute.py:1800), so every hybrid run with
streamflow nudging switched on reaches this function with a non-empty frame.
"""
import math
from datetime import timedelta

import numpy as np
import pandas as pd

FRNW_COL = 20


def _fake(seg):
    """id of the ghost segment that stands for the bottom node of a reach (:825, :829)"""
    return int(str(seg) + "2")


def _decompose(tw, rconn, breaks):
    """Reaches of the domain with their junction order, in the reference's processing order.

    A segment joins the reach of its only upstream neighbour unless exactly one of the two is a break segment (a
    Muskingum-Cunge tributary); junctions and headwaters start a reach
    (nhd_network.split_at_waterbodies_and_junctions :340-359).  Order = number of reaches between a reach and the tailwater
    reach.  Listed in post-order from the tailwater with upstream neighbours taken in `rconn` order -- the order of
    nhd_network.dfs_decomposition_depth_tuple (:362-418) -- every reach as [upstream ... downstream]."""
    def joins_upstream(n):
        up = rconn.get(n, [])
        return len(up) == 1 and up[0] in rconn and ((up[0] in breaks) == (n in breaks))

    out = []
    stack = [(tw, 0, None, None)]            # (tail segment, order, chain, iterator over the head's upstream neighbours)
    while stack:
        tail, order, chain, it = stack.pop()
        if chain is None:
            chain = [tail]
            while joins_upstream(chain[-1]):
                chain.append(rconn[chain[-1]][0])
            it = iter([u for u in rconn.get(chain[-1], []) if u in rconn])
        nxt = next(it, None)
        if nxt is None:
            out.append((order, chain[::-1]))
        else:
            stack.append((tail, order, chain, it))
            stack.append((nxt, order + 1, None, None))
    return out


def _coastal_boundary(tw, coastal_boundary_depth_df, t0, t0_g, tfin_g):
    """Tailwater depth series at the spacing of the boundary file (:576-657): missing hours interpolated over at most six
    columns, non-positive depths replaced by the smallest positive one; any remaining gap switches the solver to the
    normal-depth boundary (option 2)."""
    if coastal_boundary_depth_df is None or coastal_boundary_depth_df.empty:
        dt_db = 3600.0
        n = int((tfin_g - t0_g) * 3600.0 / dt_db) + 1
        return dt_db, 2, n, np.zeros(n)
    cols = coastal_boundary_depth_df.columns
    dt_db = (cols[1] - cols[0]).total_seconds()
    n = int((tfin_g - t0_g) * 3600.0 / dt_db) + 1
    step = timedelta(minutes=dt_db / 60.0)
    stamps = pd.date_range(t0, t0 + step * (n - 1), freq=step)
    # the reference looks every stamp up by str(timestamp) among the column labels (:620-623)
    table = np.full((len(coastal_boundary_depth_df), n), np.nan)
    for k, ts in enumerate(stamps):
        if str(ts) in cols:
            table[:, k] = np.asarray(coastal_boundary_depth_df[str(ts)]).reshape(len(coastal_boundary_depth_df), -1)[:, 0]
    df = pd.DataFrame(table, index=coastal_boundary_depth_df.index, columns=stamps)
    smallest = df.where(df > 0).min(axis=1).loc[tw]
    row = df.loc[tw].copy()
    row[row <= 0] = smallest
    df.loc[tw] = row
    filled = df.interpolate(axis="columns", limit_direction="both", limit=6)
    if filled.isnull().values.any():
        return dt_db, 2, n, np.zeros(n)
    return dt_db, 1, n, np.asarray(filled.loc[tw].values, dtype=np.float64)


def _gage_arrays(flat, usgs_df, nrch_g, t0, nsteps, dt_da_g, t0_g, tfin_g):
    """Observed flows per reach at the DA spacing (fp_da_map :512-574): column j of `usgs_da_g` holds the series of the
    gage on reach j (Fortran reach order), -4444 where there is no observation; `usgs_da_reach_g[j]` = j + 1 marks the reach.
    When a reach carries several gaged segments the one furthest downstream wins (the reference overwrites in segment order)."""
    nts_da_g = int((tfin_g - t0_g) * 3600.0 / dt_da_g) + 1
    usgs_da_g = np.full((nts_da_g, nrch_g), -4444.0)
    usgs_da_reach_g = np.zeros(nrch_g, dtype="i4")
    if usgs_df is None or usgs_df.empty:
        return nts_da_g, usgs_da_g, usgs_da_reach_g
    step = timedelta(minutes=dt_da_g / 60.0)
    stamps = pd.date_range(t0, t0 + step * nsteps, freq=step)
    table = usgs_df.reindex(columns=stamps).fillna(-4444.0)
    gaged = set(table.index)
    for j, (_, r) in enumerate(flat):
        hit = [s for s in r["segments_list"] if s in gaged]
        if hit:
            usgs_da_g[:, j] = np.asarray(table.loc[hit[-1]].values, dtype=np.float64)[:nts_da_g]
            usgs_da_reach_g[j] = j + 1
    return nts_da_g, usgs_da_g, usgs_da_reach_g


def diffusive_input_data_v02(
    tw, connections, rconn, reach_list, mainstem_seg_list, trib_seg_list, diffusive_parameters, param_df, qlat,
    initial_conditions, junction_inflows, qts_subdivisions, t0, nsteps, dt, waterbodies_df, topobathy_bytw, usgs_df,
    refactored_diffusive_domain, refactored_reaches, coastal_boundary_depth_df, unrefactored_topobathy_bytw,
):
    if refactored_diffusive_domain:
        raise NotImplementedError("refactored hydrofabric: the reference no longer fills its crosswalk "
                                  "(diffusive_utils_v02.py:1036-1041)")
    # ---- clocks (:709-738) and solver parameters (:741-753)
    dt_ql_g, dt_ub_g, dt_qtrib_g, dt_da_g, saveinterval = 3600.0, dt, dt, dt, dt
    t0_g = 0.0
    tfin_g = (dt * nsteps) / 60 / 60
    timestep_ar_g = np.zeros(10)
    timestep_ar_g[[0, 1, 2, 3, 4, 5, 7, 8, 9]] = [dt, t0_g, tfin_g, saveinterval, dt_ql_g, dt_ub_g, dt_qtrib_g, dt_da_g, 10.0]
    paradim = 11
    para_ar_g = np.array([0.95, 0.5, 10.0, 10000.0, -15.0, -10.0, 1.0, 0.02831, 0.0001, 1.0, 2.0])
    nrch_g = len(reach_list)
    mxncomp_g = max(len(r) for r in reach_list) + 1

    # ---- reaches by junction order (:803-855)
    mainstem = set(mainstem_seg_list)
    tribs = set(trib_seg_list)
    tuples = sorted(_decompose(tw, rconn, set(junction_inflows.index.to_list())), key=lambda x: x[0])
    mx_jorder = tuples[-1][0]
    ordered_reaches, rchbottom_head = {}, {}
    for order, rch in tuples:
        segs = rch + [_fake(rch[-1])]
        ordered_reaches.setdefault(order, []).append([rch[0], {
            "number_segments": len(segs), "segments_list": segs,
            "upstream_bottom_segments": [_fake(u) for u in rconn[rch[0]]],
            "downstream_head_segment": connections[rch[-1]],
        }])
        rchbottom_head.setdefault(segs[-1], rch[0])
    dbfksegID = _fake(tw)
    flat = [(h, r) for x in range(mx_jorder, -1, -1) for h, r in ordered_reaches[x]]       # Fortran reach order (:887-893)
    pynw = {j: h for j, (h, _) in enumerate(flat)}
    index_of_head = {h: j for j, h in pynw.items()}
    if len(flat) != nrch_g:
        raise ValueError(f"reach_list has {nrch_g} reaches, the domain decomposes into {len(flat)}")

    # ---- network map (:55-166)
    frnw_g = np.zeros((nrch_g, FRNW_COL), dtype="int32")
    for j, (head, r) in enumerate(flat):
        ups = [index_of_head[rchbottom_head[b]] for b in r["upstream_bottom_segments"]]
        frnw_g[j, 0] = r["number_segments"]
        frnw_g[j, 2] = len(ups)
        frnw_g[j, 3:3 + len(ups)] = np.asarray(ups, dtype="int32") + 1
        if head in mainstem:
            frnw_g[j, 3 + len(ups)] = 555
        if head in tribs:
            frnw_g[j, 3 + len(ups)] = -555
        if dbfksegID in r["segments_list"]:
            frnw_g[j, 1] = -100 + 1
        else:
            frnw_g[j, 1] = index_of_head[r["downstream_head_segment"][0]] + 1

    # ---- per-node gather tables: node k of reach j reads segment geo[k, j]; the bottom node repeats the last segment
    ncomp = np.asarray([r["number_segments"] for _, r in flat])
    node = np.arange(mxncomp_g)[:, None]
    live = node < ncomp[None, :]
    geo_ids = np.zeros((mxncomp_g, nrch_g), dtype=np.int64)
    for j, (_, r) in enumerate(flat):
        s = r["segments_list"]
        geo_ids[: len(s), j] = s[:-1] + [s[-2]]
    rows = param_df.index.get_indexer(geo_ids[live])
    if (rows < 0).any():
        raise KeyError("a diffusive segment is missing from param_df")

    def gather(values):
        a = np.zeros((mxncomp_g, nrch_g))
        a[live] = np.asarray(values, dtype=np.float64)[rows]
        return a
    bo_ar_g, tw_ar_g, twcc_ar_g = gather(param_df["bw"].values), gather(param_df["tw"].values), gather(param_df["twcc"].values)
    mann_ar_g, manncc_ar_g = gather(param_df["n"].values), gather(param_df["ncc"].values)
    so_ar_g, dx_ar_g = gather(param_df["s0"].values), gather(param_df["dx"].values)
    traps_ar_g = gather(1 / param_df["cs"].values)
    # bed elevation (:10-52): own altitude; bottom node = altitude of the head of the downstream reach, or for the
    # tailwater reach the last segment's altitude lowered by s0 * dx
    z_ar_g = gather(param_df["alt"].values)
    for j, (_, r) in enumerate(flat):
        n, s = r["number_segments"], r["segments_list"]
        if dbfksegID in s:
            z_ar_g[n - 1, j] = z_ar_g[n - 2, j] - param_df.loc[s[-2], "s0"] * param_df.loc[s[-2], "dx"]
        else:
            z_ar_g[n - 1, j] = float(param_df.loc[r["downstream_head_segment"], "alt"].iloc[0])

    # ---- initial flow (:938-957), lateral inflow per metre (:242-289), tributary hydrographs (:1003-1013)
    ic_rows = initial_conditions.index.get_indexer(geo_ids[live])
    iniq = np.zeros((mxncomp_g, nrch_g))
    iniq[live] = np.maximum(np.asarray(initial_conditions["qu0"].values, dtype=np.float64)[ic_rows], 0.0001)
    nts_ql_g = math.ceil((tfin_g - t0_g) * 3600.0 / dt_ql_g)
    qlat_g = np.zeros((nts_ql_g, mxncomp_g, nrch_g))
    seg_nodes = node < (ncomp[None, :] - 1)                      # nodes that carry a real segment
    qrows = qlat.index.get_indexer(geo_ids[seg_nodes])
    qvals = np.asarray(qlat[list(range(nts_ql_g))].values, dtype=np.float64)[qrows]          # (n_seg_nodes, nts_ql)
    qlat_g[:, seg_nodes] = (qvals / dx_ar_g[seg_nodes][:, None]).T
    nts_ub_g = int((tfin_g - t0_g) * 3600.0 / dt_ub_g)
    ubcd_g = np.zeros((nts_ub_g, nrch_g))
    dt_db_g, dsbd_option, nts_db_g, dbcd_g = _coastal_boundary(tw, coastal_boundary_depth_df, t0, t0_g, tfin_g)
    timestep_ar_g[6] = dt_db_g
    para_ar_g[10] = dsbd_option
    nts_qtrib_g = int((tfin_g - t0_g) * 3600.0 / dt_qtrib_g) + 1
    qtrib_g = np.zeros((nts_qtrib_g, nrch_g))
    for j, (head, _) in enumerate(flat):
        if head not in mainstem:
            qtrib_g[1:, j] = junction_inflows.loc[head]
            qtrib_g[0, j] = initial_conditions.loc[head, "qu0"]

    # ---- surveyed cross sections (:394-510): a node takes the section of its segment; the bottom node that of the head of
    # the downstream reach, or at the tailwater the last section lowered by s0 * dx
    if topobathy_bytw is not None and not topobathy_bytw.empty:
        mxnbathy_g = int(topobathy_bytw.index.value_counts().max())
        x_bathy_g = np.zeros((mxnbathy_g, mxncomp_g, nrch_g)); z_bathy_g = np.zeros_like(x_bathy_g)
        mann_bathy_g = np.zeros_like(x_bathy_g)
        size_bathy_g = np.zeros((mxncomp_g, nrch_g), dtype="i4")
        xc, zc, nc = ("relative_dist", "Z", "roughness") if "cs_id" in topobathy_bytw.columns else ("xid_d", "z", "n")
        groups = {k: g for k, g in topobathy_bytw.groupby(level=0, sort=False)}
        x_of_reach = {h: x for x in ordered_reaches for h, _ in ordered_reaches[x]}
        for j, (head, r) in enumerate(flat):
            if head not in mainstem:
                continue
            s, n = r["segments_list"], r["number_segments"]
            for k, seg in enumerate(s):
                lower = 0.0
                if k == n - 1 and x_of_reach[head] > 0:
                    src = r["downstream_head_segment"][0]
                elif seg == dbfksegID:
                    src = s[k - 1]
                    lower = param_df.loc[src].s0 * param_df.loc[src].dx
                else:
                    src = seg
                g = groups[src]
                m = len(g)
                size_bathy_g[k, j] = m
                x_bathy_g[:m, k, j] = g[xc].values
                z_bathy_g[:m, k, j] = g[zc].values - lower
                mann_bathy_g[:m, k, j] = g[nc].values
    else:
        mxnbathy_g = 0
        x_bathy_g = np.array([]).reshape(0, 0, 0); z_bathy_g = np.array([]).reshape(0, 0, 0)
        mann_bathy_g = np.array([]).reshape(0, 0, 0); size_bathy_g = np.array([], dtype="i4").reshape(0, 0)

    # ---- gage arrays (:512-574; the solver ignores them), crosswalk placeholders (:1036-1041)
    nts_da_g, usgs_da_g, usgs_da_reach_g = _gage_arrays(flat, usgs_df, nrch_g, t0, nsteps, dt_da_g, t0_g, tfin_g)
    empty2 = np.array([]).reshape(0, 0)
    return {
        "timestep_ar_g": timestep_ar_g, "nts_ql_g": nts_ql_g, "nts_ub_g": nts_ub_g, "nts_db_g": nts_db_g,
        "nts_qtrib_g": nts_qtrib_g, "ntss_ev_g": int((tfin_g - t0_g) * 3600.0 / dt) + 1, "nts_da_g": nts_da_g,
        "mxncomp_g": mxncomp_g, "nrch_g": nrch_g, "z_ar_g": z_ar_g, "bo_ar_g": bo_ar_g, "traps_ar_g": traps_ar_g,
        "tw_ar_g": tw_ar_g, "twcc_ar_g": twcc_ar_g, "mann_ar_g": mann_ar_g, "manncc_ar_g": manncc_ar_g, "so_ar_g": so_ar_g,
        "dx_ar_g": dx_ar_g, "frnw_col": FRNW_COL, "frnw_g": frnw_g, "qlat_g": qlat_g, "ubcd_g": ubcd_g, "dbcd_g": dbcd_g,
        "qtrib_g": qtrib_g, "paradim": paradim, "para_ar_g": para_ar_g, "mxnbathy_g": mxnbathy_g, "x_bathy_g": x_bathy_g,
        "z_bathy_g": z_bathy_g, "mann_bathy_g": mann_bathy_g, "size_bathy_g": size_bathy_g, "iniq": iniq, "pynw": pynw,
        "ordered_reaches": ordered_reaches, "usgs_da_g": usgs_da_g,
        "usgs_da_reach_g": usgs_da_reach_g, "rdx_ar_g": empty2, "cwnrow_g": 0, "cwncol_g": 0,
        "crosswalk_g": empty2, "z_thalweg_g": empty2,
    }


def unpack_output(pynw, ordered_reaches, out_q, out_elv):
    """Solver output (nts, mxncomp, nrch) -> rows per segment [q, NaN, elevation] x nts, segments listed reach by reach in
    ascending junction order (:1156-1215).  A segment reads the node at its DOWNSTREAM end (nodes 1 .. n of its reach)."""
    j_of_head = {h: j for j, h in pynw.items()}
    nts = np.asarray(out_q).shape[0]
    ids, blocks = [], []
    for order in ordered_reaches.keys():
        for head, r in ordered_reaches[order]:
            segs = r["segments_list"]
            j = j_of_head[head]
            block = np.full((len(segs) - 1, nts * 3), np.nan)
            block[:, ::3] = np.asarray(out_q)[:, 1:len(segs), j].T
            block[:, 2::3] = np.asarray(out_elv)[:, 1:len(segs), j].T
            ids.extend(segs[:-1]); blocks.append(block)
    return np.asarray(ids, dtype=np.intp), np.asarray(np.concatenate(blocks), dtype="float32")
