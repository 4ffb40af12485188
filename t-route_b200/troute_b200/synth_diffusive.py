"""Synthetic diffusive-wave domains in the reference's `diff_inputs` format.

A diffusive domain is what compute_diffusive_routing (/root/reference/src/troute-routing/troute/routing/compute.py:1740-1884)
hands to diffusive.compute_diffusive for ONE tailwater: the mainstem reaches that are routed with the diffusive wave,
plus the tributary reaches whose Muskingum-Cunge flows enter at the junctions, all packed into Fortran-ordered arrays
by diffusive_input_data_v02 (diffusive_utils_v02.py:659-1153).  There is no network for the LowerColorado topobathy here,
so tests and the bench build domains of the same shape: a mainstem of `n_mainstem` reaches in series (optionally a second
mainstem branch joining half way down), one tributary reach at most junctions, synthetic trapezoid + floodplain cross
sections (mxnbathy_g = 0), hourly lateral inflow, `dt`-second tributary hydrographs with a flood pulse.

Array conventions are the reference's: reaches are numbered upstream to downstream (fp_network_map,
diffusive_utils_v02.py:55-166); frnw_g[j] = (ncomp, downstream reach (1-based, -99 = terminal), number of upstream
reaches, their 1-based indices ..., 555 mainstem | -555 tributary); node arrays are (mxncomp_g, nrch_g); time series
are (nts, ..) with time first.
"""
import numpy as np


def diffusive_domain(n_mainstem=6, nodes=(4, 9), n_branch=0, nsteps=72, dt=300.0, seed=16, trib_every=1, q_head=40.0,
                     pulse=2.5, dsbc_option=2, slope=5e-4):
    """-> diff_inputs dict (keys of diffusive_utils_v02.py:1104-1153 that the solver reads)."""
    rng = np.random.default_rng(seed)
    # ---- topology: branch reaches (a second mainstem arm), then the main arm; tributaries are appended per junction
    reaches = []        # dicts: kind 'm' | 't', ncomp, ds (index into reaches or None), ups list
    def add(kind, ncomp):
        reaches.append(dict(kind=kind, ncomp=int(ncomp), ds=None, ups=[]))
        return len(reaches) - 1
    lo, hi = nodes
    head_tribs = set()
    def mainstem_reach(prev, with_trib, head):
        # tributary reaches of the junction at the head of a mainstem reach come first (they are upstream of it); the head
        # reach of an arm always has one: it carries the Muskingum-Cunge flow entering the diffusive domain
        tribs = []
        if head or with_trib:
            t = add("t", 2)
            tribs.append(t)
            if head:
                head_tribs.add(t)
        m = add("m", rng.integers(lo, hi + 1))
        if prev is not None:
            reaches[prev]["ds"] = m; reaches[m]["ups"].append(prev)
        for t in tribs:
            reaches[t]["ds"] = m; reaches[m]["ups"].append(t)
        return m
    branch = []
    for b in range(n_branch):
        branch.append(mainstem_reach(branch[-1] if branch else None, False, b == 0))
    join_at = max(1, n_mainstem // 2) if n_branch else -1
    main = []
    for r in range(n_mainstem):
        m = mainstem_reach(main[-1] if main else None, r > 0 and (r % trib_every == 0), r == 0)
        if r == join_at and branch:
            reaches[branch[-1]]["ds"] = m; reaches[m]["ups"].append(branch[-1])
        main.append(m)
    # upstream-to-downstream numbering: a reach index is larger than every reach upstream of it.  The construction order
    # above guarantees it except for the branch arm, which was created first and drains into a later reach: fine.
    nrch = len(reaches)
    mx = max(r["ncomp"] for r in reaches)
    frnw_col = 15
    frnw = np.zeros((nrch, frnw_col), dtype=np.int32)
    for j, r in enumerate(reaches):
        frnw[j, 0] = r["ncomp"]
        frnw[j, 1] = -99 if r["ds"] is None else r["ds"] + 1
        frnw[j, 2] = len(r["ups"])
        for k, u in enumerate(r["ups"]):
            frnw[j, 3 + k] = u + 1
        frnw[j, 3 + len(r["ups"])] = 555 if r["kind"] == "m" else -555

    # ---- geometry (node arrays, Fortran shape (mxncomp, nrch))
    def arr(fill=0.0):
        return np.full((mx, nrch), fill, dtype=np.float64)
    z, bo, traps, tw, twcc, mann, manncc, so, dx, iniq = (arr() for _ in range(10))
    # elevations: walk upstream from the outlet so that the last node of a reach sits on the first node of its downstream reach
    z_head = {}
    order = sorted(range(nrch), key=lambda j: -j)
    outlet_z = 10.0
    for j in order:
        r = reaches[j]
        n = r["ncomp"]
        d = rng.uniform(600.0, 2500.0, n - 1)
        dx[: n - 1, j] = d
        z_tail = outlet_z if r["ds"] is None else z_head[r["ds"]]
        s = slope * rng.uniform(0.6, 1.6)
        zz = z_tail + np.concatenate([np.cumsum((d * s)[::-1])[::-1], [0.0]])
        z[:n, j] = zz
        z_head[j] = zz[0]
        so[:n, j] = s
        width = rng.uniform(25.0, 60.0) if r["kind"] == "m" else rng.uniform(8.0, 15.0)
        bo[:n, j] = width
        traps[:n, j] = rng.uniform(1.0, 2.5)
        hbf = rng.uniform(2.0, 4.0)
        tw[:n, j] = bo[:n, j] + 2.0 * traps[:n, j] * hbf
        twcc[:n, j] = 3.0 * tw[:n, j]
        mann[:n, j] = rng.uniform(0.03, 0.05)
        manncc[:n, j] = 2.0 * mann[:n, j]
    # ---- forcing
    tfin_hr = dt * nsteps / 3600.0
    nts_ql = int(np.ceil(tfin_hr)) + 1
    nts_qtrib = nsteps + 1
    nts_db = nts_ql
    tq = np.arange(nts_qtrib) * dt / 3600.0
    qtrib = np.zeros((nts_qtrib, nrch))
    base = {}
    for j, r in enumerate(reaches):
        if r["kind"] == "t":
            base[j] = q_head if j in head_tribs else rng.uniform(3.0, 12.0)
            peak_hr = rng.uniform(0.25, 0.6) * tfin_hr
            qtrib[:, j] = base[j] * (1.0 + pulse * np.exp(-((tq - peak_hr) ** 2) / (2 * (0.12 * tfin_hr + 0.2) ** 2)))
    qlat = np.zeros((nts_ql, mx, nrch))
    for j, r in enumerate(reaches):
        if r["kind"] == "m":
            n = r["ncomp"]
            per_m = rng.uniform(2e-5, 2e-4, n - 1)                 # m2/s = (m3/s of the segment) / dx
            storm = 1.0 + 1.5 * np.exp(-((np.arange(nts_ql) - 0.4 * nts_ql) ** 2) / 6.0)
            qlat[:, : n - 1, j] = storm[:, None] * per_m[None, :]
    # initial flows: accumulate the tributary base flows downstream (reach indices grow downstream except for the branch
    # arm, which drains into a later reach: a second pass settles it)
    acc = np.zeros(nrch)
    for _ in range(2):
        for j, r in enumerate(reaches):
            acc[j] = base[j] if r["kind"] == "t" else sum(acc[u] for u in r["ups"])
    for j, r in enumerate(reaches):
        iniq[: r["ncomp"], j] = acc[j]
    timestep = np.zeros(10)
    timestep[:] = [dt, 0.0, tfin_hr, dt, 3600.0, dt, 3600.0, dt, dt, 10.0]
    para = np.array([0.95, 0.5, 10.0, 10000.0, -15.0, -10.0, 1.0, 0.02831, 0.0001, 1.0, float(dsbc_option)])
    dbcd = 1.5 + 0.5 * np.sin(np.arange(nts_db) / max(1, nts_db - 1) * np.pi)       # tailwater depth [m], dsbc_option 1
    e3 = np.zeros((0, mx, nrch))
    return dict(
        timestep_ar_g=timestep, nts_ql_g=nts_ql, nts_ub_g=nts_qtrib, nts_db_g=nts_db, ntss_ev_g=nsteps + 1,
        nts_qtrib_g=nts_qtrib, nts_da_g=1, mxncomp_g=mx, nrch_g=nrch, z_ar_g=z, bo_ar_g=bo, traps_ar_g=traps, tw_ar_g=tw,
        twcc_ar_g=twcc, mann_ar_g=mann, manncc_ar_g=manncc, so_ar_g=so, dx_ar_g=dx, iniq=iniq, frnw_col=frnw_col, frnw_g=frnw,
        qlat_g=qlat, ubcd_g=np.zeros((nts_qtrib, nrch)), dbcd_g=dbcd, qtrib_g=qtrib, paradim=11, para_ar_g=para,
        mxnbathy_g=0, x_bathy_g=e3, z_bathy_g=e3, mann_bathy_g=e3, size_bathy_g=np.zeros((mx, nrch), dtype=np.int32),
        usgs_da_g=np.zeros((1, nrch)), usgs_da_reach_g=np.zeros(nrch, dtype=np.int32), rdx_ar_g=dx.copy(), cwnrow_g=0,
        cwncol_g=0, crosswalk_g=np.zeros((0, 0)), z_thalweg_g=z.copy(),
        mainstem=[j for j, r in enumerate(reaches) if r["kind"] == "m"],
    )


def uniform_channel(n_reaches=3, ncomp=6, q=60.0, nsteps=48, dt=300.0, slope=8e-4, dx=1200.0):
    """A prismatic channel carrying a constant flow with normal depth at the outlet: uniform flow is a steady state of the
    diffusive wave, so the solver must hold flow = q and depth = normal depth at every node (property test)."""
    d = diffusive_domain(n_mainstem=n_reaches, nodes=(ncomp, ncomp), nsteps=nsteps, dt=dt, seed=1, trib_every=10 ** 6,
                         q_head=q, pulse=0.0)
    nrch = d["nrch_g"]
    for key, val in (("bo_ar_g", 40.0), ("traps_ar_g", 2.0), ("tw_ar_g", 40.0 + 2 * 2.0 * 3.0), ("twcc_ar_g", 3 * 52.0),
                     ("mann_ar_g", 0.035), ("manncc_ar_g", 0.07), ("so_ar_g", slope)):
        d[key][:] = val
    d["dx_ar_g"][:] = 0.0
    d["dx_ar_g"][: ncomp - 1, :] = dx
    d["rdx_ar_g"] = d["dx_ar_g"].copy()
    z_tail = 10.0
    for j in range(nrch - 1, -1, -1):
        d["z_ar_g"][:, j] = z_tail + slope * dx * np.arange(ncomp - 1, -1, -1)
        z_tail = d["z_ar_g"][0, j]
    d["z_thalweg_g"] = d["z_ar_g"].copy()
    d["qlat_g"][:] = 0.0
    d["iniq"][:] = q
    d["qtrib_g"][:, 0] = q          # reach 0 is the head tributary carrying the constant inflow
    return d


def with_natural_sections(d, mxnbathy=24, seed=5):
    """Replace the synthetic trapezoids of a domain by surveyed ("natural") cross sections: per mainstem node 9..mxnbathy
    vertices (x, z, Manning n) -- levee / floodplain / bank / irregular bed / bank / floodplain / levee -- in the arrays
    fp_naturalxsec_map builds from the topobathy table (diffusive_utils_v02.py:394-510): x_bathy_g, z_bathy_g, mann_bathy_g
    (mxnbathy_g, mxncomp_g, nrch_g) and size_bathy_g (mxncomp_g, nrch_g).  The lowest vertex sits on the node's z_ar_g.
    Some floodplain roughness values exceed 0.15 (the solver caps them, diffusive.f90:1811-1814)."""
    rng = np.random.default_rng(seed)
    d = dict(d)
    mx, nrch = d["mxncomp_g"], d["nrch_g"]
    xb = np.zeros((mxnbathy, mx, nrch)); zb = np.zeros_like(xb); mb = np.zeros_like(xb)
    size = np.zeros((mx, nrch), dtype=np.int32)
    for j in d["mainstem"]:
        for i in range(d["frnw_g"][j, 0]):
            nb = int(rng.integers(9, mxnbathy + 1))
            nbed = nb - 6
            bw, tw, twcc = d["bo_ar_g"][i, j], d["tw_ar_g"][i, j], d["twcc_ar_g"][i, j]
            hbf = (tw - bw) / (2.0 * d["traps_ar_g"][i, j])
            z0 = d["z_ar_g"][i, j]
            fp = (twcc - tw) / 2.0
            bed_x = fp + (tw - bw) / 2.0 + np.sort(rng.uniform(0.0, bw, nbed))
            bed_z = z0 + rng.uniform(0.0, 0.25 * hbf, nbed)
            bed_z[rng.integers(0, nbed)] = z0
            x = np.concatenate([[0.0, 0.05 * fp, fp], bed_x, [fp + tw, fp + tw + 0.95 * fp, twcc]])
            z = np.concatenate([[z0 + 3.5 * hbf, z0 + 1.3 * hbf, z0 + hbf], bed_z, [z0 + hbf, z0 + 1.2 * hbf, z0 + 3.2 * hbf]])
            n = np.concatenate([[0.2, 0.12, 0.09], rng.uniform(0.03, 0.05, nbed), [0.09, 0.16, 0.1]])
            x += rng.uniform(100.0, 5000.0)                          # surveys do not start at x = 0 (:1804-1806)
            xb[:nb, i, j], zb[:nb, i, j], mb[:nb, i, j] = x, z, n
            size[i, j] = nb
    d.update(mxnbathy_g=mxnbathy, x_bathy_g=xb, z_bathy_g=zb, mann_bathy_g=mb, size_bathy_g=size)
    return d
