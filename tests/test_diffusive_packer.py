"""troute_b200.routing.diffusive_utils against the reference's own packer: tests/golden/diffusive_inputs.npz holds, for five
random diffusive domains, the inputs and every array that /root/reference/.../diffusive_utils_v02.py returned for them
(tests/golden/make_golden_diffusive.py, run in the build container).  Exact equality, including the reach numbering.
Case 5 carries a non-empty `usgs_df` (gage observations with gaps): the gage arrays of diffusive DA are packed like the
reference packs them, although its Fortran ignores them."""
import datetime
import os

import numpy as np
import pandas as pd
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "diffusive_inputs.npz")
ARRAYS = ["timestep_ar_g", "z_ar_g", "bo_ar_g", "traps_ar_g", "tw_ar_g", "twcc_ar_g", "mann_ar_g", "manncc_ar_g", "so_ar_g",
          "dx_ar_g", "frnw_g", "qlat_g", "ubcd_g", "dbcd_g", "qtrib_g", "para_ar_g", "x_bathy_g", "z_bathy_g", "mann_bathy_g",
          "size_bathy_g", "iniq", "usgs_da_g", "usgs_da_reach_g", "rdx_ar_g", "crosswalk_g", "z_thalweg_g"]
SCALARS = ["nts_ql_g", "nts_ub_g", "nts_db_g", "nts_qtrib_g", "ntss_ev_g", "nts_da_g", "mxncomp_g", "nrch_g", "frnw_col", "paradim",
           "mxnbathy_g", "cwnrow_g", "cwncol_g"]


def rebuild(g, c):
    """the DataFrames / dicts of case c, as compute_diffusive_routing passes them (compute.py:1826-1850)"""
    p = f"c{c}_"
    segs = g[p + "segs"].tolist()
    n_main = int(g[p + "n_main"])
    main, tribs = segs[:n_main], segs[n_main:]
    connections = {s: ([int(d)] if d >= 0 else []) for s, d in zip(segs, g[p + "down"].tolist())}
    rconn, k = {}, 0
    flat = g[p + "rconn_flat"].tolist()
    for s, cnt in zip(segs, g[p + "rconn_cnt"].tolist()):
        rconn[s] = flat[k:k + cnt]; k += cnt
    cols = ["dx", "bw", "tw", "twcc", "n", "ncc", "cs", "s0", "alt"]
    param_df = pd.DataFrame(g[p + "param"], index=pd.Index(segs), columns=cols)
    qv = g[p + "qlat"]
    qlat = pd.DataFrame(qv, index=pd.Index(segs), columns=range(qv.shape[1]))
    ic = pd.DataFrame(g[p + "ic"], index=pd.Index(segs), columns=["qu0", "qd0", "h0"])
    ji = pd.DataFrame(g[p + "junction_inflows"], index=pd.Index(tribs))
    topo = pd.DataFrame()
    if p + "topo" in g:
        topo = pd.DataFrame(g[p + "topo"], columns=["relative_dist", "Z", "roughness", "cs_id"],
                            index=pd.Index(g[p + "topo_id"], name="hy_id"))
    t0 = datetime.datetime(2023, 4, 2, 0, 0, 0)
    coastal = pd.DataFrame()
    if p + "coastal" in g:
        v = g[p + "coastal"]
        coastal = pd.DataFrame(v, index=pd.Index([main[0]]), columns=[t0 + datetime.timedelta(hours=h) for h in range(v.shape[1])])
    usgs = pd.DataFrame()
    if p + "usgs" in g:
        v, dt = g[p + "usgs"], float(g[p + "dt"])
        usgs = pd.DataFrame(v, index=pd.Index(g[p + "usgs_id"]), columns=[t0 + datetime.timedelta(seconds=dt * k) for k in range(v.shape[1])])
    return dict(tw=main[0], connections=connections, rconn=rconn, main=main, tribs=tribs, param_df=param_df, qlat=qlat, ic=ic,
                ji=ji, topo=topo, coastal=coastal, usgs=usgs, t0=t0, nsteps=int(g[p + "nsteps"]), dt=float(g[p + "dt"]))


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def pack(d):
    from troute_b200.routing import diffusive_utils as du
    # the reach list the caller passes is only counted (nrch_g, mxncomp_g); build it with the packer's own decomposition
    reaches = [r for _, r in du._decompose(d["tw"], d["rconn"], set(d["tribs"]))]
    return du.diffusive_input_data_v02(
        d["tw"], d["connections"], d["rconn"], reaches, d["main"], d["tribs"], None, d["param_df"], d["qlat"], d["ic"], d["ji"],
        12, d["t0"], d["nsteps"], d["dt"], pd.DataFrame(), d["topo"], d["usgs"], None, None, d["coastal"], pd.DataFrame())


@pytest.mark.parametrize("c", range(6))
def test_packer_reproduces_the_reference_arrays(gold, c):
    d = rebuild(gold, c)
    ins = pack(d)
    p = f"c{c}_"
    assert [ins["pynw"][j] for j in range(len(ins["pynw"]))] == gold[p + "out_pynw"].tolist()          # reach numbering
    assert [int(ins[k]) for k in SCALARS] == gold[p + "out_scalars"].tolist()
    for k in ARRAYS:
        want, got = gold[p + "out_" + k], np.asarray(ins[k])
        assert got.shape == want.shape, (k, got.shape, want.shape)
        assert got.dtype.kind == want.dtype.kind, (k, got.dtype, want.dtype)
        assert np.array_equal(got, want, equal_nan=True), k


@pytest.mark.parametrize("c", range(6))
def test_unpack_output_reproduces_the_reference(gold, c):
    from troute_b200.routing import diffusive_utils as du
    d = rebuild(gold, c)
    ins = pack(d)
    nev, mx, nr = int(ins["ntss_ev_g"]), int(ins["mxncomp_g"]), int(ins["nrch_g"])
    tt, ii, jj = np.meshgrid(np.arange(nev), np.arange(mx), np.arange(nr), indexing="ij")
    q = 1000.0 * tt + 10.0 * ii + jj / 100.0
    ids, dat = du.unpack_output(ins["pynw"], ins["ordered_reaches"], q, q + 0.5)
    p = f"c{c}_"
    assert ids.tolist() == gold[p + "unpack_ids"].tolist()
    assert dat.dtype == np.float32 and np.array_equal(dat, gold[p + "unpack_dat"], equal_nan=True)


def test_packed_domain_routes_through_the_solver_source(gold):
    """The packer's dict is what the solver consumes: route case 0 (synthetic sections) and case 2 (surveyed sections) with
    the oracle and with the host build of the product's solver source -- bit-equal, finite, every mainstem segment present
    in the unpacked result."""
    import helpers_diffusive as HD
    from oracle import diffusive as od
    from troute_b200.routing import diffusive_utils as du
    od.build()
    for c in (0, 2):
        d = rebuild(gold, c)
        ins = pack(d)
        ref = od.compute_diffusive(ins, od.POW_DET)
        got = HD.replica_compute_diffusive(ins)
        for a, b in zip(ref, got):
            HD.assert_bits64(b, a, f"case {c}")
        ids, dat = du.unpack_output(ins["pynw"], ins["ordered_reaches"], ref[0], ref[2])
        keep = np.isin(ids, d["main"])
        assert sorted(ids[keep].tolist()) == sorted(d["main"])
        assert np.isfinite(dat[keep][:, 0::3]).all()       # (random bed elevations: no claim about the hydraulics here)


def test_refused_inputs():
    from troute_b200.routing import diffusive_utils as du
    with pytest.raises(NotImplementedError):
        du.diffusive_input_data_v02(1, {}, {}, [], [], [], None, None, None, None, None, 12, None, 1, 300.0, None, None,
                                    pd.DataFrame(), {"x": 1}, None, None, None)


def hybrid_case(gold, cases=(0, 4)):
    """Inputs of compute_diffusive_routing for several tailwater domains at once: `results` is what
    compute_nhd_routing_v02 returned for the Muskingum-Cunge part (here: the recorded junction inflows as q, zeros for
    v and d), `diffusive_network_data` what MCwithDiffusive.update_routing_domain builds (AbstractRouting.py:274-312)."""
    from troute_b200.routing import diffusive_utils as du
    dnd, results, q0, qlats, params = {}, [], [], [], []
    nsteps = None
    for c in cases:
        d = rebuild(gold, c)
        assert nsteps in (None, d["nsteps"]) or True
        nsteps = d["nsteps"] if nsteps is None else min(nsteps, d["nsteps"])
    for c in cases:
        d = rebuild(gold, c)
        dnd[d["tw"]] = dict(connections=d["connections"], rconn=d["rconn"], mainstem_segs=d["main"], tributary_segments=d["tribs"],
                            reaches=[r for _, r in du._decompose(d["tw"], d["rconn"], set(d["tribs"]))], param_df=d["param_df"])
        fvd = np.zeros((len(d["tribs"]), 3 * nsteps), dtype=np.float32)
        fvd[:, ::3] = d["ji"].values[:, :nsteps]
        results.append((np.asarray(d["tribs"], dtype=np.int64), fvd, 0))
        q0.append(d["ic"]); qlats.append(d["qlat"])
    return dnd, results, pd.concat(q0), pd.concat(qlats).fillna(0.0), nsteps


def test_compute_diffusive_routing_glue(gold, monkeypatch):
    """compute_diffusive_routing (compute.py:1740-1884 mirrored) with the device call replaced by the host build of the
    solver source: junction inflows are pulled out of the MC results, every domain is packed, routed and unpacked; the
    returned tuples hold exactly the mainstem segments with [q, NaN, depth] per output step after the initial one."""
    import helpers_diffusive as HD
    from troute_b200.routing import compute
    from troute_b200.routing.fast_reach import diffusive
    seen = []

    def fake_batch(list_of_inputs):
        seen.extend(list_of_inputs)
        return [HD.replica_compute_diffusive(d) for d in list_of_inputs]
    monkeypatch.setattr(diffusive, "compute_diffusive_batch", fake_batch)
    dnd, results, q0, qlats, nsteps = hybrid_case(gold)
    t0 = datetime.datetime(2023, 4, 2)
    out = compute.compute_diffusive_routing(results, dnd, None, t0, 300.0, nsteps, q0, qlats, 12, pd.DataFrame(), pd.DataFrame(),
                                            {}, pd.DataFrame(), pd.DataFrame(), None, None, pd.DataFrame(), pd.DataFrame())
    assert len(out) == len(dnd) == len(seen)
    for (tw, net), tup, ins in zip(dnd.items(), out, seen):
        assert len(tup) == 10 and tup[2] == 0
        assert sorted(tup[0].tolist()) == sorted(net["mainstem_segs"])
        assert tup[1].shape == (len(net["mainstem_segs"]), 3 * nsteps) and tup[1].dtype == np.float32
        assert np.isnan(tup[1][:, 1::3]).all() and tup[6].shape == (len(net["mainstem_segs"]), nsteps)
        assert tup[8].shape == (0, nsteps + 1)
        # the junction inflows of the packed dict are the MC flows of the tributary heads
        j_of_head = {h: j for j, h in ins["pynw"].items()}
        for r in results:
            for seg, row in zip(r[0].tolist(), r[1][:, ::3]):
                if seg in net["tributary_segments"]:
                    np.testing.assert_array_equal(ins["qtrib_g"][1:, j_of_head[seg]], row.astype(np.float64))
        # flows of a mainstem segment = solver output at the node at its downstream end
        q = HD.replica_compute_diffusive(ins)[0]
        head = net["mainstem_segs"][0]
        for order in ins["ordered_reaches"]:
            for h, r in ins["ordered_reaches"][order]:
                if head in r["segments_list"]:
                    k = r["segments_list"].index(head)
                    row = tup[1][tup[0].tolist().index(head), 0::3]
                    np.testing.assert_array_equal(row, q[1:, k + 1, j_of_head[h]].astype(np.float32))
