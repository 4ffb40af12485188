#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ from the reference tree (/root/reference).

Run in the build container (the GPU box has no /root/reference):   python tests/golden/make_golden.py

Nothing is copied verbatim: the script READS the reference's own test files, extracts their
known-answer vectors, and imports the reference's pure-python graph module to record its outputs on
small networks.  The Fortran kernels cannot be built here (no Fortran compiler), so the only numeric
pins of the arithmetic are the reference's hard-coded expected values, recorded below.

Writes
  mc_demo_kat.json        single-segment MC known answer   (src/kernel/muskingum/mc_sseg_stime_NOLOOP_demo.py:173-248)
  mc_suite_seed16.npy     [5000, 15] float32 kernel inputs the reference's cross-implementation test feeds to
                          reach.compute_reach_kernel (demo.py:344-349 -> compare_methods :351-368 with
                          generate_conus_MC_parameters(5000, 16), test_suite_parameters.py:57-94).  NOTE the
                          reference passes the generator's tuple positionally into a differently ordered
                          signature (dt <- dx sample, qup <- bw sample, ...); the rows here are what the kernel
                          actually receives.
  levelpool_kats.json     3 level-pool fixtures + inflow series + expected final (outflow, elevation)
                          (src/troute-network/troute/network/reservoirs/test/test_compute_kernel.py:28-110,
                           :376-505, :508-637, :640-949)
  simple_da_kat.json      nudging known answer (src/troute-routing/troute/routing/test_compute.py:33-42)
  lowercolorado_v4_lakes.npz  the same network with its 19 level-pool reservoirs collapsed to nodes by the reference's graph code
  lowercolorado_v4.npz    LowerColorado_TX NextGen hydrofabric (test/LowerColorado_TX_v4/domain/*.gpkg): ids, downstream
                          ids, channel parameters, 49 h of lateral inflow (channel_forcing/*.csv) and the reach lists
                          the reference's own nhd_network.dfs_decomposition yields for it (BASELINE config 0)
  nhd_graph.json          graph fixture of troute/test_nhd_network.py:1-142 with the outputs of the reference's
                          nhd_network.{reverse_network, dfs_decomposition, build_subnetworks,
                          reachable_network} on it and on a seeded random forest
"""
import ast
import importlib.util
import json
import os
import re
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def mc_demo_kat():
    path = f"{REF}/src/kernel/muskingum/mc_sseg_stime_NOLOOP_demo.py"
    src = open(path).read()
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "single_vs_double")
    vals = {}
    # the function assigns the same names in the single and in the double branch; walk in source order and keep
    # the single-precision branch (first `if precision == "single"` body) + the top-level assignments
    def grab(stmts, into):
        for st in stmts:
            if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
                try:
                    into[st.targets[0].id] = ast.literal_eval(st.value)
                except Exception:
                    pass
    grab(fn.body, vals)
    branch = next(st for st in fn.body if isinstance(st, ast.If))
    single = {}
    grab(branch.body, single)
    double = {}
    grab(branch.orelse[0].body, double)
    inputs = dict(dt=vals["dt"], dx=vals["dx"], bw=vals["bw"], tw=vals["tw"], twcc=vals["twcc"], n=vals["n_manning"],
                  ncc=vals["n_manning_cc"], cs=vals["cs"], s0=vals["s0"], ql=vals["qlat"])
    out = {
        "source": "src/kernel/muskingum/mc_sseg_stime_NOLOOP_demo.py:173-248",
        "channel": inputs,
        "single": {"qup": single["qup"], "quc": single["quc"], "qdp": single["qdp"], "depthp": single["depthp"],
                   "velp": single["velp"],
                   "expected": {"qdc": vals["qdc_expected_sngl"], "velc": vals["velc_expected_sngl"],
                                "depthc": vals["depthc_expected_sngl"]}},
        "double": {"qup": double["qup"], "quc": double["quc"], "qdp": double["qdp"], "depthp": double["depthp"],
                   "velp": double["velp"],
                   "expected": {"qdc": vals["qdc_expected_dbl"], "velc": vals["velc_expected_dbl"],
                                "depthc": vals["depthc_expected_dbl"]}},
    }
    # the 8-row trace in the docstring (:193-200): k, i, q, vel, depth
    rows = re.findall(r"^\s+([01]) ([0-3]) (\d\.\d+) (\d\.\d+) (\d\.\d+)\s*$", src, flags=re.M)
    out["trace_single"] = [[int(a), int(b), float(c), float(d), float(e)] for a, b, c, d, e in rows[:8]]
    out["trace_double"] = [[int(a), int(b), float(c), float(d), float(e)] for a, b, c, d, e in rows[8:16]]
    json.dump(out, open(f"{OUT}/mc_demo_kat.json", "w"), indent=1)
    return out


def mc_suite():
    gen = _load(f"{REF}/src/kernel/muskingum/test_suite_parameters.py", "ref_test_suite_parameters")
    rows = []
    for p in gen.generate_conus_MC_parameters(5000, 16):
        # compare_methods("single", *p) (demo.py:346-349) binds p positionally to
        # (dt, qup, quc, qdp, qlat, dx, bw, tw, twcc, n_manning, n_manning_cc, cs, s0, depthp)   :351-368
        dt, qup, quc, qdp, qlat, dx, bw, tw, twcc, n, ncc, cs, s0, depthp = p
        # reach.compute_reach_kernel(dt, qup, quc, qdp, qlat, dx, bw, tw, twcc, n, ncc, cs, s0, 0, depthp)  :437-454
        rows.append([dt, qup, quc, qdp, qlat, dx, bw, tw, twcc, n, ncc, cs, s0, 0.0, depthp])
    arr = np.asarray(rows, dtype=np.float32)
    np.save(f"{OUT}/mc_suite_seed16.npy", arr)
    return arr


def levelpool_kats():
    path = f"{REF}/src/troute-network/troute/network/reservoirs/test/test_compute_kernel.py"
    tree = ast.parse(open(path).read())
    fns = {n.name: n for n in tree.body if isinstance(n, ast.FunctionDef)}

    def assigns(fn):
        d = {}
        for st in ast.walk(fn):
            if isinstance(st, ast.Assign) and len(st.targets) == 1 and isinstance(st.targets[0], ast.Name):
                try:
                    d[st.targets[0].id] = ast.literal_eval(st.value)
                except Exception:
                    pass
        return d

    out = {"source": "src/troute-network/troute/network/reservoirs/test/test_compute_kernel.py", "cases": []}
    for fixture, test in (("lp_reservoir", "test_lp_run"), ("lp_reservoir2", "test_lp2_run"),
                          ("lp_reservoir3", "test_lp3_run")):
        f = assigns(fns[fixture])
        t = assigns(fns[test])
        # args order of the fixture: lake_area, max_depth, orifice_area, orifice_coefficient, orifice_elevation,
        # weir_coefficient, weir_elevation, weir_length, initial_fractional_depth, 0.0, water_elevation
        wbody_row = [f["lake_area"], f["max_depth"], f["orifice_area"], f["orifice_coefficient"],
                     f["orifice_elevation"], f["weir_coefficient"], f["weir_elevation"], f["weir_length"],
                     f["initial_fractional_depth"], 0.0, f["water_elevation"]]
        out["cases"].append({
            "fixture": fixture, "test": test, "wbody_row": wbody_row, "routing_period": t["routing_period"],
            "inflow": t["inflow_list"],
            "expected_final_outflow": t["expected_final_outflow"],
            "expected_final_water_elevation": t["expected_final_water_elevation"],
        })
    json.dump(out, open(f"{OUT}/levelpool_kats.json", "w"))
    return out


def simple_da_kat():
    path = f"{REF}/src/troute-routing/troute/routing/test_compute.py"
    tree = ast.parse(open(path).read())
    mod = {}
    for st in tree.body:                                   # module-level literals (:13-31)
        if isinstance(st, ast.Assign) and isinstance(st.targets[0], ast.Name):
            try:
                mod[st.targets[0].id] = ast.literal_eval(st.value)
            except Exception:
                pass
    # test_simple_da (:33-42): lo = lastobs_old["obs"], lt = lastobs_old["time"], m = modeled_low[2]
    out = {"source": "src/troute-routing/troute/routing/test_compute.py:33-42",
           "last_valid_obs": mod["lastobs_old"]["obs"], "minutes_since_last_valid": mod["lastobs_old"]["time"],
           "model_val": mod["modeled_low"][2], "decay_coeff": mod["decay_coeff"],
           "expected": 10.483673095703125, "rel": 2.3e-06}
    src = open(path).read()
    assert "expected = 10.483673095703125" in src and "2.3e-06" in src
    json.dump(out, open(f"{OUT}/simple_da_kat.json", "w"), indent=1)
    return out


def nhd_graph():
    # nhd_network.py needs `toolz.pluck` and `deprecated`; shim toolz (absent in this image)
    if "toolz" not in sys.modules:
        tz = types.ModuleType("toolz")

        def pluck(ind, seqs):
            return (s[ind] for s in seqs)
        tz.pluck = pluck
        sys.modules["toolz"] = tz
    nn = _load(f"{REF}/src/troute-network/troute/nhd_network.py", "ref_nhd_network")
    tsrc = open(f"{REF}/src/troute-network/troute/test_nhd_network.py").read()
    head = tsrc.split("import pandas as pd")[0]
    ns = {}
    exec(compile(head, "test_nhd_network_head", "exec"), ns)

    def record(connections, target_size):
        connections = {int(k): [int(x) for x in v] for k, v in connections.items()}
        rconn = nn.reverse_network(connections)
        rec = {"connections": {str(k): v for k, v in connections.items()},
               "rconn": {str(k): sorted(v) for k, v in rconn.items()}}
        # reaches per tailwater, the way AbstractNetwork.reaches_by_tailwater builds them for MC-only runs
        # (AbstractNetwork.py:241-257 -> nhd_network.dfs_decomposition with split_at_junction)
        from functools import partial
        path_func = partial(nn.split_at_junction, rconn)
        indep = nn.reachable_network(rconn)
        rec["independent_networks"] = {str(tw): {str(k): v for k, v in net.items()} for tw, net in indep.items()}
        rec["reaches_bytw"] = {str(tw): nn.dfs_decomposition(net, path_func) for tw, net in indep.items()}
        # compute.py:557-559 calls build_subnetworks(connections, rconn, subnetwork_target_size): all tailwaters
        sn_all = nn.build_subnetworks(connections, rconn, target_size)
        subn = {str(tw): {str(order): {str(sn_tw): sorted(int(x) for x in segs) for sn_tw, segs in d.items()}
                          for order, d in sn.items()} for tw, sn in sn_all.items()}
        rec["subnetworks_target_size"] = target_size
        rec["subnetworks"] = subn
        return rec

    out = {"source": "src/troute-network/troute/test_nhd_network.py:1-142 + outputs of troute/nhd_network.py"}
    out["fixture"] = record(ns["expected_connections"], 5)
    out["fixture"]["expected_rconn_reference_literal"] = {str(k): v for k, v in ns["expected_rconn"].items()}
    # seeded random forest: node i>0 drains to a random lower id with prob .95, else it is an outlet
    rng = np.random.default_rng(16)
    n = 300
    conn = {}
    for i in range(n):
        if i == 0 or rng.random() < 0.03:
            conn[i] = []
        else:
            conn[i] = [int(rng.integers(max(0, i - 12), i))]
    out["forest300"] = record(conn, 25)
    json.dump(out, open(f"{OUT}/nhd_graph.json", "w"))
    return out


def lowercolorado_v4():
    """BASELINE config 0 plumbing case: the LowerColorado_TX NextGen hydrofabric (6,971 flowpaths) with its own channel
    forcing (49 hourly CSV files), read with troute_b200.hyfeatures (sqlite3; the reference needs geopandas), and the reach
    decomposition the REFERENCE's own code produces for an MC-only run: nhd_network.extract_connections ->
    reverse_network -> reachable_network -> dfs_decomposition(split_at_junction)  (AbstractNetwork.py:212-257)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(OUT)), "t-route_b200"))
    from functools import partial
    from troute_b200 import hyfeatures as hy
    if "toolz" not in sys.modules:
        tz = types.ModuleType("toolz")
        tz.pluck = lambda ind, seqs: (s_[ind] for s_ in seqs)
        sys.modules["toolz"] = tz
    nn = _load(f"{REF}/src/troute-network/troute/nhd_network.py", "ref_nhd_network")
    base = f"{REF}/test/LowerColorado_TX_v4"
    df = hy.read_flowpaths(f"{base}/domain/LowerColorado_NGEN_v201.gpkg")
    terminal_codes = {0} | set(df[~df["downstream"].isin(df.index)]["downstream"].values.tolist())
    conn = nn.extract_connections(df, "downstream", terminal_codes=terminal_codes)
    assert {int(k): [int(x) for x in v] for k, v in conn.items()} == hy.connections(df)
    rconn = nn.reverse_network(conn)
    indep = nn.reachable_network(rconn)
    reaches = []
    tw_of_reach = []
    for tw, net in indep.items():
        for r in nn.dfs_decomposition(net, partial(nn.split_at_junction, net)):
            reaches.append([int(x) for x in r]); tw_of_reach.append(int(tw))
    qlat = hy.read_channel_forcing(f"{base}/channel_forcing", index=df.index)
    params = hy.param_frame(df, 300.0)
    np.savez_compressed(
        f"{OUT}/lowercolorado_v4.npz",
        ids=df.index.values.astype(np.int64), downstream=df["downstream"].values.astype(np.int64),
        param_cols=np.array(params.columns.tolist()), params=params.values.astype(np.float32),
        qlat=qlat.values.astype(np.float32),
        reach_len=np.asarray([len(r) for r in reaches], dtype=np.int64),
        reach_ids=np.asarray([x for r in reaches for x in r], dtype=np.int64),
        reach_tw=np.asarray(tw_of_reach, dtype=np.int64))
    return dict(n=len(df), reaches=len(reaches), tailwaters=len(indep), qlat_cols=qlat.shape[1],
                longest_reach=max(len(r) for r in reaches))


def lowercolorado_v4_lakes():
    """The same hydrofabric with its level-pool reservoirs broken out (break_network_at_waterbodies: True): lake table and
    flowpath -> lake map read by troute_b200.hyfeatures (sqlite3), then the REFERENCE's own graph code collapses every
    lake into one node and cuts the reaches at lakes and junctions: nhd_network.replace_waterbodies_connections
    (HYFeaturesNetwork.py:520-524) -> reverse_network -> reachable_network ->
    dfs_decomposition(split_at_waterbodies_and_junctions)  (AbstractNetwork.py:241-257 with break segments = lake ids)."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(OUT)), "t-route_b200"))
    from functools import partial
    from troute_b200 import hyfeatures as hy
    if "toolz" not in sys.modules:
        tz = types.ModuleType("toolz")
        tz.pluck = lambda ind, seqs: (s_[ind] for s_ in seqs)
        sys.modules["toolz"] = tz
    nn = _load(f"{REF}/src/troute-network/troute/nhd_network.py", "ref_nhd_network")
    gpkg = f"{REF}/test/LowerColorado_TX_v4/domain/LowerColorado_NGEN_v201.gpkg"
    df = hy.read_flowpaths(gpkg)
    wb = hy.read_lakes(gpkg)
    wbody_conn = hy.waterbody_connections(df, wb)
    wb, wbody_conn = hy.drop_inconsistent_lakes(df, wb, wbody_conn)
    terminal_codes = {0} | set(df[~df["downstream"].isin(df.index)]["downstream"].values.tolist())
    conn = nn.extract_connections(df, "downstream", terminal_codes=terminal_codes)
    new_conn, link_lake = nn.replace_waterbodies_connections(conn, wbody_conn)
    new_conn = {int(k): [int(x) for x in v] for k, v in new_conn.items()}
    rconn = nn.reverse_network(new_conn)
    indep = nn.reachable_network(rconn)
    lakes = set(wbody_conn.values())
    reaches, tw_of_reach = [], []
    for tw, net in indep.items():
        for r in nn.dfs_decomposition(net, partial(nn.split_at_waterbodies_and_junctions, lakes, net)):
            reaches.append([int(x) for x in r]); tw_of_reach.append(int(tw))
    # Where one lake drains straight into another, replace_waterbodies_connections leaves the downstream lake's entry
    # flowpath as the outlet of the upstream lake although that flowpath is no longer a key of the graph
    # (reservoir_shore only excludes the lake's OWN flowpaths): reverse_network then makes it a node without a downstream
    # neighbour and the reference routes it as a one-segment tail-water.  The fixture keeps what the reference does.
    nodes = np.asarray(sorted(set(new_conn) | {x for v in new_conn.values() for x in v}), dtype=np.int64)
    down = np.asarray([new_conn[int(k)][0] if new_conn.get(int(k)) else -1 for k in nodes], dtype=np.int64)
    assert all(len(v) <= 1 for v in new_conn.values())
    np.savez_compressed(
        f"{OUT}/lowercolorado_v4_lakes.npz",
        nodes=nodes, downstream=down,
        reach_len=np.asarray([len(r) for r in reaches], dtype=np.int64),
        reach_ids=np.asarray([x for r in reaches for x in r], dtype=np.int64),
        reach_tw=np.asarray(tw_of_reach, dtype=np.int64),
        wbody_seg=np.asarray(sorted(wbody_conn), dtype=np.int64),
        wbody_lake=np.asarray([wbody_conn[k] for k in sorted(wbody_conn)], dtype=np.int64),
        lake_ids=wb.index.values.astype(np.int64), lake_cols=np.array(wb.columns.tolist()), lake_table=wb.values.astype(np.float64),
        link_lake_lake=np.asarray(sorted(link_lake), dtype=np.int64),
        link_lake_seg=np.asarray([link_lake[k] for k in sorted(link_lake)], dtype=np.int64))
    return dict(nodes=len(nodes), lakes=len(wb), lake_flowpaths=len(wbody_conn), reaches=len(reaches),
                lake_reaches=sum(1 for r in reaches if set(r) & lakes))


def lowercolorado_v4_topobathy():
    """Surveyed cross sections of the coastal diffusive domain of the shipped hybrid configuration
    (test_AnA_V4_HYFeature.yaml:81-88: use_natl_xsections True, topobathy_domain domain/troute_test.parquet, domain file
    domain/coastal_domain_tw.yaml): read and completed by troute_b200.hyfeatures (read_topobathy, complete_topobathy) from
    the reference's parquet; the table after completion and the flowpaths without any usable section are stored."""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(OUT)), "t-route_b200"))
    sys.path.insert(0, os.path.dirname(OUT))
    from troute_b200 import hyfeatures as hy
    from troute_b200.routing import diffusive_domain
    import test_lowercolorado_hybrid as TH
    base = f"{REF}/test/LowerColorado_TX_v4"
    df = hy.read_flowpaths(f"{base}/domain/LowerColorado_NGEN_v201.gpkg")
    conn = hy.connections(df)
    dnd, _, _ = diffusive_domain.build_diffusive_network_data({TH.TW: {"headwater": list(TH.HEADS)}}, conn, df)
    links = dnd[TH.TW]["mainstem_segs"]
    raw = hy.read_topobathy(f"{base}/domain/troute_test.parquet", links)
    full, bad = hy.complete_topobathy(raw, links, df)
    np.savez_compressed(
        f"{OUT}/lowercolorado_v4_topobathy.npz", hy_id=full.index.values.astype(np.int64),
        relative_dist=full["relative_dist"].values.astype(np.float64), Z=full["Z"].values.astype(np.float64),
        roughness=full["roughness"].values.astype(np.float64), cs_id=full["cs_id"].values.astype(np.float64),
        bad_links=np.asarray(sorted(bad), dtype=np.int64), mainstem=df["mainstem"].reindex(links).values.astype(np.float64),
        links=np.asarray(links, dtype=np.int64))
    return dict(domain_links=len(links), with_data=int(raw.index.nunique()), after_fill=int(full.index.nunique()),
                bad=len(bad), rows=len(full), max_vertices=int(full.groupby(level=0).size().max()))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "lowercolorado":
        print("LowerColorado v4:", lowercolorado_v4())
        print("LowerColorado v4 with lakes:", lowercolorado_v4_lakes())
        print("LowerColorado v4 topobathy:", lowercolorado_v4_topobathy())
        sys.exit(0)
    k = mc_demo_kat()
    print("mc demo KAT:", k["single"]["expected"])
    a = mc_suite()
    print("mc suite:", a.shape, a.dtype)
    lp = levelpool_kats()
    print("level pool:", [(c["fixture"], len(c["inflow"]), c["expected_final_outflow"]) for c in lp["cases"]])
    print("simple_da:", simple_da_kat())
    g = nhd_graph()
    print("graph: reaches", {k: len(v) for k, v in g["fixture"]["reaches_bytw"].items()},
          "forest tw", len(g["forest300"]["reaches_bytw"]))
    print("LowerColorado v4:", lowercolorado_v4())
    print("LowerColorado v4 with lakes:", lowercolorado_v4_lakes())
    print("LowerColorado v4 topobathy:", lowercolorado_v4_topobathy())
