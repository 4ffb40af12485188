"""troute_b200.hyfeatures on a hand-made GeoPackage (an SQLite file with the two attribute tables of a NextGen
hydrofabric) and hand-made forcing CSVs: the reader reproduces what HYFeaturesNetwork.read_geopkg + preprocess_network
(HYFeaturesNetwork.py:33-107, :369-444) hand to the routing path.  CPU only, no reference tree needed."""
import os
import sqlite3

import numpy as np
import pytest

from troute_b200 import hyfeatures as hy


def _make_gpkg(path):
    con = sqlite3.connect(path)
    con.execute("CREATE TABLE flowpaths (fid INTEGER, geom BLOB, id TEXT, toid TEXT, mainstem REAL)")
    con.execute("CREATE TABLE flowpath_attributes (fid INTEGER, id TEXT, rl_gages TEXT, rl_NHDWaterbodyComID REAL, Qi REAL,"
                " MusK REAL, MusX REAL, n REAL, So REAL, ChSlp REAL, BtmWdth REAL, time REAL, Kchan REAL, nCC REAL,"
                " TopWdthCC REAL, TopWdth REAL, length_m REAL)")
    # 5 -> 3 -> 1 -> terminal nexus;  4 -> 3;  2 -> 1   (wb-X drains to nex-Y, nexus nex-Y feeds wb-Y)
    topo = {5: "nex-3", 4: "nex-3", 3: "nex-1", 2: "nex-1", 1: "tnx-1000000001"}
    for i, (k, to) in enumerate(sorted(topo.items(), reverse=True)):
        con.execute("INSERT INTO flowpaths VALUES (?, NULL, ?, ?, ?)", (i, f"wb-{k}", to, 77.0))
        con.execute("INSERT INTO flowpath_attributes VALUES (?, ?, NULL, NULL, 0, 3600, 0.2, ?, ?, ?, ?, 0, 0, ?, ?, ?, ?)",
                    (i, f"wb-{k}", 0.05 + 0.001 * k, 0.001 * k, 0.5, 2.0 * k, 0.1 + 0.001 * k, 30.0 * k, 10.0 * k, 1000.0 * k))
    con.commit(); con.close()


def test_read_flowpaths_and_connections(tmp_path):
    g = str(tmp_path / "toy.gpkg")
    _make_gpkg(g)
    df = hy.read_flowpaths(g)
    assert df.index.tolist() == [1, 2, 3, 4, 5]                     # numeric ids, sorted
    assert df.loc[5, "downstream"] == 3 and df.loc[1, "downstream"] == 1000000001
    assert df.attrs["terminal_rows"] == 1
    assert df.loc[3, "dx"] == 3000.0 and df.loc[3, "bw"] == 6.0 and df.loc[3, "tw"] == 30.0 and df.loc[3, "twcc"] == 90.0
    assert (df["alt"] == 1.0).all() and "gages" not in df.columns
    conn = hy.connections(df)
    assert conn == {1: [], 2: [1], 3: [1], 4: [3], 5: [3]}
    p = hy.param_frame(df, 300.0)
    assert p.columns.tolist() == ["dt", "bw", "tw", "twcc", "dx", "n", "ncc", "cs", "s0", "alt"] and (p["dt"] == 300.0).all()
    assert p.values.dtype == np.float32


def test_read_channel_forcing(tmp_path):
    d = tmp_path / "forcing"; d.mkdir()
    for h, vals in (("202304010000", {1: 0.5, 3: 1.5}), ("202304010100", {1: 0.25, 2: 2.0, 3: 1.0})):
        with open(d / f"{h}.CHRTOUT_DOMAIN1.csv", "w") as f:
            f.write(f"feature_id,{h}\n" + "".join(f"{k},{v}\n" for k, v in vals.items()))
    q = hy.read_channel_forcing(str(d), index=[1, 2, 3, 4])
    assert q.shape == (4, 2) and q.values.dtype == np.float32
    assert q.loc[2].tolist() == [0.0, 2.0] and q.loc[4].tolist() == [0.0, 0.0] and q.loc[3].tolist() == [1.5, 1.0]
    with pytest.raises(FileNotFoundError):
        hy.read_channel_forcing(str(tmp_path / "nothing"))


def _add_lakes(path):
    """two lakes: lake 900 covers flowpaths 4 and 3 (drains to 1: consistent), lake 901 covers 5 and 2 -- 5 drains into
    lake 900, 2 drains to flowpath 1: two different downstream nodes, the hydrofabric defect bandaid() drops"""
    con = sqlite3.connect(path)
    con.execute("CREATE TABLE lakes (fid INTEGER, geom BLOB, id TEXT, toid TEXT, hl_id REAL, hl_reference TEXT, hl_link TEXT,"
                " hl_uri TEXT, Dam_Length REAL, ifd REAL, LkArea REAL, LkMxE REAL, OrificeA REAL, OrificeC REAL,"
                " OrificeE REAL, time REAL, WeirC REAL, WeirE REAL, WeirL REAL)")
    rows = [(1, "nex-3", "wb-3", "900", 0.9, 4.5, 120.0, 1.0, 0.1, 100.0, 0.4, 118.0, 10.0),
            (2, "nex-2", "wb-2", "901.0", 0.9, 2.5, 220.0, 1.0, 0.1, 200.0, 0.4, 218.0, 10.0),
            (3, "nex-2", "wb-2", "901.0", 0.9, 2.5, 220.0, 1.0, 0.1, 200.0, 0.4, 218.0, 10.0),      # duplicate row
            (4, "nex-9", "wb-9", "902", 0.9, None, 320.0, 1.0, 0.1, 300.0, 0.4, 318.0, 10.0)]        # missing parameter
    for fid, nid, toid, link, ifd, area, mxe, oa, oc, oe, wc, we, wl in rows:
        con.execute("INSERT INTO lakes VALUES (?, NULL, ?, ?, 1, 'WBOut', ?, 'x', 10.0, ?, ?, ?, ?, ?, ?, 0, ?, ?, ?)",
                    (fid, nid, toid, link, ifd, area, mxe, oa, oc, oe, wc, we, wl))
    for k, lake in ((4, "900"), (3, "900"), (5, "901,555"), (2, "901")):
        con.execute("UPDATE flowpath_attributes SET rl_NHDWaterbodyComID = ? WHERE id = ?", (lake, f"wb-{k}"))
    con.commit(); con.close()


def test_read_lakes_and_waterbody_connections(tmp_path):
    g = str(tmp_path / "toy.gpkg")
    _make_gpkg(g)
    _add_lakes(g)
    df = hy.read_flowpaths(g)
    wb = hy.read_lakes(g)
    assert wb.index.tolist() == [900, 901] and wb.index.name == "lake_id"          # duplicate and incomplete rows dropped
    assert wb.columns.tolist() == ["ifd", "LkArea", "LkMxE", "OrificeA", "OrificeC", "OrificeE", "WeirC", "WeirE", "WeirL", "id"]
    assert wb.loc[900, "LkArea"] == 4.5 and wb.loc[901, "WeirE"] == 218.0 and wb.loc[900, "id"] == 3
    wc = hy.waterbody_connections(df, wb)
    assert wc == {2: 901, 3: 900, 4: 900, 5: 901}                                  # 555 is not a lake of the table
    wb2, wc2 = hy.drop_inconsistent_lakes(df, wb, wc)
    assert wb2.index.tolist() == [900] and wc2 == {3: 900, 4: 900}                 # lake 901 would have two outlets


def test_topobathy_read_and_completion(tmp_path):
    """read_topobathy + complete_topobathy on a hand-made cross-section table: 5 -> 3 -> 1 is mainstem 77 (2 and 4 are side
    arms on other mainstems).  1 has no section and borrows the most downstream one (largest cs_id) of 3; 4 has nothing
    upstream on its mainstem -> bad link; 5 has two sections, the one with the smaller cs_id is kept; incomplete rows are
    dropped on reading."""
    import pandas as pd
    df = pd.DataFrame({"downstream": [1000000001, 1, 1, 3, 3], "mainstem": [77.0, 12.0, 77.0, 13.0, 77.0]},
                      index=pd.Index([1, 2, 3, 4, 5], name="key"))
    rows = []
    for hy_id, cs, pts in (("wb-3", 1.0, 3), ("wb-3", 2.0, 4), ("wb-5", 4.0, 3), ("wb-5", 2.0, 5), ("wb-2", 1.0, 3), ("wb-9", 1.0, 3)):
        for k in range(pts):
            rows.append(dict(hy_id=hy_id, cs_id=cs, pt_id=float(k + 1), X=0.0, Y=0.0, Z=100.0 + cs - 0.5 * min(k, pts - 1 - k),
                             Z_source="x", roughness=0.05 + 0.01 * cs, relative_dist=10.0 * k, pt_measure=0.0))
    rows.append(dict(hy_id="wb-3", cs_id=1.0, pt_id=9.0, X=0.0, Y=0.0, Z=None, Z_source="x", roughness=0.06, relative_dist=99.0,
                     pt_measure=0.0))                                               # incomplete: dropped
    path = str(tmp_path / "xs.parquet")
    pd.DataFrame(rows).to_parquet(path)
    links = [1, 3, 4, 5]
    tb = hy.read_topobathy(path, links)
    assert sorted(tb.index.unique().tolist()) == [3, 5] and len(tb) == 3 + 4 + 3 + 5          # wb-2 / wb-9 not asked for
    assert tb.columns.tolist() == ["relative_dist", "Z", "roughness", "cs_id"]
    full, bad = hy.complete_topobathy(tb, links, df)
    assert bad == [4]
    assert sorted(full.index.unique().tolist()) == [1, 3, 5]
    assert len(full.loc[3]) == 3 and (full.loc[3, "cs_id"] == 1.0).all()                      # smallest cs_id of 3
    assert len(full.loc[5]) == 5 and (full.loc[5, "cs_id"] == 2.0).all()                      # smallest cs_id of 5
    borrowed = full.loc[1]
    assert len(borrowed) == 4 and (borrowed["cs_id"] == 1.0).all()                            # section 2 of flowpath 3, relabelled
    assert borrowed["roughness"].iloc[0] == pytest.approx(0.07) and borrowed["relative_dist"].tolist() == [0.0, 10.0, 20.0, 30.0]


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="pins complete_topobathy against the reference tree, present only in the build container")
def test_complete_topobathy_equals_the_reference_fill_in_on_the_real_domain():
    """complete_topobathy against the reference's own _fill_in_missing_topo_data (AbstractRouting.py:503-526, compiled out
    of its module, which imports xarray) and the selection rules around it (:404-426), on the real LowerColorado hydrofabric
    and cross-section table: same borrowed sections for the same flowpaths, same list of flowpaths without any."""
    import ast
    import pandas as pd
    path = f"{REF}/src/troute-network/troute/AbstractRouting.py"
    tree = ast.parse(open(path).read())
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "_fill_in_missing_topo_data")
    ns = {"pd": pd, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    fill = ns["_fill_in_missing_topo_data"]

    base = f"{REF}/test/LowerColorado_TX_v4/domain"
    df = hy.read_flowpaths(f"{base}/LowerColorado_NGEN_v201.gpkg")
    have = pd.read_parquet(f"{base}/troute_test.parquet", columns=["hy_id"]).dropna()["hy_id"].unique()
    have = sorted(int(str(x).split("-")[-1]) for x in have)
    have = [k for k in have if k in df.index]
    # the coastal diffusive domain of the shipped configuration (787 flowpaths, recorded in the fixture) plus every flowpath
    # with data and what lies within 3 flowpaths downstream of one
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lowercolorado_v4_topobathy.npz")
    links = set(np.load(gold)["links"].tolist()) | set(have)
    for k in have:
        cur = k
        for _ in range(3):
            cur = int(df.loc[cur, "downstream"])
            if cur not in df.index:
                break
            links.add(cur)
    links = sorted(links)
    raw = hy.read_topobathy(f"{base}/troute_test.parquet", links)
    full, bad = hy.complete_topobathy(raw, links, df)
    # the reference, step by step (:404-426)
    frame = df.reset_index().rename(columns={"index": "key"}).set_index("key")
    missing = sorted(set(links) - set(raw.index))
    pieces = [fill(k, frame, raw) for k in missing]
    new_topo = pd.concat([p for p in pieces if not p.empty]) if any(not p.empty for p in pieces) else pd.DataFrame()
    ref_bad = sorted(set(missing) - set(new_topo.index))
    assert len(missing) > 100 and len(new_topo.index.unique()) >= 3 and len(ref_bad) > 100      # all three cases occur
    assert bad == ref_bad
    both = pd.concat([raw, new_topo])
    r = both.reset_index()
    r = r[r["cs_id"] == r.groupby("hy_id")["cs_id"].transform("min")].set_index("hy_id")
    cols = ["relative_dist", "Z", "roughness", "cs_id"]
    a = full[cols].sort_index(kind="stable")
    b = r[cols].sort_index(kind="stable")
    assert a.index.tolist() == b.index.tolist()
    assert np.array_equal(a.to_numpy(), b.to_numpy())


@pytest.mark.skipif(not os.path.isdir(REF), reason="pins the lake readers against the reference tree, present only in the build container")
def test_lake_readers_equal_the_reference_preprocessing_on_the_real_hydrofabric():
    """read_lakes / waterbody_connections / drop_inconsistent_lakes against HYFeaturesNetwork.preprocess_waterbodies and
    bandaid themselves (HYFeaturesNetwork.py:456-560, :819-856), compiled out of their module (which imports geopandas /
    xarray) and run on a stand-in object with the `lakes` and `nexus` layers read through sqlite3: same waterbody table, same
    flowpath -> lake map; and the collapsed graph / outlet crosswalk the reference derives from them equal the committed
    fixture tests/golden/lowercolorado_v4_lakes.npz."""
    import ast
    import importlib.util
    import sys
    import types
    import pandas as pd
    src = f"{REF}/src/troute-network/troute"
    if "toolz" not in sys.modules:
        tz = types.ModuleType("toolz")
        tz.pluck = lambda ind, seqs: (s_[ind] for s_ in seqs)
        sys.modules["toolz"] = tz
    spec = importlib.util.spec_from_file_location("ref_nhd_network_wb", f"{src}/nhd_network.py")
    nn = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(nn)
    tree = ast.parse(open(f"{src}/HYFeaturesNetwork.py").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "HYFeaturesNetwork")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name in ("preprocess_waterbodies", "bandaid")]
    ns = {"pd": pd, "np": np, "replace_waterbodies_connections": nn.replace_waterbodies_connections}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "reference", "exec"), ns)

    gpkg = f"{REF}/test/LowerColorado_TX_v4/domain/LowerColorado_NGEN_v201.gpkg"
    con = sqlite3.connect(f"file:{gpkg}?mode=ro", uri=True)
    try:
        lakes = pd.read_sql_query("SELECT * FROM lakes", con).drop(columns=["geom"])
        nexus = pd.read_sql_query("SELECT id, toid, hl_uri FROM nexus", con)
    finally:
        con.close()
    df = hy.read_flowpaths(gpkg)

    class Stand:
        waterbody_dataframe = property(lambda self: self._waterbody_df)
        dataframe = property(lambda self: self._dataframe)
        connections = property(lambda self: self._connections)
        waterbody_connections = property(lambda self: self._waterbody_connections)
    me = Stand()
    me._dataframe = df.copy()
    me._dataframe.index.name = "key"
    me._connections = {k: list(v) for k, v in hy.connections(df).items()}
    me.waterbody_parameters = {"break_network_at_waterbodies": True}
    me.output_parameters = {}
    me.bandaid = types.MethodType(ns["bandaid"], me)
    ns["preprocess_waterbodies"](me, lakes, nexus)

    wb = hy.read_lakes(gpkg)
    wc = hy.waterbody_connections(df, wb)
    wb, wc = hy.drop_inconsistent_lakes(df, wb, wc)
    assert me._waterbody_df.index.tolist() == wb.index.tolist() and len(wb) == 19
    assert np.array_equal(me._waterbody_df[wb.columns.tolist()].to_numpy(dtype=float), wb.to_numpy(dtype=float))
    assert me._waterbody_connections == wc and len(wc) == 248
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lowercolorado_v4_lakes.npz"))
    assert dict(zip(z["link_lake_lake"].tolist(), z["link_lake_seg"].tolist())) == {int(k): int(v) for k, v in me._link_lake_crosswalk.items()}
    fixture = {int(k): ([int(d)] if d >= 0 else []) for k, d in zip(z["nodes"], z["downstream"])}
    ref_conn = {int(k): [int(x) for x in v] for k, v in me._connections.items()}
    assert {k: v for k, v in fixture.items() if k in ref_conn} == ref_conn
    assert all(v == [] for k, v in fixture.items() if k not in ref_conn)                      # the three phantom outlets


@pytest.mark.skipif(not os.path.isdir(REF), reason="needs /root/reference")
def test_gage_reader_equals_the_reference_method():
    """hyfeatures.read_gages == the statements of HYFeaturesNetwork.preprocess_data_assimilation that build `self._gages`
    (:606-637; the lake-gage crosswalks that follow use a pandas < 2 groupby signature and are outside the routing path),
    compiled out of the reference module (which imports geopandas / xarray and cannot be imported here) and run on a
    stand-in object with the `network` layer read through sqlite3: the same {flowpath id: USGS gage id} dictionary on the
    real hydrofabric (86 gages)."""
    import ast
    import sqlite3
    import types
    import pandas as pd
    src = f"{REF}/src/troute-network/troute"
    tree = ast.parse(open(f"{src}/HYFeaturesNetwork.py").read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "HYFeaturesNetwork")
    fns = [n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "preprocess_data_assimilation"]
    branch = fns[0].body[0]                                                   # if not network.empty:
    assert isinstance(branch, ast.If)
    last = next(i for i, st in enumerate(branch.body) if isinstance(st, ast.Assign) and ast.unparse(st.targets[0]) == "self._gages")
    branch.body = branch.body[:last + 1]
    ns = {"pd": pd, "np": np}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "reference", "exec"), ns)
    gpkg = f"{REF}/test/LowerColorado_TX_v4/domain/LowerColorado_NGEN_v201.gpkg"
    con = sqlite3.connect(f"file:{gpkg}?mode=ro", uri=True)
    try:
        network = pd.read_sql_query("SELECT * FROM network", con)
    finally:
        con.close()
    df = hy.read_flowpaths(gpkg)
    wb = hy.read_lakes(gpkg)
    wc = hy.waterbody_connections(df, wb)

    class Stand:
        waterbody_connections = property(lambda self: self._wc)
    me = Stand()
    me._wc = wc
    me.data_assimilation_parameters = {}
    me._waterbody_types_df = pd.DataFrame()
    ns["preprocess_data_assimilation"](me, network)
    want = {int(k): str(v) for k, v in me._gages["gages"].items()}
    got = hy.read_gages(gpkg)
    assert got == want and len(got) == 86
    assert all(k in df.index for k in got)                                   # every gage sits on a flowpath of the domain
