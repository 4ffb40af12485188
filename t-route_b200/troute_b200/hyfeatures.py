"""Minimal HYFeatures (NextGen hydrofabric v2.x GeoPackage) and channel-forcing readers on the standard library.

The reference reads the GeoPackage with geopandas/fiona (troute/HYFeaturesNetwork.py:33-107 read_geopkg) and renames /
re-indexes it in preprocess_network (:369-444).  A GeoPackage is an SQLite file, so the two attribute tables the
routing path needs can be read with `sqlite3` -- no geometry, no GDAL.  This covers the MC-only plumbing configuration
(BASELINE config 0: test/LowerColorado_TX_v4) and its level-pool reservoirs (the `lakes` layer: read_lakes,
waterbody_connections, drop_inconsistent_lakes -- what preprocess_waterbodies :456-526 and bandaid :819-856 derive) and
the surveyed cross sections of a diffusive domain (read_topobathy, complete_topobathy -- AbstractRouting.py:57-82, :390-428,
:503-526; parquet through pandas / pyarrow) and the stream gages of streamflow data assimilation (read_gages -- the
`network` layer, preprocess_data_assimilation :606-637); coastal boundaries are not read here.
"""
import glob
import os
import sqlite3

import numpy as np
import pandas as pd

# network_topology_parameters.supernetwork_parameters.columns of the shipped v4 configs
# (test/LowerColorado_TX_v4/test_AnA_V4_HYFeature.yaml:13-29): standard name -> hydrofabric column
DEFAULT_COLUMNS = {"key": "id", "downstream": "toid", "dx": "length_m", "n": "n", "ncc": "nCC", "s0": "So",
                   "bw": "BtmWdth", "waterbody": "rl_NHDWaterbodyComID", "gages": "rl_gages", "tw": "TopWdth",
                   "twcc": "TopWdthCC", "musk": "MusK", "musx": "MusX", "cs": "ChSlp", "alt": "alt",
                   "mainstem": "mainstem"}


def _numeric_id(s):
    """'wb-2420800' / 'nex-2420801' / 'tnx-1000000123' -> 2420800 (HYFeaturesNetwork.numeric_id)."""
    return int(str(s).split("-")[-1])


def read_flowpaths(gpkg_path, columns=None):
    """flowpaths JOIN flowpath_attributes ON id, renamed to the standard parameter names, indexed by the numeric
    flowpath id and sorted (read_geopkg :92-98 + preprocess_network :373-408).  The `downstream` column holds the
    numeric id of the nexus the flowpath drains to, which is the id of the flowpath below that nexus; ids that are not
    in the index (terminal nexuses, 'tnx-*') are terminal codes."""
    columns = dict(DEFAULT_COLUMNS if columns is None else columns)
    con = sqlite3.connect(f"file:{gpkg_path}?mode=ro", uri=True)
    try:
        fp = pd.read_sql_query('SELECT id, toid, mainstem FROM flowpaths', con)
        at = pd.read_sql_query('SELECT * FROM flowpath_attributes', con)
    finally:
        con.close()
    if "link" in at.columns:
        at = at.rename(columns={"link": "id"})
    df = pd.merge(fp, at.drop(columns=[c for c in ("fid",) if c in at.columns]), on="id", how="inner")
    keep = [v for v in columns.values() if v in df.columns]
    df = df[keep].rename(columns={v: k for k, v in columns.items()})
    terminal = df["downstream"].astype(str).str.startswith("tnx")
    df["key"] = df["key"].map(_numeric_id)
    df["downstream"] = df["downstream"].map(_numeric_id)
    df = df.set_index("key").sort_index()
    if "alt" not in df.columns:
        df["alt"] = 1.0                                                     # :404-405
    if "gages" in df.columns:
        df = df.drop(columns="gages")
    df.attrs["terminal_rows"] = int(terminal.sum())
    return df


def connections(df):
    """nhd_network.extract_connections (nhd_network.py:26-53): id -> [downstream id] ([] at a terminal)."""
    index = set(df.index.tolist())
    return {int(k): ([int(d)] if int(d) in index else []) for k, d in df["downstream"].items()}


def read_channel_forcing(folder, pattern="*.CHRTOUT_DOMAIN1.csv", index=None):
    """One CSV per forcing hour, columns (feature_id, <timestamp>): -> DataFrame indexed by segment id, one column per
    file in name order (the qlat frame the reference assembles, nhd_io.get_ql_from_csv)."""
    files = sorted(glob.glob(os.path.join(folder, pattern)))
    if not files:
        raise FileNotFoundError(f"no forcing files matching {pattern} in {folder}")
    frames = [pd.read_csv(f, index_col=0) for f in files]
    q = pd.concat(frames, axis=1)
    q.index = q.index.astype("int64")
    if index is not None:
        q = q.reindex(index).fillna(0.0)
    return q.astype("float32")


def param_frame(df, dt):
    """The parameter table compute_nhd_routing_v02 slices (compute.py:1443-1446 column set), float32."""
    p = df[["bw", "tw", "twcc", "dx", "n", "ncc", "cs", "s0", "alt"]].copy()
    p.insert(0, "dt", float(dt))
    return p.astype("float32")


# ---- level-pool reservoirs of the hydrofabric -----------------------------------------------------------------------
LAKE_COLUMNS = ["ifd", "LkArea", "LkMxE", "OrificeA", "OrificeC", "OrificeE", "WeirC", "WeirE", "WeirL"]


def read_lakes(gpkg_path):
    """The `lakes` layer as the waterbody table of the routing path: indexed by lake id (`hl_link`), the nine level-pool
    parameters plus `id` = numeric id of the flowpath the lake outlet belongs to; duplicates and lakes with a missing
    parameter dropped, sorted (HYFeaturesNetwork.preprocess_waterbodies :459-472)."""
    con = sqlite3.connect(f"file:{gpkg_path}?mode=ro", uri=True)
    try:
        lk = pd.read_sql_query("SELECT hl_link, id, " + ", ".join(LAKE_COLUMNS) + " FROM lakes", con)
    finally:
        con.close()
    lk = lk.dropna(subset=["hl_link"])
    out = pd.DataFrame({c: lk[c].astype("float64") for c in LAKE_COLUMNS})
    out["id"] = lk["id"].map(_numeric_id).astype("int64")
    out.index = pd.Index(lk["hl_link"].astype(float).astype("int64"), name="lake_id")
    out = out[~out.index.duplicated(keep="first")].dropna().sort_index()
    return out[["ifd", "LkArea", "LkMxE", "OrificeA", "OrificeC", "OrificeE", "WeirC", "WeirE", "WeirL", "id"]]


def waterbody_connections(df, waterbodies_df):
    """{flowpath id: lake id} for every flowpath whose `waterbody` attribute (rl_NHDWaterbodyComID, possibly a
    comma-separated list) names a lake of the table (preprocess_waterbodies :483-515).  When a flowpath lists several
    known lakes the last one wins, as the reference's merge + to_dict does."""
    known = set(int(x) for x in waterbodies_df.index)
    out = {}
    col = df["waterbody"] if "waterbody" in df.columns else pd.Series(dtype=object)
    for key, val in col.dropna().items():
        for tok in str(val).split(","):
            tok = tok.strip()
            if not tok:
                continue
            try:
                lake = int(float(tok))
            except ValueError:
                continue
            if lake in known:
                out[int(key)] = lake
    return out


def drop_inconsistent_lakes(df, waterbodies_df, wbody_conn):
    """Lakes whose flowpaths, collapsed into one node, would drain to more than one downstream node are not simulated as
    reservoirs -- their flowpaths stay ordinary Muskingum-Cunge segments (HYFeaturesNetwork.bandaid :819-849: the
    hydrofabric misses some of the flowpaths under such a lake).  Returns (waterbodies_df, wbody_conn) without them."""
    index = set(int(k) for k in df.index)
    outlets = {}
    for key, down in df["downstream"].items():
        key, down = int(key), int(down)
        src = wbody_conn.get(key, key)
        dst = wbody_conn.get(down, down) if down in index else down
        if src != dst and key in wbody_conn:
            outlets.setdefault(src, set()).add(dst)
    bad = sorted(lake for lake, dsts in outlets.items() if len(dsts) > 1)
    if not bad:
        return waterbodies_df, dict(wbody_conn)
    keep = waterbodies_df.drop(index=[b for b in bad if b in waterbodies_df.index])
    return keep, {k: v for k, v in wbody_conn.items() if v not in set(bad)}


# ---- surveyed ("natural") cross sections of a diffusive domain ------------------------------------------------------
def read_topobathy(parquet_path, links):
    """Cross-section vertices of the given flowpaths from a hydrofabric cross-section table: rows (relative_dist, Z,
    roughness, cs_id) indexed by the numeric flowpath id, incomplete rows dropped (AbstractRouting.read_parquet :57-82 and
    the index handling of :396-400)."""
    want = ["wb-" + str(int(s)) for s in links]
    df = pd.read_parquet(parquet_path, columns=["hy_id", "relative_dist", "Z", "roughness", "cs_id"],
                         filters=[("hy_id", "in", want)]).dropna()
    df["hy_id"] = df["hy_id"].map(lambda x: int(str(x).split("-")[-1]))
    return df.set_index("hy_id")


def complete_topobathy(topobathy_df, links, dataframe):
    """One cross section per flowpath of the domain (AbstractRouting.py:404-426, _fill_in_missing_topo_data :503-526).

    A flowpath without data borrows the most downstream section (largest cs_id, relabelled 1) of the nearest flowpath
    upstream of it ON THE SAME MAINSTEM that has data; flowpaths for which that fails are returned as `bad_links` -- the
    domain builder ends the diffusive domain below them.  Where a flowpath has several sections the one with the smallest
    cs_id is kept.  `dataframe`: flowpath table indexed by id with `downstream` and `mainstem` columns (read_flowpaths)."""
    have = set(int(x) for x in topobathy_df.index.unique())
    up_on_mainstem = {}
    for key, (down, stem) in dataframe[["downstream", "mainstem"]].iterrows():
        up_on_mainstem.setdefault((int(down), stem), []).append(int(key))
    pieces, bad = [topobathy_df.reset_index()], []
    for key in sorted(set(int(x) for x in links) - have):
        stem = dataframe.loc[key, "mainstem"]
        donor, cur, seen = None, key, set()
        while donor is None and (cur, stem) in up_on_mainstem and cur not in seen:
            seen.add(cur)
            ups = up_on_mainstem[(cur, stem)]
            donor = next((u for u in ups if u in have), None)
            cur = ups[0]
        if donor is None:
            bad.append(key)
            continue
        rows = topobathy_df.loc[[donor]].reset_index()
        rows = rows[rows["cs_id"] == rows["cs_id"].max()].copy()
        rows["cs_id"] = 1.0
        rows["hy_id"] = key
        pieces.append(rows)
    df = pd.concat(pieces, ignore_index=True)
    df = df[df["cs_id"] == df.groupby("hy_id")["cs_id"].transform("min")]
    return df.set_index("hy_id"), bad


def read_gages(gpkg_path):
    """{flowpath id: USGS gage id} for streamflow nudging -- what HYFeaturesNetwork.preprocess_data_assimilation (:606-637)
    leaves in `network.gages["gages"]`: the hydrolocations of the `network` layer whose `hl_uri` is `Gages-<id>` or `NID-<id>`
    (several ids separated by blanks are several gages), numeric ids only (USGS; the alphanumeric ones are USACE / NID
    reservoir gages), and when a gage is listed on more than one flowpath the one with the largest `hydroseq` -- the
    reference's "furthest downstream" rule (`sort_values('hydroseq').drop_duplicates(keep='last')`) -- wins.  The gage ids
    are the strings of the hydrofabric ('08121000'), the keys the numeric flowpath ids."""
    con = sqlite3.connect(f"file:{gpkg_path}?mode=ro", uri=True)
    try:
        net = pd.read_sql_query("SELECT id, hl_uri, hydroseq FROM network", con)
    finally:
        con.close()
    g = net.drop_duplicates()
    g = g[~g["hl_uri"].isnull() & ~g["hydroseq"].isnull()]
    if g.empty:
        return {}
    kind = g["hl_uri"].str.split("-", n=1).str[0]
    g = g[kind.isin(["Gages", "NID"])]
    g = pd.DataFrame({"id": g["id"].map(_numeric_id).to_numpy(), "hydroseq": g["hydroseq"].to_numpy(),
                      "value": g["hl_uri"].str.split("-", n=1).str[1].str.split(" ").to_numpy()})
    g = g.explode("value")
    g = g[g["value"].str.isnumeric()]
    g = g.sort_values("hydroseq", kind="stable").drop_duplicates(["value"], keep="last")
    return dict(zip(g["id"].astype(int).tolist(), g["value"].tolist()))
