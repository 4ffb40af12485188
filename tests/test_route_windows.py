"""State hand-off between consecutive routing windows (troute_b200.nwm_routing.new_q0 / update_waterbody_water_elevation /
new_lastobs / route_windows, mirroring AbstractNetwork.py:177-198, DataAssimilation.py:1506-1551 and the loop of
nwm_routing.__main__:150-330): W windows routed one after the other, each started from the state derived from the previous
window's results, equal ONE call over all steps -- flows, depths, reservoir states and the nudging of gages whose last
observation lies in an earlier window.  CPU: the oracle stands in for the device call; GPU: the product's
compute_network_structured with its cached device network."""
import numpy as np
import pandas as pd
import pytest

import helpers as H
import test_gpu_api as A


def _windowed(fn, c, gages, W):
    """route c in W windows through route_windows; returns ([n, 3*nsteps] in id order, final lastobs_df)"""
    from troute_b200 import nwm_routing
    n_w = c["nsteps"] // W
    qcols = n_w // c["qts"]
    ids = c["ids"]
    q0 = pd.DataFrame(c["q0"], index=ids, columns=["qu0", "qd0", "h0"])
    wb_cols = ["LkArea", "LkMxE", "OrificeA", "OrificeC", "OrificeE", "WeirC", "WeirE", "WeirL", "ifd", "qd0", "h0"]
    waterbodies_df = pd.DataFrame(np.array(c["wbody"], dtype=np.float64), index=c["lake_numbers"], columns=wb_cols)
    lastobs_df = pd.DataFrame()
    if gages is not None:
        gage_ids = ids[gages["usgs_positions"]]
        lastobs_df = pd.DataFrame({"time_since_lastobs": gages["time_since_lastobs_init"],
                                   "lastobs_discharge": gages["lastobs_values_init"]}, index=gage_ids)

    def route_window(w, q0_df, wb_df, lo_df):
        part = dict(c)
        part["nsteps"] = n_w
        part["qlat"] = c["qlat"][:, w * qcols:(w + 1) * qcols]
        part["q0"] = q0_df.loc[ids].to_numpy(dtype=np.float32)
        part["wbody"] = wb_df.loc[c["lake_numbers"]].to_numpy(dtype=np.float64)
        g = None
        if gages is not None:
            g = dict(gages)
            # column j of a call is the observation at j routing periods after its start (column 0 replaces the initial
            # flow, mc_reach.pyx:403-411; step t reads column t, simple_da.pyx:47): a window needs n_w + 1 columns
            g["usgs_values"] = np.ascontiguousarray(gages["usgs_values"][:, w * n_w:(w + 1) * n_w + 1])
            lo = lo_df.loc[gage_ids]
            g["lastobs_values_init"] = lo["lastobs_discharge"].to_numpy(dtype=np.float32)
            g["time_since_lastobs_init"] = lo["time_since_lastobs"].to_numpy(dtype=np.float32)
        return [A._call(fn, part, gages=g)], None

    per_window, q0, waterbodies_df, lastobs_df = nwm_routing.route_windows(route_window, range(W), q0, waterbodies_df,
                                                                           lastobs_df, 300.0, n_w)
    pieces = []
    for results in per_window:
        (r,) = results
        pieces.append(r[1][np.argsort(r[0])])
    return np.concatenate(pieces, axis=1), lastobs_df, q0


def _case_with_gages():
    c = A._reference_style_case(n=3000, seed=13, n_lp=8, nsteps=48)
    g = A._gage_inputs(c, n_gages=60, seed=5, obs_steps=48)
    # observations stop after step 20 (gage_maxtimestep is the window's column count, so later columns are simply NaN):
    # windows 2.. decay from a last observation made in window 1
    g["usgs_values"][:, 20:] = np.nan
    # last observations carried in from before the run: whole routing periods ago, so that (t * dt - time) is the same
    # float32 number whichever window t is counted from
    since = g["time_since_lastobs_init"]
    g["time_since_lastobs_init"] = np.where(np.isnan(since), since, -300.0 * np.round(-since / 300.0)).astype(np.float32)
    return c, g


def test_state_handoff_helpers():
    from troute_b200 import nwm_routing
    fvd = np.arange(2 * 9, dtype=np.float32).reshape(2, 9)                  # two segments, three steps of (q, v, d)
    res = [(np.array([7, 3]), fvd, 0, (np.array([7]), np.array([600.0], np.float32), np.array([2.5], np.float32)))]
    q0 = nwm_routing.new_q0(res)
    assert q0.columns.tolist() == ["qu0", "qd0", "h0"] and q0.index.tolist() == [7, 3]
    assert q0.loc[7].tolist() == [6.0, 6.0, 8.0] and q0.loc[3].tolist() == [15.0, 15.0, 17.0]
    wb = pd.DataFrame({"LkArea": [1.0], "qd0": [0.0], "h0": [0.0]}, index=[3])
    nwm_routing.update_waterbody_water_elevation(wb, q0)
    assert wb.loc[3, "qd0"] == 15.0 and wb.loc[3, "h0"] == 17.0 and wb.loc[3, "LkArea"] == 1.0
    lo = nwm_routing.new_lastobs(res, 900.0)
    assert lo.columns.tolist() == ["time_since_lastobs", "lastobs_discharge"]
    assert lo.loc[7].tolist() == [-300.0, 2.5]                              # observed 300 s before the next window starts


@pytest.mark.parametrize("W", [2, 4])
def test_windows_equal_one_call_on_the_oracle(oracle, W):
    c, g = _case_with_gages()
    for gages in (None, g):
        full = A._call(oracle.compute_network_structured, c, gages=gages)
        got, lastobs_df, _ = _windowed(oracle.compute_network_structured, c, gages, W)
        H.assert_bit_equal(got, full[1][np.argsort(full[0])], f"{W} windows vs one call (gages: {gages is not None})")
        if gages is not None:
            # last-observation state after the last window == after the single call (its times are already relative to
            # the end of the run, mc_reach.pyx:822-836; new_lastobs re-bases the window's the same way)
            H.assert_bit_equal(lastobs_df["lastobs_discharge"].to_numpy(np.float32), full[3][2], "lastobs values")
            H.assert_bit_equal(lastobs_df["time_since_lastobs"].to_numpy(np.float32),
                               np.asarray(full[3][1], np.float32) - np.float32(300.0 * c["nsteps"]), "lastobs times")


@pytest.mark.gpu
@pytest.mark.parametrize("W", [2, 4])
def test_windows_equal_one_call_on_the_device(oracle, W):
    from troute_b200.routing.fast_reach import mc_reach
    c, g = _case_with_gages()
    try:
        for gages in (None, g):
            full = A._call(oracle.compute_network_structured, c, gages=gages)
            got, _, _ = _windowed(mc_reach.compute_network_structured, c, gages, W)
            H.assert_bit_equal(got, full[1][np.argsort(full[0])], f"device: {W} windows vs the oracle's single call")
        assert len(mc_reach._NET_CACHE) == 1
    finally:
        mc_reach.clear_network_cache()


REF = "/root/reference/src/troute-network/troute"


def _reference_function(path, name, cls=None):
    """compile ONE function out of a reference module (the modules themselves import xarray / netCDF4, absent here)"""
    import ast
    tree = ast.parse(open(path).read())
    body = tree.body
    if cls is not None:
        body = next(n for n in body if isinstance(n, ast.ClassDef) and n.name == cls).body
    fn = next(n for n in body if isinstance(n, ast.FunctionDef) and n.name == name)
    ns = {"pd": pd, "np": np}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), path, "exec"), ns)
    return ns[name]


@pytest.mark.skipif(not __import__("os").path.isdir(REF), reason="pins the helpers against the reference tree, present only in the build container")
def test_state_handoff_helpers_equal_the_reference_functions():
    """new_q0 / new_lastobs / update_waterbody_water_elevation against the reference's own source (AbstractNetwork.new_q0,
    update_waterbody_water_elevation :177-198; DataAssimilation.new_lastobs :1506-1551), compiled function by function."""
    import warnings
    from troute_b200 import nwm_routing
    rng = np.random.default_rng(5)
    results = []
    for k, n in enumerate((7, 4)):
        ids = np.arange(100 * k + 1, 100 * k + 1 + n)
        fvd = rng.uniform(0, 5, (n, 9)).astype(np.float32)
        gi = ids[::2]
        results.append((ids, fvd, 0, (gi, rng.uniform(0, 900, gi.size).astype(np.float32), rng.uniform(0, 9, gi.size).astype(np.float32))))
    ref_new_q0 = _reference_function(f"{REF}/AbstractNetwork.py", "new_q0", cls="AbstractNetwork")
    ref_update = _reference_function(f"{REF}/AbstractNetwork.py", "update_waterbody_water_elevation", cls="AbstractNetwork")
    ref_lastobs = _reference_function(f"{REF}/DataAssimilation.py", "new_lastobs")

    class Net:
        pass
    net = Net()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")                               # pandas deprecates the reference's copy=False
        q0_ref = ref_new_q0(net, results)
        lo_ref = ref_lastobs(results, 900.0)
    q0 = nwm_routing.new_q0(results)
    pd.testing.assert_frame_equal(q0, q0_ref, check_dtype=False)
    lo = nwm_routing.new_lastobs(results, 900.0)
    pd.testing.assert_frame_equal(lo, lo_ref, check_dtype=False, check_exact=True)
    wb = pd.DataFrame({"LkArea": [1.0, 2.0, 3.0], "qd0": [0.0, 0.0, 0.0], "h0": [-1.0, -1.0, -1.0]}, index=[3, 102, 999])
    net._waterbody_df = wb.copy()
    net._q0 = q0_ref
    ref_update(net)
    got = nwm_routing.update_waterbody_water_elevation(wb.copy(), q0)
    pd.testing.assert_frame_equal(got, net._waterbody_df, check_dtype=False)


def test_resume_from_lite_restart_files_on_the_oracle(oracle, tmp_path):
    """Checkpoint / resume the way T-Route does it (nwm_routing/__main__.py:269-277 writes, AbstractNetwork.py:591,687
    read): after two of four windows the channel and waterbody states go to lite restart files
    (output.write_lite_restart), a NEW loop starts from what output.read_lite_restart returns, and the four windows
    together equal the uninterrupted run bit for bit."""
    import datetime
    from troute_b200 import nwm_routing, output
    c = A._reference_style_case(n=2500, seed=21, n_lp=6, nsteps=48)
    fn = oracle.compute_network_structured
    full, _, q0_end = _windowed(fn, c, None, 4)

    # the same loop, cut in two halves of two windows with files in between
    ids = c["ids"]
    n_w, qcols = c["nsteps"] // 4, (c["nsteps"] // 4) // c["qts"]
    wb_cols = ["LkArea", "LkMxE", "OrificeA", "OrificeC", "OrificeE", "WeirC", "WeirE", "WeirL", "ifd", "qd0", "h0"]
    wb_static = pd.DataFrame(np.array(c["wbody"], dtype=np.float64), index=c["lake_numbers"], columns=wb_cols)

    def route_window(w, q0_df, wb_df, lo_df):
        part = dict(c)
        part["nsteps"] = n_w
        part["qlat"] = c["qlat"][:, w * qcols:(w + 1) * qcols]
        part["q0"] = q0_df.loc[ids].to_numpy(dtype=np.float32)
        part["wbody"] = wb_df.loc[c["lake_numbers"]].to_numpy(dtype=np.float64)
        return [A._call(fn, part)], None

    q0 = pd.DataFrame(c["q0"], index=ids, columns=["qu0", "qd0", "h0"])
    first, q0_mid, wb_mid, _ = nwm_routing.route_windows(route_window, range(0, 2), q0, wb_static.copy(), pd.DataFrame(), 300.0, n_w)
    t_mid = datetime.datetime(2023, 4, 2) + datetime.timedelta(seconds=300.0 * n_w * 2)
    written = output.write_lite_restart(q0_mid, wb_mid, t_mid, {"lite_restart_output_directory": str(tmp_path)})
    assert len(written) == 2
    del q0_mid, wb_mid

    q0_r, t_r = output.read_lite_restart(written[0])
    wb_states, t_w = output.read_lite_restart(written[1])
    assert t_r == t_w == t_mid
    wb_r = wb_static.copy()                               # static lake parameters from the hydrofabric, states from the file
    wb_r.loc[wb_states.index, ["qd0", "h0"]] = wb_states[["qd0", "h0"]]     # (AbstractNetwork.py:591-606)
    second, q0_fin, _, _ = nwm_routing.route_windows(route_window, range(2, 4), q0_r, wb_r, pd.DataFrame(), 300.0, n_w)

    pieces = [r[0][1][np.argsort(r[0][0])] for r in first + second]
    H.assert_bit_equal(np.concatenate(pieces, axis=1), full, "resumed run vs uninterrupted run")
    H.assert_bit_equal(q0_fin.loc[ids].to_numpy(np.float32), q0_end.loc[ids].to_numpy(np.float32), "final state")
