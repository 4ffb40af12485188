"""dataflow_park_kernel (csrc/routing_kernels.cu): parked stragglers and early publication give the same bits.

The kernel decides per wavefront stage, from the stage's width, whether the last unfinished secant solves of a tile are
parked and finished later in batches of 32 (wide stages) or whether every lane publishes as soon as its own solve ends
(narrow stages).  Neither changes a single operation of a solve, so every setting must reproduce the oracle bit for bit
(reference loop: mc_reach.pyx:492-800); the test networks are small, so the thresholds are forced either way."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu

SETTINGS = {
    "round2_kernel": {"park_max": 0, "early_max_tiles": 0},                          # dataflow_kernel, for reference
    "park8_everywhere": {"park_max": 8, "park_min_tiles": 0, "early_max_tiles": 0},
    "park31_everywhere": {"park_max": 31, "park_min_tiles": 0, "early_max_tiles": 0},
    "park1_everywhere": {"park_max": 1, "park_min_tiles": 0, "early_max_tiles": 0},
    "early_everywhere": {"park_max": 0, "early_max_tiles": 1 << 30},
    "park_and_early": {"park_max": 8, "park_min_tiles": 0, "early_max_tiles": 1 << 30},
    "auto_thresholds": {"park_max": 8, "park_min_tiles": -1, "early_max_tiles": -1},
}


def _networks():
    from troute_b200 import synth
    return {
        "tree4095": (synth.binary_tree(4095), 0),
        "hack20k_lp": (synth.hack_tree(20000, seed=5), 25),
        "forest": (synth.conus_like(n_total=30000, n_basins=160, seed=4), 10),
    }


@pytest.mark.parametrize("setting", list(SETTINGS))
@pytest.mark.parametrize("name", ["tree4095", "hack20k_lp", "forest"])
@pytest.mark.parametrize("short_ts", [False, True])
@pytest.mark.parametrize("mode", [2, 4])
def test_park_settings_give_the_same_bits(oracle, name, short_ts, mode, setting):
    down, n_lp = _networks()[name]
    case = H.make_case(down, nsteps=36, n_lp=n_lp, warm=(name != "forest"))
    ref, upref, _ = H.oracle_route(oracle, case, short_ts)
    out, up, stats = H.engine_route(case, short_ts, mode=mode, options=SETTINGS[setting])
    H.assert_bit_equal(out, ref, f"{name} {setting} fvd")
    if n_lp:
        H.assert_bit_equal(up[case["lp_rows"]], upref[case["lp_rows"]], f"{name} {setting} reservoir inflow")
    assert stats["lane_steps"] == case["n"] * case["nsteps"]


@pytest.mark.parametrize("setting", ["park8_everywhere", "park31_everywhere", "park_and_early"])
def test_parking_with_time_chunks_and_small_grids(oracle, setting):
    """time-chunked trt_route (every chunk ends with a final drain of the pools) and a 2-CTA grid (16 warps: every warp
    sees many tiles per stage, the regime parking is meant for)"""
    from troute_b200 import synth
    case = H.make_case(synth.conus_like(n_total=60000, n_basins=300, seed=8), nsteps=48, n_lp=15, warm=True)
    ref, _, _ = H.oracle_route(oracle, case, False)
    for extra in ({"route_chunks": 5}, {"grid_blocks": 2}, {"grid_blocks": 2, "route_chunks": 3}):
        out, _, _ = H.engine_route(case, False, mode=4, options={**SETTINGS[setting], **extra})
        H.assert_bit_equal(out, ref, f"{setting} {extra}")


@pytest.mark.parametrize("setting", ["park8_everywhere", "early_everywhere", "park_and_early"])
@pytest.mark.parametrize("short_ts", [False, True])
def test_parking_with_gages(oracle, setting, short_ts):
    """streamflow nudging (simple_da.pyx:21-89): gage lanes are never parked; flows, nudges, last-observation state"""
    import test_gpu_api as A
    from troute_b200.routing.fast_reach import mc_reach
    c = A._reference_style_case(n=6000, seed=11, n_lp=10, nsteps=48)
    gages = A._gage_inputs(c, n_gages=150, seed=3, obs_steps=30)
    ref = A._call(oracle.compute_network_structured, c, assume_short_ts=short_ts, gages=gages)
    mc_reach.DEFAULT_OPTIONS = {"mode": 2, **SETTINGS[setting]}
    try:
        got = A._call(mc_reach.compute_network_structured, c, assume_short_ts=short_ts, gages=gages)
    finally:
        mc_reach.DEFAULT_OPTIONS = {}
        mc_reach.clear_network_cache()
    H.assert_bit_equal(got[1], ref[1], "flowveldepth with nudging")
    H.assert_bit_equal(got[8], ref[8], "nudge")
    H.assert_bit_equal(got[3][2], ref[3][2], "lastobs_values")


def test_trip_counters_do_not_depend_on_parking(oracle):
    """collect_trips: a parked solve reports the trips of all its pieces"""
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    case = H.make_case(synth.hack_tree(12000, seed=3), nsteps=24, warm=True)
    res = []
    for setting in ("round2_kernel", "park8_everywhere", "early_everywhere"):
        net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
        try:
            net.set_option("mode", 2)
            for k, v in SETTINGS[setting].items():
                net.set_option(k, v)
            net.collect_trips()
            net.route(case["nsteps"], case["qts"], case["qlat"], case["q0"])
            res.append(np.asarray(net.trip_counts()))
        finally:
            net.close()
    assert res[0].sum() > 0
    assert np.array_equal(res[0], res[1]) and np.array_equal(res[0], res[2])
