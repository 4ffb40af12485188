"""The troute_model-shaped window driver (troute_b200.model.DeviceResidentModel, the device-resident form of
src/troute_model.py:138-345 `run` as driven by bmi_troute.update_until): W consecutive `run(values, until)` calls equal ONE
routing call over all steps, bit for bit -- with level pools and with gages whose last observation lies in an earlier
window -- while the only model state that crosses PCIe is the q0 of the first window."""
import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as g
    g.build()
    from troute_b200 import _lib
    assert _lib.lib().trt_device_count() >= 1, "no CUDA device: the routing path has no CPU fallback"
    return _lib.lib()


@pytest.mark.parametrize("short_ts", [False, True])
def test_bmi_windows_equal_one_call_and_move_no_state(eng, short_ts):
    from troute_b200 import synth
    from troute_b200.model import DeviceResidentModel
    from troute_b200.network import RoutingNetwork
    W, Tw = 4, 12                                                    # four hourly coupling windows at dt = 300 s
    T = W * Tw
    case = H.make_case(synth.conus_like(n_total=15000, n_basins=18, seed=43, style="nhd"), nsteps=T, n_lp=10, warm=True)
    n = case["n"]
    ids = (np.arange(n, dtype=np.int64) * 3 + 1000)                  # segment ids are not row numbers
    rng = np.random.default_rng(9)
    G = 30
    grow = np.sort(rng.choice(np.nonzero(case["kind"] == 0)[0], size=G, replace=False)).astype(np.int32)
    usgs = rng.uniform(0.2, 20.0, size=(G, T + 1)).astype(np.float32)
    usgs[rng.random(usgs.shape) < 0.4] = np.nan
    usgs[: G // 2, Tw + 3:] = np.nan                                 # half of the gages fall silent in the second window
    lastobs = rng.uniform(0.2, 20.0, G).astype(np.float32)
    since = -rng.uniform(0.0, 3600.0, G).astype(np.float32)
    base = dict(usgs_positions=grow, lastobs_values_init=lastobs, time_since_lastobs_init=since, da_decay_coefficient=120.0)

    one = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
    one.set_levelpools(case["lp_rows"], case["wbody"])
    one.set_gages(dict(base, usgs_values=usgs, usgs_positions_reach=grow, usgs_positions_gage=np.arange(G, dtype=np.int32),
                       reach_len=np.ones(n, dtype=np.int64), seg_rows=np.arange(n)), T)
    ref, _ = one.route(T, 12, case["qlat"], case["q0"], assume_short_ts=short_ts)
    nudge_ref, _, lv_ref = one.download_gages()
    one.close()

    for full in (True, False):
        m = DeviceResidentModel(ids, case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"], case["q0"],
                                lp_rows=case["lp_rows"], wbody_cols=case["wbody"], gages=base, assume_short_ts=short_ts)
        # the coupler hands over the lateral inflows of MORE ids than the network has, in its own order (troute_model.py:155-166)
        perm = rng.permutation(n)
        extra_ids = np.asarray([7, 11], dtype=np.int64)
        outs, nudges = [], []
        for w in range(W):
            ql = case["qlat"][:, w * Tw // 12:(w + 1) * Tw // 12]
            values = {"land_surface_water_source__volume_flow_rate": np.concatenate([ql[perm], np.ones((2, ql.shape[1]), np.float32)]),
                      "land_surface_water_source__id": np.concatenate([ids[perm], extra_ids]),
                      "usgs_values": usgs[:, w * Tw:(w + 1) * Tw + 1]}
            m.run(values, until=Tw * 300, full_output=full)
            last = ref[:, 3 * ((w + 1) * Tw - 1):3 * (w + 1) * Tw]
            H.assert_bit_equal(values["channel_exit_water_x-section__volume_flow_rate"], last[:, 0], f"window {w} flow")
            H.assert_bit_equal(values["channel_water_flow__speed"], last[:, 1], f"window {w} velocity")
            H.assert_bit_equal(values["channel_water__mean_depth"], last[:, 2], f"window {w} depth")
            H.assert_bit_equal(values["lake_surface__elevation"], last[case["lp_rows"], 2], f"window {w} lake elevation")
            assert np.array_equal(values["q0_index"], ids)
            if full:
                outs.append(values["fvd_results"].reshape(n, 3 * Tw))
                nudges.append(values["nudging"].reshape(G, Tw))
        if full:
            H.assert_bit_equal(np.concatenate(outs, axis=1), ref, "windows through the model vs one call")
            H.assert_bit_equal(np.concatenate(nudges, axis=1), nudge_ref[:, 1:], "nudge series")
            H.assert_bit_equal(values["lastobs_df"].reshape(G, 2)[:, 1], lv_ref, "last observation values")
        # the state crossed PCIe once, on the way in; nothing came back to be sent in again
        assert m.pcie["state_h2d"] == case["q0"].nbytes and m.pcie["state_d2h"] == 0
        assert m.windows == W and m.time == T * 300.0
        if not full:
            assert m.pcie["results_d2h"] == W * n * 12               # the last timestep of every segment, nothing else
        m.close()
