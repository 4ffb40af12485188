"""Device-resident hand-off between routing windows (trt_continue) and the result checksum (trt_result_hash), through the
C ABI.  W windows started from the state the previous window left ON THE DEVICE == one call over all steps, bit for bit,
with level pools and with gages whose last observation lies in an earlier window -- the property the reference's window loop
has through new_q0 / update_waterbody_water_elevation / new_lastobs (AbstractNetwork.py:177-198, DataAssimilation.py:1506-1551)."""
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


def _gage_case(n, T, seed=5, G=40):
    rng = np.random.default_rng(seed)
    grow = np.sort(rng.choice(n, size=G, replace=False)).astype(np.int32)
    usgs = rng.uniform(0.2, 20.0, size=(G, T + 1)).astype(np.float32)
    usgs[rng.random(usgs.shape) < 0.4] = np.nan
    usgs[:, T // 2:] = np.where(rng.random((G, 1)) < 0.5, np.nan, usgs[:, T // 2:])     # gages that fall silent half way
    lastobs = rng.uniform(0.2, 20.0, G).astype(np.float32)
    since = -rng.uniform(0.0, 3600.0, G).astype(np.float32)
    lastobs[::5] = np.nan; since[::5] = np.nan
    return grow, usgs, lastobs, since


@pytest.mark.parametrize("mode", [1, 2, 3, 4])
@pytest.mark.parametrize("short_ts", [False, True])
def test_windows_on_the_device_equal_one_call(eng, oracle, mode, short_ts):
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    W, Tw = 3, 24
    T = W * Tw
    case = H.make_case(synth.conus_like(n_total=12000, n_basins=15, seed=41, style="nhd"), nsteps=T, n_lp=8, warm=True)
    n = case["n"]
    grow, usgs, lastobs, since = _gage_case(n, T)
    gages = lambda table: dict(usgs_values=table, usgs_positions=grow, usgs_positions_reach=grow, usgs_positions_gage=np.arange(grow.size, dtype=np.int32),
                               lastobs_values_init=lastobs, time_since_lastobs_init=since, da_decay_coefficient=120.0,
                               reach_len=np.ones(n, dtype=np.int64), seg_rows=np.arange(n))

    def make():
        net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
        net.set_levelpools(case["lp_rows"], case["wbody"])
        net.set_option("mode", mode); net.set_option("deep_lanes", 1500)
        return net

    one = make()
    one.set_gages(gages(usgs), T)
    ref, upref = one.route(T, 12, case["qlat"], case["q0"], assume_short_ts=short_ts, want_upstream=True)
    nudge_ref, lt_ref, lv_ref = one.download_gages()
    whole_hash = one.result_hash()
    one.close()
    assert whole_hash == H.result_hash(ref)                       # the device checksum is the numpy one

    net = make()
    net.set_gages(gages(usgs[:, :Tw + 1]), Tw)
    outs, ups, nudges = [], [], []
    for w in range(W):
        ql = case["qlat"][:, w * Tw // 12:(w + 1) * Tw // 12]
        if w == 0:
            net.upload(Tw, 12, ql, case["q0"])
        else:
            net.continue_window(Tw, 12, ql, usgs_values=usgs[:, w * Tw:(w + 1) * Tw + 1])
        net.run(short_ts)
        o, u = net.download(want_upstream=True)
        outs.append(o); ups.append(u)
        nudges.append(net.download_gages())
    net.close()
    H.assert_bit_equal(np.concatenate(outs, axis=1), ref, f"mode {mode}: {W} device-resident windows vs one call")
    H.assert_bit_equal(np.concatenate(ups, axis=1)[case["lp_rows"]], upref[case["lp_rows"]], "reservoir inflow")
    got_nudge = np.concatenate([nudges[0][0]] + [x[0][:, 1:] for x in nudges[1:]], axis=1)
    H.assert_bit_equal(got_nudge, nudge_ref, "nudge series")
    # last-observation state after the last window, re-based to the start of the one call
    H.assert_bit_equal(nudges[-1][2], lv_ref, "last observation values")
    same = np.isnan(lt_ref) | np.isclose(nudges[-1][1] + np.float32((W - 1) * Tw * 300.0), lt_ref, rtol=0, atol=1e-2)
    assert same.all()


def test_result_hash_adds_up_over_row_subsets(eng):
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    case = H.make_case(synth.binary_tree(4095), nsteps=24, warm=True)
    net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
    out, _ = net.route(24, 12, case["qlat"], case["q0"])
    ids = (np.arange(case["n"], dtype=np.int64) * 7919 + 13)
    total = net.result_hash(ids=ids)
    assert total == H.result_hash(out, ids)
    a = np.arange(0, case["n"], 2, dtype=np.int64); b = np.arange(1, case["n"], 2, dtype=np.int64)
    part = (net.result_hash(rows=a, ids=ids[a]) + net.result_hash(rows=b, ids=ids[b])) % (1 << 64)
    net.close()
    assert part == total
    out2 = out.copy(); out2[17, 5] = np.nextafter(out2[17, 5], np.float32(np.inf))
    assert H.result_hash(out2, ids) != total                      # one ulp in one value changes it
