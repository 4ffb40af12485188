"""GPU parity: the CUDA path through the C ABI vs the CPU oracle on identical inputs.

Bar (BASELINE.json north_star): streamflow within 1e-5 relative of the reference arithmetic.  What is asserted
here is stronger: BIT equality of q, v, d with the oracle's deterministic-powf build (the numerics contract of
include/trt_detmath.h), and <= 1e-5 relative agreement with the oracle's platform-libm build on all but the
secant-termination flips a 1-ulp powf difference causes (measured and bounded below)."""
import json
import os

import numpy as np
import pytest

import helpers as H

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REL_TOL = 1e-5   # north_star tolerance on streamflow


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as g
    g.build()
    from troute_b200 import network
    from troute_b200 import _lib
    assert _lib.lib().trt_device_count() >= 1, "no CUDA device: the routing path has no CPU fallback"
    return network


def test_fastpath_division_is_the_ieee_quotient_on_the_device(eng):
    """McDivFast (csrc/mc_device.cuh), what the marching lanes divide with: seeded by the device's MUFU.RCP, the inline fast
    path returns the IEEE quotient for every pair inside its window of exponents (numpy float32 division is IEEE)."""
    rng = np.random.default_rng(7)
    n = 1 << 24
    for kind in range(3):
        if kind == 0:
            a = (rng.standard_normal(n) * np.exp2(rng.uniform(-64, 64, n))).astype(np.float32)
            d = (rng.standard_normal(n) * np.exp2(rng.uniform(-64, 64, n))).astype(np.float32)
        elif kind == 1:      # quotients next to rounding boundaries
            a = rng.integers(1, 1 << 24, n).astype(np.float32) * np.exp2(rng.integers(-30, 30, n)).astype(np.float32)
            d = rng.integers(1, 1 << 12, n).astype(np.float32) * np.exp2(rng.integers(-30, 30, n)).astype(np.float32)
        else:                # the magnitudes of the solve, zero dividends of either sign, operands outside the window
            a = np.exp(rng.uniform(np.log(1e-8), np.log(1e8), n)).astype(np.float32) * rng.choice([-1.0, 1.0], n).astype(np.float32)
            d = np.exp(rng.uniform(np.log(1e-8), np.log(1e8), n)).astype(np.float32)
            a[:1000] = 0.0
            a[1000:2000] = -0.0
            d[2000:2010] = [0.0, -0.0, np.inf, np.nan, 1e-40, 3e38, 1e-30, 1e30, -np.inf, 1e-45]
            a[2010:2016] = [np.inf, np.nan, 1e-40, 3e38, -np.inf, 1e-45]
        q, inside = eng.fdiv_batch(a, d)
        with np.errstate(all="ignore"):
            want = a / d
        assert inside.mean() > 0.8
        assert np.array_equal(q[inside].view(np.int32), want[inside].view(np.int32)), kind
        if kind == 2:
            assert not inside[2000:2016].any()             # zero / non-finite / denormal / huge operands leave the window


def test_powf_det_bits(eng, oracle):
    rng = np.random.default_rng(0)
    n = 1 << 20
    x = np.exp(rng.uniform(np.log(1e-30), np.log(1e30), n)).astype(np.float32)
    x[:8] = [0.0, -0.0, np.inf, -1.0, np.nan, 1.0, 1e-45, 3.4e38]
    y = np.array([2 / 3, 5 / 3, 0.5, 1.5], dtype=np.float32)[rng.integers(0, 4, n)]
    H.assert_bit_equal(eng.powf_batch(x, y), oracle.powf(x, y, oracle.POW_DET), "powf_det")


def test_mc_demo_kat_on_gpu(eng):
    k = json.load(open(os.path.join(GOLD, "mc_demo_kat.json")))
    c, s = k["channel"], k["single"]
    row = np.array([[c["dt"], s["qup"], s["quc"], s["qdp"], c["ql"], c["dx"], c["bw"], c["tw"], c["twcc"], c["n"],
                     c["ncc"], c["cs"], c["s0"], s["velp"], s["depthp"]]], dtype=np.float32)
    out = eng.mc_segment_batch(row)[0]
    e = s["expected"]
    assert out[2] == np.float32(e["depthc"])
    assert abs(float(out[0]) - e["qdc"]) / e["qdc"] < 1e-6      # 1 ulp, see tests/test_oracle_kat.py
    assert abs(float(out[1]) - e["velc"]) / e["velc"] < 1e-6


def test_mc_suite_bits(eng, oracle):
    """The reference's 5000 randomized kernel inputs: all six outputs and the iteration count, bit for bit."""
    in15 = np.load(os.path.join(GOLD, "mc_suite_seed16.npy"))
    got, it_g = eng.mc_segment_batch(in15, want_iters=True)
    ref, it_r = oracle.mc_segment_batch(in15, pow_mode=oracle.POW_DET)
    H.assert_bit_equal(got, ref, "mc suite")
    assert np.array_equal(it_g, it_r)
    # vs the platform-libm build: within 1e-5 relative except on termination flips
    lib, _ = oracle.mc_segment_batch(in15, pow_mode=oracle.POW_LIBM)
    rel = np.abs(got[:, 0] - lib[:, 0]) / np.maximum(np.abs(lib[:, 0]), 1e-6)
    assert (rel <= REL_TOL).mean() >= 0.999


def test_mc_edge_inputs_bits(eng, oracle):
    """Edge rows: zero flows, cs == 0, bw > tw, bw == tw, twcc == 0 (NWM 3.0 exception), huge flows (retry ladder),
    tiny depth, negative lateral inflow."""
    base = np.array([300, 1, 1, 1, 0.1, 1000, 5, 8, 24, 0.06, 0.12, 0.6, 0.01, 0, 0.5], dtype=np.float32)
    rows = []
    def add(**kw):
        r = base.copy()
        names = ["dt", "qup", "quc", "qdp", "ql", "dx", "bw", "tw", "twcc", "n", "ncc", "cs", "s0", "velp", "depthp"]
        for k2, v in kw.items():
            r[names.index(k2)] = v
        rows.append(r)
    add(qup=0, quc=0, qdp=0, ql=0)
    add(cs=0)
    add(bw=10, tw=8)
    add(bw=8, tw=8)
    add(twcc=0, depthp=30, qup=5000, quc=5000, qdp=5000)
    add(ncc=0, depthp=30, qup=5000, quc=5000, qdp=5000)
    add(qup=70000, quc=70000, qdp=70000, ql=70000, depthp=100)
    add(depthp=1e-6, qup=1e-6, quc=1e-6, qdp=1e-6, ql=1e-7)
    add(ql=-5.0)
    add(ql=-0.5, qup=0.1, quc=0.1, qdp=0.1)
    add(dx=1, s0=4.6)
    add(dx=95714, s0=1e-5, depthp=0)
    add(depthp=-1.0)
    rng = np.random.default_rng(3)
    for _ in range(3000):
        r = base.copy()
        r[1:5] = np.exp(rng.uniform(np.log(1e-5), np.log(7e4), 4))
        r[4] *= rng.choice([1, 1, 1, -1e-3, 0])
        r[5] = np.exp(rng.uniform(0, np.log(95714)))
        r[6] = np.exp(rng.uniform(np.log(0.135), np.log(230)))
        r[7] = r[6] * rng.choice([1 / 0.6, 1.0, 0.9, 3.0])
        r[8] = r[7] * rng.choice([3.0, 0.0, 1.0])
        r[9] = rng.uniform(0.02, 0.2); r[10] = r[9] * rng.choice([2.0, 1.0, 0.0])
        r[11] = rng.choice([0.0, 0.0846, 0.5857, 2.254]); r[12] = np.exp(rng.uniform(np.log(1e-5), np.log(4.6)))
        r[14] = rng.choice([0.0, 1e-3, 0.3, 3.0, 40.0])
        rows.append(r)
    in15 = np.asarray(rows, dtype=np.float32)
    got, it_g = eng.mc_segment_batch(in15, want_iters=True)
    ref, it_r = oracle.mc_segment_batch(in15, pow_mode=oracle.POW_DET)
    H.assert_bit_equal(got, ref, "mc edge rows")
    assert np.array_equal(it_g, it_r)


def test_levelpool_kats_on_gpu(eng, oracle):
    k = json.load(open(os.path.join(GOLD, "levelpool_kats.json")))
    for c in k["cases"]:
        q, h = eng.levelpool_series(c["wbody_row"], c["inflow"], 0.0, c["routing_period"])
        assert q[-1] == np.float32(c["expected_final_outflow"]), c["fixture"]
        assert h[-1] == np.float32(c["expected_final_water_elevation"]), c["fixture"]
        qo, ho = oracle.levelpool_series(c["wbody_row"], c["inflow"], 0.0, c["routing_period"], pow_mode=oracle.POW_DET)
        H.assert_bit_equal(q, qo, "lp outflow series")
        H.assert_bit_equal(h, ho, "lp elevation series")


def _networks():
    from troute_b200 import synth
    return {
        "tree4095": (synth.binary_tree(4095), 0),
        "chain300": (synth.chain(300), 0),
        "single": (synth.chain(1), 0),
        "hack20k_lp": (synth.hack_tree(20000, seed=5), 25),
        "forest": (synth.conus_like(n_total=30000, n_basins=160, seed=4), 10),
    }


@pytest.mark.parametrize("name", ["tree4095", "chain300", "single", "hack20k_lp", "forest"])
@pytest.mark.parametrize("short_ts", [False, True])
@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4])
def test_network_bits(eng, oracle, name, short_ts, mode):
    down, n_lp = _networks()[name]
    case = H.make_case(down, nsteps=36, n_lp=n_lp, warm=(name != "forest"))
    ref, upref, extras = H.oracle_route(oracle, case, short_ts)
    out, up, stats = H.engine_route(case, short_ts, mode=mode)
    assert np.isfinite(out).all()
    H.assert_bit_equal(out, ref, f"{name} fvd")
    if n_lp:
        H.assert_bit_equal(up[case["lp_rows"]], upref[case["lp_rows"]], f"{name} reservoir inflow")
        assert (np.delete(up, case["lp_rows"], axis=0) == 0).all()
    assert stats["lane_steps"] == case["n"] * case["nsteps"]


def test_network_vs_libm_oracle_tolerance(eng, oracle):
    """Against the platform-libm arithmetic (what a gfortran build of the reference computes on THIS machine).

    The secant solve stops at a 1 % tolerance far from its fixed point, so the reference amplifies a 1-ulp difference
    between two powf implementations: its own two arithmetic builds (platform powf vs the bit-specified powf) agree
    bit for bit on ~97.5 % of the segment-timesteps of this case and within 1e-5 relative on 99.6 %; the remainder
    are lanes whose iteration count flipped and everything downstream of them (measured in
    tests/test_oracle_network.py, recorded in DESIGN.md "Numerics contract").  The GPU reproduces the bit-specified
    build exactly, so its distance to the libm build must be that same set -- not one lane more."""
    from troute_b200 import synth
    case = H.make_case(synth.hack_tree(20000, seed=9), nsteps=48, warm=False)
    ref, _, _ = H.oracle_route(oracle, case, False, pow_mode=oracle.POW_LIBM)
    det, _, _ = H.oracle_route(oracle, case, False, pow_mode=oracle.POW_DET)
    out, _, _ = H.engine_route(case, False)
    H.assert_bit_equal(out, det, "GPU vs bit-specified oracle")
    q_ref, q_out = ref[:, 0::3], out[:, 0::3]
    rel = np.abs(q_out - q_ref) / np.maximum(np.abs(q_ref), 1e-3)
    assert (rel <= REL_TOL).mean() >= 0.99, float((rel <= REL_TOL).mean())
    assert (q_ref == q_out).mean() >= 0.95
    assert np.median(rel) == 0.0


def test_restart_chunks_equal_one_run(eng, oracle):
    """Checkpoint/resume contract (AbstractNetwork.new_q0, AbstractNetwork.py:177-191): two 24-step calls with the
    state hand-off q0 = fvd[:, [-3, -3, -1]] reproduce one 48-step call bit for bit."""
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    case = H.make_case(synth.hack_tree(8000, seed=2), nsteps=48, warm=True)
    full, _, _ = H.engine_route(case, False)
    net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
    a, _ = net.route(24, 12, case["qlat"][:, :2], case["q0"])
    q0b = a[:, [-3, -3, -1]].copy()
    b, _ = net.route(24, 12, case["qlat"][:, 2:4], q0b)
    net.close()
    H.assert_bit_equal(np.concatenate([a, b], axis=1), full, "chunked run")


def test_boundary_rows_prescribed(eng, oracle):
    """upstream_results injection (mc_reach.pyx:458-469): cutting a network at a segment and prescribing that
    segment's flow series reproduces the uncut downstream results."""
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork, TRT_KIND_BOUNDARY
    down = synth.hack_tree(5000, seed=11)
    case = H.make_case(down, nsteps=30, warm=True)
    full, _, _ = H.engine_route(case, False)
    sizes = synth.subtree_sizes(down)
    cut = int(np.argmin(np.abs(sizes - 1500)))           # a segment with ~1500 segments upstream
    # rows upstream of `cut` (excluded from the cut network)
    up_ptr, up_rows = case["up_ptr"], case["up_rows"]
    mask = np.zeros(case["n"], dtype=bool)
    stack = list(up_rows[up_ptr[cut]:up_ptr[cut + 1]])
    while stack:
        r = int(stack.pop())
        mask[r] = True
        stack.extend(up_rows[up_ptr[r]:up_ptr[r + 1]].tolist())
    keep = np.nonzero(~mask)[0]
    remap = -np.ones(case["n"], dtype=np.int64); remap[keep] = np.arange(keep.size)
    down2 = np.where(down[keep] >= 0, remap[np.maximum(down[keep], 0)], -1)
    p2, r2 = synth.upstream_csr(down2)
    kind = np.zeros(keep.size, dtype=np.uint8); kind[remap[cut]] = TRT_KIND_BOUNDARY
    net = RoutingNetwork(p2, r2, kind, case["params"][keep], case["cols"])
    out, _ = net.route(30, 12, case["qlat"][keep], case["q0"][keep], bnd_rows=[remap[cut]], bnd_fvd=full[cut][None, :])
    net.close()
    H.assert_bit_equal(out, full[keep], "cut network")


def test_argument_errors(eng):
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    case = H.make_case(synth.binary_tree(63), nsteps=24)
    net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
    with pytest.raises(ValueError):      # mc_reach.pyx:246-247
        net.route(24, 12, case["qlat"][:, :1], case["q0"])
    with pytest.raises(ValueError):      # mc_reach.pyx:243-244
        net.route(24, 12, case["qlat"][:-1], case["q0"])
    net.close()
    # a cycle is not a river network
    with pytest.raises(ValueError):
        RoutingNetwork(np.array([0, 1, 2]), np.array([1, 0]), np.zeros(2, np.uint8), case["params"][:2], case["cols"])
    # empty network / zero steps
    empty = RoutingNetwork(np.zeros(1, np.int64), np.zeros(0, np.int64), np.zeros(0, np.uint8),
                           np.zeros((0, 10), np.float32), case["cols"])
    fvd, _ = empty.route(12, 12, np.zeros((0, 1), np.float32), np.zeros((0, 3), np.float32))
    assert fvd.shape == (0, 36)
    empty.close()


def test_full_size_config2_properties(eng, oracle):
    """BASELINE config 2 (binary tree, 1,048,576 segments) at full width, 48 steps: (i) the dataflow kernel, the
    grid-barrier persistent kernel and the launch-per-stage schedule agree bit for bit; (ii) a complete sub-tree (4,095 segments upstream of one node)
    is independent of the rest of the network, so the oracle run on it alone must reproduce those rows exactly."""
    from troute_b200 import synth
    N, T = 1_048_576, 48
    down = synth.binary_tree(N)
    case = H.make_case(down, nsteps=T)
    a, _, st = H.engine_route(case, False, mode=2, want_upstream=False)
    b, _, _ = H.engine_route(case, False, mode=0, want_upstream=False)
    assert np.array_equal(a.view(np.int32), b.view(np.int32))
    b, _, _ = H.engine_route(case, False, mode=1, want_upstream=False)
    assert np.array_equal(a.view(np.int32), b.view(np.int32))
    del b
    assert np.isfinite(a).all() and st["stages"] == 20 + T
    for m in (3, 4):
        b, _, _ = H.engine_route(case, False, mode=m, want_upstream=False)
        assert np.array_equal(a.view(np.int32), b.view(np.int32)), f"mode {m}"
    root = 300                                            # heap index; its subtree has 2^12 - 1 nodes at N = 2^20
    ids = [root]
    frontier = [root]
    while frontier:
        nxt = [c for f in frontier for c in (2 * f + 1, 2 * f + 2) if c < N]
        ids.extend(nxt); frontier = nxt
    ids = np.asarray(sorted(ids), dtype=np.int64)
    remap = -np.ones(N, dtype=np.int64); remap[ids] = np.arange(ids.size)
    d2 = np.where(ids == root, -1, remap[np.maximum(down[ids], 0)])
    sub = dict(case)
    p2, r2 = synth.upstream_csr(d2)
    sub.update(n=ids.size, down=d2, params=case["params"][ids], qlat=case["qlat"][ids], q0=case["q0"][ids],
               up_ptr=p2, up_rows=r2, kind=case["kind"][ids])
    ref, _, _ = H.oracle_route(oracle, sub, False)
    H.assert_bit_equal(a[ids], ref, "sub-tree of the full-size run")


@pytest.mark.parametrize("short_ts", [False, True])
def test_marching_schedule_options_do_not_change_results(eng, oracle, short_ts):
    """Marching lanes (mode 3) and the dataflow + marching hybrid (mode 4): any lanes-per-warp packing, any split level,
    a 2-CTA grid (units far outnumber the resident warps, so later units start long after their upstream units
    finished) -- same bits as the oracle."""
    from troute_b200 import synth
    case = H.make_case(synth.conus_like(n_total=25000, n_basins=40, seed=8), nsteps=30, n_lp=12, warm=True)
    ref, upref, _ = H.oracle_route(oracle, case, short_ts)
    trials = [dict(mode=3, march_group=1), dict(mode=3, march_group=5), dict(mode=3, march_group=32),
              dict(mode=3, march_group=32, grid_blocks=2), dict(mode=4, deep_level=0), dict(mode=4, deep_level=3),
              dict(mode=4, deep_level=40, march_group=4), dict(mode=4, deep_level=100000),
              dict(mode=4, deep_lanes=500, march_group=2), dict(mode=4, deep_lanes=0),
              dict(mode=4, deep_lanes=20000), dict(mode=4, deep_lanes=3000, grid_blocks=3), dict(mode=2, grid_blocks=1),
              # the marching kernel BESIDE the dataflow kernel (one un-chunked call: both phases in one launch sequence)
              dict(mode=4, deep_lanes=3000, route_chunks=1, overlap_march=1),
              dict(mode=4, deep_level=3, route_chunks=1, overlap_march=1, march_group=2),
              dict(mode=4, deep_lanes=20000, route_chunks=1, overlap_march=1)]
    for opts in trials:
        out, up, _ = H.engine_route(case, short_ts, options=opts)
        H.assert_bit_equal(out, ref, f"{opts}")
        H.assert_bit_equal(up[case["lp_rows"]], upref[case["lp_rows"]], f"{opts} reservoir inflow")


@pytest.mark.parametrize("short_ts", [False, True])
def test_time_chunked_route_call(eng, oracle, short_ts):
    """trt_route cuts the call into time chunks and copies the finished columns home while the next chunk runs: any
    number of chunks (also one that does not divide the step count) gives the bits of the unchunked run."""
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    case = H.make_case(synth.conus_like(n_total=20000, n_basins=30, seed=3, style="nhd"), nsteps=50, n_lp=10, warm=True)
    ref, upref, _ = H.oracle_route(oracle, case, short_ts)
    for mode, chunks in ((4, 1), (4, 2), (4, 3), (4, 7), (2, 4), (3, 2), (4, 50), (1, 2), (0, 3)):
        net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
        net.set_levelpools(case["lp_rows"], case["wbody"])
        net.set_option("mode", mode); net.set_option("route_chunks", chunks); net.set_option("deep_lanes", 3000)
        out, up = net.route_call(50, 12, case["qlat"], case["q0"], assume_short_ts=short_ts, want_upstream=True)
        net.close()
        H.assert_bit_equal(out, ref, f"mode {mode}, {chunks} chunks")
        H.assert_bit_equal(up[case["lp_rows"]], upref[case["lp_rows"]], f"mode {mode}, {chunks} chunks: reservoir inflow")


def test_within_level_order_and_trip_counts(eng, oracle):
    """Any order of the segments inside a wavefront level gives the same bits; the trip counts the engine collects add
    up to the oracle's iteration histogram; ordering by them (what bench.py does after a calibration call) is one such
    order."""
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    case = H.make_case(synth.conus_like(n_total=20000, n_basins=30, seed=9, style="nhd"), nsteps=30, n_lp=10, warm=True)
    ref, upref, extras = H.oracle_route(oracle, case, False)
    hist = extras["iter_hist"]
    net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
    net.set_levelpools(case["lp_rows"], case["wbody"])
    net.set_option("mode", 2); net.set_option("collect_trips", 1)
    out, _ = net.route(30, 12, case["qlat"], case["q0"])
    trips = net.trip_counts()
    net.close()
    H.assert_bit_equal(out, ref, "collecting run")
    # oracle histogram buckets: 0..7 exact, then 8-15, 16-31, 32-63, 64-127, 128-255, 256-511, 512+ (troute_oracle.c)
    lo = np.array([0, 1, 2, 3, 4, 5, 6, 7, 8, 16, 32, 64, 128, 256, 512, 0])
    hi = np.array([0, 1, 2, 3, 4, 5, 6, 7, 15, 31, 63, 127, 255, 511, 760, 0])
    assert int((hist * lo).sum()) <= int(trips.sum()) <= int((hist * hi).sum())
    assert int((trips > 0).sum()) == case["n"] - case["lp_rows"].size
    assert (trips[case["lp_rows"]] == 0).all()
    rng = np.random.default_rng(0)
    for key in (trips, rng.integers(0, 5, case["n"]).astype(np.int32)):
        net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"], order_key=key)
        net.set_levelpools(case["lp_rows"], case["wbody"])
        out, up = net.route(30, 12, case["qlat"], case["q0"], want_upstream=True)
        net.close()
        H.assert_bit_equal(out, ref, "ordered network")
        H.assert_bit_equal(up[case["lp_rows"]], upref[case["lp_rows"]], "ordered network: reservoir inflow")


@pytest.mark.parametrize("nsteps,qts", [(1, 1), (7, 3), (13, 12), (25, 5), (2, 12)])
def test_ragged_step_counts(eng, oracle, nsteps, qts):
    """Step counts that are not multiples of qts_subdivisions, a single step, fewer steps than time chunks or marching
    lanes: qlat column (t - 1) // qts (mc_reach.pyx:723), every schedule, direct and time-chunked."""
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    case = H.make_case(synth.conus_like(n_total=6000, n_basins=10, seed=31, style="nhd"), nsteps=nsteps, qts=qts, n_lp=6,
                       warm=True)
    for short_ts in (False, True):
        ref, upref, _ = H.oracle_route(oracle, case, short_ts)
        for mode in (0, 1, 2, 3, 4):
            net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
            net.set_levelpools(case["lp_rows"], case["wbody"])
            net.set_option("mode", mode); net.set_option("deep_lanes", 1500)
            out, up = net.route(nsteps, qts, case["qlat"], case["q0"], assume_short_ts=short_ts, want_upstream=True)
            H.assert_bit_equal(out, ref, f"mode {mode} direct")
            out2, up2 = net.route_call(nsteps, qts, case["qlat"], case["q0"], assume_short_ts=short_ts, want_upstream=True)
            net.close()
            H.assert_bit_equal(out2, ref, f"mode {mode} chunked")
            H.assert_bit_equal(up2[case["lp_rows"]], upref[case["lp_rows"]], f"mode {mode} reservoir inflow")


def test_gate_and_grid_options_do_not_change_results(eng, oracle):
    """Dataflow schedule knobs: run-ahead gate 1 / 50, tiny grid (2 CTAs) -- same bits."""
    from troute_b200 import synth
    from troute_b200.network import RoutingNetwork
    case = H.make_case(synth.hack_tree(12000, seed=21), nsteps=30, n_lp=8, warm=True)
    ref, _, _ = H.oracle_route(oracle, case, False)
    for gate, grid in ((1, 0), (50, 0), (3, 2)):
        net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"])
        net.set_levelpools(case["lp_rows"], case["wbody"])
        net.set_option("mode", 2); net.set_option("gate", gate); net.set_option("grid_blocks", grid)
        out, _ = net.route(30, 12, case["qlat"], case["q0"])
        net.close()
        H.assert_bit_equal(out, ref, f"gate={gate} grid={grid}")


@pytest.mark.parametrize("P", [2, 3])
@pytest.mark.parametrize("short_ts", [False, True])
@pytest.mark.parametrize("split", ["all-march", "dataflow+march", "dataflow"])
def test_sharded_on_one_gpu_concurrent_kernels(eng, oracle, P, short_ts, split):
    """The sharding path without a second GPU: P shard handles on the SAME device, wired through each other's flow
    arrays exactly as peers are wired through CUDA IPC, their kernels running concurrently on P streams with a
    quarter-size grid each.  Results of all shards together == the unsharded oracle, bit for bit."""
    from troute_b200 import synth, partition, hostgraph
    from troute_b200.network import RoutingNetwork
    down = synth.conus_like(n_total=40000, n_basins=30, seed=12)
    case = H.make_case(down, nsteps=30, n_lp=20, warm=True)
    ref, upref, _ = H.oracle_route(oracle, case, short_ts)
    level = hostgraph.levels(down, case["up_ptr"])
    shard, plans, stats = partition.plan_shards(down, case["up_ptr"], case["up_rows"], case["kind"], P,
                                                pieces_per_shard=6, level=level)
    assert stats["n_cut_edges"] > 0
    nets = []
    lp_index = {int(r): i for i, r in enumerate(case["lp_rows"])}
    for p in plans:
        net = RoutingNetwork(p.up_ptr, p.up_rows, p.kind, case["params"][p.rows], case["cols"], levels=p.levels)
        loc_lp = np.nonzero(p.kind == 1)[0]
        net.set_levelpools(loc_lp, case["wbody"][[lp_index[int(g)] for g in p.rows[loc_lp]]] if loc_lp.size else np.zeros((0, 11)))
        net.set_option("grid_blocks", 74)
        if split == "dataflow+march":
            # every shard splits at the same level (multigpu.global_deep_level): the deepest <= 1500 segments of any shard march
            from troute_b200 import multigpu
            net.set_option("deep_level", multigpu.global_deep_level(level, shard, P, 1500))
        elif split == "dataflow":
            net.set_option("mode", 2)
        net.set_imports(p.imports)
        net.upload(30, 12, case["qlat"][p.rows], case["q0"][p.rows])
        nets.append(net)
    pos = [net.positions() for net in nets]
    for p, net in zip(plans, nets):
        rows, dst, glob = p.exports
        for d in sorted(set(dst.tolist())):
            net.set_peer_ptr(d, nets[d].state_ptr(), plans[d].rows.size)
        loc = [dict(zip(q.rows.tolist(), range(q.rows.size))) for q in plans]
        peer_pos = [pos[int(d)][loc[int(d)][int(g)]] for d, g in zip(dst, glob)]
        net.set_exports(rows, dst.astype(np.int32), np.asarray(peer_pos, dtype=np.int64))
    for net in nets:
        net.prepare()
    for net in nets:
        net.run_async(short_ts)
    for net in nets:
        net.sync()
    for p, net in zip(plans, nets):
        out, up = net.download(want_upstream=True)
        H.assert_bit_equal(out[p.own], ref[p.rows[p.own]], f"shard {p.rank}")
        loc_lp = np.nonzero(p.kind == 1)[0]
        H.assert_bit_equal(up[loc_lp], upref[p.rows[loc_lp]], f"shard {p.rank} reservoir inflow")
        net.close()


@pytest.mark.parametrize("P", [2, 3])
def test_sharded_nudging_on_one_gpu(eng, oracle, P):
    """Streamflow nudging in a sharded run: every shard assimilates the gages on its own segments, and a nudged flow
    that crosses a cut edge reaches the downstream shard as the nudged value (it is exported after the replacement).
    P shard handles on one device with concurrent kernels, as in test_sharded_on_one_gpu_concurrent_kernels.
    (Round 1 listed this test as an open issue: it timed out whenever it was the FIRST sharded run of a process.  The cause
    was CUDA's lazy module loading, not the nudging -- the second handle's first launch had to load a kernel while the
    first handle's kernel was already spinning on its output; trt_network_create now loads every kernel up front,
    preload_routing_kernels in csrc/routing_kernels.cu, and tools/dbg_sharded_nudging.py shows every variant green.)"""
    from troute_b200 import synth, partition, hostgraph, multigpu
    from troute_b200.network import RoutingNetwork
    T = 30
    down = synth.conus_like(n_total=30000, n_basins=20, seed=14, style="nhd")
    case = H.make_case(down, nsteps=T, warm=True)
    n = case["n"]
    rng = np.random.default_rng(4)
    G = 150
    grow = np.sort(rng.choice(n, size=G, replace=False)).astype(np.int32)
    usgs = rng.uniform(0.2, 20.0, size=(G, 18)).astype(np.float32)
    usgs[rng.random(usgs.shape) < 0.3] = np.nan
    lastobs = rng.uniform(0.2, 20.0, G).astype(np.float32)
    since = -rng.uniform(0.0, 3600.0, G).astype(np.float32)
    lastobs[::7] = np.nan; since[::7] = np.nan
    level = hostgraph.levels(down, case["up_ptr"])
    inv_order = np.empty(n, dtype=np.int64)
    inv_order[np.argsort(level, kind="stable")] = np.arange(n)          # oracle_route lists one-segment reaches in this order
    g_oracle = dict(usgs_values=usgs, usgs_positions=grow, usgs_positions_reach=inv_order[grow].astype(np.int32),
                    usgs_positions_gage=np.arange(G, dtype=np.int32), lastobs_values_init=lastobs,
                    time_since_lastobs_init=since, da_decay_coefficient=120.0)
    ref, _, extras = H.oracle_route(oracle, case, False, gages=g_oracle)
    plain, _, _ = H.oracle_route(oracle, case, False)
    assert not np.array_equal(ref, plain)
    shard, plans, stats = partition.plan_shards(down, case["up_ptr"], case["up_rows"], case["kind"], P,
                                                pieces_per_shard=6, level=level)
    assert stats["n_cut_edges"] > 0
    deep = multigpu.global_deep_level(level, shard, P, 2000)
    # A gage with an observation at step 0 replaces the INITIAL flow of its segment (mc_reach.pyx:403-411).  The owning
    # shard does that on the device (reset_gages_kernel); a shard that imports the segment reads q[u, 0] from its own
    # import row, which is initialised from the q0 it was given -- so the planner hands every shard the replaced value.
    q0_eff = case["q0"].copy()
    obs0 = ~np.isnan(usgs[:, 0])
    q0_eff[grow[obs0], 0] = usgs[obs0, 0]
    nets, gsel = [], []
    for p in plans:
        net = RoutingNetwork(p.up_ptr, p.up_rows, p.kind, case["params"][p.rows], case["cols"], levels=p.levels)
        net.set_option("grid_blocks", 74); net.set_option("deep_level", deep)
        net.set_imports(p.imports)
        own_rows = p.rows[p.own]
        sel = np.nonzero(np.isin(grow, own_rows))[0]                      # gages of this shard
        loc = np.searchsorted(p.rows, grow[sel]).astype(np.int32)
        nloc = p.rows.size
        net.set_gages(dict(usgs_values=usgs[sel], usgs_positions=loc, usgs_positions_reach=loc,
                           usgs_positions_gage=np.arange(sel.size, dtype=np.int32), lastobs_values_init=lastobs[sel],
                           time_since_lastobs_init=since[sel], da_decay_coefficient=120.0,
                           reach_len=np.ones(nloc, dtype=np.int64), seg_rows=np.arange(nloc)), T, routing_period=300.0)
        net.upload(T, 12, case["qlat"][p.rows], q0_eff[p.rows])
        nets.append(net); gsel.append(sel)
    assert sum(s.size for s in gsel) == G
    pos = [net.positions() for net in nets]
    loc_of = [dict(zip(q.rows.tolist(), range(q.rows.size))) for q in plans]
    for p, net in zip(plans, nets):
        rows, dst, glob = p.exports
        for d in sorted(set(dst.tolist())):
            net.set_peer_ptr(d, nets[d].state_ptr(), plans[d].rows.size)
        peer_pos = [pos[int(d)][loc_of[int(d)][int(g)]] for d, g in zip(dst, glob)]
        net.set_exports(rows, dst.astype(np.int32), np.asarray(peer_pos, dtype=np.int64))
    for net in nets:
        net.prepare()
    for net in nets:
        net.run_async(False)
    for net in nets:
        net.sync()
    for p, net, sel in zip(plans, nets, gsel):
        out, _ = net.download()
        H.assert_bit_equal(out[p.own], ref[p.rows[p.own]], f"shard {p.rank} flows")
        nudge, lt, lv = net.download_gages()
        H.assert_bit_equal(nudge, extras["nudge"][sel], f"shard {p.rank} nudge")
        H.assert_bit_equal(lt, extras["lastobs_times"][sel], f"shard {p.rank} lastobs_times")
        H.assert_bit_equal(lv, extras["lastobs_values"][sel], f"shard {p.rank} lastobs_values")
        net.close()
