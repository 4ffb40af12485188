"""GPU parity of the diffusive-wave mainstem solver: trt_c_diffnw / trt_diffnw_batch (CUDA, through the C ABI and the
Python mirror of fast_reach/diffusive.pyx) against oracle/diffusive_oracle.c in its bit-specified-pow build.

Bar: BIT equality of q_ev_g, elv_ev_g, depth_ev_g (binary64).  The device code keeps the operand order of every Fortran
expression, is compiled with -fmad=false and evaluates x**y with trt_pow64_det (include/trt_detmath64.h); the host build of
the same source already equals the oracle bit for bit (tests/test_diffusive_replica.py), so any difference here is a
device-side scheduling or code-generation problem.  The north_star tolerance (1e-5 relative) is asserted as well, against
the oracle's platform-libm build (what a gfortran build of the reference computes), on the cases where the reference itself
is insensitive to its libm (see test_diffusive_oracle.py::test_libm_and_bit_specified_pow_builds_agree).

(File name: runs after the Muskingum-Cunge GPU tests; this path had no GPU run in round 1.)"""
import numpy as np
import pytest

import helpers_diffusive as HD

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(600, method="thread")]
REL_TOL = 1e-5


@pytest.fixture(scope="module")
def gpu():
    import __graft_entry__ as g
    g.build()
    from troute_b200 import _lib
    assert _lib.lib().trt_device_count() >= 1, "no CUDA device: the diffusive path has no CPU fallback"
    from troute_b200.routing.fast_reach import diffusive
    return diffusive


@pytest.fixture(scope="module")
def od():
    from oracle import diffusive
    diffusive.build()
    return diffusive


@pytest.mark.parametrize("case", sorted(HD.CASES))
def test_gpu_equals_oracle_bit_for_bit(gpu, od, case):
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(**HD.CASES[case])
    ref = od.compute_diffusive(d, od.POW_DET)
    got = gpu.compute_diffusive(d)
    for name, a, b in zip(("q_ev_g", "elv_ev_g", "depth_ev_g"), ref, got):
        HD.assert_bits64(b, a, f"{case}: {name}")
    table_ms, loop_ms, launches = gpu.last_run()
    assert launches == 4 and loop_ms > 0.0


@pytest.mark.parametrize("case", ["small", "tailwater-depth", "flashy"])
def test_gpu_within_tolerance_of_the_libm_build(gpu, od, case):
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(**HD.CASES[case])
    ref = od.compute_diffusive(d, od.POW_LIBM)
    got = gpu.compute_diffusive(d)
    m = HD.mainstem_nodes(d)
    for a, b in zip(ref, got):
        rel = np.abs(a[:, m] - b[:, m]) / np.maximum(np.abs(a[:, m]), 1e-30)
        assert rel.max() < REL_TOL, rel.max()


@pytest.mark.parametrize("case", ["small", "branched", "tailwater-depth"])
def test_surveyed_cross_sections_on_gpu(gpu, od, case):
    """mxnbathy_g > 0 (what the reference's LowerColorado hybrid configuration uses, use_natl_xsections: True): the vertex
    roughness, table and smoothing kernels, then the same time loop -- bit-equal to the oracle."""
    from troute_b200 import synth_diffusive as sd
    d = sd.with_natural_sections(sd.diffusive_domain(**HD.CASES[case]))
    ref = od.compute_diffusive(d, od.POW_DET)
    got = gpu.compute_diffusive(d)
    for name, a, b in zip(("q_ev_g", "elv_ev_g", "depth_ev_g"), ref, got):
        HD.assert_bits64(b, a, f"{case}: {name}")
    assert gpu.last_run()[2] == 6


def test_uniform_flow_on_gpu(gpu):
    from troute_b200 import synth_diffusive as sd
    d = sd.uniform_channel(q=60.0)
    q, elv, dep = gpu.compute_diffusive(d)
    m = HD.mainstem_nodes(d)
    assert np.abs(q[:, m] - 60.0).max() < 1e-9
    assert dep[:, m].max() - dep[:, m].min() < 1e-6


def test_batch_of_domains_equals_single_calls(gpu, od):
    """One CTA per domain, 12 domains of different shapes in one launch: every domain's result equals its own single call
    and the oracle."""
    from troute_b200 import synth_diffusive as sd
    doms = [sd.diffusive_domain(n_mainstem=3 + k % 5, nodes=(3, 6 + k % 4), n_branch=k % 3, nsteps=36 + 12 * (k % 3), seed=100 + k,
                                dsbc_option=1 + k % 2) for k in range(12)]
    batch = gpu.compute_diffusive_batch(doms)
    assert len(batch) == len(doms)
    for k, (d, got) in enumerate(zip(doms, batch)):
        ref = od.compute_diffusive(d, od.POW_DET)
        for name, a, b in zip(("q", "elv", "depth"), ref, got):
            HD.assert_bits64(b, a, f"domain {k}: {name}")
    single = gpu.compute_diffusive(doms[5])
    for a, b in zip(single, batch[5]):
        HD.assert_bits64(a, b, "single call vs batch")


def test_repeated_calls_are_deterministic(gpu):
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(**HD.CASES["branched"])
    a = gpu.compute_diffusive(d)
    b = gpu.compute_diffusive(d)
    for x, y in zip(a, b):
        HD.assert_bits64(x, y, "second call")


def test_compute_diffusive_routing_on_gpu(gpu, od):
    """The reference-level entry point (compute.py:1740-1884 mirrored): two tailwater domains recorded from the reference's
    own packer tests, junction inflows taken from Muskingum-Cunge style results, packed by
    troute_b200.routing.diffusive_utils, routed in one trt_diffnw_batch launch, unpacked -- equal to the oracle run on the
    same packed inputs."""
    import datetime
    import pandas as pd
    import test_diffusive_packer as TP
    from troute_b200.routing import compute, diffusive_utils
    gold = np.load(TP.GOLD)
    dnd, results, q0, qlats, nsteps = TP.hybrid_case(gold)
    out = compute.compute_diffusive_routing(results, dnd, None, datetime.datetime(2023, 4, 2), 300.0, nsteps, q0, qlats, 12,
                                            pd.DataFrame(), pd.DataFrame(), {}, pd.DataFrame(), pd.DataFrame(), None, None,
                                            pd.DataFrame(), pd.DataFrame())
    assert len(out) == len(dnd)
    for (tw, net), tup in zip(dnd.items(), out):
        fvd = np.zeros((len(net["tributary_segments"]), 3 * nsteps), dtype=np.float32)
        ji = pd.DataFrame(np.concatenate([r[1][np.isin(r[0], net["tributary_segments"])][:, ::3] for r in results]),
                          index=np.concatenate([r[0][np.isin(r[0], net["tributary_segments"])] for r in results]))
        dq = qlats.copy(); dq.columns = range(dq.shape[1])
        ins = diffusive_utils.diffusive_input_data_v02(
            tw, net["connections"], net["rconn"], net["reaches"], net["mainstem_segs"], net["tributary_segments"], None,
            net["param_df"], dq, q0, ji, 12, datetime.datetime(2023, 4, 2), nsteps, 300.0, pd.DataFrame(), pd.DataFrame(),
            pd.DataFrame(), None, None, pd.DataFrame(), pd.DataFrame())
        ref_q, _, ref_depth = od.compute_diffusive(ins, od.POW_DET)
        ids, dat = diffusive_utils.unpack_output(ins["pynw"], ins["ordered_reaches"], ref_q, ref_depth)
        keep = ~np.isin(ids, net["tributary_segments"])
        assert ids[keep].tolist() == tup[0].tolist()
        assert np.array_equal(dat[keep][:, 3:], tup[1], equal_nan=True)


def test_lowercolorado_shipped_hybrid_configuration_on_gpu(gpu, od, oracle):
    """The shipped hybrid configuration (use_natl_xsections: True): the coastal domain with the surveyed cross sections of the
    hydrofabric's cross-section table (up to 500 vertices per section), truncated where the sections end, through
    compute_diffusive_routing on the device == the oracle (surveyed-section table kernels + time loop on real data)."""
    import datetime
    import pandas as pd
    import test_lowercolorado_hybrid as LH
    from troute_b200.routing import compute, diffusive_utils
    c, dnd, results, q0, qlats, topo, bad, _ = LH.natural_inputs(oracle)
    out = compute.compute_diffusive_routing(results, dnd, None, datetime.datetime(2023, 4, 2), 300.0, LH.NTS, q0, qlats, 12,
                                            pd.DataFrame(), pd.DataFrame(), {}, pd.DataFrame(), topo, None, None,
                                            pd.DataFrame(), pd.DataFrame())
    ins = LH.pack_natural(dnd, results, q0, qlats, topo)
    ref_q, _, ref_depth = od.compute_diffusive(ins, od.POW_DET)
    ids, dat = diffusive_utils.unpack_output(ins["pynw"], ins["ordered_reaches"], ref_q, ref_depth)
    keep = ~np.isin(ids, dnd[LH.TW]["tributary_segments"])
    assert ids[keep].tolist() == out[0][0].tolist()
    assert np.array_equal(dat[keep][:, 3:], out[0][1], equal_nan=True)


def test_lowercolorado_hybrid_domain_on_gpu(gpu, od, oracle):
    """BASELINE configs[3] on the real hydrofabric: the coastal diffusive domain of LowerColorado_TX_v4 (787 mainstem
    segments, 640 reaches, 7 Muskingum-Cunge tributaries) through compute_diffusive_routing on the device == the oracle."""
    import datetime
    import pandas as pd
    import test_lowercolorado_hybrid as LH
    from troute_b200.routing import compute, diffusive_utils
    c, dnd, results, q0, qlats, _, _ = LH.hybrid_inputs(oracle)
    out = compute.compute_diffusive_routing(results, dnd, None, datetime.datetime(2023, 4, 2), 300.0, LH.NTS, q0, qlats, 12,
                                            pd.DataFrame(), pd.DataFrame(), {}, pd.DataFrame(), pd.DataFrame(), None, None,
                                            pd.DataFrame(), pd.DataFrame())
    ins = LH.pack(dnd, results, q0, qlats)
    ref_q, _, ref_depth = od.compute_diffusive(ins, od.POW_DET)
    ids, dat = diffusive_utils.unpack_output(ins["pynw"], ins["ordered_reaches"], ref_q, ref_depth)
    keep = ~np.isin(ids, dnd[LH.TW]["tributary_segments"])
    assert ids[keep].tolist() == out[0][0].tolist()
    assert np.array_equal(dat[keep][:, 3:], out[0][1], equal_nan=True)


def test_lowercolorado_hybrid_end_to_end_on_gpu(gpu, od, oracle):
    """BASELINE configs[3], both halves on the device through the reference-level entry points, the way nwm_route chains
    them (__main__.py:1215-1290): compute_nhd_routing_v02 on the network without the diffusive mainstem, then
    compute_diffusive_routing fed by its results.  MC half bit-equal to the MC oracle, diffusive half equal to the diffusive
    oracle run on the same junction inflows."""
    import datetime
    import pandas as pd
    import test_lowercolorado as LC
    import test_lowercolorado_hybrid as LH
    from troute_b200.routing import compute, diffusive_utils
    from troute_b200.routing.fast_reach.mc_reach import clear_network_cache
    c, dnd, _, q0, qlats, df_mc, conn_mc = LH.hybrid_inputs(oracle)
    sub, reaches_bytw, indep = LH.reduced_mc_case(c, conn_mc, df_mc)
    ids = sub["ids"]
    param_df = pd.DataFrame(sub["params"][:, 1:], index=ids, columns=sub["cols"][1:])
    # one initial-condition frame for both halves: cold start on the MC network, 0.5 m3/s on the diffusive mainstem
    q0 = q0.copy()
    q0.loc[ids.tolist()] = 0.0
    empty = pd.DataFrame()
    t0 = datetime.datetime(2023, 4, 2)
    # one call of the driver-level entry point routes both halves; the diffusive half needs forcing / initial state of the
    # mainstem segments too, so the full-network frames are passed (rows outside the MC network are ignored by the MC half)
    from troute_b200.nwm_routing import nwm_route
    try:
        both, _ = nwm_route(
            sub["connections"], sub["rconn"], {}, reaches_bytw, "by-network", "V02-structured", 10000, 1, t0, 300.0, LH.NTS, 12,
            indep, param_df, q0.astype(np.float32), qlats, empty, empty, empty, empty, empty, empty, empty, empty, empty, empty,
            empty, {}, True, False, empty, {}, empty, False, dnd, empty, None, None, [None, None], empty, empty)
    finally:
        clear_network_cache()
    results, out = both[:-1], both[-1:]
    # MC half vs the MC oracle (same reduced network, first LH.NTS steps)
    ref_mc = LC._oracle_call(oracle, sub, True)
    ref_fvd = ref_mc[1].reshape(ids.shape[0], LC.NTS, 3)[:, :LH.NTS, :].reshape(ids.shape[0], -1)
    got_ids = np.concatenate([r[0] for r in results]); got_fvd = np.concatenate([r[1] for r in results])
    order = np.argsort(got_ids)
    assert np.array_equal(got_ids[order], ref_mc[0])
    assert np.array_equal(got_fvd[order].view(np.int32), ref_fvd.view(np.int32))
    # diffusive half
    ins = LH.pack(dnd, [(ref_mc[0], ref_fvd, 0)], q0, qlats)
    ref_q, _, ref_depth = od.compute_diffusive(ins, od.POW_DET)
    seg_ids, dat = diffusive_utils.unpack_output(ins["pynw"], ins["ordered_reaches"], ref_q, ref_depth)
    keep = ~np.isin(seg_ids, dnd[LH.TW]["tributary_segments"])
    assert seg_ids[keep].tolist() == out[0][0].tolist()
    assert np.array_equal(dat[keep][:, 3:], out[0][1], equal_nan=True)
