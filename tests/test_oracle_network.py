"""Network-level behaviour of the CPU oracle (CPU only): loop-order independence, reach grouping independence, the
by-subnetwork job decomposition, and the measured distance between its two arithmetic builds."""
import numpy as np
import pytest

import helpers as H
from troute_b200 import hostgraph, synth


@pytest.fixture(scope="module")
def case():
    return H.make_case(synth.hack_tree(6000, seed=4), nsteps=30, n_lp=6, warm=True)


@pytest.mark.parametrize("short_ts", [False, True])
def test_reach_grouping_does_not_change_results(oracle, case, short_ts):
    """One-segment reaches in level order vs the reference's multi-segment reaches (chains between junctions): the
    segment walk inside a reach (mc_reach.pyx:133-138) is the same data flow as an upstream connection."""
    a, upa, _ = H.oracle_route(oracle, case, short_ts)
    down = case["down"]
    is_lp = case["kind"] == 1
    up_ptr, up_rows = case["up_ptr"], case["up_rows"]
    indeg = np.diff(up_ptr)
    level = synth.levels_from_down(down)

    def starts(i):
        return indeg[i] != 1 or is_lp[i] or is_lp[up_rows[up_ptr[i]]]
    reaches = []
    for h in sorted((i for i in range(down.size) if starts(i)), key=lambda i: (int(level[i]), i)):
        r = [h]
        cur = h
        while not is_lp[h]:
            d = int(down[cur])
            if d < 0 or starts(d):
                break
            r.append(d); cur = d
        reaches.append(r)
    assert max(len(r) for r in reaches) > 1
    ups = {i: up_rows[up_ptr[i]:up_ptr[i + 1]].tolist() for i in range(down.size)}
    lakes = case["lp_rows"].tolist()
    flat = oracle.flatten_reaches([(r, 1 if is_lp[r[0]] else 0) for r in reaches], ups, np.arange(down.size), lakes)
    b, upb, _ = H.oracle_route(oracle, case, short_ts, reaches=flat)
    H.assert_bit_equal(a, b, "reach grouping")
    H.assert_bit_equal(upa[case["lp_rows"]], upb[case["lp_rows"]], "reservoir inflow")


def test_subnetwork_job_order_equals_reference_order(oracle):
    """Routing ALL timesteps of an upstream sub-network before the downstream one (compute.py:975-1199) gives the
    same bits as the time-outer loop (mc_reach.pyx:492-493) -- the property the wavefront schedule relies on."""
    c = H.make_case(synth.conus_like(n_total=20000, n_basins=25, seed=6), nsteps=24, warm=False)
    a, _, _ = H.oracle_route(oracle, c, False)
    r = hostgraph.segment_reaches_level_order(c["down"], c["up_ptr"], c["up_rows"])
    jobs = hostgraph.subnetwork_jobs(c["down"], c["up_ptr"], c["up_rows"], r["order"], target_size=1500)
    assert jobs["n_orders"] > 1
    b, _, _ = H.oracle_route(oracle, c, False, jobs=jobs, nthreads=4)
    H.assert_bit_equal(a, b, "job decomposition")


def test_libm_vs_det_distance_is_bounded(oracle):
    """How far the reference's arithmetic moves when powf changes by <= 1 ulp (platform libm vs trt_powf_det)."""
    c = H.make_case(synth.hack_tree(20000, seed=9), nsteps=48, warm=False)
    a, _, ea = H.oracle_route(oracle, c, False, pow_mode=oracle.POW_LIBM)
    b, _, eb = H.oracle_route(oracle, c, False, pow_mode=oracle.POW_DET)
    qa, qb = a[:, 0::3], b[:, 0::3]
    rel = np.abs(qa - qb) / np.maximum(np.abs(qa), 1e-3)
    assert (qa == qb).mean() >= 0.95
    assert (rel <= 1e-5).mean() >= 0.99
    assert np.median(rel) == 0.0
    # iteration histograms are nearly identical: the flips are rare events, not a systematic shift
    ha, hb = ea["iter_hist"].astype(float), eb["iter_hist"].astype(float)
    assert np.abs(ha - hb).sum() / ha.sum() < 1e-3


def test_qlat_column_check(oracle, case):
    with pytest.raises(ValueError):
        c = dict(case); c["qlat"] = case["qlat"][:, :1]
        H.oracle_route(oracle, c, False)
