"""oracle/diffusive_oracle.c -- the CPU restatement of diffusive.f90.  PARITY UNPINNED (the reference ships no vectors for
this solver and its Fortran cannot be built here); what is checked are properties the PDE and the reference's own
construction guarantee."""
import numpy as np
import pytest

import helpers_diffusive as HD


@pytest.fixture(scope="module")
def od():
    from oracle import diffusive
    diffusive.build()
    return diffusive


def test_uniform_flow_is_a_fixed_point(od):
    """A prismatic channel at normal depth carrying a constant flow: the diffusive wave must hold it (Q = const, depth =
    Manning's normal depth at every node, every output time)."""
    from troute_b200 import synth_diffusive as sd
    q0, slope = 60.0, 8e-4
    d = sd.uniform_channel(q=q0, slope=slope)
    q, elv, dep = od.compute_diffusive(d)
    m = HD.mainstem_nodes(d)
    assert np.abs(q[:, m] - q0).max() < 1e-9
    y = dep[:, m]
    assert y.max() - y.min() < 1e-6
    # Manning in the trapezoid (bed lowered by the 1 cm notch the table builder cuts, diffusive.f90:2253-2254)
    h = float(np.median(y)) - 0.01
    bw, z, n = 40.0, 2.0, 0.035
    area = (bw + z * h) * h
    peri = bw + 2 * h * np.sqrt(1 + z * z)
    q_manning = area * (area / peri) ** (2 / 3) * np.sqrt(slope) / n
    assert abs(q_manning / q0 - 1) < 1e-2          # notch triangle + linear interpolation in a 501-row table


def test_constant_forcing_converges_to_the_sum_of_inflows(od):
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(n_mainstem=5, nsteps=1440, pulse=0.0, seed=4)
    d["qlat_g"][:] = d["qlat_g"][0]                                # steady lateral inflow
    q, elv, dep = od.compute_diffusive(d)
    tribs = [j for j in range(d["nrch_g"]) if j not in d["mainstem"]]
    inflow = d["qtrib_g"][0, tribs].sum()
    for j in d["mainstem"]:
        n = d["frnw_g"][j, 0]
        inflow += (d["qlat_g"][0, : n - 1, j] * d["dx_ar_g"][: n - 1, j]).sum()
    last = d["mainstem"][-1]
    out = q[-1, d["frnw_g"][last, 0] - 1, last]
    assert abs(out / inflow - 1) < 2e-3, (out, inflow)
    assert np.isfinite(q).all() and np.isfinite(dep).all()


@pytest.mark.parametrize("case", ["small", "tailwater-depth", "flashy"])
def test_flood_wave_is_attenuated_and_depths_are_positive(od, case):
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(**HD.CASES[case])
    q, elv, dep = od.compute_diffusive(d)
    m = HD.mainstem_nodes(d)
    assert np.isfinite(q[:, m]).all() and (dep[1:, m] > 0).all() and (q[:, m] >= d["para_ar_g"][7]).all()
    tribs = [j for j in range(d["nrch_g"]) if j not in d["mainstem"]]
    last = d["mainstem"][-1]
    out = q[:, d["frnw_g"][last, 0] - 1, last]
    assert out.max() < d["qtrib_g"][:, tribs].sum(axis=1).max() * 1.05     # no amplification of the inflow peak
    # tributary rows carry the tributary hydrograph at the save times (diffusive.f90:611-633)
    t = tribs[0]
    np.testing.assert_allclose(q[:, 0, t], d["qtrib_g"][:, t], rtol=1e-12)


def test_every_output_row_is_written(od):
    """The adaptive step (calculateDT, diffusive.f90:942-991) is clipped so that every save time is hit exactly: all
    ntss_ev_g rows of the mainstem nodes are written, the first one with the initial state (:801-821, t0 = 0)."""
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(nsteps=48)
    q, elv, dep = od.compute_diffusive(d)
    m = HD.mainstem_nodes(d)
    assert (q[:, m] != 0).all() and (elv[:, m] != 0).all()
    np.testing.assert_allclose(q[0][m], d["iniq"][m], rtol=0, atol=0)


def test_libm_and_bit_specified_pow_builds_agree(od):
    """The reference's own sensitivity to its libm: the two arithmetic builds differ in the last bits of pow only."""
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(**HD.CASES["small"])
    a = od.compute_diffusive(d, od.POW_LIBM)
    b = od.compute_diffusive(d, od.POW_DET)
    m = HD.mainstem_nodes(d)
    for x, y in zip(a, b):
        rel = np.abs(x[:, m] - y[:, m]) / np.maximum(np.abs(x[:, m]), 1e-30)
        assert rel.max() < 1e-9, rel.max()


def test_inconsistent_cross_section_counts_are_refused(od):
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain()
    d["mxnbathy_g"] = 4                                                 # surveyed sections announced, none given
    with pytest.raises(RuntimeError):
        od.compute_diffusive(d)


def test_crosswalk_onto_the_same_hydrofabric_is_the_identity(od):
    """diffusive.f90:837-903 maps results from a refactored hydrofabric back by linear interpolation along every refactored
    segment; a crosswalk that maps every segment onto itself (one link, fraction 1) must reproduce flows and elevations."""
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(nsteps=24)
    plain = od.compute_diffusive(d)
    rows = []
    for j in d["mainstem"]:
        for i in range(d["frnw_g"][j, 0] - 1):
            rows.append([i + 1, j + 1, 1, i + 1, j + 1, 1.0])
    d2 = dict(d)
    d2["crosswalk_g"] = np.asarray(rows, dtype=np.float64)
    d2["cwnrow_g"], d2["cwncol_g"] = d2["crosswalk_g"].shape
    cw = od.compute_diffusive(d2)
    m = HD.mainstem_nodes(d)
    np.testing.assert_allclose(cw[0][:, m], plain[0][:, m], rtol=1e-12)
    # elevations come back on the thalweg of the ORIGINAL bed (z_thalweg_g), i.e. without the 1 cm notch
    np.testing.assert_allclose(cw[1][:, m], plain[1][:, m] + 0.01, rtol=0, atol=1e-6)
