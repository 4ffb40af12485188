"""The product's diffusive solver source (t-route_b200/csrc/diffusive_device.cuh) compiled for the HOST and run
single-threaded (tests/native/diffusive_replica.cpp) against the oracle's bit-specified-pow build: bit equality.

This is what can be checked without a GPU: the table construction, the arithmetic row look-ups that replace the reference's
linear scans and bisections, the hoisted Newton invariants and the re-ordered sweeps of dw_time_loop compute exactly what
the reference's loop order computes.  tests/test_zz_gpu_diffusive.py repeats the comparison through the C ABI on the
device."""
import ctypes as C
import os

import numpy as np
import pytest

import helpers_diffusive as HD


@pytest.fixture(scope="module")
def od():
    from oracle import diffusive
    diffusive.build()
    return diffusive


@pytest.mark.parametrize("case", sorted(HD.CASES))
def test_host_build_of_the_solver_equals_the_oracle(od, case):
    from troute_b200 import synth_diffusive as sd
    d = sd.diffusive_domain(**HD.CASES[case])
    ref = od.compute_diffusive(d, od.POW_DET)
    got = HD.replica_compute_diffusive(d)
    for name, a, b in zip(("q_ev_g", "elv_ev_g", "depth_ev_g"), ref, got):
        HD.assert_bits64(b, a, f"{case}: {name}")


def test_uniform_channel_through_the_solver_source(od):
    from troute_b200 import synth_diffusive as sd
    d = sd.uniform_channel()
    ref = od.compute_diffusive(d, od.POW_DET)
    got = HD.replica_compute_diffusive(d)
    for a, b in zip(ref, got):
        HD.assert_bits64(b, a, "uniform channel")


def test_locate_with_a_hint_is_the_bisection():
    """dw_locate_hint (galloping from a guess) == the reference's `locate` (diffusive.f90:2701-2742) for every hint, on
    columns with ties, and for arguments on, between, below and above the rows."""
    lib = HD.replica_lib()
    if not hasattr(lib, "trt_replica_locate"):
        pytest.skip("replica built without the locate probe")
    lib.trt_replica_locate.restype = C.c_int
    rng = np.random.default_rng(5)

    def bisect_ref(xx, x):                       # the Fortran, transcribed
        n = len(xx)
        ascnd = xx[-1] >= xx[0]
        jl, ju = 0, n + 1
        while ju - jl > 1:
            jm = (ju + jl) // 2
            if ascnd == (x >= xx[jm - 1]):
                jl = jm
            else:
                ju = jm
        if x == xx[0]:
            return 1
        if x == xx[-1]:
            return n - 1
        return jl

    for n in (2, 3, 17, 501):
        xx = np.sort(rng.uniform(0, 10, n))
        if n > 5:
            xx[3] = xx[2]                        # a tie
        probes = np.concatenate([xx, (xx[:-1] + xx[1:]) / 2, [xx[0] - 1, xx[-1] + 1, np.nan]])
        for x in probes:
            want = bisect_ref(xx, x)
            for hint in (-3, 1, n // 2, n - 1, n + 5):
                got = lib.trt_replica_locate(xx.ctypes.data_as(C.c_void_p), C.c_int(n), C.c_double(x), C.c_int(hint))
                assert got == want, (n, x, hint, got, want)


def test_c_abi_refuses_unsupported_inputs_before_touching_the_device():
    """trt_c_diffnw validates on the host first: inconsistent cross-section counts, the refactored-hydrofabric crosswalk
    and a network without a mainstem are refused with TRT_ERR_INVALID (ValueError), not silently computed as something
    else."""
    import __graft_entry__ as g
    g.build()
    from troute_b200 import synth_diffusive as sd
    from troute_b200.routing.fast_reach import diffusive
    d = sd.diffusive_domain()
    d["mxnbathy_g"] = 3                                                 # surveyed sections announced, none given
    with pytest.raises(ValueError, match="size_bathy_g"):
        diffusive.compute_diffusive(d)
    d = sd.diffusive_domain()
    d["cwnrow_g"], d["cwncol_g"], d["crosswalk_g"] = 1, 6, np.ones((1, 6))
    with pytest.raises(ValueError, match="crosswalk"):
        diffusive.compute_diffusive(d)
    d = sd.diffusive_domain()
    d["frnw_g"] = d["frnw_g"].copy()
    d["frnw_g"][:, 3:] = 0                                              # no reach flagged 555
    d["frnw_g"][:, 2] = 0
    with pytest.raises(ValueError, match="no mainstem"):
        diffusive.compute_diffusive(d)


@pytest.mark.parametrize("case", ["small", "branched", "tailwater-depth"])
def test_surveyed_cross_sections_host_build_equals_the_oracle(od, case):
    """mxnbathy_g > 0: tables from surveyed vertices (readXsection_natural_mann_vertices, diffusive.f90:1756-2091) --
    vertex roughness capped at 0.15, conveyance and dK/dA made monotone -- then the same time loop."""
    from troute_b200 import synth_diffusive as sd
    d = sd.with_natural_sections(sd.diffusive_domain(**HD.CASES[case]))
    ref = od.compute_diffusive(d, od.POW_DET)
    got = HD.replica_compute_diffusive(d)
    for name, a, b in zip(("q_ev_g", "elv_ev_g", "depth_ev_g"), ref, got):
        HD.assert_bits64(b, a, f"{case}: {name}")
    m = HD.mainstem_nodes(d)
    assert np.isfinite(ref[0][:, m]).all() and (ref[2][1:, m] > 0).all()


def test_surveyed_trapezoid_table_matches_the_geometry():
    """A surveyed section that IS a trapezoid with one roughness: area, wetted perimeter, top width and conveyance of
    every table row equal the closed-form values; the monotone smoothing leaves a monotone table alone."""
    lib = HD.replica_lib()
    lib.trt_replica_table_natural.restype = C.c_int
    bw, side, depth, n, z0 = 30.0, 2.0, 5.0, 0.04, 100.0
    x = np.array([0.0, side * depth, side * depth + bw, 2 * side * depth + bw]) + 1234.5
    z = np.array([z0 + depth, z0, z0, z0 + depth])
    mann = np.full(4, n)
    out = np.zeros((8, 501)); zout = C.c_double()
    lib.trt_replica_table_natural(4, x.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), mann.ctypes.data_as(C.c_void_p),
                                  out.ctypes.data_as(C.c_void_p), C.byref(zout))
    assert zout.value == z0
    el, area, peri, conv, topw, dkda, _, skk = out
    assert (np.diff(el) > 0).all() and abs(el[-1] - (z0 + 4 * depth)) < 1e-9        # timesDepth = 4 (:214)
    inb = el <= z0 + depth                                                         # rows inside the surveyed banks
    h = el[inb] - z0
    np.testing.assert_allclose(area[inb], (bw + side * h) * h, rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(topw[inb], bw + 2 * side * h, rtol=1e-10)
    p_full = bw + 2 * h * np.sqrt(1 + side * side)
    np.testing.assert_allclose(peri[inb], p_full, rtol=1e-10)
    # every wetted side carries n (a side takes the roughness of its LEFT vertex, :1905-1911), so the equivalent n is n
    e23 = np.float64(np.float32(2.0) / np.float32(3.0))
    np.testing.assert_allclose(skk[inb][1:], 1.0 / n, rtol=1e-6)
    np.testing.assert_allclose(conv[inb][1:], (area[inb] * (area[inb] / p_full) ** e23 / n)[1:], rtol=1e-6)
    assert (np.diff(conv) > 0).all() and (np.diff(dkda) >= 0).all()


def test_surveyed_section_with_rough_floodplains_is_made_monotone():
    """Wide rough floodplains make the single-section conveyance DROP when the water leaves the channel; the table the
    solver reads must be monotone in elevation (diffusive.f90:1951-2008), with roughness capped at 0.15."""
    lib = HD.replica_lib()
    lib.trt_replica_table_natural.restype = C.c_int
    x = np.array([0.0, 5.0, 400.0, 410.0, 440.0, 450.0, 900.0, 905.0])
    z = np.array([20.0, 13.2, 13.0, 10.0, 10.0, 13.0, 13.1, 20.0])
    mann = np.array([0.3, 0.3, 0.3, 0.03, 0.03, 0.03, 0.3, 0.3])
    out = np.zeros((8, 501)); zout = C.c_double()
    lib.trt_replica_table_natural(8, x.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), mann.ctypes.data_as(C.c_void_p),
                                  out.ctypes.data_as(C.c_void_p), C.byref(zout))
    el, area, peri, conv, topw, dkda, _, skk = out
    assert (np.diff(conv) >= 0).all() and (np.diff(dkda) >= 0).all()
    assert skk.min() >= 1 / 0.15 - 1e-9                                            # equivalent n never above the cap
    raw_like = area * (area / peri) ** (2 / 3) * skk                              # conveyance WITHOUT the smoothing
    assert (np.diff(raw_like) < 0).any()                                           # ... is not monotone here
