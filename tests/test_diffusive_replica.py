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
    """trt_c_diffnw validates on the host first: natural cross sections and the refactored-hydrofabric crosswalk are
    refused with TRT_ERR_INVALID (ValueError), not silently computed as something else."""
    import __graft_entry__ as g
    g.build()
    from troute_b200 import synth_diffusive as sd
    from troute_b200.routing.fast_reach import diffusive
    d = sd.diffusive_domain()
    d["mxnbathy_g"] = 3
    with pytest.raises(ValueError, match="natural cross sections"):
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
