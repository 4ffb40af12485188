"""The product's per-lane Muskingum-Cunge / level-pool device source (t-route_b200/csrc/mc_device.cuh) compiled for the
HOST (tests/native/mc_replica.cpp) against the oracle's bit-specified-powf build: bit equality on the reference's 5000-row
randomized suite, on random rows over the CONUS parameter ranges and on the level-pool known-answer fixtures -- both as the
one-call solve the dataflow kernel uses and as the prepare / begin / iterate / outflow / velocity pieces the marching
kernel interleaves.  This is the no-GPU gate for changes to the device physics; the GPU tests repeat it on the device."""
import ctypes as C
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "mc_replica.cpp")
LIB = os.path.join(HERE, "native", "libmc_replica.so")
DEPS = [SRC, os.path.join(HERE, "native", "shim", "cuda_runtime.h"), os.path.join(ROOT, "t-route_b200", "csrc", "mc_device.cuh"),
        os.path.join(ROOT, "include", "trt_detmath.h"), os.path.join(ROOT, "include", "trt_detmath_tables.h")]
GOLD = os.path.join(HERE, "golden")


def _build(lib, extra=()):
    if not os.path.exists(lib) or any(os.path.getmtime(d) > os.path.getmtime(lib) for d in DEPS):
        flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
        cmd = ["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-I", os.path.join(HERE, "native", "shim"),
               "-shared", "-fPIC", "-o", lib, SRC, *extra]
        if " fma" in flags:
            cmd.insert(1, "-mfma")
        subprocess.run(cmd, check=True)
    return C.CDLL(lib)


@pytest.fixture(scope="module")
def rep():
    return _build(LIB)


@pytest.fixture(scope="module", params=[-1, 0, 1])
def rep_seed(request):
    """the same source with the reciprocal seed of the fast-path division moved by -1 / 0 / +1 ulp: the device's MUFU.RCP is
    within one ulp of 1/d, and the sequence must not care which way"""
    u = request.param
    return _build(os.path.join(HERE, "native", f"libmc_replica_seed{u + 1}.so"), (f"-DTRT_RCP_SEED_ULPS=({u})",))


def run_rows(rep, rows, resumable):
    rows = np.ascontiguousarray(rows, dtype=np.float32).reshape(-1, 15)
    out = np.zeros((rows.shape[0], 6), dtype=np.float32)
    iters = np.zeros(rows.shape[0], dtype=np.int32)
    rep.trt_replica_mc_segment_batch(C.c_long(rows.shape[0]), rows.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p),
                                     iters.ctypes.data_as(C.c_void_p), C.c_int(int(resumable)))
    return out, iters


def random_rows(n, seed):
    """dt,qup,quc,qdp,ql,dx,bw,tw,twcc,n,ncc,cs,s0,velp,depthp over the ranges of test_suite_parameters.py:4-13, with dry
    segments, zero side slope, bw >= tw and flat beds mixed in"""
    rng = np.random.default_rng(seed)
    r = np.empty((n, 15), dtype=np.float32)
    r[:, 0] = 300.0
    r[:, 1:4] = np.exp(rng.uniform(np.log(1e-4), np.log(3e3), (n, 3)))
    r[:, 4] = np.exp(rng.uniform(np.log(1e-6), np.log(50.0), n))
    r[:, 5] = np.exp(rng.uniform(np.log(1.0), np.log(9.5e4), n))
    r[:, 6] = np.exp(rng.uniform(np.log(0.135), np.log(230.0), n))
    r[:, 7] = r[:, 6] / 0.6
    r[:, 8] = 3 * r[:, 7]
    r[:, 9] = rng.uniform(0.04, 0.06, n); r[:, 10] = 2 * r[:, 9]
    r[:, 11] = rng.uniform(0.085, 2.25, n)
    r[:, 12] = np.exp(rng.uniform(np.log(1e-5), np.log(0.5), n))
    r[:, 13] = 0.0
    r[:, 14] = np.exp(rng.uniform(np.log(1e-3), np.log(12.0), n))
    k = n // 20
    r[0 * k:1 * k, 1:5] = 0.0                    # no flow at all (:171-178)
    r[1 * k:2 * k, 11] = 0.0                     # cs = 0 -> z = 1 (:49-53)
    r[2 * k:3 * k, 7] = r[2 * k:3 * k, 6] * 0.5  # bw > tw (:55-57)
    r[3 * k:4 * k, 7] = r[3 * k:4 * k, 6]        # bw == tw
    r[4 * k:5 * k, 8] = 0.0                      # no floodplain width (:400-403)
    r[5 * k:6 * k, 14] = 0.0                     # dry start
    return r


@pytest.mark.parametrize("resumable", [0, 1, 2])
def test_reference_suite_rows(rep, oracle, resumable):
    rows = np.load(os.path.join(GOLD, "mc_suite_seed16.npy"))
    want, wi = oracle.mc_segment_batch(rows, oracle.POW_DET)
    got, gi = run_rows(rep, rows, resumable)
    cols = [0, 1, 2, 5] if resumable == 1 else [0, 1, 2, 3, 4, 5]        # the marching pieces never evaluate the Courant diagnostics
    assert np.array_equal(got[:, cols].view(np.int32), want[:, cols].view(np.int32))
    assert np.array_equal(gi, wi)


@pytest.mark.parametrize("resumable", [0, 1, 2])
def test_random_rows_over_the_parameter_ranges(rep, oracle, resumable):
    rows = random_rows(200_000, 7)
    want, wi = oracle.mc_segment_batch(rows, oracle.POW_DET)
    got, gi = run_rows(rep, rows, resumable)
    cols = [0, 1, 2, 5] if resumable == 1 else [0, 1, 2, 3, 4, 5]
    bad = (got[:, cols].view(np.int32) != want[:, cols].view(np.int32)).any(axis=1)
    assert not bad.any(), (int(bad.sum()), rows[bad][0].tolist(), got[bad][0].tolist(), want[bad][0].tolist())
    assert np.array_equal(gi, wi) and wi.max() > 5               # the retry ladder is reached


def test_levelpool_kats(rep, oracle):
    k = json.load(open(os.path.join(GOLD, "levelpool_kats.json")))
    for c in k["cases"]:
        a = np.ascontiguousarray(c["wbody_row"], dtype=np.float64)
        inflow = np.ascontiguousarray(c["inflow"], dtype=np.float32)
        q = np.zeros_like(inflow); h = np.zeros_like(inflow)
        rep.trt_replica_levelpool_series(a.ctypes.data_as(C.c_void_p), C.c_long(inflow.size), inflow.ctypes.data_as(C.c_void_p),
                                         C.c_float(0.0), C.c_float(c["routing_period"]), q.ctypes.data_as(C.c_void_p),
                                         h.ctypes.data_as(C.c_void_p))
        wq, wh = oracle.levelpool_series(c["wbody_row"], c["inflow"], 0.0, c["routing_period"], pow_mode=oracle.POW_DET)
        assert np.array_equal(q.view(np.int32), wq.view(np.int32)) and np.array_equal(h.view(np.int32), wh.view(np.int32))
        assert q[-1] == np.float32(c["expected_final_outflow"]) and h[-1] == np.float32(c["expected_final_water_elevation"])


def test_fastpath_division_is_the_ieee_quotient(rep_seed):
    """McDivFast (mc_device.cuh): inside its window of exponents the inline fast path -- reciprocal seed, one Newton step,
    quotient, exact remainder, correction -- returns the correctly rounded quotient, whatever the last bit of the seed."""
    rng = np.random.default_rng(5)
    n = 20_000_000
    total_in = 0
    for kind in range(3):
        if kind == 0:        # anything inside (and a little outside) the window
            a = (rng.standard_normal(n) * np.exp2(rng.uniform(-64, 64, n))).astype(np.float32)
            d = (rng.standard_normal(n) * np.exp2(rng.uniform(-64, 64, n))).astype(np.float32)
        elif kind == 1:      # quotients next to rounding boundaries: small integers over small integers, scaled
            a = (rng.integers(1, 1 << 24, n).astype(np.float32)) * np.exp2(rng.integers(-30, 30, n)).astype(np.float32)
            d = (rng.integers(1, 1 << 12, n).astype(np.float32)) * np.exp2(rng.integers(-30, 30, n)).astype(np.float32)
        else:                # the magnitudes of the solve: flows, areas, depths, dt / dx
            a = np.exp(rng.uniform(np.log(1e-8), np.log(1e8), n)).astype(np.float32) * rng.choice([-1.0, 1.0], n).astype(np.float32)
            d = np.exp(rng.uniform(np.log(1e-8), np.log(1e8), n)).astype(np.float32)
            a[: n // 50] = 0.0
            a[n // 50: n // 25] = -0.0
        inside = C.c_long(0)
        rep_seed.trt_replica_fdiv_check.restype = C.c_long
        bad = rep_seed.trt_replica_fdiv_check(C.c_long(n), a.ctypes.data_as(C.c_void_p), d.ctypes.data_as(C.c_void_p), C.byref(inside))
        assert bad == 0, (kind, bad)
        total_in += inside.value
    assert total_in > 0.8 * 3 * n


def test_marching_pieces_with_fastpath_division(rep_seed, oracle):
    """the marching kernel's form of the solve with McDivFast (restart with IEEE divisions when a trip leaves the window) ==
    oracle, bit for bit, and nearly every row stays on the fast path"""
    rows = np.concatenate([np.load(os.path.join(GOLD, "mc_suite_seed16.npy")), random_rows(100_000, seed=21)])
    want, wi = oracle.mc_segment_batch(rows, oracle.POW_DET)
    got, gi = run_rows(rep_seed, rows, 3)
    assert np.array_equal(got[:, [0, 1, 2, 5]].view(np.int32), want[:, [0, 1, 2, 5]].view(np.int32))
    assert np.array_equal(gi, wi)
    assert got[:, 3].mean() > 0.9, float(got[:, 3].mean())
