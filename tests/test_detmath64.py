"""The bit-specified binary64 pow of the diffusive path (include/trt_detmath64.h) against numpy / glibc pow."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def powdet():
    from oracle import diffusive as od
    L = od.lib()
    L.trt_oracle_pow64_det_array.restype = None
    def f(x, y):
        x = np.ascontiguousarray(x, dtype=np.float64); y = np.ascontiguousarray(y, dtype=np.float64)
        out = np.empty_like(x)
        L.trt_oracle_pow64_det_array(C.c_long(x.size), x.ctypes.data_as(C.c_void_p), y.ctypes.data_as(C.c_void_p),
                                     out.ctypes.data_as(C.c_void_p))
        return out
    return f


def test_accuracy_on_the_solver_exponents(powdet):
    rng = np.random.default_rng(0)
    n = 2_000_000
    x = np.exp(rng.uniform(np.log(1e-8), np.log(1e8), n))
    ys = np.array([3.0, np.float32(2.0) / np.float32(3.0), np.float32(0.3), np.float32(0.4), np.float32(0.6), 0.5],
                  dtype=np.float64)
    y = ys[rng.integers(0, ys.size, n)]
    rel = np.abs(powdet(x, y) - np.power(x, y)) / np.power(x, y)
    assert rel.max() < 4.5e-16, rel.max()          # <= 2 ulp of glibc's (nearly correctly rounded) pow


def test_wide_range_bound(powdet):
    rng = np.random.default_rng(1)
    n = 500_000
    x = np.exp(rng.uniform(np.log(1e-200), np.log(1e200), n))
    y = rng.uniform(-1.5, 1.5, n)
    y[y == 0] = 0.5
    ref = np.power(x, y)
    rel = np.abs(powdet(x, y) - ref) / ref
    bound = 2.0 ** -52 * (2.0 + np.abs(y * np.log(x)))          # the header's stated bound
    assert (rel <= bound).all(), float((rel / bound).max())


def test_special_values(powdet):
    x = np.array([0.0, 0.0, 0.0, 1.0, 2.0, np.inf, np.nan, -1.0, 5e-324, 4.0, 1e-310])
    y = np.array([3.0, 0.0, -1.0, 7.3, 0.0, 0.5, 0.5, 0.5, 0.5, 0.5, 0.5])
    got = powdet(x, y)
    assert got[0] == 0.0 and got[1] == 1.0 and np.isposinf(got[2]) and got[3] == 1.0 and got[4] == 1.0
    assert np.isposinf(got[5]) and np.isnan(got[6]) and np.isnan(got[7])
    assert abs(got[8] / np.sqrt(5e-324) - 1) < 1e-15 and got[9] == 2.0
    assert abs(got[10] / np.sqrt(1e-310) - 1) < 1e-15
