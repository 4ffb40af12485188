"""Test helpers of the diffusive-wave path: builds / loads the host replica of the product's solver source
(tests/native/diffusive_replica.cpp, TEST INFRASTRUCTURE) and compares outputs."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "native", "diffusive_replica.cpp")
LIB = os.path.join(HERE, "native", "libdiffusive_replica.so")
DEPS = [SRC, os.path.join(ROOT, "t-route_b200", "csrc", "diffusive_device.cuh"),
        os.path.join(ROOT, "t-route_b200", "csrc", "diffusive_setup.h"), os.path.join(ROOT, "include", "trt_detmath64.h")]
_lib = None


def replica_lib():
    """g++ -O2 -ffp-contract=off: the flags of the oracle (no FMA contraction, no fast-math)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or any(os.path.getmtime(d) > os.path.getmtime(LIB) for d in DEPS):
            flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
            cmd = ["g++", "-O2", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-shared", "-fPIC", "-o", LIB, SRC]
            if " fma" in flags:
                cmd.insert(1, "-mfma")
            subprocess.run(cmd, check=True)
        _lib = C.CDLL(LIB)
        _lib.trt_replica_diffnw.restype = C.c_int
    return _lib


def replica_compute_diffusive(diff_inputs):
    from oracle import diffusive as od
    args, keep, shape = od.marshal(diff_inputs)
    outs = [np.zeros(shape, dtype=np.float64, order="F") for _ in range(3)]
    rc = replica_lib().trt_replica_diffnw(*args, *[o.ctypes.data_as(C.c_void_p) for o in outs])
    if rc != 0:
        raise RuntimeError(f"trt_replica_diffnw failed with status {rc}")
    return tuple(np.ascontiguousarray(o) for o in outs)


def assert_bits64(a, b, what):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    bad = a.view(np.int64) != b.view(np.int64)
    if bad.any():
        idx = tuple(np.argwhere(bad)[0])
        raise AssertionError(f"{what}: {int(bad.sum())} of {bad.size} values differ; first at {idx}: {a[idx]!r} vs {b[idx]!r}")


def mainstem_nodes(d):
    """boolean mask (mxncomp, nrch) of the nodes the solver computes"""
    m = np.zeros((d["mxncomp_g"], d["nrch_g"]), dtype=bool)
    for j in d["mainstem"]:
        m[: d["frnw_g"][j, 0], j] = True
    return m


CASES = {
    "small": dict(),
    "branched": dict(n_branch=3, n_mainstem=8, nsteps=144),
    "tailwater-depth": dict(dsbc_option=1),
    "flashy": dict(seed=3, pulse=6.0, nsteps=200, slope=2e-4),
    "long": dict(n_mainstem=24, nodes=(5, 12), nsteps=288, n_branch=4, seed=7),
}
