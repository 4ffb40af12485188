"""The SPMD structure of the diffusive time loop, checked without a GPU: tests/native/diffusive_cta_main.cpp runs the
product's dw_time_loop as an emulated CTA -- N host threads that meet at a barrier at every __syncthreads() of the device
code -- under ThreadSanitizer.  A phase that reads what another thread of the same phase writes (a missing barrier, a lane
touching another reach's slot) is a data race TSan reports; and the result must still equal the oracle bit for bit, whatever
the interleaving.  (What this cannot see: device code generation and memory-model effects below the barrier level.)"""
import os
import subprocess

import numpy as np
import pytest

import helpers_diffusive as HD

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "diffusive_cta_main.cpp")
EXE = os.path.join(HERE, "native", "diffusive_cta_tsan")


def build_exe():
    deps = [SRC] + HD.DEPS[1:]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        flags = open("/proc/cpuinfo").read() if os.path.exists("/proc/cpuinfo") else ""
        cmd = ["g++", "-O1", "-g", "-ffp-contract=off", "-fno-fast-math", "-std=c++17", "-DDW_EMULATE_CTA", "-fsanitize=thread",
               "-pthread", "-o", EXE, SRC]
        if " fma" in flags:
            cmd.insert(1, "-mfma")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            pytest.skip("ThreadSanitizer build not available here: " + r.stderr[-300:])
    return EXE


def write_inputs(path, d):
    from oracle import diffusive as od
    with open(path, "wb") as f:
        for name, kind in od.ARGS:
            v = d[name]
            a = np.asarray([int(v)], dtype=np.int32) if kind == "i" else np.asfortranarray(v, dtype=np.int32 if kind == "I" else np.float64)
            flat = a.ravel(order="F")
            f.write(np.asarray([flat.size], dtype=np.int64).tobytes())
            f.write(flat.tobytes())


def run_emulated(d, nthreads, tmp_path):
    exe = build_exe()
    inp, out = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
    write_inputs(inp, d)
    env = dict(os.environ, TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=66")
    r = subprocess.run(["setarch", "-R", exe, str(nthreads), inp, out], capture_output=True, text=True, env=env, timeout=900)
    if "unexpected memory mapping" in r.stderr or "FATAL: ThreadSanitizer" in r.stderr:
        pytest.skip("ThreadSanitizer cannot run in this sandbox: " + r.stderr[-200:])
    shape = (int(d["ntss_ev_g"]), int(d["mxncomp_g"]), int(d["nrch_g"]))
    n = shape[0] * shape[1] * shape[2]
    raw = np.fromfile(out, dtype=np.float64) if os.path.exists(out) else np.zeros(0)
    return r, [np.ascontiguousarray(raw[k * n:(k + 1) * n].reshape(shape, order="F")) for k in range(3)] if raw.size == 3 * n else None


@pytest.mark.parametrize("case,nthreads", [("small", 64), ("branched", 256), ("tailwater-depth", 32)])
def test_emulated_cta_is_race_free_and_equals_the_oracle(tmp_path, case, nthreads):
    from oracle import diffusive as od
    from troute_b200 import synth_diffusive as sd
    od.build()
    kw = dict(HD.CASES[case]); kw["nsteps"] = min(kw.get("nsteps", 72), 48)
    d = sd.diffusive_domain(**kw)
    ref = od.compute_diffusive(d, od.POW_DET)
    r, got = run_emulated(d, nthreads, tmp_path)
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0, (r.returncode, r.stderr[-500:])
    for name, a, b in zip(("q_ev_g", "elv_ev_g", "depth_ev_g"), ref, got):
        HD.assert_bits64(b, a, f"{case} / {nthreads} threads: {name}")


def test_emulated_cta_with_surveyed_sections(tmp_path):
    from oracle import diffusive as od
    from troute_b200 import synth_diffusive as sd
    od.build()
    d = sd.with_natural_sections(sd.diffusive_domain(n_branch=2, n_mainstem=6, nsteps=36))
    ref = od.compute_diffusive(d, od.POW_DET)
    r, got = run_emulated(d, 128, tmp_path)
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[:3000]
    assert r.returncode == 0, (r.returncode, r.stderr[-500:])
    for a, b in zip(ref, got):
        HD.assert_bits64(b, a, "surveyed sections, 128 threads")
