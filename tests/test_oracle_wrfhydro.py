"""Second pin of the oracle: oracle/mc_wrfhydro.c restates the WRF-Hydro ORIGINAL of the kernel,
/root/reference/src/kernel/muskingum/MUSKINGCUNGE.f90, independently of oracle/mc_kernel.inc (the restatement of
MCsingleSegStime_f2py_NOLOOP.f90 the GPU path is checked against).  The two Fortran sources perform the same float32
operations in the same order wherever their logic coincides, so the two restatements must agree BIT FOR BIT there -- in-bank
evaluations, compound-channel evaluations with a floodplain and a positive celerity, first attempt of the retry ladder -- and
every row on which they do not agree must carry one of the flags that name a difference between the two Fortran FILES
(oracle.WRF_*; the header of mc_wrfhydro.c lists them with line numbers).  The reference itself holds one known-answer
vector for this kernel (in-bank, no retries, tests/test_oracle_kat.py); this widens the pinned region to the compound
channel branch (:248-258) and to thousands of inputs."""
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NAMES = ["dt", "qup", "quc", "qdp", "ql", "dx", "bw", "tw", "twcc", "n", "ncc", "cs", "s0", "velp", "depthp"]


def bankfull_depth(r):
    bw, tw, cs = r[:, 6], r[:, 7], r[:, 11]
    z = np.where(cs == 0, 1.0, 1.0 / np.where(cs == 0, 1.0, cs))
    return np.where(bw > tw, bw / 1e-5, np.where(bw == tw, bw / (2 * z), (tw - bw) / (2 * z)))


def edge_rows():
    """Compound channels in flood, channels without a floodplain, vertical banks, tiny and huge flows, losing reaches."""
    base = np.array([300, 1, 1, 1, 0.1, 1000, 5, 8, 24, 0.06, 0.12, 0.6, 0.01, 0, 0.5], dtype=np.float32)
    rng = np.random.default_rng(3)
    rows = []
    for _ in range(4000):
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
    return np.asarray(rows, dtype=np.float32)


def compare(oracle, in15, pow_mode):
    a, ia = oracle.mc_segment_batch(in15, pow_mode=pow_mode)
    b, flags, ib = oracle.wrfhydro_mc_batch(in15, pow_mode)
    # the original leaves velc undefined on its no-flow branch (:326-329); compare q and depth there
    noflow = ~((in15[:, 4] > 0) | (in15[:, 1] > 0) | (in15[:, 3] > 0))
    same = (a[:, [0, 2]].view(np.int32) == b[:, [0, 2]].view(np.int32)).all(axis=1)
    same &= noflow | (a[:, 1].view(np.int32) == b[:, 1].view(np.int32)) | (np.isnan(a[:, 1]) & np.isnan(b[:, 1]))
    return a, b, flags, same, ia, ib


@pytest.mark.parametrize("pow_mode", ["det", "libm"])
def test_reference_suite_in_bank_and_compound_rows_agree_bit_for_bit(oracle, pow_mode):
    pm = oracle.POW_DET if pow_mode == "det" else oracle.POW_LIBM
    in15 = np.load(os.path.join(GOLD, "mc_suite_seed16.npy"))       # test_suite_parameters.py, seed 16, n = 5000
    a, b, flags, same, ia, ib = compare(oracle, in15, pm)
    clean = flags == 0
    assert same[clean].all(), f"{int((~same & clean).sum())} unflagged rows differ"
    assert clean.sum() >= 4990                                       # the two sources coincide on (almost) the whole suite
    assert np.array_equal(ia[clean], ib[clean])                      # same secant trip counts
    # the pinned region includes the compound-channel branch: rows that END above bankfull depth with a floodplain
    over = clean & (a[:, 2] > bankfull_depth(in15)) & (in15[:, 8] > 0) & (in15[:, 10] > 0)
    assert over.sum() >= 100, int(over.sum())
    # and every disagreement is explained by a difference between the two Fortran files
    assert (flags[~same] != 0).all()


def test_edge_rows_every_disagreement_is_a_known_difference_of_the_sources(oracle):
    in15 = edge_rows()
    a, b, flags, same, ia, ib = compare(oracle, in15, oracle.POW_DET)
    clean = flags == 0
    assert same[clean].all(), f"{int((~same & clean).sum())} unflagged rows differ"
    assert clean.sum() >= 2000
    counts = {name: int(((flags & bit) != 0).sum()) for name, bit in (
        ("retry ladder (Qj_0 reset)", oracle.WRF_RETRY), ("no floodplain above bankfull", oracle.WRF_NO_FLOODPLAIN),
        ("compound with zero celerity", oracle.WRF_ZERO_CELERITY), ("zero wetted perimeter", oracle.WRF_ZERO_PERIMETER),
        ("only quc positive", oracle.WRF_ONLY_QUC))}
    print("rows by difference of the two sources:", counts, "| flagged rows that still agree:", int((same & ~clean).sum()))
    assert counts["no floodplain above bankfull"] > 50              # the NWM 3.0 exception is exercised ...
    # ... and it is a real difference: some of those rows give different answers
    assert (~same[(flags & oracle.WRF_NO_FLOODPLAIN) != 0]).any()


def test_demo_vector_through_the_original(oracle):
    """The reference's demo inputs (mc_sseg_stime_NOLOOP_demo.py:173-209) are in-bank: the WRF-Hydro original gives the same
    bits as the NOLOOP restatement and the published answer (depthc exact, qdc / velc within 1 ulp), i.e. the one KAT the
    reference holds pins both restatements."""
    import json
    kat = json.load(open(os.path.join(GOLD, "mc_demo_kat.json")))
    c, s = kat["channel"], kat["single"]
    d = dict(c); d.update({k: s[k] for k in ("qup", "quc", "qdp", "depthp", "velp")})
    row = np.asarray([[d[k] for k in NAMES]], dtype=np.float32)
    a, b, flags, same, _, _ = compare(oracle, row, oracle.POW_LIBM)
    assert flags[0] == 0 and same[0]
    e = s["expected"]
    assert b[0, 2] == np.float32(e["depthc"])
    for got, want in ((b[0, 0], e["qdc"]), (b[0, 1], e["velc"])):
        assert abs(int(np.float32(got).view(np.int32)) - int(np.float32(want).view(np.int32))) <= 1
