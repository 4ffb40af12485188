"""Warm restarts: what the BMI / nwm_route drivers do between loops (troute_model.py:138-345, __main__.py:255-330): route
`until` seconds, take the last column of the results as the next initial condition (AbstractNetwork.new_q0: qu0 = qd0 =
q_last, h0 = depth_last; reservoirs keep their last outflow and water elevation, update_waterbody_water_elevation) and call
again with the next forcing.  Two half-length calls must equal one full-length call, bit for bit -- on the oracle (CPU) and
on the device, where the second call also reuses the cached device network (the `subnetwork_list` precedent,
troute_model.py:75)."""
import numpy as np
import pytest

import helpers as H
import test_gpu_api as A


def split_run(fn, c, halves=2):
    """route c["nsteps"] steps in `halves` calls of equal length, carrying the state like the reference's drivers"""
    n_half = c["nsteps"] // halves
    q_cols = n_half // c["qts"]
    ids = c["ids"]
    lake_rows = np.searchsorted(ids, np.asarray(c["lake_numbers"], dtype=np.int64))
    q0 = c["q0"].copy()
    wbody = np.array(c["wbody"], dtype=np.float64, copy=True)
    pieces = []
    for h in range(halves):
        part = dict(c)
        part["nsteps"] = n_half
        part["qlat"] = c["qlat"][:, h * q_cols:(h + 1) * q_cols]
        part["q0"] = q0
        part["wbody"] = wbody
        out = A._call(fn, part)
        order = np.argsort(out[0])
        assert np.array_equal(out[0][order], ids)
        fvd = out[1][order]
        pieces.append(fvd)
        q0 = np.stack([fvd[:, -3], fvd[:, -3], fvd[:, -1]], axis=1).astype(np.float32)             # new_q0
        if len(lake_rows):
            wbody = wbody.copy()
            wbody[:, 9] = fvd[lake_rows, -3]                                                      # qd0 <- last outflow
            wbody[:, 10] = fvd[lake_rows, -1]                                                     # h0  <- last water elevation
    return np.concatenate(pieces, axis=1)


def full_run(fn, c):
    out = A._call(fn, c)
    return out[1][np.argsort(out[0])]


def test_split_run_equals_full_run_on_the_oracle(oracle):
    c = A._reference_style_case(n=3000, seed=13, n_lp=8, nsteps=48)
    H.assert_bit_equal(split_run(oracle.compute_network_structured, c), full_run(oracle.compute_network_structured, c),
                       "oracle: two warm-started halves vs one call")


@pytest.mark.gpu
@pytest.mark.parametrize("halves", [2, 4])
def test_split_run_equals_full_run_on_the_device(oracle, halves):
    from troute_b200.routing.fast_reach import mc_reach
    c = A._reference_style_case(n=3000, seed=13, n_lp=8, nsteps=48)
    ref = full_run(oracle.compute_network_structured, c)
    try:
        got = split_run(mc_reach.compute_network_structured, c, halves)
        assert len(mc_reach._NET_CACHE) == 1                     # every call after the first reused the device network
    finally:
        mc_reach.clear_network_cache()
    H.assert_bit_equal(got, ref, f"device: {halves} warm-started calls vs the oracle's single call")
