"""Shared builders for the parity tests: one synthetic case -> inputs for the oracle AND for the engine."""
import numpy as np

from troute_b200 import synth
from troute_b200.network import TRT_KIND_LEVELPOOL, TRT_KIND_MC


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def assert_bit_equal(a, b, what=""):
    a = np.ascontiguousarray(a, dtype=np.float32)
    b = np.ascontiguousarray(b, dtype=np.float32)
    assert a.shape == b.shape, (what, a.shape, b.shape)
    same = (bits(a) == bits(b)) | (np.isnan(a) & np.isnan(b))
    if not same.all():
        idx = np.argwhere(~same)
        i = tuple(idx[0])
        raise AssertionError(f"{what}: {idx.shape[0]} of {a.size} values differ; first at {i}: {a[i]!r} vs {b[i]!r}")


def make_case(down, nsteps=48, qts=12, seed=16, n_lp=0, warm=False):
    """Inputs of one routing call for a network given by down[]: every segment is its own reach unless
    it is the single upstream of its downstream neighbour (then it is chained into the same reach)."""
    n = down.shape[0]
    params = synth.channel_params(down, seed=seed)
    qlat = synth.lateral_inflow(n, nsteps, qts, seed=seed)
    rng = np.random.default_rng(seed + 7)
    if warm:
        q0 = np.stack([rng.uniform(0.0, 2.0, n), rng.uniform(0.0, 2.0, n), rng.uniform(0.0, 1.0, n)], axis=1).astype(np.float32)
    else:
        q0 = np.zeros((n, 3), dtype=np.float32)
    up_ptr, up_rows = synth.upstream_csr(down)
    kind = np.zeros(n, dtype=np.uint8)
    lp_rows = np.zeros(0, dtype=np.int64)
    wbody = np.zeros((0, 11))
    if n_lp:
        # reservoirs replace in-line segments that have at least one upstream neighbour
        cand = np.nonzero(np.diff(up_ptr) > 0)[0]
        lp_rows = np.sort(rng.choice(cand, size=min(n_lp, cand.size), replace=False)).astype(np.int64)
        kind[lp_rows] = TRT_KIND_LEVELPOOL
        wbody = synth.levelpool_params(lp_rows.size, seed=seed)
    return dict(n=n, down=down, params=params, cols=synth.PARAM_COLS, qlat=qlat, q0=q0, up_ptr=up_ptr,
                up_rows=up_rows, kind=kind, lp_rows=lp_rows, wbody=wbody, nsteps=nsteps, qts=qts)


def oracle_route(o, case, assume_short_ts, pow_mode=None, reaches=None, jobs=None, nthreads=0, gages=None):
    """Run the oracle in the reference's loop order.  By default each segment is a single-segment reach
    listed in level order (a valid upstream-first order); `reaches` overrides that."""
    n = case["n"]
    if pow_mode is None:
        pow_mode = o.POW_DET
    scols = np.asarray(o.column_mapper(case["cols"]), dtype=np.int32)
    if reaches is None:
        level = synth.levels_from_down(case["down"])
        order = np.argsort(level, kind="stable").astype(np.int64)
        reach_ptr = np.arange(n + 1, dtype=np.int64)
        reach_rows = order
        counts = np.diff(case["up_ptr"])[order]
        reach_up_ptr = np.zeros(n + 1, dtype=np.int64)
        np.cumsum(counts, out=reach_up_ptr[1:])
        starts = case["up_ptr"][order]
        reach_up_rows = np.concatenate([case["up_rows"][s:s + c] for s, c in zip(starts, counts)]) if n < 200000 else _gather(case, order, counts, starts)
        reach_type = case["kind"][order].astype(np.int32)
        lp_index = {int(r): i for i, r in enumerate(case["lp_rows"])}
        reach_wbody = np.asarray([lp_index.get(int(r), -1) for r in order], dtype=np.int32) if len(lp_index) else np.full(n, -1, np.int32)
    else:
        reach_ptr, reach_rows, reach_type, reach_up_ptr, reach_up_rows, reach_wbody = reaches
    fvd, up, extras = o.route_network_flat(
        case["nsteps"], 300.0, case["qts"], n, reach_ptr, reach_rows, reach_type, reach_up_ptr, reach_up_rows,
        case["params"], scols, case["q0"], case["qlat"], assume_short_ts=assume_short_ts, reach_wbody=reach_wbody,
        wbody_cols=case["wbody"], pow_mode=pow_mode, jobs=jobs, nthreads=nthreads, want_hist=True, gages=gages)
    return fvd[:, 1:, :].reshape(n, -1), up[:, 1:], extras


def _gather(case, order, counts, starts):
    total = int(counts.sum())
    out = np.empty(total, dtype=np.int64)
    # vectorised ragged gather
    offs = np.zeros(order.shape[0] + 1, dtype=np.int64)
    np.cumsum(counts, out=offs[1:])
    idx = np.arange(total, dtype=np.int64) - np.repeat(offs[:-1], counts) + np.repeat(starts, counts)
    out[:] = case["up_rows"][idx]
    return out


def engine_route(case, assume_short_ts, mode=None, device=0, want_upstream=True, options=None):
    from troute_b200.network import RoutingNetwork
    net = RoutingNetwork(case["up_ptr"], case["up_rows"], case["kind"], case["params"], case["cols"], device=device)
    try:
        if case["lp_rows"].size:
            net.set_levelpools(case["lp_rows"], case["wbody"])
        if mode is not None:
            net.set_option("mode", mode)
        for k, v in (options or {}).items():
            net.set_option(k, v)
        fvd, up = net.route(case["nsteps"], case["qts"], case["qlat"], case["q0"], assume_short_ts=assume_short_ts,
                            want_upstream=want_upstream)
        stats = net.last_run_stats()
        stats["levels"] = net.num_levels
    finally:
        net.close()
    return fvd, up, stats


def _mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrap-around arithmetic), the hash of csrc/routing_kernels.cu::mix64"""
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return x ^ (x >> np.uint64(31))


def result_hash(fvd, ids=None):
    """trt_result_hash in numpy: sum over rows of mix(sum_j mix(j << 32 | bits[row, j]) ^ mix(id)), modulo 2^64."""
    with np.errstate(over="ignore"):
        b = np.ascontiguousarray(fvd, dtype=np.float32).view(np.uint32).astype(np.uint64)
        j = (np.arange(b.shape[1], dtype=np.uint64) << np.uint64(32))[None, :]
        h = _mix64(j | b).sum(axis=1, dtype=np.uint64)
        ids = np.arange(b.shape[0], dtype=np.uint64) if ids is None else np.asarray(ids).astype(np.uint64)
        return int(_mix64(h ^ _mix64(ids)).sum(dtype=np.uint64))
