"""Host-side cost of the Python mirror `compute_network_structured` per call, without a GPU: the cache key of the device
network (`_fingerprint`), the flattening of the reaches (first call only) and the assembly of the result tuple, on a
reference-style description (list of (reach, type), dict of upstream connections) of an NHD-like forest.

    python tools/mirror_overhead.py [n_segments] [nsteps]

What a T-Route caller pays around the C-ABI call (VERDICT r01 "the Python mirrors' per-call cost")."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "t-route_b200")):
    sys.path.insert(0, p)


def reference_style(n, seed=16):
    from troute_b200 import synth
    down = synth.conus_like(n_total=n, n_basins=max(4, n // 185), seed=seed, style="nhd")
    n = down.shape[0]
    ids = (np.arange(n, dtype=np.int64) * 3 + 11)
    up_ptr, up_rows = synth.upstream_csr(down)
    indeg = np.diff(up_ptr)
    level = synth.levels_from_down(down)
    head = indeg != 1                                   # a segment with exactly one upstream neighbour continues its reach
    order = np.lexsort((np.arange(n), level))
    reaches = []
    dl = down.tolist(); hl = head.tolist(); idl = ids.tolist()
    for h in order[head[order]].tolist():
        r = [idl[h]]
        d = dl[h]
        while d >= 0 and not hl[d]:
            r.append(idl[d]); d = dl[d]
        reaches.append((r, 0))
    upl = up_rows.tolist(); pl = up_ptr.tolist()
    conn = {idl[i]: [idl[u] for u in upl[pl[i]:pl[i + 1]]] for i in range(n)}
    params = synth.channel_params(down, seed=seed)
    return n, ids, reaches, conn, params, list(synth.PARAM_COLS)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_729_077
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 288
    from troute_b200.routing.fast_reach import mc_reach
    t0 = time.time()
    n, ids, reaches, conn, params, cols = reference_style(n)
    print(f"{n} segments, {len(reaches)} reaches; building the reference-style description took {time.time() - t0:.1f} s (the caller's cost)")
    t0 = time.time()
    mc_reach._fingerprint(reaches, conn, ids, cols, params, 0)
    print(f"full fingerprint (first call on a topology, or VERIFY_TOPOLOGY_EVERY_CALL): {time.time() - t0:.3f} s")
    mc_reach._NET_CACHE[mc_reach._network_key(reaches, conn, ids, cols, params, 0)] = dict(net=None)   # as after a first call
    for rep in range(2):
        t0 = time.time()
        mc_reach._network_key(reaches, conn, ids, cols, params, 0)
        print(f"cache key of a repeated call (same topology objects): {time.time() - t0:.3f} s")
    mc_reach._NET_CACHE.clear(); mc_reach._QUICK.clear()
    t0 = time.time()
    flat = mc_reach.flatten_network(reaches, conn, ids)
    print(f"flatten_network (first call on a network): {time.time() - t0:.3f} s")
    fvd = np.zeros((n, 3 * T), dtype=np.float32)
    up = np.zeros((n, T), dtype=np.float32) if hasattr(mc_reach, "_assemble") else None
    t0 = time.time()
    mask = np.ones(n, dtype=bool)
    a = ids.astype(np.intp)[mask]; b = fvd[mask]
    print(f"result tuple with boolean-mask copies (round-1 form): {time.time() - t0:.3f} s")
    if hasattr(mc_reach, "_take_rows"):
        t0 = time.time()
        a = mc_reach._take_rows(ids.astype(np.intp), mask); b = mc_reach._take_rows(fvd, mask)
        print(f"result tuple, no copy when every row is returned: {time.time() - t0:.3f} s")


if __name__ == "__main__" and not (len(sys.argv) > 3 and sys.argv[3] in ("frames", "gpu")):
    main()


def frames_level(n=1_000_000, T=288, calls=3):
    """compute_nhd_routing_v02 (DataFrames in, tuples out) around a stub compute function: what the frame handling and the
    reach bookkeeping of the mirror cost per call, first call and repeated calls with `subnetwork_list` handed back."""
    import pandas as pd
    from troute_b200.routing import compute
    from troute_b200 import synth
    n, ids, reaches, conn, params, cols = reference_style(n)
    idl = ids.tolist()
    tw = [r[0][-1] for r in reaches if True][-1]
    reaches_bytw = {tw: [r for r, _ in reaches]}
    param_df = pd.DataFrame(params, index=ids, columns=cols)
    if "alt" not in param_df:
        param_df["alt"] = 0.0
    q0 = pd.DataFrame(np.zeros((n, 3), dtype=np.float32), index=ids, columns=["qu0", "qd0", "h0"])
    qlats = pd.DataFrame(np.zeros((n, T // 12), dtype=np.float32), index=ids)
    seen = {}

    def stub(*a, **k):
        seen["rows"] = len(a[5])
        return None
    compute._compute_func_map["stub"] = stub
    sub = [None, None, None]
    import datetime
    for c in range(calls):
        t0 = time.time()
        _, sub = compute.compute_nhd_routing_v02(
            None, conn, None, reaches_bytw, "stub", "serial", 1, None, datetime.datetime(2021, 8, 23, 13), 300.0, T, 12,
            {tw: conn}, param_df, q0, qlats, pd.DataFrame(), pd.DataFrame(), pd.DataFrame(), pd.DataFrame(), pd.DataFrame(),
            pd.DataFrame(), pd.DataFrame(), pd.DataFrame(), pd.DataFrame(), pd.DataFrame(), pd.DataFrame(), {}, False, False,
            pd.DataFrame(), {}, pd.DataFrame(), False, sub)
        print(f"compute_nhd_routing_v02 around a stub kernel, call {c + 1}: {time.time() - t0:.3f} s ({seen['rows']} rows)")


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "frames":
    frames_level(int(sys.argv[1]), int(sys.argv[2]))


def gpu_level(n=2_729_077, T=288):
    """the mirror end to end on a GPU: compute_network_structured on reference-style inputs, first call (flatten, device
    network, trip-count collection, re-ordered rebuild) and repeated calls, with and without the pinned result buffer"""
    from troute_b200 import synth
    from troute_b200.routing.fast_reach import mc_reach
    n, ids, reaches, conn, params, cols = reference_style(n)
    qlat = synth.lateral_inflow(n, T, 12, seed=16)
    q0 = np.zeros((n, 3), dtype=np.float32)
    e_f = np.zeros(0, np.float32); e_i = np.zeros(0, np.int32); e_f2 = np.zeros((0, 0), np.float32)

    def call():
        return mc_reach.compute_network_structured(
            T, 300.0, 12, reaches, conn, ids, np.array(cols, dtype=object), params, q0, qlat, [], np.zeros((0, 11)), {},
            np.zeros((0, 1), np.int32), False, "2021-08-23_13:00:00", e_f2, e_i, e_i, e_i, e_f, e_f, 0.0,
            e_f2, e_i, e_f, e_f, e_f, e_f, e_f, e_f2, e_i, e_f, e_f, e_f, e_f, e_f, e_f2, e_i, e_i, [], e_i, e_i, e_f, e_i, e_i,
            e_i, e_i, e_f, e_i, e_f, e_i, e_i, e_f2, {}, False, False)
    for reuse in (False, True):
        mc_reach.RESULT_POOL_BYTES = (32 << 30) if reuse else 0
        for c in range(4):
            t0 = time.time()
            out = call()
            dt = time.time() - t0
            print(f"compute_network_structured, {n} segments x {T} steps, pinned result pool {'on' if reuse else 'off'}, call {c + 1}: "
                  f"{dt:.3f} s  ({n * T / dt:.3g} segment-timesteps/s)  checksum {float(out[1][::1000, -3].sum()):.6g}", flush=True)
            del out
    mc_reach.clear_network_cache()


if __name__ == "__main__" and len(sys.argv) > 3 and sys.argv[3] == "gpu":
    gpu_level(int(sys.argv[1]), int(sys.argv[2]))
