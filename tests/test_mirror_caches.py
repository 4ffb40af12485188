"""Host-side caches of the Python mirrors (no GPU): the per-call cost of `compute_network_structured` /
`compute_nhd_routing_v02` at CONUS scale is the cache key and the frame handling, not the routing (tools/mirror_overhead.py),
so both keep what depends on the topology alone -- and must never hand a stale topology to the device."""
import numpy as np
import pandas as pd
import pytest

from troute_b200.routing import compute
from troute_b200.routing.fast_reach import mc_reach


def _topology(n=400, seed=0):
    rng = np.random.default_rng(seed)
    ids = np.sort(rng.choice(10 ** 6, size=n, replace=False)).astype(np.int64)
    reaches, conn = [], {}
    i = 0
    while i < n:
        k = int(rng.integers(1, 4))
        seg = [int(x) for x in ids[i:i + k]]
        reaches.append((seg, 0))
        conn[seg[0]] = [int(x) for x in ids[max(0, i - 2):i][:int(rng.integers(0, 3))]]
        for a, b in zip(seg[1:], seg[:-1]):
            conn[a] = [b]
        i += k
    vals = rng.uniform(0.1, 10.0, (n, 9)).astype(np.float32)
    return ids, reaches, conn, vals, ["dt", "bw", "tw", "twcc", "dx", "n", "ncc", "cs", "s0"]


def test_reach_type_list_equals_the_reference_expression():
    """compute.py:41-47: `1 if (set(reaches) & wbodies_segs) else 0` per reach"""
    rng = np.random.default_rng(1)
    reach_list = [[int(x) for x in rng.integers(0, 500, int(rng.integers(1, 6)))] for _ in range(300)]
    for wb in (set(), {3, 17, 255}, set(range(0, 500, 7))):
        want = [(r, 1 if (set(r) & wb) else 0) for r in reach_list]
        assert compute._build_reach_type_list(reach_list, wb) == want
    big_then_empty = set(range(200000)).symmetric_difference(range(200000))    # empty, with the table of a large set
    assert compute._build_reach_type_list(reach_list, big_then_empty) == [(r, 0) for r in reach_list]


def test_network_key_recognises_the_same_objects_and_nothing_else():
    ids, reaches, conn, vals, cols = _topology()
    mc_reach.clear_network_cache()
    full = mc_reach._fingerprint(reaches, conn, ids, cols, vals, 0)
    k1 = mc_reach._network_key(reaches, conn, ids, cols, vals, 0)
    assert k1 == full
    mc_reach._NET_CACHE[k1] = dict(net=_Closable())              # as after a first call
    calls = []
    orig = mc_reach._fingerprint
    mc_reach._fingerprint = lambda *a: calls.append(1) or orig(*a)
    try:
        assert mc_reach._network_key(reaches, conn, ids, cols, vals, 0) == k1 and not calls      # quick key: no full walk
        # same containers, different parameters / ids / device: a different network
        v2 = vals.copy(); v2[7, 3] += 1.0
        assert mc_reach._network_key(reaches, conn, ids, cols, v2, 0) != k1 and len(calls) == 1
        assert mc_reach._network_key(reaches, conn, ids, cols, vals, 1) != k1
        # a copy of the topology with one confluence rewired: new objects, full fingerprint, different key
        r2 = [(list(r), t) for r, t in reaches]
        c2 = {k: list(v) for k, v in conn.items()}
        head = next(r[0] for r, _ in r2 if len(c2[r[0]]) == 0)
        c2[head] = [int(ids[0])]
        assert mc_reach._network_key(r2, c2, ids, cols, vals, 0) != k1
        # an identical copy maps to the same device network
        r3 = [(list(r), t) for r, t in reaches]
        assert mc_reach._network_key(r3, dict(conn), ids, cols, vals, 0) == k1
        # callers that edit a topology in place ask for the full check on every call
        mc_reach.VERIFY_TOPOLOGY_EVERY_CALL = True
        n = len(calls)
        reaches[len(reaches) // 2 + 1][0].reverse()
        assert mc_reach._network_key(reaches, conn, ids, cols, vals, 0) != k1 or len(reaches[len(reaches) // 2 + 1][0]) == 1
        assert len(calls) == n + 1
    finally:
        mc_reach._fingerprint = orig
        mc_reach.VERIFY_TOPOLOGY_EVERY_CALL = False
        mc_reach.clear_network_cache()


class _Closable:
    def close(self):
        pass


def test_take_rows_copies_only_when_rows_are_dropped():
    a = np.arange(12, dtype=np.float32).reshape(4, 3)
    m = np.ones(4, dtype=bool)
    assert mc_reach._take_rows(a, m) is a
    m[2] = False
    out = mc_reach._take_rows(a, m)
    assert out.shape == (3, 3) and np.array_equal(out, a[[0, 1, 3]])


def test_frames_cache_is_keyed_on_the_tables(monkeypatch):
    """compute_nhd_routing_v02 keeps reach lists / sub-frames per topology; a changed parameter table or another
    reaches_bytw object rebuilds them, and `subnetwork_list` comes back untouched (the reference's serial branch)."""
    ids, reaches, conn, vals, cols = _topology(300, seed=3)
    reaches_bytw = {int(ids[-1]): [r for r, _ in reaches]}
    param_df = pd.DataFrame(vals[:, 1:], index=ids, columns=cols[1:])
    param_df["alt"] = 0.0
    q0 = pd.DataFrame(np.zeros((ids.size, 3), np.float32), index=ids, columns=["qu0", "qd0", "h0"])
    qlats = pd.DataFrame(np.ones((ids.size, 2), np.float32), index=ids)
    seen = []

    def stub(*a, **k):
        seen.append((a[3], a[7].copy()))
        return None
    monkeypatch.setitem(compute._compute_func_map, "stub", stub)
    compute._TOPO_CACHE.clear()
    e = pd.DataFrame()
    import datetime

    def call(pdf, rb):
        sl = [None, None, None]
        _, out = compute.compute_nhd_routing_v02(None, conn, None, rb, "stub", "serial", 1, None, datetime.datetime(2021, 1, 1), 300.0, 24,
                                                 12, {1: conn}, pdf, q0, qlats, e, e, e, e, e, e, e, e, e, e, e, {}, False, False, e, {}, e,
                                                 False, sl)
        assert out is sl and sl == [None, None, None]
    call(param_df.copy(), reaches_bytw)
    call(param_df.copy(), reaches_bytw)
    assert seen[1][0] is seen[0][0] and len(compute._TOPO_CACHE) == 1          # the same reach list object: the device cache key hits
    p2 = param_df.copy(); p2.iloc[5, 2] *= 2.0
    call(p2, reaches_bytw)
    assert seen[2][0] is not seen[0][0] and not np.array_equal(seen[2][1], seen[0][1]) and len(compute._TOPO_CACHE) == 2
    call(param_df.copy(), dict(reaches_bytw))
    assert seen[3][0] is not seen[0][0] and np.array_equal(seen[3][1], seen[0][1])
    compute._TOPO_CACHE.clear()


class _FakeBlock:
    """host-memory stand-in for network._PinnedBlock (trt_host_alloc needs a CUDA device)"""
    live = 0

    def __init__(self, nbytes):
        import ctypes
        self._mem = (ctypes.c_uint8 * nbytes)()
        self.ptr = ctypes.c_void_p(ctypes.addressof(self._mem))
        self.nbytes = nbytes
        _FakeBlock.live += 1

    def free(self):
        if self.ptr:
            self.ptr = None
            _FakeBlock.live -= 1


def test_pinned_pool_reuses_a_block_only_after_its_array_is_gone(monkeypatch):
    """network.PinnedPool: the logic behind mc_reach.RESULT_POOL_BYTES, with host memory standing in for pinned blocks"""
    import gc
    from troute_b200 import network
    monkeypatch.setattr(network.PinnedPool, "_block_type", _FakeBlock)
    _FakeBlock.live = 0
    pool = network.PinnedPool()
    a = pool.take((100, 30), np.float32, 30000)
    a[:] = 1.0
    b = pool.take((100, 30), np.float32, 30000)               # `a` is held: a second block
    assert b is not None and not np.shares_memory(a, b) and pool.pinned_bytes == 24000
    assert pool.take((100, 30), np.float32, 30000) is None     # a third would exceed the limit: the caller goes pageable
    view = a[10:20]
    addr = a.ctypes.data
    del a
    gc.collect()
    assert pool.take((100, 30), np.float32, 24000) is None     # a view still holds the block
    del view
    gc.collect()
    c = pool.take((100, 30), np.float32, 24000)
    assert c is not None and c.ctypes.data == addr and pool.pinned_bytes == 24000      # the released block, nothing new pinned
    # another shape: the idle block of the old shape is given up to stay under the limit
    del c
    gc.collect()
    d = pool.take((50, 100), np.float32, 36000)
    assert d is not None and pool.pinned_bytes == 12000 + 20000 and _FakeBlock.live == 2
    del b, d
    gc.collect()
    pool.clear()
    assert pool.pinned_bytes == 0 and _FakeBlock.live == 0
