"""Host logic of the drop-in boundary (CPU only): reach lists -> CSR, argument checks, caching key."""
import numpy as np
import pytest

from troute_b200 import synth
from troute_b200.routing.fast_reach import mc_reach
from troute_b200.network import TRT_KIND_BOUNDARY, TRT_KIND_LEVELPOOL, TRT_KIND_MC


def test_binary_find_matches_reference_semantics():
    arr = np.array([2, 5, 9, 40], dtype=np.int64)
    assert mc_reach.binary_find(arr, [9, 2]).tolist() == [2, 0]
    assert mc_reach.binary_find(arr, []).size == 0
    with pytest.raises(ValueError):          # mc_reach.pyx:65
        mc_reach.binary_find(arr, [3])
    with pytest.raises(ValueError):
        mc_reach.binary_find(arr, [41])


def test_column_mapper():
    cols = ["dt", "bw", "tw", "twcc", "dx", "n", "ncc", "cs", "s0", "alt"]
    assert mc_reach.column_mapper(cols) == [0, 4, 1, 2, 3, 5, 6, 7, 8]


@pytest.mark.parametrize("seed", [1, 2])
def test_flatten_reaches_reproduces_the_graph(seed):
    down = synth.hack_tree(3000, seed=seed)
    reaches, ups = synth.reaches_from_down(down)
    assert max(len(r) for r in reaches) > 1
    # ids are not rows: scramble ids so that binary_find does real work
    ids = np.sort(np.random.default_rng(seed).choice(10**6, size=down.size, replace=False)).astype(np.int64)
    reaches_ids = [([int(ids[s]) for s in r], 0) for r in reaches]
    ups_ids = {int(ids[k]): [int(ids[u]) for u in v] for k, v in ups.items()}
    up_ptr, up_rows, kind, seg_rows, reach_len, reach_type = mc_reach.flatten_network(reaches_ids, ups_ids, ids)
    p2, r2 = synth.upstream_csr(down)
    assert np.array_equal(up_ptr, p2)
    for r in range(down.size):
        assert sorted(up_rows[up_ptr[r]:up_ptr[r + 1]].tolist()) == sorted(r2[p2[r]:p2[r + 1]].tolist())
    assert (kind == TRT_KIND_MC).all()
    assert int(reach_len.sum()) == down.size


def test_flatten_kinds_and_errors():
    data_idx = np.array([10, 20, 30, 40, 50], dtype=np.int64)
    reaches = [([10, 20], 0), ([30], 1), ([40], 0)]            # 50 belongs to no reach -> boundary row
    ups = {10: [], 30: [20, 50], 40: [30]}
    up_ptr, up_rows, kind, *_ = mc_reach.flatten_network(reaches, ups, data_idx)
    assert kind.tolist() == [TRT_KIND_MC, TRT_KIND_MC, TRT_KIND_LEVELPOOL, TRT_KIND_MC, TRT_KIND_BOUNDARY]
    assert up_rows[up_ptr[2]:up_ptr[3]].tolist() == [1, 4]     # upstream list order is kept (summation order)
    assert up_rows[up_ptr[1]:up_ptr[2]].tolist() == [0]
    with pytest.raises(ValueError):
        mc_reach.flatten_network([([10, 99], 0)], {}, data_idx)
    with pytest.raises(ValueError):
        mc_reach.flatten_network([([10], 0), ([10], 0)], {}, data_idx)
    with pytest.raises(ValueError):
        mc_reach.flatten_network([([10, 20], 1)], {}, data_idx)


def test_network_cache_key_sees_the_connectivity():
    """Same ids, same parameters, same reach lengths / types / end points -- but a different confluence wiring or a
    different interior membership of the reaches -- must not share a cached device network (mc_reach._fingerprint)."""
    data_idx = np.arange(10, 90, 10, dtype=np.int64)                      # 10 .. 80
    cols = ["dt", "bw", "tw", "twcc", "dx", "n", "ncc", "cs", "s0", "alt"]
    vals = np.ones((data_idx.size, len(cols)), dtype=np.float32)
    reaches = [([10, 20, 30], 0), ([40, 50, 60], 0), ([70], 0), ([80], 0)]
    ups = {10: [], 40: [], 70: [30], 80: [60]}
    key = mc_reach._fingerprint(reaches, ups, data_idx, cols, vals, 0)
    assert key == mc_reach._fingerprint([(list(r), t) for r, t in reaches], dict(ups), data_idx.copy(), list(cols), vals.copy(), 0)
    rewired = {10: [], 40: [], 70: [60], 80: [30]}                        # the two confluences swapped
    assert mc_reach._fingerprint(reaches, rewired, data_idx, cols, vals, 0) != key
    swapped = [([10, 50, 30], 0), ([40, 20, 60], 0), ([70], 0), ([80], 0)]   # interior segments exchanged between reaches
    assert mc_reach._fingerprint(swapped, ups, data_idx, cols, vals, 0) != key
    assert mc_reach._fingerprint(reaches, ups, data_idx, cols, vals, 1) != key
