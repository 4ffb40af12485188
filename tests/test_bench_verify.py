"""bench.py `verify.hash`: the per-rank, per-window checksums of a sharded run combine to the value an unsharded run prints
(N = 1/2/4/8 must show the same hash).  The device checksum itself is tested on the GPU
(tests/test_gpu_continue.py); this is the host-side combination rule."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_rank_checksums_add_up_per_window_whatever_the_sharding():
    import bench
    rng = np.random.default_rng(1)
    W = 7
    rows = rng.integers(0, 1 << 63, size=(1000, W), dtype=np.int64).astype(np.uint64)      # checksum of every row and window
    whole = [[int(rows[:, w].sum(dtype=np.uint64)) for w in range(W)]]
    want = bench.Verifier.combine(whole)
    for shards in (2, 4, 8):
        owner = rng.integers(0, shards, size=rows.shape[0])
        per_rank = [[int(rows[owner == r, w].sum(dtype=np.uint64)) for w in range(W)] for r in range(shards)]
        assert bench.Verifier.combine(per_rank) == want
    # the window index is part of the value: swapping two windows changes it
    swapped = [list(whole[0])]
    swapped[0][0], swapped[0][1] = swapped[0][1], swapped[0][0]
    assert bench.Verifier.combine(swapped) != want
