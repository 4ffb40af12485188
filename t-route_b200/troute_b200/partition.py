"""Sub-basin sharding of a river network over P GPUs.

The reference parallelises over independent networks (compute.py `by-network`, bounded by the Mississippi holding
half of all segments, doc/AGU_Poster.md:208-214) or over ordered sub-networks that hand complete tail-water time series
from one order to the next (compute.py:909-1209).  Here every shard routes all timesteps concurrently; what crosses a
cut edge u -> s is the single float q[u, t], written by the upstream shard's kernel directly into the downstream GPU's
flow array (network.RoutingNetwork.set_exports / set_imports).

Partitioning (host, numpy):
  1. pieces: walk the forest upstream-first accumulating the residual size r(v) = 1 + sum r(children); when r(v)
     reaches `piece_target` the residual sub-basin below v is cut off as a piece (its outlet edge becomes a cut edge);
     what is left at an outlet is a piece as well.  Pieces are connected sub-basins, so cut edges = pieces - outlets.
  2. shards: longest-processing-time bin packing of the pieces by segment count (every segment costs the same:
     one lane-step per timestep).
The quotient graph of shards may contain cycles (A feeds B feeds A through different pieces); that is fine, the
dependency is per (segment, timestep), not per shard.
"""
import heapq

import numpy as np

from . import hostgraph


def make_pieces(down, up_ptr, level, piece_target):
    """-> piece id per segment (0..n_pieces-1) and the root segment of every piece."""
    n = down.shape[0]
    resid = np.ones(n, dtype=np.int64)
    is_root = down < 0
    order = np.argsort(level, kind="stable")
    lvl_sorted = level[order]
    nlev = int(lvl_sorted[-1]) + 1 if n else 0
    bounds = np.searchsorted(lvl_sorted, np.arange(nlev + 1))
    for l in range(nlev):                                  # children (lower levels) before parents
        v = order[bounds[l]:bounds[l + 1]]
        cut = (resid[v] >= piece_target) & (down[v] >= 0)
        is_root[v[cut]] = True
        keep = v[~cut & (down[v] >= 0)]
        np.add.at(resid, down[keep], resid[keep])
    roots = np.nonzero(is_root)[0]
    piece = np.full(n, -1, dtype=np.int64)
    piece[roots] = np.arange(roots.size)
    for l in range(nlev - 1, -1, -1):                      # parents before children
        v = order[bounds[l]:bounds[l + 1]]
        v = v[piece[v] < 0]
        piece[v] = piece[down[v]]
    return piece, roots


def assign_shards(piece, n_pieces, n_shards):
    """LPT bin packing: biggest piece first into the lightest shard."""
    sizes = np.bincount(piece, minlength=n_pieces)
    heap = [(0, s) for s in range(n_shards)]
    heapq.heapify(heap)
    shard_of_piece = np.empty(n_pieces, dtype=np.int64)
    for p in np.argsort(-sizes, kind="stable"):
        load, s = heapq.heappop(heap)
        shard_of_piece[p] = s
        heapq.heappush(heap, (load + int(sizes[p]), s))
    return shard_of_piece


class ShardPlan:
    """Everything rank `r` needs to build its RoutingNetwork and wire it to its peers.

    rows        global rows held by this shard (own segments + import rows), ascending
    own         mask over `rows`: True for segments this shard routes
    up_ptr/up_rows  local CSR (import rows have no upstream)
    kind        local kinds (imports are TRT_KIND_BOUNDARY)
    levels      levels of the WHOLE network for the local rows
    imports     local rows written by peers;  import_src_shard[i] = shard that owns the segment
    exports     (local row, destination shard, global id of the segment) -- the destination's local row of that
                global id is resolved after an all-gather of every shard's `rows`
    """

    def __init__(self, rank, rows, own, up_ptr, up_rows, kind, levels, imports, import_src, exports):
        self.rank = rank
        self.rows = rows
        self.own = own
        self.up_ptr = up_ptr
        self.up_rows = up_rows
        self.kind = kind
        self.levels = levels
        self.imports = imports
        self.import_src = import_src
        self.exports = exports


def plan_shards(down, up_ptr, up_rows, kind, n_shards, pieces_per_shard=16, level=None):
    """Partition the network; returns (shard_of_segment, [ShardPlan for every rank], stats)."""
    n = down.shape[0]
    if level is None:
        level = hostgraph.levels(down, up_ptr)
    if n_shards == 1:
        shard = np.zeros(n, dtype=np.int64)
    else:
        target = max(1, n // (n_shards * pieces_per_shard))
        piece, roots = make_pieces(down, up_ptr, level, target)
        shard = assign_shards(piece, roots.size, n_shards)[piece]
    src = np.nonzero(down >= 0)[0]
    cut = src[shard[src] != shard[down[src]]]              # cut edges u -> down[u]
    plans = []
    for r in range(n_shards):
        own_rows = np.nonzero(shard == r)[0]
        imp_global = cut[shard[down[cut]] == r]            # segments of other shards draining into this one
        rows = np.union1d(own_rows, imp_global)
        own = np.isin(rows, own_rows, assume_unique=True)
        loc = -np.ones(n, dtype=np.int64)
        loc[rows] = np.arange(rows.size)
        # local CSR: upstream lists of own rows, in the global (reference) order; imports have none
        cnt = np.where(own, np.diff(up_ptr)[rows], 0)
        lptr = np.zeros(rows.size + 1, dtype=np.int64)
        np.cumsum(cnt, out=lptr[1:])
        total = int(lptr[-1])
        starts = up_ptr[rows]
        idx = np.arange(total, dtype=np.int64) - np.repeat(lptr[:-1], cnt) + np.repeat(starts, cnt)
        lup = loc[up_rows[idx]] if total else np.zeros(0, np.int64)
        assert (lup >= 0).all()
        lkind = np.where(own, kind[rows], 2).astype(np.uint8)
        exp_global = cut[shard[cut] == r]
        exports = (loc[exp_global], shard[down[exp_global]], exp_global)
        plans.append(ShardPlan(r, rows, own, lptr, lup, lkind, level[rows].astype(np.int32), loc[imp_global],
                               shard[imp_global], exports))
    sizes = np.bincount(shard, minlength=n_shards)
    stats = dict(n_cut_edges=int(cut.size), shard_sizes=sizes.tolist(),
                 imbalance=float(sizes.max() / max(1.0, sizes.mean())))
    return shard, plans, stats
