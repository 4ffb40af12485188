"""Host-side graph utilities on flat arrays (numpy): what the reference does with dict-of-lists in
troute/nhd_network.py, restated for networks of millions of segments.

A network is `down[i]` = row of the segment that row i drains into (-1 = outlet), plus the upstream CSR
(`up_ptr`, `up_rows`).  Nothing here touches the GPU; the arithmetic path never depends on these decompositions
(the engine levels the graph itself) -- they exist for the CPU baseline's job decomposition and for sharding.
"""
from collections import deque

import numpy as np


def levels(down, up_ptr):
    """Longest path (in segments) from a headwater; a Kahn sweep, vectorised per frontier."""
    n = down.shape[0]
    remaining = np.diff(up_ptr).astype(np.int64)
    level = np.zeros(n, dtype=np.int32)
    frontier = np.nonzero(remaining == 0)[0]
    while frontier.size:
        d = down[frontier]
        ok = d >= 0
        np.maximum.at(level, d[ok], level[frontier[ok]] + 1)
        np.subtract.at(remaining, d[ok], 1)
        cand = np.unique(d[ok])
        frontier = cand[remaining[cand] == 0]
    return level


def segment_reaches_level_order(down, up_ptr, up_rows):
    """Every segment as a one-segment reach, listed level by level (a valid upstream-first order, which is all
    mc_reach.pyx:493 needs).  Returns the flat reach arrays of oracle.route_network_flat plus `order`."""
    n = down.shape[0]
    lvl = levels(down, up_ptr)
    order = np.argsort(lvl, kind="stable").astype(np.int64)
    counts = np.diff(up_ptr)[order]
    reach_up_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=reach_up_ptr[1:])
    total = int(counts.sum())
    starts = up_ptr[order]
    idx = np.arange(total, dtype=np.int64) - np.repeat(reach_up_ptr[:-1], counts) + np.repeat(starts, counts)
    return dict(order=order, level=lvl, reach_ptr=np.arange(n + 1, dtype=np.int64), reach_rows=order,
                reach_type=np.zeros(n, dtype=np.int32), reach_up_ptr=reach_up_ptr,
                reach_up_rows=up_rows[idx] if total else np.zeros(0, np.int64))


def build_subnetworks(down, up_ptr, up_rows, target_size, sources=None):
    """Truncated breadth-first decomposition into ordered sub-networks -- the algorithm of
    nhd_network.build_subnetworks (nhd_network.py:691-771) on flat arrays.

    From every tailwater, grow a sub-network upstream breadth-first; the junction depth `y` increases by one when
    a segment with more than one upstream neighbour is crossed; once more than `target_size` segments have been
    reached the current depth becomes the stop depth.  Segments on the fringe that are not real headwaters become
    the sources (tailwaters) of the next order.  Returns a list over orders (0 = most downstream) of dicts
    {tailwater_row: np.ndarray of member rows}."""
    n = down.shape[0]
    upp = up_ptr.tolist()
    upr = up_rows.tolist()
    if sources is None:
        sources = np.nonzero(down < 0)[0].tolist()
    indeg = np.diff(up_ptr)
    orders = []
    new_sources = list(sources)
    in_sub = np.zeros(n, dtype=bool)     # scratch: membership of the sub-network being grown
    while new_sources:
        this_order = {}
        next_sources = []
        for h in new_sources:
            reachable = []
            Q = deque([(h, 0)])
            stop_depth = 1000000
            while Q:
                x, y = Q.popleft()
                reachable.append(x)
                in_sub[x] = True
                a, b = upp[x], upp[x + 1]
                us_depth = y + 1 if b - a > 1 else y
                if len(reachable) > target_size:
                    stop_depth = y
                if us_depth <= stop_depth:
                    for e in range(a, b):
                        Q.append((upr[e], us_depth))
            members = np.asarray(reachable, dtype=np.int64)
            # apparent headwaters of the sub-network: members none of whose upstream rows is a member.  Real
            # headwaters stay; the others seed the next order and leave this sub-network (:752-761)
            cnt = np.zeros(members.shape[0], dtype=np.int64)
            deg = indeg[members]
            has_up = deg > 0
            if has_up.any():
                m = members[has_up]
                starts = up_ptr[m]
                c = deg[has_up]
                offs = np.zeros(c.shape[0] + 1, dtype=np.int64)
                np.cumsum(c, out=offs[1:])
                idx = np.arange(int(offs[-1]), dtype=np.int64) - np.repeat(offs[:-1], c) + np.repeat(starts, c)
                inside = in_sub[up_rows[idx]]
                cnt[has_up] = np.add.reduceat(inside.astype(np.int64), offs[:-1])
            fringe = has_up & (cnt == 0)
            in_sub[members] = False
            srcs = members[fringe]
            next_sources.extend(srcs.tolist())
            this_order[int(h)] = members[~fringe]
        orders.append(this_order)
        new_sources = next_sources
    return orders


def subnetwork_jobs(down, up_ptr, up_rows, order, target_size=10000, cluster_fraction=0.65):
    """Job decomposition of compute.py's `by-subnetwork-jit-clustered` mode (compute.py:553-907): sub-networks from
    build_subnetworks, grouped by order, orders run from the most upstream to the most downstream (:975), the
    sub-networks of one order clustered until a cluster holds >= cluster_fraction * target_size segments
    (:604, :631-646).  `order` is the reach list's row order (one-segment reaches); the result indexes into it.

    Returns dict(order_ptr, job_ptr, job_reaches) for oracle.route_network_flat(jobs=...)."""
    n = down.shape[0]
    pos_in_order = np.empty(n, dtype=np.int64)
    pos_in_order[order] = np.arange(n, dtype=np.int64)
    orders = build_subnetworks(down, up_ptr, up_rows, target_size)
    order_ptr = [0]
    job_ptr = [0]
    job_reaches = []
    njobs = 0
    for subs in reversed(orders):
        cluster = []
        size = 0
        for tw, members in subs.items():
            if members.size == 0:
                continue
            cluster.append(members)
            size += members.size
            if size >= cluster_fraction * target_size:
                r = np.sort(pos_in_order[np.concatenate(cluster)])
                job_reaches.append(r); job_ptr.append(job_ptr[-1] + r.size); njobs += 1
                cluster, size = [], 0
        if cluster:
            r = np.sort(pos_in_order[np.concatenate(cluster)])
            job_reaches.append(r); job_ptr.append(job_ptr[-1] + r.size); njobs += 1
        order_ptr.append(njobs)
    return dict(order_ptr=np.asarray(order_ptr, dtype=np.int64), job_ptr=np.asarray(job_ptr, dtype=np.int64),
                job_reaches=np.concatenate(job_reaches) if job_reaches else np.zeros(0, np.int64),
                n_orders=len(orders))
