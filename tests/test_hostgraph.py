"""Host graph utilities vs outputs recorded from the reference's troute/nhd_network.py (tests/golden/nhd_graph.json,
made by tests/golden/make_golden.py).  CPU only."""
import json
import os

import numpy as np
import pytest

from troute_b200 import hostgraph, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _flat(connections):
    ids = np.asarray(sorted(int(k) for k in connections), dtype=np.int64)
    row = {int(s): i for i, s in enumerate(ids)}
    down = np.full(ids.size, -1, dtype=np.int64)
    for k, v in connections.items():
        if v:
            assert len(v) == 1
            down[row[int(k)]] = row[int(v[0])]
    up_ptr, up_rows = synth.upstream_csr(down)
    return ids, down, up_ptr, up_rows


@pytest.mark.parametrize("name", ["fixture", "forest300"])
def test_reverse_network_matches_reference(name):
    g = json.load(open(os.path.join(GOLD, "nhd_graph.json")))[name]
    ids, down, up_ptr, up_rows = _flat(g["connections"])
    for k, ups in g["rconn"].items():
        r = int(np.searchsorted(ids, int(k)))
        assert sorted(ids[up_rows[up_ptr[r]:up_ptr[r + 1]]].tolist()) == sorted(ups)
    if name == "fixture":      # the literal expected_rconn of test_nhd_network.py:139
        for k, ups in g["expected_rconn_reference_literal"].items():
            r = int(np.searchsorted(ids, int(k)))
            assert sorted(ids[up_rows[up_ptr[r]:up_ptr[r + 1]]].tolist()) == sorted(ups)


@pytest.mark.parametrize("name", ["fixture", "forest300"])
def test_build_subnetworks_matches_reference(name):
    g = json.load(open(os.path.join(GOLD, "nhd_graph.json")))[name]
    ids, down, up_ptr, up_rows = _flat(g["connections"])
    orders = hostgraph.build_subnetworks(down, up_ptr, up_rows, g["subnetworks_target_size"])
    mine = {}
    for o, subs in enumerate(orders):
        for tw, members in subs.items():
            mine[(o, int(ids[tw]))] = sorted(ids[members].tolist())
    ref = {}
    for tw, by_order in g["subnetworks"].items():
        for o, d in by_order.items():
            for sn_tw, segs in d.items():
                ref[(int(o), int(sn_tw))] = sorted(segs)
    assert mine == ref


def test_reaches_and_levels_respect_reference_reach_order():
    """Every reach of the reference's dfs_decomposition (reaches_bytw) is a chain in `down`, and the level order is
    upstream-first: a segment's level exceeds that of everything upstream of it."""
    g = json.load(open(os.path.join(GOLD, "nhd_graph.json")))["forest300"]
    ids, down, up_ptr, up_rows = _flat(g["connections"])
    lvl = hostgraph.levels(down, up_ptr)
    row = {int(s): i for i, s in enumerate(ids)}
    seen = set()
    for tw, reaches in g["reaches_bytw"].items():
        for reach in reaches:
            for a, b in zip(reach[:-1], reach[1:]):
                assert down[row[a]] == row[b] and lvl[row[b]] == lvl[row[a]] + 1
            seen.update(reach)
    assert seen == set(int(x) for x in ids)
    d = down >= 0
    assert (lvl[down[d]] > lvl[d.nonzero()[0]]).all()


def test_jobs_cover_every_segment_once_and_respect_orders():
    down = synth.hack_tree(30000, seed=3)
    up_ptr, up_rows = synth.upstream_csr(down)
    r = hostgraph.segment_reaches_level_order(down, up_ptr, up_rows)
    jobs = hostgraph.subnetwork_jobs(down, up_ptr, up_rows, r["order"], target_size=2000)
    assert np.array_equal(np.sort(jobs["job_reaches"]), np.arange(30000))
    # order index of every segment; a segment's upstream neighbours are in the same job or an earlier order
    order_of_job = np.repeat(np.arange(len(jobs["order_ptr"]) - 1), np.diff(jobs["order_ptr"]))
    job_of_reach = np.repeat(np.arange(len(jobs["job_ptr"]) - 1), np.diff(jobs["job_ptr"]))
    job_of_row = np.empty(30000, dtype=np.int64)
    job_of_row[r["order"][jobs["job_reaches"]]] = job_of_reach
    d = np.nonzero(down >= 0)[0]
    same = job_of_row[d] == job_of_row[down[d]]
    earlier = order_of_job[job_of_row[d]] < order_of_job[job_of_row[down[d]]]
    assert (same | earlier).all()
