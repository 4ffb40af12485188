"""Diffusive-domain builder: which segments are routed with the diffusive wave, which Muskingum-Cunge segments feed them, and
what is left for the Muskingum-Cunge engine.  Mirror of MCwithDiffusive.update_routing_domain
(/root/reference/src/troute-network/troute/AbstractRouting.py:209-328) on plain dicts and one DataFrame; the reference
version lives in the hydrofabric classes (xarray / geopandas), this one only needs the connection dicts.

    build_diffusive_network_data(diffusive_domain, connections, dataframe, waterbody_ids=(), bad_topobathy_links=())
        -> (diffusive_network_data, dataframe_mc, connections_mc)

`diffusive_domain` is the content of the domain file (:14-35): {tailwater id: {"headwater": [ids ...], ...}}.  A headwater
list containing 999999 means "walk upstream from the tailwater until a headwater, a waterbody or one of the other listed
ids" (:224-243); otherwise its first entry is the upstream end of a single mainstem (:361-380) and becomes a tributary
boundary itself (:266-271).  Per tailwater the result holds what compute_diffusive_routing reads: mainstem_segs,
tributary_segments, connections, rconn, reaches, param_df, upstream_boundary_link (:274-312).  The returned Muskingum-Cunge
connections have the mainstem removed and every tributary segment turned into a tailwater (:314-328).
"""
from itertools import chain

from . import diffusive_utils


def _upstream_closure(rconn, source, targets):
    """segments reachable upstream of `source`, not walking past `targets` (nhd_network.reachable with targets)"""
    targets = set(targets)
    seen, stack = [], [source]
    mark = set()
    while stack:
        n = stack.pop()
        if n in mark:
            continue
        mark.add(n); seen.append(n)
        if n not in targets:
            stack.extend(rconn.get(n, []))
    return seen


def build_diffusive_network_data(diffusive_domain, connections, dataframe, waterbody_ids=(), bad_topobathy_links=()):
    """`bad_topobathy_links`: segments of the domain for which no surveyed cross section can be found (hyfeatures.
    complete_topobathy); the reference stops the upstream walk at them and leaves them to Muskingum-Cunge (:258-264)."""
    connections = {k: list(v) for k, v in connections.items()}
    rconn0 = {k: [] for k in connections}
    for k, dsts in connections.items():
        for d in dsts:
            rconn0.setdefault(d, []).append(k)
    wbody_ids = list(waterbody_ids)
    outlet_ids = list(chain.from_iterable(connections.get(w, []) for w in wbody_ids))
    bad = list(bad_topobathy_links)
    excluded = set(wbody_ids) | set(outlet_ids) | set(bad)
    out = {}
    for tw, spec in diffusive_domain.items():
        heads = list(spec["headwater"] if isinstance(spec, dict) else spec)
        boundary_links = []
        if 999999 in heads:
            targets = [h for h in heads if h != 999999] + wbody_ids + bad
            links = _upstream_closure(rconn0, tw, targets)
        else:
            # single mainstem between a given head and the tailwater
            links, n = [heads[0]], heads[0]
            while n != tw:
                n = connections[n][0]
                links.append(n)
            boundary_links = [heads[0]]
        mainstem = [s for s in links if s not in excluded and s not in boundary_links]
        in_main = set(mainstem)
        tribs = [u for s in mainstem for u in rconn0.get(s, []) if u not in in_main]
        dconn = {k: list(connections[k]) for k in mainstem + tribs}
        dconn[tw] = []
        rconn = {k: [] for k in dconn}
        for k, dsts in dconn.items():
            for d in dsts:
                rconn[d].append(k)
        reaches = [r for _, r in diffusive_utils._decompose(tw, rconn, set(tribs))]
        out[tw] = dict(mainstem_segs=mainstem, tributary_segments=tribs, connections=dconn, rconn=rconn, reaches=reaches,
                       param_df=dataframe.filter(mainstem + tribs, axis=0), upstream_boundary_link=boundary_links)
        dataframe = dataframe.drop([s for s in mainstem if s in dataframe.index])
        for s in mainstem:
            connections.pop(s)
        for t in tribs:
            connections[t] = []
    return out, dataframe, connections
