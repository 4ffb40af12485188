"""Synthetic river networks and forcing for tests and bench.py.

Networks are described by ONE array, `down[i]` = id of the segment that segment i drains into
(-1 for an outlet).  Segment ids are 0..n-1, so the sorted segment index `data_idx` of the reference
is arange(n) and "row" == id.  Parameters follow the CONUS NWM 2.1 statistics the reference's own
kernel test-suite quotes (/root/reference/src/kernel/muskingum/test_suite_parameters.py:4-13).
"""
import numpy as np

PARAM_COLS = ["dt", "bw", "tw", "twcc", "dx", "n", "ncc", "cs", "s0", "alt"]   # compute.py:1443-1446 column order


# -------------------------------------------------------------------------------------------------
# topology
# -------------------------------------------------------------------------------------------------
def binary_tree(n):
    """Balanced binary tree: node i drains to (i-1)//2, node 0 is the outlet (config 2)."""
    down = (np.arange(n, dtype=np.int64) - 1) // 2
    down[0] = -1
    return down


def chain(n):
    """A single reach of n segments in series: i+1 drains into i."""
    down = np.arange(n, dtype=np.int64) - 1
    return down


def hack_tree(n, seed=16, hack_c=1.6, hack_h=0.55, trib_shape=1.15):
    """One river basin of n segments with Hack's-law depth: mainstem length ~ hack_c * n**hack_h.

    Built top-down with an explicit stack.  A (sub)basin of size m draining to `parent` gets a mainstem
    of length l = clip(round(hack_c * m**hack_h), 1, m); the remaining m - l segments are split into
    tributary sub-basins with heavy-tailed (Pareto, shape `trib_shape`) sizes, each attached to a random
    mainstem node.  Returns down[n].
    """
    rng = np.random.default_rng(seed)
    down = np.full(n, -1, dtype=np.int64)
    next_id = 0
    stack = [(n, -1)]
    while stack:
        m, parent = stack.pop()
        length = int(min(m, max(1, round(hack_c * m ** hack_h))))
        ids = np.arange(next_id, next_id + length, dtype=np.int64)
        next_id += length
        down[ids[0]] = parent
        if length > 1:
            down[ids[1:]] = ids[:-1]
        rest = m - length
        if rest <= 0:
            continue
        # tributary sizes: Pareto draws until `rest` is used up
        sizes = []
        cap = max(1, rest // 2) if rest > 3 else rest
        while rest > 0:
            s = int(min(rest, cap, np.floor(rng.pareto(trib_shape) + 1.0)))
            s = max(s, 1)
            sizes.append(s)
            rest -= s
        sizes = np.asarray(sizes, dtype=np.int64)
        attach = ids[rng.integers(0, length, size=sizes.shape[0])]
        for s, a in zip(sizes.tolist(), attach.tolist()):
            stack.append((s, a))
    assert next_id == n
    return down


# tributaries per interior main-stem node -> in-degree 1 + k.  Tuned so that a whole forest reproduces the in-degree
# mix of the LowerColorado v4 hydrofabric quoted in SURVEY.md 8(d): 0: 48 %, 1: 19 %, 2: 23 %, 3: 6 %, >= 4: 4 %.
_NHD_TRIB_P = np.array([0.365, 0.44, 0.115, 0.055, 0.025])


def nhd_tree(n, seed=16, hack_c=1.15, hack_h=0.55, trib_shape=1.15, trib_p=_NHD_TRIB_P):
    """One river basin of n segments with Hack's-law depth AND NHD-like confluences: every interior main-stem node
    receives 0..4 tributaries (mostly 0 or 1: binary confluences), never dozens as hack_tree() allows.  The remaining
    m - l segments of a (sub)basin are split over its tributaries with heavy-tailed (Pareto) sizes, each tributary
    being a basin of the same kind.  Returns down[n]."""
    rng = np.random.default_rng(seed)
    down = np.full(n, -1, dtype=np.int64)
    next_id = 0
    stack = [(n, -1)]
    kmax = trib_p.shape[0] - 1
    trib_cdf = np.cumsum(trib_p)[:-1]
    while stack:
        m, parent = stack.pop()
        length = int(min(m, max(1, round(hack_c * m ** hack_h))))
        if m >= 2:
            length = max(length, 2)                       # a basin with tributaries needs an interior node
        ids = np.arange(next_id, next_id + length, dtype=np.int64)
        next_id += length
        down[ids[0]] = parent
        if length > 1:
            down[ids[1:]] = ids[:-1]
        rest = m - length
        if rest <= 0:
            continue
        # ids[0] is the outlet, ids[-1] the headwater of this main stem; interior nodes = all but the headwater
        k = np.searchsorted(trib_cdf, rng.random(length - 1), side="right")
        K = int(k.sum())
        if K > rest:                                      # thin basin: drop tributaries at random
            slots = np.repeat(np.arange(length - 1), k)
            keep = rng.choice(slots.shape[0], size=rest, replace=False)
            k = np.bincount(slots[keep], minlength=length - 1)
            K = rest
        elif K == 0:
            k[rng.integers(0, length - 1)] = 1
            K = 1
        w = rng.pareto(trib_shape, size=K) + 1.0
        sizes = 1 + np.floor(w / w.sum() * (rest - K)).astype(np.int64)
        sizes[np.argmax(w)] += rest - int(sizes.sum())
        attach = np.repeat(ids[:-1], k)
        rng.shuffle(sizes)
        one = sizes == 1                                  # single-segment tributaries: no recursion needed
        c1 = int(one.sum())
        if c1:
            down[next_id:next_id + c1] = attach[one]
            next_id += c1
        for s_, a_ in zip(sizes[~one].tolist(), attach[~one].tolist()):
            stack.append((s_, a_))
    assert next_id == n
    return down


def conus_like(n_total=2_729_077, n_basins=14_713, largest_frac=0.5, seed=16, hack_c=None, hack_h=0.55, style="hack"):
    """CONUS-scale forest (config 3): n_basins independent basins, the largest holding ~largest_frac of
    all segments (doc/AGU_Poster.md:35-41, :208-214).  style "hack": hack_tree() basins (main stems collect dozens
    of tributaries per node -- a stress case for the upstream gather); style "nhd": nhd_tree() basins (NHD-like
    confluences and in-degree mix, the bench workload).  Returns down[n_total]."""
    tree = hack_tree if style == "hack" else nhd_tree
    if hack_c is None:
        hack_c = 1.6 if style == "hack" else 1.15
    rng = np.random.default_rng(seed)
    big = int(n_total * largest_frac)
    rest = n_total - big
    # remaining basins: Pareto sizes normalised to `rest`
    w = rng.pareto(0.9, size=n_basins - 1) + 1.0
    sizes = np.maximum(1, np.floor(w / w.sum() * rest).astype(np.int64))
    diff = rest - int(sizes.sum())
    sizes[np.argmax(sizes)] += diff
    assert sizes.min() >= 1 and int(sizes.sum()) == rest
    parts = [tree(big, seed=seed + 1, hack_c=hack_c, hack_h=hack_h)]
    offs = [0]
    off = big
    for i, s in enumerate(sizes.tolist()):
        t = tree(s, seed=seed + 2 + i, hack_c=hack_c, hack_h=hack_h)
        t = np.where(t >= 0, t + off, -1)
        parts.append(t)
        offs.append(off)
        off += s
    return np.concatenate(parts)


def basin_of(down):
    """Outlet segment of the basin every segment belongs to (pointer jumping: log2(depth) vectorised rounds)."""
    n = down.shape[0]
    nxt = np.where(down >= 0, down, np.arange(n, dtype=down.dtype))
    while True:
        nn = nxt[nxt]
        if np.array_equal(nn, nxt):
            return nxt
        nxt = nn


def upstream_csr(down):
    """CSR of upstream ids per segment, upstream ids ascending (the order nhd_network.reverse_network
    yields for sorted keys)."""
    n = down.shape[0]
    src = np.nonzero(down >= 0)[0].astype(np.int64)
    dst = down[src]
    order = np.argsort(dst, kind="stable")
    up_rows = src[order]
    counts = np.bincount(dst, minlength=n)
    up_ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(counts, out=up_ptr[1:])
    return up_ptr, up_rows


def subtree_sizes(down):
    """Number of segments draining through each segment (itself included)."""
    n = down.shape[0]
    up_ptr, up_rows = upstream_csr(down)
    indeg = np.diff(up_ptr).astype(np.int64)
    size = np.ones(n, dtype=np.int64)
    frontier = np.nonzero(indeg == 0)[0]
    remaining = indeg.copy()
    while frontier.size:
        d = down[frontier]
        ok = d >= 0
        np.add.at(size, d[ok], size[frontier[ok]])
        np.subtract.at(remaining, d[ok], 1)
        cand = np.unique(d[ok])
        frontier = cand[remaining[cand] == 0]
    return size


def levels_from_down(down):
    """Longest path (in segments) from a headwater to each segment."""
    n = down.shape[0]
    up_ptr, _ = upstream_csr(down)
    remaining = np.diff(up_ptr).astype(np.int64)
    level = np.zeros(n, dtype=np.int64)
    frontier = np.nonzero(remaining == 0)[0]
    while frontier.size:
        d = down[frontier]
        ok = d >= 0
        np.maximum.at(level, d[ok], level[frontier[ok]] + 1)
        np.subtract.at(remaining, d[ok], 1)
        cand = np.unique(d[ok])
        frontier = cand[remaining[cand] == 0]
    return level


def reaches_from_down(down):
    """Reference-style structures for SMALL networks: (reaches, upstream_connections).

    Reaches are maximal chains of in-degree-1 / out-degree-1 segments (dfs_decomposition semantics,
    nhd_network.py:503-557), listed upstream-first; upstream_connections maps a segment id to the list
    of its upstream ids (ascending)."""
    n = down.shape[0]
    up_ptr, up_rows = upstream_csr(down)
    ups = {int(i): up_rows[up_ptr[i]:up_ptr[i + 1]].tolist() for i in range(n)}
    indeg = np.diff(up_ptr)
    level = levels_from_down(down)
    heads = [i for i in range(n) if indeg[i] != 1]
    # a segment starts a reach when it is a headwater or a junction (in-degree != 1)
    reaches = []
    for h in sorted(heads, key=lambda i: (int(level[i]), i)):
        r = [h]
        cur = h
        while True:
            d = int(down[cur])
            if d < 0 or indeg[d] != 1:
                break
            r.append(d)
            cur = d
        reaches.append(r)
    # order reaches so that every reach comes after the reaches feeding its head
    reaches.sort(key=lambda r: int(level[r[0]]))
    return reaches, ups


# -------------------------------------------------------------------------------------------------
# parameters and forcing
# -------------------------------------------------------------------------------------------------
def _lognormal(rng, mean, sd, size):
    sigma2 = np.log(1.0 + (sd / mean) ** 2)
    mu = np.log(mean) - 0.5 * sigma2
    return rng.lognormal(mu, np.sqrt(sigma2), size)


def channel_params(down, dt=300.0, seed=16):
    """[n, 10] float32 parameter table with columns PARAM_COLS, CONUS NWM 2.1 statistics clipped to
    [min, max] (test_suite_parameters.py:4-13); bottom width grows with contributing segments."""
    rng = np.random.default_rng(seed)
    n = down.shape[0]
    acc = subtree_sizes(down).astype(np.float64)
    dx = np.clip(_lognormal(rng, 1947.776, 1965.625, n), 1.0, 95714.0)
    bw = np.clip(0.8 * acc ** 0.4 * _lognormal(rng, 1.0, 0.35, n), 0.135, 230.035)
    tw = bw / 0.6
    twcc = 3.0 * tw
    nman = np.clip(rng.normal(0.058, 0.003, n), 0.040, 0.060)
    ncc = 2.0 * nman
    cs = np.clip(rng.normal(0.5857, 0.1945, n), 0.0846, 2.254)
    s0 = np.clip(_lognormal(rng, 0.02150, 0.04585, n), 0.00001, 4.6)
    alt = np.zeros(n)
    table = np.stack([np.full(n, dt), bw, tw, twcc, dx, nman, ncc, cs, s0, alt], axis=1)
    return table.astype(np.float32)


def lateral_inflow(n, nsteps, qts_subdivisions=12, seed=16, storms=((8.0, 2.0, 8.0),)):
    """[n, ceil(nsteps/qts)] float32: per-segment base U(0.01, 0.1) m3/s times storm pulses
    1 + sum_i a_i exp(-(hour - c_i)^2 / w_i); `storms` = ((c, a, w), ...), default one pulse 1 + 2 exp(-(hour-8)^2/8)."""
    rng = np.random.default_rng(seed + 1000)
    ncol = int(np.ceil(nsteps / qts_subdivisions))
    base = rng.uniform(0.01, 0.1, n)
    hour = np.arange(ncol, dtype=np.float64)
    pulse = np.ones(ncol)
    for c, a, w in storms:
        pulse = pulse + a * np.exp(-((hour - c) ** 2) / w)
    return (base[:, None] * pulse[None, :]).astype(np.float32)


def levelpool_params(n_lp, seed=16):
    """[n_lp, 11] float64 waterbody rows in the fixture ranges of
    reservoirs/test/test_compute_kernel.py:28-110 (LkArea km2, LkMxE m, OrificeA, OrificeC, OrificeE, WeirC,
    WeirE, WeirL, ifd, qd0, h0)."""
    rng = np.random.default_rng(seed + 2000)
    area = rng.uniform(0.2, 30.0, n_lp)
    base = rng.uniform(5.0, 500.0, n_lp)
    orifice_e = base
    weir_e = base + rng.uniform(3.0, 12.0, n_lp)
    max_e = weir_e + rng.uniform(1.0, 4.0, n_lp)
    h0 = orifice_e + (weir_e - orifice_e) * rng.uniform(0.5, 1.05, n_lp)
    rows = np.stack([area, max_e, rng.uniform(0.5, 3.0, n_lp), np.full(n_lp, 0.1), orifice_e,
                     np.full(n_lp, 0.4), weir_e, rng.uniform(5.0, 40.0, n_lp), np.full(n_lp, 0.9),
                     rng.uniform(0.0, 5.0, n_lp), h0], axis=1)
    return rows.astype(np.float64)
