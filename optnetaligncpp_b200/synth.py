"""Seeded synthetic inputs of the shapes BASELINE.json names.

Host-side input synthesis only: power-law PPI-like network pairs, sparse BLAST bit-score
tables and GO-style annotation sets.  Everything is produced in the numbering the
reference's parser would assign (node ids in first-appearance order of the edge list,
Network.cpp:59-83), so the same arrays can be fed to the GPU engine, to the C oracle and --
written out as text -- to the compiled reference.
"""
from dataclasses import dataclass, field

import numpy as np

DEFAULT_SEED = 20260000


@dataclass
class NetPair:
    """Two networks in reference numbering + optional score data (all host numpy)."""
    n1: int
    n2: int
    edges1: np.ndarray          # [E1,2] int32, file order (u,v)
    edges2: np.ndarray          # [E2,2] int32
    rp1: np.ndarray = None      # CSR of Network::adjList, adjList order
    col1: np.ndarray = None
    rp2: np.ndarray = None
    col2: np.ndarray = None
    bit_rows: np.ndarray = None  # COO of the non-zero bit scores
    bit_cols: np.ndarray = None
    bit_vals: np.ndarray = None
    ann1: list = None           # per node: sorted int array of terms (empty = unannotated)
    ann2: list = None
    ortholog: np.ndarray = None  # planted image of every net1 node (for information)
    meta: dict = field(default_factory=dict)

    @property
    def E1(self):
        # net1->edges.size(): the reference's edge SET (Network.cpp:87) -- duplicate lines count once
        return self.meta.get("E1", len(self.edges1))

    @property
    def E2(self):
        return self.meta.get("E2", len(self.edges2))


def adjlist_csr(n, edges):
    """CSR of Network::adjList: for every line (u,v) push v to u's list and, unless u==v,
    u to v's list, in file order (Network.cpp:89-92)."""
    u = edges[:, 0].astype(np.int64)
    v = edges[:, 1].astype(np.int64)
    m = len(edges)
    # interleave so that the per-node order equals the reference's push_back order
    src = np.empty(2 * m, np.int64)
    dst = np.empty(2 * m, np.int64)
    src[0::2], dst[0::2] = u, v
    src[1::2], dst[1::2] = v, u
    keep = np.ones(2 * m, bool)
    keep[1::2] = u != v
    src, dst = src[keep], dst[keep]
    order = np.argsort(src, kind="stable")
    col = dst[order].astype(np.int32)
    counts = np.bincount(src, minlength=n)
    rp = np.zeros(n + 1, np.int32)
    np.cumsum(counts, out=rp[1:])
    return rp, col


def canonicalize(edges):
    """Relabel nodes in first-appearance order of the flattened (u0,v0,u1,v1,...) list,
    which is the numbering Network.cpp:59-83 assigns.  Returns (edges', old->new map)."""
    flat = edges.reshape(-1)
    _, first = np.unique(flat, return_index=True)
    order = np.sort(first)
    old_ids = flat[order]
    remap = -np.ones(int(flat.max()) + 1, np.int64)
    remap[old_ids] = np.arange(len(old_ids))
    return remap[edges].astype(np.int32), remap


def _weights(n, gamma, m, max_deg):
    """Chung-Lu weights w_i ~ (i+i0)^(-1/(gamma-1)); i0 tuned so the hub's expected degree
    is about max_deg for m edges."""
    alpha = 1.0 / (gamma - 1.0)
    lo, hi = 0.0, float(n)
    for _ in range(60):
        i0 = 0.5 * (lo + hi)
        w = (np.arange(n) + 1.0 + i0) ** (-alpha)
        d0 = 2.0 * m * w[0] / w.sum()
        if d0 > max_deg:
            lo = i0
        else:
            hi = i0
    w = (np.arange(n) + 1.0 + hi) ** (-alpha)
    return w / w.sum()


def _sample_unique_edges(rng, n, m, w, taken=None):
    """m distinct undirected non-loop edges with endpoints ~ w, avoiding `taken` keys."""
    keys = set() if taken is None else taken
    out = []
    cdf = np.cumsum(w)
    cdf[-1] = 1.0
    while len(out) < m:
        need = m - len(out)
        a = np.searchsorted(cdf, rng.random(2 * need + 16))
        b = np.searchsorted(cdf, rng.random(2 * need + 16))
        for x, y in zip(a, b):
            if x == y:
                continue
            lo, hi = (x, y) if x < y else (y, x)
            k = int(lo) * n + int(hi)
            if k in keys:
                continue
            keys.add(k)
            out.append((int(x), int(y)))
            if len(out) == m:
                break
    return out


def _finish(rng, n, edges, keys, loops, w):
    """add `loops` self-loops on distinct nodes, give every isolated node one edge, shuffle."""
    if loops:
        for u in rng.choice(n, size=loops, replace=False):
            edges.append((int(u), int(u)))
    deg = np.zeros(n, np.int64)
    for u, v in edges:
        deg[u] += 1
        deg[v] += 1
    cdf = np.cumsum(w)
    cdf[-1] = 1.0
    for u in np.nonzero(deg == 0)[0]:
        while True:
            v = int(np.searchsorted(cdf, rng.random()))
            lo, hi = (u, v) if u < v else (v, u)
            if v != u and int(lo) * n + int(hi) not in keys:
                keys.add(int(lo) * n + int(hi))
                edges.append((int(u), v))
                break
    e = np.array(edges, np.int64)
    return e[rng.permutation(len(e))]


def make_pair(n1=6000, n2=8000, m1=30000, m2=30000, gamma=2.5, max_deg=300, loop_frac=0.02,
              keep_frac=0.7, seed=DEFAULT_SEED + 1, bits=True, bit_density=1e-3,
              bit_decimals=False, annotations=False, n_terms=5000, terms_per_node=10,
              annotated_frac=0.9):
    """Synthetic PPI-like pair: net2 contains a relabelled copy of `keep_frac` of net1's
    edges (so a non-trivial optimum exists) plus fresh power-law edges.  No duplicate or
    reversed-duplicate lines; about loop_frac self-loops; every node has an edge."""
    assert n1 <= n2
    rng = np.random.default_rng(seed)
    w1 = _weights(n1, gamma, m1, max_deg)
    l1 = int(round(loop_frac * m1))
    keys1 = set()
    e1 = _sample_unique_edges(rng, n1, m1 - l1, w1, keys1)
    e1 = _finish(rng, n1, e1, keys1, l1, w1)

    # plant net1 inside net2 through a random injection
    inj = rng.permutation(n2)[:n1]
    w2 = np.empty(n2)
    w2[inj] = w1 * n1
    rest = np.setdiff1d(np.arange(n2), inj)
    w2[rest] = rng.permutation(_weights(n2, gamma, m2, max_deg) * n2)[: len(rest)]
    w2 /= w2.sum()
    l2 = int(round(loop_frac * m2))
    nonloop1 = e1[e1[:, 0] != e1[:, 1]]
    kept = nonloop1[rng.random(len(nonloop1)) < keep_frac]
    keys2 = set()
    e2 = []
    for u, v in inj[kept]:
        lo, hi = (u, v) if u < v else (v, u)
        k = int(lo) * n2 + int(hi)
        if k not in keys2:
            keys2.add(k)
            e2.append((int(u), int(v)))
    e2 = e2[: max(0, m2 - l2)]
    e2 += _sample_unique_edges(rng, n2, max(0, m2 - l2 - len(e2)), w2, keys2)
    e2 = _finish(rng, n2, e2, keys2, l2, w2)

    e1c, map1 = canonicalize(e1)
    e2c, map2 = canonicalize(e2)
    assert e1c.max() + 1 == n1 and e2c.max() + 1 == n2
    inv1 = np.empty(n1, np.int64)
    inv1[map1[:n1]] = np.arange(n1)
    orth = map2[inj[inv1]].astype(np.int32)     # image of canonical net1 node u

    pair = NetPair(n1=n1, n2=n2, edges1=e1c, edges2=e2c, ortholog=orth,
                   meta=dict(seed=seed, gamma=gamma, max_deg=max_deg))
    pair.rp1, pair.col1 = adjlist_csr(n1, e1c)
    pair.rp2, pair.col2 = adjlist_csr(n2, e2c)

    if bits:
        nrand = int(round(bit_density * n1 * n2))
        r = np.concatenate([np.arange(n1), rng.integers(0, n1, nrand)])
        c = np.concatenate([orth.astype(np.int64), rng.integers(0, n2, nrand)])
        key, first = np.unique(r * n2 + c, return_index=True)
        r, c = r[first], c[first]
        val = np.exp(rng.normal(4.3, 0.6, len(r)))
        if bit_decimals:
            val = np.round(val, 1)
        else:
            val = np.round(val * 2.0) / 2.0
        val = np.clip(val, 20.0, 3000.0)
        pair.bit_rows, pair.bit_cols, pair.bit_vals = r.astype(np.int32), c.astype(np.int32), val

    if annotations:
        zipf_p = 1.0 / np.arange(1, n_terms + 1) ** 1.1
        zipf_p /= zipf_p.sum()
        ann1 = []
        for u in range(n1):
            if rng.random() < annotated_frac:
                k = max(1, int(rng.poisson(terms_per_node)))
                ann1.append(np.unique(rng.choice(n_terms, size=k, p=zipf_p)).astype(np.int32))
            else:
                ann1.append(np.empty(0, np.int32))
        ann2 = [None] * n2
        for u in range(n1):
            t = ann1[u]
            if len(t) and rng.random() < 0.8:
                keep = t[rng.random(len(t)) < 0.8]
                extra = rng.choice(n_terms, size=int(rng.integers(0, 4)), p=zipf_p)
                t2 = np.unique(np.concatenate([keep, extra])).astype(np.int32)
                ann2[orth[u]] = t2 if len(t2) else t[:1].copy()
        for v in range(n2):
            if ann2[v] is None:
                if rng.random() < annotated_frac:
                    k = max(1, int(rng.poisson(terms_per_node)))
                    ann2[v] = np.unique(rng.choice(n_terms, size=k, p=zipf_p)).astype(np.int32)
                else:
                    ann2[v] = np.empty(0, np.int32)
        pair.ann1, pair.ann2 = ann1, ann2
    return pair


def dense_bits(pair):
    """the dense n1 x n2 matrix loadBLASTInfo builds (blastinfo.cpp:20-42)."""
    S = np.zeros((pair.n1, pair.n2), np.float64)
    if pair.bit_rows is not None:
        S[pair.bit_rows, pair.bit_cols] = pair.bit_vals
    return S


def goc_counts(pair):
    """(intersection, union) term counts of every annotated pair: the two ints loadGOC
    divides (goc.cpp:44-48).  Unannotated nodes are absent from the reference's maps, so
    their rows/columns stay zero."""
    import scipy.sparse as sp

    def inc(ann, n):
        rows = np.concatenate([np.full(len(t), i, np.int64) for i, t in enumerate(ann)] or [[]])
        cols = np.concatenate([t.astype(np.int64) for t in ann] or [[]])
        nt = int(cols.max()) + 1 if len(cols) else 1
        return rows, cols, nt

    r1, c1, t1 = inc(pair.ann1, pair.n1)
    r2, c2, t2 = inc(pair.ann2, pair.n2)
    T = max(t1, t2)
    A1 = sp.csr_matrix((np.ones(len(r1), np.int32), (r1, c1)), shape=(pair.n1, T))
    A2 = sp.csr_matrix((np.ones(len(r2), np.int32), (r2, c2)), shape=(pair.n2, T))
    inter = np.asarray((A1 @ A2.T).todense(), dtype=np.int32)
    s1 = np.array([len(t) for t in pair.ann1], np.int32)
    s2 = np.array([len(t) for t in pair.ann2], np.int32)
    union = s1[:, None] + s2[None, :] - inter
    both = (s1[:, None] > 0) & (s2[None, :] > 0)
    inter = np.where(both, inter, 0)
    union = np.where(both, union, 0)
    return inter, union


def dense_goc(pair):
    """the dense n1 x n2 Jaccard matrix loadGOC builds: double(inter)/double(union)."""
    inter, union = goc_counts(pair)
    G = np.zeros((pair.n1, pair.n2), np.float64)
    nz = union > 0
    G[nz] = inter[nz].astype(np.float64) / union[nz].astype(np.float64)
    return G


def random_population(pair, P, seed, total=True):
    """P random alignments: aln[p] a permutation of V2; mask all-true when total."""
    rng = np.random.default_rng(seed)
    aln = np.empty((P, pair.n2), np.int32)
    for p in range(P):
        aln[p] = rng.permutation(pair.n2)
    mask = None
    if not total:
        mask = (rng.random((P, pair.n2)) < 0.5).astype(np.uint8)
    return aln, mask


def random_candidates(pair, P, K, seed):
    """[P,K,2] proposals with the supports correctHillClimb uses (localSearch.cpp:21-28):
    x uniform over the real nodes of net1, y uniform over the other n2-1 positions."""
    rng = np.random.default_rng(seed)
    x = rng.integers(0, pair.n1, (P, K))
    y = rng.integers(0, pair.n2 - 1, (P, K))
    y = y + (y >= x)
    return np.stack([x, y], axis=-1).astype(np.int32)


def with_duplicate_lines(pair, n_dup, seed=0):
    """The same pair with n_dup extra edge LINES per network, each a copy of an earlier line --
    half of them reversed ("b a" after "a b"), self-loops included -- inserted at random places
    after their original, so that the parser's node numbering is unchanged.  The reference keeps
    every copy in adjList but not in its `edges` set or adjMatrix (Network.cpp:87-92)."""
    import copy
    rng = np.random.default_rng(seed)
    out = copy.copy(pair)
    out.meta = dict(pair.meta)
    out.meta["E1"] = pair.E1
    out.meta["E2"] = pair.E2
    for tag, n in (("1", pair.n1), ("2", pair.n2)):
        edges = [tuple(int(z) for z in e) for e in getattr(pair, "edges" + tag)]
        src = rng.integers(0, len(edges), n_dup)
        # insert from the back so earlier insert positions stay valid; a line copied twice gets two copies
        ins = sorted(((int(rng.integers(k + 1, len(edges) + 1)), k) for k in src), reverse=True)
        orig = list(edges)
        for pos, k in ins:
            u, v = orig[k]
            edges.insert(pos, (v, u) if rng.random() < 0.5 else (u, v))
        e = np.array(edges, np.int32)
        setattr(out, "edges" + tag, e)
        rp, col = adjlist_csr(n, e)
        setattr(out, "rp" + tag, rp)
        setattr(out, "col" + tag, col)
    return out


def write_text(pair, prefix, p1="a", p2="b"):
    """write <prefix>1.net, <prefix>2.net (+ .sim, .annos) in the reference's text formats
    (manual.tex:76-87).  Returns the dict of paths."""
    paths = {}
    for tag, edges, pre in (("1", pair.edges1, p1), ("2", pair.edges2, p2)):
        path = "%s%s.net" % (prefix, tag)
        with open(path, "w") as f:
            for u, v in edges:
                f.write("%s%d\t%s%d\n" % (pre, u, pre, v))
        paths["net" + tag] = path
    if pair.bit_rows is not None:
        path = prefix + ".sim"
        with open(path, "w") as f:
            for r, c, v in zip(pair.bit_rows, pair.bit_cols, pair.bit_vals):
                f.write("%s%d %s%d %r\n" % (p1, r, p2, c, float(v)))
        paths["sim"] = path
    if pair.ann1 is not None:
        for tag, ann, pre in (("1", pair.ann1, p1), ("2", pair.ann2, p2)):
            path = "%s%s.annos" % (prefix, tag)
            with open(path, "w") as f:
                for u, t in enumerate(ann):
                    if len(t):
                        f.write("%s%d %s\n" % (pre, u, " ".join(str(int(z)) for z in t)))
            paths["annos" + tag] = path
    return paths
