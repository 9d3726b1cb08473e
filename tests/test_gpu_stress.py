"""GPU: randomized differential test against the oracle on many tiny, nasty inputs -- dense
graphs (every swap hits self-loops / the x-y edge / the a_x-a_y edge), partial masks, dummy and
equal positions, every objective, ragged sizes around the 32/1024 boundaries."""
import os

import numpy as np
import pytest

from optnetaligncpp_b200 import engine, synth
from oracle import port

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["classic-walk", "fast-walk"])
def swap_delta_walk(request, monkeypatch):
    """every test runs with both walks of the swap-delta kernel: the classic one (both neighbour
    lists of x and y) and the fast one on precomputed node counts (normally chosen from 4096
    candidates per individual on; ONA_FAST_MIN_K is read when an engine is created)"""
    monkeypatch.setenv("ONA_FAST_MIN_K", "-1" if request.param == "classic-walk" else "0")
    # ... and with both from-scratch kernels: edge slices over several CTAs (small batches) and one
    # CTA per network (from 96 individuals on), which also leaves the node counts behind
    monkeypatch.setenv("ONA_FF_SPLIT_MIN", "1000000" if request.param == "classic-walk" else "1")

ALL = list(range(10))


def tiny_pair(rng, n1, n2, p_edge, loops):
    def edges(n):
        e = [(u, v) for u in range(n) for v in range(u + 1, n) if rng.random() < p_edge]
        e += [(u, u) for u in range(n) if rng.random() < loops]
        deg = np.zeros(n, int)
        for u, v in e:
            deg[u] += 1
            deg[v] += 1
        for u in np.nonzero(deg == 0)[0]:            # every node must appear in a line
            e.append((int(u), int(u)))
        e = np.array(e, np.int64)
        return e[rng.permutation(len(e))]
    e1, _ = synth.canonicalize(edges(n1))
    e2, _ = synth.canonicalize(edges(n2))
    pair = synth.NetPair(n1=n1, n2=n2, edges1=e1, edges2=e2)
    pair.rp1, pair.col1 = synth.adjlist_csr(n1, e1)
    pair.rp2, pair.col2 = synth.adjlist_csr(n2, e2)
    S = np.where(rng.random((n1, n2)) < 0.3, np.round(rng.random((n1, n2)) * 50, 1), 0.0)
    G = np.where(rng.random((n1, n2)) < 0.6, rng.integers(1, 5, (n1, n2)) / rng.integers(5, 9, (n1, n2)), 0.0)
    return pair, S, G


@pytest.mark.parametrize("seed", range(int(os.environ.get("ONA_STRESS_SEEDS", "6"))))
def test_random_tiny_inputs_match_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    n1 = int(rng.integers(2, 40))
    n2 = n1 + int(rng.integers(0, 12))
    pair, S, G = tiny_pair(rng, n1, n2, p_edge=float(rng.choice([0.08, 0.3, 0.8])), loops=0.3)
    prob = port.Problem(n1, n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, G)
    names = list(rng.permutation(ALL)[: int(rng.integers(1, 11))])
    P = int(rng.integers(1, 5))
    K = int(rng.choice([1, 31, 33, 1023, 1025, 2500]))
    aln = np.stack([rng.permutation(n2) for _ in range(P)]).astype(np.int32)
    mask = None if seed % 2 == 0 else (rng.random((P, n2)) < 0.7).astype(np.uint8)
    e = engine.Engine(n1, n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, names, bits=S, goc=G)
    e.upload(aln, mask)
    r = e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    xy = rng.integers(0, n2, (P, K, 2)).astype(np.int32)            # any two positions, x == y included
    objs = rng.integers(-1, len(names), P)
    orc = [port.Aln(prob, aln[p], None if mask is None else mask[p]) for p in range(P)]
    for p in range(P):
        c = orc[p].counters()
        assert (r["cons_ref"][p], r["ind"][p], r["bits"][p], r["goc"][p]) == c
        assert np.array_equal(r["fit"][p], orc[p].compute_fitness(names), equal_nan=True)
        d = e.swap_deltas(xy[p:p + 1], obj=int(objs[p]), first=p)
        o = orc[p].swap_delta_records(xy[p], names, obj=int(objs[p]), want_snd=True)
        for k in ("cons", "ind", "improved"):
            assert np.array_equal(d[k][0], o[k]), (p, k)
        for k in ("bits", "goc", "fit", "snd"):
            assert np.array_equal(d[k][0], o[k], equal_nan=True), (p, k)
    if n2 >= 3:
        best = [e.steepest_step(p) for p in range(P)]
        grid = np.stack(np.meshgrid(np.arange(n1), np.arange(n2), indexing="ij"), -1).reshape(-1, 2)
        for p in range(P):
            o = orc[p].clone()
            o.compute_fitness(names)
            rec = o.swap_delta_records(grid, names, want_snd=True)
            b, bx, by = -1.0, -1, -1
            for (x, y), v in zip(grid, rec["snd"]):
                if v > b:
                    b, bx, by = v, int(x), int(y)
            assert best[p] == (b, bx, by) or (np.isinf(b) and np.isinf(best[p][0]) and best[p][1:] == (bx, by))
    # exact hill climb on general proposals (dummy x too), then unconditional swaps
    iters = int(rng.choice([1, 40, 700]))
    hxy = rng.integers(0, n2, (P, iters, 2)).astype(np.int32)
    acc, nacc = e.hill_climb(xy=hxy, objs=objs, want_accepted=True)
    e.apply_swaps(hxy[:, : min(iters, 50)])
    a, inv, m = e.download()
    c = e.counters()
    for p in range(P):
        oacc = orc[p].hill_climb_replay(hxy[p], names, obj=int(objs[p]), want_accepted=True)
        assert np.array_equal(acc[p], oacc), p
        for x, y in hxy[p, : min(iters, 50)]:
            orc[p].do_swap(int(x), int(y))
        st = orc[p].state()
        assert np.array_equal(a[p], st[0]) and np.array_equal(inv[p], st[1]) and np.array_equal(m[p], st[2])
        assert (c["cons"][p], c["ind"][p], c["bits"][p], c["goc"][p]) == orc[p].counters()
