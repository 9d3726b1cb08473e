"""GPU (-m gpu): the CUDA engine, called through the C ABI, against the oracle and the golden
vectors.  Integer counts, accept bits and permutations must be bit-exact; fp64 accumulators /
objectives are compared exactly wherever the engine replicates the reference's operation order
(everything incremental) and within 1e-9 relative for the tree-summed from-scratch sums."""
import numpy as np
import pytest

from conftest import dense_from_coo
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
    return request.param

ALL = list(range(10))
RTOL = 1e-9          # north_star: "S3/ICS ratios and summed bitscores agree within 1e-9 relative"


def close(a, b):
    return np.allclose(a, b, rtol=RTOL, atol=0.0, equal_nan=True)


def engine_for_golden(g, names, S=None, G=None):
    return engine.Engine(g["n1"], g["n2"], g["E1"], g["rp1"], g["col1"], g["rp2"], g["col2"], names,
                         bits=S, goc=G)


# ------------------------------------------------------------------ golden: small pair
def test_small_golden_full_fitness(small_golden):
    g = small_golden
    e = engine_for_golden(g, ALL, g["S"], g["G"])
    e.upload(g["aln"], g["mask"])
    for mode in (engine.SEED_REFCOMPAT, engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL):
        r = e.full_fitness(seed_mode=mode)
        assert np.array_equal(r["cons_ref"], g["cons"])
        assert np.array_equal(r["ind"], g["ind"])
        assert (r["cons_true"] >= r["cons_ref"]).all() and (r["cons_true"] - r["cons_ref"] <= 1).all()
        if mode & engine.SUM_SEQUENTIAL:
            assert np.array_equal(r["bits"], g["bits"]) and np.array_equal(r["goc"], g["goc"])
            assert np.array_equal(r["fit"], g["fitness"], equal_nan=True)
        else:
            assert close(r["bits"], g["bits"]) and close(r["goc"], g["goc"])
            assert close(r["fit"], g["fitness"])
        c = e.counters()
        assert np.array_equal(c["cons"], g["cons"]) and np.array_equal(c["ind"], g["ind"])
    a, inv, m = e.download()
    assert np.array_equal(a, g["aln"]) and np.array_equal(m, g["mask"])
    assert all(np.array_equal(inv[p][a[p]], np.arange(int(g["n2"]))) for p in range(len(a)))


def test_small_golden_swap_deltas_bit_exact(small_golden):
    g = small_golden
    e = engine_for_golden(g, ALL, g["S"], g["G"])
    e.upload(g["aln"], g["mask"])
    e.set_counters(g["cons"], g["ind"], g["bits"], g["goc"])
    for i in range(4):
        xy = g["cand_xy_%d" % i][None]
        r = e.swap_deltas(xy, obj=-1, first=i)
        for k in ("cons", "ind", "improved"):
            assert np.array_equal(r[k][0], g["cand_%s_%d" % (k, i)]), (i, k)
        for k in ("bits", "goc", "fit", "snd"):
            assert np.array_equal(r[k][0], g["cand_%s_%d" % (k, i)], equal_nan=True), (i, k)


def test_small_golden_hill_climb_trajectories(small_golden):
    g = small_golden
    e = engine_for_golden(g, ALL, g["S"], g["G"])
    e.upload(g["aln"], g["mask"])
    e.set_counters(g["cons"], g["ind"], g["bits"], g["goc"])
    xy = np.stack([g["hc_xy_%d" % i] for i in range(4)])
    objs = np.array([int(g["hc_obj_%d" % i]) for i in range(4)])
    acc, nacc = e.hill_climb(xy=xy, objs=objs, want_accepted=True)
    a, inv, m = e.download()
    c = e.counters()
    for i in range(4):
        assert np.array_equal(acc[i], g["hc_accepted_%d" % i]), i
        assert nacc[i] == g["hc_accepted_%d" % i].sum()
        assert np.array_equal(a[i], g["hc_final_aln_%d" % i])
        assert np.array_equal(m[i], g["hc_final_mask_%d" % i])
        fc = g["hc_final_counters_%d" % i]
        assert (c["cons"][i], c["ind"][i]) == (fc[0], fc[1])
        assert c["bits"][i] == fc[2] and c["goc"][i] == fc[3]      # incl. H3 undo residue
        assert np.array_equal(inv[i][a[i]], np.arange(int(g["n2"])))


# --------------------------------------------------------------- golden: example pair
@pytest.fixture
def example_engine(example_golden, swap_delta_walk):
    g = example_golden
    e = engine_for_golden(g, g["names"], (g["bit_rows"], g["bit_cols"], g["bit_vals"]))
    return e


def test_example_golden_full_fitness(example_golden, example_engine):
    g, e = example_golden, example_engine
    e.upload(g["aln"])
    r = e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    assert np.array_equal(r["cons_ref"], g["cons"]) and np.array_equal(r["ind"], g["ind"])
    assert np.array_equal(r["bits"], g["bits"])         # multiples of 0.5: any order is exact
    assert np.array_equal(r["fit"], g["fitness"])
    assert r["cons_ref"][0] == 140 and r["ind"][0] == 12177 and r["bits"][0] == 360.0
    ident = np.arange(int(g["n2"]), dtype=np.int32)[None]
    e.upload(ident, first=0)
    r = e.full_fitness(first=0, count=1)
    assert (r["cons_ref"][0], r["ind"][0], r["bits"][0]) == (584, 27789, 685.0)


@pytest.mark.parametrize("obj", [-1, 0, 1])
def test_example_golden_swap_deltas(example_golden, example_engine, obj):
    g, e = example_golden, example_engine
    e.upload(g["aln"])
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    r = e.swap_deltas(g["cand_xy"][None], obj=obj, first=0)
    assert np.array_equal(r["improved"][0], g["cand_improved_obj%d" % obj])
    if obj == -1:
        for k in ("cons", "ind"):
            assert np.array_equal(r[k][0], g["cand_%s_obj-1" % k])
        for k in ("bits", "goc", "fit", "snd"):
            assert np.array_equal(r[k][0], g["cand_%s_obj-1" % k], equal_nan=True), k


def test_example_golden_hill_climb(example_golden, example_engine):
    g, e = example_golden, example_engine
    e.upload(g["aln"])
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    acc, nacc = e.hill_climb(xy=g["hc_xy"][None], first=0, count=1, want_accepted=True)
    assert np.array_equal(np.packbits(acc[0]), g["hc_accepted"])
    a, inv, m = e.download(0, 1)
    assert np.array_equal(a[0], g["hc_final_aln"])
    c = e.counters(0, 1)
    fc = g["hc_final_counters"]
    assert (c["cons"][0], c["ind"][0], c["bits"][0]) == (fc[0], fc[1], fc[2])
    assert np.array_equal(c["fit"][0], g["hc_final_fitness"])
    acc1, _ = e.hill_climb(xy=g["hc1_xy"][None], objs=[1], first=1, count=1, want_accepted=True)
    assert np.array_equal(np.packbits(acc1[0]), g["hc1_accepted"])
    assert np.array_equal(e.download(1, 1)[0][0], g["hc1_final_aln"])
    fc = g["hc1_final_counters"]
    c = e.counters(1, 1)
    assert (c["cons"][0], c["ind"][0], c["bits"][0]) == (fc[0], fc[1], fc[2])


# ---------------------------------------------------- seeded inputs against the oracle
@pytest.fixture(scope="module")
def mid():
    pair = synth.make_pair(n1=600, n2=800, m1=3000, m2=3200, max_deg=80, seed=123,
                           annotations=True, n_terms=300, bit_density=0.005, bit_decimals=True)
    S, G = synth.dense_bits(pair), synth.dense_goc(pair)
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, G)
    return pair, S, G, prob


@pytest.mark.parametrize("total", [True, False])
def test_engine_matches_oracle_on_seeded_pair(mid, total):
    pair, S, G, prob = mid
    names = [engine.S3, engine.BITSCORE_SUM, engine.GOC, engine.ICS, engine.SIZE]
    P, K = 12, 1500
    aln, mask = synth.random_population(pair, P, seed=5, total=total)
    e = engine.Engine.from_pair(pair, names, use_goc=True, goc_dense=G)
    e.upload(aln, mask)
    r = e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    orc = [port.Aln(prob, aln[p], None if mask is None else mask[p]) for p in range(P)]
    for p in range(P):
        c = orc[p].counters()
        assert (r["cons_ref"][p], r["ind"][p]) == (c[0], c[1])
        assert r["cons_true"][p] == orc[p].conserved_true()
        assert r["bits"][p] == c[2] and r["goc"][p] == c[3]
        assert np.array_equal(r["fit"][p], orc[p].compute_fitness(names), equal_nan=True)
    # general candidates: any two positions (dummies included, as mutate/crossover produce)
    rng = np.random.default_rng(9)
    xy = rng.integers(0, pair.n2, (P, K, 2)).astype(np.int32)
    for obj in (-1, 2):
        d = e.swap_deltas(xy, obj=obj)
        for p in range(P):
            o = orc[p].swap_delta_records(xy[p], names, obj=obj, want_snd=True)
            for k in ("cons", "ind", "improved"):
                assert np.array_equal(d[k][p], o[k]), (p, k)
            for k in ("bits", "goc", "fit", "snd"):
                assert np.array_equal(d[k][p], o[k], equal_nan=True), (p, k)
    # hill climb, mixed objectives per individual, proposals from the host
    iters = 6000
    hxy = synth.random_candidates(pair, P, iters, seed=77)
    objs = np.array([-1, 0, 1, 2] * 3)
    acc, nacc = e.hill_climb(xy=hxy, objs=objs, want_accepted=True)
    a, inv, m = e.download()
    c = e.counters()
    for p in range(P):
        oacc = orc[p].hill_climb_replay(hxy[p], names, obj=int(objs[p]), want_accepted=True)
        assert np.array_equal(acc[p], oacc), p
        st = orc[p].state()
        assert np.array_equal(a[p], st[0]) and np.array_equal(inv[p], st[1]) and np.array_equal(m[p], st[2])
        oc = orc[p].counters()
        assert (c["cons"][p], c["ind"][p], c["bits"][p], c["goc"][p]) == oc
        assert np.array_equal(c["fit"][p], orc[p].compute_fitness(names), equal_nan=True)
    assert nacc.sum() == acc.sum() and acc.sum() > 0


def test_device_rng_proposals_and_climb(mid):
    pair, S, G, prob = mid
    names = [engine.S3, engine.BITSCORE_SUM]
    P, iters = 6, 5000
    aln, _ = synth.random_population(pair, P, seed=1)
    e = engine.Engine.from_pair(pair, names)
    e.set_island_offset(40)
    e.upload(aln)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    base = e.counters()
    xy = e.rng_proposals(iters, seed=2024)
    assert np.array_equal(xy, engine.rng_proposals_host(2024, pair.n1, pair.n2, 40, P, iters))
    _, nacc = e.hill_climb(iters=iters, seed=2024)
    a, _, _ = e.download()
    c = e.counters()
    for p in range(P):
        o = port.Aln(prob, aln[p])
        assert o.counters()[:3] == (base["cons"][p], base["ind"][p], base["bits"][p])
        oacc = o.hill_climb_replay(xy[p], names, want_accepted=True)
        assert oacc.sum() == nacc[p]
        assert np.array_equal(a[p], o.state()[0])
        assert (c["cons"][p], c["ind"][p], c["bits"][p]) == o.counters()[:3]
    # swap_deltas_rng: accept counts + arg-max swapNormalizedDelta over the same proposals
    K = 3000
    ni, bs, bxy = e.swap_deltas_rng(K, seed=7, best=True)
    pxy = e.rng_proposals(K, seed=7)
    d = e.swap_deltas(pxy, obj=-1)
    assert np.array_equal(ni, d["improved"].sum(1))
    for p in range(P):
        snd = np.where(np.isnan(d["snd"][p]), -np.inf, d["snd"][p])
        k = int(np.argmax(snd))
        assert bs[p] == snd[k] and tuple(bxy[p]) == tuple(pxy[p, k])


def test_apply_swaps_matches_doSwap_sequence(mid):
    pair, S, G, prob = mid
    names = [engine.EC, engine.BITSCORE_SUM, engine.GOC, engine.SIZE]
    P, n = 5, 300
    aln, mask = synth.random_population(pair, P, seed=8, total=False)
    e = engine.Engine.from_pair(pair, names, use_goc=True, goc_dense=G)
    e.upload(aln, mask)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    rng = np.random.default_rng(3)
    xy = rng.integers(0, pair.n2, (P, n, 2)).astype(np.int32)
    xy[:, ::7, 0] = -1                                   # ragged lists: skipped entries
    e.apply_swaps(xy)
    a, inv, m = e.download()
    c = e.counters()
    for p in range(P):
        o = port.Aln(prob, aln[p], mask[p])
        for x, y in xy[p]:
            if x >= 0:
                o.do_swap(int(x), int(y))
        st = o.state()
        assert np.array_equal(a[p], st[0]) and np.array_equal(m[p], st[2])
        assert (c["cons"][p], c["ind"][p], c["bits"][p], c["goc"][p]) == o.counters()
        assert c["size"][p] == o.aln_size()


def test_wide_id_path_uint32_state():
    # n2 >= 32768 switches the shared-memory state to 32-bit entries
    pair = synth.make_pair(n1=1500, n2=33000, m1=5000, m2=40000, max_deg=100, seed=31, bits=False)
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2)
    names = [engine.S3, engine.ICS]
    aln, _ = synth.random_population(pair, 3, seed=2)
    e = engine.Engine.from_pair(pair, names, use_bits=False)
    e.upload(aln)
    r = e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    xy = synth.random_candidates(pair, 3, 3000, seed=4)
    d = e.swap_deltas(xy, want=("cons", "ind", "improved", "fit"))
    acc, _ = e.hill_climb(xy=xy, want_accepted=True)
    a, _, _ = e.download()
    for p in range(3):
        o = port.Aln(prob, aln[p])
        assert (r["cons_ref"][p], r["ind"][p]) == o.counters()[:2]
        rec = o.swap_delta_records(xy[p], names)
        assert np.array_equal(d["cons"][p], rec["cons"]) and np.array_equal(d["ind"][p], rec["ind"])
        assert np.array_equal(d["improved"][p], rec["improved"])
        assert np.array_equal(d["fit"][p], rec["fit"], equal_nan=True)
        assert np.array_equal(acc[p], o.hill_climb_replay(xy[p], names, want_accepted=True))
        assert np.array_equal(a[p], o.state()[0])


# ------------------------------------------------ full BASELINE size: invariant families
def test_full_size_invariants_yeast_fly():
    """C2 shape (6k/8k nodes, 30k edges, P=200): size-independent properties the reference's own
    suite checks (tests.cpp: *_consistent_after_doSwap, alnInv_consistent_*, fast_ics_match_*)."""
    pair = synth.make_pair(seed=synth.DEFAULT_SEED + 1)
    names = [engine.S3, engine.BITSCORE_SUM]
    P = 200
    aln, _ = synth.random_population(pair, P, seed=11)
    e = engine.Engine.from_pair(pair, names)
    e.upload(aln)
    r0 = e.full_fitness(seed_mode=engine.SEED_TRUE)
    _, nacc = e.hill_climb(iters=20000, seed=99)
    assert (nacc > 0).all()
    inc = e.counters()
    scratch = e.full_fitness(seed_mode=engine.SEED_NONE)
    assert np.array_equal(inc["cons"], scratch["cons_true"])       # incremental == from scratch
    assert np.array_equal(inc["ind"], scratch["ind"])
    assert np.array_equal(inc["bits"], scratch["bits"])            # 0.5-multiples: exact
    assert (inc["fit"] >= r0["fit"]).all()                         # obj=-1 never worsens anything
    a, inv, m = e.download()
    ar = np.arange(pair.n2)
    assert all(np.array_equal(inv[p][a[p]], ar) for p in range(P)) and m.all()
    # a swap evaluated as a delta equals committing it and rescoring from scratch
    xy = synth.random_candidates(pair, P, 1, seed=5)
    d = e.swap_deltas(xy, want=("cons", "ind", "bits"))
    e.apply_swaps(xy)
    s2 = e.full_fitness(seed_mode=engine.SEED_NONE)
    assert np.array_equal(d["cons"][:, 0], s2["cons_true"]) and np.array_equal(d["ind"][:, 0], s2["ind"])
    assert np.array_equal(d["bits"][:, 0], s2["bits"])
    # swapping back restores everything (idempotence of the double swap)
    e.apply_swaps(xy)
    c3 = e.counters()
    assert np.array_equal(c3["cons"], inc["cons"]) and np.array_equal(c3["ind"], inc["ind"])
    assert np.array_equal(e.download()[0], a)


def test_error_behaviour():
    pair = synth.make_pair(n1=60, n2=80, m1=200, m2=260, max_deg=20, seed=4)
    e = engine.Engine.from_pair(pair, [engine.S3])
    bad = np.zeros((1, pair.n2), np.int32)
    with pytest.raises(engine.OnaError, match="permutation"):
        e.upload(bad)
    with pytest.raises(engine.OnaError):
        e.upload(np.full((1, pair.n2), pair.n2, np.int32))
    aln, _ = synth.random_population(pair, 2, seed=1)
    e.upload(aln)
    with pytest.raises(engine.OnaError):
        e.full_fitness(first=1, count=5)
    with pytest.raises(engine.OnaError):
        e.swap_deltas(np.full((2, 3, 2), pair.n2, np.int32))
    with pytest.raises(engine.OnaError):
        e.hill_climb(iters=10, objs=[3, 3])
    col = pair.col1.copy()
    col[3] = pair.n1                                     # column id outside the network
    with pytest.raises(engine.OnaError, match="out of range"):
        engine.Engine(pair.n1, pair.n2, pair.E1, pair.rp1, col, pair.rp2, pair.col2, [engine.S3])
    # empty inputs are no-ops
    e.hill_climb(iters=0)
    e.swap_deltas(np.empty((2, 0, 2), np.int32))
    # device proposal stream: a generator position beyond 624 or an empty range is refused, zero
    # iterations leave the states alone
    st = np.stack([port.mt_state(1), port.mt_state(2)])
    bad_st = st.copy()
    bad_st[1, 624] = 625
    with pytest.raises(engine.OnaError, match="position"):
        e.hill_climb_mt(10, bad_st)
    with pytest.raises(engine.OnaError):
        e.mt_proposals(5, 1, 10, st)
    with pytest.raises(engine.OnaError):
        e.hill_climb_mt(10, st, objs=[0, 7])
    _, nacc, after = e.hill_climb_mt(0, st)
    assert np.array_equal(after, st) and not nacc.any()


def test_forked_queue_and_per_individual_lists(mid):
    """ona_engine_fork: a second queue on the same device data (own slots / stream / lock) gives the
    same results as its parent, concurrently with it; ona_hill_climb_lists == ona_hill_climb on the
    gathered lists."""
    import threading
    pair, S, G, prob = mid
    names = [engine.S3, engine.BITSCORE_SUM, engine.GOC]
    P, iters = 6, 4000
    aln, mask = synth.random_population(pair, P, seed=31, total=False)
    e = engine.Engine.from_pair(pair, names, use_goc=True, goc_dense=G)
    f = e.fork()
    hxy = synth.random_candidates(pair, P, iters, seed=32)
    out = {}

    def climb(tag, q, how):
        q.upload(aln, mask)
        q.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
        if how == "lists":
            nacc = q.hill_climb_lists([hxy[p].copy() for p in range(P)])
        else:
            _, nacc = q.hill_climb(xy=hxy)
        out[tag] = (q.download(), q.counters(), nacc)

    ts = [threading.Thread(target=climb, args=("parent", e, "flat")), threading.Thread(target=climb, args=("fork", f, "lists"))]
    [t.start() for t in ts]
    [t.join() for t in ts]
    (sa, ca, na), (sb, cb, nb) = out["parent"], out["fork"]
    assert all(np.array_equal(x, y) for x, y in zip(sa, sb)) and np.array_equal(na, nb)
    for k in ca:
        assert np.array_equal(ca[k], cb[k], equal_nan=True), k
    for p in range(P):
        o = port.Aln(prob, aln[p], mask[p])
        o.hill_climb_replay(hxy[p], names)
        assert np.array_equal(sb[0][p], o.state()[0])
        assert (cb["cons"][p], cb["ind"][p], cb["bits"][p], cb["goc"][p]) == o.counters()
    f.close()
    # the parent keeps working after its fork is gone (and the other way round)
    e.full_fitness(seed_mode=engine.SEED_NONE)
    g = e.fork()
    e.close()
    g.upload(aln, mask)
    r = g.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    assert r["ind"][0] == port.Aln(prob, aln[0], mask[0]).counters()[1]
    g.close()


def test_resize_keeps_uploaded_slots():
    """growing the population (explicitly or through upload(first=...)) keeps earlier slots:
    permutation, mask, counters and objective rows."""
    pair = synth.make_pair(n1=120, n2=150, m1=500, m2=560, max_deg=30, seed=14, bit_density=0.02)
    e = engine.Engine.from_pair(pair, [engine.S3, engine.BITSCORE_SUM])
    aln, mask = synth.random_population(pair, 7, seed=2, total=False)
    e.upload(aln[:3], mask[:3])
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    c0, s0 = e.counters(), e.download()
    e.upload(aln[3:], mask[3:], first=3)                 # grows 3 -> 7
    assert e.P == 7
    e.full_fitness(first=3, count=4, seed_mode=engine.SEED_REFCOMPAT)
    c1, s1 = e.counters(), e.download()
    for k in c0:
        assert np.array_equal(c1[k][:3], c0[k], equal_nan=True), k
    assert all(np.array_equal(a[:3], b) for a, b in zip(s1, s0))
    assert np.array_equal(s1[0][3:], aln[3:]) and np.array_equal(s1[2][3:], mask[3:])
    e.resize(5)                                          # shrink keeps the first 5
    c2 = e.counters()
    for k in c1:
        assert np.array_equal(c2[k], c1[k][:5], equal_nan=True), k
    # the whole population scored at once equals the piecewise result
    r = e.full_fitness(seed_mode=engine.SEED_NONE)
    assert np.array_equal(r["cons_ref"], c2["cons"]) and np.array_equal(r["ind"], c2["ind"])


@pytest.mark.parametrize("total", [True, False])
def test_duplicate_edge_lines_match_oracle(total):
    """Edge lists with "a b" + "b a" and repeated lines (self-loops too) are legal reference input:
    adjList keeps every copy, so conservedCount / inducedCount / fastICSDenominator weigh such an
    edge by its multiplicity while edges.size(), adjMatrix and the set-based ics() see it once
    (Network.cpp:87-92).  The oracle is pinned on exactly this input against the
    compiled reference (tests/test_oracle_vs_reference.py, "duplicate-lines")."""
    base = synth.make_pair(n1=300, n2=380, m1=1500, m2=1700, max_deg=60, seed=15, loop_frac=0.06,
                           annotations=True, n_terms=120, bit_density=0.02)
    pair = synth.with_duplicate_lines(base, 200, seed=4)
    assert len(pair.col1) > len(base.col1) and pair.E1 == base.E1
    S, G = synth.dense_bits(pair), synth.dense_goc(pair)
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, G)
    names = [engine.S3, engine.ICS, engine.EC, engine.BITSCORE_SUM, engine.GOC, engine.SIZE]
    P, K = 10, 3000
    aln, mask = synth.random_population(pair, P, seed=6, total=total)
    e = engine.Engine.from_pair(pair, names, use_goc=True, goc_dense=G)
    e.upload(aln, mask)
    r = e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    orc = [port.Aln(prob, aln[p], None if mask is None else mask[p]) for p in range(P)]
    differ = 0
    for p in range(P):
        c = orc[p].counters()
        assert (r["cons_ref"][p], r["ind"][p]) == (c[0], c[1])
        assert r["cons_true"][p] == orc[p].conserved_true() and r["ics"][p] == orc[p].ics()
        assert r["ind"][p] == orc[p].induced_true()
        assert np.array_equal(r["fit"][p], orc[p].compute_fitness(names), equal_nan=True)
        differ += int(r["ics"][p] * r["ind"][p] != r["cons_true"][p])
    assert differ > 0                                    # the multiplicities do matter on this input
    # candidates biased towards the duplicated edges' endpoints (self-loops and the a_x - a_y edge)
    rng = np.random.default_rng(11)
    xy = rng.integers(0, pair.n2, (P, K, 2)).astype(np.int32)
    d = e.swap_deltas(xy, obj=-1)
    for p in range(P):
        o = orc[p].swap_delta_records(xy[p], names, obj=-1, want_snd=True)
        for k in ("cons", "ind", "improved"):
            assert np.array_equal(d[k][p], o[k]), (p, k)
        for k in ("bits", "goc", "fit", "snd"):
            assert np.array_equal(d[k][p], o[k], equal_nan=True), (p, k)
    iters = 8000
    hxy = synth.random_candidates(pair, P, iters, seed=78)
    objs = np.array([-1, 0, 1, 3, 4] * 2)
    acc, nacc = e.hill_climb(xy=hxy, objs=objs, want_accepted=True)
    a, inv, m = e.download()
    c = e.counters()
    for p in range(P):
        oacc = orc[p].hill_climb_replay(hxy[p], names, obj=int(objs[p]), want_accepted=True)
        assert np.array_equal(acc[p], oacc), p
        st = orc[p].state()
        assert np.array_equal(a[p], st[0]) and np.array_equal(inv[p], st[1]) and np.array_equal(m[p], st[2])
        assert (c["cons"][p], c["ind"][p], c["bits"][p], c["goc"][p]) == orc[p].counters()
    assert acc.sum() > 0
    # variation operators on the same graphs
    ops = np.zeros((P, 60, 4), np.int32)
    ops[..., 0] = rng.integers(0, 4, (P, 60))
    ops[..., 1] = rng.integers(0, pair.n2, (P, 60))
    ops[..., 2] = rng.integers(0, pair.n2, (P, 60))
    ops[..., 3] = rng.integers(0, 8, (P, 60))
    e.apply_edits(ops)
    c = e.counters()
    for p in range(P):
        orc[p].apply_edits(ops[p])
        assert (c["cons"][p], c["ind"][p], c["bits"][p], c["goc"][p]) == orc[p].counters(), p
    # the throughput climb keeps the counters equal to a (multiplicity-weighted) recount
    e.full_fitness(seed_mode=engine.SEED_TRUE)
    e.hill_climb_batched(iters=4096, seed=5)
    c = e.counters()
    r = e.full_fitness(seed_mode=engine.SEED_NONE)
    assert np.array_equal(c["cons"], r["cons_true"]) and np.array_equal(c["ind"], r["ind"])


def test_steepest_step_matches_exhaustive_oracle_scan():
    """ona_steepest_step == the arg-max of swapNormalizedDelta over the whole n1 x n2 grid with
    the reference's tie rule (localSearch.cpp:130-188)."""
    pair = synth.make_pair(n1=60, n2=80, m1=220, m2=260, max_deg=20, seed=6, bit_density=0.05)
    S = synth.dense_bits(pair)
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, None)
    names = [engine.S3, engine.BITSCORE_SUM]
    aln, _ = synth.random_population(pair, 2, seed=3)
    e = engine.Engine.from_pair(pair, names)
    e.upload(aln)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    for p in range(2):
        o = port.Aln(prob, aln[p])
        o.compute_fitness(names)
        grid = np.stack(np.meshgrid(np.arange(pair.n1), np.arange(pair.n2), indexing="ij"), -1).reshape(-1, 2)
        rec = o.swap_delta_records(grid, names, want_snd=True)
        best, bx, by = -1.0, -1, -1
        for (x, y), v in zip(grid, rec["snd"]):
            if v > best:
                best, bx, by = v, int(x), int(y)
        assert e.steepest_step(p) == (best, bx, by)


def test_human_scale_pair_matches_oracle():
    """C4 shape (n1 = n2 = 20 000, 200 k edges each, --ics --blastsum): no dummy positions, so the
    induced count is constant and ICS follows EC; a sample of the population against the oracle."""
    pair = synth.make_pair(n1=20000, n2=20000, m1=200000, m2=200000, seed=synth.DEFAULT_SEED + 4,
                           bit_density=2e-4)
    S = synth.dense_bits(pair)                       # untouched pages stay virtual
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, None)
    names = [engine.ICS, engine.BITSCORE_SUM]
    P, K, iters = 6, 3000, 3000
    aln, _ = synth.random_population(pair, P, seed=21)
    e = engine.Engine.from_pair(pair, names)
    e.upload(aln)
    r = e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    xy = synth.random_candidates(pair, P, K, seed=22)
    d = e.swap_deltas(xy, want=("cons", "ind", "bits", "fit", "improved"))
    hxy = synth.random_candidates(pair, P, iters, seed=23)
    acc, _ = e.hill_climb(xy=hxy, want_accepted=True)
    a, _, _ = e.download()
    c = e.counters()
    for p in range(P):
        o = port.Aln(prob, aln[p])
        assert (r["cons_ref"][p], r["ind"][p], r["bits"][p]) == o.counters()[:3]
        assert r["cons_true"][p] == o.conserved_true()
        rec = o.swap_delta_records(xy[p], names)
        for k in ("cons", "ind", "improved"):
            assert np.array_equal(d[k][p], rec[k]), k
        assert np.array_equal(d["bits"][p], rec["bits"]) and np.array_equal(d["fit"][p], rec["fit"], equal_nan=True)
        assert (d["ind"][p] == r["ind"][p]).all()            # n1 == n2: I never changes
        assert np.array_equal(acc[p], o.hill_climb_replay(hxy[p], names, want_accepted=True))
        assert np.array_equal(a[p], o.state()[0])
        assert (c["cons"][p], c["ind"][p], c["bits"][p]) == o.counters()[:3]


def test_goc_objective_yeast_fly_shape():
    """C3 shape: the yeast-fly pair with GO annotations (dense 6000 x 8000 Jaccard table on the
    device) and the GOC objective; sampled individuals against the oracle, fp64 accumulators
    bit for bit (incremental updates in the reference's order)."""
    pair = synth.make_pair(seed=synth.DEFAULT_SEED + 1, annotations=True)
    G = synth.dense_goc(pair)
    S = synth.dense_bits(pair)
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, G)
    names = [engine.S3, engine.BITSCORE_SUM, engine.GOC]
    P, K, iters = 4, 4000, 8000
    aln, _ = synth.random_population(pair, P, seed=31)
    e = engine.Engine.from_pair(pair, names, use_goc=True, goc_dense=G)
    e.upload(aln)
    r = e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    xy = synth.random_candidates(pair, P, K, seed=32)
    d = e.swap_deltas(xy)
    hxy = synth.random_candidates(pair, P, iters, seed=33)
    acc, _ = e.hill_climb(xy=hxy, want_accepted=True)
    a, _, _ = e.download()
    c = e.counters()
    for p in range(P):
        o = port.Aln(prob, aln[p])
        assert (r["cons_ref"][p], r["ind"][p], r["bits"][p], r["goc"][p]) == o.counters()
        rec = o.swap_delta_records(xy[p], names, want_snd=True)
        for k in ("cons", "ind", "improved"):
            assert np.array_equal(d[k][p], rec[k]), k
        for k in ("bits", "goc", "fit", "snd"):
            assert np.array_equal(d[k][p], rec[k], equal_nan=True), k
        assert np.array_equal(acc[p], o.hill_climb_replay(hxy[p], names, want_accepted=True))
        assert np.array_equal(a[p], o.state()[0])
        assert (c["cons"][p], c["ind"][p], c["bits"][p], c["goc"][p]) == o.counters()


def _batched_round_reference(pair, o, xy, names, obj):
    """Host restatement of one round of ona_hill_climb_batched (selection rule in
    include/ona_b200.h / csrc): decisions against the base state come from the oracle."""
    n1, n2 = pair.n1, pair.n2
    aln, _, mask = o.state()
    rec = o.swap_delta_records(xy, names, obj=obj)
    acc = [bool(rec["improved"][j]) and xy[j][0] != xy[j][1] for j in range(len(xy))]
    BIG = 1 << 30
    wP, rP, wV, rV = (np.full(n2, BIG) for _ in range(4))
    sets = []
    for j, (x, y) in enumerate(xy):
        if not acc[j]:
            sets.append(None)
            continue
        ax, ay = int(aln[x]), int(aln[y])
        xr, yr, mx, my = x < n1, y < n1, bool(mask[x]), bool(mask[y])
        RP = []
        if xr:
            RP += list(pair.col1[pair.rp1[x]:pair.rp1[x + 1]])
        if yr:
            RP += list(pair.col1[pair.rp1[y]:pair.rp1[y + 1]])
        RV = []
        if (xr and mx) != (yr and mx) or (yr and my) != (xr and my):
            RV = list(pair.col2[pair.rp2[ax]:pair.rp2[ax + 1]]) + list(pair.col2[pair.rp2[ay]:pair.rp2[ay + 1]])
        sets.append((x, y, ax, ay, RP, RV))
        for n in (x, y):
            wP[n] = min(wP[n], j)
        for n in (ax, ay):
            wV[n] = min(wV[n], j)
        for n in RP:
            rP[n] = min(rP[n], j)
        for n in RV:
            rV[n] = min(rV[n], j)
    applied = 0
    for j, st in enumerate(sets):
        if st is None:
            continue
        x, y, ax, ay, RP, RV = st
        ok = all(wP[n] == j and rP[n] >= j for n in (x, y)) and all(wV[n] == j and rV[n] >= j for n in (ax, ay))
        ok = ok and all(wP[n] >= j for n in RP) and all(wV[n] >= j for n in RV)
        if ok:
            o.do_swap(int(x), int(y))
            applied += 1
    o.compute_fitness(names)
    return applied


@pytest.mark.parametrize("total", [True, False])
def test_batched_hill_climb_rule_and_exact_counters(mid, total):
    pair, S, G, prob = mid
    names = [engine.S3, engine.BITSCORE_SUM, engine.GOC]
    P, iters = 4, 2500                       # rounds of 1024, 1024, 452
    aln, mask = synth.random_population(pair, P, seed=15, total=total)
    e = engine.Engine.from_pair(pair, names, use_goc=True, goc_dense=G)
    e.upload(aln, mask)
    e.full_fitness(seed_mode=engine.SEED_TRUE | engine.SUM_SEQUENTIAL)
    xy = synth.random_candidates(pair, P, iters, seed=16)
    objs = np.array([-1, 0, -1, 2])
    napp = e.hill_climb_batched(xy=xy, objs=objs)
    a, inv, m = e.download()
    c = e.counters()
    scratch = e.full_fitness(seed_mode=engine.SEED_NONE | engine.SUM_SEQUENTIAL)
    for p in range(P):
        o = port.Aln(prob, aln[p], None if mask is None else mask[p])
        o.set_counters(int(o.conserved_true()), o.counters()[1], o.counters()[2], o.counters()[3])
        o.compute_fitness(names)
        applied = 0
        for r0 in range(0, iters, 1024):
            applied += _batched_round_reference(pair, o, xy[p, r0:r0 + 1024], names, int(objs[p]))
        st = o.state()
        assert applied == napp[p] and applied > 0
        assert np.array_equal(a[p], st[0]) and np.array_equal(inv[p], st[1]) and np.array_equal(m[p], st[2])
        oc = o.counters()
        assert (c["cons"][p], c["ind"][p]) == (oc[0], oc[1])
        assert close(c["bits"][p], oc[2]) and close(c["goc"][p], oc[3])
        # incremental == from scratch (the applied swaps' deltas are exact and additive)
        assert c["cons"][p] == scratch["cons_true"][p] and c["ind"][p] == scratch["ind"][p]
        assert close(c["bits"][p], scratch["bits"][p]) and close(c["goc"][p], scratch["goc"][p])


def test_goc_table_built_on_device_from_term_sets(mid):
    """ona_score_table's term-set form: the engine builds loadGOC's Jaccard matrix (goc.cpp:16-52)
    on the device; everything downstream must be bit-identical to handing it the dense matrix."""
    pair, S, G, prob = mid
    names = [engine.GOC, engine.S3]
    P, K = 5, 2000
    aln, _ = synth.random_population(pair, P, seed=51)
    e_dense = engine.Engine.from_pair(pair, names, use_goc=True, goc_dense=G)
    e_sets = engine.Engine(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, names,
                           bits=(pair.bit_rows, pair.bit_cols, pair.bit_vals),
                           goc={"ann1": pair.ann1, "ann2": pair.ann2})
    xy = synth.random_candidates(pair, P, K, seed=52)
    res = []
    for e in (e_dense, e_sets):
        e.upload(aln)
        r = e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
        d = e.swap_deltas(xy, want=("goc", "fit", "improved"))
        e.hill_climb(iters=3000, seed=5)
        res.append((r, d, e.counters(), e.download()[0]))
    (r0, d0, c0, a0), (r1, d1, c1, a1) = res
    assert np.array_equal(r0["goc"], r1["goc"]) and np.array_equal(r0["fit"], r1["fit"])
    for k in d0:
        assert np.array_equal(d0[k], d1[k], equal_nan=True), k
    assert np.array_equal(a0, a1) and np.array_equal(c0["goc"], c1["goc"])
    for p in range(P):
        assert r1["goc"][p] == port.Aln(prob, aln[p]).counters()[3]
    # the stand-alone form a replacement loadGOC calls (host/goc_b200.cpp): the dense host matrix itself,
    # with nodes that have an entry but no terms (0/0 = NaN against each other, goc.cpp:47) and absent nodes
    assert np.array_equal(engine.goc_matrix(pair.ann1, pair.ann2), G)
    a1 = [np.array(t) for t in ([1, 5, 9], [], [5], [], [2, 3])]
    a2 = [np.array(t) for t in ([5, 9], [], [], [1, 2, 3, 5, 9, 11])]
    M = engine.goc_matrix(a1, a2, present1=[1, 1, 1, 0, 1], present2=[1, 1, 0, 1])
    want = np.zeros((5, 4))
    for i, (s1, q1) in enumerate(zip(a1, [1, 1, 1, 0, 1])):
        for j, (s2, q2) in enumerate(zip(a2, [1, 1, 0, 1])):
            if q1 and q2:
                inter, union = len(set(s1) & set(s2)), len(set(s1) | set(s2))
                want[i, j] = np.float64(inter) / np.float64(union) if union else np.nan
    assert np.array_equal(M, want, equal_nan=True) and np.isnan(M[1, 1])
    with pytest.raises(engine.OnaError, match="sorted"):
        engine.goc_matrix([np.array([3, 1])], [np.array([1])])


# ------------------------------------------------------------------ edit lists (partial operators)
class _EngineObj:
    def __init__(self, e, aln, mask, ctr):
        self.e = e
        e.upload(aln[None], mask[None])
        e.set_counters([int(ctr[0])], [int(ctr[1])], [float(ctr[2])], [float(ctr[3])])

    def apply_edits(self, ops):
        self.e.apply_edits(np.asarray(ops, np.int32).reshape(1, -1, 4))

    def snapshot(self):
        a, _, m = self.e.download()
        c = self.e.counters()
        return a[0], m[0], (int(c["cons"][0]), int(c["ind"][0]), float(c["bits"][0]), float(c["goc"][0]))


@pytest.mark.parametrize("c", [0, 1, 2])
def test_edit_lists_golden(edits_golden, small_golden, c):
    """ona_apply_edits: mutate's mask flips, UPMX with mask AND/OR, onBit -- permutation, mask,
    integer counters and fp64 accumulators exactly as the reference left them."""
    from test_oracle_golden import replay_edits_golden
    g = small_golden
    e = engine_for_golden(g, ALL, g["S"], g["G"])
    replay_edits_golden(edits_golden, lambda aln, mask, ctr: _EngineObj(e, aln, mask, ctr), c)


def test_edit_lists_random_against_oracle():
    """random mixed edit lists (self-loops, dummy positions, i == j UPMX steps, ragged rows) on a
    batch of partial individuals, against orc_apply_edits; objectives refreshed from the result."""
    pair = synth.make_pair(n1=300, n2=380, m1=1500, m2=1700, max_deg=60, seed=5, loop_frac=0.08,
                           annotations=True, n_terms=120, bit_density=0.02)
    S, G = synth.dense_bits(pair), synth.dense_goc(pair)
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, G)
    e = engine.Engine.from_pair(pair, ALL, use_goc=True, goc_dense=G)
    rng = np.random.default_rng(3)
    P, nops = 12, 90
    aln, _ = synth.random_population(pair, P, seed=9)
    mask = (rng.random((P, pair.n2)) < 0.8).astype(np.uint8)
    e.upload(aln, mask)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    c0 = e.counters()
    ops = np.zeros((P, nops, 4), np.int32)
    ops[..., 0] = rng.integers(-1, 4, (P, nops))
    ops[..., 1] = rng.integers(0, pair.n2, (P, nops))
    ops[..., 2] = rng.integers(0, pair.n2, (P, nops))
    ops[..., 3] = rng.integers(0, 8, (P, nops))
    same = rng.random((P, nops)) < 0.1           # UPMX step whose wanted image is already there
    ops[..., 2] = np.where(same & (ops[..., 0] == 3), aln[np.arange(P)[:, None], ops[..., 1]], ops[..., 2])
    e.apply_edits(ops)
    a, inv, m = e.download()
    c1 = e.counters()
    for p in range(P):
        o = port.Aln(prob, aln[p], mask[p])
        o.set_counters(int(c0["cons"][p]), int(c0["ind"][p]), float(c0["bits"][p]), float(c0["goc"][p]))
        o.apply_edits(ops[p])
        oa, oi, om = o.state()
        assert np.array_equal(a[p], oa) and np.array_equal(inv[p], oi) and np.array_equal(m[p], om), p
        assert (int(c1["cons"][p]), int(c1["ind"][p]), float(c1["bits"][p]), float(c1["goc"][p])) == o.counters(), p
        assert np.array_equal(c1["fit"][p], o.compute_fitness(ALL), equal_nan=True), p
    with pytest.raises(engine.OnaError):
        e.apply_edits(np.array([[[7, 0, 0, 0]]], np.int32))
    with pytest.raises(engine.OnaError):
        e.apply_edits(np.array([[[0, 0, pair.n2, 0]]], np.int32))


# ------------------------------------------------------------------ every hill-climb instantiation
@pytest.mark.parametrize("group", ["4", "45", "8", "16"])      # 45: 4 lanes, register budget for five CTAs per SM
@pytest.mark.parametrize("total", [True, False])
def test_hill_climb_every_group_size(monkeypatch, group, total):
    """k_hill_climb<T, G> for G = 4 (what populations above 3 x SMs get), 8 and 16, forced through
    ONA_GROUP, on a small dense pair where almost every commit touches later proposals of its round
    (in-place corrections, shared positions, dummies, self-loops, masks): accept bits, permutation,
    counters and accumulators equal the oracle's sequential loop."""
    monkeypatch.setenv("ONA_GROUP", group)
    pair = synth.make_pair(n1=70, n2=90, m1=700, m2=900, max_deg=40, seed=31, loop_frac=0.1,
                           annotations=True, n_terms=40, bit_density=0.2, bit_decimals=True)
    S, G = synth.dense_bits(pair), synth.dense_goc(pair)
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, G)
    names = [engine.S3, engine.BITSCORE_SUM, engine.GOC, engine.ICS]
    P, iters = 6, 3000
    aln, mask = synth.random_population(pair, P, seed=32, total=total)
    e = engine.Engine.from_pair(pair, names, use_goc=True, goc_dense=G)
    e.upload(aln, mask)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    rng = np.random.default_rng(33)
    xy = rng.integers(0, pair.n2, (P, iters, 2)).astype(np.int32)     # any positions, x == y included
    objs = np.array([-1, 0, 1, 2, 3, -1])
    acc, nacc = e.hill_climb(xy=xy, objs=objs, want_accepted=True)
    a, inv, m = e.download()
    c = e.counters()
    for p in range(P):
        o = port.Aln(prob, aln[p], None if mask is None else mask[p])
        want = o.hill_climb_replay(xy[p], names, obj=int(objs[p]), want_accepted=True)
        assert np.array_equal(acc[p], want), (group, p)
        st = o.state()
        assert np.array_equal(a[p], st[0]) and np.array_equal(inv[p], st[1]) and np.array_equal(m[p], st[2])
        assert (int(c["cons"][p]), int(c["ind"][p]), float(c["bits"][p]), float(c["goc"][p])) == o.counters()


# ------------------------------------------------- the reference's own proposal stream on the device
def test_mt_proposal_stream_on_device(mid, mt_golden, swap_delta_walk):
    """ona_mt_proposals / ona_hill_climb_mt: std::mt19937 + libstdc++'s uniform_int_distribution
    (localSearch.cpp:21-28) generated on the device -- proposals and the generator state left behind,
    against golden vectors recorded from the real ones and the oracle's restatement; the climb they
    drive equals the climb on the host-drawn list."""
    if swap_delta_walk != "classic-walk":
        pytest.skip("independent of the swap-delta walk")
    pair, S, G, prob = mid
    names = [engine.S3, engine.BITSCORE_SUM]
    e = engine.Engine.from_pair(pair, names)
    g = mt_golden
    for i, (seed, n1, n2, iters) in enumerate(g["cases"].tolist()):
        xy, after = e.mt_proposals(n1, n2, iters, g["c%d_state" % i][None])
        assert np.array_equal(xy[0], g["c%d_xy" % i]), i
        assert np.array_equal(after[0], g["c%d_after" % i]), i
    # many generators at once, odd lengths (an x left pending across a state refill), states that start
    # in the middle of a block, ranges where every third draw is rejected and drawn again
    rng = np.random.default_rng(3)
    for n1, n2, iters in [(600, 800, 1), (600, 800, 311), (600, 800, 312), (600, 800, 313), (600, 800, 2501),
                          (1, 2, 700), (1, 3, 700), (1_500_000_000, 2_000_000_000, 1999),
                          (2_147_483_647, 2_147_483_647, 1000), (1_073_741_825, 1_073_741_826, 1777)]:
        states = np.stack([port.mt_state(int(s)) for s in rng.integers(0, 2**32, 9)])
        for p in range(9):          # advance each by a different amount first
            states[p] = port.mt_xy_stream(states[p], 77, 99, int(rng.integers(0, 700)))[1]
        xy, after = e.mt_proposals(n1, n2, iters, states)
        for p in range(9):
            xo, so = port.mt_xy_stream(states[p], n1, n2, iters)
            assert np.array_equal(xy[p], xo), (n1, n2, iters, p)
            assert np.array_equal(after[p], so), (n1, n2, iters, p)
        xy2, after2 = e.mt_proposals(n1, n2, 5, after)      # and the states are resumable
        for p in range(9):
            assert np.array_equal(xy2[p], port.mt_xy_stream(after[p], n1, n2, 5)[0])
    # a long stream (481 state refills): only the last proposals and the end state are compared in full
    states = np.stack([port.mt_state(s) for s in (5, 6, 7)])
    xy, after = e.mt_proposals(6000, 8000, 150001, states)
    for p in range(3):
        xo, so = port.mt_xy_stream(states[p], 6000, 8000, 150001)
        assert np.array_equal(xy[p], xo) and np.array_equal(after[p], so)
    # the climb
    P, iters = 5, 7001
    aln, _ = synth.random_population(pair, P, seed=11)
    states = np.stack([port.mt_state(1000 + p) for p in range(P)])
    e.upload(aln)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT | engine.SUM_SEQUENTIAL)
    acc, nacc, after = e.hill_climb_mt(iters, states, objs=[-1, 0, 1, -1, 0], want_accepted=True)
    a, _, _ = e.download()
    c = e.counters()
    for p in range(P):
        xy, so = port.mt_xy_stream(states[p], pair.n1, pair.n2, iters)
        assert np.array_equal(after[p], so)
        o = port.Aln(prob, aln[p])
        o.compute_fitness(names)
        oacc = o.hill_climb_replay(xy, names, obj=[-1, 0, 1, -1, 0][p], want_accepted=True)
        assert np.array_equal(acc[p], oacc) and nacc[p] == oacc.sum()
        assert np.array_equal(a[p], o.state()[0])
        assert (c["cons"][p], c["ind"][p], c["bits"][p]) == o.counters()[:3]
