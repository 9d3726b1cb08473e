"""CPU: the C restatement against the live compiled reference (oracle/_ref), plus the
reference test-suite's invariant families (tests.cpp) re-run on synthetic
inputs.  Skipped where the prebuilt _ref library did not travel."""
import os
import tempfile

import numpy as np
import pytest

from optnetaligncpp_b200 import synth
from oracle import port, ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref/libona_ref.so not built")

NAMES = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]


@pytest.fixture(scope="module", params=["plain", "duplicate-lines"])
def world(request):
    pair = synth.make_pair(n1=250, n2=330, m1=1200, m2=1300, max_deg=50, seed=77,
                           annotations=True, n_terms=150, bit_density=0.01)
    if request.param == "duplicate-lines":
        # "a b" + "b a" and repeated lines (self-loops too): every copy stays in adjList, the
        # `edges` set / adjMatrix / ics() see the edge once (Network.cpp:87-92)
        pair = synth.with_duplicate_lines(pair, 90, seed=3)
    d = tempfile.mkdtemp()
    paths = synth.write_text(pair, os.path.join(d, "w"))
    n1, n2 = ref.Net(paths["net1"]), ref.Net(paths["net2"])
    B = ref.Dict.load_blast(n1, n2, paths["sim"])
    G = ref.Dict.load_goc(n1, n2, paths["annos1"], paths["annos2"])
    S, Gd = synth.dense_bits(pair), synth.dense_goc(pair)
    prob = port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, Gd)
    return dict(pair=pair, n1=n1, n2=n2, B=B, G=G, S=S, Gd=Gd, prob=prob)


def test_loader_numbering_and_adjacency(world):
    # adjacency_list_correct / adjList_adjMatrix_consistent (tests.cpp:443-508)
    pair, n1, n2 = world["pair"], world["n1"], world["n2"]
    assert (n1.n, n1.E, n2.n, n2.E) == (pair.n1, pair.E1, pair.n2, pair.E2)
    for net, rp, col in ((n1, pair.rp1, pair.col1), (n2, pair.rp2, pair.col2)):
        r, c = net.csr()
        assert np.array_equal(r, rp) and np.array_equal(c, col)
    rng = np.random.default_rng(0)
    dense = np.zeros((pair.n2, pair.n2), bool)
    for u in range(pair.n2):
        dense[u, pair.col2[pair.rp2[u]:pair.rp2[u + 1]]] = True
    for u, v in rng.integers(0, pair.n2, (2000, 2)):
        assert n2.adj_bit(int(u), int(v)) == int(dense[u, v])
    assert np.array_equal(world["B"].dense(), world["S"])
    assert np.array_equal(world["G"].dense(), world["Gd"])      # double(i)/double(u), goc.cpp:47


@pytest.mark.parametrize("total", [True, False])
def test_from_scratch_and_incremental_match_reference(world, total):
    w = world
    for seed in range(6):
        a = ref.Aln(w["n1"], w["n2"], w["B"], w["G"])
        a.shuf(seed, total=total)
        al, inv, m = a.state()
        o = port.Aln(w["prob"], al, m)
        assert a.counters() == o.counters()
        assert a.ics() == o.ics() and a.fast_ics_denominator() == o.induced_true()
        assert a.sum_blast() == o.sum_blast() and a.sum_goc() == o.sum_goc()
        assert a.aln_size() == o.aln_size()
        assert np.array_equal(a.compute_fitness(NAMES), o.compute_fitness(NAMES), equal_nan=True)
        rng = np.random.default_rng(seed)
        for _ in range(200):      # doSwap with arbitrary positions (mutate path, Alignment.cpp:464)
            x, y = (int(v) for v in rng.integers(0, w["pair"].n2, 2))
            if x == y:
                continue
            assert a.conserved_count(x, int(al[x]), 1, y) == o.conserved_count(x, int(al[x]), 1, y)
            a.do_swap(x, y)
            o.do_swap(x, y)
            al = a.state()[0]
        assert a.counters() == o.counters()
        assert all(np.array_equal(p, q) for p, q in zip(a.state(), o.state()))
        # do_swap_consistent_ics / currInducedCount_consistent_after_doSwap (tests.cpp:670-697,993+)
        assert o.induced_true() == o.counters()[1]
        assert a.fast_ics_denominator() == a.counters()[1]
        assert abs(o.sum_blast() - o.counters()[2]) < 1e-9 * max(1.0, abs(o.counters()[2]))
        assert abs(o.sum_goc() - o.counters()[3]) < 1e-9 * max(1.0, abs(o.counters()[3]))


@pytest.mark.parametrize("obj", [-1, 2, 5, 6])
def test_hill_climb_replay_is_correctHillClimb(world, obj):
    w = world
    names = [6, 2, 5]
    o_idx = {-1: -1, 6: 0, 2: 1, 5: 2}[obj]
    a = ref.Aln(w["n1"], w["n2"], w["B"], w["G"])
    a.shuf(3)
    al, _, m = a.state()
    o = port.Aln(w["prob"], al, m)
    real = a.clone()
    real.correct_hill_climb(99, 30000, names, obj=o_idx)
    xy = ref.gen_xy_stream(99, w["pair"].n1, w["pair"].n2, 30000)
    acc_r = a.hill_climb_replay(xy, names, obj=o_idx, want_accepted=True)
    acc_o = o.hill_climb_replay(xy, names, obj=o_idx, want_accepted=True)
    assert np.array_equal(acc_r, acc_o)
    assert real.counters() == a.counters() == o.counters()
    assert all(np.array_equal(p, q) for p, q in zip(real.state(), o.state()))


def test_swap_normalized_delta_and_records(world):
    w = world
    a = ref.Aln(w["n1"], w["n2"], w["B"], w["G"])
    a.shuf(5, total=False)
    al, _, m = a.state()
    o = port.Aln(w["prob"], al, m)
    xy = ref.gen_xy_stream(1, w["pair"].n1, w["pair"].n2, 1500)
    r1 = a.swap_delta_records(xy, NAMES, want_snd=True)
    r2 = o.swap_delta_records(xy, NAMES, want_snd=True)
    for k in r1:
        assert np.array_equal(r1[k], r2[k], equal_nan=True), k


def test_refseed_formula_matches_compiled_reference(world):
    # H1: the truncated seed equals (int)((C/I)*I) with the TRUE counts
    w = world
    hits = 0
    for seed in range(40):
        a = ref.Aln(w["n1"], w["n2"], w["B"], w["G"])
        a.shuf(seed)
        al, _, m = a.state()
        o = port.Aln(w["prob"], al, m)
        C, I = o.conserved_true(), o.induced_true()
        assert a.counters()[0] == port.lib().orc_refseed_from_ics(o.ics(), I)
        if not w["pair"].meta.get("E1"):      # without duplicate lines ics() == C / I
            assert a.counters()[0] == port.lib().orc_refseed_conserved(C, I)
        hits += int(a.counters()[0] != C)
    assert hits >= 0


# ------------------------------------------------------------------ variation operators
@pytest.mark.parametrize("total", [False, True])
def test_edit_replay_is_mutate_crossover_onbit(world, total):
    """the generated edit lists (ref_gen_mutate_ops / ref_gen_cross_draws) replayed by
    orc_apply_edits == the real Alignment::mutate / UPMX constructor / onBit: permutation,
    mask, counters and fp64 accumulators, exactly."""
    w = world
    n2 = w["pair"].n2

    def mirror(a):
        st = a.state()
        o = port.Aln(w["prob"], st[0], st[2])
        o.set_counters(*a.counters())
        return o

    def same(a, o, tag):
        assert all(np.array_equal(x, y) for x, y in zip(a.state(), o.state())), tag
        assert a.counters() == o.counters(), (tag, a.counters(), o.counters())

    for t in range(8):
        a = ref.Aln(w["n1"], w["n2"], w["B"], w["G"])
        a.shuf(100 + t, total=total)
        b = ref.Aln(w["n1"], w["n2"], w["B"], w["G"])
        b.shuf(200 + t, total=total)
        o = mirror(a)
        ops = ref.gen_mutate_ops(300 + t, n2, 0.03, total)
        a.mutate(300 + t, 0.03, total)
        o.apply_edits(ops)
        same(a, o, ("mutate", t))
        first, cops = ref.gen_cross_ops(400 + t, 0.2, (a.state()[0], a.state()[2]),
                                        (b.state()[0], b.state()[2]), total)
        child = ref.Aln.crossover(400 + t, 0.2, a, b, total)
        oc = mirror(a if first else b)
        oc.apply_edits(cops)
        same(child, oc, ("cross", t))
        m = child.state()[2]
        for x in list(np.flatnonzero(m == 0)[:3]) + [0, n2 - 1]:
            child.on_bit(int(x))
            oc.apply_edits([[2, int(x), 0, 0]])
            same(child, oc, ("onbit", t, x))


# ------------------------------------------------------------------ NSGA-II ranking
def test_gnu_sort_restatement_matches_std_sort():
    """orc_gnu_sort_by_key places tied keys exactly where this toolchain's std::sort does
    (setCrowdingDists sorts fronts with it, nsga-ii.cpp:98-101); the heap-sort fallback of the
    depth limit is pinned separately against std::partial_sort."""
    rng = np.random.default_rng(5)
    for t in range(600):
        n = int(rng.integers(1, 400)) if t % 10 else int(rng.integers(1000, 5000))
        kind = t % 5
        if kind == 0:
            k = rng.random(n)
        elif kind == 1:
            k = rng.integers(0, 5, n).astype(np.float64)
        elif kind == 2:
            k = rng.integers(0, max(2, n // 4), n).astype(np.float64)
        elif kind == 3:
            k = np.sort(rng.integers(0, 50, n)).astype(np.float64)
        else:
            k = np.sort(rng.integers(0, 50, n))[::-1].astype(np.float64)
        idx = rng.permutation(n) if t % 2 else np.arange(n)
        assert np.array_equal(ref.std_sort_by_key(k, idx), port.gnu_sort_by_key(k, idx)), t
        if n <= 400:
            assert np.array_equal(ref.std_sort_by_key(k, idx, heap=True),
                                  port.gnu_sort_by_key(k, idx, heap=True)), t


def test_nsga2_matches_reference_live():
    import nsga_cases
    rng = np.random.default_rng(6)
    for t in range(120):
        kind = nsga_cases.KINDS[t % len(nsga_cases.KINDS)]
        P, M = int(rng.integers(1, 420)), int(rng.integers(1, 5))
        f = nsga_cases.make_case(kind, P, M, 500 + t)
        hn, nm = nsga_cases.stale_norm(P, M, t) if t % 3 == 0 else (None, None)
        nsga_cases.assert_same(port.nsga2(f, hn, nm), ref.nsga2(f, hn, nm), (t, kind, P, M))
        a, b = f[int(rng.integers(0, P))], f[int(rng.integers(0, P))]
        assert port.dominates(a, b) == ref.dominates(a, b)


def test_nsga2_edge_shapes_match_reference():
    """one row, identical rows, 16 objectives, a population that is one long chain or one flat front"""
    import nsga_cases
    rng = np.random.default_rng(8)
    cases = [np.array([[0.3, 0.7]]), np.full((37, 3), 0.25), rng.random((90, 16)),
             np.round(rng.random((64, 16)), 1), nsga_cases.make_case("chain", 130, 3, 1),
             nsga_cases.make_case("single_front", 257, 4, 2), np.zeros((20, 2)), -rng.random((50, 2))]
    for i, f in enumerate(cases):
        nsga_cases.assert_same(port.nsga2(f), ref.nsga2(f), i)


MT_RANGES = [(6000, 8000), (5499, 9633), (20000, 20000), (1, 2), (3, 5), (1, 3),
             (1_500_000_000, 2_000_000_000),      # about every third draw is thrown away and repeated
             (2_147_483_647, 2_147_483_647), (1_073_741_825, 1_073_741_826)]


def test_mt_proposal_stream_matches_libstdcxx():
    """orc_mt_seed / orc_mt_xy_stream (MT19937 + Lemire's mapping as libstdc++ does it) against the
    real std::mt19937 and uniform_int_distribution<int> in the compiled shim: proposals AND the
    generator state left behind, from fresh seeds and from states in the middle of a block."""
    for seed in (0, 1, 7, 42, 5489, 0xFFFFFFFF):
        assert np.array_equal(port.mt_state(seed), ref.mt_state(seed)), seed
    for t, (n1, n2) in enumerate(MT_RANGES):
        st = ref.mt_state(100 + t)
        for iters in (1, 311, 312, 313, 5000, 1):      # odd totals leave an x pending across blocks
            xr, sr = ref.gen_xy_stream_state(st, n1, n2, iters)
            xo, so = port.mt_xy_stream(st, n1, n2, iters)
            assert np.array_equal(xr, xo), (n1, n2, iters)
            assert np.array_equal(sr, so), (n1, n2, iters)
            st = sr
    assert np.array_equal(ref.gen_xy_stream(7, 600, 800, 4000), port.mt_xy_stream(port.mt_state(7), 600, 800, 4000)[0])
