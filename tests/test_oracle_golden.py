"""CPU: the C restatement (oracle/ona_oracle.c) against golden vectors that
tests/golden/make_golden.py generated from the unmodified compiled reference."""
import numpy as np
import pytest

from conftest import dense_from_coo
from oracle import port


@pytest.fixture(scope="module")
def example_problem(example_golden):
    g = example_golden
    S = dense_from_coo(int(g["n1"]), int(g["n2"]), g["bit_rows"], g["bit_cols"], g["bit_vals"])
    return port.Problem(g["n1"], g["n2"], g["E1"], g["rp1"], g["col1"], g["rp2"], g["col2"], S, None)


@pytest.fixture(scope="module")
def small_problem(small_golden):
    g = small_golden
    return port.Problem(g["n1"], g["n2"], g["E1"], g["rp1"], g["col1"], g["rp2"], g["col2"],
                        g["S"], g["G"])


def test_example_identity_known_answers(example_golden, example_problem):
    # identity alignment of the example pair: C=584, I=27789, bits=685
    n2 = int(example_golden["n2"])
    a = port.Aln(example_problem, np.arange(n2, dtype=np.int32))
    c = a.counters()
    assert list(example_golden["ident_counters"][:3]) == [584, 27789, 685]
    assert (c[0], c[1], c[2]) == (584, 27789, 685.0)


def test_example_full_fitness(example_golden, example_problem):
    g = example_golden
    for i in range(len(g["seeds"])):
        a = port.Aln(example_problem, g["aln"][i])
        c = a.counters()
        assert c[0] == g["cons"][i] and c[1] == g["ind"][i] and c[2] == g["bits"][i]
        assert a.induced_true() == g["denom"][i]
        assert a.ics() == g["ics"][i]
        assert np.array_equal(a.compute_fitness(g["names"]), g["fitness"][i])
    # seed 42 known answer
    assert g["cons"][0] == 140 and g["ind"][0] == 12177 and g["bits"][0] == 360.0
    assert g["fitness"][0][0] == 0.0031865255491066347


@pytest.mark.parametrize("obj", [-1, 0, 1])
def test_example_candidate_records(example_golden, example_problem, obj):
    g = example_golden
    a = port.Aln(example_problem, g["aln"][0])
    r = a.swap_delta_records(g["cand_xy"], g["names"], obj=obj, want_snd=True)
    assert np.array_equal(r["improved"], g["cand_improved_obj%d" % obj])
    if obj == -1:
        for k in ("cons", "ind", "bits", "goc", "fit", "snd"):
            assert np.array_equal(r[k], g["cand_%s_obj-1" % k], equal_nan=True), k


def test_example_hill_climb_trajectory(example_golden, example_problem):
    g = example_golden
    a = port.Aln(example_problem, g["aln"][0])
    acc = a.hill_climb_replay(g["hc_xy"], g["names"], obj=-1, want_accepted=True)
    assert np.array_equal(np.packbits(acc), g["hc_accepted"])
    assert np.array_equal(a.state()[0], g["hc_final_aln"])
    assert list(a.counters()) == list(g["hc_final_counters"])
    assert np.array_equal(a.compute_fitness(g["names"]), g["hc_final_fitness"])
    b = port.Aln(example_problem, g["aln"][1])
    acc = b.hill_climb_replay(g["hc1_xy"], g["names"], obj=1, want_accepted=True)
    assert np.array_equal(np.packbits(acc), g["hc1_accepted"])
    assert np.array_equal(b.state()[0], g["hc1_final_aln"])
    assert list(b.counters()) == list(g["hc1_final_counters"])


def test_small_pair_all_objectives_and_masks(small_golden, small_problem):
    g = small_golden
    names = g["names"]
    for i in range(4):
        a = port.Aln(small_problem, g["aln"][i], g["mask"][i])
        c = a.counters()
        assert (c[0], c[1]) == (g["cons"][i], g["ind"][i])
        assert c[2] == g["bits"][i] and c[3] == g["goc"][i]
        assert np.array_equal(a.compute_fitness(names), g["fitness"][i], equal_nan=True)
        r = a.swap_delta_records(g["cand_xy_%d" % i], names, obj=-1, want_snd=True)
        for k in ("cons", "ind", "bits", "goc", "fit", "improved", "snd"):
            assert np.array_equal(r[k], g["cand_%s_%d" % (k, i)], equal_nan=True), (i, k)
        acc = a.hill_climb_replay(g["hc_xy_%d" % i], names, obj=int(g["hc_obj_%d" % i]),
                                  want_accepted=True)
        assert np.array_equal(acc, g["hc_accepted_%d" % i])
        st = a.state()
        assert np.array_equal(st[0], g["hc_final_aln_%d" % i])
        assert np.array_equal(st[2], g["hc_final_mask_%d" % i])
        assert list(a.counters()) == list(g["hc_final_counters_%d" % i])


def test_refseed_truncation_hazard():
    # (int)((C/I)*I) lands one below C for some pairs
    assert port.lib().orc_refseed_conserved(15, 22) == 14
    assert port.lib().orc_refseed_conserved(104, 11500) == 103
    assert port.lib().orc_refseed_conserved(140, 12177) == 140
    assert port.lib().orc_refseed_conserved(5, 0) == 0


# ------------------------------------------------------------------ variation operators' edit lists
def replay_edits_golden(g, make, c):
    """shared by the oracle test here and the GPU test: `make(aln, mask, counters)` returns an
    object with apply_edits(ops) and snapshot() -> (aln, mask, counters)"""
    def check(obj, tag):
        al, m, ctr = obj.snapshot()
        assert np.array_equal(al, g[tag + "_aln"]) and np.array_equal(m, g[tag + "_mask"]), tag
        assert tuple(ctr) == tuple(g[tag + "_counters"]), (tag, ctr, g[tag + "_counters"])
    a = make(g["c%d_start_aln" % c], g["c%d_start_mask" % c], g["c%d_start_counters" % c])
    a.apply_edits(g["c%d_mutate_ops" % c])
    check(a, "c%d_mutated" % c)
    src = "c%d_mutated" % c if int(g["c%d_cross_first" % c]) else "c%d_other" % c
    child = make(g[src + "_aln"], g[src + "_mask"], g[src + "_counters"])
    child.apply_edits(g["c%d_cross_ops" % c])
    check(child, "c%d_child" % c)
    child.apply_edits([[2, int(x), 0, 0] for x in g["c%d_onbit_x" % c]])
    check(child, "c%d_onbit" % c)


@pytest.mark.parametrize("c", [0, 1, 2])
def test_edit_lists_golden(edits_golden, small_problem, c):
    """mutate with mask flips, UPMX with mask AND/OR, onBit (Alignment.cpp:152-256, :451-486,
    :522-533): state, integer counters and fp64 accumulators exactly as the reference left them."""
    class Obj:
        def __init__(self, aln, mask, ctr):
            self.a = port.Aln(small_problem, aln, mask)
            self.a.set_counters(int(ctr[0]), int(ctr[1]), float(ctr[2]), float(ctr[3]))

        def apply_edits(self, ops):
            self.a.apply_edits(ops)

        def snapshot(self):
            al, _, m = self.a.state()
            return al, m, self.a.counters()
    replay_edits_golden(edits_golden, Obj, c)


# ------------------------------------------------------------------ NSGA-II ranking (nsga-ii.cpp:14-118)
NSGA_KEYS = ("dom_rank", "front_order", "front_offsets", "order_after", "crowd_dist", "norm")


def test_nsga2_golden_cases(nsga_golden):
    import nsga_cases
    for i, (kind, P, M, seed, stale) in enumerate(nsga_cases.GOLDEN_CASES):
        f = nsga_cases.make_case(kind, P, M, seed)
        hn, nm = nsga_cases.stale_norm(P, M, seed) if stale else (None, None)
        got = port.nsga2(f, hn, nm)
        want = {k: nsga_golden["c%d_%s" % (i, k)] for k in NSGA_KEYS}
        nsga_cases.assert_same(got, want, (i, kind))


def test_nsga2_reference_suite_known_answers(nsga_golden):
    """tests.cpp:355-408 (two fronts of three) and :410-437 (crowding 1.0, 2/3, 1.0)."""
    import nsga_cases
    g = nsga_golden
    r = port.nsga2(g["ka1_fitness"])
    assert list(r["front_offsets"]) == [0, 3, 6]
    assert (g["ka1_fitness"][r["front_order"][:3]] > 0.6).all()
    assert (g["ka1_fitness"][r["front_order"][3:]] < 0.7).all()
    nsga_cases.assert_same(r, {k: g["ka1_" + k] for k in NSGA_KEYS}, "ka1")
    r = port.nsga2(g["ka2_fitness"])
    assert list(r["front_offsets"]) == [0, 5]
    assert np.allclose(r["crowd_dist"][1:4], [1.0, 2.0 / 3.0, 1.0], rtol=1e-12)
    assert r["crowd_dist"][0] == r["crowd_dist"][4] == np.finfo(np.float64).max
    nsga_cases.assert_same(r, {k: g["ka2_" + k] for k in NSGA_KEYS}, "ka2")


# ------------------------------------------------------------------ proposal stream (localSearch.cpp:21-28)
def test_mt_proposal_stream_golden(mt_golden):
    """the C restatement of std::mt19937 + libstdc++'s uniform_int_distribution against proposals and
    end states recorded from the real ones (make_golden.py:mt_stream)"""
    g = mt_golden
    for i, (seed, n1, n2, iters) in enumerate(g["cases"].tolist()):
        st = port.mt_state(seed)
        assert np.array_equal(st, g["c%d_state" % i])
        xy, after = port.mt_xy_stream(st, n1, n2, iters)
        assert np.array_equal(xy, g["c%d_xy" % i]) and np.array_equal(after, g["c%d_after" % i]), i
        # resumable: the stream drawn in two calls equals the stream drawn in one
        h = iters // 3
        xa, sa = port.mt_xy_stream(st, n1, n2, h)
        xb, sb = port.mt_xy_stream(sa, n1, n2, iters - h)
        assert np.array_equal(np.concatenate([xa, xb]), xy) and np.array_equal(sb, after)
