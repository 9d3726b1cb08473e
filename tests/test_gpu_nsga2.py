"""GPU (-m gpu): ona_nsga2_rank -- nonDominatedSort, normalizeFitnesses and setCrowdingDists
(nsga-ii.cpp:14-118) on the device -- against the golden vectors of the compiled reference and
against the oracle.  Everything is compared exactly: ranks, the order of the members inside
every front before and after the crowding sorts (ties included), and the fp64 distances."""
import numpy as np
import pytest

import nsga_cases
from optnetaligncpp_b200 import engine, synth
from oracle import port

pytestmark = pytest.mark.gpu

KEYS = ("dom_rank", "front_order", "front_offsets", "order_after", "crowd_dist", "norm")


@pytest.fixture(scope="module")
def eng():
    pair = synth.make_pair(n1=60, n2=80, m1=150, m2=170, max_deg=12, seed=3)
    return engine.Engine.from_pair(pair, [1, 6])


@pytest.fixture(params=["one_cta", "multi_kernel"])
def small_path(request, monkeypatch):
    """populations of <= 1024 rows are ranked by one CTA; ONA_NSGA_MULTI forces the per-front
    kernel pairs that larger populations use, so both paths see every small case"""
    if request.param == "multi_kernel":
        monkeypatch.setenv("ONA_NSGA_MULTI", "1")
    return request.param


def test_golden_cases(eng, nsga_golden, small_path):
    for i, (kind, P, M, seed, stale) in enumerate(nsga_cases.GOLDEN_CASES):
        f = nsga_cases.make_case(kind, P, M, seed)
        hn, nm = nsga_cases.stale_norm(P, M, seed) if stale else (None, None)
        got = eng.nsga2_rank(f, hn, nm)
        nsga_cases.assert_same(got, {k: nsga_golden["c%d_%s" % (i, k)] for k in KEYS}, (i, kind))


def test_reference_suite_known_answers(eng, nsga_golden, small_path):
    g = nsga_golden
    for name in ("ka1", "ka2"):  # tests.cpp:355-408, :410-437
        r = eng.nsga2_rank(g[name + "_fitness"])
        nsga_cases.assert_same(r, {k: g["%s_%s" % (name, k)] for k in KEYS}, name)
    assert np.allclose(r["crowd_dist"][1:4], [1.0, 2.0 / 3.0, 1.0], rtol=1e-12)


@pytest.mark.parametrize("kind", nsga_cases.KINDS)
def test_random_cases_against_oracle(eng, kind, small_path):
    rng = np.random.default_rng(nsga_cases.KINDS.index(kind))
    for t in range(12):
        P, M = int(rng.integers(1, 1025 if t % 4 == 0 else 700)), int(rng.integers(1, 6))
        f = nsga_cases.make_case(kind, P, M, 900 + t)
        hn, nm = nsga_cases.stale_norm(P, M, t) if t % 3 == 0 else (None, None)
        nsga_cases.assert_same(eng.nsga2_rank(f, hn, nm), port.nsga2(f, hn, nm), (kind, t, P, M))


@pytest.mark.parametrize("kind,P,M", [("continuous", 4096, 3), ("counter_ratios", 3000, 3),
                                      ("clones", 2500, 2), ("chain", 300, 2)])
def test_larger_populations_against_oracle(eng, kind, P, M):
    f = nsga_cases.make_case(kind, P, M, 42)
    nsga_cases.assert_same(eng.nsga2_rank(f), port.nsga2(f), (kind, P, M))


def test_front_larger_than_shared_memory(eng):
    """one front of 15 000 members (> 14 336): the crowding CTA works in HBM scratch; with and
    without ties (the replayed std::sort)."""
    P = 15000
    f = nsga_cases.make_case("single_front", P, 2, 7)
    nsga_cases.assert_same(eng.nsga2_rank(f), port.nsga2(f), "tie-free")
    f2 = np.round(f, 3)
    f2[:, 1] = 1.0 - f2[:, 0]
    nsga_cases.assert_same(eng.nsga2_rank(f2), port.nsga2(f2), "ties")


def test_full_size_invariants(eng):
    """P = 16 384 (C5's upper population): the fronts partition the population, no member of a
    front is dominated inside it, every member of front k+1 is dominated by one of front k."""
    P, M = 16384, 3
    f = nsga_cases.make_case("continuous", P, M, 11)
    r = eng.nsga2_rank(f)
    offs, order, rank = r["front_offsets"], r["front_order"], r["dom_rank"]
    assert offs[0] == 0 and offs[-1] == P and np.array_equal(np.sort(order), np.arange(P))
    assert np.array_equal(np.sort(r["order_after"]), np.arange(P))
    for k in range(len(offs) - 1):
        mem = order[offs[k]:offs[k + 1]]
        assert (rank[mem] == k).all()
        assert set(r["order_after"][offs[k]:offs[k + 1]]) == set(mem)

    def dominated_by(A, B):  # for each row of A: does some row of B dominate it
        out = np.zeros(len(A), bool)
        for s in range(0, len(A), 512):
            a = A[s:s + 512, None, :]
            out[s:s + 512] = ((B[None] >= a).all(2) & (B[None] > a).any(2)).any(1)
        return out
    for k in range(min(6, len(offs) - 1)):
        cur = f[order[offs[k]:offs[k + 1]]]
        assert not dominated_by(cur, cur).any()
        if k + 2 < len(offs):
            assert dominated_by(f[order[offs[k + 1]:offs[k + 2]]], cur).all()
    # tie-free continuous objectives: the distances follow from the sorted neighbours alone
    mem = order[offs[0]:offs[1]]
    want = np.zeros(len(mem))
    for m in range(M):
        v = r["norm"][mem, m]
        o = np.argsort(v, kind="stable")
        d = np.empty(len(mem))
        d[o[1:-1]] = (v[o[2:]] - v[o[:-2]]) / (v[o[-1]] - v[o[0]])
        want += d
        want[o[0]] = want[o[-1]] = np.finfo(np.float64).max
    assert np.array_equal(r["crowd_dist"][mem], want)


def test_edge_shapes(eng, small_path):
    """one row, identical rows, 16 objectives, all-zero and all-negative objectives (normalisation by a
    maximum that stays 0.0: inf / NaN exactly as the reference produces them)"""
    rng = np.random.default_rng(8)
    cases = [np.array([[0.3, 0.7]]), np.full((37, 3), 0.25), rng.random((90, 16)),
             np.round(rng.random((64, 16)), 1), np.zeros((20, 2)), -rng.random((50, 2))]
    for i, f in enumerate(cases):
        nsga_cases.assert_same(eng.nsga2_rank(f), port.nsga2(f), i)


def test_device_buffers_and_island_gather(eng):
    """ona_nsga2_rank_dev on torch device tensors, as used on the all-gathered objective rows."""
    import torch
    P, M = 777, 3
    f = nsga_cases.make_case("counter_ratios", P, M, 5)
    want = port.nsga2(f)
    d = torch.device("cuda:0")
    ft = torch.from_numpy(f).to(d)
    rank, order, after = (torch.empty(P, dtype=torch.int32, device=d) for _ in range(3))
    offs = torch.empty(P + 1, dtype=torch.int32, device=d)
    crowd = torch.empty(P, dtype=torch.float64, device=d)
    norm = torch.empty(P, M, dtype=torch.float64, device=d)
    nf = eng.nsga2_rank_dev(P, M, ft.data_ptr(), None, norm.data_ptr(), rank.data_ptr(), order.data_ptr(),
                            offs.data_ptr(), after.data_ptr(), crowd.data_ptr())
    eng.synchronize()
    got = {"dom_rank": rank.cpu().numpy(), "front_order": order.cpu().numpy(),
           "front_offsets": offs.cpu().numpy()[:nf + 1], "order_after": after.cpu().numpy(),
           "crowd_dist": crowd.cpu().numpy(), "norm": norm.cpu().numpy()}
    nsga_cases.assert_same(got, want, "dev")
    # outputs are optional
    assert eng.nsga2_rank_dev(P, M, ft.data_ptr()) == nf


def test_argument_errors(eng):
    with pytest.raises(engine.OnaError):
        eng.nsga2_rank(np.zeros((4, 17)))
    with pytest.raises(engine.OnaError):  # P = 0
        eng._check(eng.L.ona_nsga2_rank(eng.h, 0, 2, None, None, None, None, None, None, None, None, None))
