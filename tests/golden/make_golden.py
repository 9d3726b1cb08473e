"""Generates the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the dev container only (needs /root/reference and oracle/_ref/libona_ref.so):

    python tests/golden/make_golden.py

Outputs
  example_pair.npz   the bundled example pair as the reference's loaders see it
                     (Network.cpp:39-112 node numbering / adjList order, edges.size(),
                     blastinfo.cpp:11-45 non-zeros incl. the "unknown name -> node 0" rule H4)
                     + reference results on it: shuf(mt19937(seed)) individuals, their seeded
                     counters, objectives, per-candidate doSwap records, a correctHillClimb
                     trajectory (proposals from mt19937(7), localSearch.cpp:21-28).
  small_pair.npz     a 120/150-node synthetic pair with bit scores AND GO annotations, with
                     partial (masked) individuals and every objective name.
  nsga2.npz          nonDominatedSort / normalizeFitnesses / setCrowdingDists (nsga-ii.cpp:14-118)
                     results of the compiled reference on the seeded matrices of
                     tests/nsga_cases.py:GOLDEN_CASES, plus the two known-answer cases of the
                     reference's own suite (tests.cpp:355-437).
  edits.npz          partial-alignment variation operators on the small pair: the edit lists that
                     Alignment::mutate / the UPMX constructor / onBit perform for given seeds
                     (Alignment.cpp:152-256, :451-486, :522-533) and the reference's state, counters
                     and fp64 accumulators afterwards (network arrays: small_pair.npz).
Nothing here is read at test time except the .npz files.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref  # noqa: E402
from optnetaligncpp_b200 import synth  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))
ALL_NAMES = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9]


def records(a, xy, names, obj):
    r = a.swap_delta_records(xy, names, obj=obj, want_snd=True)
    return {k: v for k, v in r.items()}


def example_pair():
    n1 = ref.Net(os.path.join(REF, "example1.net"))
    n2 = ref.Net(os.path.join(REF, "example2.net"))
    B = ref.Dict.load_blast(n1, n2, os.path.join(REF, "example.sim"))
    rp1, col1 = n1.csr()
    rp2, col2 = n2.csr()
    S = B.dense()
    br, bc = np.nonzero(S)
    out = dict(n1=n1.n, n2=n2.n, E1=n1.E, E2=n2.E, rp1=rp1, col1=col1, rp2=rp2, col2=col2,
               bit_rows=br.astype(np.int32), bit_cols=bc.astype(np.int32), bit_vals=S[br, bc])
    names = np.array([6, 2], np.int32)          # --s3 --blastsum   (argParsing.h:212-271 order)
    out["names"] = names
    seeds = [42, 43, 44]
    alns, ctr, fit, ics, denom = [], [], [], [], []
    for s in seeds:
        a = ref.Aln(n1, n2, B, None)
        a.shuf(s)
        al, _, _ = a.state()
        alns.append(al)
        ctr.append(a.counters())
        fit.append(a.compute_fitness(names))
        ics.append(a.ics())
        denom.append(a.fast_ics_denominator())
    out["seeds"] = np.array(seeds)
    out["aln"] = np.array(alns, np.int32)
    out["cons"] = np.array([c[0] for c in ctr], np.int32)
    out["ind"] = np.array([c[1] for c in ctr], np.int32)
    out["bits"] = np.array([c[2] for c in ctr])
    out["fitness"] = np.array(fit)
    out["ics"] = np.array(ics)
    out["denom"] = np.array(denom)
    # identity alignment (constructor 1) known answers
    ident = ref.Aln(n1, n2, B, None)
    out["ident_counters"] = np.array(ident.counters())
    # per-candidate records on individual 0, obj=-1 and obj=0
    a = ref.Aln(n1, n2, B, None)
    a.shuf(seeds[0])
    xy = ref.gen_xy_stream(11, n1.n, n2.n, 3000)
    out["cand_xy"] = xy
    for obj in (-1, 0, 1):
        r = records(a, xy, names, obj)
        for k, v in r.items():
            if k == "improved" or obj == -1:
                out["cand_%s_obj%d" % (k, obj)] = v
    # correctHillClimb trajectory
    iters = 60000
    xy = ref.gen_xy_stream(7, n1.n, n2.n, iters)
    b = a.clone()
    b.correct_hill_climb(7, iters, names)          # the real thing
    acc = a.hill_climb_replay(xy, names, want_accepted=True)
    assert a.counters() == b.counters()
    assert all((p == q).all() for p, q in zip(a.state(), b.state()))
    out["hc_seed"] = 7
    out["hc_xy"] = xy
    out["hc_accepted"] = np.packbits(acc)
    out["hc_final_aln"] = a.state()[0]
    out["hc_final_counters"] = np.array(a.counters())
    out["hc_final_fitness"] = a.compute_fitness(names)
    # single-objective climb (obj = 1: bit scores only)
    c = ref.Aln(n1, n2, B, None)
    c.shuf(seeds[1])
    xy1 = ref.gen_xy_stream(8, n1.n, n2.n, 20000)
    acc1 = c.hill_climb_replay(xy1, names, obj=1, want_accepted=True)
    out["hc1_xy"] = xy1
    out["hc1_accepted"] = np.packbits(acc1)
    out["hc1_final_aln"] = c.state()[0]
    out["hc1_final_counters"] = np.array(c.counters())
    np.savez_compressed(os.path.join(OUT, "example_pair.npz"), **out)
    print("example_pair.npz written; identity", out["ident_counters"], "seed42", ctr[0], fit[0],
          "after climb", a.counters())


def small_pair():
    pair = synth.make_pair(n1=120, n2=150, m1=500, m2=560, max_deg=30, seed=synth.DEFAULT_SEED + 9,
                           annotations=True, n_terms=60, terms_per_node=5, bit_density=0.01,
                           bit_decimals=True)
    d = tempfile.mkdtemp()
    paths = synth.write_text(pair, os.path.join(d, "s"))
    n1, n2 = ref.Net(paths["net1"]), ref.Net(paths["net2"])
    B = ref.Dict.load_blast(n1, n2, paths["sim"])
    G = ref.Dict.load_goc(n1, n2, paths["annos1"], paths["annos2"])
    rp1, col1 = n1.csr()
    rp2, col2 = n2.csr()
    assert (rp1 == pair.rp1).all() and (col1 == pair.col1).all()
    out = dict(n1=n1.n, n2=n2.n, E1=n1.E, E2=n2.E, rp1=rp1, col1=col1, rp2=rp2, col2=col2,
               S=B.dense(), G=G.dense(), names=np.array(ALL_NAMES, np.int32))
    names = ALL_NAMES
    alns, masks, ctr, fit, cand, hc = [], [], [], [], [], []
    for i, (seed, total) in enumerate([(1, True), (2, False), (3, False), (4, True)]):
        a = ref.Aln(n1, n2, B, G)
        a.shuf(seed, total=total)
        al, _, m = a.state()
        alns.append(al)
        masks.append(m)
        ctr.append(a.counters())
        fit.append(a.compute_fitness(names))
        # general doSwap support: x may be a dummy position too (mutate, Alignment.cpp:464)
        rng = np.random.default_rng(100 + i)
        xy = np.stack([rng.integers(0, n2.n, 400), rng.integers(0, n2.n, 400)], -1).astype(np.int32)
        r = records(a, xy, names, -1)
        out["cand_xy_%d" % i] = xy
        for k, v in r.items():
            out["cand_%s_%d" % (k, i)] = v
        xyh = ref.gen_xy_stream(50 + i, n1.n, n2.n, 4000)
        obj = [-1, 6, -1, 5][i]
        acc = a.hill_climb_replay(xyh, names, obj=obj, want_accepted=True)
        out["hc_xy_%d" % i] = xyh
        out["hc_obj_%d" % i] = obj
        out["hc_accepted_%d" % i] = acc
        out["hc_final_aln_%d" % i] = a.state()[0]
        out["hc_final_mask_%d" % i] = a.state()[2]
        out["hc_final_counters_%d" % i] = np.array(a.counters())
    out["aln"] = np.array(alns, np.int32)
    out["mask"] = np.array(masks, np.uint8)
    out["cons"] = np.array([c[0] for c in ctr], np.int32)
    out["ind"] = np.array([c[1] for c in ctr], np.int32)
    out["bits"] = np.array([c[2] for c in ctr])
    out["goc"] = np.array([c[3] for c in ctr])
    out["fitness"] = np.array(fit)
    np.savez_compressed(os.path.join(OUT, "small_pair.npz"), **out)
    print("small_pair.npz written", ctr)


def edits():
    pair = synth.make_pair(n1=120, n2=150, m1=500, m2=560, max_deg=30, seed=synth.DEFAULT_SEED + 9,
                           annotations=True, n_terms=60, terms_per_node=5, bit_density=0.01,
                           bit_decimals=True)
    d = tempfile.mkdtemp()
    paths = synth.write_text(pair, os.path.join(d, "s"))
    n1, n2 = ref.Net(paths["net1"]), ref.Net(paths["net2"])
    B = ref.Dict.load_blast(n1, n2, paths["sim"])
    G = ref.Dict.load_goc(n1, n2, paths["annos1"], paths["annos2"])
    out = {}

    def put(tag, a):
        al, _, m = a.state()
        out[tag + "_aln"], out[tag + "_mask"], out[tag + "_counters"] = al, m, np.array(a.counters())

    for c, total in enumerate([False, False, True]):
        a = ref.Aln(n1, n2, B, G)
        a.shuf(2 + c, total=total)
        b = ref.Aln(n1, n2, B, G)
        b.shuf(12 + c, total=total)
        put("c%d_start" % c, a)
        put("c%d_other" % c, b)
        out["c%d_mutate_ops" % c] = ref.gen_mutate_ops(20 + c, n2.n, 0.05, total)
        a.mutate(20 + c, 0.05, total)
        put("c%d_mutated" % c, a)
        first, cops = ref.gen_cross_ops(30 + c, 0.25, (a.state()[0], a.state()[2]),
                                        (b.state()[0], b.state()[2]), total)
        child = ref.Aln.crossover(30 + c, 0.25, a, b, total)
        out["c%d_cross_first" % c] = np.array(int(first))
        out["c%d_cross_ops" % c] = cops
        put("c%d_child" % c, child)
        m = child.state()[2]
        xs = np.array(list(np.flatnonzero(m == 0)[:4]) + [0, n2.n - 1], np.int32)
        for x in xs:
            child.on_bit(int(x))
        out["c%d_onbit_x" % c] = xs
        put("c%d_onbit" % c, child)
    np.savez_compressed(os.path.join(OUT, "edits.npz"), **out)
    print("edits.npz written;", [len(out["c%d_cross_ops" % c]) for c in range(3)], "UPMX steps,",
          [len(out["c%d_mutate_ops" % c]) for c in range(3)], "mutate edits")


def nsga2():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import nsga_cases
    out = {}
    for i, (kind, P, M, seed, stale) in enumerate(nsga_cases.GOLDEN_CASES):
        f = nsga_cases.make_case(kind, P, M, seed)
        hn, nm = nsga_cases.stale_norm(P, M, seed) if stale else (None, None)
        r = ref.nsga2(f, hn, nm)
        for k in ("dom_rank", "front_order", "front_offsets", "order_after", "crowd_dist", "norm"):
            out["c%d_%s" % (i, k)] = r[k]
    # tests.cpp:355-408 (non_dominated_sort) and :410-437 (crowding_dist_assignment)
    ka1 = np.array([[0.8, 0.7], [0.75, 0.75], [0.4, 0.6], [0.7, 0.8], [0.6, 0.4], [0.5, 0.5]])
    ka2 = np.array([[0.6, 0.9], [0.7, 0.8], [0.75, 0.75], [0.8, 0.7], [0.9, 0.6]])
    for name, f in (("ka1", ka1), ("ka2", ka2)):
        r = ref.nsga2(f)
        out[name + "_fitness"] = f
        for k in ("dom_rank", "front_order", "front_offsets", "order_after", "crowd_dist", "norm"):
            out["%s_%s" % (name, k)] = r[k]
    np.savez_compressed(os.path.join(OUT, "nsga2.npz"), **out)
    print("nsga2.npz written;", len(nsga_cases.GOLDEN_CASES), "cases; ka1 fronts", out["ka1_front_offsets"],
          "ka2 crowd", out["ka2_crowd_dist"])


MT_CASES = [(7, 6000, 8000, 3000), (42, 5499, 9633, 2000), (1, 3, 5, 1000),
            (9, 1_500_000_000, 2_000_000_000, 1500)]


def mt_stream():
    """the (x, y) proposals of localSearch.cpp:21-28 drawn by the compiled reference toolchain's own
    std::mt19937 / uniform_int_distribution<int>, and the generator state they leave behind"""
    out = {"cases": np.array(MT_CASES, np.int64)}
    for i, (seed, n1, n2, iters) in enumerate(MT_CASES):
        st = ref.mt_state(seed)
        xy, after = ref.gen_xy_stream_state(st, n1, n2, iters)
        assert np.array_equal(xy, ref.gen_xy_stream(seed, n1, n2, iters))
        out["c%d_state" % i], out["c%d_xy" % i], out["c%d_after" % i] = st, xy, after
    np.savez_compressed(os.path.join(OUT, "mt_stream.npz"), **out)
    print("mt_stream.npz written;", [out["c%d_xy" % i][0].tolist() for i in range(len(MT_CASES))])


if __name__ == "__main__":
    assert ref.available(), "build oracle/_ref first (make -C oracle)"
    if sys.argv[1:] == ["mt_stream"]:
        mt_stream()
        sys.exit(0)
    mt_stream()
    nsga2()
    edits()
    small_pair()
    example_pair()
