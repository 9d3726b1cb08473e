"""GPU: reference callers work unchanged on the drop-in Alignment.cpp / localSearch.cpp.

optnetaligncpp_b200/host/_build/dropin_demo is a driver in the shape of the reference's
hillclimber.cpp, compiled against the reference's UNCHANGED headers and loader sources and
linked with this repo's replacement translation units (built in the dev container, where the
reference tree is).  It must reproduce, bit for bit, what the unmodified reference
(oracle/_ref) computes for the same seeds: shuf -> computeFitness -> correctHillClimb (several
host threads, combined into one batched GPU climb) -> UPMX crossover -> mutate -> ics()."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from optnetaligncpp_b200 import synth
from oracle import ref

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "optnetaligncpp_b200", "host", "_build", "dropin_demo")


@pytest.mark.skipif(not os.path.exists(DEMO) or not ref.available(),
                    reason="drop-in demo / compiled reference not built (dev container only)")
@pytest.mark.parametrize("lines", ["plain", "duplicate-lines"])
@pytest.mark.parametrize("with_goc", [False, True])
def test_dropin_matches_reference(with_goc, lines):
    pair = synth.make_pair(n1=400, n2=520, m1=2000, m2=2300, max_deg=60, seed=41,
                           annotations=with_goc, n_terms=200, bit_density=0.01)
    if lines == "duplicate-lines":      # "a b" + "b a", repeated lines: legal input upstream
        pair = synth.with_duplicate_lines(pair, 150, seed=8)
    d = tempfile.mkdtemp()
    paths = synth.write_text(pair, os.path.join(d, "w"))
    if with_goc and lines == "duplicate-lines":
        # annotation files as untidy as the reference's loader tolerates (goc.cpp:55-89): names the
        # network does not know (they land on node 0, the last one wins), a node listed twice (the last
        # line replaces the first), a line without terms (an empty set), an empty line, repeated terms on one line -- the replacement loadGOC / loadAnnotations
        # (host/goc_b200.cpp, pair loop on the GPU) must give the reference's matrix value for value
        for tag, node in (("annos1", "a7"), ("annos2", "b9")):
            with open(paths[tag], "a") as f:
                f.write("nobody%s 3 5 7 11\n" % tag)
                f.write("%s 2 3 3 5 5 5\n" % node)
                if tag == "annos1":      # (an empty set in BOTH files would put a NaN = 0/0 into the matrix, and
                    f.write("%s\n" % (node[0] + "11"))   # the reference's ranking is undefined on NaN objectives)
                    f.write("\n")
                f.write("somebodyelse 1 2\n")
    seed, iters, nthreads = 77, 4000, 3
    a1 = paths.get("annos1", "-") if with_goc else "-"
    a2 = paths.get("annos2", "-") if with_goc else "-"
    # with GOC: correctHillClimb's proposals come from the device (the caller's mt19937 state travels,
    # ona_hill_climb_mt) whatever the iteration count; without: drawn on the host below 8192 iterations
    env = dict(os.environ, ONA_HOST_MT_MIN="0") if with_goc else dict(os.environ)
    out = subprocess.run([DEMO, paths["net1"], paths["net2"], paths["sim"], a1, a2, str(seed), str(iters),
                          str(nthreads), os.path.join(d, "out")], capture_output=True, text=True, timeout=300,
                         env=env)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.strip().splitlines()]

    n1, n2 = ref.Net(paths["net1"]), ref.Net(paths["net2"])
    B = ref.Dict.load_blast(n1, n2, paths["sim"])
    G = ref.Dict.load_goc(n1, n2, a1, a2) if with_goc else None
    names = [6, 2] + ([5] if with_goc else [])
    climbed = []
    for t in range(nthreads):
        a = ref.Aln(n1, n2, B, G)
        a.shuf(seed + t)
        a.correct_hill_climb(seed + 1000 + t, iters, names)
        climbed.append(a)
        got = lines[t]
        assert got[0] == "climb" and int(got[1]) == t
        c = a.counters()
        assert (int(got[3]), int(got[5])) == (c[0], c[1])
        assert float(got[7]) == c[2] and float(got[9]) == c[3]
        assert [float(v) for v in got[10:]] == list(a.compute_fitness(names))
        # Alignment::save output == the reference's permutation
        al = a.state()[0]
        saved = [l.split() for l in open(os.path.join(d, "out_%d.aln" % t))]
        assert len(saved) == pair.n1
        assert all(int(u[1:]) == i and int(v[1:]) == al[i] for i, (u, v) in enumerate(saved))
    # the archive's result files (steadystate.cpp:339-395): header with '*' on the optimised
    # objectives, one TSV row per alignment in the stream's default %g formatting, '?' without data
    info = [l.rstrip("\n").split("\t") for l in open(os.path.join(d, "out_res.info"))]
    starred = {"S3", "BitscoreSum"} | ({"GOC"} if with_goc else set())
    cols = ["EC", "ICS", "S3", "GOC", "BitscoreSum", "EvalsSum", "Size", "ICS*EC", "S3Variant"]
    assert info[0] == ["filename"] + [c + ("*" if c in starred else "") for c in cols]
    assert len(info) == 1 + nthreads
    for t in range(nthreads):
        a = climbed[t]
        # (on a clone: computeFitness never shrinks Alignment::fitness, and swapNormalizedDelta
        # below reads whatever that vector holds, localSearch.cpp:90)
        f = dict(zip(range(10), a.clone().compute_fitness(list(range(10)))))   # fitnessName order, Network.h:31
        c = a.counters()
        want = [os.path.join(d, "out_res_%d.aln" % t), "%g" % f[1], "%g" % f[0], "%g" % f[6],
                "%g" % c[3] if with_goc else "?", "%g" % c[2], "%g" % c[2], "%g" % a.aln_size(),
                "%g" % f[8], "%g" % f[9]]
        assert info[1 + t] == want, (info[1 + t], want)
        assert open(os.path.join(d, "out_res_%d.aln" % t)).read() == open(os.path.join(d, "out_%d.aln" % t)).read()
    for t in range(nthreads):
        child = ref.Aln.crossover(seed + 2000 + t, 0.1, climbed[t], climbed[0], True)
        child.mutate(seed + 3000 + t, 0.005, True)
        got = lines[nthreads + t]
        assert got[0] == "child"
        c = child.counters()
        assert (int(got[3]), int(got[5])) == (c[0], c[1])
        assert float(got[7]) == c[2] and float(got[9]) == c[3]
        assert float(got[11]) == child.ics()
        assert [float(v) for v in got[12:]] == list(child.compute_fitness(names))
    # one generation's NSGA-II ranking, checked inside the demo against the reference's own
    # nsga-ii.cpp (linked unchanged): fronts, member order, domRank, crowdDist, fitnessNormalized
    got = lines[2 * nthreads]
    assert got[0] == "nsga2" and int(got[2]) >= 1 and got[3] == "identical", got
    # proportionalSearch (localSearch.cpp:60-84)
    ps = climbed[0].clone()
    ps.proportional_search(seed + 9000, 6, names, 0, 0.5)
    got = lines[2 * nthreads + 1]
    c = ps.counters()
    assert got[0] == "propsearch" and (int(got[2]), int(got[4])) == (c[0], c[1])
    assert float(got[6]) == c[2] and float(got[8]) == c[3]
    assert [float(v) for v in got[9:]] == list(ps.compute_fitness(names))
    snd = climbed[0].swap_normalized_delta(names, 3, 7)
    got = float(lines[2 * nthreads + 2][1])
    assert got == snd or (np.isnan(got) and np.isnan(snd))

    # partial alignments (mask flips, mask AND/OR, onBit go through ona_apply_edits): exact too
    rest = lines[2 * nthreads + 3:]

    def check(row, tag, a):
        assert row[0] == tag, row
        c = a.counters()
        assert (int(row[2]), int(row[4])) == (c[0], c[1]), (tag, row, c)
        assert float(row[6]) == c[2] and float(row[8]) == c[3], (tag, row, c)
        assert float(row[10]) == a.aln_size()
        assert [float(v) for v in row[11:]] == list(a.compute_fitness(names))

    parts = []
    for t in range(2):
        a = ref.Aln(n1, n2, B, G)
        a.shuf(seed + 5000 + t, total=False)
        check(rest[2 * t], "pshuf", a)
        a.mutate(seed + 6000 + t, 0.01, False)
        check(rest[2 * t + 1], "pmut", a)
        parts.append(a)
    child = ref.Aln.crossover(seed + 7000, 0.2, parts[0], parts[1], False)
    check(rest[4], "pcx", child)
    mask = child.state()[2]
    off = int(np.argmin(mask[:pair.n1])) if not mask[:pair.n1].all() else pair.n1
    assert rest[5][0] == "pbit" and int(rest[5][1]) == off
    child.on_bit(off)
    check(rest[5][2:], "pon", child)
    child.correct_hill_climb(seed + 8000, iters, names)
    check(rest[6], "pclimb", child)
    # greedyMatch and the size-controlled random starts (steadystate.cpp:92-118)
    gm = ref.Aln(n1, n2, B, G)
    gm.greedy_match(True)
    check(rest[7], "greedy", gm)
    # --total --seeding: the greedyMatch individual (mask bits off on the dummies) as a UPMX parent;
    # under --total the reference leaves mask bits in place (Alignment.cpp:219-227), doSwap would move them
    as_par1 = 0
    for v in range(4):
        gchild = ref.Aln.crossover(seed + 9700 + v, 0.3, gm, climbed[0], True)
        check(rest[8 + v], "gcx", gchild)
        as_par1 += int(not gchild.state()[2].all())
    assert 0 < as_par1 < 4          # both parent orders were exercised
    for v, tag in enumerate(("uniformsize", "smallstart")):
        a = ref.Aln(n1, n2, B, G)
        a.shuf(seed + 9500 + v, uniformsize=(v == 0), smallstart=(v == 1), total=False)
        check(rest[12 + v], tag, a)
    # Alignment::save of a partial alignment + the alignment-file constructor
    loaded = ref.Aln.load(n1, n2, os.path.join(d, "out_partial.aln"), B, G)
    check(rest[14], "pload", loaded)
    sa, sl = child.state(), loaded.state()
    live = sa[2][:pair.n1] != 0
    assert np.array_equal(sl[0][:pair.n1][live], sa[0][:pair.n1][live])


@pytest.mark.skipif(not os.path.exists(DEMO) or not ref.available(),
                    reason="drop-in demo / compiled reference not built (dev container only)")
def test_dropin_steepest_ascent_matches_reference():
    """steepestAscentHillClimb (localSearch.cpp:122-203) through the drop-in: every scan of the
    n1 x n2 grid is one ona_steepest_step; same swaps, same final alignment as the reference's
    threaded CPU scan (integer bit scores, so the reference's swap/un-swap leaves no fp residue)."""
    pair = synth.make_pair(n1=60, n2=80, m1=200, m2=260, max_deg=15, seed=43, bit_density=0.03)
    d = tempfile.mkdtemp()
    paths = synth.write_text(pair, os.path.join(d, "w"))
    seed = 5
    out = subprocess.run([DEMO, paths["net1"], paths["net2"], paths["sim"], "-", "-", str(seed), "0", "2",
                          os.path.join(d, "out"), "steepest"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    got = out.stdout.strip().splitlines()[-1].split()
    n1, n2 = ref.Net(paths["net1"]), ref.Net(paths["net2"])
    B = ref.Dict.load_blast(n1, n2, paths["sim"])
    a = ref.Aln(n1, n2, B, None)
    a.shuf(seed)
    start = a.counters()
    a.steepest_ascent([6, 2], nthreads=2)
    c = a.counters()
    assert c != start                       # the search did move
    assert got[0] == "steepest" and (int(got[2]), int(got[4])) == (c[0], c[1])
    assert float(got[6]) == c[2]
    assert [float(v) for v in got[9:]] == list(a.compute_fitness([6, 2]))
    saved = [l.split() for l in open(os.path.join(d, "out_steepest.aln"))]
    al = a.state()[0]
    assert all(int(u[1:]) == i and int(v[1:]) == al[i] for i, (u, v) in enumerate(saved))
