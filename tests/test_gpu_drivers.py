"""GPU: the reference's three drivers, UNMODIFIED, linked against the drop-in translation units.

optnetaligncpp_b200/host/_build/{optnetalign,nsgaii,hillclimber} are steadystate.cpp,
nsgaiimain.cpp and hillclimber.cpp compiled where they lie under the reference tree (with
argParsing.h and every other caller unchanged) against this repo's Alignment.cpp / localSearch.cpp
and the CUDA engine; oracle/_ref/*_ref are the same drivers built entirely from reference sources.
Both sets are built in the dev container (Boost.ProgramOptions / TBB stand-ins:
optnetaligncpp_b200/host/compat) and travel to the GPU box.

The drivers seed their worker PRNGs with clock() (steadystate.cpp:133, nsgaiimain.cpp:78,195), so
whole-program output is not reproducible even for the reference itself.  What is
checked is what must hold for ANY run: the CLI contract (same flags, same files), that every
alignment written is a valid alignment, and that the objective values the driver reports for it --
maintained incrementally on the GPU across crossover, mutation and hill climbing -- are the values
the reference computes from scratch for that very alignment file."""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from optnetaligncpp_b200 import synth
from oracle import ref

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GPU = {n: os.path.join(ROOT, "optnetaligncpp_b200", "host", "_build", n) for n in ("optnetalign", "nsgaii", "hillclimber")}
CPU = {n: os.path.join(ROOT, "oracle", "_ref", n + "_ref") for n in GPU}
built = all(os.path.exists(p) for p in list(GPU.values()) + list(CPU.values())) and ref.available()
needs_build = pytest.mark.skipif(not built, reason="driver binaries / compiled reference not built (dev container only)")


@pytest.fixture(scope="module")
def world():
    pair = synth.make_pair(n1=400, n2=520, m1=2000, m2=2300, max_deg=60, seed=41,
                           annotations=True, n_terms=200, bit_density=0.01)
    d = tempfile.mkdtemp()
    paths = synth.write_text(pair, os.path.join(d, "w"))
    n1, n2 = ref.Net(paths["net1"]), ref.Net(paths["net2"])
    B = ref.Dict.load_blast(n1, n2, paths["sim"])
    G = ref.Dict.load_goc(n1, n2, paths["annos1"], paths["annos2"])
    return dict(pair=pair, dir=d, paths=paths, n1=n1, n2=n2, B=B, G=G)


def run(binary, args, cwd):
    out = subprocess.run([binary] + args, cwd=cwd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "error" not in out.stderr.lower(), out.stderr[-2000:]
    return out.stdout


def fmt(v):
    return "%g" % v       # the drivers print doubles with the stream's default precision


def check_results(w, prefix, starred, total, with_goc, with_bits=True):
    """<prefix>.info + <prefix>_<i>.aln against the reference's own from-scratch scoring"""
    info = [l.rstrip("\n").split("\t") for l in open(prefix + ".info")]
    cols = ["EC", "ICS", "S3", "GOC", "BitscoreSum", "EvalsSum", "Size", "ICS*EC", "S3Variant"]
    assert info[0] == ["filename"] + [c + ("*" if c in starred else "") for c in cols]
    assert len(info) >= 2
    pair = w["pair"]
    best_s3 = 0.0
    for row in info[1:]:
        path = row[0] if os.path.isabs(row[0]) else os.path.join(os.path.dirname(prefix), os.path.basename(row[0]))
        lines = [l.split() for l in open(path)]
        us = [int(u[1:]) for u, _ in lines]
        vs = [int(v[1:]) for _, v in lines]
        assert len(set(us)) == len(us) and len(set(vs)) == len(vs)           # injective
        assert all(0 <= u < pair.n1 for u in us) and all(0 <= v < pair.n2 for v in vs)
        if total:
            assert len(us) == pair.n1
        # the reference scores the written file from scratch (alignment-file constructor,
        # Alignment.cpp:49-138); its ics() / sums / size are exact, its cached conserved count is the
        # truncated seed (H1), so the driver's incrementally maintained count may sit one above it
        a = ref.Aln.load(w["n1"], w["n2"], path, w["B"] if with_bits else None, w["G"] if with_goc else None)
        C_seed, I = a.counters()[0], a.counters()[1]
        E1 = pair.E1
        ok = False
        for C in (C_seed, C_seed + 1, C_seed - 1):
            ec, ics = C / E1, (C / I if I else 0.0)
            s3 = C / (E1 + I - C)
            want = [fmt(ec), fmt(ics), fmt(s3)]
            if row[1:4] == want:
                ok = True
                best_s3 = max(best_s3, s3)
                assert row[8] == fmt(ics * ec)
        if not total and not ok:
            # Partial alignments: the REFERENCE's incrementally maintained currInducedCount drifts a few
            # units away from fastICSDenominator() once mask flips and UPMX mask AND/OR are in play --
            # the unmodified CPU build of this driver (oracle/_ref/optnetalign_ref) reports ICS values
            # that are off by the same kind of margin in about 2 % of its rows.  The engine replays the
            # reference's own update sequence (ona_apply_edits, bit-exact against the reference in
            # tests/test_gpu_parity.py), so it inherits that; here the reported pair must merely be
            # explained by a conserved count within 1 and an induced count within 10 of the recount.
            ec_rep, ics_rep = float(row[1]), float(row[2])
            C_rep = round(ec_rep * E1)
            I_rep = C_rep / ics_rep if ics_rep else I
            ok = abs(C_rep - C_seed) <= 1 and abs(I_rep - I) <= 10 and fmt(C_rep / E1) == row[1]
        assert ok, (row, C_seed, I)
        assert row[4] == (fmt(a.sum_goc()) if with_goc else "?"), (row, a.sum_goc())
        if with_bits:
            assert row[5] == fmt(a.sum_blast()) and row[6] == fmt(a.sum_blast())
        else:
            assert row[5] == "?" and row[6] == "?"
        assert row[7] == fmt(a.aln_size())
    return len(info) - 1, best_s3


@needs_build
def test_optnetalign_total_s3_blastsum(world):
    """BASELINE.json configs[1] flags: --total --s3 --bitscores F --blastsum --popsize --hillclimbiters"""
    w, p = world, world["paths"]
    args = ["--net1", p["net1"], "--net2", p["net2"], "--total", "--s3", "--bitscores", p["sim"], "--blastsum",
            "--popsize", "12", "--generations", "4", "--hillclimbiters", "20", "--nthreads", "4",
            "--randseed", "5", "--finalstats"]
    out = run(GPU["optnetalign"], args + ["--outprefix", os.path.join(w["dir"], "g1")], w["dir"])
    assert len(out.strip().split("\t")) == 25                  # the --finalstats line (steadystate.cpp:320-337)
    n_gpu, s3_gpu = check_results(w, os.path.join(w["dir"], "g1"), {"S3", "BitscoreSum"}, True, False)
    run(CPU["optnetalign"], args + ["--outprefix", os.path.join(w["dir"], "c1")], w["dir"])
    n_cpu, s3_cpu = check_results(w, os.path.join(w["dir"], "c1"), {"S3", "BitscoreSum"}, True, False)
    # a random total alignment of this pair has S3 ~ 0.003; 2 x 10 000 hill-climb iterations per
    # child lift it by more than an order of magnitude in either build
    assert s3_gpu > 0.03 and s3_cpu > 0.03 and 0.5 < s3_gpu / s3_cpu < 2.0, (s3_gpu, s3_cpu)
    assert n_gpu >= 1 and n_cpu >= 1


@needs_build
def test_optnetalign_partial_goc(world):
    """no --total: SizeFit is appended (argParsing.h:269-271), masks are randomised, the variation
    operators flip mask bits; --annotations1/2 --goc adds the GOC objective"""
    w, p = world, world["paths"]
    args = ["--net1", p["net1"], "--net2", p["net2"], "--ics", "--ec", "--bitscores", p["sim"], "--blastsum",
            "--annotations1", p["annos1"], "--annotations2", p["annos2"], "--goc",
            "--popsize", "10", "--generations", "3", "--hillclimbiters", "6", "--nthreads", "3",
            "--mutrate", "0.5", "--cxrate", "0.9"]
    run(GPU["optnetalign"], args + ["--outprefix", os.path.join(w["dir"], "g2")], w["dir"])
    n, _ = check_results(w, os.path.join(w["dir"], "g2"), {"ICS", "EC", "BitscoreSum", "GOC", "Size"}, False, True)
    assert n >= 1


@needs_build
def test_nsgaii_driver(world):
    w, p = world, world["paths"]
    args = ["--net1", p["net1"], "--net2", p["net2"], "--total", "--s3", "--bitscores", p["sim"], "--blastsum",
            "--popsize", "16", "--generations", "3", "--hillclimbiters", "2000"]
    prefix = os.path.join(w["dir"], "g3")
    run(GPU["nsgaii"], args + ["--outprefix", prefix], w["dir"])
    # nsgaiimain.cpp:268-295: header = "filename" + the fitnessName enum values, one row per member
    # of the first front with its fitness vector
    info = [l.rstrip("\n").split("\t") for l in open(prefix + ".info")]
    assert info[0] == ["filename", "6", "2"]                  # S3Fit, BitscoreSumFit (Network.h:31-32)
    assert len(info) >= 2
    best = 0.0
    for row in info[1:]:
        lines = [l.split() for l in open(row[0])]
        assert len(lines) == w["pair"].n1 and len({v for _, v in lines}) == len(lines)
        a = ref.Aln.load(w["n1"], w["n2"], row[0], w["B"], None)
        C_seed, I = a.counters()[0], a.counters()[1]
        E1 = w["pair"].E1
        # the loader's seed and the driver's running count may each sit one below the true count (H1)
        assert row[1] in [fmt(C / (E1 + I - C)) for C in (C_seed - 1, C_seed, C_seed + 1)], (row, C_seed, I)
        assert row[2] == fmt(a.sum_blast())
        best = max(best, float(row[1]))
    assert best > 0.01                                         # 3 generations x 2000 climb iterations


@needs_build
def test_hillclimber_driver_and_cli_errors(world):
    w, p = world, world["paths"]
    prefix = os.path.join(w["dir"], "g4")
    out = run(GPU["hillclimber"], ["--net1", p["net1"], "--net2", p["net2"], "--total", "--ics",
                                   "--generations", "40", "--outprefix", prefix], w["dir"])
    assert "Generation 39 complete." in out                   # hillclimber.cpp:66-75 counts from 0
    saved = prefix + "_localTest.aln"                          # :79
    lines = [l.split() for l in open(saved)]
    assert len(lines) == w["pair"].n1
    a = ref.Aln.load(w["n1"], w["n2"], saved, None, None)
    ics_lines = [l for l in out.splitlines() if l.startswith("current ICS is")]
    reported = float(ics_lines[-1].split()[-1])
    assert abs(reported - a.ics()) <= 2.0 / max(1, a.counters()[1]) + 1e-6      # within the truncated seed (H1)
    assert a.ics() > 0.05                                                        # 20 000 iterations did climb
    # argument errors are reported like the reference reports them (caught in main, exit status 0)
    for binary in (GPU["hillclimber"], CPU["hillclimber"]):
        bad = subprocess.run([binary, "--net1", p["net1"], "--net2", p["net2"], "--total"], capture_output=True, text=True)
        assert bad.returncode == 0 and "error" in (bad.stderr + bad.stdout).lower()
        bad = subprocess.run([binary, "--net1", p["net1"], "--nonsense"], capture_output=True, text=True)
        assert bad.returncode == 0 and "nonsense" in (bad.stderr + bad.stdout)
