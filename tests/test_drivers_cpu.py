"""CPU: the reference's drivers built ENTIRELY from unmodified reference sources against the
Boost.ProgramOptions / TBB stand-ins of optnetaligncpp_b200/host/compat (oracle/_ref/*_ref, built in
the dev container).  This pins the stand-ins: the reference's own argParsing.h has to parse its
documented invocations (manual.tex:131-142) through them, reject what Boost rejects, and the
drivers' result files have to agree with the reference's from-scratch scoring of the alignments
they contain.  tests/test_gpu_drivers.py runs the GPU-engine builds of the same drivers through
the same checks."""
import os
import subprocess
import tempfile

import pytest

from optnetaligncpp_b200 import synth
from oracle import ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CPU = {n: os.path.join(ROOT, "oracle", "_ref", n + "_ref") for n in ("optnetalign", "nsgaii", "hillclimber")}
pytestmark = pytest.mark.skipif(not (all(os.path.exists(p) for p in CPU.values()) and ref.available()),
                                reason="reference drivers not built (dev container only)")


@pytest.fixture(scope="module")
def world():
    pair = synth.make_pair(n1=150, n2=200, m1=600, m2=700, max_deg=30, seed=9, annotations=True,
                           n_terms=80, bit_density=0.02)
    d = tempfile.mkdtemp()
    paths = synth.write_text(pair, os.path.join(d, "w"))
    n1, n2 = ref.Net(paths["net1"]), ref.Net(paths["net2"])
    return dict(pair=pair, dir=d, p=paths, n1=n1, n2=n2, B=ref.Dict.load_blast(n1, n2, paths["sim"]))


def run(binary, args):
    return subprocess.run([binary] + args, capture_output=True, text=True, timeout=300)


def test_reference_cli_through_the_standins(world):
    w, p = world, world["p"]
    prefix = os.path.join(w["dir"], "o")
    out = run(CPU["optnetalign"], ["--net1", p["net1"], "--net2=" + p["net2"], "--total", "--s3", "--bitscores", p["sim"],
                                   "--blastsum", "--popsize", "8", "--generations", "3", "--hillclimbiters", "4",
                                   "--nthreads", "2", "--outprefix", prefix, "--verbose"])
    assert out.returncode == 0 and "Writing to disk!" in out.stdout, out.stderr
    info = [l.rstrip("\n").split("\t") for l in open(prefix + ".info")]
    assert info[0][:4] == ["filename", "EC", "ICS", "S3*"] and info[0][5] == "BitscoreSum*"
    E1 = w["pair"].E1
    for row in info[1:]:
        a = ref.Aln.load(w["n1"], w["n2"], row[0], w["B"], None)
        C, I = a.counters()[0], a.counters()[1]
        assert row[1] in ["%g" % (c / E1) for c in (C - 1, C, C + 1)]          # truncated seeds (H1)
        assert row[5] == "%g" % a.sum_blast() and row[7] == "%g" % w["pair"].n1
    # what Boost.ProgramOptions rejects: unknown options, missing / malformed arguments, positionals
    for bad in (["--net1", p["net1"], "--nonsense"], ["--net1"], ["--net1", p["net1"], "--popsize", "ten"],
                ["stray"], ["--total=1", "--net1", p["net1"]]):
        r = run(CPU["optnetalign"], bad)
        assert r.returncode == 0 and "error" in (r.stderr + r.stdout).lower(), bad     # caught in main, exit 0
    # argParsing.h's own validation still fires behind the parser (argParsing.h:153-200, 273-286)
    r = run(CPU["optnetalign"], ["--net1", p["net1"], "--net2", p["net2"], "--total", "--generations", "1",
                                 "--outprefix", prefix])
    assert "error" in (r.stderr + r.stdout).lower()                             # no objective given
    r = run(CPU["hillclimber"], ["--help"])
    assert "--hillclimbiters arg" in r.stdout and "--total" in r.stdout


def test_nsgaii_and_hillclimber_reference_builds_run(world):
    w, p = world, world["p"]
    out = run(CPU["nsgaii"], ["--net1", p["net1"], "--net2", p["net2"], "--total", "--ics", "--popsize", "6",
                              "--generations", "2", "--hillclimbiters", "50", "--outprefix", os.path.join(w["dir"], "n")])
    assert out.returncode == 0 and os.path.exists(os.path.join(w["dir"], "n.info")), out.stderr
    out = run(CPU["hillclimber"], ["--net1", p["net1"], "--net2", p["net2"], "--total", "--ics", "--generations", "3",
                                   "--outprefix", os.path.join(w["dir"], "h")])
    assert out.returncode == 0 and "Generation 2 complete." in out.stdout
    assert len(open(os.path.join(w["dir"], "h_localTest.aln")).readlines()) == w["pair"].n1
