"""CPU: the C-ABI library loads, exports every symbol include/ona_b200.h declares, and
refuses to run without a GPU (no CPU fallback).  No compute calls here."""
import os
import re

import numpy as np
import pytest

from optnetaligncpp_b200 import engine, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ona_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ona_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = engine.load_library()
    syms = header_symbols()
    assert len(syms) >= 20
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    assert sorted(engine.EXPORTS) == syms


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    pair = synth.make_pair(n1=40, n2=50, m1=120, m2=140, max_deg=12, seed=3)
    with pytest.raises(engine.OnaError, match="no CPU fallback"):
        engine.Engine.from_pair(pair, [engine.S3, engine.BITSCORE_SUM])


def test_host_rng_mirror_supports():
    xy = engine.rng_proposals_host(seed=5, n1=60, n2=80, first_global=3, count=4, K=5000)
    x, y = xy[..., 0], xy[..., 1]
    assert x.min() >= 0 and x.max() < 60 and y.min() >= 0 and y.max() < 80
    assert (x != y).all()
    assert len(np.unique(x)) == 60 and len(np.unique(y)) == 80
    # streams are keyed by the GLOBAL individual index (island sharding keeps them)
    a = engine.rng_proposals_host(5, 60, 80, first_global=4, count=1, K=100)
    assert np.array_equal(a[0], xy[1, :100])
    b = engine.rng_proposals_host(5, 60, 80, first_global=3, count=1, K=50, start=50)
    assert np.array_equal(b[0], xy[0, 50:100])


def test_synth_pair_is_reference_loadable_shape():
    pair = synth.make_pair(n1=200, n2=260, m1=900, m2=1000, max_deg=40, seed=11)
    for n, rp, col, edges in ((pair.n1, pair.rp1, pair.col1, pair.edges1),
                              (pair.n2, pair.rp2, pair.col2, pair.edges2)):
        assert rp[0] == 0 and rp[-1] == len(col) and (np.diff(rp) >= 1).all()
        # no duplicate / reversed duplicate lines
        key = np.minimum(edges[:, 0], edges[:, 1]).astype(np.int64) * n + np.maximum(edges[:, 0], edges[:, 1])
        assert len(np.unique(key)) == len(edges)
        # first-appearance numbering (Network.cpp:59-83)
        flat = edges.reshape(-1)
        _, first = np.unique(flat, return_index=True)
        assert (np.diff(first) > 0).all()
    assert (edges[:, 0] == edges[:, 1]).sum() > 0      # self-loops are exercised
