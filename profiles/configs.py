#!/usr/bin/env python
"""Throughput of the hot path on the other BASELINE.json shapes (1 GPU): configs[2] (yeast-fly pair
+ GOC objective, 256 individuals = one island of the 2048 / 8 GPUs, 5 M-iteration climbs sampled by
50 000) and configs[3] (human-scale pair, ~20k nodes / ~200k edges, --ics --blastsum, P = 1024).
    python profiles/configs.py out.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from optnetaligncpp_b200 import engine, synth

dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
res = {}
def timed(fn):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); r = fn(); b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b), r
for tag, mk, names, P, K in (
        ("c3_yeast_fly_goc", lambda: synth.make_pair(n1=6000, n2=8000, m1=30000, m2=30000, seed=20260002, annotations=True),
         [engine.S3, engine.BITSCORE_SUM, engine.GOC], 256, 16384),
        ("c4_human_scale", lambda: synth.make_pair(n1=20000, n2=20000, m1=200000, m2=200000, seed=20260003),
         [engine.ICS, engine.BITSCORE_SUM], 1024, 4096)):
    pair = mk()
    use_goc = engine.GOC in names
    e = engine.Engine(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, names,
                      bits=(pair.bit_rows, pair.bit_cols, pair.bit_vals),
                      goc=dict(ann1=pair.ann1, ann2=pair.ann2) if use_goc else None)   # Jaccard table built on the device
    e.set_stream(stream.cuda_stream)
    aln, _ = synth.random_population(pair, P, seed=7)
    e.upload(aln)
    ms_ff, _ = timed(lambda: e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT))
    ms_ff, _ = timed(lambda: e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT))
    nimp = torch.zeros(P, dtype=torch.int32, device=dev)
    e.swap_deltas_dev(K, seed=1, num_improved=nimp.data_ptr())
    ms_sd, _ = timed(lambda: [e.swap_deltas_dev(K, seed=2 + i, num_improved=nimp.data_ptr()) for i in range(5)])
    iters = 50000
    e.hill_climb(iters=1000, seed=3)
    ms_hc, (_, nacc) = timed(lambda: e.hill_climb(iters=iters, seed=4))
    c = e.counters()
    r = e.full_fitness(seed_mode=engine.SEED_NONE)
    assert np.array_equal(c["cons"] - r["cons_true"], (c["cons"] - r["cons_true"]).clip(-1, 0))   # seed truncation only (H1)
    assert np.array_equal(c["ind"], r["ind"])
    res[tag] = dict(n1=pair.n1, n2=pair.n2, P=P, K=K, objectives=[int(n) for n in names],
                    full_fitness_evals_per_sec=P / (ms_ff * 1e-3),
                    swap_deltas_per_sec=5 * P * K / (ms_sd * 1e-3),
                    hill_climb_iters_per_sec=P * iters / (ms_hc * 1e-3), accept_rate=float(nacc.sum()) / (P * iters))
    print(tag, res[tag], flush=True)
    e.close()
json.dump(res, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), sys.argv[1] if len(sys.argv) > 1 else "configs_1gpu.json"), "w"), indent=1)
