#!/usr/bin/env python
"""From-scratch evaluations per second by population size on the C2 pair (k_full_fitness_split from 32
individuals on, k_full_fitness below).   python profiles/ff_sweep.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from optnetaligncpp_b200 import engine, synth
pair = bench.make_workload()
e = engine.Engine.from_pair(pair, bench.OBJECTIVES)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); e.set_stream(stream.cuda_stream)
res = {}
for P in (1, 8, 31, 32, 100, 200, 222, 400, 444, 1024, 4096, 16384):
    aln, _ = synth.random_population(pair, P, seed=5)
    e.resize(P); e.upload(aln)
    for _ in range(3):
        e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT)
    torch.cuda.synchronize()
    reps = 20 if P <= 1024 else 5
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT)
    b.record(stream); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    res[P] = dict(ms=ms, evals_per_sec=P / (ms * 1e-3))
    print("P=%6d  %8.3f ms  %.3e evals/s" % (P, ms, P / (ms * 1e-3)), flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(os.path.join(ROOT, sys.argv[1]), "w"), indent=1)
