#!/usr/bin/env python
"""One traced host-buffer swap-delta call (ONA_E2E_TRACE=1 ONA_E2E_GRAPH=0 python profiles/e2e_trace.py)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from optnetaligncpp_b200 import engine, synth
pair = bench.make_workload()
P, K = bench.POP_PER_GPU, bench.CAND_PER_IND
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
e = engine.Engine.from_pair(pair, bench.OBJECTIVES); e.set_stream(stream.cuda_stream)
aln, _ = synth.random_population(pair, P, seed=1000)
e.upload(aln); e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
xy = torch.from_numpy(synth.random_candidates(pair, P, K, seed=50)).pin_memory()
out = dict(improved=torch.empty((P, K), dtype=torch.uint8).pin_memory().numpy(),
           snd=torch.empty((P, K), dtype=torch.float64).pin_memory().numpy())
for i in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    e.swap_deltas(xy.numpy(), obj=-1, out=out)
