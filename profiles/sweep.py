"""BASELINE.json configs[4]: swap-delta throughput sweep on the yeast-fly pair,
popsize 256..16384 x candidates/individual 1k..64k (device proposals, accept-count summary only)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from optnetaligncpp_b200 import engine, synth
import bench

pair = bench.make_workload()
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
e = engine.Engine.from_pair(pair, bench.OBJECTIVES)
e.set_stream(stream.cuda_stream)
rows = []
for P in (256, 1024, 4096, 16384):
    aln, _ = synth.random_population(pair, P, seed=3)
    e.resize(P)
    e.upload(aln)
    e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT)
    nimp = torch.zeros(P, dtype=torch.int32, device=dev)
    for K in (1024, 4096, 16384, 65536):
        if P * K > 2 ** 30:
            continue
        reps = max(3, min(20, int(2e8 / (P * K))))
        for w in range(2):
            e.swap_deltas_dev(K, seed=100 + w, num_improved=nimp.data_ptr())
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(stream)
        for r in range(reps):
            e.swap_deltas_dev(K, seed=r, num_improved=nimp.data_ptr())
        ev[1].record(stream)
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / reps
        rows.append(dict(P=P, K=K, ms=ms, swap_deltas_per_sec=P * K / (ms * 1e-3)))
        print("P=%6d K=%6d  %8.3f ms  %.3e swap-deltas/s" % (P, K, ms, P * K / (ms * 1e-3)), flush=True)
out = sys.argv[1] if len(sys.argv) > 1 else "sweep_1gpu.json"
json.dump(rows, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), out), "w"), indent=1)
