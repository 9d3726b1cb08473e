"""Small driver for profiling the hill-climb kernel: C2 pair, P individuals, `iters` proposals each."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from optnetaligncpp_b200 import engine, synth
import bench

P = int(sys.argv[1]) if len(sys.argv) > 1 else 200
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
pair = bench.make_workload()
e = engine.Engine.from_pair(pair, bench.OBJECTIVES)
aln, _ = synth.random_population(pair, P, seed=1000)
e.upload(aln)
e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
e.hill_climb(iters=500, seed=1)
t0 = time.perf_counter()
_, nacc = e.hill_climb(iters=iters, seed=2)
dt = time.perf_counter() - t0
print("P=%d iters=%d: %.3e iterations/s, accept rate %.3f" % (P, iters, P * iters / dt, nacc.sum() / (P * iters)))

e.upload(aln)
e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
e.hill_climb_batched(iters=2048, seed=1)
t0 = time.perf_counter()
napp = e.hill_climb_batched(iters=iters * 10, seed=2)
dt = time.perf_counter() - t0
print("batched: P=%d iters=%d: %.3e proposals/s, applied fraction %.3f" % (P, iters * 10, P * iters * 10 / dt, napp.sum() / (P * iters * 10)))
