"""Hill-climb throughput late in a climb (after `pre` proposals per individual), exact vs batched."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from optnetaligncpp_b200 import engine, synth
import bench

P = int(sys.argv[1]) if len(sys.argv) > 1 else 200
pre = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
iters = 100000
pair = bench.make_workload()
e = engine.Engine.from_pair(pair, bench.OBJECTIVES)
aln, _ = synth.random_population(pair, P, seed=1000)
e.upload(aln)
e.full_fitness(seed_mode=engine.SEED_TRUE)      # exact seed, so incremental counters can be checked
f0 = e.counters()["fit"].mean(0)
t0 = time.perf_counter(); napp = e.hill_climb_batched(iters=pre, seed=1); dt = time.perf_counter() - t0
print("pre-climb (batched) %d proposals/individual: %.3e proposals/s, applied %.3f" % (pre, P * pre / dt, napp.sum() / (P * pre)))
f1 = e.counters()["fit"].mean(0)
print("mean objectives before", f0, "after", f1)
t0 = time.perf_counter(); _, nacc = e.hill_climb(iters=iters, seed=2); dt = time.perf_counter() - t0
print("exact   late: %.3e iterations/s, accept rate %.3f" % (P * iters / dt, nacc.sum() / (P * iters)))
t0 = time.perf_counter(); napp = e.hill_climb_batched(iters=iters * 4, seed=3); dt = time.perf_counter() - t0
print("batched late: %.3e proposals/s, applied %.3f" % (P * iters * 4 / dt, napp.sum() / (P * iters * 4)))
c = e.counters(); s = e.full_fitness(seed_mode=engine.SEED_NONE)
print("incremental == from scratch:", bool((c["cons"] == s["cons_true"]).all()), bool((c["ind"] == s["ind"]).all()), bool((c["bits"] == s["bits"]).all()),
      "max |dcons|", int(abs(c["cons"] - s["cons_true"]).max()))
