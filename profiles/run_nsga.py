"""Small driver for profiling the NSGA-II ranking kernels: rows x objectives of continuous or
counter-ratio (tie-heavy) values, host API (ona_nsga2_rank)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import nsga_cases
from optnetaligncpp_b200 import engine, synth

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
M = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "continuous"
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
pair = synth.make_pair(n1=60, n2=80, m1=150, m2=170, max_deg=12, seed=3)
e = engine.Engine.from_pair(pair, [1, 6])
f = nsga_cases.make_case(kind, P, M, 11)
r = e.nsga2_rank(f)
t0 = time.perf_counter()
for _ in range(reps):
    r = e.nsga2_rank(f)
dt = (time.perf_counter() - t0) / reps
sizes = r["front_offsets"][1:] - r["front_offsets"][:-1]
print("%s P=%d M=%d: %d fronts (largest %d), %.3f ms per call incl. host copies" % (
    kind, P, M, len(sizes), sizes.max(), dt * 1e3))
