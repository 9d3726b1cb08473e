#!/usr/bin/env python
"""Where the host-buffer swap-delta call (bench.py's e2e leg) spends its time: the same call with
different subsets of the per-candidate outputs copied back, and the copies alone."""
import os, sys, time
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
outs = dict(improved=torch.empty((P, K), dtype=torch.uint8).pin_memory().numpy(),
            snd=torch.empty((P, K), dtype=torch.float64).pin_memory().numpy(),
            cons=torch.empty((P, K), dtype=torch.int32).pin_memory().numpy())
def run(want, n=10):
    o = {k: outs[k] for k in want}
    for _ in range(3): e.swap_deltas(xy.numpy(), obj=-1, out=o, want=want)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): e.swap_deltas(xy.numpy(), obj=-1, out=o, want=want)
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for want in (("improved", "snd"), ("snd",), ("improved",), ("cons",), ()):
    print("outputs %-22s %.3f ms" % (",".join(want) or "(none)", run(want)))
d_xy = torch.empty((P, K, 2), dtype=torch.int32, device="cuda")
d_snd = torch.empty((P, K), dtype=torch.float64, device="cuda")
h_snd = torch.from_numpy(outs["snd"])
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("H2D 26 MB alone           %.3f ms" % t(lambda: d_xy.copy_(xy, non_blocking=True)))
print("D2H 26 MB (snd) alone     %.3f ms" % t(lambda: h_snd.copy_(d_snd, non_blocking=True)))
