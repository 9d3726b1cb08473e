import sys, os, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from optnetaligncpp_b200 import engine, synth
import bench
pair = bench.make_workload()
P, K = 200, 16384
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
e = engine.Engine.from_pair(pair, bench.OBJECTIVES); e.set_stream(stream.cuda_stream)
aln, _ = synth.random_population(pair, P, seed=1000); e.upload(aln); e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
xy = torch.from_numpy(synth.random_candidates(pair, P, K, seed=50)).pin_memory()
imp = torch.empty((P, K), dtype=torch.uint8).pin_memory()
snd = torch.empty((P, K), dtype=torch.float64).pin_memory()
imp2 = torch.empty((P, K), dtype=torch.uint8).pin_memory()
snd2 = torch.empty((P, K), dtype=torch.float64).pin_memory()
# reference result via the copy pipeline
out = dict(improved=imp2.numpy(), snd=snd2.numpy())
e.swap_deltas(xy.numpy(), obj=-1, out=out)
def zc():
    e.swap_deltas_dev(K, obj=-1, xy_dev=xy.data_ptr(), improved=imp.data_ptr(), snd=snd.data_ptr())
    e.synchronize()
for _ in range(3): zc()
t0 = time.perf_counter()
for _ in range(10): zc()
dt = (time.perf_counter() - t0) / 10
print("zero-copy: %.3f ms  %.3e /s" % (dt * 1e3, P * K / dt))
print("equal:", torch.equal(imp, imp2), torch.equal(torch.nan_to_num(snd), torch.nan_to_num(snd2)))
t0 = time.perf_counter()
for _ in range(10): e.swap_deltas(xy.numpy(), obj=-1, out=out)
dt = (time.perf_counter() - t0) / 10
print("pipelined copies: %.3f ms  %.3e /s" % (dt * 1e3, P * K / dt))
# hybrid: zero-copy input, D2H copy of outputs
