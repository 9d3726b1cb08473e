import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from optnetaligncpp_b200 import engine, synth
pair = bench.make_workload()
P, K = 200, 16384
dev = torch.device("cuda:0")
stream = torch.cuda.Stream()
e = engine.Engine.from_pair(pair, bench.OBJECTIVES)
e.set_stream(stream.cuda_stream)
aln, _ = synth.random_population(pair, P, seed=1000)
e.upload(aln); e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT)
o_imp = torch.empty((P, K), dtype=torch.uint8, device=dev)
o_snd = torch.empty((P, K), dtype=torch.float64, device=dev)
xy = torch.from_numpy(synth.random_candidates(pair, P, K, seed=50)).to(dev)
for S in (1, 2, 4, 8, 16):
    per = P // S
    def run():
        for s in range(S):
            f = s * per
            e.swap_deltas_dev(K, first=f, count=per, xy_dev=xy.data_ptr() + f * K * 8,
                              improved=o_imp.data_ptr() + f * K, snd=o_snd.data_ptr() + f * K * 8)
    for _ in range(3): run()
    e.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        a.record(stream)
        for _ in range(10): run()
        b.record(stream)
    e.synchronize()
    print("S=%2d slices back to back, device-resident: %.3f ms" % (S, a.elapsed_time(b) / 10))
