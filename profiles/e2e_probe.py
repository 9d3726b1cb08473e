"""End-to-end probe of ona_swap_deltas with pinned host buffers (C2 workload, one B200):
time per call for the outputs named on the command line (default: improved,snd)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from optnetaligncpp_b200 import engine, synth

outs = (sys.argv[1] if len(sys.argv) > 1 else "improved,snd").split(",")
pair = bench.make_workload()
P, K = 200, 16384
stream = torch.cuda.Stream()
e = engine.Engine.from_pair(pair, bench.OBJECTIVES)
e.set_stream(stream.cuda_stream)
aln, _ = synth.random_population(pair, P, seed=1000)
e.upload(aln)
e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT)
xy_host = torch.from_numpy(synth.random_candidates(pair, P, K, seed=50)).pin_memory()
dt = dict(improved=torch.uint8, snd=torch.float64, cons=torch.int32, ind=torch.int32)
h_out = {k: torch.empty((P, K), dtype=dt[k]).pin_memory().numpy() for k in outs if k}
for _ in range(3):
    e.swap_deltas(xy_host.numpy(), obj=-1, out=h_out)
t0 = time.perf_counter()
for _ in range(20):
    e.swap_deltas(xy_host.numpy(), obj=-1, out=h_out)
dt_ = (time.perf_counter() - t0) / 20
print("outputs=%s streamed=%s slices=%s/%s: %.3f ms per call -> %.3e swap-deltas/s" % (
    ",".join(outs), "no" if os.environ.get("ONA_NO_STREAMED") else "yes", os.environ.get("ONA_STREAM_SLICES", "-"),
    os.environ.get("ONA_PIPE_SLICES", "-"), dt_ * 1e3, P * K / dt_))
