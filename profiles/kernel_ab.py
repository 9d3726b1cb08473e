#!/usr/bin/env python
"""A/B timing of prebuilt engine variants on the bench workload (C2 pair, 200 x 16384).

    python profiles/kernel_ab.py lib_a.so lib_b.so ...      (paths relative to the repo root)

Each variant runs in its own process (ONA_LIB_PATH); prints the kernel times of the swap-delta
launch with the bench's outputs (L2 flushed in between), full fitness and the exact hill climb.
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def child():
    sys.path.insert(0, ROOT)
    import numpy as np
    import torch
    import bench
    from optnetaligncpp_b200 import engine, synth
    pair = bench.make_workload()
    P, K, M = bench.POP_PER_GPU, bench.CAND_PER_IND, len(bench.OBJECTIVES)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    e = engine.Engine.from_pair(pair, bench.OBJECTIVES)
    e.set_stream(stream.cuda_stream)
    aln, _ = synth.random_population(pair, P, seed=1000)
    e.upload(aln)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    o_cons = torch.empty((P, K), dtype=torch.int32, device=dev)
    o_ind = torch.empty((P, K), dtype=torch.int32, device=dev)
    o_fit = torch.empty((P, K, M), dtype=torch.float64, device=dev)
    o_imp = torch.empty((P, K), dtype=torch.uint8, device=dev)
    o_nimp = torch.zeros((P,), dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    res = {}

    noflush = bool(os.environ.get("ONA_AB_NOFLUSH"))   # warm-L2 timing, to see what the cold misses cost

    def timed(fn, n, do_flush=True):
        do_flush = do_flush and not noflush
        ts = []
        for i in range(n + 3):
            if do_flush:
                flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            fn(i)
            b.record(stream)
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(a.elapsed_time(b))
        return float(np.median(ts)), float(np.min(ts))

    res["swap_deltas_ms"] = timed(lambda i: e.swap_deltas_dev(
        K, seed=100 + i, obj=-1, cons=o_cons.data_ptr(), ind=o_ind.data_ptr(), fit=o_fit.data_ptr(),
        improved=o_imp.data_ptr(), num_improved=o_nimp.data_ptr()), 20)
    res["swap_deltas_count_only_ms"] = timed(lambda i: e.swap_deltas_dev(
        K, seed=100 + i, obj=-1, num_improved=o_nimp.data_ptr()), 20)
    res["nimp_sum"] = int(o_nimp.sum().item())
    # end-to-end leg as bench.py times it: pinned host proposals in, accept bit + snd out
    import time
    xy_host = torch.from_numpy(synth.random_candidates(pair, P, K, seed=50)).pin_memory()
    h_out = dict(improved=torch.empty((P, K), dtype=torch.uint8).pin_memory().numpy(),
                 snd=torch.empty((P, K), dtype=torch.float64).pin_memory().numpy())
    for i in range(3):
        e.swap_deltas(xy_host.numpy(), obj=-1, out=h_out)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(10):
        e.swap_deltas(xy_host.numpy(), obj=-1, out=h_out)
    torch.cuda.synchronize()
    res["e2e_ms"] = (time.perf_counter() - t0) * 100.0
    res["e2e_checksum"] = [int(h_out["improved"].sum()), float(np.nan_to_num(h_out["snd"], nan=0.0, posinf=0.0).sum())]
    res["full_fitness_ms"] = timed(lambda i: e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT), 20)
    iters = 20000
    e.upload(aln)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e.hill_climb(iters=200, seed=1)
    e.upload(aln)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    a.record(stream)
    _, nacc = e.hill_climb(iters=iters, seed=7)
    b.record(stream)
    torch.cuda.synchronize()
    res["hill_climb_ms"] = a.elapsed_time(b)
    res["hill_climb_acc"] = int(nacc.sum())
    c = e.counters()
    res["checksum"] = [int(c["cons"].sum()), int(c["ind"].sum()), float(c["bits"].sum())]
    print("AB " + json.dumps(res))


def main():
    if os.environ.get("ONA_AB_CHILD"):
        return child()
    for lib in sys.argv[1:]:
        extra = {}
        if "@" in lib:                      # lib.so@VAR=value,VAR2=value2
            lib, _, kv = lib.partition("@")
            extra = dict(item.split("=", 1) for item in kv.split(","))
        env = dict(os.environ, ONA_AB_CHILD="1", ONA_LIB_PATH=os.path.join(ROOT, lib), **extra)
        out = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("AB ")]
        print("%-40s %s %s" % (lib, extra or "", line[0][3:] if line else "FAILED: " + out.stderr[-600:]))


if __name__ == "__main__":
    main()
