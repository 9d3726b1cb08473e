"""BASELINE.json configs[4]: swap-delta throughput sweep on the yeast-fly pair,
popsize 256..16384 x candidates/individual 1k..64k (device proposals, accept-count summary only), at 1 GPU
or -- launched with torchrun, one process per GPU -- with every population split into islands over the GPUs
(strong scaling at fixed P x K; times are CUDA events between barriers, max over ranks).
    python profiles/sweep.py ../gpurun_out/sweep_1gpu.json
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29535 \
        profiles/sweep.py ../gpurun_out/sweep_8gpu.json"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from optnetaligncpp_b200 import engine, synth, islands
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
pair = bench.make_workload()
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
e = engine.Engine.from_pair(pair, bench.OBJECTIVES, device=local)
e.set_stream(stream.cuda_stream)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


rows = []
for P in (256, 1024, 4096, 16384):
    first, cnt = islands.shard(P, rank, world)
    aln, _ = synth.random_population(pair, P, seed=3)
    e.resize(cnt)
    e.set_island_offset(first)
    e.upload(aln[first:first + cnt])
    e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT)
    nimp = torch.zeros(cnt, dtype=torch.int32, device=dev)
    for K in (1024, 4096, 16384, 65536):
        if P * K > 2 ** 30:
            continue
        reps = max(3, min(20, int(2e8 * world / (P * K))))
        for w in range(2):
            e.swap_deltas_dev(K, seed=100 + w, num_improved=nimp.data_ptr())
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record(stream)
        for r in range(reps):
            e.swap_deltas_dev(K, seed=r, num_improved=nimp.data_ptr())
        ev[1].record(stream)
        barrier()
        t = torch.tensor([ev[0].elapsed_time(ev[1]) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        rows.append(dict(P=P, K=K, n_gpus=world, ms=ms, swap_deltas_per_sec=P * K / (ms * 1e-3)))
        if rank == 0:
            print("P=%6d K=%6d  %8.3f ms  %.3e swap-deltas/s" % (P, K, ms, P * K / (ms * 1e-3)), flush=True)
if rank == 0:
    out = sys.argv[1] if len(sys.argv) > 1 else "sweep_%dgpu.json" % world
    json.dump(rows, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), out), "w"), indent=1)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
