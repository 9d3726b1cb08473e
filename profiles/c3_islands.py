#!/usr/bin/env python
"""BASELINE.json configs[2] as written: the yeast-fly pair with the GOC objective, popsize 2048 split into
islands over the GPUs of one box (256 per GPU at 8), hillclimbiters 10000 (5 M-iteration climbs, sampled
by CLIMB_ITERS per individual).  One process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29533 profiles/c3_islands.py gpurun_out/c3_8gpu.json

Per generation-shaped round every island evaluates its individuals, climbs them (exact correctHillClimb
semantics, device proposals keyed by the GLOBAL individual id), scores K candidate swaps per individual,
exports its objective rows; one NCCL all-gather makes the [2048, 3] matrix on every rank, which ranks it
in place (ona_nsga2_rank_dev).  Times are CUDA events, max over ranks; rank 0 prints / writes one JSON."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from optnetaligncpp_b200 import engine, synth, islands

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
P_TOTAL = int(os.environ.get("C3_POP", "2048")); K = 16384; CLIMB_ITERS = int(os.environ.get("CLIMB_ITERS", "50000"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
first, P = islands.shard(P_TOTAL, rank, world)
pair = synth.make_pair(n1=6000, n2=8000, m1=30000, m2=30000, seed=20260002, annotations=True)
names = [engine.S3, engine.BITSCORE_SUM, engine.GOC]
e = engine.Engine(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, names,
                  bits=(pair.bit_rows, pair.bit_cols, pair.bit_vals), goc=dict(ann1=pair.ann1, ann2=pair.ann2),
                  device=local)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); e.set_stream(stream.cuda_stream)
e.set_island_offset(first)
aln, _ = synth.random_population(pair, P_TOTAL, seed=7)
e.upload(aln[first:first + P])
M = len(names)
fit_local = torch.empty((P, M), dtype=torch.float64, device=dev)
nimp = torch.zeros(P, dtype=torch.int32, device=dev)

def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()

def timed(fn):
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream); r = fn(); b.record(stream)
    barrier()
    t = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()), r

def exchange_and_rank():
    e.export_fitness_dev(fit_local.data_ptr())
    allrows = islands.all_gather_objectives(fit_local, P_TOTAL, world)
    return e.nsga2_rank_dev(P_TOTAL, M, allrows.data_ptr()), allrows

e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT)                       # warm-up of every stage
e.hill_climb(iters=1000, seed=3); e.swap_deltas_dev(K, seed=1, num_improved=nimp.data_ptr()); exchange_and_rank()
e.upload(aln[first:first + P])
ms_ff, _ = timed(lambda: e.full_fitness_dev(seed_mode=engine.SEED_REFCOMPAT))
ms_hc, (_, nacc) = timed(lambda: e.hill_climb(iters=CLIMB_ITERS, seed=4))
ms_ff2, _ = timed(lambda: e.full_fitness_dev(seed_mode=engine.SEED_NONE))   # leaves the node counts of the climbed state
ms_sd, _ = timed(lambda: [e.swap_deltas_dev(K, seed=10 + i, num_improved=nimp.data_ptr()) for i in range(5)])
ms_x, (nfronts, allrows) = timed(exchange_and_rank)
# invariants: incremental counters of the climbed state equal a from-scratch count (up to the reference's seed
# truncation, H1), and every rank holds the same gathered matrix
c = e.counters(); r = e.full_fitness(seed_mode=engine.SEED_NONE)
d = c["cons"] - r["cons_true"]
assert np.array_equal(d, d.clip(-1, 0)) and np.array_equal(c["ind"], r["ind"])
chk = allrows.sum(0)
if world > 1:
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    assert torch.equal(lo, hi)
acc = torch.tensor([float(nacc.sum())], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(acc)
if rank == 0:
    res = dict(config="BASELINE configs[2]: yeast-fly + GOC, popsize %d as %d islands" % (P_TOTAL, world), n_gpus=world,
               individuals_per_gpu=P, objectives=["S3", "BitscoreSum", "GOC"], climb_iters_per_individual=CLIMB_ITERS,
               full_fitness_evals_per_sec=P_TOTAL / (ms_ff * 1e-3),
               hill_climb_iters_per_sec=P_TOTAL * CLIMB_ITERS / (ms_hc * 1e-3), accept_rate=float(acc.item()) / (P_TOTAL * CLIMB_ITERS),
               swap_deltas_per_sec=5.0 * P_TOTAL * K / (ms_sd * 1e-3),
               exchange_and_rank_ms=ms_x, fronts=int(nfronts), rows_ranked=P_TOTAL,
               ms=dict(full_fitness=ms_ff, hill_climb=ms_hc, swap_deltas_x5=ms_sd))
    print(json.dumps(res), flush=True)
    if len(sys.argv) > 1:
        json.dump(res, open(sys.argv[1], "w"), indent=1)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
