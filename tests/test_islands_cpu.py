"""CPU: the N>1 island logic (world_size 2, gloo): population sharding, global individual ids
for the proposal streams, and the all-gather of per-individual objective rows that feeds the
host NSGA-II sort (nsga-ii.cpp:14-118).  No GPU: the objective rows are produced by the oracle."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from optnetaligncpp_b200 import engine, islands, synth  # noqa: E402


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import port as oracle_port
    pair = synth.make_pair(n1=80, n2=100, m1=300, m2=340, max_deg=20, seed=12)
    S = synth.dense_bits(pair)
    prob = oracle_port.Problem(pair.n1, pair.n2, pair.E1, pair.rp1, pair.col1, pair.rp2, pair.col2, S, None)
    names = [engine.S3, engine.BITSCORE_SUM]
    P = 10                                                    # global population
    first, count = islands.shard(P, rank, world)
    aln, _ = synth.random_population(pair, P, seed=5)         # same on every rank
    rows = np.stack([oracle_port.Aln(prob, aln[p]).compute_fitness(names) for p in range(first, first + count)])
    gathered = islands.all_gather_objectives(torch.from_numpy(rows), P, world)
    full = np.stack([oracle_port.Aln(prob, aln[p]).compute_fitness(names) for p in range(P)])
    ok = np.array_equal(gathered.numpy(), full)
    # proposal streams are keyed by the GLOBAL individual id
    mine = engine.rng_proposals_host(9, pair.n1, pair.n2, first, count, 50)
    whole = engine.rng_proposals_host(9, pair.n1, pair.n2, 0, P, 50)
    ok = ok and np.array_equal(mine, whole[first:first + count])
    out[rank] = int(ok)
    dist.destroy_process_group()


def test_shard_covers_population_without_overlap():
    for P in (1, 7, 200, 2048):
        for world in (1, 2, 3, 8):
            spans = [islands.shard(P, r, world) for r in range(world)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == P
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_two_rank_gloo_all_gather_of_objective_rows():
    world = 2
    port = 29500 + (os.getpid() % 2000)
    out = mp.Manager().dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert [out[r] for r in range(world)] == [1, 1]
