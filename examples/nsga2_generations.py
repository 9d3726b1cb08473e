"""A few NSGA-II generations on one GPU with nothing but the engine's batched entry points --
the shape of nsgaiimain.cpp:142-238 (R_t = parents + children -> non-dominated sort -> crowding
-> next parents; children = tournament copies + mutation + hill climb), with the population
resident in the engine's slots.  Not a port of the reference's drivers (their PRNG seeds come
from clock(), so no two runs of the reference agree either): it shows how the calls compose
and is what tests/test_gpu_generation.py runs.

    python examples/nsga2_generations.py [popsize] [generations] [hillclimb iterations]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from optnetaligncpp_b200 import engine, synth  # noqa: E402


def next_parents(rank, popsize):
    """fill the new population front by front, the split front by descending crowding distance
    (nsgaiimain.cpp:161-177)"""
    offs, after, crowd = rank["front_offsets"], rank["order_after"], rank["crowd_dist"]
    chosen = []
    for f in range(len(offs) - 1):
        members = after[offs[f]:offs[f + 1]]
        if len(chosen) + len(members) <= popsize:
            chosen.extend(members)
        else:
            by_crowd = sorted(members, key=lambda i: -crowd[i])       # crowdedComp inside one front
            chosen.extend(by_crowd[:popsize - len(chosen)])
        if len(chosen) == popsize:
            break
    return np.array(chosen, np.int64)


def run(pair, names, popsize=64, generations=4, hillclimb_iters=2000, mutswappb=0.005, seed=1, log=None):
    rng = np.random.default_rng(seed)
    e = engine.Engine.from_pair(pair, names)
    P = popsize
    aln, _ = synth.random_population(pair, 2 * P, seed=seed)            # slots [0,P) parents, [P,2P) children
    e.upload(aln)
    e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
    e.hill_climb(iters=hillclimb_iters, seed=seed, first=0, count=2 * P)
    history = []
    for gen in range(generations):
        fit = e.counters()["fit"]                                       # [2P, M] objective rows
        rank = e.nsga2_rank(fit)
        parents = next_parents(rank, P)
        history.append(fit[parents].mean(axis=0))
        if log:
            log("generation %d: %d fronts, parents' mean objectives %s" % (
                gen, len(rank["front_offsets"]) - 1, np.array2string(history[-1], precision=5)))
        # children: binary tournaments on (rank, crowding), copies of the winners, mutation swaps
        # (host-drawn ordered lists, Alignment.cpp:451-466), then the device hill climb
        a, _, _ = e.download()
        cand = rng.integers(0, P, (P, 2))
        key = lambda i: (rank["dom_rank"][parents[i]], -rank["crowd_dist"][parents[i]])
        winners = np.array([c[0] if key(c[0]) <= key(c[1]) else c[1] for c in cand])
        e.upload(np.concatenate([a[parents], a[parents[winners]]]))
        e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
        nswaps = max(1, int(mutswappb * pair.n2))
        swaps = np.stack([rng.integers(0, pair.n2, (P, nswaps)), rng.integers(0, pair.n2, (P, nswaps))], -1)
        e.apply_swaps(swaps.astype(np.int32), first=P)
        e.hill_climb(iters=hillclimb_iters, seed=seed + 100 + gen, first=P, count=P)
    fit = e.counters()["fit"]
    rank = e.nsga2_rank(fit)
    parents = next_parents(rank, P)
    history.append(fit[parents].mean(axis=0))
    return e, parents, np.array(history)


if __name__ == "__main__":
    popsize = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    gens = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5000
    pair = synth.make_pair(n1=600, n2=800, m1=3000, m2=3200, max_deg=80, seed=123, bit_density=0.005)
    run(pair, [engine.S3, engine.BITSCORE_SUM], popsize, gens, iters, log=print)
