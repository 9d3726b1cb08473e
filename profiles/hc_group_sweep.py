#!/usr/bin/env python
"""Exact hill climb (k_hill_climb) on the C2 pair: iterations/s by population size and lanes per proposal
(ONA_GROUP = 4 / 8; unset = the engine's own choice).   python profiles/hc_group_sweep.py [out.json]"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if os.environ.get("HC_CHILD"):
    sys.path.insert(0, ROOT)
    import numpy as np, torch
    import bench
    from optnetaligncpp_b200 import engine, synth
    pair = bench.make_workload()
    e = engine.Engine.from_pair(pair, bench.OBJECTIVES)
    out = {}
    for P in [int(v) for v in os.environ.get("HC_POPS", "200,296,320,455,592,700,1024,1223,2048").split(",")]:
        aln, _ = synth.random_population(pair, P, seed=1000)
        e.resize(P); e.upload(aln); e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
        e.hill_climb(iters=500, seed=1)
        e.upload(aln); e.full_fitness(seed_mode=engine.SEED_REFCOMPAT)
        iters = 20000
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); _, nacc = e.hill_climb(iters=iters, seed=7); b.record(); torch.cuda.synchronize()
        out[P] = P * iters / (a.elapsed_time(b) * 1e-3)
    print("HC " + json.dumps(out))
    sys.exit(0)
res = {}
for grp in [g if g != "auto" else None for g in os.environ.get("HC_GROUPS", "4,8,auto").split(",")]:
    env = dict(os.environ, HC_CHILD="1")
    if grp:
        env["ONA_GROUP"] = grp
    else:
        env.pop("ONA_GROUP", None)
    o = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True)
    line = [l for l in o.stdout.splitlines() if l.startswith("HC ")]
    res[grp or "auto"] = json.loads(line[0][3:]) if line else o.stderr[-400:]
    print(grp or "auto", res[grp or "auto"], flush=True)
if len(sys.argv) > 1:
    json.dump(res, open(os.path.join(ROOT, sys.argv[1]), "w"), indent=1)
