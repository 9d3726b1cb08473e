import sys, numpy as np
sys.path.insert(0, '.')
from optnetaligncpp_b200 import engine, synth
from oracle import port
pair = synth.make_pair(n1=60, n2=80, m1=200, m2=260, max_deg=20, seed=4)
e = engine.Engine.from_pair(pair, [engine.S3])
rng = np.random.default_rng(99)
bad = 0
for t in range(300):
    kind = t % 3
    if kind == 0:
        n2 = int(rng.integers(2, 40000)); n1 = int(rng.integers(1, n2 + 1))
    elif kind == 1:
        n2 = int(rng.integers(2**30, 2**31 - 1)); n1 = int(rng.integers(2**29, n2))
    else:
        n2 = int(rng.integers(2, 12)); n1 = int(rng.integers(1, n2 + 1))
    cnt = int(rng.integers(1, 6)); iters = int(rng.integers(1, 3000))
    states = np.stack([port.mt_state(int(s)) for s in rng.integers(0, 2**32, cnt)])
    for p in range(cnt):
        states[p] = port.mt_xy_stream(states[p], 7, 9, int(rng.integers(0, 1000)))[1]
    xy, after = e.mt_proposals(n1, n2, iters, states)
    for p in range(cnt):
        xo, so = port.mt_xy_stream(states[p], n1, n2, iters)
        if not (np.array_equal(xy[p], xo) and np.array_equal(after[p], so)):
            bad += 1; print("MISMATCH", t, n1, n2, iters, p)
print("mt stress done, mismatches:", bad)
