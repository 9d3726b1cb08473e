#!/usr/bin/env python
"""loadGOC (goc.cpp:16-52) on the C3 pair (6 000 x 8 000 nodes, ~10 terms per annotated node): the compiled
reference's host loop vs ona_goc_matrix on the GPU, values compared bit for bit.
    python profiles/goc_build.py [out.json]"""
import ctypes as C, json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from optnetaligncpp_b200 import engine, synth
from oracle import ref
pair = synth.make_pair(n1=6000, n2=8000, m1=30000, m2=30000, seed=20260002, annotations=True)
d = tempfile.mkdtemp()
paths = synth.write_text(pair, os.path.join(d, "c3"))
n1, n2 = ref.Net(paths["net1"]), ref.Net(paths["net2"])
t0 = time.perf_counter()
G = ref.Dict.load_goc(n1, n2, paths["annos1"], paths["annos2"])
t_ref = time.perf_counter() - t0
want = G.dense()
def rows(ann, n):
    rp = np.zeros(n + 1, np.int32)
    rp[1:] = np.cumsum([len(a) for a in ann])
    return rp, (np.concatenate(ann).astype(np.int32) if rp[-1] else np.zeros(0, np.int32)), np.array([len(a) > 0 for a in ann], np.uint8)
rp1, t1, p1 = rows(pair.ann1, pair.n1)
rp2, t2, p2 = rows(pair.ann2, pair.n2)
L = engine.load_library()
out = np.empty((pair.n1, pair.n2), np.float64)
ptr = lambda a: a.ctypes.data_as(C.c_void_p)
ts = []
for _ in range(3):
    t0 = time.perf_counter()
    rc = L.ona_goc_matrix(0, pair.n1, pair.n2, ptr(rp1), ptr(t1), ptr(p1), ptr(rp2), ptr(t2), ptr(p2), ptr(out))
    ts.append(time.perf_counter() - t0)
    assert rc == 0, L.ona_last_error()
same = np.array_equal(out, want, equal_nan=True)
res = dict(n1=pair.n1, n2=pair.n2, nonzero=int(np.count_nonzero(want)), reference_loadGOC_s=t_ref,
           ona_goc_matrix_s=ts, bit_identical=bool(same))
print(json.dumps(res))
assert same
if len(sys.argv) > 1:
    json.dump(res, open(os.path.join(ROOT, sys.argv[1]), "w"), indent=1)
