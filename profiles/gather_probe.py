#!/usr/bin/env python
"""Scattered-sector read rate of the memory system for tables of several sizes (ona_probe_gather_rate):
the ceiling of the engine's gather-bound kernels.   python profiles/gather_probe.py [out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from optnetaligncpp_b200 import engine, synth
pair = synth.make_pair(n1=300, n2=400, m1=1500, m2=1600, max_deg=60, seed=5)
e = engine.Engine.from_pair(pair, [engine.S3])
res = {}
for mb in (1, 8, 32, 64, 96, 256, 1024):
    r, ms = e.probe_gather_rate(mb << 20)
    res["%d MiB" % mb] = {"sectors_per_s": r, "GB_per_s_of_32B_sectors": r * 32 / 1e9, "ms": ms}
    print("%5d MiB table: %.3e sectors/s = %7.1f GB/s of 32-byte sectors (%.3f ms)" % (mb, r, r * 32 / 1e9, ms))
if len(sys.argv) > 1:
    json.dump(res, open(os.path.join(ROOT, sys.argv[1]), "w"), indent=1)
