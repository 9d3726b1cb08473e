#!/usr/bin/env python
"""Does PCIe run both directions at once for SLICED pinned copies on two streams?"""
import time, torch
P, K = 200, 16384
xy = torch.empty((P, K, 2), dtype=torch.int32).pin_memory()
snd = torch.empty((P, K), dtype=torch.float64).pin_memory()
d_xy = torch.empty((P, K, 2), dtype=torch.int32, device="cuda")
d_snd = torch.empty((P, K), dtype=torch.float64, device="cuda")
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
def run(S, h2d=True, d2h=True, chain=False):
    per = P // S
    evs = [torch.cuda.Event() for _ in range(S)]
    for s in range(S):
        f, l = s * per, (s + 1) * per
        if h2d:
            with torch.cuda.stream(sa):
                d_xy[f:l].copy_(xy[f:l], non_blocking=True)
                evs[s].record(sa)
        if d2h:
            with torch.cuda.stream(sb):
                if chain and h2d:
                    sb.wait_event(evs[s])
                snd[f:l].copy_(d_snd[f:l], non_blocking=True)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
for S in (1, 8, 25):
    print("S=%2d  H2D %.3f  D2H %.3f  both %.3f  both, D2H slice i after H2D slice i %.3f ms" % (
        S, t(lambda: run(S, True, False)), t(lambda: run(S, False, True)), t(lambda: run(S)), t(lambda: run(S, chain=True))))

# the same with a kernel per slice between the two copies (three streams), like the engine's pipeline
sk = torch.cuda.Stream()
d_imp = torch.empty((P, K), dtype=torch.uint8, device="cuda")
imp = torch.empty((P, K), dtype=torch.uint8).pin_memory()
def pipe(S, two_outputs):
    per = P // S
    for s in range(S):
        f, l = s * per, (s + 1) * per
        e1, e2 = torch.cuda.Event(), torch.cuda.Event()
        with torch.cuda.stream(sa):
            d_xy[f:l].copy_(xy[f:l], non_blocking=True)
            e1.record(sa)
        with torch.cuda.stream(sk):
            sk.wait_event(e1)
            torch.add(d_xy[f:l, :, 0], d_xy[f:l, :, 1], out=d_snd[f:l])     # a short kernel on the slice
            e2.record(sk)
        with torch.cuda.stream(sb):
            sb.wait_event(e2)
            if two_outputs:
                imp[f:l].copy_(d_imp[f:l], non_blocking=True)
            snd[f:l].copy_(d_snd[f:l], non_blocking=True)
for S in (8, 16):
    print("S=%2d  copy -> kernel -> copy pipeline: one output %.3f ms, two outputs %.3f ms" % (
        S, t(lambda: pipe(S, False)), t(lambda: pipe(S, True))))
