"""What bounds the end-to-end leg: pinned-memory copy bandwidth of this host <-> GPU link,
each direction alone and both at once (the sizes of one bench step: 26.2 MB of proposals in,
29.5 MB of results out).  Prints one JSON line."""
import json
import torch

dev = torch.device("cuda:0")
h_in = torch.empty(26214400, dtype=torch.uint8).pin_memory()
h_out = torch.empty(29491200, dtype=torch.uint8).pin_memory()
d_in = torch.empty_like(h_in, device=dev)
d_out = torch.empty_like(h_out, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)


def both():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
    torch.cuda.current_stream().wait_stream(s2)


t_in, t_out, t_both = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"h2d_ms": t_in, "h2d_gbs": h_in.numel() / t_in / 1e6, "d2h_ms": t_out,
                  "d2h_gbs": h_out.numel() / t_out / 1e6, "both_ms": t_both,
                  "both_gbs_total": (h_in.numel() + h_out.numel()) / t_both / 1e6,
                  "swap_deltas_per_s_at_the_duplex_bound": 200 * 16384 / (t_both * 1e-3)}))
