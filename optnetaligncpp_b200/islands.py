"""Island sharding of the population over the GPUs of one node.

Individuals never interact during scoring or hill climbing (each correctHillClimb touches one
Alignment, localSearch.cpp:18-55), so the population is split into contiguous islands, one per
rank; networks and score tables are replicated.  The only exchange is the per-generation
all-gather of the [P, M] objective rows that the host NSGA-II sort consumes
(nsga-ii.cpp:14-118) -- a few KB, latency-bound, NCCL over NVLink on GPUs and gloo in CPU tests.
"""
import torch
import torch.distributed as dist


def shard(P, rank, world):
    """contiguous island [first, first+count) of rank `rank`; sizes differ by at most one."""
    base, rem = divmod(P, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_gather_objectives(rows, P, world):
    """rows: this rank's [count, M] objective rows (any device).  Returns the [P, M] matrix in
    global individual order on every rank."""
    if world == 1:
        return rows
    M = rows.shape[1]
    counts = [shard(P, r, world)[1] for r in range(world)]
    if len(set(counts)) == 1:
        out = torch.empty((P, M), dtype=rows.dtype, device=rows.device)
        dist.all_gather_into_tensor(out, rows.contiguous())
        return out
    pad = max(counts)
    buf = torch.zeros((pad, M), dtype=rows.dtype, device=rows.device)
    buf[: rows.shape[0]] = rows
    parts = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(parts, buf)
    return torch.cat([parts[r][: counts[r]] for r in range(world)], 0)
