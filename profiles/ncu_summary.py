#!/usr/bin/env python
"""Summarise a .ncu-rep (read with `ncu -i`) into the few numbers DESIGN.md / bench.py cite.

    python profiles/ncu_summary.py gpurun_out/prof.ncu-rep [top_n] > profiles/<name>.txt
"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed_op_local_ld.sum", "smsp__inst_executed_op_local_st.sum",
]


def ncu(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    kn = hdr.index("Kernel Name")
    print("# %s" % rep)
    for r in rows[2:]:
        print("kernel:", r[kn][:110])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print("%-66s %s  [%s]" % (k, " | ".join(r[i] for r in rows[2:]), units[i]))
    src = list(csv.reader(io.StringIO(ncu(["-i", rep, "--page", "source", "--csv"]))))
    # first kernel only
    start = [i for i, r in enumerate(src) if r and r[0] == "Address"]
    if not start:
        return
    h = src[start[0]]
    end = start[1] - 1 if len(start) > 1 else len(src)
    body = [r for r in src[start[0] + 1:end] if len(r) == len(h)]
    si = h.index("# Samples")
    stall_cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
    tot = {h[i]: sum(int(r[i] or 0) for r in body) for i in stall_cols}
    allsamp = sum(tot.values()) or 1
    print("\nstall samples (all, first launch):")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:9]:
        print("  %-24s %8d  %5.1f%%" % (k, v, 100.0 * v / allsamp))
    ie = h.index("Instructions Executed")
    print("\ninstructions executed (warp-level) by opcode class:")
    cls = {}
    for r in body:
        op = r[1].split()[0] if r[1].split() else "?"
        if op.startswith("@"):
            op = r[1].split()[1]
        op = op.split(".")[0]
        cls[op] = cls.get(op, 0) + int(r[ie] or 0)
    tot_i = sum(cls.values()) or 1
    for k, v in sorted(cls.items(), key=lambda kv: -kv[1])[:14]:
        print("  %-10s %12d  %5.1f%%" % (k, v, 100.0 * v / tot_i))
    print("\ntop %d SASS instructions by stall samples:" % top)
    order = sorted(range(len(body)), key=lambda i: -int(body[i][si] or 0))[:top]
    for i in sorted(order):
        r = body[i]
        dom = max(stall_cols, key=lambda c: int(r[c] or 0))
        print("  %5d  %-70s samples=%-6s exec=%-9s thr=%-5s %s" % (i, r[1].strip()[:70], r[si], r[ie],
                                                              r[h.index("Avg. Threads Executed")], h[dom]))


if __name__ == "__main__":
    main()
