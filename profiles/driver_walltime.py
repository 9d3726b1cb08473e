#!/usr/bin/env python
"""Wall-clock of the reference's own optnetalign driver (steadystate.cpp, unmodified) on the C2
pair: the all-reference CPU build at the host's thread count vs the GPU-engine build at several
--nthreads (thread count = batch size of the combined hill-climb launches).  Same CLI, same flags;
prints generated alignments per second and correctHillClimb iterations per second.
    python profiles/driver_walltime.py out.json"""
import json, os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from optnetaligncpp_b200 import synth

pair = bench.make_workload()
d = tempfile.mkdtemp()
p = synth.write_text(pair, os.path.join(d, "c2"))
popsize, gens, hci = 200, 1, int(os.environ.get("HCI", "2000"))
# steadystate.cpp:137 lets every worker start another child while fewer than popsize*generations have been
# FINISHED, so a run with T workers makes popsize*generations + T - 1 children (215 at 16 threads, 455 at 256,
# 1 223 at 1 024 -- confirmed by the call counts ONA_HOST_STATS prints); each child is climbed twice for
# 500*hillclimbiters iterations (:243-253).  Runs at different --nthreads therefore do different amounts of
# work: compare children per second, not wall-clock.
def work(nthreads):
    children = popsize * gens + nthreads - 1
    return children, children * 2 * 500 * hci
base = ["--net1", p["net1"], "--net2", p["net2"], "--total", "--s3", "--bitscores", p["sim"], "--blastsum",
        "--popsize", str(popsize), "--generations", str(gens), "--hillclimbiters", str(hci), "--nooutput", "--randseed", "1"]
ncpu = os.cpu_count() or 1
res = {}
def run(tag, binary, nthreads, args=None, reps=1):
    # wall-clock of a whole process; run-to-run spread is large for the GPU build (how the worker threads'
    # calls happen to fall into batched launches), so those are repeated and the median is kept
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        out = subprocess.run([binary] + (args or base) + ["--nthreads", str(nthreads)], capture_output=True, text=True, timeout=3000)
        ts.append(time.perf_counter() - t0)
        assert out.returncode == 0, out.stderr[-500:]
    dt = sorted(ts)[len(ts) // 2]
    res[tag] = dict(nthreads=nthreads, wall_s=dt, wall_s_all=ts)
    print("%-28s nthreads %4d  wall %7.2f s  %s" % (tag, nthreads, dt, ["%.2f" % t for t in ts]), flush=True)
# fixed cost of a run (process start, loaders incl. the 384 MB dense BLAST matrix, for the GPU build
# also CUDA context + engine creation): one child, no hill climbing
fixed = [a for a in base]
fixed[fixed.index("--popsize") + 1] = "1"
fixed[fixed.index("--hillclimbiters") + 1] = "0"
REF, GPU = os.path.join(ROOT, "oracle/_ref/optnetalign_ref"), os.path.join(ROOT, "optnetaligncpp_b200/host/_build/optnetalign")
run("reference_cpu_fixed_cost", REF, 1, fixed)
run("gpu_engine_fixed_cost_first_process", GPU, 1, fixed)   # the first CUDA process on a fresh box pays extra
run("gpu_engine_fixed_cost", GPU, 1, fixed, reps=3)
run("reference_cpu", REF, ncpu)
for nt in (ncpu, 64, 256, 1024):
    run("gpu_engine_nthreads_%d" % nt, GPU, nt, reps=3)
for k, v in res.items():
    if "fixed_cost" in k:
        continue
    f = res["gpu_engine_fixed_cost" if k.startswith("gpu") else "reference_cpu_fixed_cost"]["wall_s"]
    v["children"], v["climb_iterations"] = work(v["nthreads"])
    v["children_per_s_wall"] = v["children"] / v["wall_s"]
    v["iterations_per_s_net_of_fixed_cost"] = v["climb_iterations"] / max(v["wall_s"] - f, 1e-9)
for k, v in res.items():
    if k.startswith("gpu_engine_nthreads"):
        v["children_per_s_vs_reference_cpu"] = v["children_per_s_wall"] / res["reference_cpu"]["children_per_s_wall"]
        print("%-28s %5d children  %6.1f children/s wall (%.1fx the all-CPU build)  %.2e climb iterations/s net of start-up" % (
            k, v["children"], v["children_per_s_wall"], v["children_per_s_vs_reference_cpu"], v["iterations_per_s_net_of_fixed_cost"]), flush=True)
print("%-28s %5d children  %6.1f children/s wall  %.2e climb iterations/s net of start-up" % (
    "reference_cpu", res["reference_cpu"]["children"], res["reference_cpu"]["children_per_s_wall"],
    res["reference_cpu"]["iterations_per_s_net_of_fixed_cost"]), flush=True)
json.dump(dict(host_threads=ncpu, popsize=popsize, generations=gens, hillclimbiters=hci, runs=res),
          open(os.path.join(ROOT, "profiles", sys.argv[1] if len(sys.argv) > 1 else "driver_walltime.json"), "w"), indent=1)
