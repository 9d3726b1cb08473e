cd /root/repo
python - <<'PY'
import sys, os
sys.path.insert(0,'/root/repo')
import bench
from optnetaligncpp_b200 import synth
pair = bench.make_workload()
os.makedirs('/tmp/c2', exist_ok=True)
synth.write_text(pair, '/tmp/c2/c2')
PY
for i in 1 2; do
( time ONA_TRACE_CREATE=1 ONA_HOST_STATS=1 optnetaligncpp_b200/host/_build/optnetalign --net1 /tmp/c2/c21.net --net2 /tmp/c2/c22.net --total --s3 --bitscores /tmp/c2/c2.sim --blastsum --popsize 1 --generations 1 --hillclimbiters 0 --nooutput --randseed 1 --nthreads 1 ) 2>&1 | tail -16
done
( time oracle/_ref/optnetalign_ref --net1 /tmp/c2/c21.net --net2 /tmp/c2/c22.net --total --s3 --bitscores /tmp/c2/c2.sim --blastsum --popsize 1 --generations 1 --hillclimbiters 0 --nooutput --randseed 1 --nthreads 1 ) 2>&1 | tail -4
