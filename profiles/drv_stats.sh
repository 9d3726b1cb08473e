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
for nt in ${NTS:-16 64 256}; do
echo "== nthreads $nt"
( time ONA_HOST_STATS=1 optnetaligncpp_b200/host/_build/optnetalign --net1 /tmp/c2/c21.net --net2 /tmp/c2/c22.net --total --s3 --bitscores /tmp/c2/c2.sim --blastsum --popsize 200 --generations 1 --hillclimbiters 2000 --nooutput --randseed 1 --nthreads $nt ) 2>&1 | tail -14
done
if [ -n "$SMALLOPS" ]; then
for nt in 1 16 64; do
echo "== small ops only (hillclimbiters 0), nthreads $nt"
( time ONA_HOST_STATS=1 optnetaligncpp_b200/host/_build/optnetalign --net1 /tmp/c2/c21.net --net2 /tmp/c2/c22.net --total --s3 --bitscores /tmp/c2/c2.sim --blastsum --popsize 200 --generations 5 --hillclimbiters 0 --nooutput --randseed 1 --nthreads $nt ) 2>&1 | tail -10
done
fi
