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
for nt in ${NTS:-1 8 32}; do
ONA_HOST_STATS=1 optnetaligncpp_b200/host/_build/dropin_demo /tmp/c2/c21.net /tmp/c2/c22.net /tmp/c2/c2.sim - - 5 1000000 $nt /tmp/c2/out smallops 2>&1 | grep -E "smallops|evaluations  |swap lists  |waiting"
done
