#!/bin/bash
# round 2, second half: the tree-walk kernel with bodies specialised per node shape
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-b15}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -q --no-header -rf -k "walk or chain" > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
for v in "C3_WALK=1" "C3_WALK=0" "C3_WALK=1 C3_TIMELINE=1"; do
  env $v timeout 300 python bench.py --config c1 --steps 300 --warmup 5 --no-extra --no-dropin --no-cpu > $OUT/${TAG}_c1.json 2> $OUT/${TAG}_c1.err
  python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_c1.json').read().strip().splitlines()[-1])
print('$v', 'ms', round(d['ms_per_step'],4), 'mirror_e2e', round(d['e2e_mirror']['ms_per_step'],4), 'many_small', round(d['many_small']['evals_per_s']), 'partial', round(d['partial_update']['evals_per_s']), [(x['kind'], round(x['ms']*1e3,1)) for x in d['roofline']['per_launch']])
PY
  grep "kind 112" $OUT/${TAG}_c1.err | head -2
done
env C3_WALK=1 timeout 300 python bench.py --config c2 --columns 2000 --steps 300 --warmup 5 --no-extra --no-dropin --no-cpu > $OUT/${TAG}_c2small_walk.json 2> /dev/null
env C3_WALK=0 timeout 300 python bench.py --config c2 --columns 2000 --steps 300 --warmup 5 --no-extra --no-dropin --no-cpu > $OUT/${TAG}_c2small_nowalk.json 2> /dev/null
python - <<PY
import json
for n in ('walk','nowalk'):
    d=json.loads(open('$OUT/${TAG}_c2small_'+n+'.json').read().strip().splitlines()[-1])
    print('c2 x 2000 columns (K = 4)', n, 'ms', round(d['ms_per_step'],4), [(x['kind'], round(x['ms']*1e3,1)) for x in d['roofline']['per_launch']])
PY
