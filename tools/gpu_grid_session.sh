#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-g}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_rec.py tests/test_gpu_parity.py tests/test_gpu_rescaling.py -q --no-header -rf > $OUT/${TAG}_pytest.log 2>&1
tail -4 $OUT/${TAG}_pytest.log
B="python bench.py --steps 200 --warmup 5 --no-extra --no-dropin --no-cpu"
run() { local name=$1; shift
  env "$@" timeout -k 10 240 $B ${CFG} > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],3), 'all', round(r['all_pruning_launches']['frac'],3), 'mirror_e2e_ms', round(d['e2e_mirror']['ms_per_step'],4), 'launches', d['gpu_launches'], 'replays', d['graph_replays'], 'lnL', d['lnL'])
    print('     ', [(x['kind'], round(x['ms']*1e3,1)) for x in r['per_launch']])
    if d.get('partial_update'): print('      partial', round(d['partial_update']['evals_per_s']), d['partial_update']['launches_of_one_update'])
except Exception as e: print('no line', e)
PY
)"
}
CFG="--config c2"
run c2_norec C3_NO_REC=1
run c2_rec C3_X=1
run c2_rec1m C3_REC_UNITS=1000000
run c2_rec150k C3_REC_UNITS=150000
CFG="--config c3"
run c3_norec C3_NO_REC=1
run c3_rec C3_X=1
run c3_rec1m C3_REC_UNITS=1000000
CFG="--config c5 --columns 1250000"
run c5s_norec C3_NO_REC=1
run c5s_rec C3_X=1
CFG="--config c2 --columns 100000"
run c2s_norec C3_NO_REC=1
run c2s_rec C3_X=1
run c2s_rec1m C3_REC_UNITS=1000000
