#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-g}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_streams.py tests/test_gpu_parity.py tests/test_gpu_rescaling.py tests/test_gpu_fullsize.py -q --no-header -rf > $OUT/${TAG}_pytest.log 2>&1
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
run c2_st0 C3_STREAMS=0
run c2_st1 C3_STREAMS=1
run c2_st1b C3_STREAMS=1
CFG="--config c3"
run c3_st0 C3_STREAMS=0
run c3_st1 C3_STREAMS=1
CFG="--config c4"
run c4_st0 C3_STREAMS=0
run c4_st1 C3_STREAMS=1
CFG="--config c5 --columns 1250000"
run c5s_st0 C3_STREAMS=0
run c5s_st1 C3_STREAMS=1
CFG="--config c2 --columns 100000"
run c2s_st0 C3_STREAMS=0
run c2s_st1 C3_STREAMS=1
CFG="--config c1"
run c1_st1 C3_STREAMS=1
