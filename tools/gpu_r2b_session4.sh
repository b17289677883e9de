#!/bin/bash
# round 2, second half, session 4: fused chain launch (one launch per single-branch update)
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-b5}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -q --no-header -rf -x > $OUT/${TAG}_pytest_parity.log 2>&1
tail -5 $OUT/${TAG}_pytest_parity.log
B="python bench.py --steps 300 --warmup 5 --no-extra --no-cpu"
run() { local name=$1; shift
  env "$@" timeout -k 10 300 $B ${CFG} > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('ms', round(d['ms_per_step'],4), 'mirror_e2e_ms', round(d['e2e_mirror']['ms_per_step'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4), 'lnL', d['lnL'])
    if d.get('partial_update'): print('      partial', round(d['partial_update']['evals_per_s']), d['partial_update'].get('launches_of_one_update'))
    if d.get('many_small'): print('      many_small', round(d['many_small']['evals_per_s']), round(d['many_small']['one_at_a_time_evals_per_s']))
    print('      lf_evals', d['e2e'].get('lf_evals'))
    print('      optimise', d['e2e'].get('optimise'))
except Exception as e: print('no line', e)
PY
)"
  grep "c3 hostprof" $OUT/${TAG}_${name}.err | sed -n 1,3p
  grep "kind 11[12]" $OUT/${TAG}_${name}.err | tail -4
}
CFG="--config c1"
run c1_walk C3_HOSTPROF=1
run c1_walk_tl C3_TIMELINE=1
run c1_nowalk C3_WALK=0 C3_HOSTPROF=1
