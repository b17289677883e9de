#!/bin/bash
# round 2, second half, session 2: chain kernel (C1 latency path), root instance with 3 rows per lane
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-b2}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q --no-header -rf -x > $OUT/${TAG}_pytest_parity.log 2>&1
tail -5 $OUT/${TAG}_pytest_parity.log
B="python bench.py --steps 300 --warmup 5 --no-extra --no-dropin --no-cpu"
run() { local name=$1; shift
  env "$@" timeout -k 10 240 $B ${CFG} > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],3), 'all', round(r['all_pruning_launches']['frac'],3), 'mirror_e2e_ms', round(d['e2e_mirror']['ms_per_step'],4), 'launches', d['gpu_launches'], 'lnL', d['lnL'])
    print('     ', [(x['kind'], round(x['ms']*1e3,1)) for x in r['per_launch']])
    if d.get('partial_update'): print('      partial', round(d['partial_update']['evals_per_s']), d['partial_update'].get('launches_of_one_update'))
    if d.get('many_small'): print('      many_small', d['many_small'])
except Exception as e: print('no line', e)
PY
)"
  grep "c3 timeline" $OUT/${TAG}_${name}.err | tail -12
}
CFG="--config c2"
run c2_root2 C3_ROOT_U=2
run c2_root3 C3_ROOT_U=3
run c2_root3_full C3_ROOT_U=3 C3_LEAN_ROOT=0
CFG="--config c5 --columns 1250000"
run c5s_root3 C3_ROOT_U=3
CFG="--config c1"
run c1_chain C3_CHAIN=1
run c1_nochain C3_CHAIN=0
run c1_chain_tl C3_CHAIN=1 C3_TIMELINE=1
