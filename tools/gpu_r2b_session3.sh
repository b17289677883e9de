#!/bin/bash
# round 2, second half, session 3: fence-free result publication, parallel tile reduction in the chain kernel,
# host profile of the blocking path; then the whole GPU suite + the default bench line
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-b3}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -q --no-header -rf -x > $OUT/${TAG}_pytest_parity.log 2>&1
tail -3 $OUT/${TAG}_pytest_parity.log
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
    if d.get('many_small'): print('      many_small', round(d['many_small']['evals_per_s']), round(d['many_small']['one_at_a_time_evals_per_s']))
except Exception as e: print('no line', e)
PY
)"
  grep "c3 hostprof" $OUT/${TAG}_${name}.err | tail -4
  grep "kind 111" $OUT/${TAG}_${name}.err | tail -3
}
CFG="--config c1"
run c1_chain C3_CHAIN=1 C3_HOSTPROF=1
run c1_chain_tl C3_CHAIN=1 C3_TIMELINE=1
run c1_nochain C3_CHAIN=0 C3_HOSTPROF=1
CFG="--config c2"
run c2 C3_HOSTPROF=1
timeout 1700 python -m pytest tests -m gpu -q --no-header -rf -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench rc=$?"
tail -c 1500 $OUT/${TAG}_bench_default.json
