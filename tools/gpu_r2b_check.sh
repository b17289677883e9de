#!/bin/bash
# a short record: the whole GPU suite and the default bench line
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-g1}
python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 1700 python -m pytest tests -m gpu -q --no-header -rf > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -4 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_bench_default.json').read().strip().splitlines()[-1])
print('value',d['value'],'ms',d['ms_per_step'],'pipelined',d.get('pipelined'))
e=d['e2e']; print('e2e', e['value'], e['ms_per_step'], e['setup'])
r=d['roofline']; print(r['frac'], r['all_pruning_launches']['frac'], r['whole_step']['frac'])
c1=d['c1']; print('c1', c1['us_per_full_evaluation'], c1['partial_update']['evals_per_s'], c1['e2e']['ms_per_step'], c1['e2e'].get('optimise'), (d.get('many_small') or c1.get('many_small') or {}).get('evals_per_s'))
c4=d['c4']; print('c4', c4['value'], c4['ms_per_step'], c4['e2e']['ms_per_step'], c4['roofline']['frac'])
c5=d['c5_strong']; print('c5', c5['ms_per_step'], c5['hbm_frac_per_gpu'])
PY
