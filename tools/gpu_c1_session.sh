#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-c}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
for v in pv0 pv1 pv2 nopath; do
  E="C3_X=0"; [ $v = nopath ] && E="C3_NO_PATH=1"; [ $v = pv0 ] && E="C3_PATH_VARIANT=0"; [ $v = pv1 ] && E="C3_PATH_VARIANT=1"; [ $v = pv2 ] && E="C3_PATH_VARIANT=2"
  env $E timeout -k 10 300 python -m pytest tests/test_gpu_wave.py -q --no-header -k path 2>&1 | tail -1
  env $E timeout -k 10 300 python bench.py --config c1 --steps 2000 --warmup 20 --no-cpu > $OUT/${TAG}_c1_$v.json 2> $OUT/${TAG}_c1_$v.err
  python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_c1_$v.json').read().strip().splitlines()[-1])
    print("$v", "full eval us", round(d['ms_per_step']*1e3,2), 'partial evals/s', round(d['partial_update']['evals_per_s']), 'e2e ms', round(d['e2e']['ms_per_step'],4), 'lf_evals', d['e2e'].get('lf_evals'), 'many', d['many_small'] and round(d['many_small']['evals_per_s']))
except Exception as e: print('$v no line', e)
PY
done
tail -3 $OUT/${TAG}_c1_pv1.err
