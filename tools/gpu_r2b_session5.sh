#!/bin/bash
# round 2, second half, session: the frozen-argument fix (chain launch under graph replay), dynamic strips
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-b11}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_reference_suite_dropin.py -m gpu -q --no-header -rf > $OUT/${TAG}_pytest.log 2>&1
tail -5 $OUT/${TAG}_pytest.log
B="python bench.py --steps 300 --warmup 5 --no-extra --no-dropin --no-cpu"
run() { local name=$1; shift
  env "$@" timeout -k 10 240 $B ${CFG} > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],3), 'all', round(r['all_pruning_launches']['frac'],3), 'whole', round(r.get('whole_step',{}).get('frac',0),3), 'lnL', d['lnL'])
    print('     ', [(x['kind'], round(x['ms']*1e3,1)) for x in r['per_launch']])
except Exception as e: print('no line', e)
PY
)"
  grep "c3 timeline" $OUT/${TAG}_${name}.err | sed -n 2,10p
}
CFG="--config c2"
run c2_static C3_DYN=0
run c2_dyn C3_DYN=1
run c2_dyn_tl C3_DYN=1 C3_TIMELINE=1
run c2_static2 C3_DYN=0
CFG="--config c5 --columns 1250000"
run c5s_static C3_DYN=0
run c5s_dyn C3_DYN=1
CFG="--config c3"
run c3_static C3_DYN=0
run c3_dyn C3_DYN=1
