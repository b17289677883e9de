#!/bin/bash
# round 2, second half, session 1: lean root / 16-warp regpipe2 / launch timeline / TMA-gather microbenchmark
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-b1}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -q --no-header -rf -x > $OUT/${TAG}_pytest_parity.log 2>&1
tail -3 $OUT/${TAG}_pytest_parity.log
B="python bench.py --steps 300 --warmup 5 --no-extra --no-dropin --no-cpu"
run() { local name=$1; shift
  env "$@" timeout -k 10 240 $B ${CFG} > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],3), 'all', round(r['all_pruning_launches']['frac'],3), 'launches', d['gpu_launches'], 'lnL', d['lnL'])
    print('     ', [(x['kind'], round(x['ms']*1e3,1)) for x in r['per_launch']])
except Exception as e: print('no line', e)
PY
)"
  grep "c3 timeline" $OUT/${TAG}_${name}.err | head -40
}
CFG="--config c2"
run c2_full C3_LEAN_ROOT=0
run c2_lean C3_LEAN_ROOT=1
run c2_lean_u2 C3_LEAN_ROOT=1 C3_REG2_U2=1
run c2_full_u2 C3_LEAN_ROOT=0 C3_REG2_U2=1
run c2_lean_tl C3_LEAN_ROOT=1 C3_TIMELINE=1
run c2_lean_u2_tl C3_LEAN_ROOT=1 C3_REG2_U2=1 C3_TIMELINE=1
CFG="--config c5 --columns 1250000"
run c5s_lean C3_LEAN_ROOT=1
run c5s_lean_u2 C3_LEAN_ROOT=1 C3_REG2_U2=1
CFG="--config c4"
run c4_tl C3_TIMELINE=1
timeout -k 5 300 tools/_tma_gather_bench 2000000 > $OUT/${TAG}_tma_gather.txt 2>&1; echo "tma rc=$?"; cat $OUT/${TAG}_tma_gather.txt
