#!/bin/bash
# wavefront kernel: parity tests, then C2 / C3 bench lines with and without it, a few lags
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-w}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 600 python -m pytest tests/test_gpu_wave.py -q --no-header -rf -x > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -4 $OUT/${TAG}_pytest.log
B="python bench.py --steps 200 --warmup 5 --no-extra --no-dropin --no-cpu"
run() { # name, env...
  local name=$1; shift
  env "$@" C3_WAVE_VERBOSE=1 timeout -k 10 240 $B ${CFG} > $OUT/${TAG}_${name}.json 2> $OUT/${TAG}_${name}.err
  echo "$name rc=$? $(python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_${name}.json').read().strip().splitlines()[-1])
    r=d['roofline']
    print('ms', round(d['ms_per_step'],4), 'frac', round(r['frac'],3), 'all', round(r['all_pruning_launches']['frac'],3), 'mirror_e2e_ms', round(d['e2e_mirror']['ms_per_step'],4), 'launches', d['gpu_launches'], 'lnL', d['lnL'])
except Exception as e: print('no line', e)
PY
)"
}
CFG="--config c2"
run c2_wave0 C3_WAVE=0
run c2_g1 C3_WAVE=1
run c2_g4 C3_WAVE=1 C3_WAVE_GROUP=4
run c2_g8 C3_WAVE=1 C3_WAVE_GROUP=8
run c2_g16 C3_WAVE=1 C3_WAVE_GROUP=16
run c2_g8l8 C3_WAVE=1 C3_WAVE_GROUP=8 C3_WAVE_LAG=8
run c2_p1k C3_WAVE=1 C3_WAVE_PREFIX=1200
run c2_p3k C3_WAVE=1 C3_WAVE_PREFIX=3000
run c2_p10k C3_WAVE=1 C3_WAVE_PREFIX=10000
run c2_p10kg2 C3_WAVE=1 C3_WAVE_PREFIX=10000 C3_WAVE_GROUP=2
run c2_p10kl3 C3_WAVE=1 C3_WAVE_PREFIX=10000 C3_WAVE_LAG=3
CFG="--config c3"
run c3_wave0 C3_WAVE=0
run c3_g8 C3_WAVE=1 C3_WAVE_GROUP=8
run c3_p3k C3_WAVE=1 C3_WAVE_PREFIX=3000
run c3_p10k C3_WAVE=1 C3_WAVE_PREFIX=10000
CFG="--config c5 --columns 1250000"
run c5s_wave0 C3_WAVE=0
run c5s_g8 C3_WAVE=1 C3_WAVE_GROUP=8
run c5s_p10k C3_WAVE=1 C3_WAVE_PREFIX=10000
grep -h "wave schedule" $OUT/${TAG}_*.err | sort | uniq -c | sort -rn | head -12
for f in c2_wave0 c2_p10k c2_g8; do python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_$f.json').read().strip().splitlines()[-1])
print('$f', [(x['kind'], x['ms']) for x in d['roofline']['per_launch']])
PY
done
