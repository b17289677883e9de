#!/bin/bash
# latency-path trims: parity / graph / batch tests, then the C1 block
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-b16}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multigpu.py tests/test_gpu_dropin.py tests/test_gpu_rescaling.py -m gpu -q --no-header -rf > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
for i in 1 2; do
env C3_HOSTPROF=1 timeout 300 python bench.py --config c1 --steps 300 --warmup 5 --no-extra --no-cpu > $OUT/${TAG}_c1.json 2> $OUT/${TAG}_c1.err
python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_c1.json').read().strip().splitlines()[-1])
print('ms', round(d['ms_per_step'],4), 'partial', round(d['partial_update']['evals_per_s']), 'single-param', round(d['e2e']['lf_evals']['single_parameter_real_changes_per_s']), 'optimise', d['e2e']['optimise']['cuda']['seconds'], d['e2e']['optimise']['cpu']['seconds'], 'many_small', round(d['many_small']['evals_per_s']))
PY
grep "c3 hostprof: 436" $OUT/${TAG}_c1.err
done
