#!/bin/bash
# round 2, second half: one copy kernel for the parameter blocks of a batch graph
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-b14}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py tests/test_gpu_multigpu.py -m gpu -q --no-header -rf -k "batch or many or lockstep or multi or sharded or loci or fused" > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
for v in 1 0 1 0; do
  env C3_BATCH_ONE_COPY=$v timeout 300 python bench.py --config c1 --steps 300 --warmup 5 --no-extra --no-dropin --no-cpu > $OUT/${TAG}_c1_onecopy$v.json 2> $OUT/${TAG}_c1_onecopy$v.err
  python - <<PY
import json
d=json.loads(open('$OUT/${TAG}_c1_onecopy$v.json').read().strip().splitlines()[-1])
print('one_copy=$v', d.get('many_small'))
PY
done
