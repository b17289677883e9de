#!/bin/bash
# N-GPU bench line (torchrun), c1 latency line
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-n}; N=${2:-2}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 \
  bench.py --gpus $N --steps 200 --warmup 5 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err
echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d=json.loads(open('$OUT/${TAG}_bench_n$N.json').read().strip().splitlines()[-1])
    print('value', d['value'], 'ms', d['ms_per_step'], 'e2e ms', d['e2e']['ms_per_step'], d['e2e'].get('fused_allreduce'), d['e2e'].get('graph_replays'), 'mirror', d['e2e_mirror']['ms_per_step'])
    print('allreduce', d['allreduce'])
    c5=d['c5_strong']; print('c5', {k:c5.get(k) for k in ('ms_per_step','value','lnL','allreduce','U_gp_per_step_all_ranks','error','skipped')})
except Exception as e: print('no line', e)
PY
tail -5 $OUT/${TAG}_bench_n$N.err
timeout -k 10 600 python -m pytest tests/test_gpu_multigpu.py tests/test_gpu_dropin.py -q --no-header -rf -k "multigpu or sharded or between_processes or graph_replay" > $OUT/${TAG}_pytest.log 2>&1
tail -3 $OUT/${TAG}_pytest.log
