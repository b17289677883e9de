#!/bin/bash
# one gpurun call: GPU test-suite, default bench line, host facts.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-s}
{ free -g | head -2; nproc; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; } > $OUT/${TAG}_host.txt 2>&1
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q --no-header -rf > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
timeout 900 python bench.py --steps 200 --warmup 5 > $OUT/${TAG}_bench_c2.json 2> $OUT/${TAG}_bench_c2.err
echo "bench rc=$?" >> $OUT/${TAG}_bench_c2.err
tail -3 $OUT/${TAG}_pytest.log; tail -c 600 $OUT/${TAG}_bench_c2.json
