#!/bin/bash
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-t}
python __graft_entry__.py > $OUT/${TAG}_build.log 2>&1
timeout 1700 python -m pytest tests -m gpu -q --no-header -rf --durations=15 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
