#!/bin/bash
# round 2, second half: the record -- smoke, the whole GPU suite, the default bench line, the reference arm, the ncu
# launch list of the default command's C2 part and a full capture of the dominant kernel's launches of one step
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
TAG=${1:-f2}
python __graft_entry__.py --smoke > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 $OUT/${TAG}_smoke.log
timeout 1700 python -m pytest tests -m gpu -q --no-header -rf --durations=8 > $OUT/${TAG}_pytest.log 2>&1
echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log; tail -14 $OUT/${TAG}_pytest.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.json 2> $OUT/${TAG}_bench_default.err; echo "bench rc=$?"
( time timeout 1700 python bench.py --impl reference --steps 10 --warmup 1 > $OUT/${TAG}_bench_reference.json 2> $OUT/${TAG}_bench_reference.err ) 2> $OUT/${TAG}_reference_time.txt; echo "reference rc=$?"; tail -3 $OUT/${TAG}_reference_time.txt
tail -c 600 $OUT/${TAG}_bench_reference.json; echo
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-dropin --no-cpu"
$CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:prune|expm|lnl_reduce' -c 300 --csv --log-file $OUT/r2b_launches_c2.csv $CMD > $OUT/${TAG}_ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/${TAG}_ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:prune4_reg2 -s 21 -c 7 -o $OUT/r2b_prune4_reg2 -f $CMD > $OUT/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"
env C3_WALK=1 timeout 300 python bench.py --config c1 --steps 300 --warmup 5 --no-extra --no-dropin --no-cpu > $OUT/${TAG}_c1_walk.json 2> $OUT/${TAG}_c1_walk.err; echo "c1 walk rc=$?"
ls -la $OUT/*.ncu-rep | tail -3
