#!/bin/bash
# round-2 profiles: launch list of the default bench command, full captures of the dominant 4-state kernel
# (one step's launches) and of the wavefront experiment
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
python __graft_entry__.py > $OUT/ncu_build.log 2>&1
CMD="python bench.py --steps 2 --warmup 3 --no-extra --no-dropin --no-cpu"
$CMD > $OUT/ncu_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/r2_launches_c2.csv $CMD > $OUT/ncu_launch.log 2>&1
echo "launch list rc=$?"
$CMD > $OUT/ncu_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:prune4_reg2 -s 21 -c 7 -o $OUT/r2_prune4_reg2 -f $CMD > $OUT/ncu_full.log 2>&1
echo "full capture rc=$?"
env C3_WAVE=1 C3_WAVE_GROUP=8 $CMD > $OUT/ncu_plain3.log 2>&1 &&
env C3_WAVE=1 C3_WAVE_GROUP=8 ncu --set full --clock-control none --import-source on -k regex:prune4_wave -s 2 -c 1 -o $OUT/r2_prune4_wave -f $CMD > $OUT/ncu_wave.log 2>&1
echo "wave capture rc=$?"
ls -la $OUT/*.ncu-rep
