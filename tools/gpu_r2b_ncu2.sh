#!/bin/bash
# ncu captures for the record: the TMA-gather microbenchmark's ring kernels, the codon (DMMA) launches of one C4 step
cd "$(dirname "$0")/.." || exit 1
OUT=gpurun_out; mkdir -p $OUT
python __graft_entry__.py > $OUT/ncu2_build.log 2>&1
timeout 120 tools/_tma_gather_bench 2000000 > $OUT/ncu2_tma_plain.txt 2>&1 &&
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_ring -c 14 -o $OUT/r2_tma_ring -f tools/_tma_gather_bench 2000000 > $OUT/ncu2_tma.log 2>&1
echo "tma capture rc=$?"
CMD="python bench.py --config c4 --steps 2 --warmup 3 --no-extra --no-dropin --no-cpu"
$CMD > $OUT/ncu2_c4_plain.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:prune_dmma_edge -s 18 -c 9 -o $OUT/r2b_dmma_edge -f $CMD > $OUT/ncu2_c4.log 2>&1
echo "c4 capture rc=$?"
ls -la $OUT/*.ncu-rep | tail -4
