#!/bin/bash
#  Third gpurun call (what is left of the round's GPU budget: ~3 minutes).
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/round1d.log) 2>&1
date
echo "== a. A/B count kernel x digit histogram";  date +%T
timeout 60 python tools/ab_bench.py 8 5 countopts 2>&1 | tee $OUT/ab_8gbp_countopts.jsonl | cut -c1-420
echo "== b. union-sum + whole-path tests";  date +%T
timeout 40 python -m pytest tests/test_gpu_union.py -x -q -k "not large" 2>&1 | tail -2
timeout 50 python -m pytest tests/test_gpu_count.py -x -q -k "golden or k_sweep or forward_reverse" 2>&1 | tail -2
echo "== c. union-sum next to the reference";  date +%T
timeout 45 python tools/union_bench.py 0.25 2>&1 | tee $OUT/union_bench2.json | cut -c1-900
echo "== d. ncu: DRAM bytes of the default sweep kernel";  date +%T
timeout 45 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none -k regex:"sweep|hist2|suffix32b" -c 9 --csv --log-file $OUT/dram_v7.csv \
    python bench.py --gbp 1 --workspace-gb 6 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/ncu_dram_v7.log 2>&1
grep -c sweep $OUT/dram_v7.csv
date
