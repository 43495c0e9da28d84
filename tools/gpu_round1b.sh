#!/bin/bash
#  One gpurun call, ordered by priority (the round has ~16 GPU-minutes left): every step has its own timeout and
#  writes to gpurun_out/ as it goes.
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/round1b.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
echo "== a. stage tests: both sweep kernels, look-back";  date +%T
timeout 200 python -m pytest tests/test_gpu_stages.py -x -q -k "sort or lookback" 2>&1 | tail -4
echo "== b. A/B: sweep v1/v2 x filter, 8 Gbp in 5 passes";  date +%T
timeout 240 python tools/ab_bench.py 8 5 2>&1 | tee $OUT/ab_8gbp.jsonl | cut -c1-400
echo "== c. union-sum + multi-pass filter tests";  date +%T
timeout 200 python -m pytest tests/test_gpu_union.py -x -q 2>&1 | tail -4
timeout 200 python -m pytest tests/test_gpu_count.py -x -q -k "prefilter or multi_pass or golden or forward_reverse" 2>&1 | tail -4
echo "== d. bench 100 Gbp, sweep v2";  date +%T
MB_SORT=2 timeout 420 python bench.py --steps 1 --warmup 1 > $OUT/bench_100gbp_v2.json 2> $OUT/bench_100gbp_v2.err; cut -c1-1500 $OUT/bench_100gbp_v2.json; tail -3 $OUT/bench_100gbp_v2.err
echo "== e. ncu: launch list + DRAM bytes of the sweep kernel (1 Gbp, v2)";  date +%T
MB_SORT=2 timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_v2.csv \
    python bench.py --gbp 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/ncu_launch_v2.log 2>&1
MB_SORT=2 timeout 150 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread \
    --clock-control none -k regex:"sweep|extract_kernel|count_suffix32" -c 10 --csv --log-file $OUT/dram_v2.csv \
    python bench.py --gbp 0.5 --workspace-gb 4 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/ncu_dram_v2.log 2>&1
tail -12 $OUT/dram_v2.csv | cut -c1-300
echo "== f. the rest of the GPU suite";  date +%T
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
date
