#!/bin/bash
#  Round 2, GPU call B (one B200, ~25 min): what was written after call A, tested and measured -- the dense extraction form,
#  the joint digit histogram, the grouped run-length count, L2 prefetch distances, the streamed db writer, the arena cache.
#  usage (from the repo root):  gpurun --timeout 2400 -- 'bash tools/gpu_r2_b.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/r2b.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2; df -h /dev/shm | tail -1
echo "== a. the GPU suite, experimental tests included, no -x";  date +%T
MB_TEST_EXPERIMENTAL=1 timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -25
echo "== b. A/B at 8 Gbp in 5 passes (one process, same reads)";  date +%T
MB_AB_ENVS='[{}, {"MB_EXTRACT_DENSE": "0"}, {"MB_DENSE_ITEMS": "16"}, {"MB_HIST_SINGLE": "1"}, {"MB_COUNT_GROUP": "1"}, {"MB_COUNT_GROUP": "2"}, {"MB_COUNT_GROUP": "8"}, {"MB_COUNT_GROUP": "1", "MB_COUNT_PREFETCH": "1184"}, {"MB_SWEEP_PREFETCH": "444"}, {"MB_SWEEP_PREFETCH": "1332"}]' \
  timeout 400 python tools/ab_bench.py 8 5 env 2>&1 | tee $OUT/r2b_ab_8gbp_env.jsonl | cut -c1-420
echo "== c. bench 100 Gbp, all legs";  date +%T
timeout 900 python bench.py --steps 2 --warmup 1 > $OUT/r2b_bench_100gbp.json 2> $OUT/r2b_bench_100gbp.err; cut -c1-3000 $OUT/r2b_bench_100gbp.json; tail -2 $OUT/r2b_bench_100gbp.err
echo "== d. ncu --set full with source: the new kernels (0.5 Gbp, 5 file-aligned passes)";  date +%T
MB_MAX_PASS_KEYS=88500000 timeout 600 ncu --set full --import-source on --clock-control none -k regex:"extract_dense|sort_hist2|count_suffix32g|sortx_sweep" -c 12 -f -o $OUT/r2b_full \
    python bench.py --gbp 0.5 --workspace-gb 6 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2b_ncu_full.log 2>&1
[ -f $OUT/r2b_full.ncu-rep ] && ncu -i $OUT/r2b_full.ncu-rep --page raw --csv > $OUT/r2b_full_raw.csv 2>/dev/null
[ -f $OUT/r2b_full.ncu-rep ] && ncu -i $OUT/r2b_full.ncu-rep --page source --csv > $OUT/r2b_full_source.csv 2>/dev/null
rm -f $OUT/r2b_full.ncu-rep
echo "== d2. ncu launch list (1 Gbp, 5 passes)";  date +%T
MB_MAX_PASS_KEYS=177000000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/r2b_launches.csv \
    python bench.py --gbp 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2b_ncu_launch.log 2>&1
echo "== e. union-sum next to the reference: 0.5 Gbp, then 3 Gbp per input";  date +%T
timeout 240 python tools/union_bench.py 0.5 2>&1 | tee $OUT/r2b_union_bench_0.5.json | cut -c1-1000
timeout 600 python tools/union_bench.py 3 2>&1 | tee $OUT/r2b_union_bench_3.json | cut -c1-1000
ls -la $OUT | tail -12
date
