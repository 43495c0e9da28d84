#!/bin/bash
#  First gpurun call of the next round: everything that was written without a GPU at hand gets tested and measured, and
#  the default kernels get the ncu captures that round 1 could not afford (launch list + --set full with source).
#  usage (from the repo root):  gpurun --timeout 1500 -- 'bash tools/gpu_round2_first.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/round2a.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
echo "== a. the GPU suite, experimental tests included";  date +%T
MB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4
echo "== b. A/B: experiments (rank by match, ...) at 8 Gbp in 5 passes";  date +%T
timeout 120 python tools/ab_bench.py 8 5 experiments 2>&1 | tee $OUT/ab_8gbp_experiments.jsonl | cut -c1-420
echo "== b2. A/B: the sweep shapes, incl. the unmeasured ones (5: 512x6, 6: 256x6 at 128 registers, 7: persistent with prefetch)";  date +%T
timeout 240 python tools/ab_bench.py 8 5 sortonly 2>&1 | tee $OUT/ab_8gbp_shapes_r2.jsonl | cut -c1-420
echo "== c. bench 100 Gbp: default, then with the copy-back overlapped";  date +%T
timeout 420 python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_100gbp_default.json 2> $OUT/bench_100gbp_default.err; cut -c1-1500 $OUT/bench_100gbp_default.json
MB_OVERLAP_D2H=1 timeout 420 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > $OUT/bench_100gbp_overlap.json 2> $OUT/bench_100gbp_overlap.err; cut -c1-1500 $OUT/bench_100gbp_overlap.json
echo "== d. ncu: launch list of the default kernels (1 Gbp), then --set full with source of the sweep and the extraction";  date +%T
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_r2a.csv \
    python bench.py --gbp 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/ncu_launch_r2a.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"sortx_sweep|extract_kernel|sort_hist_kernel|count_suffix32_kernel" -c 6 -o $OUT/full_r2a \
    python bench.py --gbp 0.25 --workspace-gb 3 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/ncu_full_r2a.log 2>&1
ls -la $OUT | tail -12
echo "== e. union-sum next to the reference (staging ring; then with the warp-per-block decoder)";  date +%T
timeout 200 python tools/union_bench.py 0.5 2>&1 | tee $OUT/union_bench_r2a.json | cut -c1-900
MB_DECODE=warp timeout 200 python tools/union_bench.py 0.5 2>&1 | tee $OUT/union_bench_r2a_warp.json | cut -c1-900
echo "== f. meryl-lookup -existence next to the reference; ingest through the device parser next to the host parser";  date +%T
timeout 300 python tools/lookup_bench.py 300 2>&1 | tee $OUT/lookup_bench_r2a.jsonl | cut -c1-400
date
