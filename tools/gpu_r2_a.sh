#!/bin/bash
#  Round 2, GPU call A (one B200, ~30 min): the whole GPU suite WITHOUT -x (incl. the tests that never ran on hardware and the
#  headline-plan parity test), A/B of what was written without a GPU, the 100 Gbp bench, ncu launch list + --set full
#  captures (with source) of the default kernels, union-sum and the command line tools next to the reference.
#  usage (from the repo root):  gpurun --timeout 2100 -- 'bash tools/gpu_r2_a.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/r2a.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
echo "== a. the GPU suite, experimental tests included, no -x";  date +%T
MB_TEST_EXPERIMENTAL=1 timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -25
echo "== b. A/B: extraction ranks (atomic vs match) at 8 Gbp in 5 passes";  date +%T
timeout 150 python tools/ab_bench.py 8 5 experiments 2>&1 | tee $OUT/r2a_ab_8gbp_experiments.jsonl | cut -c1-420
echo "== b2. A/B: sweep shapes 3 (default), 5 (512x6), 6 (256x6, 128 regs), 7 (persistent + prefetch)";  date +%T
MB_AB_SHAPES=3,5,6,7 timeout 300 python tools/ab_bench.py 8 5 sortonly 2>&1 | tee $OUT/r2a_ab_8gbp_shapes.jsonl | cut -c1-420
echo "== c. bench 100 Gbp: default, then the e2e leg with the copy-back overlapped";  date +%T
timeout 420 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --cli-gbp 0 > $OUT/r2a_bench_100gbp.json 2> $OUT/r2a_bench_100gbp.err; cut -c1-2200 $OUT/r2a_bench_100gbp.json; tail -2 $OUT/r2a_bench_100gbp.err
MB_OVERLAP_D2H=1 timeout 420 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --cli-gbp 0 > $OUT/r2a_bench_100gbp_overlap.json 2> $OUT/r2a_bench_100gbp_overlap.err; cut -c1-2200 $OUT/r2a_bench_100gbp_overlap.json
echo "== d. ncu: launch list of the default kernels (1 Gbp, 5 passes)";  date +%T
MB_MAX_PASS_KEYS=177000000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/r2a_launches.csv \
    python bench.py --gbp 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2a_ncu_launch.log 2>&1
echo "== d2. ncu --set full with source: sweep / digit histogram / count (0.25 Gbp, one pass)";  date +%T
timeout 500 ncu --set full --import-source on --clock-control none -k regex:"sortx_sweep|sort_hist_kernel|count_suffix32_kernel" -c 8 -f -o $OUT/r2a_full_sort \
    python bench.py --gbp 0.25 --workspace-gb 3 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2a_ncu_full_sort.log 2>&1
echo "== d3. ncu --set full with source: extraction as the 100 Gbp run uses it (0.5 Gbp, 5 file-aligned passes)";  date +%T
MB_MAX_PASS_KEYS=88500000 timeout 500 ncu --set full --import-source on --clock-control none -k regex:"extract_kernel" -c 3 -f -o $OUT/r2a_full_extract \
    python bench.py --gbp 0.5 --workspace-gb 6 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2a_ncu_full_extract.log 2>&1
for f in r2a_full_sort r2a_full_extract; do
  [ -f $OUT/$f.ncu-rep ] && ncu -i $OUT/$f.ncu-rep --page raw --csv > $OUT/${f}_raw.csv 2>/dev/null
  [ -f $OUT/$f.ncu-rep ] && ncu -i $OUT/$f.ncu-rep --page source --csv > $OUT/${f}_source.csv 2>/dev/null
done
ls -la $OUT | tail -20
echo "== e. union-sum next to the reference (staging ring; then the warp-per-block decoder)";  date +%T
timeout 240 python tools/union_bench.py 0.5 2>&1 | tee $OUT/r2a_union_bench.json | cut -c1-900
MB_DECODE=warp timeout 240 python tools/union_bench.py 0.5 2>&1 | tee $OUT/r2a_union_bench_warp.json | cut -c1-900
echo "== f. the command line tools side by side on one 5 Gbp FASTQ file";  date +%T
timeout 400 python -c "
import bench, json
print(json.dumps(bench.cli_leg(5.0, 0)))" 2>&1 | tee $OUT/r2a_cli_leg.json | cut -c1-1200
date
