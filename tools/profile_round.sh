#!/bin/bash
#  Run on the GPU box (under gpurun): ncu launch list + section capture of the hot kernels.
#  Outputs land in gpurun_out/; tools/summarize_ncu.py turns them into the CSVs kept under profiles/.
#  usage: tools/profile_round.sh <tag>
set -u
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches_$TAG.csv \
    python bench.py --gbp 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/ncu_launch_$TAG.log 2>&1
#  a 3 GB arena keeps ncu's save/restore of device memory between replays short
timeout 420 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section Occupancy --section WarpStateStats \
    --section LaunchStats --section SchedulerStats --clock-control none \
    -k regex:"sort_sweep|sortx_sweep|extract_kernel|count_suffix32|fields_kernel|sort_hist" -c 12 \
    -o $OUT/prof_$TAG python bench.py --gbp 0.1 --workspace-gb 3 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e > $OUT/ncu_sections_$TAG.log 2>&1
tail -2 $OUT/ncu_sections_$TAG.log | cut -c1-200
