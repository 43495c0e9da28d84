#!/bin/bash
#  The call round 2 had planned last and could not afford (its GPU minutes ended inside the 8-GPU call): the committed state on one
#  B200, ~22 min -- the whole GPU suite, the bench with all its legs as the driver runs it, the reference arm, smoke(), the launch
#  list and the DRAM traffic of the default sweep kernel, the A/Bs of the prefetch distance and of the dense extraction's members
#  per thread.  First call of the next round.
#  usage (from the repo root):  gpurun --timeout 2400 -- 'bash tools/gpu_next_first.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/r2f.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
echo "== a. the GPU suite, no -x";  date +%T
timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -8
echo "== b. A/B at 8 Gbp: prefetch distance of the sweep (5 passes forced); members per thread of the dense extraction (3 passes)";  date +%T
MB_AB_ENVS='[{}, {"MB_SWEEP_PREFETCH": "0"}, {"MB_SWEEP_PREFETCH": "148"}, {"MB_SWEEP_PREFETCH": "296"}, {"MB_SWEEP_PREFETCH": "592"}, {"MB_SWEEP_PREFETCH": "888"}]' \
  timeout 300 python tools/ab_bench.py 8 5 env 2>&1 | tee $OUT/r2f_ab_8gbp_prefetch.jsonl | cut -c1-400
MB_AB_ENVS='[{}, {"MB_DENSE_ITEMS": "12"}, {"MB_DENSE_ITEMS": "8"}]' \
  timeout 200 python tools/ab_bench.py 8 3 env 2>&1 | tee $OUT/r2f_ab_8gbp_dense_items.jsonl | cut -c1-400
echo "== c. bench.py as the driver runs it (fewer steps), then the reference arm";  date +%T
timeout 1200 python bench.py --steps 3 --warmup 3 > $OUT/r2f_bench_100gbp.json 2> $OUT/r2f_bench_100gbp.err; cut -c1-4000 $OUT/r2f_bench_100gbp.json; tail -2 $OUT/r2f_bench_100gbp.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/r2f_bench_reference.json 2> $OUT/r2f_bench_reference.err; cut -c1-1500 $OUT/r2f_bench_reference.json; tail -2 $OUT/r2f_bench_reference.err
echo "== d. smoke()";  date +%T
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== e. ncu: launch list of the default kernels (1 Gbp, 3 passes), DRAM traffic of the sweep (1 Gbp, one pass)";  date +%T
MB_MAX_PASS_KEYS=295000000 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $OUT/r2f_launches.csv \
    python bench.py --gbp 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2f_ncu_launch.log 2>&1
MB_GROUP_KEYS=900000000 timeout 400 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"sortx_sweep" -c 6 --csv --log-file $OUT/r2f_sweep_dram.csv \
    python bench.py --gbp 1 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2f_ncu_dram.log 2>&1
ls -la $OUT | tail -10
date
