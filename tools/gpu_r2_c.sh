#!/bin/bash
#  Round 2, GPU call C (one B200, ~25 min): the in-place group-wise sort (3 passes instead of 5 at 100 Gbp), the dense
#  extraction at four CTAs per SM, the headline-plan parity test, L2 prefetch distances of the sweep, the device parser next to
#  the host parser, union-sum at the size of BASELINE.json configs[4], launch lists of the k = 31 and k = 51 kernels.
#  usage (from the repo root):  gpurun --timeout 2700 -- 'bash tools/gpu_r2_c.sh'
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/r2c.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader; nproc; free -g | head -2
echo "== a. the GPU suite, experimental tests included, no -x";  date +%T
MB_TEST_EXPERIMENTAL=1 timeout 1800 python -m pytest tests -q -m gpu -p no:cacheprovider 2>&1 | tail -12
echo "== b. A/B at 8 Gbp in 5 passes (one process, same reads)";  date +%T
MB_AB_ENVS='[{}, {"MB_SWEEP_RUNS": "1"}, {"MB_SORT_PINGPONG": "1"}, {"MB_SWEEP_PREFETCH": "444"}, {"MB_SWEEP_PREFETCH": "1332"}, {"MB_SWEEP_RUNS": "1", "MB_SWEEP_PREFETCH": "444"}, {"MB_COUNT_GROUP": "1"}, {"MB_EXTRACT_DENSE": "0"}]' \
  timeout 400 python tools/ab_bench.py 8 5 env 2>&1 | tee $OUT/r2c_ab_8gbp_env.jsonl | cut -c1-420
echo "== b2. parity with the copy-out by digit runs (MB_SWEEP_RUNS=1): sort stages + the headline plan";  date +%T
MB_SWEEP_RUNS=1 timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_headline.py -q -p no:cacheprovider 2>&1 | tail -3
echo "== c0. bench 100 Gbp with MB_SWEEP_RUNS=1 (device-resident only)";  date +%T
MB_SWEEP_RUNS=1 timeout 500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2c_bench_100gbp_runs.json 2> $OUT/r2c_bench_100gbp_runs.err; cut -c1-1700 $OUT/r2c_bench_100gbp_runs.json; tail -2 $OUT/r2c_bench_100gbp_runs.err
echo "== c. bench 100 Gbp: the round-1 ping-pong layout (device-resident only), then the default with all legs";  date +%T
MB_SORT_PINGPONG=1 timeout 500 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2c_bench_100gbp_pingpong.json 2> $OUT/r2c_bench_100gbp_pingpong.err; cut -c1-1700 $OUT/r2c_bench_100gbp_pingpong.json; tail -2 $OUT/r2c_bench_100gbp_pingpong.err
timeout 900 python bench.py --steps 3 --warmup 2 > $OUT/r2c_bench_100gbp.json 2> $OUT/r2c_bench_100gbp.err; cut -c1-3600 $OUT/r2c_bench_100gbp.json; tail -2 $OUT/r2c_bench_100gbp.err
echo "== d. the device parser next to the host parser; meryl-lookup -existence next to the reference";  date +%T
timeout 400 python tools/lookup_bench.py 1500 2>&1 | tee $OUT/r2c_lookup_bench.jsonl | cut -c1-400
echo "== e. union-sum at the size of BASELINE.json configs[4] (two databases of about 3 G distinct k-mers)";  date +%T
timeout 900 python tools/union_bench.py 9 2>&1 | tee $OUT/r2c_union_bench_9.json | cut -c1-1200
echo "== f. launch lists of the k = 31 (u64 keys) and k = 51 (u128 keys) kernels, 1 Gbp";  date +%T
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/r2c_launches_k31.csv \
    python bench.py --gbp 1 --k 31 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2c_ncu_launch_k31.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $OUT/r2c_launches_k51.csv \
    python bench.py --gbp 1 --k 51 --read-len 15000 --err 0.005 --steps 1 --warmup 0 --no-cpu-baseline --no-e2e --cli-gbp 0 > $OUT/r2c_ncu_launch_k51.log 2>&1
ls -la $OUT | tail -12
date
