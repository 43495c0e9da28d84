#!/bin/bash
#  Round 2, 8-GPU call (charged 8x: keep it short): multi-GPU byte parity at world 8 on real GPUs, then one bench line
#  each for BASELINE.json configs[1] (k=21, 100 Gbp), configs[2] (k=31, 100 Gbp) and configs[3] (k=51, 50 Gbp of 15 kb reads).
#  usage (from the repo root):  gpurun --gpus 8 --timeout 1200 -- 'bash tools/gpu_r2_multi8.sh'
set -u
N=8
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/r2_multi_8.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -$N; nproc; free -g | head -2
echo "== a. multi-GPU byte parity at world 8";  date +%T
MB_TEST_WORLDS=8 timeout 700 python -m pytest tests/test_gpu_multi.py -v -p no:cacheprovider -k "not nccl-u64-51" 2>&1 | grep -E "PASS|FAIL|ERROR|passed|failed" | tee $OUT/r2_multi_8_pytest.log | tail -30
run() {   #  name, bench args
  local name=$1; shift
  echo "== bench $name";  date +%T
  timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --no-cpu-baseline "$@" > $OUT/r2_bench_8gpu_$name.json 2> $OUT/r2_bench_8gpu_$name.err
  grep '^{' $OUT/r2_bench_8gpu_$name.json | cut -c1-2400; tail -3 $OUT/r2_bench_8gpu_$name.err | cut -c1-400
}
run c2_k21 --steps 3 --warmup 2
run c3_k31 --steps 2 --warmup 1 --k 31 --no-e2e
run c4_k51 --steps 2 --warmup 1 --k 51 --read-len 15000 --err 0.005 --gbp 50 --no-e2e
date
