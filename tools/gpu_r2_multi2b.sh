#!/bin/bash
#  Round 2, second 2-GPU call: the multi-GPU defaults as they are now (u32 suffix keys + extraction storing into the
#  owner's HBM, arenas and peer mappings kept between jobs) -- byte parity, then bench lines for k = 21 / 31 / 51.
#  usage (from the repo root):  gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2_multi2b.sh'
set -u
N=2
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/r2_multi_2b.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -$N
echo "== a. multi-GPU byte parity at world 2 (all routes)";  date +%T
MB_TEST_WORLDS=2 timeout 600 python -m pytest tests/test_gpu_multi.py -v -p no:cacheprovider 2>&1 | grep -E "PASS|FAIL|ERROR|passed|failed" | tee $OUT/r2_multi_2b_pytest.log | tail -30
run() {   #  name, bench args
  local name=$1; shift
  echo "== bench $name";  date +%T
  MB_TRACE=1 timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --no-cpu-baseline "$@" > $OUT/r2b_bench_2gpu_$name.json 2> $OUT/r2b_bench_2gpu_$name.err
  grep '^{' $OUT/r2b_bench_2gpu_$name.json | cut -c1-2400; grep -v "peer scatter" $OUT/r2b_bench_2gpu_$name.err | tail -3 | cut -c1-400
}
run c2_k21 --steps 3 --warmup 1
run c3_k31 --steps 1 --warmup 1 --k 31 --gbp 50 --no-e2e
run c4_k51 --steps 1 --warmup 1 --k 51 --read-len 15000 --err 0.005 --gbp 12 --no-e2e
date
