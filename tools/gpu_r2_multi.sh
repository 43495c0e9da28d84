#!/bin/bash
#  Round 2, multi-GPU call: byte parity of every multi-GPU form on real GPUs (log kept under profiles/) and an A/B of
#  the forms written without GPUs at hand (u32 suffix keys routed by bin; extraction fused with the exchange over peer
#  memory) against the round-1 default (u64 keys, NCCL all-to-all) on the 100 Gbp job.
#  usage (from the repo root):  gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_r2_multi.sh 2'
#         (N GPUs are charged N x the box time)        ... 'bash tools/gpu_r2_multi.sh 8 "winner env" extra'
set -u
export MB_EXTRACT_DENSE=${MB_EXTRACT_DENSE:-0}     #  (first multi-GPU call: the extraction form that round 1 measured)
N=${1:-2}
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/r2_multi_$N.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -$N; nvidia-smi topo -m | head -12
echo "== a. multi-GPU byte parity, experimental forms included";  date +%T
MB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_multi.py -v -p no:cacheprovider 2>&1 | grep -E "PASS|FAIL|SKIP|ERROR|passed|failed" | tee $OUT/r2_multi_${N}_pytest.log | tail -40
run() {   #  name, extra bench args, env...
  local name=$1; shift
  local extra=$1; shift
  echo "== bench $name";  date +%T
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline $extra > $OUT/r2_bench_${N}gpu_$name.json 2> $OUT/r2_bench_${N}gpu_$name.err
  cut -c1-2000 $OUT/r2_bench_${N}gpu_$name.json; tail -3 $OUT/r2_bench_${N}gpu_$name.err
}
if [ "$N" = "2" ]; then
  run default     ""       MB_NOP=1
  run suffix32p2p ""       MB_MULTI_SUFFIX32=1 MB_P2P_SCATTER=1
  run suffix32    "--no-e2e" MB_MULTI_SUFFIX32=1
  run p2p         "--no-e2e" MB_P2P_SCATTER=1
fi
date
