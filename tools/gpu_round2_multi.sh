#!/bin/bash
#  Multi-GPU call of the next round: the two multi-GPU forms written without GPUs at hand (u32 suffix keys routed by bin,
#  extraction fused with the exchange over peer memory) get their parity tests and an A/B against the default
#  (u64 keys, NCCL all-to-all), at N GPUs of one box.
#  usage (from the repo root):  gpurun --gpus 2 --timeout 1500 -- 'bash tools/gpu_round2_multi.sh 2'
#  (a call on N GPUs is charged N x its box time: develop at N=2, confirm the winner once at N=8 with a short call, e.g.
#   gpurun --gpus 8 --timeout 600 -- 'bash tools/gpu_round2_multi.sh 8 quick' -- parity tests and two bench lines only)
set -u
N=${1:-2}
QUICK=${2:-}
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/round2_multi_$N.log) 2>&1
date; nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -$N
echo "== a. multi-GPU parity, experimental forms included";  date +%T
MB_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -5
run() {   #  name, env...
  local name=$1; shift
  echo "== bench $name";  date +%T
  env "$@" timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) \
      bench.py --gpus $N --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_${N}gpu_$name.json 2> $OUT/bench_${N}gpu_$name.err
  cut -c1-1800 $OUT/bench_${N}gpu_$name.json; tail -3 $OUT/bench_${N}gpu_$name.err
}
run default     MB_NOP=1
run suffix32p2p MB_MULTI_SUFFIX32=1 MB_P2P_SCATTER=1
if [ -n "$QUICK" ]; then date; exit 0; fi
run suffix32    MB_MULTI_SUFFIX32=1
run p2p         MB_P2P_SCATTER=1
#  the default all-to-all moved 75.8 GB per GPU at ~270 GB/s in round 1: does NCCL want more point-to-point channels?
run default_p2pch32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
date
