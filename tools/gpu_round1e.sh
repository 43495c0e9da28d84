#!/bin/bash
#  Last gpurun call of the round (~1.5 GPU-minutes left): the operators added after call d, and the smoke entry point.
set -u
OUT=gpurun_out
mkdir -p $OUT
exec > >(tee $OUT/round1e.log) 2>&1
date
timeout 45 python -m pytest tests/test_gpu_union.py -x -q -k "set_operations or value_operations or golden" 2>&1 | tail -2
timeout 25 python -m pytest tests/test_gpu_count.py -x -q -k "compress or golden" 2>&1 | tail -2
timeout 25 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
date
