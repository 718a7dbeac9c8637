#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c16.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c16.log; }
for k in 16 24 32; do step b16_$k 600 env SX_K7_MIN_BLOCKS=$k python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --no-e2e; done
step t16_enum 900 env SX_K7_MIN_BLOCKS=32 python -m pytest tests/test_zz_gpu_enumerate.py -q -x
cat gpurun_out/r2c16.log
