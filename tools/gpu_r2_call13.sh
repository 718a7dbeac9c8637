#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c13.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c13.log; }
step b13_w2 900 python bench.py --loci 500000 --steps 2 --warmup 1 --no-legs --e2e-workers 2 --cpu-sample-loci 40
step b13_w3 900 python bench.py --loci 500000 --steps 2 --warmup 1 --no-legs --e2e-workers 3 --cpu-sample-loci 40
step b13_l2 600 env SX_K7_LOCAL_BLOCKS_PER_SM=8 SX_K2A_BLOCKS_PER_SM=4 python bench.py --loci 400000 --steps 2 --warmup 1 --no-legs --no-e2e --lanes 2
step b13_l2b 600 env SX_K2A_BLOCKS_PER_SM=4 python bench.py --loci 400000 --steps 2 --warmup 1 --no-legs --no-e2e --lanes 2
cat gpurun_out/r2c13.log
