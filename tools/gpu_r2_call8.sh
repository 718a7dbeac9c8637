#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c8.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c8.log; }
step t_k6k2a 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_enumerate.py -q -x -k "k6 or chain or k2a"
step t_window 900 python -m pytest tests/test_zzzz_gpu_window.py -q -x
step b_list 600 python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --no-e2e
step b_staged 600 env SX_K6_STAGED=1 python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --no-e2e
for k in 8 4 2; do step b_k7l$k 600 env SX_K7_LOCAL_BLOCKS_PER_SM=$k python bench.py --loci 200000 --steps 2 --warmup 1 --no-legs --no-e2e; done
tail -2 gpurun_out/t_*.log
cat gpurun_out/r2c8.log
