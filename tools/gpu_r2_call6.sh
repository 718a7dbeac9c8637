#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c6.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c6.log; }
step t_k2a 600 python -m pytest tests/test_gpu_parity.py -q -x -k "k2a or k4_pileup_feeds"
step t_full 900 python -m pytest tests/test_gpu_fullsize.py -q -x
step t_k7 600 python -m pytest tests/test_zz_gpu_enumerate.py -q -x
step t_window 900 python -m pytest tests/test_zzzz_gpu_window.py -q -x
step b_l1 600 python bench.py --loci 200000 --steps 2 --warmup 1 --no-legs --no-e2e --lanes 1
step b_l3 600 python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --lanes 3
step b_old_k2a 600 env SX_K2A_BATCH4=1 python bench.py --loci 200000 --steps 2 --warmup 1 --no-legs --no-e2e --lanes 1
tail -4 gpurun_out/t_*.log
cat gpurun_out/r2c6.log
