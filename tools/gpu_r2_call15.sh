#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c15.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c15.log; }
step t15_par 900 python -m pytest tests/test_gpu_parity.py -q -x -k "k2a or germline or eprob"
step t15_window 900 python -m pytest tests/test_zzzz_gpu_window.py -q -x
step b15 600 python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --no-e2e
tail -n 3 gpurun_out/t15_*.log
cat gpurun_out/r2c15.log
