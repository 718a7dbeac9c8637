#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c10.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c10.log; }
step t10_k2a 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_enumerate.py -q -x -k "k2a or germline or chain or link or k7b"
step t10_window 900 python -m pytest tests/test_zzzz_gpu_window.py -q -x
step b10 600 python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --no-e2e
tail -n 3 gpurun_out/t10_k2a.log gpurun_out/t10_window.log
cat gpurun_out/r2c10.log
