#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c11.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c11.log; }
step t11_enum 900 python -m pytest tests/test_zz_gpu_enumerate.py -q -x
step t11_window 900 python -m pytest tests/test_zzzz_gpu_window.py -q -x
step b11 900 python bench.py --loci 500000 --steps 2 --warmup 1 --no-legs
tail -n 3 gpurun_out/t11_enum.log gpurun_out/t11_window.log
cat gpurun_out/r2c11.log
