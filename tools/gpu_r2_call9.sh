#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c9.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c9.log; }
step b9 600 python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --no-e2e
step ncu9 900 ncu --set full --import-source on --clock-control none --kernel-name 'regex:k2a_germline12|k7_search_local|k7a_count|k7a_write|k8_' -c 12 -f -o gpurun_out/r2_c9 python bench.py --loci 50000 --tile-loci 50000 --steps 1 --warmup 0 --no-legs --no-e2e
cat gpurun_out/r2c9.log
