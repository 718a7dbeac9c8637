#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c4.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c4.log; }
step b_tiny 300 python bench.py --config tiny --steps 2 --warmup 1 --no-legs
step b_ref 300 python bench.py --impl reference --steps 3 --warmup 1
step b_full 900 python bench.py --steps 3 --warmup 2 --no-legs
step legs_k2b 300 python tools/site_legs.py k2b
step legs_k5 300 python tools/site_legs.py k5
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_whole_launches.csv python bench.py --loci 100000 --steps 1 --warmup 1 --no-legs --no-e2e
tail -c 3000 gpurun_out/b_full.log; tail -3 gpurun_out/b_*.err
cat gpurun_out/r2c4.log
