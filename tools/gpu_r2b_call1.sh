#!/usr/bin/env bash
# re-entry call 1: baseline of the restored build + one full capture (with source) of the step's kernels at 50k loci
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2b1.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2b1.log; }
step b_base 600 python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --no-e2e
step f_full 900 ncu --set full --import-source on --clock-control none --kernel-name 'regex:k2a_germline12|k7_search_local|k4_fill|k6_score_list|k1_score_kernel|k9_choose|k7a_count|k8_write|k7_gather' -c 9 -f -o gpurun_out/r2b_full python bench.py --loci 50000 --tile-loci 50000 --steps 1 --warmup 0 --no-legs --no-e2e
ls -la gpurun_out
cat gpurun_out/r2b1.log
tail -c 1500 gpurun_out/b_base.log
