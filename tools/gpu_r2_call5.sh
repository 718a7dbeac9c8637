#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c5.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c5.log; }
step t_k7 600 python -m pytest tests/test_zz_gpu_enumerate.py tests/test_zzz_gpu_enumerate_fast.py -q -x
step t_window 900 python -m pytest tests/test_zzzz_gpu_window.py -q -x -k "synthetic or reference-0 or reference-7"
step t_demo 600 python -m pytest tests/test_zzzzz_gpu_demo_vcf.py -q -x
step b_200k 600 python bench.py --loci 200000 --steps 2 --warmup 1 --no-legs --no-e2e
step ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"k2a_germline4_kernel|k6_score_kernel|k4_fill_kernel|k7_search_local_kernel|k7_search_arena_kernel|k1_score_kernel|k1q_score_kernel" -c 8 -o gpurun_out/r2_whole_full python bench.py --loci 50000 --tile-loci 50000 --steps 1 --warmup 0 --no-legs --no-e2e
tail -c 1500 gpurun_out/b_200k.log; tail -5 gpurun_out/t_*.log
cat gpurun_out/r2c5.log
