#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c7.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c7.log; }
step t_k6 600 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_enumerate.py -q -x -k "k6 or chain"
step b_l1 600 python bench.py --loci 300000 --steps 2 --warmup 1 --no-legs --lanes 1
step ncu_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_whole_launches2.csv python bench.py --loci 100000 --steps 1 --warmup 0 --no-legs --no-e2e --lanes 1
step ncu_full 900 ncu --set full --clock-control none --import-source on -k regex:"k2a_germline12_kernel|k6_score_kernel|k4_fill_kernel|k7_search_local_kernel|k7_search_arena_kernel|k7_gather_kernel|k7a_count_kernel|k8_region_kernel" -c 9 -o gpurun_out/r2_whole_full2 python bench.py --loci 50000 --tile-loci 50000 --steps 1 --warmup 0 --no-legs --no-e2e --lanes 1
tail -3 gpurun_out/t_k6.log
cat gpurun_out/r2c7.log
