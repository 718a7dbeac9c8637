#!/usr/bin/env bash
# the round's closing single-GPU call: whole GPU suite, smoke, both bench arms at default flags, launch list, one full capture
set -u
mkdir -p gpurun_out
: > gpurun_out/r2bfinal.log
step() { local name=$1 secs=$2; shift 2; echo "== $name" >> gpurun_out/r2bfinal.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2bfinal.log; }
step f_tests 1500 python -m pytest tests -q -x -m gpu
step f_smoke 600 python -c "import __graft_entry__ as g; g.smoke()"
step f_ref 600 python bench.py --impl reference --steps 2 --warmup 1
step f_bench 1200 python bench.py
step f_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2b_launches_whole_path.csv python bench.py --loci 200000 --steps 1 --warmup 1 --no-legs --no-e2e --no-cpu
step f_full 900 ncu --set full --import-source on --clock-control none --kernel-name 'regex:k2a_germline12|k7_search_local|k4_fill|k6_score_list|k7_gather|k1_score_kernel' -c 6 -f -o gpurun_out/r2b_final python bench.py --loci 100000 --tile-loci 100000 --steps 1 --warmup 0 --no-legs --no-e2e --no-cpu
tail -n 3 gpurun_out/f_tests.log >> gpurun_out/r2bfinal.log
cat gpurun_out/r2bfinal.log
