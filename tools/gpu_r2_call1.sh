#!/usr/bin/env bash
# Round 2, GPU call 1: evidence for the chain kernels the round-1 verdict asked for (launch list + full captures of the K7 fast plan,
# K9, K6-in-chain, K7b), before any change.  Every step bounded; outputs under gpurun_out/.
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c1.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c1.log; }
nvidia-smi -q | grep -iE "product name|numa|cpu affinity" > gpurun_out/r2c1_env.txt 2>&1
nproc >> gpurun_out/r2c1_env.txt; python -c "import os;print('affinity',len(os.sched_getaffinity(0)),'cpu_count',os.cpu_count())" >> gpurun_out/r2c1_env.txt
cat /sys/fs/cgroup/cpu.max >> gpurun_out/r2c1_env.txt 2>&1; lscpu | head -30 >> gpurun_out/r2c1_env.txt 2>&1; numactl -H >> gpurun_out/r2c1_env.txt 2>&1
step ncu_chain_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_chain_fast_launches.csv \
    python tools/chain_leg.py 100000 30 150 6572.2 fast
step ncu_chain_full 900 ncu --set full --clock-control none --import-source on -k regex:"k7_search_local_kernel|k7_gather_kernel|k9_choose_kernel|k6_score_kernel|k8_write_kernel|k1q_score_kernel|k1_score_kernel" -s 6 -c 8 -o gpurun_out/r2_chain_full \
    python tools/chain_leg.py 100000 30 150 6572.2 fast
cat gpurun_out/r2c1.log
