#!/usr/bin/env bash
# The first GPU call of a round, in the order that loses least if the call is cut short: everything here was last verified on the
# CPU only (host-compiled device bodies, see DESIGN.md "what ran where") unless the step says otherwise.  Run from the repo root:
#   gpurun --timeout 1500 -- 'bash tools/gpu_first_call.sh'
# Every step has its own time limit and writes under gpurun_out/; a failing step does not stop the later ones.
set -u
mkdir -p gpurun_out
step() { # name, seconds, command...
    local name=$1 secs=$2
    shift 2
    echo "== $name" | tee -a gpurun_out/first_call.log
    timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"
    echo "   exit $?" | tee -a gpurun_out/first_call.log
}
# 1. parity: the GPU-verified kernels first (K1..K6, K7, K7b), then the files holding the paths that have only run on the CPU
step tests_verified 900 python -m pytest tests -m gpu -x -q --ignore=tests/test_zz_gpu_enumerate.py --ignore=tests/test_zzz_gpu_enumerate_fast.py
step tests_k7_family 900 python -m pytest tests/test_zz_gpu_enumerate.py -m gpu -q
step tests_k7_fast 900 python -m pytest tests/test_zzz_gpu_enumerate_fast.py -m gpu -q
# 2. the legs, each a process of its own: K7 in both launch plans, then the device-resident chain in both
step k7_leg_original 300 python tools/k7_leg.py 200000 30 150 6572.2 original
step k7_leg_fast 300 python tools/k7_leg.py 200000 30 150 6572.2 fast
step chain_leg_original 300 python tools/chain_leg.py 100000 30 150 6572.2 original
step chain_leg_fast 300 python tools/chain_leg.py 100000 30 150 6572.2 fast
# 3. the whole bench line
step bench 900 python bench.py
# 4. launch lists (ncu serialises and replays: the shares, not the absolute times, are what these are for)
step ncu_chain_launches 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_chain_fast_launches.csv \
    python tools/chain_leg.py 20000 30 150 6572.2 fast
# 5. one full capture of the fast plan's search kernel
step ncu_k7_fast_full 600 ncu --set full --clock-control none --import-source on -k regex:k7_search_local_kernel -c 1 -o gpurun_out/r2_k7_search_local \
    python tools/k7_leg.py 50000 30 150 6572.2 fast
cat gpurun_out/first_call.log
