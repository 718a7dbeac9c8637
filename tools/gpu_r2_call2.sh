#!/usr/bin/env bash
# Round 2, GPU call 2: the new paths first (window pass, chain vs reference, K4 buffer order), then the whole GPU suite.
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c2.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c2.log; }
step t_k4 300 python -m pytest tests/test_gpu_parity.py -q -x -k "k4 or k6"
step t_window 1200 python -m pytest tests/test_zzzz_gpu_window.py -q -x
step t_chain_ref 600 python -m pytest tests/test_zz_gpu_enumerate.py -q -x -k "references_realignAndScoreRead"
step t_all 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_zzzz_gpu_window.py
tail -5 gpurun_out/t_*.log
cat gpurun_out/r2c2.log
