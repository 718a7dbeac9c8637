#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zzzz_gpu_window.py -x -q -k "regrouped or synthetic" > gpurun_out/t_rg.log 2>&1; echo "exit $?"; tail -3 gpurun_out/t_rg.log
