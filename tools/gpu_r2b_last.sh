#!/usr/bin/env bash
# the last single-GPU call of the round: whole GPU suite + smoke on the final build (K2b's float log-sum mirrors included)
set -u
mkdir -p gpurun_out
: > gpurun_out/r2blast.log
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/l_tests.log 2> gpurun_out/l_tests.err; echo "tests exit $?" >> gpurun_out/r2blast.log; tail -2 gpurun_out/l_tests.log >> gpurun_out/r2blast.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/r2blast.log
cat gpurun_out/r2blast.log
