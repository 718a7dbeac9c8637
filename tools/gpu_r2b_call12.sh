#!/usr/bin/env bash
# re-entry call 12: K2b's float strand log-sum on the expf / log1p / logf restatements: strand states and strandBias bit for bit
set -u
mkdir -p gpurun_out
: > gpurun_out/r2b12.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_zzzzz_gpu_demo_vcf.py -x -q -k "k2b or somatic or demo or cfg3" > gpurun_out/t_k2b.log 2>&1; echo "k2b parity exit $?" >> gpurun_out/r2b12.log; tail -3 gpurun_out/t_k2b.log >> gpurun_out/r2b12.log
timeout 300 python tools/site_legs.py k2b > gpurun_out/k2b_leg2.log 2>&1; tail -1 gpurun_out/k2b_leg2.log | cut -c1-420 >> gpurun_out/r2b12.log
cat gpurun_out/r2b12.log
