#!/usr/bin/env bash
# re-entry call 11: K2a's double epilogue on the libm mirrors (exp, log10): ref_pprob bit for bit; K2b / K5 qualities through the same log10
set -u
mkdir -p gpurun_out
: > gpurun_out/r2b11.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_zzzz_gpu_window.py tests/test_zzzzz_gpu_demo_vcf.py -x -q -k "k2a or germline or somatic or k2b or k5 or indel_gl or window or demo or libm or fullsize or cfg" > gpurun_out/t_mir.log 2>&1; echo "parity exit $?" >> gpurun_out/r2b11.log; tail -3 gpurun_out/t_mir.log >> gpurun_out/r2b11.log
timeout 400 python bench.py --loci 600000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu > gpurun_out/b_mir.log 2> gpurun_out/b_mir.err; echo "bench exit $?" >> gpurun_out/r2b11.log
python - <<'PY' >> gpurun_out/r2b11.log
import json
try:
    l=[x for x in open("gpurun_out/b_mir.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step", {k:round(v) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print("   no line", e)
PY
timeout 300 python tools/site_legs.py k2b > gpurun_out/k2b_leg.log 2>&1; tail -1 gpurun_out/k2b_leg.log | cut -c1-160 >> gpurun_out/r2b11.log
cat gpurun_out/r2b11.log
