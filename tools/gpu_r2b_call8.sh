#!/usr/bin/env bash
# re-entry call 8: K4 plan 2 with the lean base pass + the pipelined gather: parity, bench against the windowed fill, one capture
set -u
mkdir -p gpurun_out
echo "== K4 plan 2 (lean base pass) parity" > gpurun_out/r2b8.log
SX_K4_PLAN=2 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zzzz_gpu_window.py tests/test_gpu_fullsize.py -x -q -k "k4 or pileup or window" > gpurun_out/t_k4c.log 2>&1; echo "   exit $?" >> gpurun_out/r2b8.log; tail -3 gpurun_out/t_k4c.log >> gpurun_out/r2b8.log
B="python bench.py --loci 600000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu"
run() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/r2b8.log; timeout 300 env "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2b8.log;
        python - "$name" <<'PY' >> gpurun_out/r2b8.log
import json,sys
try:
    l=[x for x in open(f"gpurun_out/{sys.argv[1]}.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step", {k:round(v) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print("   no line", e)
PY
}
run k4_new SX_K4_PLAN=2 $B
run k4_new_generic SX_K4_PLAN=2 SX_K4_GENERIC_BASES=1 $B
SX_K4_PLAN=2 timeout 600 ncu --set full --import-source on --clock-control none --kernel-name 'regex:k4_bases|k4_gather' -c 2 -f -o gpurun_out/r2b_k4c python bench.py --loci 50000 --tile-loci 50000 --steps 1 --warmup 0 --no-legs --no-e2e --no-cpu > gpurun_out/n_k4c.log 2>&1
cat gpurun_out/r2b8.log
