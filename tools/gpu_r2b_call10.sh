#!/usr/bin/env bash
# re-entry call 10: windows of a step on several contexts at once (resident: --lanes, end to end: --e2e-workers) at the full size; K5 with the padded tile
set -u
mkdir -p gpurun_out
: > gpurun_out/r2b10.log
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "k5 or indel_gl" > gpurun_out/t_k5.log 2>&1; echo "k5 parity exit $?" >> gpurun_out/r2b10.log; tail -1 gpurun_out/t_k5.log >> gpurun_out/r2b10.log
timeout 300 python tools/site_legs.py k5 > gpurun_out/k5_leg.log 2>&1; tail -1 gpurun_out/k5_leg.log | cut -c1-200 >> gpurun_out/r2b10.log
run() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/r2b10.log; timeout 500 env "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2b10.log;
        python - "$name" <<'PY' >> gpurun_out/r2b10.log
import json,sys
try:
    l=[x for x in open(f"gpurun_out/{sys.argv[1]}.log") if x.startswith("{")][-1]; d=json.loads(l)
    e=d.get("e2e") or {}
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step (host clock", round(d["timing"]["host_clock_ms_per_step_rank0"],1), "); e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),1), "cpu_s", e.get("host_cpu_seconds_per_step_rank0"))
except Exception as e: print("   no line", e)
PY
}
E="python bench.py --steps 3 --warmup 1 --no-legs --no-cpu"
run l1w3 X=1 $E --lanes 1 --e2e-workers 3
run l5w5 X=1 $E --lanes 5 --e2e-workers 5
run l10w10 X=1 $E --lanes 10 --e2e-workers 10
run l3w4 X=1 $E --lanes 3 --e2e-workers 4
nvidia-smi --query-gpu=memory.used,memory.total --format=csv >> gpurun_out/r2b10.log
cat gpurun_out/r2b10.log
