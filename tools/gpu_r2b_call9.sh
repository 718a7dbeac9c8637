#!/usr/bin/env bash
# re-entry call 9: host threads of the end-to-end leg
set -u
mkdir -p gpurun_out
: > gpurun_out/r2b9.log
run() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/r2b9.log; timeout 400 env "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2b9.log;
        python - "$name" <<'PY' >> gpurun_out/r2b9.log
import json,sys
try:
    l=[x for x in open(f"gpurun_out/{sys.argv[1]}.log") if x.startswith("{")][-1]; d=json.loads(l)
    e=d.get("e2e") or {}
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step; e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),1), "cpu_s", e.get("host_cpu_seconds_per_step_rank0"))
except Exception as e: print("   no line", e)
PY
}
E="python bench.py --loci 600000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-cpu"
run w2 X=1 $E --e2e-workers 2
run w3 X=1 $E --e2e-workers 3
run w4 X=1 $E --e2e-workers 4
run w6 X=1 $E --e2e-workers 6
cat gpurun_out/r2b9.log
