#!/usr/bin/env bash
# re-entry call 6: the whole GPU suite on the regrouped lists (K7, K6, K9, K7a), smoke, bench with / without the K9 and K7a regrouping
set -u
mkdir -p gpurun_out
echo "== whole GPU suite" > gpurun_out/r2b6.log
timeout 1800 python -m pytest tests -x -q -m gpu > gpurun_out/t_all.log 2>&1; echo "   exit $?" >> gpurun_out/r2b6.log; tail -3 gpurun_out/t_all.log >> gpurun_out/r2b6.log
SX_K4_PLAN=2 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_zzzz_gpu_window.py -x -q -k "k4 or pileup or window" > gpurun_out/t_k4b.log 2>&1; echo "   k4 plan 2 exit $?" >> gpurun_out/r2b6.log; tail -1 gpurun_out/t_k4b.log >> gpurun_out/r2b6.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "   smoke exit $?" >> gpurun_out/r2b6.log
B="python bench.py --loci 600000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu"
run() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/r2b6.log; timeout 400 env "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2b6.log;
        python - "$name" <<'PY' >> gpurun_out/r2b6.log
import json,sys
try:
    l=[x for x in open(f"gpurun_out/{sys.argv[1]}.log") if x.startswith("{")][-1]; d=json.loads(l)
    e=d.get("e2e") or {}
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step; e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),1), "cpu_s", e.get("host_cpu_seconds_per_step_rank0"), {k:round(v) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print("   no line", e)
PY
}
run all_on X=1 $B
run k9k7a_off SX_K9_NO_CLASS_SORT=1 SX_K7A_NO_CLASS_SORT=1 $B
cat gpurun_out/r2b6.log
