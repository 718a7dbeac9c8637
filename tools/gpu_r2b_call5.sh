#!/usr/bin/env bash
# re-entry call 5: K6 list regrouped by alignment count, K7 gather over the active list (parity + bench), wait policy switched for the e2e leg
set -u
mkdir -p gpurun_out
echo "== parity (K6 regrouped list, K7 gather over the list)" > gpurun_out/r2b5.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_zz_gpu_enumerate.py tests/test_zzz_gpu_enumerate_fast.py tests/test_zzzz_gpu_window.py -x -q -k "k6 or chain or window or enumerate or fast or k7" > gpurun_out/t_k6.log 2>&1; echo "   exit $?" >> gpurun_out/r2b5.log; tail -3 gpurun_out/t_k6.log >> gpurun_out/r2b5.log
B="python bench.py --loci 600000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu"
run() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/r2b5.log; timeout 400 env "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2b5.log;
        python - "$name" <<'PY' >> gpurun_out/r2b5.log
import json,sys
try:
    l=[x for x in open(f"gpurun_out/{sys.argv[1]}.log") if x.startswith("{")][-1]; d=json.loads(l)
    e=d.get("e2e") or {}
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step; e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),1), "cpu_s", e.get("host_cpu_seconds_per_step_rank0"), d["config"].get("host_wait"), {k:round(v) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print("   no line", e)
PY
}
run k6_plain SX_K6_NO_CLASS_SORT=1 $B
run k6_class X=1 $B
E="python bench.py --loci 300000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-cpu"
run e2e_free X=1 $E
run e2e_free_sw SX_BLOCKING_WAIT=e2e $E
run e2e_2cpu_sw SX_BLOCKING_WAIT=e2e taskset -c 0,1 $E
cat gpurun_out/r2b5.log
