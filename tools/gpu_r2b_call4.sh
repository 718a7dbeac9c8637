#!/usr/bin/env bash
# re-entry call 4: K7 active list regrouped by shape class (parity + bench with / without), host wait policy under CPU oversubscription
set -u
mkdir -p gpurun_out
echo "== K7 class-sorted list parity" > gpurun_out/r2b4.log
timeout 1200 python -m pytest tests/test_zz_gpu_enumerate.py tests/test_zzz_gpu_enumerate_fast.py tests/test_zzzz_gpu_window.py -x -q > gpurun_out/t_k7.log 2>&1; echo "   exit $?" >> gpurun_out/r2b4.log; tail -3 gpurun_out/t_k7.log >> gpurun_out/r2b4.log
B="python bench.py --loci 600000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu"
run() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/r2b4.log; timeout 400 env "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2b4.log;
        python - "$name" <<'PY' >> gpurun_out/r2b4.log
import json,sys
try:
    l=[x for x in open(f"gpurun_out/{sys.argv[1]}.log") if x.startswith("{")][-1]; d=json.loads(l)
    e=d.get("e2e") or {}
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step; e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),1), {k:round(v) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print("   no line", e)
PY
}
run k7_plain SX_K7_NO_CLASS_SORT=1 $B
run k7_class X=1 $B
E="python bench.py --loci 300000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-cpu"
run e2e_free X=1 $E
run e2e_2cpu_spin X=1 taskset -c 0,1 $E
run e2e_2cpu_block SX_BLOCKING_WAIT=1 taskset -c 0,1 $E
run e2e_free_block SX_BLOCKING_WAIT=1 $E
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name 'regex:k7_search_local|k7_class' -c 4 -f -o gpurun_out/r2b_k7c python bench.py --loci 50000 --tile-loci 50000 --steps 1 --warmup 0 --no-legs --no-e2e --no-cpu > gpurun_out/n_k7c.log 2>&1
cat gpurun_out/r2b4.log
