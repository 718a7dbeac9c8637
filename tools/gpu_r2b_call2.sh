#!/usr/bin/env bash
# re-entry call 2: two contexts with staggered windows and capped persistent grids (K7 beside K2a), and full captures of K2b / K5
set -u
mkdir -p gpurun_out
B="python bench.py --loci 600000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu"
run() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/r2b2.log; timeout 300 env "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2b2.log;
        python - "$name" <<'PY' >> gpurun_out/r2b2.log
import json,sys
try:
    l=[x for x in open(f"gpurun_out/{sys.argv[1]}.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step", {k:round(v) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print("   no line", e)
PY
}
echo "== parity (K2a grouping up to 64 calls, K7 one-word segments)" >> gpurun_out/r2b2.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zzz_gpu_enumerate_fast.py tests/test_zzzz_gpu_window.py -x -q -k "k2a or germline or fast or window" > gpurun_out/t_parity.log 2>&1; echo "   exit $?" >> gpurun_out/r2b2.log; tail -2 gpurun_out/t_parity.log >> gpurun_out/r2b2.log
run l1 X=1 $B --lanes 1
run l2_o0 X=1 $B --lanes 2
run l2_o0_k5 SX_K2A_BLOCKS_PER_SM=5 SX_K7_LOCAL_BLOCKS_PER_SM=9 $B --lanes 2
run l2_o50_k5 SX_K2A_BLOCKS_PER_SM=5 SX_K7_LOCAL_BLOCKS_PER_SM=9 $B --lanes 2 --lane-offset-ms 50
run l2_o50_k6 SX_K2A_BLOCKS_PER_SM=6 SX_K7_LOCAL_BLOCKS_PER_SM=6 $B --lanes 2 --lane-offset-ms 50
run l2_o50_k4 SX_K2A_BLOCKS_PER_SM=4 SX_K7_LOCAL_BLOCKS_PER_SM=12 $B --lanes 2 --lane-offset-ms 50
run l2_o50_full X=1 $B --lanes 2 --lane-offset-ms 50
run l3_o35_k5 SX_K2A_BLOCKS_PER_SM=5 SX_K7_LOCAL_BLOCKS_PER_SM=9 $B --lanes 3 --lane-offset-ms 35
echo "== ncu k2b" >> gpurun_out/r2b2.log
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name 'regex:k2b_somatic' -s 3 -c 1 -f -o gpurun_out/r2b_k2b python tools/site_legs.py k2b > gpurun_out/n_k2b.log 2>&1
timeout 400 ncu --set full --import-source on --clock-control none --kernel-name 'regex:k5_indel_gl' -s 3 -c 1 -f -o gpurun_out/r2b_k5 python tools/site_legs.py k5 > gpurun_out/n_k5.log 2>&1
cat gpurun_out/r2b2.log
