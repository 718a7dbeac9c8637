#!/usr/bin/env bash
# re-entry call 7: K7 local-tier launch shape again, now that its warps are of one class (resident blocks, local slots)
set -u
mkdir -p gpurun_out
: > gpurun_out/r2b7.log
B="python bench.py --loci 600000 --tile-loci 100000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu"
run() { local name=$1; shift; echo "== $name: $*" >> gpurun_out/r2b7.log; timeout 400 env "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" >> gpurun_out/r2b7.log;
        python - "$name" <<'PY' >> gpurun_out/r2b7.log
import json,sys
try:
    l=[x for x in open(f"gpurun_out/{sys.argv[1]}.log") if x.startswith("{")][-1]; d=json.loads(l)
    print("   ", round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step", {k:round(v) for k,v in d["kernel_ms_per_step"].items()})
except Exception as e: print("   no line", e)
PY
}
run base X=1 $B
run mb16 SX_K7_MIN_BLOCKS=16 $B
run mb32 SX_K7_MIN_BLOCKS=32 $B
run lb12 SX_K7_LOCAL_BLOCKS_PER_SM=12 $B
run lb16 SX_K7_LOCAL_BLOCKS_PER_SM=16 $B
run la16 SX_K7_LOCAL_ALNS=16 $B
run tile50 X=1 python bench.py --loci 600000 --tile-loci 50000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu
run tile200 X=1 python bench.py --loci 600000 --tile-loci 200000 --steps 2 --warmup 1 --no-legs --no-e2e --no-cpu
cat gpurun_out/r2b7.log
