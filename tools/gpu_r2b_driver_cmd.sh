#!/usr/bin/env bash
# the driver's own command line for the one-GPU bench (steps 20, warm-up 5), timed
set -u
mkdir -p gpurun_out
s=$(date +%s)
timeout 700 python3 bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/b_drv.log 2> gpurun_out/b_drv.err
echo "exit $? after $(( $(date +%s) - s )) s"
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/b_drv.log") if x.startswith("{")][-1]; d=json.loads(l)
    e=d.get("e2e") or {}
    print(round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step; e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),1), "cpu", round(d["cpu_baseline"]["value"]), "legs", [k for k in ("scoring_only_step","k2b_somatic_cfg3","k5_indel_gl") if isinstance(d.get(k),dict) and "error" not in d[k]])
except Exception as ex: print("no line", ex)
PY
grep -c "Error\|Traceback" gpurun_out/b_drv.err; nvidia-smi --query-gpu=memory.used --format=csv,noheader
