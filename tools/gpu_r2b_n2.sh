#!/usr/bin/env bash
# two ranks at the default size (1M loci per GPU), final build; both arms
set -u
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 1 > gpurun_out/b_n2.log 2> gpurun_out/b_n2.err
echo "exit $?"
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/b_n2.log") if x.startswith("{")][-1]; d=json.loads(l)
    e=d.get("e2e") or {}
    print(round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step; e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),1), "cpu_s", e.get("host_cpu_seconds_per_step_rank0"), d["config"].get("host_wait"), d["config"].get("gen_seconds"))
except Exception as ex: print("no line", ex)
PY
grep -c "Error\|Traceback" gpurun_out/b_n2.err
