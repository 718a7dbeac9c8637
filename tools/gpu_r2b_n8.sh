#!/usr/bin/env bash
# 8 ranks on one box, a short run (300k loci per GPU): does the end-to-end leg scale with blocking host waits under the 16-CPU quota?
set -u
mkdir -p gpurun_out
{ cat /sys/fs/cgroup/memory.max /sys/fs/cgroup/cpu.max; free -g | head -2; nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; } > gpurun_out/n8_env.txt 2>&1
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --loci 300000 --steps 3 --warmup 1 --no-legs > gpurun_out/b_n8.log 2> gpurun_out/b_n8.err
echo "exit $?" >> gpurun_out/n8_env.txt
cat gpurun_out/n8_env.txt
python - <<'PY'
import json
try:
    l=[x for x in open("gpurun_out/b_n8.log") if x.startswith("{")][-1]; d=json.loads(l)
    e=d.get("e2e") or {}
    print(round(d["value"]), "loci/s", round(d["ms_per_step"],1), "ms/step; e2e", round(e.get("value",0)), round(e.get("ms_per_step",0),1), "cpu_s", e.get("host_cpu_seconds_per_step_rank0"), d["config"].get("host_wait"), d["config"].get("gen_seconds"))
except Exception as ex: print("no line", ex)
PY
tail -5 gpurun_out/b_n8.err
