#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
step() { local name=$1 secs=$2; shift 2; echo "== $name" | tee -a gpurun_out/r2c14.log; timeout "$secs" "$@" > "gpurun_out/$name.log" 2> "gpurun_out/$name.err"; echo "   exit $?" | tee -a gpurun_out/r2c14.log; }
step t14_par 900 python -m pytest tests/test_gpu_parity.py -q -x -k "k2a or germline or eprob"
step t14_window 900 python -m pytest tests/test_zzzz_gpu_window.py -q -x
step b14_n2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --loci 300000 --steps 2 --warmup 1 --no-legs
step r14_n2 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 2 --warmup 1
tail -n 3 gpurun_out/t14_*.log
tail -n 5 gpurun_out/b14_n2.err
cat gpurun_out/r2c14.log
