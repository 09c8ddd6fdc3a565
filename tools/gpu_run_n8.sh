#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${1:-8}
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench n$N exit $?"; wc -c gpurun_out/bench_n$N.json; tail -c 600 gpurun_out/bench_n$N.err; python -c "
import json,sys; d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','gpu_launches')})"
