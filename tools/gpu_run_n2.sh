#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','e2e','scaling','clocks')})"; tail -5 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "ref n2 exit $?"; cut -c1-300 gpurun_out/bench_ref_n2.json
