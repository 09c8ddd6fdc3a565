#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
echo "bench n2 exit $?"
wc -c gpurun_out/bench_n2.json gpurun_out/bench_n2.err
tail -c 1500 gpurun_out/bench_n2.err
head -c 1500 gpurun_out/bench_n2.json
