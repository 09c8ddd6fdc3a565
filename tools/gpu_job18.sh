#!/bin/bash
mkdir -p gpurun_out
echo "box env:"; env | grep -i "nccl\|TORCH_" | head
NCCL_DEBUG=INFO timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus 2 --steps 1 --warmup 3 --no-c3 > gpurun_out/r02_bench_n2c.json 2> gpurun_out/r02_bench_n2c.err; echo "bench N=2 (NCCL_DEBUG=INFO from the caller) exit $?"
echo "stderr NCCL INFO lines: $(grep -c 'NCCL INFO' gpurun_out/r02_bench_n2c.err)"; grep -m3 "nranks" gpurun_out/r02_bench_n2c.err | cut -c1-200
echo "stdout lines: $(wc -l < gpurun_out/r02_bench_n2c.json), json lines: $(grep -c '^{' gpurun_out/r02_bench_n2c.json)"; grep -v '^{' gpurun_out/r02_bench_n2c.json | head -3 | cut -c1-200
