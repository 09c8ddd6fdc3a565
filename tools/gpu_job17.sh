#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 2 --steps 1 --warmup 3 --no-c3 > gpurun_out/r02_bench_n2b.json 2> gpurun_out/r02_bench_n2b.err; echo "bench N=2 exit $?"
grep -c "NCCL INFO" gpurun_out/r02_bench_n2b.err; grep -m4 "nranks" gpurun_out/r02_bench_n2b.err | cut -c1-220; grep -c "^{" gpurun_out/r02_bench_n2b.json; wc -l gpurun_out/r02_bench_n2b.json
