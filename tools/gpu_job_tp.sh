#!/bin/bash
# N-GPU job: tensor-parallel decode with the fused NVLink collectives: parity (TP-N ids == TP-1 ids), then the c5-shaped bench.  usage: gpu_job_tp.sh N
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 tools/tp_run.py --check --comm p2p > gpurun_out/r02_tp${N}_check_p2p.log 2>&1; echo "check p2p exit $?"; grep '"mode"' gpurun_out/r02_tp${N}_check_p2p.log | cut -c1-200; tail -3 gpurun_out/r02_tp${N}_check_p2p.log | cut -c1-300
timeout 420 $TR --master-port 29553 tools/tp_run.py --bench --new 512 --comm p2p > gpurun_out/r02_tp${N}_bench_p2p.log 2>&1; echo "bench p2p exit $?"; grep '"mode"' gpurun_out/r02_tp${N}_bench_p2p.log | cut -c1-900
