#!/bin/bash
# 2-GPU job: tensor-parallel decode parity (TP-2 ids == TP-1 ids) with the fused NVLink collectives and with NCCL, then the c5-shaped bench
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29541 tools/tp_run.py --check --comm p2p > gpurun_out/r02_tp2_check_p2p.log 2>&1; echo "check p2p exit $?"; grep '"mode"' gpurun_out/r02_tp2_check_p2p.log | cut -c1-300; tail -5 gpurun_out/r02_tp2_check_p2p.log | cut -c1-300
timeout 600 $TR --master-port 29542 tools/tp_run.py --check --comm nccl > gpurun_out/r02_tp2_check_nccl.log 2>&1; echo "check nccl exit $?"; grep '"mode"' gpurun_out/r02_tp2_check_nccl.log | cut -c1-300; tail -3 gpurun_out/r02_tp2_check_nccl.log | cut -c1-300
timeout 900 $TR --master-port 29543 tools/tp_run.py --bench --new 256 --comm p2p > gpurun_out/r02_tp2_bench_p2p.log 2>&1; echo "bench p2p exit $?"; grep '"mode"' gpurun_out/r02_tp2_bench_p2p.log | cut -c1-700
timeout 900 $TR --master-port 29544 tools/tp_run.py --bench --new 256 --comm nccl > gpurun_out/r02_tp2_bench_nccl.log 2>&1; echo "bench nccl exit $?"; grep '"mode"' gpurun_out/r02_tp2_bench_nccl.log | cut -c1-700
timeout 300 python tools/ab.py --rounds 2 --cmd "python tools/microbench.py attn" A: B:SRGPT_ATTN_PP=-1 > gpurun_out/r02_ab_attn_pp4.txt 2>&1; tail -6 gpurun_out/r02_ab_attn_pp4.txt | head -3
