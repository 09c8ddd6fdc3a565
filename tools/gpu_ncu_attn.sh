#!/bin/bash
# ncu --set full capture of the tcgen05 prefill attention kernel (64 images x 16 heads x 1024 tokens) + attention micro-benchmark
mkdir -p gpurun_out
timeout 300 python tools/microbench.py attn > gpurun_out/microbench_attn.jsonl 2>&1; cut -c1-170 gpurun_out/microbench_attn.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 15 -c 1 -o gpurun_out/attn_tc3 -f python tools/microbench.py attn > gpurun_out/ncu_attn.log 2>&1; echo "ncu exit $?"; ls -la gpurun_out/attn_tc3.ncu-rep
