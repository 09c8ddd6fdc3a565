#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/microbench.py maskpool > gpurun_out/microbench_maskpool.jsonl 2>&1; echo "maskpool exit $?"; cut -c1-200 gpurun_out/microbench_maskpool.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 15 -c 1 -o gpurun_out/attn_tc2 -f python tools/microbench.py attn > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 172 -c 3 -o gpurun_out/gemm_vit -f python tools/microbench.py gemm > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"; tail -3 gpurun_out/ncu_gemm.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
