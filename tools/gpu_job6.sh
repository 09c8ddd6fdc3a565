#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -p no:cacheprovider -k "attention" 2>&1 | tail -8 > gpurun_out/r02_pytest6.log; tail -4 gpurun_out/r02_pytest6.log
timeout 400 python tools/ab.py --rounds 2 --cmd "python tools/microbench.py attn" A: B:SRGPT_ATTN_PP=-1 > gpurun_out/r02_ab_attn_pp2.txt 2>&1; tail -7 gpurun_out/r02_ab_attn_pp2.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_vit_pp -s 15 -c 1 -o gpurun_out/r02_attn_pp2 -f python tools/microbench.py attn > gpurun_out/ncu_attn_pp.log 2>&1; echo "ncu attn exit $?"
