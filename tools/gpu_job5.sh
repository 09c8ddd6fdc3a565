#!/bin/bash
# round-2 GPU job 5: ping-pong ViT attention (parity + A/B), c3 breakdown, batched-decode launch list, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider -k "attention or stages or batch" 2>&1 | tail -30 > gpurun_out/r02_pytest5.log; tail -8 gpurun_out/r02_pytest5.log
timeout 400 python tools/ab.py --rounds 2 --cmd "python tools/microbench.py attn" A: B:SRGPT_ATTN_PP=-1 > gpurun_out/r02_ab_attn_pp.txt 2>&1; tail -8 gpurun_out/r02_ab_attn_pp.txt
timeout 300 python tools/prefill_breakdown.py 32 4 > gpurun_out/r02_prefill_breakdown_b32_pp.txt 2>&1; tail -11 gpurun_out/r02_prefill_breakdown_b32_pp.txt | head -9
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_batched_decode.csv python tools/batch_decode_once.py 32 3 > gpurun_out/ncu_bd.log 2>&1; echo "ncu batched decode exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_vit_pp -s 3 -c 1 -o gpurun_out/r02_attn_pp -f python tools/microbench.py attn > gpurun_out/ncu_attn_pp.log 2>&1; echo "ncu attn exit $?"
