#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/microbench.py attn gemm > gpurun_out/microbench_attn_gemm.jsonl 2>&1; echo "microbench exit $?"; cut -c1-200 gpurun_out/microbench_attn_gemm.jsonl
SRGPT_ATTN_MMA_SYNC=1 timeout 300 python tools/microbench.py attn > gpurun_out/microbench_attn_mmasync.jsonl 2>&1; cut -c1-200 gpurun_out/microbench_attn_mmasync.jsonl
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout=600 -p no:cacheprovider -k batched > gpurun_out/pytest_batched.log 2>&1; echo "pytest batched exit $?"; tail -3 gpurun_out/pytest_batched.log
