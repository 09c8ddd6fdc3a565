#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "attention or gemm" --timeout=600 -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1; echo "pytest attn/gemm exit $?"; tail -5 gpurun_out/pytest_attn.log
timeout 600 python tools/microbench.py attn > gpurun_out/microbench_attn.jsonl 2>&1; cut -c1-170 gpurun_out/microbench_attn.jsonl
timeout 600 python tools/prefill_breakdown.py 32 4 > gpurun_out/prefill_breakdown_b32.log 2>&1; echo "breakdown exit $?"; tail -10 gpurun_out/prefill_breakdown_b32.log | head -9
