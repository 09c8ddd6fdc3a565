#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "attention" --timeout=300 -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1; echo "pytest attn exit $?"; tail -25 gpurun_out/pytest_attn.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest gpu exit $?"; tail -8 gpurun_out/pytest_gpu_full.log
timeout 600 python tools/prefill_breakdown.py 32 4 > gpurun_out/prefill_breakdown_b32.log 2>&1; echo "breakdown exit $?"; tail -12 gpurun_out/prefill_breakdown_b32.log | head -10
