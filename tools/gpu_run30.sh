#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -m gpu -q -k "attention or batched or stages" --timeout=600 -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_attn.log
timeout 600 python tools/microbench.py attn 2>&1 | cut -c1-170
timeout 600 python tools/microbench.py attn 2>&1 | cut -c1-170
