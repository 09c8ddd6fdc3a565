#!/bin/bash
# First GPU bring-up: every stage in its own process (a trapped kernel poisons the CUDA context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; free -g | head -2 >> gpurun_out/gpu.txt
timeout 300 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1; echo "gemm_diag exit $?"
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "gemm" -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "not gemm" -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1; echo "pytest ops exit $?"
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_pipeline.log 2>&1; echo "pytest pipeline exit $?"
timeout 600 python tools/microbench.py > gpurun_out/microbench.log 2>&1; echo "microbench exit $?"
tail -3 gpurun_out/gemm_diag.log; tail -3 gpurun_out/pytest_gemm.log; tail -3 gpurun_out/pytest_ops.log; tail -3 gpurun_out/pytest_pipeline.log; tail -5 gpurun_out/microbench.log
