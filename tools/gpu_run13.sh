#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1; echo "gemm_diag exit $?"; tail -12 gpurun_out/gemm_diag.log
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "gemm" -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -5 gpurun_out/pytest_gemm.log
echo "== gemm auto (clusters)"; timeout 600 python tools/microbench.py gemm 2>&1 | cut -c1-160 | tee gpurun_out/microbench_gemm_cl2.log
echo "== gemm CL=1"; SRGPT_GEMM_CL=1 timeout 600 python tools/microbench.py gemm 2>&1 | cut -c1-160 | tee gpurun_out/microbench_gemm_cl1.log
