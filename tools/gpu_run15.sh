#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gemm_diag.py > gpurun_out/gemm_diag.log 2>&1; echo "gemm_diag exit $?"; tail -4 gpurun_out/gemm_diag.log
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=300 -k "gemm" -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -5 gpurun_out/pytest_gemm.log
echo "== gemm (tall for M=259)"; timeout 600 python tools/microbench.py gemm 2>&1 | grep " 259x" | cut -c1-170 | tee gpurun_out/microbench_gemm_tall.log
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest gpu exit $?"; tail -4 gpurun_out/pytest_gpu_full.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','prefill')}); print(d['roofline'])"; tail -3 gpurun_out/bench_n1.err
