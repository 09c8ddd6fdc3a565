#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q --timeout=120 -k "gemv or lm_head or decode or gemm_epilogues" -p no:cacheprovider > gpurun_out/pytest_ops.log 2>&1; echo "pytest ops exit $?"; tail -3 gpurun_out/pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_pipeline.py -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_pipeline.log 2>&1; echo "pytest pipeline exit $?"; tail -3 gpurun_out/pytest_pipeline.log
timeout 600 python tools/microbench.py gemv > gpurun_out/microbench_gemv.log 2>&1; echo "microbench exit $?"; cat gpurun_out/microbench_gemv.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -c 2500 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
