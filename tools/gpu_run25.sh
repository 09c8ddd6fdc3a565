#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_configs.py -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_pipe.log 2>&1; echo "pytest pipeline exit $?"; tail -3 gpurun_out/pytest_pipe.log
timeout 600 python tools/microbench.py gemm 2>&1 | tail -9 | cut -c1-170
timeout 600 python tools/prefill_breakdown.py 32 4 > gpurun_out/prefill_breakdown_b32.log 2>&1; echo "breakdown exit $?"; tail -10 gpurun_out/prefill_breakdown_b32.log | head -9
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e')}); print(d['prefill'])"; tail -3 gpurun_out/bench_n1.err
