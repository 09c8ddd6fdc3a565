#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest gpu exit $?"; tail -15 gpurun_out/pytest_gpu_full.log
timeout 600 python tools/prefill_breakdown.py 32 4 > gpurun_out/prefill_breakdown_b32.log 2>&1; echo "breakdown exit $?"; tail -14 gpurun_out/prefill_breakdown_b32.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','prefill')})"; tail -3 gpurun_out/bench_n1.err
