#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest gpu exit $?"; tail -6 gpurun_out/pytest_gpu_full.log
echo "== maskpool"; timeout 600 python tools/microbench.py maskpool 2>&1 | grep -E "mask_pool|mask_weights n1 M8 448->128" | cut -c1-160 | tee gpurun_out/microbench_maskpool.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','prefill')}); print(d['roofline'])"; tail -3 gpurun_out/bench_n1.err
