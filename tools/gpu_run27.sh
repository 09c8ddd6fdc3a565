#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest gpu exit $?"; tail -4 gpurun_out/pytest_gpu_full.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks')}); print(d['prefill']); print(d['roofline'])"; tail -3 gpurun_out/bench_n1.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair_kernel -s 120 -c 1 -o gpurun_out/gemm_pair -f python tools/microbench.py gemm > gpurun_out/ncu_gemm_pair.log 2>&1; echo "ncu pair exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm|attn|norm|rope|splice|mask|pool|patchify|argmax|gemv|lm_head|depth' -s 700 -c 700 --csv --log-file gpurun_out/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-c3 > gpurun_out/ncu_bench.log 2>&1; echo "ncu bench list exit $?"; wc -l gpurun_out/launches_bench.csv
