#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest gpu exit $?"; tail -4 gpurun_out/pytest_gpu_full.log
timeout 600 python tools/decode_trace.py > gpurun_out/decode_trace.log 2>&1; echo "trace exit $?"; grep -E "decode step|qkv_rope |attn |o_proj |gateup |down |lm_head  " gpurun_out/decode_trace.log | tail -7
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','prefill')}); print(d['roofline'])"; tail -3 gpurun_out/bench_n1.err
