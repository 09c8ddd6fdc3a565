#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -m gpu -q --timeout=300 -k "decode or pipeline or stages or eos or forward or text_only or lm_head or gemv" -p no:cacheprovider > gpurun_out/pytest_dec.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_dec.log
timeout 600 python tools/decode_trace.py > gpurun_out/decode_trace.log 2>&1; echo "trace exit $?"; tail -9 gpurun_out/decode_trace.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','roofline')})"; tail -3 gpurun_out/bench_n1.err
