#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py tests/test_gpu_configs.py -m gpu -q --timeout=600 -k "gemv or lm_head or stages or real_widths or cluster" -p no:cacheprovider > gpurun_out/pytest_dec.log 2>&1; echo "pytest exit $?"; tail -3 gpurun_out/pytest_dec.log
for SP in 0 1 2; do
  echo "== SRGPT_GEMV_SPRE=$SP"
  SRGPT_GEMV_SPRE=$SP timeout 600 python tools/decode_trace.py > gpurun_out/decode_trace_spre$SP.log 2>&1; grep -E "decode step|qkv_rope |attn |o_proj |gateup |down |lm_head  " gpurun_out/decode_trace_spre$SP.log | tail -7
done
