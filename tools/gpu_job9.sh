#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fp16.py tests/test_gpu_full_depth.py tests/test_gpu_tp.py tests/test_gpu_pipeline.py tests/test_gpu_checkpoint.py -m gpu -q -p no:cacheprovider -s -rf > gpurun_out/r02_pytest9_full.log 2>&1
grep -a "fp16 \|passed\|failed\|FAILED\|Error\|error:" gpurun_out/r02_pytest9_full.log | tail -40 > gpurun_out/r02_pytest9.log; tail -30 gpurun_out/r02_pytest9.log
