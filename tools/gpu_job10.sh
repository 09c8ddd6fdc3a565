#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s -rf > gpurun_out/r02_pytest10_full.log 2>&1
grep -a "fp16 \|^clip \|passed\|failed\|FAILED\|Error\|error:" gpurun_out/r02_pytest10_full.log | tail -40 > gpurun_out/r02_pytest10.log; tail -30 gpurun_out/r02_pytest10.log
