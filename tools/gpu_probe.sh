#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/umma_probe/run_probe.py > gpurun_out/umma_probe.log 2>&1; echo "probe exit $?"; cat gpurun_out/umma_probe.log
