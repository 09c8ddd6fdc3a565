#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/decode_trace.py > gpurun_out/decode_trace.log 2>&1; echo "trace exit $?"; cat gpurun_out/decode_trace.log | tail -40
