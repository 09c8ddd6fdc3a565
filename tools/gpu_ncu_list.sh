#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm|attn|norm|rope|splice|mask|pool|patchify|argmax' -s 950 -c 950 --csv --log-file gpurun_out/launches_c3.csv python tools/prefill_breakdown.py 32 4 > gpurun_out/ncu_list.log 2>&1; echo "ncu exit $?"; wc -l gpurun_out/launches_c3.csv
