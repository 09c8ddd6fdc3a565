#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_fwd_tc -s 15 -c 1 -o gpurun_out/attn_tc -f python tools/microbench.py attn > gpurun_out/ncu_attn.log 2>&1; echo "ncu exit $?"; tail -5 gpurun_out/ncu_attn.log; ls -la gpurun_out/attn_tc.ncu-rep
