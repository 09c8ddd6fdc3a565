#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider -s -rf -k "sampling or clip" > gpurun_out/r02_pytest11_full.log 2>&1
grep -a "^clip \|passed\|failed\|FAILED\|Error\|error:" gpurun_out/r02_pytest11_full.log | tail -20
timeout 900 python tools/ab.py --rounds 3 --cmd "python tools/decode_trace.py" A: B:SRGPT_GEMV_L2PF=2 C:SRGPT_GEMV_L2PF=3 > gpurun_out/r02_ab_decode_l2pf_oproj.txt 2>&1; tail -18 gpurun_out/r02_ab_decode_l2pf_oproj.txt
