#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider -rf -k "sampling or rope" 2>&1 | tail -3
# batched decode (32 sequences): one ncu --set full capture of the stream-K GEMMs of a decode step (o_proj / gate-up shapes at M = 32)
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:gemm_tall_sk -s 8 -c 4 -o gpurun_out/r02_tsk_m32 -f python tools/batch_decode_once.py 32 2 > gpurun_out/ncu_tsk.log 2>&1; echo "ncu tsk exit $?"
tail -3 gpurun_out/ncu_tsk.log
