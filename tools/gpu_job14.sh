#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/ab.py --rounds 2 --cmd "python tools/batched_decode_time.py 32" A: B:SRGPT_GEMM_TSK_WHOLE=-1 C:SRGPT_GEMM_TSK_WHOLE=-1,SRGPT_GEMM_TSK_ABOX=128 D:SRGPT_GEMM_TSK_ABOX=128 > gpurun_out/r02_ab_batched_decode_skinny.txt 2>&1; tail -8 gpurun_out/r02_ab_batched_decode_skinny.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_batched_decode_whole.csv python tools/batch_decode_once.py 32 2 > gpurun_out/ncu_bd2.log 2>&1; echo "ncu list exit $?"
