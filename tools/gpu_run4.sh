#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gpu_full.log 2>&1; echo "pytest gpu exit $?"; tail -4 gpurun_out/pytest_gpu_full.log
timeout 600 python tools/microbench.py maskpool > gpurun_out/microbench_maskpool.log 2>&1; echo "microbench exit $?"; cat gpurun_out/microbench_maskpool.log
# launch list of one eager request (8 tokens), profiler range only
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r01.csv python tools/one_request.py 8 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
# full capture of the dominant kernel (gate/up decode GEMV) and of the mask-pool kernel
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:decode_gemv_kernel -s 3 -c 3 -o gpurun_out/prof_gemv python tools/one_request.py 4 > gpurun_out/ncu_gemv.log 2>&1; echo "ncu gemv exit $?"
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:mask_pool_kernel -c 2 -o gpurun_out/prof_maskpool python tools/one_request.py 2 > gpurun_out/ncu_maskpool.log 2>&1; echo "ncu maskpool exit $?"
ls -la gpurun_out/*.ncu-rep
