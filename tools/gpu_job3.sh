#!/bin/bash
# round-2 GPU job 3: tall stream-K GEMM, sampling, PDL pooling chain: tests, micro-benchmarks, mask_pool ncu, bench
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_full_depth.py 2>&1 | tail -40 > gpurun_out/r02_pytest3.log; tail -6 gpurun_out/r02_pytest3.log
timeout 600 python tools/ab.py --rounds 2 --cmd "python tools/microbench.py gemm" A: B:SRGPT_GEMM_TSK=-1 > gpurun_out/r02_ab_gemm_tsk.txt 2>&1; tail -26 gpurun_out/r02_ab_gemm_tsk.txt
timeout 300 python tools/ab.py --rounds 2 --cmd "python tools/microbench.py maskpool" A: B:SRGPT_NO_PDL=1 > gpurun_out/r02_ab_maskpool_pdl.txt 2>&1; tail -30 gpurun_out/r02_ab_maskpool_pdl.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mask_pool_kernel -s 2 -c 1 -o gpurun_out/r02_mask_pool -f python tools/microbench.py maskpool > gpurun_out/ncu_maskpool.log 2>&1; echo "ncu maskpool exit $?"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench3.json 2> gpurun_out/r02_bench3.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench3.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','with_stop_checks')}); print(d['prefill']); print(d['roofline'])
PY
tail -3 gpurun_out/r02_bench3.err
