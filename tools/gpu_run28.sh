#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "gemm" --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -3 gpurun_out/pytest_gemm.log
run() { echo "== $1"; env $1 timeout 600 python tools/microbench.py gemm 2>&1 | python -c "
import json,sys
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    if '65536' in d['kernel'] or '8288x4096' in d['kernel']: print(f\"{d['kernel']:34s} {d['ms_median']:8.4f} ms {d['TFLOPs']:7.1f} TF {d['frac_tensor']:.3f}\")
"; }
run SRGPT_GEMM_NO_RESPF=1
run SRGPT_X=0
run SRGPT_GEMM_PAIR=1
run SRGPT_GEMM_NO_RESPF=1
run SRGPT_X=0
timeout 600 python tools/prefill_breakdown.py 32 4 > gpurun_out/prefill_breakdown_b32.log 2>&1; echo "breakdown exit $?"; tail -10 gpurun_out/prefill_breakdown_b32.log | head -9
