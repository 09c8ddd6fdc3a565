#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "gemm" --timeout=600 -p no:cacheprovider > gpurun_out/pytest_gemm.log 2>&1; echo "pytest gemm exit $?"; tail -3 gpurun_out/pytest_gemm.log
for ew in 8 0; do
  SRGPT_GEMM_EW=$ew timeout 600 python tools/microbench.py gemm > gpurun_out/microbench_gemm_ew$ew.jsonl 2>&1; echo "== SRGPT_GEMM_EW=$ew"; python - <<PY
import json
for l in open('gpurun_out/microbench_gemm_ew$ew.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(f"{d['kernel']:34s} {d['ms_median']:8.4f} ms {d['TFLOPs']:7.1f} TF {d['frac_tensor']:.3f}")
PY
done
timeout 600 python tools/prefill_breakdown.py 32 4 > gpurun_out/prefill_breakdown_b32.log 2>&1; echo "breakdown exit $?"; tail -10 gpurun_out/prefill_breakdown_b32.log | head -9
