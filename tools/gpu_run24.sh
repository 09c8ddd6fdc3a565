#!/bin/bash
mkdir -p gpurun_out
SRGPT_GEMM_PAIR=1 timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -q -x -k "gemm" --timeout=300 -p no:cacheprovider > gpurun_out/pytest_gemm_pair.log 2>&1; echo "pytest gemm (pair forced) exit $?"; tail -12 gpurun_out/pytest_gemm_pair.log | cut -c1-250
for pr in 1 -1; do
  SRGPT_GEMM_PAIR=$pr timeout 600 python tools/microbench.py gemm > gpurun_out/microbench_gemm_pair$pr.jsonl 2>&1; echo "== PAIR=$pr"; python - <<PY
import json
for l in open('gpurun_out/microbench_gemm_pair$pr.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:300]); continue
    print(f"{d['kernel']:34s} {d['ms_median']:8.4f} ms {d['TFLOPs']:7.1f} TF {d['frac_tensor']:.3f}")
PY
done
