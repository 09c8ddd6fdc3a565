#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py -m gpu -q -k "attention or gemm" --timeout=600 -p no:cacheprovider > gpurun_out/pytest_attn.log 2>&1; echo "pytest attn/gemm exit $?"; tail -5 gpurun_out/pytest_attn.log
for pf in -1 8 16; do
  SRGPT_GEMM_L2PF=$pf timeout 600 python tools/microbench.py gemm > gpurun_out/microbench_gemm_pf$pf.jsonl 2>&1; echo "== L2PF=$pf"; python - <<PY
import json
for l in open('gpurun_out/microbench_gemm_pf$pf.jsonl'):
    try: d=json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print(f"{d['kernel']:34s} {d['ms_median']:8.4f} ms {d['TFLOPs']:7.1f} TF {d['frac_tensor']:.3f}")
PY
done
timeout 600 python tools/microbench.py attn > gpurun_out/microbench_attn.jsonl 2>&1; cut -c1-170 gpurun_out/microbench_attn.jsonl
timeout 600 python tools/prefill_breakdown.py 32 4 > gpurun_out/prefill_breakdown_b32.log 2>&1; echo "breakdown exit $?"; tail -10 gpurun_out/prefill_breakdown_b32.log | head -9
