#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_pipeline.py -m gpu -q -p no:cacheprovider -k "attention or argmax or tall or batch or stages" 2>&1 | tail -8 > gpurun_out/r02_pytest7.log; tail -4 gpurun_out/r02_pytest7.log
timeout 400 python tools/ab.py --rounds 2 --cmd "python tools/microbench.py attn" A: B:SRGPT_ATTN_PP=-1 > gpurun_out/r02_ab_attn_pp3.txt 2>&1; tail -7 gpurun_out/r02_ab_attn_pp3.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_vit_pp -s 15 -c 1 -o gpurun_out/r02_attn_pp3 -f python tools/microbench.py attn > gpurun_out/ncu_attn_pp.log 2>&1; echo "ncu attn exit $?"
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench7.json 2> gpurun_out/r02_bench7.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench7.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step')}); print(d['prefill']['batch32'])
PY
