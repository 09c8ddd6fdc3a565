#!/bin/bash
# full GPU suite (beam search, rope scaling, whole-tile skinny GEMMs), then the bench (c3 batched decode) and an A/B of the skinny GEMM mode
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -s -rf > gpurun_out/r02_pytest13_full.log 2>&1
grep -a "fp16 \|^clip \|^beam \|passed\|failed\|FAILED\|Error\|error:" gpurun_out/r02_pytest13_full.log | tail -40 > gpurun_out/r02_pytest13.log; tail -30 gpurun_out/r02_pytest13.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench13.json 2> gpurun_out/r02_bench13.err; echo "bench exit $?"
SRGPT_GEMM_TSK_WHOLE=-1 timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench13_split.json 2> gpurun_out/r02_bench13_split.err; echo "bench (stream-K split) exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench13.json', 'gpurun_out/r02_bench13_split.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ('value', 'ms_per_step')}, d['prefill']['batch32'].get('ms_per_batch'), d['prefill']['batch32'].get('batched_decode'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
