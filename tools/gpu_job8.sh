#!/bin/bash
# round-2 validation: the whole GPU suite (both element types), then the default bench and an fp16 bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider -s -rf > gpurun_out/r02_pytest8_full.log 2>&1
grep -a "fp16 \|passed\|failed\|FAILED\|Error\|error:" gpurun_out/r02_pytest8_full.log | tail -40 > gpurun_out/r02_pytest8.log; tail -30 gpurun_out/r02_pytest8.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench8.json 2> gpurun_out/r02_bench8.err; echo "bench exit $?"
timeout 600 python bench.py --steps 3 --warmup 3 --dtype fp16 > gpurun_out/r02_bench8_fp16.json 2> gpurun_out/r02_bench8_fp16.err; echo "bench fp16 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench8.json', 'gpurun_out/r02_bench8_fp16.json'):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, {k: d.get(k) for k in ('value', 'ms_per_step', 'dtype')}, d.get('roofline'), d['prefill'].get('batch32'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
