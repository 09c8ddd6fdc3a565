#!/bin/bash
# round-2 final validation: whole GPU suite, smoke, default bench, reference arm, launch list of one c2 request
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider -rf > gpurun_out/r02_pytest_final_full.log 2>&1
grep -a "passed\|failed\|FAILED\|Error\|error:" gpurun_out/r02_pytest_final_full.log | tail -12 > gpurun_out/r02_pytest_final.log; cat gpurun_out/r02_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_final_reference.json 2> gpurun_out/r02_bench_final_reference.err; echo "reference arm exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_one_request_8tok.csv python tools/one_request.py 8 > gpurun_out/ncu_one.log 2>&1; echo "ncu list exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02_bench_final.json').read().strip().splitlines()[-1])
print({k: d.get(k) for k in ('value', 'ms_per_step', 'e2e', 'gpu_launches', 'clocks')}); print(d['roofline']); print(d['prefill']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
r = json.loads(open('gpurun_out/r02_bench_final_reference.json').read().strip().splitlines()[-1])
print('reference', r.get('value'), r.get('cpu_baseline', {}).get('cores'))
PY
