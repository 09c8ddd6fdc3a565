#!/bin/bash
# round-2 GPU job 4: all GPU tests (batched decode, TP shards, TMA pooling, checkpoint loader), pooling micro-benchmark + ncu, bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_full_depth.py 2>&1 | tail -60 > gpurun_out/r02_pytest4.log; tail -12 gpurun_out/r02_pytest4.log
timeout 300 python tools/microbench.py maskpool > gpurun_out/r02_microbench_maskpool.jsonl 2>&1; cut -c1-170 gpurun_out/r02_microbench_maskpool.jsonl
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mask_pool_kernel -s 2 -c 1 -o gpurun_out/r02_mask_pool_tma -f python tools/microbench.py maskpool > gpurun_out/ncu_maskpool.log 2>&1; echo "ncu maskpool exit $?"
timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/r02_bench4.json 2> gpurun_out/r02_bench4.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench4.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','e2e','with_stop_checks')}); print(d['prefill']); print(d['roofline']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
PY
tail -3 gpurun_out/r02_bench4.err
