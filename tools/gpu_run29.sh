#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench default exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench_default.json')); print({k:d[k] for k in ('value','ms_per_step','e2e','steps','warmup','gpu_launches')}); print(d['prefill']['batch32']); print(d['roofline']['frac'], d['roofline']['decode_step']); print(d['cpu_baseline']['value'], d['cpu_baseline']['cores'])"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; cut -c1-400 gpurun_out/bench_ref.json
