#!/bin/bash
# the driver's N = 2 launch of both arms (scaling run), as a pre-flight
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err; echo "bench N=2 exit $?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29556 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_bench_n2_reference.json 2> gpurun_out/r02_bench_n2_reference.err; echo "reference N=2 exit $?"
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_n2.json', 'gpurun_out/r02_bench_n2_reference.json'):
    try:
        lines = [l for l in open(f).read().strip().splitlines() if l.startswith('{')]
        d = json.loads(lines[-1])
        print(f, len(lines), 'json line(s)', {k: d.get(k) for k in ('impl', 'value', 'n_gpus', 'ms_per_step', 'scaling')}, (d.get('cpu_baseline') or {}).get('cores'))
    except Exception as e:
        print(f, 'unreadable', e)
PY
grep -c "NCCL INFO" gpurun_out/r02_bench_n2.err; grep -m2 "nranks\|NVLS" gpurun_out/r02_bench_n2.err | cut -c1-200
