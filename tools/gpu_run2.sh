#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -2 gpurun_out/smoke.log
# launch list (shares, not absolutes): one short request through generate()
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_r01.csv python tools/one_request.py > gpurun_out/ncu_launches.log 2>&1; echo "ncu launches exit $?"
