#!/bin/bash
# round-2 GPU job: full GPU test suite, decode L2-prefetch A/B, c3 stage + kernel breakdown, GEMM ncu, cuBLAS context, bench
mkdir -p gpurun_out
export SRGPT_FULL_DEPTH_REPORT=gpurun_out/r02_oracle_c2_full.json
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/r02_pytest2.log; tail -4 gpurun_out/r02_pytest2.log
timeout 600 python tools/ab.py --rounds 2 --cmd "python tools/decode_trace.py" A: B:SRGPT_GEMV_L2PF=0 > gpurun_out/r02_ab_decode_l2pf.txt 2>&1; tail -16 gpurun_out/r02_ab_decode_l2pf.txt
timeout 300 python tools/prefill_breakdown.py 32 4 > gpurun_out/r02_prefill_breakdown_b32.txt 2>&1; tail -12 gpurun_out/r02_prefill_breakdown_b32.txt
timeout 300 python tools/microbench.py cublas > gpurun_out/r02_microbench_cublas.jsonl 2>&1; cut -c1-150 gpurun_out/r02_microbench_cublas.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:gemm|attn|norm|rope|splice|mask|pool|patchify|argmax' -s 950 -c 950 --csv --log-file gpurun_out/r02_launches_c3.csv python tools/prefill_breakdown.py 32 4 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_pair -s 3 -c 1 -o gpurun_out/r02_gemm_pair_res -f python tools/gemm_one.py 65536 1152 1152 4 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench2.json 2> gpurun_out/r02_bench2.err; echo "bench exit $?"; cut -c1-400 gpurun_out/r02_bench2.json
