#!/bin/bash
mkdir -p gpurun_out
T="tests/test_gpu_configs.py -m gpu -q --timeout=600 -p no:cacheprovider -k c2_llama3"
echo "== default"; timeout 600 python -m pytest $T 2>&1 | tail -3
echo "== SRGPT_ATTN_NO_PREFETCH=1"; SRGPT_ATTN_NO_PREFETCH=1 timeout 600 python -m pytest $T 2>&1 | tail -3
echo "== SRGPT_NO_PDL=1"; SRGPT_NO_PDL=1 timeout 600 python -m pytest $T 2>&1 | tail -3
echo "== both"; SRGPT_NO_PDL=1 SRGPT_ATTN_NO_PREFETCH=1 timeout 600 python -m pytest $T 2>&1 | tail -3
