#!/bin/bash
for a in 1 2 1 2; do echo "== SRGPT_ATTN_S_AHEAD=$a"; SRGPT_ATTN_S_AHEAD=$a timeout 300 python tools/microbench.py attn 2>&1 | cut -c1-150; done
