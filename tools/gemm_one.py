"""One GEMM shape repeated a few times (the command profiled by `ncu --set full -k regex:gemm -s 3 -c 1`).
usage: python tools/gemm_one.py M N K epilogue"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spatialrgpt_b200 import ops

M, N, K, epi = (int(x) for x in sys.argv[1:5])
dev, BF = "cuda", torch.bfloat16
rnd = lambda *s: (torch.randn(*s, device=dev) * 0.05).to(BF)  # noqa: E731
a, w = rnd(M, K), rnd(N, K)
bias = rnd(N) if epi in (1, 2, 3, 4) else None
n_out = N // 2 if epi == 5 else N
res = rnd(M, n_out) if epi == 4 else None
out = torch.empty(M, n_out, dtype=BF, device=dev)
for _ in range(6):
    ops.gemm(a, w, bias=bias, residual=res, epilogue=epi, out=out)
torch.cuda.synchronize()
print("ok", float(out.float().abs().mean()))
