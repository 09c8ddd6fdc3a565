"""GPU-side diagnostic for the tcgen05 GEMM (not a test): structured inputs that expose descriptor /
swizzle / TMEM-layout mistakes, plus per-tile error maps.  Output is plain text for gpurun_out/."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from spatialrgpt_b200 import ops

dev = "cuda"
torch.manual_seed(0)


def report(tag, out, ref, bm=128, bn=128):
    out, ref = out.float().cpu(), ref.float().cpu()
    d = (out - ref).abs()
    rr = ref.pow(2).mean().sqrt().item() + 1e-12
    print(f"[{tag}] shape {tuple(out.shape)} max_err/rms {d.max().item()/rr:.3e} rms_err/rms {d.pow(2).mean().sqrt().item()/rr:.3e} finite={bool(torch.isfinite(out).all())}")
    if d.max().item() / rr > 5e-2:
        M, N = out.shape
        for mi in range(0, M, bm):
            row = []
            for ni in range(0, N, bn):
                blk = d[mi:mi + bm, ni:ni + bn]
                row.append(f"{blk.max().item()/rr:8.2e}")
            print("   tile-row", mi // bm, " ".join(row[:12]))
        bad = (d > 5e-2 * rr).nonzero()
        print("   first bad elements:", bad[:8].tolist())
        i, j = bad[0].tolist()
        print(f"   out[{i},{j}]={out[i,j]:.4f} ref={ref[i,j]:.4f}")
        return False
    return True


def onehot_probe(M, N, K):
    """A[i, i % K] = 1 -> C[i, j] must equal W[j, i % K]."""
    a = torch.zeros(M, K)
    a[torch.arange(M), torch.arange(M) % K] = 1
    w = torch.arange(N * K, dtype=torch.float32).reshape(N, K) % 251 - 125
    out = ops.gemm(a.bfloat16().to(dev), w.bfloat16().to(dev), out_fp32=True)
    ref = a @ w.t()
    ok = report(f"onehot {M}x{N}x{K}", out, ref)
    if not ok:
        o = out.cpu()
        for i in (0, 1, 8, 9, 33, 64, 127):
            if i < M:
                # which k does row i actually pick? compare with every column of W
                cand = [(k, float((o[i] - w[:, k]).abs().max())) for k in range(K)]
                best = min(cand, key=lambda t: t[1])
                print(f"   row {i}: expected k={i % K}, best matching k={best[0]} (err {best[1]:.2f}); out[:6]={o[i,:6].tolist()}")
    return ok


ok = True
for (M, N, K) in [(128, 128, 64), (128, 128, 128), (128, 256, 64), (256, 128, 192), (100, 72, 40)]:
    ok &= onehot_probe(M, N, K)
for (M, N, K) in [(128, 128, 64), (256, 384, 512), (259, 6144, 4096), (2048, 3456, 1152), (2048, 1152, 4304)]:
    a = torch.randn(M, K).bfloat16()
    w = (torch.randn(N, K) * K ** -0.5).bfloat16()
    out = ops.gemm(a.to(dev), w.to(dev), out_fp32=True)
    ok &= report(f"random {M}x{N}x{K}", out, a.float() @ w.float().t())
print("GEMM_DIAG", "PASS" if ok else "FAIL")
