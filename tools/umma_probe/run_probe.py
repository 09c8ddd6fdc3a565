#!/usr/bin/env python
"""Checks UMMA shared-memory descriptor hypotheses on the GPU (see probe.cu).  Prints one line per hypothesis."""
import ctypes as C
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "libprobe.so"))


class Load(C.Structure):
    _fields_ = [("map", C.c_int), ("ndim", C.c_int), ("c0", C.c_int), ("c1", C.c_int), ("c2", C.c_int), ("smem_off", C.c_uint32), ("bytes", C.c_uint32)]


class Mma(C.Structure):
    _fields_ = [("a_desc", C.c_uint64), ("b_desc", C.c_uint64), ("a_off", C.c_uint32), ("b_off", C.c_uint32), ("d_col", C.c_uint32),
                ("idesc", C.c_uint32), ("acc", C.c_uint32)]


class Plan(C.Structure):
    _fields_ = [("n_loads", C.c_int), ("n_mma", C.c_int), ("n_cols", C.c_int), ("manual_a", C.c_int), ("a_manual_off", C.c_uint32),
                ("a_src", C.c_void_p), ("a_ld", C.c_int), ("loads", Load * 8), ("mma", Mma * 40)]


assert lib.probe_plan_size() == C.sizeof(Plan), (lib.probe_plan_size(), C.sizeof(Plan))
DEV = "cuda"
torch.manual_seed(0)
A = (torch.randn(128, 256) * 0.5).to(torch.bfloat16).to(DEV)
V = (torch.randn(64, 256) * 0.5).to(torch.bfloat16).to(DEV)   # [kv rows (K), channels]
Bk = (torch.randn(64, 256) * 0.5).to(torch.bfloat16).to(DEV)  # [N rows, K cols]


def make_map(t, dims, strides_bytes, box, swizzle):
    buf = (C.c_uint8 * 128)()
    nd = len(dims)
    r = lib.probe_make_map(buf, C.c_void_p(t.data_ptr()), nd, (C.c_longlong * nd)(*dims), (C.c_longlong * max(nd - 1, 1))(*strides_bytes),
                           (C.c_int * nd)(*box), swizzle)
    if r != 0:
        raise RuntimeError(f"cuTensorMapEncodeTiled -> {r} for dims {dims} box {box} sw {swizzle}")
    return bytes(buf)


def desc(layout_type, lbo16, sbo16):
    return (lbo16 << 16) | (sbo16 << 32) | (1 << 46) | (layout_type << 61)


def idesc(n, a_mn=0, b_mn=0, m=128):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)


def run(name, maps, loads, mmas, n_cols, ref, manual=None):
    plan = Plan()
    plan.n_loads, plan.n_mma, plan.n_cols = len(loads), len(mmas), n_cols
    for i, l in enumerate(loads):
        plan.loads[i] = Load(*l)
    for i, m in enumerate(mmas):
        plan.mma[i] = Mma(*m)
    if manual is not None:
        plan.manual_a, plan.a_manual_off, plan.a_src, plan.a_ld = 1, manual[0], manual[1].data_ptr(), manual[2]
    pd = torch.frombuffer(bytearray(bytes(plan)), dtype=torch.uint8).to(DEV)
    mp = b"".join(maps + [bytes(128)] * (4 - len(maps)))
    D = torch.full((128, n_cols), float("nan"), device=DEV)
    rc = lib.probe_launch(mp, C.c_void_p(pd.data_ptr()), C.c_void_p(D.data_ptr()))
    if rc != 0:
        print(f"{name:58s} LAUNCH FAILED rc={rc}")
        return False
    err = (D - ref.float()).abs().max().item()
    scale = ref.float().abs().max().item()
    ok = err <= 2e-2 * max(scale, 1.0)
    print(f"{name:58s} {'OK  ' if ok else 'FAIL'} max|err| {err:.4f} (max|ref| {scale:.2f})", flush=True)
    return ok


SW128, SW32 = 3, 1
K128, K32 = desc(2, 1, 64), desc(6, 1, 16)   # K-major: SBO = 8 rows * 128 B / 8 rows * 32 B
mapA = make_map(A, [256, 128], [512], [64, 128], SW128)
mapB = make_map(Bk, [256, 64], [512], [64, 64], SW128)
Af, Bf, Vf = A.float(), Bk.float(), V.float()

# H1: sanity, both operands K-major SW128
mm = [(K128, K128, k * 32, 32768 + k * 32, 0, idesc(64), 1 if k else 0) for k in range(4)]
run("H1 K-major SW128 x K-major SW128 (known good)", [mapA, mapB], [(0, 2, 0, 0, 0, 0, 16384), (1, 2, 0, 0, 0, 32768, 8192)], mm, 64, Af[:, :64] @ Bf[:, :64].T)
# H7: A written by threads with the SW128 formula (P tile of the attention kernel)
run("H7 manual SW128 A (generic-proxy stores + fence.proxy.async)", [mapA, mapB], [(1, 2, 0, 0, 0, 32768, 8192)], mm, 64, Af[:, :64] @ Bf[:, :64].T,
    manual=(0, A, 256))

# H2: B = V tile [64 kv rows, hd cols], MN-major, SW128 boxes of 64 channels
c0 = 16 * 0 + 72 - 8  # a 16-byte aligned, non-zero channel offset (64)
mapV = make_map(V, [256, 64], [512], [64, 64], SW128)
ldV = [(0, 2, 0, 0, 0, 0, 16384), (1, 2, c0, 0, 0, 32768, 8192), (1, 2, c0 + 64, 0, 0, 40960, 8192)]
for lbo in (1, 512):
    mm = [(K128, desc(2, lbo, 64), k * 32, 32768 + k * 2048, 0, idesc(64, b_mn=1), 1 if k else 0) for k in range(4)]
    run(f"H2a MN-major SW128 B, N=64, SBO=1024 LBO16={lbo} adv=2048", [mapA, mapV], ldV, mm, 64, Af[:, :64] @ Vf[:, c0:c0 + 64])
for lbo, sbo, adv in ((512, 64, 2048), (64, 512, 2048), (512, 64, 256), (64, 512, 256)):
    mm = [(K128, desc(2, lbo, sbo), k * 32, 32768 + k * adv, 0, idesc(80, b_mn=1), 1 if k else 0) for k in range(4)]
    run(f"H2b MN-major SW128 B, N=80, LBO16={lbo} SBO16={sbo} adv={adv}", [mapA, mapV], ldV, mm, 80, Af[:, :64] @ Vf[:, c0:c0 + 80])

# H3: K-major 32B-swizzled tail boxes (16 channels) for both operands
mapA32 = make_map(A, [256, 128], [512], [16, 128], SW32)
mapB32 = make_map(Bk, [256, 64], [512], [16, 64], SW32)
run("H3 K-major SW32 tails (K=16)", [mapA32, mapB32], [(0, 2, 64, 0, 0, 16384, 4096), (1, 2, 64, 0, 0, 49152, 2048)],
    [(K32, K32, 16384, 49152, 0, idesc(64), 0)], 64, Af[:, 64:80] @ Bf[:, 64:80].T)
mm = [(K128, K128, k * 32, 32768 + k * 32, 0, idesc(64), 1 if k else 0) for k in range(4)] + [(K32, K32, 16384, 49152, 0, idesc(64), 1)]
run("H3+ K=80 = SW128 box (64) + SW32 tail (16)", [mapA, mapB, mapA32, mapB32],
    [(0, 2, 0, 0, 0, 0, 16384), (1, 2, 0, 0, 0, 32768, 8192), (2, 2, 64, 0, 0, 16384, 4096), (3, 2, 64, 0, 0, 49152, 2048)], mm, 64,
    Af[:, :80] @ Bf[:, :80].T)

# H4: 3-D maps (72-wide heads) with out-of-bounds zero fill of channels 72..79
h = 1
m3A = make_map(A, [72, 3, 128], [144, 512], [64, 1, 128], SW128)
m3B = make_map(Bk, [72, 3, 64], [144, 512], [64, 1, 64], SW128)
m3A32 = make_map(A, [72, 3, 128], [144, 512], [16, 1, 128], SW32)
m3B32 = make_map(Bk, [72, 3, 64], [144, 512], [16, 1, 64], SW32)
run("H4 3-D maps, head dim 72 -> K=80 with OOB zero fill", [m3A, m3B, m3A32, m3B32],
    [(0, 3, 0, h, 0, 0, 16384), (1, 3, 0, h, 0, 32768, 8192), (2, 3, 64, h, 0, 16384, 4096), (3, 3, 64, h, 0, 49152, 2048)], mm, 64,
    Af[:, 72:144] @ Bf[:, 72:144].T)

# H5: MN-major SW32 tail of V (16 channels): second MMA per k-step writes D columns 64..79
mapV32 = make_map(V, [256, 64], [512], [16, 64], SW32)
ld5 = [(0, 2, 0, 0, 0, 0, 16384), (1, 2, c0, 0, 0, 32768, 8192), (2, 2, c0 + 64, 0, 0, 57344, 2048)]
for lbo, sbo, adv in ((1, 16, 512), (16, 1, 512), (1, 16, 64), (16, 16, 512)):
    mm5 = []
    for k in range(4):
        mm5.append((K128, desc(2, 1, 64), k * 32, 32768 + k * 2048, 0, idesc(64, b_mn=1), 1 if k else 0))
        mm5.append((K128, desc(6, lbo, sbo), k * 32, 57344 + k * adv, 64, idesc(16, b_mn=1), 1 if k else 0))
    run(f"H5 MN-major SW128 (N=64) + SW32 tail (N=16) LBO16={lbo} SBO16={sbo} adv={adv}", [mapA, mapV, mapV32], ld5, mm5, 80, Af[:, :64] @ Vf[:, c0:c0 + 80])
# H6: V tail via a 3-D map (head dim 72, OOB zero fill) in MN-major SW32
m3V = make_map(V, [72, 3, 64], [144, 512], [64, 1, 64], SW128)
m3V32 = make_map(V, [72, 3, 64], [144, 512], [16, 1, 64], SW32)
ref6 = torch.zeros(128, 80, device=DEV)
ref6[:, :72] = Af[:, :64] @ Vf[:, 72:144]
mm6 = []
for k in range(4):
    mm6.append((K128, desc(2, 1, 64), k * 32, 32768 + k * 2048, 0, idesc(64, b_mn=1), 1 if k else 0))
    mm6.append((K128, desc(6, 1, 16), k * 32, 57344 + k * 512, 64, idesc(16, b_mn=1), 1 if k else 0))
run("H6 V head (72 ch) via 3-D maps: SW128 box + SW32 tail, MN-major", [mapA, m3V, m3V32],
    [(0, 2, 0, 0, 0, 0, 16384), (1, 3, 0, 1, 0, 32768, 8192), (2, 3, 64, 1, 0, 57344, 2048)], mm6, 80, ref6)
