#!/usr/bin/env python
"""Interleaved A/B runs of a micro-benchmark command under different environment settings, inside ONE gpurun call.

Box-to-box variation of the same build was +-5..10 % in round 1 (and single outliers of 20 %), larger than most of the effects
being decided, so variants must be compared on the same box, alternating, several times:

    python tools/ab.py --rounds 3 --cmd "python tools/microbench.py attn" A:SRGPT_ATTN_S_AHEAD=1 B:SRGPT_ATTN_S_AHEAD=2

Every JSON line with "kernel" and "ms_median" printed by the command is collected; the table shows the median over rounds per
variant and the ratio to the first variant.
"""
import argparse
import json
import os
import shlex
import statistics
import subprocess
import sys
from collections import OrderedDict


def parse_variant(spec: str):
    """'NAME:K1=V1,K2=V2' -> (NAME, {K1: V1, K2: V2}); 'NAME:' is the unmodified environment."""
    name, _, rest = spec.partition(":")
    env = {}
    for kv in filter(None, rest.split(",")):
        k, _, v = kv.partition("=")
        env[k] = v
    return name, env


def collect(stdout: str):
    out = OrderedDict()
    for line in stdout.splitlines():
        line = line.strip()
        if not line.startswith("{"):
            continue
        try:
            d = json.loads(line)
        except ValueError:
            continue
        if "kernel" in d and "ms_median" in d:
            out[d["kernel"]] = float(d["ms_median"])
    return out


def summarise(samples):
    """samples: {variant: {kernel: [ms, ...]}} -> rows (kernel, {variant: median ms}, {variant: ratio to the first variant})."""
    variants = list(samples)
    kernels = list(OrderedDict((k, None) for v in variants for k in samples[v]))
    rows = []
    for k in kernels:
        med = {v: statistics.median(samples[v][k]) for v in variants if samples[v].get(k)}
        base = med.get(variants[0])
        rows.append((k, med, {v: (m / base if base else float("nan")) for v, m in med.items()}))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cmd", required=True)
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("variants", nargs="+", help="NAME:ENV=VAL[,ENV=VAL...]")
    a = ap.parse_args()
    variants = [parse_variant(v) for v in a.variants]
    samples = OrderedDict((n, OrderedDict()) for n, _ in variants)
    for r in range(a.rounds):
        for name, env in variants:
            p = subprocess.run(shlex.split(a.cmd), env=dict(os.environ, **env), capture_output=True, text=True)
            if p.returncode != 0:
                sys.stderr.write(p.stderr[-2000:])
                sys.exit(f"variant {name} failed in round {r}")
            for k, ms in collect(p.stdout).items():
                samples[name].setdefault(k, []).append(ms)
    names = [n for n, _ in variants]
    print(f"{'kernel':60s} " + " ".join(f"{n:>12s}" for n in names) + "   ratio to " + names[0])
    for k, med, ratio in summarise(samples):
        print(f"{k[:60]:60s} " + " ".join(f"{med.get(n, float('nan')):12.4f}" for n in names) + "   "
              + " ".join(f"{ratio.get(n, float('nan')):.3f}" for n in names[1:]))


if __name__ == "__main__":
    main()
