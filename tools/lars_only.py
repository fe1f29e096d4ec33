#!/usr/bin/env python
"""Time (and offer to ncu) the LARS path kernel alone on a synthetic SPD problem:  python tools/lars_only.py [p] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dlsa_b200 import engine  # noqa: E402

p = int(sys.argv[1]) if len(sys.argv) > 1 else 500
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda").manual_seed(1)
A = torch.randn(4 * p, p, dtype=torch.float64, device="cuda", generator=g)
Sigma = A.T @ A / (4 * p) * 1e4
b = torch.randn(p, dtype=torch.float64, device="cuda", generator=g)
b[: p // 20] += 3.0
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    path = engine.lars_path(Sigma, b, False, 1e6)
    e1.record()
    torch.cuda.synchronize()
    print(f"p={p}: {e0.elapsed_time(e1):.2f} ms, {path.beta.shape[0]} knots", flush=True)
