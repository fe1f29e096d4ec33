#!/usr/bin/env python
"""Time (and offer to ncu) the batched SPD solve alone:  python tools/chol_only.py [p] [K] [reps]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dlsa_b200 import engine  # noqa: E402
from dlsa_b200._lib import check, lib  # noqa: E402

p = int(sys.argv[1]) if len(sys.argv) > 1 else 500
K = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
g = torch.Generator(device="cuda").manual_seed(2)
M = torch.randn(K, p + 8, p, dtype=torch.float64, device="cuda", generator=g)
A = M.transpose(1, 2) @ M + 0.1 * torch.eye(p, dtype=torch.float64, device="cuda")
b = torch.randn(K, p, dtype=torch.float64, device="cuda", generator=g)
x = torch.empty_like(b)
st = torch.zeros(K, dtype=torch.int32, device="cuda")
ws = engine._workspace(lib.dlsa_newton_update_workspace_bytes(K, p), A.device)
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    check(lib.dlsa_spd_solve_f64(A.data_ptr(), b.data_ptr(), K, p, x.data_ptr(), st.data_ptr(), ws.data_ptr(), ws.numel(),
                                 torch.cuda.current_stream().cuda_stream), "dlsa_spd_solve_f64")
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    flops = K * (p ** 3 / 3.0 + 2.0 * p * p)
    print(f"p={p} K={K}: {ms:.3f} ms, {flops / ms / 1e9:.2f} TFLOP/s, max residual "
          f"{(torch.einsum('kij,kj->ki', A, x) - b).abs().max().item():.2e}, not ok {int((st != 0).sum())}", flush=True)
