#!/usr/bin/env python
"""How many fused passes the Newton loop needs as a function of the step tolerance, and what the looser stop costs in
accuracy:  python tools/newton_tol.py [n] [p] [K]   (defaults: the C2 shape)"""
import dataclasses
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dlsa_b200 import engine, synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000_000
p = int(sys.argv[2]) if len(sys.argv) > 2 else 50
K = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
dev = torch.device("cuda", 0)
X, y = synth.simulate_logistic_device(n, p, 1, dev)
off = synth.equal_partition_offsets(n, K, dev)
ref = None
for tol in (1e-12, 1e-10, 1e-9, 1e-8, 1e-7, 1e-6):
    cfg = dataclasses.replace(engine.EngineConfig(), tol=tol)
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fits = engine.fit_logistic_partitions(X, y, off, fit_intercept=False, config=cfg)
        e1.record()
        torch.cuda.synchronize()
    if ref is None:
        ref = fits
    db = (fits.coef - ref.coef).abs().max().item()
    dh = ((fits.Sig_inv - ref.Sig_inv).flatten(1).norm(dim=1) / ref.Sig_inv.flatten(1).norm(dim=1)).max().item()
    print(f"tol {tol:g}: {e0.elapsed_time(e1):7.2f} ms, passes {fits.passes}, iters {torch.bincount(fits.iters).tolist()}, "
          f"max|coef - coef(1e-12)| {db:.2e}, max relfro(Sig_inv) {dh:.2e}, not ok {int(fits.n_bad)}")
