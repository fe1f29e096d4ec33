#!/usr/bin/env python
"""Whole-fit time of the C2 shape as a function of the pilot block's size:  python tools/pilot_sweep.py [n] [p] [K]"""
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
for frac, pmin in ((1 / 32, 8), (1 / 64, 8), (1 / 128, 8), (1 / 256, 4), (1 / 32, 8)):
    cfg = dataclasses.replace(engine.EngineConfig(), pilot_fraction=frac, pilot_min=pmin)
    ts = []
    for rep in range(4):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fits = engine.fit_logistic_partitions(X, y, off, fit_intercept=False, config=cfg)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    if ref is None:
        ref = fits
    db = (fits.coef - ref.coef).abs().max().item()
    print(f"pilot {max(pmin, int(K * frac))} partitions: {min(ts[1:]):7.2f} ms, full passes {fits.passes}, pilot passes {fits.pilot_passes}, "
          f"iters {torch.bincount(fits.iters).tolist()}, max|coef - coef(1/32)| {db:.2e}, not ok {int(fits.n_bad)}")
