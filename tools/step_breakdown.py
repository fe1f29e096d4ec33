#!/usr/bin/env python
"""Where one bench step (C2: n=1e8, p=50, K=1024) spends its time: CUDA-event and host wall-clock time of the pilot fit,
the main fit, the combine and the LARS path.  Diagnostic; bench.py is the benchmark."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dlsa_b200 import engine, synth  # noqa: E402

n, p, K = 100_000_000, 50, 1024
if len(sys.argv) == 4:  # python tools/step_breakdown.py <rows> <p> <partitions>
    n, p, K = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dev = torch.device("cuda", 0)
X = torch.empty(n, p, dtype=torch.float64, device=dev)
y = torch.empty(n, dtype=torch.float64, device=dev)
bt = synth.true_beta(p) * (0.2 if p > 60 else 1.0)
for s in range(0, n, 10_000_000):
    m = min(10_000_000, n - s)
    synth.simulate_logistic_device(m, p, 20210517, dev, row0=s, beta=bt, out=(X[s:s + m], y[s:s + m]))
off = synth.equal_partition_offsets(n, K, dev)


def timed(fn):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    out = fn()
    e1.record()
    t_host = time.perf_counter() - t0
    torch.cuda.synchronize()
    return out, e0.elapsed_time(e1), t_host * 1e3


for rep in range(3):
    engine.PASS_EVENTS = []
    fits, t_fit, h_fit = timed(lambda: engine.fit_logistic_partitions(X, y, off))
    ev = engine.PASS_EVENTS
    engine.PASS_EVENTS = None
    comb, t_comb, h_comb = timed(lambda: engine.combine_partitions(fits))
    path, t_lars, h_lars = timed(lambda: engine.lars_path(comb.Sig_inv_sum, comb.beta_byOLS, False, float(n)))
    (_, _, _), t_all, h_all = timed(lambda: engine.dlsa_fit(X, y, off, sample_size=n))
    passes = [round(a.elapsed_time(b), 3) for a, b, _ in ev]
    print(f"rep {rep}: fit {t_fit:.2f} ms (host {h_fit:.2f}) passes {passes} sum {sum(passes):.2f} | combine {t_comb:.2f} "
          f"(host {h_comb:.2f}) | lars {t_lars:.2f} (host {h_lars:.2f}) | whole step {t_all:.2f} (host {h_all:.2f})", flush=True)
