#!/usr/bin/env python
"""Whole DLSA fits (map -> combine -> LARS + BIC) on one GPU for the BASELINE.json configs other than the bench line:

  c1      n=1e5,  p=20,  K=16    configs[0] (the reference's own CPU-runnable case)
  c3_cat  n=1.2e8 airline-shaped, p=182 after one-hot, K=120 — the WHOLE configs[2] data set on one GPU through the
          compact categorical pass K1c (10 GB instead of 174 GB dense)
  c3      one GPU's dense shard of configs[2] at G=8: n=1.5e7, p=182, K=15
  c4      a 1e8-row chunk (K=128 of 1024 partitions... 1/10) of configs[3]: OLS, p=100 (closed form, one pass)
  c5      1024 of the 8192 partitions of configs[4]: p=500, 2500 rows each, 'lar' path + BIC

One JSON line per case: rows/s of the whole fit, passes, ms.  Diagnostic (profiles/), not the bench line.
  python tools/config_fits.py [--cases c1,c3_cat,c3,c4,c5] [--reps 3]"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from dlsa_b200 import engine, synth  # noqa: E402


def dense(n, px, law, dev, beta_scale=1.0):
    ldx = px + (px & 1)
    X = torch.zeros(n, ldx, dtype=torch.float64, device=dev)[:, :px]
    y = torch.empty(n, dtype=torch.float64, device=dev)
    bt = synth.true_beta(px) * beta_scale
    for s in range(0, n, 5_000_000):
        m = min(5_000_000, n - s)
        synth.simulate_logistic_device(m, px, 7, dev, row0=s, beta=bt, law=law, out=(X[s:s + m], y[s:s + m]))
    return X, y


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return out, sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c1,c3_cat,c3,c4,c5")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    for name in args.cases.split(","):
        if name == "c1":
            n, K = 100_000, 16
            X, y = dense(n, 20, 0, dev)
            off = synth.equal_partition_offsets(n, K, dev)
            (fits, comb, path), ms = timed(lambda: engine.dlsa_fit(X, y, off, sample_size=n), args.reps)
            p = 20
        elif name == "c3_cat":
            n, K = 120_000_000, 120
            num, codes, n_levels, y = synth.airline_shaped_device(n, 7, dev)
            off = synth.equal_partition_offsets(n, K, dev)
            p = 1 + num.shape[1] + sum(n_levels)

            def run():
                f = engine.fit_logistic_partitions_cat(num, codes, n_levels, y, off, fit_intercept=True)
                c = engine.combine_partitions(f)
                return f, c, engine.lars_path(c.Sig_inv_sum, c.beta_byOLS, True, float(n))
            (fits, comb, path), ms = timed(run, args.reps)
        elif name == "c3":
            n, K = 15_000_000, 15
            X, y = dense(n, 181, 0, dev, 0.3)  # law 0: U(-.5,.5) features, Bernoulli labels (law 1 has a linear response)
            off = synth.equal_partition_offsets(n, K, dev)
            (fits, comb, path), ms = timed(lambda: engine.dlsa_fit(X, y, off, fit_intercept=True, sample_size=n), args.reps)
            p = 182
        elif name == "c4":
            n, K = 100_000_000, 128
            X, y = dense(n, 100, 1, dev)
            off = synth.equal_partition_offsets(n, K, dev)
            (fits, comb, path), ms = timed(lambda: engine.dlsa_fit(X, y, off, sample_size=n, family="linear"), args.reps)
            p = 100
        elif name == "c5":
            n, K = 2_560_000, 1024
            X, y = dense(n, 500, 0, dev, 0.2)
            off = synth.equal_partition_offsets(n, K, dev)
            (fits, comb, path), ms = timed(lambda: engine.dlsa_fit(X, y, off, sample_size=n), args.reps)
            p = 500
        else:
            raise SystemExit(f"unknown case {name}")
        line = {"case": name, "rows": n, "p": p, "K": K, "ms_per_fit": round(ms, 3), "rows_per_s": n / (ms * 1e-3),
                "passes": fits.passes, "pilot_passes": fits.pilot_passes, "pilot_partitions": fits.pilot_partitions,
                "cold_restart": fits.cold_restart, "not_ok": int((fits.status != 0).sum().item()),
                "path_rows": int(path.beta.shape[0]), "support_BIC": int((path.beta[path.argmin_BIC] != 0).sum().item())}
        print(json.dumps(line), flush=True)
        del fits, comb, path
        if name != "c3_cat":
            del X
        else:
            del num, codes
        del y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
