#!/usr/bin/env python
"""Time the fused pass K1 alone (CUDA events on the launching stream) over the BASELINE.json shapes.

  python tools/k1_sweep.py [--cases c2,c1,c3,c4,c5] [--reps 5] [--out gpurun_out/k1_sweep.jsonl]

One JSON line per case: ms per pass, algorithmic GB/s and TFLOP/s (bytes = 8(px+1),
flops = p^2+5p per row) and their fractions of the measured HBM / FP64 peaks.  Not a bench line: bench.py is.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from dlsa_b200 import _lib, engine, synth  # noqa: E402

# name -> (rows, px, K, intercept, family, law)
CASES = {
    "c1": (100_000_000, 20, 1024, False, "logistic", 0),      # C1 shape at GPU scale (HBM-bound)
    "c2": (100_000_000, 50, 1024, False, "logistic", 0),      # the bench workload
    "c2_ncu": (20_000_000, 50, 296, False, "logistic", 0),    # the size the ncu capture uses
    "c2_k148": (100_000_000, 50, 148, False, "logistic", 0),
    "c2_k4096": (100_000_000, 50, 4096, False, "logistic", 0),
    "c2_icpt": (100_000_000, 50, 1024, True, "logistic", 0),
    "p56": (80_000_000, 56, 1024, False, "logistic", 0),
    "p32": (100_000_000, 32, 1024, False, "logistic", 0),
    "c3": (15_000_000, 181, 15, True, "logistic", 1),         # airline-shaped shard of one of 8 GPUs
    "c3_small": (1_500_000, 181, 15, True, "logistic", 1),
    "c3_lin": (15_000_000, 182, 15, False, "linear", 1),      # the same Gram without the link stage (ceiling for c3)
    "c5_lin": (2_560_000, 500, 1024, False, "linear", 1),
    "c4": (50_000_000, 100, 128, False, "linear", 1),         # OLS, p=100 (one GPU's 128 of 1024 partitions)
    "c4_logit": (50_000_000, 100, 128, False, "logistic", 0),
    "c4_logit_ncu": (10_000_000, 100, 148, False, "logistic", 0),   # the size the tiled-kernel ncu capture uses
    "c4_ncu": (10_000_000, 100, 148, False, "linear", 1),
    "c5": (2_560_000, 500, 1024, False, "logistic", 0),       # p=500, 2500-row partitions (1024 of 8192)
}


# compact categorical pass K1c on the airline-shaped design (7 numerics + 5 categorical columns, p = 182): name -> (rows, K)
CAT_CASES = {
    "c3_cat": (15_000_000, 15),        # one GPU's shard of C3 at G=8 — compare with the dense "c3" line
    "c3_cat_full": (120_000_000, 120), # the WHOLE C3 data set on one GPU (10 GB compact instead of 174 GB dense)
    "c3_cat_k1024": (60_000_000, 1024),
}


def run_cat(name, reps, hbm):
    from dlsa_b200 import synth as sy
    n, K = CAT_CASES[name]
    dev = torch.device("cuda", 0)
    numeric, codes, n_levels, y = sy.airline_shaped_device(n, 7, dev)
    off = sy.equal_partition_offsets(n, K, dev)
    p = 1 + numeric.shape[1] + sum(n_levels)
    beta = 0.05 * torch.randn(K, p, dtype=torch.float64, device=dev)
    max_rows = int((off[1:] - off[:-1]).max().item())
    g = torch.zeros(K, p, dtype=torch.float64, device=dev)
    H = torch.zeros(K, p, p, dtype=torch.float64, device=dev)
    import ctypes as C
    nl = (C.c_int * len(n_levels))(*n_levels)
    ws = engine._workspace(_lib.lib.dlsa_irls_pass_cat_workspace_bytes(K, len(n_levels), nl, max_rows), dev)

    def run():
        engine.irls_pass_cat(numeric, codes, n_levels, y, off, beta, True, None, None, None, max_rows,
                             want_loglik=False, out=(g, H, None), workspace=ws)

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    times = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = sorted(times)[len(times) // 2]
    bytes_row = 8 * numeric.shape[1] + 4 * codes.shape[1] + 8
    gbs = n * bytes_row / (ms * 1e-3) / 1e9
    return {"case": name, "kernel": "irls_cat_kernel (K1c)", "rows": n, "p": p, "K": K, "family": "logistic",
            "ms_per_pass": round(ms, 4), "ms_min": round(min(times), 4), "rows_per_s_per_pass": n / (ms * 1e-3),
            "algorithmic_bytes_per_row": bytes_row, "GBps": round(gbs, 1), "hbm_frac": round(gbs / hbm, 4),
            "dense_equivalent_bytes_per_row": 8 * p, "finite": bool(torch.isfinite(H).all().item())}


def peaks():
    hbm = 6550.0
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            hbm = float(json.load(f)["hbm_gbs"])
    except Exception:
        pass
    f64 = 37.1
    try:
        with open(os.path.join(ROOT, "profiles", "r01_microbench_fp64.json")) as f:
            d = json.load(f)
        f64 = max(v for k, v in d.items() if k.startswith("dmma884_"))
    except Exception:
        pass
    return hbm, f64


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default="c2,c2_ncu,c1,c3,c4,c5")
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    hbm, f64 = peaks()
    lines = []
    for name in args.cases.split(","):
        if name.startswith("describe"):  # the column-statistics pass (HBM-bound): describe_<px>
            px = int(name.split("_")[1]) if "_" in name else 50
            n = 4_000_000_000 // (8 * px)
            X = torch.empty(n, px, dtype=torch.float64, device=dev)
            synth.simulate_logistic_device(min(n, 5_000_000), px, 7, dev, out=(X[:min(n, 5_000_000)], torch.empty(min(n, 5_000_000), dtype=torch.float64, device=dev)))
            X[5_000_000:] = 0.25
            for _ in range(2):
                engine.column_stats(X)
            torch.cuda.synchronize()
            times = []
            for _ in range(args.reps):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                engine.column_stats(X)
                e1.record()
                torch.cuda.synchronize()
                times.append(e0.elapsed_time(e1))
            ms = sorted(times)[len(times) // 2]
            gbs = n * px * 8 / (ms * 1e-3) / 1e9
            line = {"case": name, "kernel": "colstats_partial_kernel (describe)", "rows": n, "px": px, "ms": round(ms, 4),
                    "GBps": round(gbs, 1), "hbm_frac": round(gbs / hbm, 4)}
            print(json.dumps(line), flush=True)
            lines.append(line)
            del X
            torch.cuda.empty_cache()
            continue
        if name in CAT_CASES:
            line = run_cat(name, args.reps, hbm)
            print(json.dumps(line), flush=True)
            lines.append(line)
            torch.cuda.empty_cache()
            continue
        n, px, K, icpt, fam, law = CASES[name]
        p = px + int(icpt)
        ldx = px + (px & 1)  # odd px is stored with an even row pitch (16-byte aligned rows -> TMA row copies)
        X = torch.zeros(n, ldx, dtype=torch.float64, device=dev)[:, :px]
        y = torch.empty(n, dtype=torch.float64, device=dev)
        chunk = 5_000_000
        beta_true = synth.true_beta(px) * (0.3 if px > 60 else 1.0)
        for s in range(0, n, chunk):
            m = min(chunk, n - s)
            synth.simulate_logistic_device(m, px, 7, dev, row0=s, beta=beta_true, law=law, out=(X[s:s + m], y[s:s + m]))
        off = synth.equal_partition_offsets(n, K, dev)
        beta = 0.05 * torch.randn(K, p, dtype=torch.float64, device=dev)
        family = _lib.FAMILY_LOGISTIC if fam == "logistic" else _lib.FAMILY_LINEAR
        max_rows = int((off[1:] - off[:-1]).max().item())
        g = torch.zeros(K, p, dtype=torch.float64, device=dev)
        H = torch.zeros(K, p, p, dtype=torch.float64, device=dev)
        ws = engine._workspace(_lib.lib.dlsa_irls_pass_workspace_bytes(K, p, max_rows), dev)

        def run():
            engine.irls_pass(X, y, off, beta if fam == "logistic" else None, icpt, family, None, max_rows,
                             want_loglik=False, out=(g, H, None), workspace=ws)

        for _ in range(2):
            run()
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            run()
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        gbs = n * 8 * (px + 1) / (ms * 1e-3) / 1e9
        tf = n * (p * p + 5 * p) / (ms * 1e-3) / 1e12
        line = {"case": name, "rows": n, "px": px, "p": p, "K": K, "family": fam, "ms_per_pass": round(ms, 4),
                "ms_min": round(min(times), 4), "rows_per_s_per_pass": n / (ms * 1e-3), "GBps": round(gbs, 1),
                "hbm_frac": round(gbs / hbm, 4), "TFLOPs": round(tf, 3), "fp64_frac": round(tf / f64, 4),
                "finite": bool(torch.isfinite(H).all().item())}
        print(json.dumps(line), flush=True)
        lines.append(line)
        del X, y, g, H, ws
        torch.cuda.empty_cache()
    if args.out:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        with open(args.out, "w") as f:
            for ln in lines:
                f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
