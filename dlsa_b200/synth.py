"""Synthetic data with the law of ``simulate_logistic`` (dlsa/models.py:6-40), generated on the device by a
counter-based generator keyed by (seed, global row, column): any sharding of the rows over GPUs sees the same
values.  The NumPy restatement of the device hash (bit-identical) is test infrastructure: ``oracle/synth_host.py``."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lib, check

LABEL_COL = 0xFFFFFFFF   # counter-space column of the label's uniform draw (csrc/misc.cu)
CAT_COL0 = 1000          # counter-space column of categorical column 0 in the airline-shaped stream


def true_beta(p: int) -> np.ndarray:
    """beta = 1 on the first int(0.4 p) coordinates, else 0 (dlsa/models.py:12,18-19)."""
    b = np.zeros(p)
    b[: int(p * 0.4)] = 1.0
    return b


def simulate_logistic_device(n: int, p: int, seed: int, device, row0: int = 0, beta=None, law: int = 0,
                             out=None):
    """Rows [row0, row0+n) of the global synthetic stream: X [n,p] ~ U(-.5,.5) (law 0) or N(0,1) (law 1),
    y ~ Bernoulli(sigmoid(X beta)) or X beta + N(0,1).  Returns (X, y) on ``device``."""
    if beta is None:
        beta = true_beta(p)
    bt = torch.as_tensor(np.asarray(beta, dtype=np.float64), device=device)
    if out is None:
        X = torch.empty(n, p, dtype=torch.float64, device=device)
        y = torch.empty(n, dtype=torch.float64, device=device)
    else:
        X, y = out
    check(lib.dlsa_simulate_f64(X.data_ptr(), X.stride(0), y.data_ptr(), int(row0), int(n), int(p), bt.data_ptr(),
                                int(seed), int(law), torch.cuda.current_stream().cuda_stream), "dlsa_simulate_f64")
    return X, y


def equal_partition_offsets(n: int, K: int, device=None) -> torch.Tensor:
    """Offsets of K contiguous partitions whose sizes are those of the systematic rule ``i % K``
    (dlsa/models.py:32-33; sizes differ by at most one row)."""
    base, rem = divmod(n, K)
    sizes = torch.full((K,), base, dtype=torch.int64)
    sizes[:rem] += 1
    off = torch.zeros(K + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(sizes, 0)
    return off.to(device) if device is not None else off


def group_by_partition(X: torch.Tensor, y: torch.Tensor, partition_id: torch.Tensor, K: int):
    """Make every partition's rows contiguous (stable order inside a partition): the device-side analogue of
    Spark's ``repartition(K, "partition_id")`` (projects/logistic_dlsa.py:295).  Returns (X, y, part_offsets)."""
    pid = partition_id.to(torch.int64)
    order = torch.argsort(pid, stable=True)
    counts = torch.bincount(pid, minlength=K)
    off = torch.zeros(K + 1, dtype=torch.int64, device=X.device)
    off[1:] = torch.cumsum(counts, 0)
    return X[order].contiguous(), y[order].contiguous(), off


# ---- airline-shaped synthetic design (BASELINE.json configs[2]; SURVEY.md §8d "C3") ---------------------------------
# 7 standardised numerics ~N(0,1) (Year, DayofMonth, DepTime, CRSDepTime, CRSArrTime, ActualElapsedTime, Distance) and
# the categoricals Month(12), DayOfWeek(7), UniqueCarrier(8 incl. OTHERS), Origin(76), Dest(76); the first level of each
# is the baseline (projects/logistic_dlsa.py:157-165) => 11+6+7+75+75 = 174 dummy columns, p = 182 with the intercept.
AIRLINE_LEVELS = (12, 7, 8, 76, 76)
AIRLINE_ZIPF = (False, False, True, True, True)
AIRLINE_NUM = 7


def airline_truth(seed: int = 11):
    """(beta0, beta_num [7], beta_levels [174]): N(0, 0.3^2) with 60 % zeros, intercept -1."""
    rng = np.random.default_rng(seed)
    p = AIRLINE_NUM + sum(L - 1 for L in AIRLINE_LEVELS)
    b = rng.standard_normal(p) * 0.3
    b[rng.random(p) < 0.6] = 0.0
    return -1.0, b[:AIRLINE_NUM].copy(), b[AIRLINE_NUM:].copy()


def airline_cdfs():
    """Cumulative level shares of the five categorical columns (uniform or Zipf(1)), level 0 = baseline."""
    out = []
    for L, zipf in zip(AIRLINE_LEVELS, AIRLINE_ZIPF):
        w = 1.0 / np.arange(1, L + 1) if zipf else np.ones(L)
        out.append(np.cumsum(w / w.sum()))
    return out


def airline_shaped_device(n: int, seed: int, device, row0: int = 0, out=None):
    """Rows [row0, row0+n) of the compact airline-shaped stream on ``device``: numeric [n,7] float64, codes [n,5] int32
    (level-in-block index, -1 = baseline level), n_levels (11,6,7,75,75), y [n] ~ Bernoulli(sigmoid(eta)).  Counter-based
    (``dlsa_simulate_airline_f64``): a row depends on (seed, global row index) only, so every sharding sees the same
    data.  Benchmark / test input, not part of the hot path."""
    b0, bnum, blev = airline_truth()
    n_levels = [L - 1 for L in AIRLINE_LEVELS]
    if out is None:
        numeric = torch.empty(n, AIRLINE_NUM, dtype=torch.float64, device=device)
        codes = torch.empty(n, len(n_levels), dtype=torch.int32, device=device)
        y = torch.empty(n, dtype=torch.float64, device=device)
    else:
        numeric, codes, y = out
    lev = torch.as_tensor(np.asarray(AIRLINE_LEVELS, dtype=np.int32), device=device)
    cdf = torch.as_tensor(np.concatenate(airline_cdfs()), device=device)
    bn = torch.as_tensor(bnum, device=device)
    bl = torch.as_tensor(blev, device=device)
    check(lib.dlsa_simulate_airline_f64(numeric.data_ptr(), numeric.stride(0), codes.data_ptr(), y.data_ptr(), int(row0),
                                        int(n), AIRLINE_NUM, len(n_levels), lev.data_ptr(), cdf.data_ptr(), float(b0),
                                        bn.data_ptr(), bl.data_ptr(), int(seed),
                                        torch.cuda.current_stream().cuda_stream), "dlsa_simulate_airline_f64")
    return numeric, codes, n_levels, y
