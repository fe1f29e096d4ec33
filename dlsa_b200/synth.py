"""Synthetic data with the law of ``simulate_logistic`` (dlsa/models.py:6-40), generated on the device by a
counter-based generator keyed by (seed, global row, column): any sharding of the rows over GPUs sees the same
values.  ``host_uniform`` is the NumPy restatement of the device hash (bit-identical), used by the tests."""
from __future__ import annotations

import numpy as np
import torch

from ._lib import lib, check

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_PHI = np.uint64(0x9E3779B97F4A7C15)
_C2 = np.uint64(0xD1B54A32D192ED03)
_C3 = np.uint64(0x8CB92BA72F3D8DD7)
LABEL_COL = 0xFFFFFFFF


def _mix64(z):
    z = (z ^ (z >> np.uint64(30))) * _M1
    z = (z ^ (z >> np.uint64(27))) * _M2
    return z ^ (z >> np.uint64(31))


def host_uniform(seed: int, rows: np.ndarray, cols: np.ndarray) -> np.ndarray:
    """u01(seed, row, col) of csrc/misc.cu for broadcastable uint64 arrays rows, cols -> float64 in [0,1)."""
    with np.errstate(over="ignore"):
        rows = np.asarray(rows, dtype=np.uint64)
        cols = np.asarray(cols, dtype=np.uint64)
        z = _mix64(np.uint64(seed) + rows * _PHI)
        z = _mix64(z + cols * _C2 + _C3)
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


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
