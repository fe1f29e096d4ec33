"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the device generators in dlsa_b200/csrc/misc.cu.

``host_uniform`` reproduces the counter-based hash ``u01(seed, row, col)`` bit for bit; ``simulate_logistic_host`` and
``airline_shaped_host`` rebuild the rows the device kernels (``dlsa_simulate_f64`` law 0, ``dlsa_simulate_airline_f64``)
write for the same (seed, global row).  The data law is the reference's ``simulate_logistic`` (dlsa/models.py:6-40; its
own RNG is unseeded, so only the law can be pinned) and the airline-shaped stand-in of SURVEY.md section 8d.
"""
from __future__ import annotations

import numpy as np

_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)
_PHI = np.uint64(0x9E3779B97F4A7C15)
_C2 = np.uint64(0xD1B54A32D192ED03)
_C3 = np.uint64(0x8CB92BA72F3D8DD7)
LABEL_COL = 0xFFFFFFFF
CAT_COL0 = 1000


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


def host_gauss(seed: int, rows: np.ndarray, cols: np.ndarray) -> np.ndarray:
    """gauss(seed, row, col) of csrc/misc.cu (Box-Muller on two counters); agrees with the device to a few ulp (the
    device's log / cospi are not bit-identical to NumPy's)."""
    cols = np.asarray(cols, dtype=np.uint64)
    u1 = host_uniform(seed, rows, np.uint64(2) * cols)
    u2 = host_uniform(seed, rows, np.uint64(2) * cols + np.uint64(1))
    return np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)


def simulate_logistic_host(n: int, p: int, seed: int, beta: np.ndarray, row0: int = 0):
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(row0))[:, None]
    X = host_uniform(seed, rows, np.arange(p, dtype=np.uint64)[None, :]) - 0.5
    u = host_uniform(seed, rows[:, 0], np.uint64(LABEL_COL))
    eta = X @ beta
    return X, (u < 1.0 / (1.0 + np.exp(-eta))).astype(np.float64), u, eta


def airline_shaped_host(n: int, seed: int, levels, cdfs, beta0, beta_num, beta_lev, row0: int = 0):
    """-> numeric [n, n_num], codes [n, n_cat] (level - 1), eta [n], u_label [n] (y = u_label < sigmoid(eta))."""
    rows = np.arange(n, dtype=np.uint64) + np.uint64(row0)
    n_num = len(beta_num)
    numeric = host_gauss(seed, rows[:, None], np.arange(n_num, dtype=np.uint64)[None, :])
    eta = beta0 + numeric @ np.asarray(beta_num)
    codes = np.empty((n, len(levels)), dtype=np.int32)
    lb = 0
    for c, (L, cdf) in enumerate(zip(levels, cdfs)):
        u = host_uniform(seed, rows, np.uint64(CAT_COL0 + c))
        lvl = np.minimum(np.searchsorted(np.asarray(cdf)[: L - 1], u, side="right"), L - 1)
        codes[:, c] = lvl - 1
        bl = np.concatenate([[0.0], np.asarray(beta_lev)[lb: lb + L - 1]])
        eta = eta + bl[lvl]
        lb += L - 1
    return numeric, codes, eta, host_uniform(seed, rows, np.uint64(LABEL_COL))
