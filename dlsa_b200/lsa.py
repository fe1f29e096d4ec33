"""``dlsa.lsa`` drop-in: ``lars_lsa`` (dlsa/lsa.py:90-212) computed by the on-device LARS kernels."""
from __future__ import annotations

import numpy as np
import torch

from . import engine

EPS = float(np.finfo(np.float64).eps)


def _device():
    if not torch.cuda.is_available():
        raise RuntimeError("dlsa_b200 needs a CUDA device: there is no CPU implementation of the DLSA hot path")
    return torch.device("cuda", torch.cuda.current_device())


def _to_dev(a):
    if isinstance(a, torch.Tensor):
        return a.to(_device(), torch.float64)
    return torch.as_tensor(np.array(a, dtype=np.float64, copy=True), device=_device())


def lars_lsa(Sigma0, b0, intercept, n, type="lar", eps=EPS, max_steps=None):
    """Same signature, defaults and return keys as the reference (dlsa/lsa.py:90-91,211): a dict with
    ``AIC`` [k+1], ``BIC`` [k+1], ``beta`` [(k+1) x m] and ``beta0`` [k+1] as NumPy arrays.

    Accepts ndarray / np.matrix / DataFrame / tensor inputs (the reference only runs with ``np.matrix``,
    SURVEY.md App. A-2).  ``intercept=True`` follows the R original (dlsa/R/dlsa_alasso_func.R:115-125,
    243-245); the Python reference indexes by the sample size there and cannot run (App. A-3)."""
    S = _to_dev(np.asarray(Sigma0) if not isinstance(Sigma0, torch.Tensor) else Sigma0)
    b = _to_dev(np.asarray(b0).ravel() if not isinstance(b0, torch.Tensor) else b0).view(-1)
    res = engine.lars_path(S, b, bool(intercept), float(n), type=type, eps=float(eps), max_steps=max_steps)
    return {"AIC": res.AIC.cpu().numpy(), "BIC": res.BIC.cpu().numpy(), "beta": res.beta.cpu().numpy(),
            "beta0": res.beta0.cpu().numpy()}
