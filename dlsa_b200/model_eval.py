"""``dlsa.model_eval`` drop-in: ``logistic_model_eval_sdf`` (dlsa/model_eval.py:10-42) — second streaming pass that
sums the per-partition log-likelihood of several coefficient columns."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import engine, synth
from .models import _design


def logistic_model_eval_sdf(data_sdf, par, fit_intercept, Y_name, dummy_info=[], dummy_factors_baseline=[],
                            data_info=[]):
    """``data_sdf`` is a pandas frame with a ``partition_id`` column (or a Spark-like object with ``.toPandas()``).
    Returns a one-row frame with ``par.columns``: the global log-likelihood of each estimator (:38-42)."""
    pdf = data_sdf.toPandas() if hasattr(data_sdf, "toPandas") else data_sdf
    pdf = pdf.reset_index(drop=True)
    X, y, names = _design(pdf, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info, strict=False)
    pid_vals, pid_codes = np.unique(pdf["partition_id"].to_numpy(), return_inverse=True)
    Xg, yg, off = synth.group_by_partition(X, y, torch.as_tensor(pid_codes, device=X.device), len(pid_vals))
    P = torch.as_tensor(np.ascontiguousarray(np.asarray(par, dtype=np.float64)), device=X.device)
    ll = engine.logistic_loglik(Xg, yg, off, P, fit_intercept=fit_intercept).sum(0).cpu().numpy()
    return pd.DataFrame([ll], columns=list(par.columns))
