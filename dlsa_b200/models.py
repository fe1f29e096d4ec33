"""``dlsa.models`` drop-in: ``simulate_logistic``, ``logistic_model``, ``logistic_model_eval`` (+ ``linear_model``).

Frames in, frames out, with the reference's signatures, column names and row order (dlsa/models.py); every number
is computed by the CUDA kernels behind include/dlsa_b200.h.
"""
from __future__ import annotations

import warnings

import numpy as np
import pandas as pd
import torch

from . import engine, synth
from .dummies import ChunkMissingDummies, design_columns, encode_design, encode_design_compact
from .lsa import _device


def simulate_logistic(sample_size, p, partition_method, partition_num, seed=0):
    """dlsa/models.py:6-40: frame ``[partition_id, label, x0..x{p-1}]`` with X~U(-.5,.5), beta = 1 on the first
    int(0.4 p) coordinates, label~Bernoulli(sigmoid(X beta)), partition_id = i % K.  Generated on the device
    (counter-based, seeded) instead of the reference's quadratic row loop over the global NumPy RNG."""
    if partition_method != "systematic":
        raise Exception("No such partition method implemented!")  # dlsa/models.py:34-35
    X, y = synth.simulate_logistic_device(int(sample_size), int(p), int(seed), _device())
    pid = np.arange(sample_size, dtype=np.float64) % partition_num
    data = np.concatenate([pid[:, None], y.cpu().numpy()[:, None], X.cpu().numpy()], axis=1)
    return pd.DataFrame(data, columns=["partition_id", "label"] + ["x" + str(i) for i in range(p)])


def _design(sample_df, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info, strict=True):
    dev = _device()
    X, names = encode_design(sample_df, Y_name, dummy_info, list(dummy_factors_baseline), data_info, dev, strict=strict)
    y = torch.as_tensor(np.array(sample_df[Y_name].to_numpy(dtype=np.float64), copy=True), device=dev)
    return X, y, names


def _compact_ok(sample_df, Y_name, dummy_info):
    """True when the chunk's design fits the compact categorical pass K1c (<= 7 numerics, 1..8 dummy columns, <= 192
    levels): the one-hot columns are then never materialised."""
    if dummy_info is None or len(dummy_info) == 0:
        return False
    conv = list(dummy_info["factor_selected"].keys())
    n_num = len(set(sample_df.columns) - {"partition_id", Y_name} - set(conv))
    return engine.cat_supported(n_num, [len(dummy_info["factor_selected_names"][c]) for c in conv])


def _fit_design(sample_df, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info, off_fn, config):
    """Encode one frame and fit every partition ``off_fn`` cuts it into: the compact pass K1c when the design has dummy
    blocks within its limits, else the dense pass K1 on the expanded matrix.  Returns (fits, names)."""
    dev = _device()
    y = torch.as_tensor(np.array(sample_df[Y_name].to_numpy(dtype=np.float64), copy=True), device=dev)
    if _compact_ok(sample_df, Y_name, dummy_info):
        num, codes, n_levels, (mean, std), names = encode_design_compact(sample_df, Y_name, dummy_info,
                                                                          list(dummy_factors_baseline), data_info, dev)
        if engine.cat_supported(num.shape[1], n_levels):
            order, off = off_fn(dev)
            if order is not None:
                num, codes, y = num[order].contiguous(), codes[order].contiguous(), y[order].contiguous()
            fits = engine.fit_logistic_partitions_cat(num if num.shape[1] else None, codes, n_levels, y, off,
                                                      fit_intercept=fit_intercept, mean=mean, std=std, config=config)
            return fits, names
    X, names = encode_design(sample_df, Y_name, dummy_info, list(dummy_factors_baseline), data_info, dev)
    order, off = off_fn(dev)
    if order is not None:
        X, y = X[order].contiguous(), y[order].contiguous()
    return engine.fit_logistic_partitions(X, y, off, fit_intercept=fit_intercept, config=config), names


def _frame(coef, smc, sig, names, par_id=None):
    p = coef.shape[0]
    out = pd.DataFrame(np.concatenate([coef[:, None], smc[:, None], sig], axis=1),
                       columns=["coef", "Sig_invMcoef"] + list(names))
    out.insert(0, "par_id", np.arange(p) if par_id is None else par_id)
    return out


def logistic_model(sample_df, Y_name, fit_intercept=False, dummy_info=[], dummy_factors_baseline=[], data_info=[],
                   config=None):
    """Run the logistic model on one partition (dlsa/models.py:42-147) and return its p x (3+p) frame
    ``['par_id', 'coef', 'Sig_invMcoef'] + X column names`` (intercept first), ``par_id = 0..p-1`` (:136-142).

    A chunk whose dummy columns differ from the expected set gives the reference's warning and all-zero frame
    (:84-91).  A fit that does not converge / is not positive definite raises ``ValueError`` (the reference surfaces
    sklearn's exceptions and ConvergenceWarning in that situation)."""
    sample_df = sample_df.reset_index(drop=True)
    icpt = ["intercept"] if fit_intercept else []
    n_rows = sample_df.shape[0]
    try:
        fits, names = _fit_design(sample_df, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info,
                                  lambda dev: (None, torch.tensor([0, n_rows], dtype=torch.int64, device=dev)), config)
    except ChunkMissingDummies as ex:
        names = design_columns(list(sample_df.columns), Y_name, dummy_info, dummy_factors_baseline, False)
        warnings.warn("Dummies:" + str(ex.missing) + "missing in this data chunk " + str(ex.shape)
                      + "Skip modeling this part of data.")
        return pd.DataFrame(0, index=np.arange(len(icpt + names)), columns=["par_id", "coef", "Sig_invMcoef"] + icpt + names)
    st = int(fits.status.item())
    if st != 0:
        raise ValueError(f"logistic_model: partition fit failed (status {st}: "
                         f"{'not converged' if st == 1 else 'Hessian not positive definite'}); "
                         "single-class or separable data has no finite MLE")
    out = _frame(fits.coef[0].cpu().numpy(), fits.Sig_invMcoef[0].cpu().numpy(), fits.Sig_inv[0].cpu().numpy(),
                 icpt + names)
    if out.isna().values.any():
        warnings.warn("NAs appear in the final output")  # dlsa/models.py:144-145
    return out


def logistic_model_partitions(data_df, Y_name, fit_intercept=False, dummy_info=[], dummy_factors_baseline=[],
                              data_info=[], config=None, return_fits=False):
    """The whole map phase in one call: the equivalent of
    ``data_sdf.groupby("partition_id").apply(logistic_model_udf)`` (projects/logistic_dlsa.py:305-316).
    Returns the union of the per-partition frames (K*p rows, partitions in ascending ``partition_id`` order), or
    the tensor-level ``PartitionFits`` with ``return_fits=True``."""
    data_df = data_df.reset_index(drop=True)
    icpt = ["intercept"] if fit_intercept else []
    pid_vals, pid_codes = np.unique(data_df["partition_id"].to_numpy(), return_inverse=True)
    K = len(pid_vals)

    def by_partition(dev):  # the device-side repartition(K, "partition_id"): stable order inside a partition
        pid = torch.as_tensor(pid_codes, device=dev).to(torch.int64)
        off = torch.zeros(K + 1, dtype=torch.int64, device=dev)
        off[1:] = torch.cumsum(torch.bincount(pid, minlength=K), 0)
        return torch.argsort(pid, stable=True), off

    fits, names = _fit_design(data_df, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info,
                              by_partition, config)
    if return_fits:
        return fits, icpt + names
    coef, smc, sig = fits.coef.cpu().numpy(), fits.Sig_invMcoef.cpu().numpy(), fits.Sig_inv.cpu().numpy()
    status = fits.status.cpu().numpy()
    frames = []
    for k in range(K):
        if status[k] != 0:
            warnings.warn(f"partition {pid_vals[k]}: fit failed (status {status[k]}); returning a zero frame")
            frames.append(pd.DataFrame(0, index=np.arange(coef.shape[1]), columns=["par_id", "coef", "Sig_invMcoef"] + icpt + names))
        else:
            frames.append(_frame(coef[k], smc[k], sig[k], icpt + names))
    return pd.concat(frames, axis=0, ignore_index=True)


def logistic_model_eval(sample_df, Y_name, par, fit_intercept=False, dummy_info=[], dummy_factors_baseline=[],
                        data_info=[]):
    """Log-likelihood of every column of ``par`` on one partition (dlsa/models.py:151-225): a one-row frame whose
    columns are ``par.columns``.  Chunks with missing dummy levels are evaluated with those columns at 0 (:189-197)."""
    sample_df = sample_df.reset_index(drop=True)
    try:
        X, y, names = _design(sample_df, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info)
    except ChunkMissingDummies as ex:
        warnings.warn("Dummies:" + str(ex.missing) + "missing in this data chunk " + str(ex.shape))
        X, y, names = _design(sample_df, Y_name, fit_intercept, dummy_info, dummy_factors_baseline, data_info, strict=False)
    off = torch.tensor([0, X.shape[0]], dtype=torch.int64, device=X.device)
    P = torch.as_tensor(np.ascontiguousarray(np.asarray(par, dtype=np.float64)), device=X.device)
    ll = engine.logistic_loglik(X, y, off, P, fit_intercept=fit_intercept)[0].cpu().numpy()
    return pd.DataFrame({c: [v] for c, v in zip(par.columns, ll)})


def linear_model(sample_df, Y_name, dummy_factors_baseline=[]):
    """OLS on one partition in the frame layout of ``logistic_model``: coef = (X'X)^-1 X'y,
    Sig_inv = X'X / sigma2 with sigma2 = RSS/(n-p) — R ``lm(y~x-1)`` + ``solve(vcov)`` (dlsa/R/wlse_func.R:169-182;
    per-group OLS precedent projects/results/linear_regression_dc.py:27-39)."""
    sample_df = sample_df.reset_index(drop=True)
    X, y, names = _design(sample_df, Y_name, False, [], dummy_factors_baseline, [])
    off = torch.tensor([0, X.shape[0]], dtype=torch.int64, device=X.device)
    fits = engine.fit_linear_partitions(X, y, off)
    if int(fits.status.item()) != 0:
        raise ValueError("linear_model: X'X is not positive definite")
    return _frame(fits.coef[0].cpu().numpy(), fits.Sig_invMcoef[0].cpu().numpy(), fits.Sig_inv[0].cpu().numpy(), names)


def poisson_model(sample_df, Y_name, fit_intercept=False, dummy_factors_baseline=[], config=None):
    """Poisson (log link) regression on one partition in the frame layout of ``logistic_model``: coef = the MLE of
    R ``glm(y ~ x - 1, family = "poisson")``, Sig_inv = solve(vcov) = X' diag(mu) X (dlsa/R/wlse_func.R:155-168)."""
    sample_df = sample_df.reset_index(drop=True)
    X, y, names = _design(sample_df, Y_name, fit_intercept, [], dummy_factors_baseline, [])
    off = torch.tensor([0, X.shape[0]], dtype=torch.int64, device=X.device)
    fits = engine.fit_poisson_partitions(X, y, off, fit_intercept=fit_intercept, config=config)
    if int(fits.status.item()) != 0:
        raise ValueError("poisson_model: the Newton iteration did not converge")
    return _frame(fits.coef[0].cpu().numpy(), fits.Sig_invMcoef[0].cpu().numpy(), fits.Sig_inv[0].cpu().numpy(),
                  (["intercept"] if fit_intercept else []) + names)
