"""``dlsa.dlsa`` drop-in: ``dlsa_mapred`` (alias ``dlsa_mapreduce``) and ``dlsa`` (dlsa/dlsa.py:21-107)."""
from __future__ import annotations

import numpy as np
import pandas as pd
import torch

from . import _lib, engine
from .lsa import _device, _to_dev


def _fits_from_frame(pdf: pd.DataFrame):
    """Stacked per-partition frames (K*p rows x (3+p)) -> PartitionFits tensors.  The zero frame of a skipped chunk
    (dlsa/models.py:84-91: p all-zero rows) carries no information and is dropped as a BLOCK of p consecutive rows; the
    rows that remain are grouped by ``par_id`` exactly as ``groupby('par_id')`` does (dlsa/dlsa.py:30).  A rank whose
    local frames are all zero returns K = 0 tensors: its packed buffer is zero and the "empty" error of
    dlsa/dlsa.py:36-39 is raised after the allreduce, on every rank at once (engine.combine_partitions)."""
    cols = list(pdf.columns)
    p = len(cols) - 3
    vals = pdf.to_numpy(dtype=np.float64)
    if vals.shape[0] % p != 0:
        raise ValueError("dlsa_mapred: the mapped frame must stack whole p-row partition frames")
    blocks = vals.reshape(-1, p, 3 + p)
    blocks = blocks[np.any(blocks[:, :, 1:] != 0.0, axis=(1, 2))]
    K = blocks.shape[0]
    if K and not np.array_equal(np.sort(blocks[:, :, 0], axis=1), np.broadcast_to(np.arange(p), (K, p))):
        raise ValueError("dlsa_mapred: every par_id must appear once per fitted partition")
    if K:
        order = np.argsort(blocks[:, :, 0], axis=1, kind="stable")
        blocks = np.take_along_axis(blocks, order[:, :, None], axis=1)   # rows of each frame in par_id order
    dev = _device()
    return engine.PartitionFits(
        coef=torch.as_tensor(np.ascontiguousarray(blocks[:, :, 1]), device=dev).reshape(K, p),
        Sig_inv=torch.as_tensor(np.ascontiguousarray(blocks[:, :, 3:]), device=dev).reshape(K, p, p),
        Sig_invMcoef=torch.as_tensor(np.ascontiguousarray(blocks[:, :, 2]), device=dev).reshape(K, p),
        status=torch.zeros(K, dtype=torch.int32, device=dev), iters=torch.zeros(K, dtype=torch.int32, device=dev),
        loglik=None, passes=0, n_rows=0), cols[3:]


def dlsa_mapred(model_mapped_sdf, num_partitions=None, group=None):
    """One-round combine of the mapped partitions (dlsa/dlsa.py:21-61).

    ``model_mapped_sdf`` may be: the union of per-partition frames as a pandas DataFrame; a list of such frames; a
    Spark-like object exposing ``.toPandas()`` (and ``.rdd.getNumPartitions()``); or ``(PartitionFits, names)`` /
    ``PartitionFits`` from the tensor engine.  Returns the p x (2+p) frame
    ``['beta_byOLS', 'beta_byONESHOT'] + Sig_inv column names`` with rows in ``par_id`` order (:55-59).
    ``beta_byONESHOT`` divides by the number of partitions (``getNumPartitions()``, :52): ``num_partitions`` if
    given, else the Spark object's value, else the number of frames (zero frames included)."""
    names = None
    if isinstance(model_mapped_sdf, tuple) and isinstance(model_mapped_sdf[0], engine.PartitionFits):
        fits, names = model_mapped_sdf
        K_all = fits.coef.shape[0]
    elif isinstance(model_mapped_sdf, engine.PartitionFits):
        fits = model_mapped_sdf
        K_all = fits.coef.shape[0]
    else:
        if isinstance(model_mapped_sdf, (list, tuple)):
            pdf = pd.concat(list(model_mapped_sdf), axis=0, ignore_index=True)
        elif isinstance(model_mapped_sdf, pd.DataFrame):
            pdf = model_mapped_sdf
        elif hasattr(model_mapped_sdf, "toPandas"):
            if num_partitions is None and hasattr(model_mapped_sdf, "rdd"):
                num_partitions = model_mapped_sdf.rdd.getNumPartitions()
            pdf = model_mapped_sdf.toPandas()
        else:
            raise TypeError("dlsa_mapred: unsupported input type " + type(model_mapped_sdf).__name__)
        p = pdf.shape[1] - 3
        K_all = pdf.shape[0] // p
        fits, names = _fits_from_frame(pdf)
    from . import parallel
    if num_partitions is None and parallel.world_size(group) > 1:
        denom = None  # all-reduced partition count
    else:
        denom = K_all if num_partitions is None else int(num_partitions)
    comb = engine.combine_partitions(fits, group=group, num_partitions=denom)
    p = comb.beta_byOLS.numel()
    if names is None:
        names = ["x" + str(i) for i in range(p)]
    out = np.concatenate([comb.beta_byOLS.cpu().numpy()[:, None], comb.beta_byONESHOT.cpu().numpy()[:, None],
                          comb.Sig_inv_sum.cpu().numpy()], axis=1)
    return pd.DataFrame(out, columns=["beta_byOLS", "beta_byONESHOT"] + list(names))


dlsa_mapreduce = dlsa_mapred  # name used by the dstats successor (README.md:24-28)


def dlsa(Sig_inv_, beta_, sample_size, fit_intercept=False, type="lar", verbose=False):
    """Distributed least squares approximation: LARS path on (Sig_inv_, beta_) and first-argmin AIC / BIC
    selection (dlsa/dlsa.py:70-107).  Returns the frame ``['beta_byAIC', 'beta_byBIC']`` (p rows; with an
    intercept row 0 is ``beta0 + beta_[0]`` as in the R original, dlsa/R/dlsa_alasso_func.R:54-58)."""
    S = _to_dev(np.asarray(Sig_inv_))
    b = _to_dev(np.asarray(beta_).ravel()).view(-1)
    res = engine.lars_path(S, b, bool(fit_intercept), float(sample_size), type=type)
    ia, ib = res.argmin_AIC, res.argmin_BIC
    beta = res.beta.cpu().numpy()
    if verbose:
        print({"AIC": res.AIC.cpu().numpy(), "BIC": res.BIC.cpu().numpy(), "beta": beta, "beta0": res.beta0.cpu().numpy()})
    if fit_intercept:
        beta0 = res.beta0.cpu().numpy() + float(b[0].item())
        byA = np.hstack([beta0[ia], beta[ia, :]])
        byB = np.hstack([beta0[ib], beta[ib, :]])
    else:
        byA, byB = beta[ia, :], beta[ib, :]
    return pd.DataFrame({"beta_byAIC": byA, "beta_byBIC": byB})
