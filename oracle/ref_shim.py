"""TEST INFRASTRUCTURE ONLY — runs the *reference's own functions* under in-process compat shims.

Usable where the reference tree exists: ``/root/reference`` in the build container, or the git-ignored
``baseline/_ref`` install of the unmodified reference that travels to the GPU box.  It is how the oracle
(``oracle/numpy_dlsa.py``) is pinned, how ``tests/golden/*.npz`` are minted (``oracle/make_golden.py``) and what
``bench.py --impl reference`` / the ``cpu_baseline`` leg time (``oracle/ref_cpu.py``).  Nothing in the product path,
the ``-m gpu`` tests or ``smoke()`` imports this module.

The shims do not change the reference's arithmetic:
  1. ``np.float`` / ``np.NAN`` aliases so ``dlsa/lsa.py:12,23,90`` import under numpy 2.
  2. ``LogisticRegression(penalty='none')`` -> ``penalty=None`` (sklearn >= 1.2 spelling) and an
     explicit ``tol`` (the reference never sets it, ``dlsa/models.py:110-112``; the oracle setting
     is 1e-10, see SURVEY.md §7 hard part 1).
  3. ``pd.concat(objs, 1)`` positional axis (``dlsa/models.py:142``) accepted again.
  4. ``pd.get_dummies(sparse=True)`` -> dense float (pandas 3 cannot transpose Sparse[bool]).
  5. ``lars_lsa`` is called with ``np.matrix`` inputs (``dlsa/lsa.py:108,114`` need matrix ``*``).
  6. ``dlsa_mapred``'s Spark calls are emulated with a pandas group-sum; lines
     ``dlsa/dlsa.py:44-59`` are then followed literally.
  7. ``pd.DataFrame({name: scalar})`` (``dlsa/models.py:224``: ``np.sum(DataFrame)`` is a scalar under pandas >= 2,
     a one-element Series under the pinned pandas) -> one-row frame.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
import warnings

import numpy as np
import pandas as pd

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _find_reference_root() -> str:
    """The reference tree: $DLSA_REFERENCE_ROOT, else /root/reference (the build container), else the git-ignored
    install under baseline/_ref (``pip install --target baseline/_ref`` of the unmodified reference, made by
    ``__graft_entry__.build()``; it travels to the GPU box with the snapshot so that ``bench.py --impl reference`` can
    time the reference's own functions there)."""
    cands = [os.environ.get("DLSA_REFERENCE_ROOT"), "/root/reference", os.path.join(_ROOT, "baseline", "_ref")]
    for c in cands:
        if c and os.path.isfile(os.path.join(c, "dlsa", "models.py")):
            return c
    return "/root/reference"


REFERENCE_ROOT = _find_reference_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "dlsa", "models.py"))


def _load_by_path(name: str, relpath: str) -> types.ModuleType:
    path = os.path.join(REFERENCE_ROOT, relpath)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _PandasCompat(types.ModuleType):
    """Module proxy over pandas that restores two removed call forms used by the reference."""

    def __init__(self):
        super().__init__("pandas_compat")

    def __getattr__(self, item):
        return getattr(pd, item)

    @staticmethod
    def concat(objs, axis=0, **kw):  # dlsa/models.py:142 passes axis positionally
        return pd.concat(objs, axis=axis, **kw)

    @staticmethod
    def DataFrame(data=None, *a, **kw):
        # dlsa/models.py:222-224 builds a frame from {name: np.sum(one-column DataFrame)}: a one-element Series per name
        # under the pinned pandas, a scalar under pandas >= 2 (np.sum(df) reduces over both axes) -> one-row frame
        if isinstance(data, dict) and data and all(np.isscalar(v) for v in data.values()) and not a and "index" not in kw:
            return pd.DataFrame(data, index=[0], **kw)
        return pd.DataFrame(data, *a, **kw)

    @staticmethod
    def get_dummies(**kw):  # dlsa/models.py:62-65 asks for sparse=True
        kw = dict(kw)
        kw["sparse"] = False
        kw["dtype"] = float
        return pd.get_dummies(**kw)


_models = None
_lsa = None
_dummies = None


def ref_models(tol: float = 1e-10):
    """The reference ``dlsa/models.py`` module with shims 2-4 applied."""
    global _models
    if _models is None:
        _models = _load_by_path("_ref_dlsa_models", "dlsa/models.py")
        _models.pd = _PandasCompat()
    from sklearn.linear_model import LogisticRegression as _LR

    def _lr(**kw):
        kw = dict(kw)
        if kw.get("penalty") == "none":
            kw.pop("penalty")
            kw["C"] = np.inf  # sklearn 1.9: penalty=None is spelled C=inf
        kw["tol"] = tol
        return _LR(**kw)

    _models.LogisticRegression = _lr
    return _models


def ref_lsa():
    """The reference ``dlsa/lsa.py`` module with shim 1 applied."""
    global _lsa
    if _lsa is None:
        if not hasattr(np, "float"):
            np.float = float  # type: ignore[attr-defined]
        if not hasattr(np, "NAN"):
            np.NAN = np.nan  # type: ignore[attr-defined]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            _lsa = _load_by_path("_ref_dlsa_lsa", "dlsa/lsa.py")
    return _lsa


def ref_dummies():
    global _dummies
    if _dummies is None:
        _dummies = _load_by_path("_ref_dlsa_dummies", "dlsa/dummies.py")
    return _dummies


def ref_logistic_model(sample_df, Y_name, fit_intercept=False, dummy_info=[], dummy_factors_baseline=[],
                       data_info=[], tol: float = 1e-10):
    """``dlsa.models.logistic_model`` (dlsa/models.py:42-147) on a frame with a fresh RangeIndex."""
    m = ref_models(tol)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", category=FutureWarning)
        warnings.simplefilter("ignore", category=pd.errors.PerformanceWarning)
        return m.logistic_model(sample_df.reset_index(drop=True), Y_name, fit_intercept=fit_intercept,
                                dummy_info=dummy_info, dummy_factors_baseline=dummy_factors_baseline,
                                data_info=data_info)


def ref_mapred(frames, num_partitions=None):
    """``dlsa.dlsa.dlsa_mapred`` (dlsa/dlsa.py:21-61) with the Spark group-sum emulated in pandas."""
    frames = list(frames)
    K = len(frames) if num_partitions is None else num_partitions
    cols = list(frames[0].columns)
    allrows = pd.concat(frames, axis=0, ignore_index=True)
    g = allrows.groupby("par_id").sum().sort_index()  # columns cols[1:], rows by par_id
    if g.shape[0] == 0:
        raise Exception("Zero-length grouped pandas DataFrame obtained, check the input.")
    Sig_invMcoef_sum = g.iloc[:, 1]
    Sig_inv_sum = g.iloc[:, 2:]
    beta_byOLS = np.linalg.lstsq(Sig_inv_sum, Sig_invMcoef_sum, rcond=None)[0]
    beta_byONESHOT = g["coef"] / K
    p = len(Sig_invMcoef_sum)
    return pd.DataFrame(
        np.concatenate((beta_byOLS.reshape(p, 1), np.asarray(beta_byONESHOT).reshape(p, 1), Sig_inv_sum), 1),
        columns=["beta_byOLS", "beta_byONESHOT"] + cols[3:])


def ref_lars_lsa(Sigma0, b0, intercept, n, type="lar", max_steps=None):
    """``dlsa.lsa.lars_lsa`` (dlsa/lsa.py:90-212) called the only way it runs: np.matrix Sigma.

    ``intercept=True`` is broken in the Python reference and is not offered here.
    """
    if intercept:
        raise NotImplementedError("reference lars_lsa(intercept=True) raises IndexError; R is the spec")
    lsa = ref_lsa()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = lsa.lars_lsa(np.matrix(np.asarray(Sigma0, dtype=float)), np.asarray(b0, dtype=float).ravel(),
                           False, n, type=type, max_steps=max_steps)
    return {"AIC": np.asarray(out["AIC"]).ravel(), "BIC": np.asarray(out["BIC"]).ravel(),
            "beta": np.asarray(out["beta"]), "beta0": np.asarray(out["beta0"]).ravel()}
