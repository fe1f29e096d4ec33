"""``dlsa.dummies`` / ``dlsa.sdummies`` drop-in: level bookkeeping (host dictionaries, as in the reference) and the
device-side one-hot expansion into the reference's exact column layout.

Layout contract (dlsa/models.py:56-79): X columns = sorted non-dummy feature names, then for each
dummy column (dict order of ``dummy_info['factor_selected']``) its string-sorted ``factor_selected_names`` minus
``dummy_factors_baseline``; the intercept, when fitted, is coefficient 0.
"""
from __future__ import annotations

import os
import pickle
import warnings
from collections import Counter

import numpy as np
import torch

from ._lib import lib, check


def dummy_factors_counts(pdf, dummy_columns):
    """dlsa/dummies.py:10-31: per-column ``value_counts().to_dict()`` (key order = descending count)."""
    cols = list(pdf.columns)
    if all(isinstance(c, int) for c in dummy_columns):
        names = [cols[i] for i in dummy_columns]
    else:
        names = list(dummy_columns)
    return {c: pdf[c].value_counts().to_dict() for c in names}


def cumsum_dicts(dict1, dict2):
    """dlsa/dummies.py:34-49: merge two count dictionaries (Counter addition keeps dict1's key order first)."""
    if len(dict1) == 0:
        return dict2
    if len(dict2) == 0:
        return dict1
    return {c: dict(Counter(dict1[c]) + Counter(dict2[c])) for c in dict1.keys()}


def select_dummy_factors(dummy_dict, keep_top, replace_with="000_OTHERS", pickle_file=None):
    """dlsa/dummies.py:52-108.  Levels (in dict insertion order) whose cumulative share is <= keep_top[i] are
    selected (:72-81), the rest dropped (:78,83-84); names are ``<col>_<level>`` with ``<col>_<replace_with>``
    first when anything was dropped (:87-96).  ``pickle_file`` is optional here (the reference always writes it)."""
    factor_set, factor_selected, factor_dropped, factor_selected_names = {}, {}, {}, {}
    for i, col in enumerate(list(dummy_dict)):
        levels = list(dummy_dict[col].keys())
        share = np.cumsum(np.asarray(list(dummy_dict[col].values()), dtype=np.float64))
        share = share / share[-1]
        factor_set[col] = levels
        factor_selected[col] = [lv for lv, s in zip(levels, share) if s <= keep_top[i]]
        factor_dropped[col] = [lv for lv, s in zip(levels, share) if s > keep_top[i]]
        new = ([replace_with] if factor_dropped[col] else []) + factor_selected[col]
        factor_selected_names[col] = [col + "_" + str(x) for x in new]
    info = {"factor_set": factor_set, "factor_selected": factor_selected, "factor_dropped": factor_dropped,
            "factor_selected_names": factor_selected_names}
    if pickle_file is not None:
        with open(os.path.expanduser(pickle_file), "wb") as f:
            pickle.dump(info, f)
    return info


def select_sdummy_levels(counts, keep_top, replace_with="000_OTHERS", dropLast=False):
    """The level-selection rule of ``get_sdummies`` (dlsa/sdummies.py:41-83) without Spark: ``counts`` maps each
    dummy column to {level: count}.  Levels sorted by descending count; kept when cumulative share <= keep_top[i]
    OR count >= 10% of the rows (:50-53); ``factor_selected_names`` are un-prefixed level names ordered by
    descending share with ``replace_with`` standing for the dropped mass (:65-81)."""
    factor_set, factor_selected, factor_dropped, factor_selected_names = {}, {}, {}, {}
    for i, col in enumerate(list(counts)):
        items = sorted(counts[col].items(), key=lambda kv: -kv[1])
        total = float(sum(v for _, v in items))
        cum = np.cumsum([v for _, v in items]) / total
        keep = [lv for (lv, v), c in zip(items, cum) if c <= keep_top[i] or v >= 0.1 * total]
        factor_set[col] = [lv for lv, _ in items]
        factor_selected[col] = keep
        factor_dropped[col] = [lv for lv in factor_set[col] if lv not in set(keep)]
        table = [(lv, v / total) for lv, v in items[: len(keep)]]
        if factor_dropped[col]:
            table.append((replace_with, 1 - sum(s for _, s in table)))
        table.sort(key=lambda kv: -kv[1])
        names = [lv for lv, _ in table]
        factor_selected_names[col] = names[:-1] if dropLast else names
    return {"factor_set": factor_set, "factor_selected": factor_selected, "factor_dropped": factor_dropped,
            "factor_selected_names": factor_selected_names}


def design_columns(raw_columns, Y_name, dummy_info, dummy_factors_baseline, fit_intercept):
    """Canonical X column names (dlsa/models.py:71-79 and dlsa/utils_spark.py:5-33 in one sorted order)."""
    feats = [c for c in raw_columns if c not in ("partition_id", Y_name)]
    if dummy_info is not None and len(dummy_info) > 0:
        conv = list(dummy_info["factor_selected"].keys())
        cols = sorted(set(feats) - set(conv))
        for c in conv:
            cols += sorted(dummy_info["factor_selected_names"][c])
        cols = [c for c in cols if c not in dummy_factors_baseline]
    else:
        cols = [c for c in feats if c not in dummy_factors_baseline]
    return (["intercept"] if fit_intercept else []) + cols


def convert_schema(usecols_x, dummy_info, fit_intercept, dummy_factors_baseline):
    """``dlsa.utils_spark.convert_schema`` (dlsa/utils_spark.py:5-33) without pyspark: the output-frame field list
    as ``(name, 'double')`` pairs, in the ONE canonical (sorted) order used for rows and columns (the reference's
    ``list(set(...))`` hash order for the numeric block is an accident)."""
    if len(dummy_info) == 0:
        names = list(usecols_x)
    else:
        conv = list(dummy_info["factor_selected"].keys())
        names = sorted(set(usecols_x) - set(conv))
        for c in conv:
            names += [n for n in sorted(dummy_info["factor_selected_names"][c]) if n not in dummy_factors_baseline]
    if fit_intercept:
        names.insert(0, "intercept")
    return [(n, "double") for n in names]


class ChunkMissingDummies(Exception):
    def __init__(self, missing, shape):
        self.missing, self.shape = missing, shape


def _encode_host(sample_df, Y_name, dummy_info, dummy_factors_baseline, data_info, strict):
    """Host-side (string) part of the design encoding shared by the dense and the compact path: returns
    (names, numeric columns, numeric ndarray [n, n_num], codes ndarray [n, n_cat] of ABSOLUTE column indices into
    ``names`` (-1 = baseline level), categorical columns, (mean, std) ndarrays or None)."""
    has_dummies = dummy_info is not None and len(dummy_info) > 0
    names = design_columns(list(sample_df.columns), Y_name, dummy_info if has_dummies else None,
                           dummy_factors_baseline, False)
    n = sample_df.shape[0]
    if has_dummies:
        conv = list(dummy_info["factor_selected"].keys())
        numeric = sorted(set(sample_df.columns) - {"partition_id", Y_name} - set(conv))
    else:
        conv = []
        numeric = names
    n_num = len(numeric)
    pos = {nm: j for j, nm in enumerate(names)}
    codes = np.full((n, len(conv)), -1, dtype=np.int32)
    present = set(numeric)
    for ci, c in enumerate(conv):
        vals = sample_df[c]
        dropped = dummy_info["factor_dropped"].get(c, [])
        if len(dropped) > 0:
            vals = vals.where(~vals.isin(dropped), "000_OTHERS")
        colnames = (c + "_" + vals.astype(str)).to_numpy()
        uniq, inv = np.unique(colnames, return_inverse=True)
        present |= set(uniq.tolist())
        lut = np.array([pos.get(u, -1) for u in uniq], dtype=np.int32)
        codes[:, ci] = lut[inv]
    if has_dummies and strict:
        missing_base = [b for b in dummy_factors_baseline if b not in present]
        if missing_base:
            raise KeyError(f"{missing_base} not found in axis")  # pandas .drop() in dlsa/models.py:68
        present -= set(dummy_factors_baseline)
        if present != set(names):
            raise ChunkMissingDummies(set(names) - present, (n, len(present)))
    num = np.ascontiguousarray(sample_df[numeric].to_numpy(dtype=np.float64)) if n_num else np.zeros((n, 0))
    stats = None
    if data_info is not None and len(data_info) > 0:
        stats = (np.array([float(data_info[c][1]) for c in numeric]), np.array([float(data_info[c][2]) for c in numeric]))
    return names, numeric, num, codes, conv, stats


def encode_design(sample_df, Y_name, dummy_info, dummy_factors_baseline, data_info, device, strict=True):
    """Build the dense float64 design matrix of one chunk on the device in the reference layout.

    Host side (strings): replace dropped levels by ``000_OTHERS`` (dlsa/models.py:60), map every categorical value
    to the index of its ``<col>_<level>`` column (-1 for a baseline level).  Device side: kernel
    ``dlsa_onehot_expand_f64`` writes the 1s and the (optionally standardised, :99-101) numerics.
    Returns (X [n, px] cuda tensor, column names).  Raises ChunkMissingDummies when the chunk's dummy set differs
    from the expected one (the caller turns it into the reference's warning + zero frame, :84-91); with
    ``strict=False`` absent levels simply leave their columns at zero (logistic_model_eval, :189-197)."""
    names, numeric, num, codes, conv, stats = _encode_host(sample_df, Y_name, dummy_info, dummy_factors_baseline,
                                                           data_info, strict)
    n, n_num = sample_df.shape[0], len(numeric)
    px = len(names)
    # an odd width is stored with an even row pitch: rows stay 16-byte aligned, so the fused pass stages them with TMA
    # row copies instead of 8-byte cp.async (4x faster at p = 181)
    X = torch.zeros(n, px + (px & 1), dtype=torch.float64, device=device)[:, :px]
    num_d = torch.as_tensor(num, device=device)
    mean_d = std_d = None
    if stats is not None:
        mean_d = torch.as_tensor(stats[0], device=device)
        std_d = torch.as_tensor(stats[1], device=device)
    # numeric block occupies names[:n_num] (sorted numerics come first); codes are absolute column indices
    codes_d = torch.as_tensor(codes, device=device)
    base_d = torch.zeros(max(len(conv), 1), dtype=torch.int32, device=device)
    check(lib.dlsa_onehot_expand_f64(num_d.data_ptr() if n_num else None, n_num,
                                     mean_d.data_ptr() if mean_d is not None else None,
                                     std_d.data_ptr() if std_d is not None else None,
                                     codes_d.data_ptr() if conv else None, len(conv),
                                     base_d.data_ptr() if conv else None, n, X.data_ptr(), X.stride(0),
                                     torch.cuda.current_stream().cuda_stream), "dlsa_onehot_expand_f64")
    return X, names


def encode_design_compact(sample_df, Y_name, dummy_info, dummy_factors_baseline, data_info, device, strict=True):
    """The same design as ``encode_design`` in the COMPACT form the categorical pass K1c consumes
    (``dlsa_irls_pass_cat_f64``): returns (numeric [n, n_num] cuda float64 (NOT yet standardised), codes [n, n_cat] cuda
    int32 with the level's index inside its categorical block or -1, n_levels list, (mean, std) cuda tensors or
    (None, None), column names).  The level blocks follow the numerics in ``names`` in ``dummy_columns`` order
    (dlsa/models.py:71-79), which is exactly the column order K1c emits."""
    names, numeric, num, codes, conv, stats = _encode_host(sample_df, Y_name, dummy_info, dummy_factors_baseline,
                                                           data_info, strict)
    n_num = len(numeric)
    n_levels, start = [], n_num
    for c in conv:
        L = len([nm for nm in dummy_info["factor_selected_names"][c] if nm not in dummy_factors_baseline])
        n_levels.append(L)
        start += L
    rel = codes.copy()
    start = n_num
    for ci, L in enumerate(n_levels):
        col = rel[:, ci]
        col[col >= 0] -= start
        start += L
    mean_d = std_d = None
    if stats is not None:
        mean_d = torch.as_tensor(stats[0], device=device)
        std_d = torch.as_tensor(stats[1], device=device)
    return (torch.as_tensor(num, device=device), torch.as_tensor(np.ascontiguousarray(rel), device=device), n_levels,
            (mean_d, std_d), names)


def describe(data, columns=None, device=None):
    """``data_sdf_i.describe().toPandas()`` on the device (projects/logistic_dlsa.py:279-283): a frame with a
    ``summary`` column (count, mean, stddev, min, max) and one string-valued column per numeric column, so that
    ``float(data_info[c][1])`` / ``float(data_info[c][2])`` (dlsa/models.py:99-101) read the mean and the sample
    standard deviation.  ``data``: a pandas frame (its numeric ``columns``) or a float64 CUDA tensor [n, px]."""
    import pandas as pd
    from . import engine
    if isinstance(data, torch.Tensor):
        X = data
        names = list(columns) if columns is not None else [f"x{i}" for i in range(X.shape[1])]
    else:
        names = list(columns) if columns is not None else [c for c in data.columns
                                                            if pd.api.types.is_numeric_dtype(data[c])]
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        X = torch.as_tensor(np.ascontiguousarray(data[names].to_numpy(dtype=np.float64)), device=dev)
    st = engine.column_stats(X).cpu().numpy()
    out = {"summary": ["count", "mean", "stddev", "min", "max"]}
    for j, c in enumerate(names):
        out[c] = [str(int(st[0, j]))] + [repr(float(v)) for v in st[1:, j]]
    return pd.DataFrame(out)
