"""Ingest for the hot path: CSV / Arrow -> device columns in the compact design the fused
categorical pass consumes, without a pandas ``get_dummies`` detour.

Mirrors what the reference's driver does before the map (projects/logistic_dlsa.py:213-237, 279-290) and what
``dlsa/dummies.py`` does out of core:

    select usecols_x + [Y_name]; dropna()                              :218-222
    label = (Y > 0)                                                    :226-228
    K = ceil(n / sample_size_per_partition)                            :231-232
    partition_id = monotonically_increasing_id() % K                   :235-237   (row index modulo K here)
    data_info = describe()                                             :279-283   (count / mean / stddev on the device)
    dummy_factors_counts + select_dummy_factors                        dlsa/dummies.py:10-108

Host side: parsing (pyarrow.csv) and the string -> dictionary-index step (Arrow dictionary encoding) — strings never
reach the GPU.  Device side: level counting (one histogram per categorical column), the level -> design-column remap
(a look-up table applied on the device: selected level -> its sorted position, dropped level -> ``000_OTHERS``, baseline
-> -1), the ``(x - mean) / std`` statistics, the repartition by ``partition_id`` and the fit itself.
"""
from __future__ import annotations

import dataclasses
import math
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import dummies, engine


@dataclasses.dataclass
class IngestedDesign:
    """A data set ready for ``fit``: rows grouped by partition (partition k = rows part_offsets[k]:part_offsets[k+1], in
    their original order), numerics raw (standardised on load when ``data_info`` is used), categoricals as level codes."""
    numeric: torch.Tensor            # [n, n_num] float64, columns in sorted-name order (dlsa/models.py:71)
    codes: torch.Tensor              # [n, n_cat] int32: position among the column's design levels, -1 = baseline level
    n_levels: List[int]              # design levels per categorical column (baselines removed)
    y: torch.Tensor                  # [n] float64 in {0, 1}
    part_offsets: torch.Tensor       # [K+1] int64
    names: List[str]                 # design column names in output order (no intercept)
    numeric_names: List[str]
    dummy_columns: List[str]
    dummy_info: dict                 # the dict dlsa.dummies.select_dummy_factors returns
    dummy_factors_baseline: List[str]
    data_info: dict                  # {numeric column: array(count, mean, stddev, min, max)} — describe() rows
    n_rows_read: int                 # rows in the file before dropna()
    level_counts: dict               # {column: {level: count}} in descending-count order (dummy_factors_counts)


def _arrow_table(source, columns: Sequence[str]):
    import pyarrow as pa
    import pyarrow.csv as pacsv
    if isinstance(source, pa.Table):
        return source.select(list(columns))
    if hasattr(source, "to_pandas") and not isinstance(source, (str, bytes)):
        return source.select(list(columns))
    try:
        import pandas as pd
        if isinstance(source, pd.DataFrame):
            return pa.Table.from_pandas(source[list(columns)], preserve_index=False)
    except ImportError:
        pass
    return pacsv.read_csv(source, convert_options=pacsv.ConvertOptions(include_columns=list(columns)))


def level_counts_device(codes: torch.Tensor, n_dict: int) -> torch.Tensor:
    """Histogram of dictionary indices on the device (the map side of dummy_factors_counts, dlsa/dummies.py:10-31)."""
    return torch.bincount(codes.to(torch.int64), minlength=n_dict)


def ingest(source, usecols_x: Sequence[str], Y_name: str, dummy_columns: Sequence[str] = (),
           dummy_keep_top: Sequence[float] = (), dummy_factors_baseline: Optional[Sequence[str]] = None,
           fit_intercept: bool = True, sample_size_per_partition: int = 1_000_000, partition_num: Optional[int] = None,
           binarize_label: bool = True, replace_with: str = "000_OTHERS", device=None) -> IngestedDesign:
    """CSV path / Arrow table / pandas frame -> ``IngestedDesign`` on ``device``.

    ``dummy_factors_baseline`` defaults to the first design level of every dummy column when an intercept is fitted
    (the driver's choice, projects/logistic_dlsa.py:157-165, spelled per column there) and to none otherwise."""
    import pyarrow as pa
    import pyarrow.compute as pc
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    usecols_x, dummy_columns = list(usecols_x), list(dummy_columns)
    if len(dummy_keep_top) != len(dummy_columns):
        raise ValueError("dummy_keep_top needs one entry per dummy column")
    tab = _arrow_table(source, usecols_x + [Y_name])
    n_read = tab.num_rows
    # dropna(): a row survives when every selected column is non-null (projects/logistic_dlsa.py:222)
    keep = None
    for c in tab.column_names:
        v = pc.is_valid(tab[c])
        keep = v if keep is None else pc.and_(keep, v)
    if keep is not None and pc.sum(pc.invert(keep)).as_py():
        tab = tab.filter(keep)
    n = tab.num_rows
    if n == 0:
        raise ValueError("ingest: no complete rows")
    numeric_names = sorted(c for c in usecols_x if c not in dummy_columns)
    # ---- host -> device ------------------------------------------------------------------------------------------
    num_h = np.empty((n, len(numeric_names)), dtype=np.float64)
    for j, c in enumerate(numeric_names):
        num_h[:, j] = tab[c].to_numpy(zero_copy_only=False).astype(np.float64, copy=False)
    numeric = torch.as_tensor(num_h, device=dev)
    yv = torch.as_tensor(np.array(tab[Y_name].to_numpy(zero_copy_only=False), dtype=np.float64), device=dev)
    y = (yv > 0).to(torch.float64) if binarize_label else yv          # F.when(Y > 0, 1).otherwise(0), :226-228
    dict_codes, dict_levels = [], []
    for c in dummy_columns:
        enc = pc.dictionary_encode(tab[c].combine_chunks())
        dict_levels.append([str(v) for v in enc.dictionary.to_pylist()])
        dict_codes.append(torch.as_tensor(enc.indices.to_numpy(zero_copy_only=False).astype(np.int32), device=dev))
    # ---- level counts on the device, selection rule on the host (a few hundred integers) -----------------------
    counts = {}
    for c, cd, lv in zip(dummy_columns, dict_codes, dict_levels):
        h = level_counts_device(cd, len(lv)).cpu().numpy()
        order = np.argsort(-h, kind="stable")                          # value_counts(): descending count
        counts[c] = {lv[i]: int(h[i]) for i in order}
    info = dummies.select_dummy_factors(counts, list(dummy_keep_top), replace_with) if dummy_columns else {}
    if dummy_factors_baseline is None:
        dummy_factors_baseline = ([sorted(info["factor_selected_names"][c])[0] for c in dummy_columns]
                                  if (fit_intercept and dummy_columns) else [])
    dummy_factors_baseline = list(dummy_factors_baseline)
    # ---- dictionary index -> design level (device LUT) ---------------------------------------------------------
    names = list(numeric_names)
    n_levels, code_cols = [], []
    for c, cd, lv in zip(dummy_columns, dict_codes, dict_levels):
        design = [nm for nm in sorted(info["factor_selected_names"][c]) if nm not in dummy_factors_baseline]
        pos = {nm: i for i, nm in enumerate(design)}
        dropped = set(str(v) for v in info["factor_dropped"][c])
        lut = np.empty(len(lv), dtype=np.int32)
        for i, level in enumerate(lv):
            nm = c + "_" + (replace_with if level in dropped else level)
            lut[i] = pos.get(nm, -1)                                   # baseline (or unknown) level -> no column
        code_cols.append(torch.as_tensor(lut, device=dev)[cd.to(torch.int64)])
        n_levels.append(len(design))
        names += design
    codes = (torch.stack(code_cols, dim=1).contiguous() if code_cols
             else torch.empty(n, 0, dtype=torch.int32, device=dev))
    # ---- describe() on the device -------------------------------------------------------------------------------
    data_info = {}
    if numeric_names:
        st = engine.column_stats(numeric).cpu().numpy()
        data_info = {c: st[:, j].copy() for j, c in enumerate(numeric_names)}
    # ---- partition_id = row % K, rows of a partition contiguous (repartition(K, "partition_id"), :295) ----------
    K = int(partition_num) if partition_num else max(1, math.ceil(n / float(sample_size_per_partition)))
    idx = torch.arange(n, device=dev, dtype=torch.int64)
    order = torch.argsort(idx % K, stable=True)
    sizes = torch.full((K,), n // K, dtype=torch.int64)
    sizes[: n % K] += 1
    off = torch.zeros(K + 1, dtype=torch.int64)
    off[1:] = torch.cumsum(sizes, 0)
    return IngestedDesign(numeric=numeric[order].contiguous(), codes=codes[order].contiguous(), n_levels=n_levels,
                          y=y[order].contiguous(), part_offsets=off.to(dev), names=names, numeric_names=numeric_names,
                          dummy_columns=dummy_columns, dummy_info=info, dummy_factors_baseline=dummy_factors_baseline,
                          data_info=data_info, n_rows_read=n_read, level_counts=counts)


def fit(design: IngestedDesign, fit_intercept: bool = True, standardize: bool = True, config=None, return_fits=False):
    """The map phase on an ingested design: the same stacked p x (3+p) frames as
    ``models.logistic_model_partitions`` (== ``groupby("partition_id").apply(logistic_model_udf)``,
    projects/logistic_dlsa.py:305-316), computed by the compact categorical pass when the design has dummy blocks."""
    import pandas as pd
    from .models import _frame
    dev = design.numeric.device
    mean = std = None
    if standardize and design.numeric_names:
        mean = torch.as_tensor(np.array([design.data_info[c][1] for c in design.numeric_names]), device=dev)
        std = torch.as_tensor(np.array([design.data_info[c][2] for c in design.numeric_names]), device=dev)
    n_num = design.numeric.shape[1]
    if design.dummy_columns and engine.cat_supported(n_num, design.n_levels):
        fits = engine.fit_logistic_partitions_cat(design.numeric if n_num else None, design.codes, design.n_levels,
                                                  design.y, design.part_offsets, fit_intercept=fit_intercept, mean=mean,
                                                  std=std, config=config)
    else:
        X = design.numeric if mean is None else (design.numeric - mean) / std
        if design.dummy_columns:   # outside the compact pass's limits: expand on the device
            n = X.shape[0]
            blocks, start = [X], 0
            for ci, L in enumerate(design.n_levels):
                oh = torch.zeros(n, L, dtype=torch.float64, device=dev)
                c = design.codes[:, ci].to(torch.int64)
                ok = c >= 0
                oh[torch.nonzero(ok).squeeze(1), c[ok]] = 1.0
                blocks.append(oh)
            X = torch.cat(blocks, dim=1)
        fits = engine.fit_logistic_partitions(X.contiguous(), design.y, design.part_offsets, fit_intercept=fit_intercept,
                                              config=config)
    names = (["intercept"] if fit_intercept else []) + design.names
    if return_fits:
        return fits, names
    coef, smc, sig = fits.coef.cpu().numpy(), fits.Sig_invMcoef.cpu().numpy(), fits.Sig_inv.cpu().numpy()
    return pd.concat([_frame(coef[k], smc[k], sig[k], names) for k in range(coef.shape[0])], axis=0, ignore_index=True)
