"""``dlsa.sdummies`` / ``dlsa.utils`` names of the reference package.  Only what sits on the hot path is provided; the rest
raises a clear NotImplementedError instead of an ImportError, so a caller learns which piece is out of scope
(components 9 and 11) rather than that "the module does not exist"."""
from __future__ import annotations

import types

import numpy as np

from . import dummies as _dummies


def _out_of_scope(name, why):
    def fn(*a, **k):
        raise NotImplementedError(f"dlsa.{name} is outside the hot path this engine replaces: {why}")
    fn.__name__ = name.split(".")[-1]
    return fn


def insert_partition_id_pdf(data_pdf, partition_num, partition_method):
    """dlsa/utils.py:93-107: systematic partition ids ``RangeIndex % partition_num`` as a leading column."""
    if partition_method != "systematic":
        raise Exception("No such partition method implemented!")   # same message as dlsa/models.py:35
    out = data_pdf.copy()
    out.insert(0, "partition_id", np.arange(out.shape[0]) % int(partition_num))
    return out


sdummies = types.ModuleType("dlsa.sdummies")
sdummies.__doc__ = "dlsa/sdummies.py: the level-selection rule is select_sdummy_levels; the Spark ML pipeline is out of scope."
sdummies.select_sdummy_levels = _dummies.select_sdummy_levels
sdummies.get_sdummies = _out_of_scope(
    "sdummies.get_sdummies", "it builds a Spark ML StringIndexer/OneHotEncoder pipeline (dlsa/sdummies.py:95-116); use "
    "dlsa.sdummies.select_sdummy_levels for its level-selection rule and dlsa.models.logistic_model(dummy_info=...) for the fit")

utils = types.ModuleType("dlsa.utils")
utils.__doc__ = "dlsa/utils.py: insert_partition_id_pdf is provided; the airline CSV cleaner is out of scope."
utils.insert_partition_id_pdf = insert_partition_id_pdf
utils.clean_airlinedata = _out_of_scope(
    "utils.clean_airlinedata", "CSV ingest / cleaning of the airline files (dlsa/utils.py:8-90) is not part of the fit -> "
    "combine -> LARS path; hand logistic_model a cleaned frame")
