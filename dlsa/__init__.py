"""Compatibility namespace: ``import dlsa`` / ``from dlsa.models import logistic_model`` resolve to the B200 engine
(dlsa_b200) with the reference package's module and function names (feng-li/dlsa: dlsa/models.py, dlsa/dlsa.py,
dlsa/lsa.py, dlsa/dummies.py, dlsa/model_eval.py, dlsa/utils_spark.py)."""
import sys as _sys

from dlsa_b200 import __version__  # noqa: F401
from dlsa_b200 import dlsa as _dlsa_mod, dummies, lsa, model_eval, models

from dlsa_b200 import compat_unsupported as _unsupported

dlsa = _dlsa_mod
utils_spark = dummies  # convert_schema lives with the layout code
sdummies = _unsupported.sdummies   # level-selection rule only; the Spark ML pipeline is out of scope
utils = _unsupported.utils         # insert_partition_id_pdf; the airline CSV cleaner is out of scope (#11)
for _name, _mod in (("models", models), ("dlsa", _dlsa_mod), ("lsa", lsa), ("dummies", dummies),
                    ("model_eval", model_eval), ("utils_spark", dummies), ("sdummies", sdummies), ("utils", utils)):
    _sys.modules[__name__ + "." + _name] = _mod
