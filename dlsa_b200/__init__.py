"""dlsa_b200 — B200-native engine for the DLSA hot path (per-partition GLM fit -> one-round combine ->
LARS/DBIC), behind the function names of feng-li/dlsa.  See DESIGN.md and include/dlsa_b200.h."""
__version__ = "0.1.0"

from . import _lib  # noqa: F401  (fails loudly when the CUDA library is not built)
from .engine import (PartitionFits, CombineResult, LarsResult, fit_logistic_partitions, fit_linear_partitions, fit_poisson_partitions,  # noqa: F401
                     combine_partitions, lars_path, dlsa_fit, dlsa_fit_host, logistic_loglik, column_stats, EngineConfig)
