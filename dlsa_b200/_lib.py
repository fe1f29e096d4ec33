"""ctypes binding of libdlsa_b200.so (the C ABI declared in include/dlsa_b200.h).

The library is the product: there is no CPU or PyTorch fallback.  Importing this module on a machine
where the shared object has not been built raises immediately (run ``python -c "import
__graft_entry__ as g; g.build()"`` or ``make -C dlsa_b200/csrc``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# DLSA_B200_LIB: tooling only (tools/k1_variants.sh times experimental builds of the same ABI side by side)
LIB_PATH = os.environ.get("DLSA_B200_LIB") or os.path.join(_HERE, "libdlsa_b200.so")

ABI_VERSION = 1
ST_OK, ST_NOT_CONVERGED, ST_NOT_SPD, ST_SKIPPED = 0, 1, 2, 4
FAMILY_LOGISTIC, FAMILY_LINEAR, FAMILY_POISSON = 0, 1, 2
LARS_LAR, LARS_LASSO = 0, 1


class DlsaLibraryError(RuntimeError):
    pass


if not os.path.isfile(LIB_PATH):
    raise DlsaLibraryError(
        f"{LIB_PATH} is missing: the CUDA library is the only implementation of the DLSA hot path "
        "(no CPU fallback). Build it with `make -C dlsa_b200/csrc` (nvcc, sm_100a).")

lib = C.CDLL(LIB_PATH)

_vp, _i, _i64, _d, _sz, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t, C.c_uint64

# name -> (restype, argtypes); mirrors include/dlsa_b200.h one to one
SIGNATURES = {
    "dlsa_abi_version": (_i, []),
    "dlsa_last_error": (C.c_char_p, []),
    "dlsa_launch_count": (C.c_ulonglong, []),
    "dlsa_irls_pass_workspace_bytes": (_sz, [_i, _i, _i64]),
    "dlsa_irls_pass_f64": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64, _vp, _sz, _vp]),
    "dlsa_irls_pass_cat_workspace_bytes": (_sz, [_i, _i, _vp, _i64]),
    "dlsa_irls_pass_cat_f64": (_i, [_vp, _i64, _i, _vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i64,
                                    _vp, _sz, _vp]),
    "dlsa_newton_update_workspace_bytes": (_sz, [_i, _i]),
    "dlsa_newton_update_f64": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _d, _d, _d, _vp, _sz, _vp]),
    "dlsa_spd_solve_f64": (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp, _sz, _vp]),
    "dlsa_batched_symv_f64": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "dlsa_ols_scale_f64": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "dlsa_partition_reduce_f64": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp]),
    "dlsa_lars_path_workspace_bytes": (_sz, [_i, _i]),
    "dlsa_lars_path_f64": (_i, [_vp, _vp, _i, _i, _d, _i, _d, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "dlsa_logistic_loglik_workspace_bytes": (_sz, [_i, _i]),
    "dlsa_logistic_loglik_f64": (_i, [_vp, _i64, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "dlsa_simulate_f64": (_i, [_vp, _i64, _vp, _i64, _i64, _i, _vp, _u64, _i, _vp]),
    "dlsa_simulate_airline_f64": (_i, [_vp, _i64, _vp, _vp, _i64, _i64, _i, _i, _vp, _vp, _d, _vp, _vp, _u64, _vp]),
    "dlsa_probe_fp64_dmma": (_i, [_i, _i, _vp, _vp, _vp]),
    "dlsa_onehot_expand_f64": (_i, [_vp, _i, _vp, _vp, _vp, _i, _vp, _i64, _vp, _i64, _vp]),
    "dlsa_irls_tiled_schedule": (_i, [_i, _i, _vp, _vp]),
    "dlsa_irls_cat_plan": (_i, [_i, _i, _i, _vp, _i64, _vp, _vp, _vp]),
    "dlsa_column_stats_workspace_bytes": (_sz, [_i]),
    "dlsa_column_stats_f64": (_i, [_vp, _i64, _i64, _i, _vp, _vp, _sz, _vp]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here = header and library disagree
    _fn.restype = _res
    _fn.argtypes = _args

if lib.dlsa_abi_version() != ABI_VERSION:
    raise DlsaLibraryError(f"libdlsa_b200.so ABI {lib.dlsa_abi_version()} != binding {ABI_VERSION}; rebuild")


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib.dlsa_last_error().decode("utf-8", "replace")
        kind = "argument error" if rc < 0 else "CUDA error"
        raise DlsaLibraryError(f"{what}: {kind} {rc}: {msg}")


def launch_count() -> int:
    return int(lib.dlsa_launch_count())
