"""Tensor-level engine: drives the C ABI (include/dlsa_b200.h) with PyTorch-owned device buffers.

PyTorch is plumbing here (device memory, streams, torch.distributed); every number is produced by the
hand-written sm_100a kernels in libdlsa_b200.so.  One process per GPU; partitions are sharded across
ranks and the only cross-rank step is one allreduce of the packed sufficient statistic (combine_partitions).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Optional

import torch

from . import _lib, parallel
from ._lib import lib, check

_F64 = torch.float64

# bench.py sets this to a list to collect (start_event, end_event, rows) around every fused pass launch,
# recorded on the stream the kernel is launched on.  None = no timing overhead.
PASS_EVENTS = None


@dataclasses.dataclass
class EngineConfig:
    """Solver settings.  Defaults follow the exact-Newton specification of dlsa/R/wlse_func.R:24-65 with the
    oracle tolerance: start at 0, damp delta*0.01 when max|delta|>5, stop on max|delta|<=tol."""
    tol: float = 1e-10
    max_iter: int = 50
    damp_above: float = 5.0
    damp_factor: float = 0.01
    want_loglik: bool = False   # True: one extra fused pass that also returns every partition's log-likelihood
    # Pooled warm start (the R fitter takes ``init.beta``, dlsa/R/wlse_func.R:24,28): with >= warm_min_partitions
    # partitions, a pilot block of the first max(pilot_min, K*pilot_fraction) partitions is fitted from 0 and the mean of
    # its coefficients starts every other partition.  The MLE each partition converges to is the same (same tol); only
    # the number of passes over X changes (C2: 6 -> 4 + the 1/32-size pilot).  Any partition the warm-started fit does not
    # bring to ST_OK sends the whole batch through the cold start (beta = 0) the reference specifies.
    warm_start: bool = True
    warm_min_partitions: int = 64
    pilot_fraction: float = 1.0 / 32.0
    pilot_min: int = 8
    pilot_tol: float = 1e-3     # the pilot only supplies a starting point: it stops at max|delta| <= max(tol, pilot_tol)


@dataclasses.dataclass
class PartitionFits:
    """Per-partition map output: the tensor form of the p x (3+p) frames of dlsa/models.py:136-142."""
    coef: torch.Tensor          # [K, p]
    Sig_inv: torch.Tensor       # [K, p, p]
    Sig_invMcoef: torch.Tensor  # [K, p]
    status: torch.Tensor        # [K] int32, _lib.ST_*
    iters: torch.Tensor         # [K] int32
    loglik: Optional[torch.Tensor]  # [K] at the returned coef (None if not requested)
    passes: int                 # fused passes launched over ALL partitions (Newton iterations; the last one's H is Sig_inv)
    n_rows: int
    pilot_passes: int = 0       # passes over the pilot block only (pooled warm start), 0 when cold-started
    pilot_partitions: int = 0
    cold_restart: bool = False  # the warm-started fit left a partition unconverged and the batch was refitted from 0
    # deferred validation (dlsa_fit reads every scalar the host needs in ONE device->host copy at the end of the step)
    n_bad: Optional[torch.Tensor] = None        # [1] int32 on the device: partitions whose status is not ST_OK
    cold_refit: Optional[Callable[[], "PartitionFits"]] = None   # set when the fit was warm-started and not yet checked


@dataclasses.dataclass
class CombineResult:
    """dlsa_mapred output (dlsa/dlsa.py:55-59) in tensor form."""
    beta_byOLS: torch.Tensor      # [p]
    beta_byONESHOT: torch.Tensor  # [p]
    Sig_inv_sum: torch.Tensor     # [p, p]
    K_eff: int
    K_total: int


@dataclasses.dataclass
class LarsResult:
    """lars_lsa output (dlsa/lsa.py:211): AIC, BIC, beta, beta0 (+ RSS, dof, argmins)."""
    AIC: torch.Tensor
    BIC: torch.Tensor
    beta: torch.Tensor
    beta0: torch.Tensor
    RSS: torch.Tensor
    dof: torch.Tensor
    argmin_AIC: int
    argmin_BIC: int


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _require_cuda(*ts: torch.Tensor) -> torch.device:
    dev = ts[0].device
    for t in ts:
        if t.device.type != "cuda":
            raise ValueError("dlsa_b200 runs on CUDA tensors only (there is no CPU path)")
        if t.device != dev:
            raise ValueError("all tensors must live on the same device")
    return dev


def _workspace(nbytes: int, dev) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 8), dtype=torch.uint8, device=dev)


def _check_inputs(X, y, part_offsets):
    dev = _require_cuda(X, y, part_offsets)
    if X.dtype != _F64 or y.dtype != _F64:
        raise TypeError("X and y must be float64 (the reference's arithmetic type)")
    if X.dim() != 2 or X.stride(1) != 1:
        raise ValueError("X must be [n, p] row-major (stride(1) == 1)")
    if part_offsets.dtype != torch.int64 or part_offsets.dim() != 1 or part_offsets.numel() < 1:
        raise TypeError("part_offsets must be int64 [K+1]")
    y = y.contiguous().view(-1)
    if y.numel() != X.shape[0]:
        raise ValueError("y and X disagree on the number of rows")
    return dev, y


def _even_pitch(X: torch.Tensor) -> torch.Tensor:
    """A wide design (p > 56: the CTA-tiled pass) with an odd row pitch is copied ONCE into a buffer with an even pitch, so
    that every row starts 16-byte aligned and the pass stages rows with TMA bulk copies instead of 8-byte cp.async
    (4x faster at p = 181); one extra pass over X against the 5+ Newton passes that follow."""
    if X.shape[1] > 56 and (X.stride(0) % 2 == 1 or X.data_ptr() % 16 != 0):
        buf = torch.empty(X.shape[0], X.shape[1] + (X.shape[1] & 1), dtype=X.dtype, device=X.device)
        view = buf[:, : X.shape[1]]
        view.copy_(X)
        return view
    return X


def irls_pass(X, y, part_offsets, beta, fit_intercept=False, family=_lib.FAMILY_LOGISTIC, active=None,
              max_part_rows=None, want_loglik=True, out=None, workspace=None, full=None):
    """One fused pass (K1) over all partitions: returns (g[K,p], H[K,p,p], loglik[K] or None)."""
    dev, y = _check_inputs(X, y, part_offsets)
    K = part_offsets.numel() - 1
    px = X.shape[1]
    p = px + (1 if fit_intercept else 0)
    if max_part_rows is None:
        max_part_rows = int((part_offsets[1:] - part_offsets[:-1]).max().item()) if K > 0 else 0
    if out is None:
        g = torch.zeros(K, p, dtype=_F64, device=dev)
        H = torch.zeros(K, p, p, dtype=_F64, device=dev)
        ll = torch.zeros(K, dtype=_F64, device=dev) if want_loglik else None
    else:
        g, H, ll = out
    need = lib.dlsa_irls_pass_workspace_bytes(K, p, max_part_rows)
    if workspace is None or workspace.numel() < need:
        workspace = _workspace(need, dev)
    if PASS_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.dlsa_irls_pass_f64(X.data_ptr(), X.stride(0), y.data_ptr(), part_offsets.data_ptr(), K, px,
                                 int(fit_intercept), family, _ptr(beta), _ptr(active), g.data_ptr(), H.data_ptr(),
                                 _ptr(ll), max_part_rows, workspace.data_ptr(), workspace.numel(), _stream()),
          "dlsa_irls_pass_f64")
    if PASS_EVENTS is not None:
        e1.record()
        PASS_EVENTS.append((e0, e1, (active is None) if full is None else bool(full)))
    return g, H, ll


def _newton_fit(pass_fn, K: int, p: int, dev, cfg: EngineConfig, n_rows: int, beta_init=None,
                whole_batch: bool = True) -> PartitionFits:
    """The lock-step Newton loop shared by the dense and the categorical map: ``pass_fn(beta, active, out, want_loglik,
    full)`` launches one fused pass (K1 / K1c) into ``out = (g, H, ll)``; K2 updates every active partition.
    ``whole_batch=False`` (pilot block) keeps its passes out of bench.py's full-pass timing."""
    beta = torch.zeros(K, p, dtype=_F64, device=dev) if beta_init is None else beta_init.to(dev, _F64).clone()
    g = torch.zeros(K, p, dtype=_F64, device=dev)
    H = torch.zeros(K, p, p, dtype=_F64, device=dev)
    ll = torch.zeros(K, dtype=_F64, device=dev) if cfg.want_loglik else None
    status = torch.full((K,), _lib.ST_NOT_CONVERGED, dtype=torch.int32, device=dev)
    active = torch.ones(K, dtype=torch.int32, device=dev)
    iters = torch.zeros(K, dtype=torch.int32, device=dev)
    n_active = torch.zeros(1, dtype=torch.int32, device=dev)
    max_delta = torch.zeros(K, dtype=_F64, device=dev)
    ws_chol = _workspace(lib.dlsa_newton_update_workspace_bytes(K, p), dev)
    # The host learns "how many partitions are still active" one iteration late: the count goes to pinned memory with an
    # asynchronous copy and is read only after the NEXT pass has been queued, so the GPU never idles on the host's
    # round trip.  The pass queued after the last partition converged finds every partition inactive and exits at once.
    flags = torch.empty(max(cfg.max_iter, 1), dtype=torch.int32).pin_memory()
    events = []
    done_at = None
    for it in range(cfg.max_iter):
        pass_fn(beta, active, (g, H, None), False, whole_batch and it == 0)
        n_active.zero_()
        check(lib.dlsa_newton_update_f64(H.data_ptr(), g.data_ptr(), K, p, beta.data_ptr(), max_delta.data_ptr(),
                                         status.data_ptr(), active.data_ptr(), iters.data_ptr(), n_active.data_ptr(),
                                         cfg.tol, cfg.damp_above, cfg.damp_factor, ws_chol.data_ptr(),
                                         ws_chol.numel(), _stream()), "dlsa_newton_update_f64")
        flags[it:it + 1].copy_(n_active, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        events.append(ev)
        if it >= 1:
            events[it - 1].synchronize()
            if int(flags[it - 1]) == 0:
                done_at = it - 1
                break
    if done_at is None and events:
        events[-1].synchronize()
        left = int(flags[len(events) - 1])
        passes = len(events)
    else:
        left = 0
        passes = (done_at + 1) if done_at is not None else 0
    # A converged partition's last step (max|delta| <= tol) is not applied by K2, so its H is already Sig_inv at the
    # returned coefficients (dlsa/models.py:114,130): no extra pass.  Only partitions that ran out of iterations get
    # one more pass at the coefficients they return.
    if left > 0:
        pass_fn(beta, active, (g, H, None), False, False)
        passes += 1
    if cfg.want_loglik:  # optional extra: every partition's log-likelihood at its own returned coefficients
        pass_fn(beta, None, (g, H, ll), True, False)
        passes += 1
    Hb = torch.empty(K, p, dtype=_F64, device=dev)
    check(lib.dlsa_batched_symv_f64(H.data_ptr(), beta.data_ptr(), K, p, Hb.data_ptr(), _stream()),
          "dlsa_batched_symv_f64")
    return PartitionFits(coef=beta, Sig_inv=H, Sig_invMcoef=Hb, status=status, iters=iters, loglik=ll, passes=passes,
                         n_rows=n_rows, n_bad=(status != _lib.ST_OK).sum(dtype=torch.int32).view(1))


def _pooled_start(pilot: PartitionFits) -> Optional[torch.Tensor]:
    """Mean coefficient vector of the pilot partitions that converged (None unless at least half of them did)."""
    ok = pilot.status == _lib.ST_OK
    n_ok = int(ok.sum().item())
    if n_ok == 0 or 2 * n_ok < pilot.coef.shape[0]:
        return None
    return (pilot.coef * ok[:, None].to(_F64)).sum(0) / n_ok


def _fit_warm(make_pass, part_offsets, K: int, p: int, dev, cfg: EngineConfig, n_rows: int, beta_init,
              pooled_start=None, defer_check: bool = False) -> PartitionFits:
    """Cold or pooled-warm-start Newton fit (EngineConfig.warm_start).  ``make_pass(offsets)`` returns the pass_fn over
    the partitions ``offsets`` names (a prefix view of ``part_offsets`` for the pilot block: same X, same rows).
    ``pooled_start`` [p]: a pooled coefficient vector from an earlier block of the same job (no pilot needed).
    ``defer_check``: do not read the warm-started fit's status here (a host round trip); the caller reads ``n_bad`` with
    its other scalars and calls ``cold_refit()`` if it is non-zero."""
    pilot_passes = Kp = 0
    warm = False
    if beta_init is None and pooled_start is not None and cfg.warm_start:
        beta_init = pooled_start.to(dev, _F64).view(1, p).expand(K, p).clone()
        warm = True
    elif beta_init is None and cfg.warm_start and K >= max(2, cfg.warm_min_partitions):
        Kp = min(K, max(int(cfg.pilot_min), int(K * cfg.pilot_fraction)))
        pcfg = dataclasses.replace(cfg, want_loglik=False, tol=max(cfg.tol, cfg.pilot_tol))
        pilot = _newton_fit(make_pass(part_offsets[:Kp + 1]), Kp, p, dev, pcfg, 0, None, whole_batch=False)
        pilot_passes = pilot.passes
        start = _pooled_start(pilot)
        if start is not None:
            beta_init = start.expand(K, p).clone()
            okp = (pilot.status == _lib.ST_OK)[:, None]
            beta_init[:Kp] = torch.where(okp, pilot.coef, beta_init[:Kp])  # pilot partitions keep their own MLE
            warm = True
    fits = _newton_fit(make_pass(part_offsets), K, p, dev, cfg, n_rows, beta_init)
    fits.pilot_passes, fits.pilot_partitions = pilot_passes, Kp
    if warm:
        warm_passes = fits.passes

        def cold_refit() -> PartitionFits:
            cold = _newton_fit(make_pass(part_offsets), K, p, dev, cfg, n_rows, None)
            cold.passes += warm_passes
            cold.pilot_passes, cold.pilot_partitions, cold.cold_restart = pilot_passes, Kp, True
            return cold

        if defer_check:
            fits.cold_refit = cold_refit
        elif int(fits.n_bad.item()) != 0:
            fits = cold_refit()
    return fits


def fit_logistic_partitions(X: torch.Tensor, y: torch.Tensor, part_offsets: torch.Tensor, fit_intercept: bool = False,
                            config: Optional[EngineConfig] = None, beta_init: Optional[torch.Tensor] = None,
                            pooled_start: Optional[torch.Tensor] = None, family: int = _lib.FAMILY_LOGISTIC,
                            defer_check: bool = False, max_part_rows: Optional[int] = None) -> PartitionFits:
    """Per-partition logistic MLE for every partition at once (the map side, dlsa/models.py:42-147).

    X [n, px] float64 row-major on the GPU with each partition's rows contiguous; ``part_offsets`` int64 [K+1].
    Newton iterations = K1 (fused pass) + K2 (batched Cholesky update) until every partition has
    max|delta| <= tol.  K2 does not apply a step that small, so the H of a partition's last pass is Sig_inv at exactly
    the coefficients it returns (dlsa/models.py:114,130); K3 then gives Sig_invMcoef (:131).  With an intercept it is coefficient 0 (:116-122).
    """
    cfg = config or EngineConfig()
    dev, y = _check_inputs(X, y, part_offsets)
    X = _even_pitch(X)
    K = part_offsets.numel() - 1
    px = X.shape[1]
    p = px + (1 if fit_intercept else 0)
    if max_part_rows is None:  # one host round trip; callers that know the largest partition pass it in
        sizes = part_offsets[1:] - part_offsets[:-1]
        max_part_rows = int(sizes.max().item()) if K > 0 else 0
    max_rows = int(max_part_rows)

    def make_pass(offsets):
        ws_pass = _workspace(lib.dlsa_irls_pass_workspace_bytes(offsets.numel() - 1, p, max_rows), dev)

        def pass_fn(beta, active, out, want_ll, full):
            irls_pass(X, y, offsets, beta, fit_intercept, family, active, max_rows,
                      want_loglik=want_ll, out=out, workspace=ws_pass, full=full)
        return pass_fn

    return _fit_warm(make_pass, part_offsets, K, p, dev, cfg, int(X.shape[0]), beta_init, pooled_start, defer_check)


def fit_poisson_partitions(X: torch.Tensor, y: torch.Tensor, part_offsets: torch.Tensor, fit_intercept: bool = False,
                           config: Optional[EngineConfig] = None, beta_init: Optional[torch.Tensor] = None
                           ) -> PartitionFits:
    """Per-partition Poisson (log-link) MLE: R ``glm(y ~ x - 1, family = "poisson")`` + ``solve(vcov)`` per block
    (dlsa/R/wlse_func.R:155-168).  Same Newton loop and kernels as the logistic map with the link stage switched:
    mu = exp(eta), g = X'(y - mu), Sig_inv = X' diag(mu) X at the returned coefficients (IRLS with the canonical link
    is exactly this Newton iteration; R starts it from mustart = y + 0.1, here from 0 or the pooled warm start)."""
    return fit_logistic_partitions(X, y, part_offsets, fit_intercept, config, beta_init, None, _lib.FAMILY_POISSON)


def _check_cat_inputs(numeric, codes, n_levels, y, part_offsets):
    dev = _require_cuda(codes, y, part_offsets)
    n_levels = [int(v) for v in n_levels]
    if codes.dtype != torch.int32 or codes.dim() != 2 or not codes.is_contiguous() or codes.shape[1] != len(n_levels):
        raise TypeError("codes must be a contiguous int32 [n, n_cat] tensor with one column per entry of n_levels")
    n = codes.shape[0]
    if numeric is None:
        n_num = 0
    else:
        _require_cuda(numeric, codes)
        if numeric.dtype != _F64 or numeric.dim() != 2 or numeric.shape[0] != n or (numeric.shape[1] and numeric.stride(1) != 1):
            raise TypeError("numeric must be float64 [n, n_num] row-major")
        n_num = numeric.shape[1]
    if y.dtype != _F64 or y.numel() != n:
        raise TypeError("y must be float64 [n]")
    if part_offsets.dtype != torch.int64 or part_offsets.dim() != 1 or part_offsets.numel() < 1:
        raise TypeError("part_offsets must be int64 [K+1]")
    return dev, y.contiguous().view(-1), n_levels, n_num


def irls_pass_cat(numeric, codes, n_levels, y, part_offsets, beta, fit_intercept=False, mean=None, std=None, active=None,
                  max_part_rows=None, want_loglik=True, out=None, workspace=None, full=None):
    """One fused pass over the COMPACT design [numeric | categorical codes] (K1c, ``dlsa_irls_pass_cat_f64``): the
    (g, H, loglik) of ``irls_pass`` on the one-hot expanded matrix, without materialising it.
    ``codes[i, c]`` in [0, n_levels[c]) or -1 for a baseline level; columns = [intercept] numerics, then each
    categorical's levels (dlsa/models.py:71-79)."""
    dev, y, n_levels, n_num = _check_cat_inputs(numeric, codes, n_levels, y, part_offsets)
    K = part_offsets.numel() - 1
    p = int(fit_intercept) + n_num + sum(n_levels)
    if max_part_rows is None:
        max_part_rows = int((part_offsets[1:] - part_offsets[:-1]).max().item()) if K > 0 else 0
    if out is None:
        g = torch.zeros(K, p, dtype=_F64, device=dev)
        H = torch.zeros(K, p, p, dtype=_F64, device=dev)
        ll = torch.zeros(K, dtype=_F64, device=dev) if want_loglik else None
    else:
        g, H, ll = out
    import ctypes as C
    nl = (C.c_int * len(n_levels))(*n_levels)
    need = lib.dlsa_irls_pass_cat_workspace_bytes(K, len(n_levels), nl, max_part_rows)
    if workspace is None or workspace.numel() < need:
        workspace = _workspace(need, dev)
    if PASS_EVENTS is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    check(lib.dlsa_irls_pass_cat_f64(_ptr(numeric) if n_num else None, numeric.stride(0) if n_num else 0, n_num,
                                     _ptr(mean), _ptr(std), codes.data_ptr(), len(n_levels), nl, y.data_ptr(),
                                     part_offsets.data_ptr(), K, int(fit_intercept), beta.data_ptr(), _ptr(active),
                                     g.data_ptr(), H.data_ptr(), _ptr(ll), max_part_rows, workspace.data_ptr(),
                                     workspace.numel(), _stream()), "dlsa_irls_pass_cat_f64")
    if PASS_EVENTS is not None:
        e1.record()
        PASS_EVENTS.append((e0, e1, (active is None) if full is None else bool(full)))
    return g, H, ll


def cat_supported(n_num: int, n_levels) -> bool:
    """Shape limits of the compact pass (include/dlsa_b200.h, K1c); outside them expand and use the dense pass."""
    n_levels = list(n_levels)
    return n_num <= 7 and 1 <= len(n_levels) <= 8 and sum(n_levels) <= 192 and min(n_levels) >= 1


def fit_logistic_partitions_cat(numeric, codes, n_levels, y, part_offsets, fit_intercept: bool = False, mean=None,
                                std=None, config: Optional[EngineConfig] = None,
                                beta_init: Optional[torch.Tensor] = None, pooled_start: Optional[torch.Tensor] = None,
                                defer_check: bool = False, max_part_rows: Optional[int] = None) -> PartitionFits:
    """``fit_logistic_partitions`` for a design with dummy blocks kept compact: numeric [n, n_num] float64 (optionally
    standardised on load with mean/std, dlsa/models.py:99-101) and codes [n, n_cat] int32.  Same outputs, same column
    order as fitting the one-hot expanded matrix (the dummy branch of logistic_model, dlsa/models.py:56-91)."""
    cfg = config or EngineConfig()
    dev, y, n_levels, n_num = _check_cat_inputs(numeric, codes, n_levels, y, part_offsets)
    K = part_offsets.numel() - 1
    p = int(fit_intercept) + n_num + sum(n_levels)
    if max_part_rows is None:
        sizes = part_offsets[1:] - part_offsets[:-1]
        max_part_rows = int(sizes.max().item()) if K > 0 else 0
    max_rows = int(max_part_rows)
    import ctypes as C
    nl = (C.c_int * len(n_levels))(*n_levels)
    if mean is not None:
        mean = mean.to(dev, _F64).contiguous()
        std = std.to(dev, _F64).contiguous()

    def make_pass(offsets):
        ws_pass = _workspace(lib.dlsa_irls_pass_cat_workspace_bytes(offsets.numel() - 1, len(n_levels), nl, max_rows),
                             dev)

        def pass_fn(beta, active, out, want_ll, full):
            irls_pass_cat(numeric, codes, n_levels, y, offsets, beta, fit_intercept, mean, std, active, max_rows,
                          want_loglik=want_ll, out=out, workspace=ws_pass, full=full)
        return pass_fn

    return _fit_warm(make_pass, part_offsets, K, p, dev, cfg, int(codes.shape[0]), beta_init, pooled_start, defer_check)


def fit_linear_partitions(X: torch.Tensor, y: torch.Tensor, part_offsets: torch.Tensor,
                          max_part_rows: Optional[int] = None, events=None) -> PartitionFits:
    """Per-partition OLS in closed form (R ``lm(y~x-1)`` + ``solve(vcov)``, dlsa/R/wlse_func.R:169-182):
    one fused pass gives X'X, X'y, y'y (K8); theta_k by batched Cholesky; Sig_inv_k = X'X / sigma2_k.
    ``events``: optional (event,) recorded right after the fused pass (bench.py's per-kernel bracket)."""
    dev, y = _check_inputs(X, y, part_offsets)
    X = _even_pitch(X)
    K = part_offsets.numel() - 1
    p = X.shape[1]
    if max_part_rows is None:
        sizes = part_offsets[1:] - part_offsets[:-1]
        max_part_rows = int(sizes.max().item()) if K > 0 else 0
    max_rows = int(max_part_rows)
    g, H, yty = irls_pass(X, y, part_offsets, None, False, _lib.FAMILY_LINEAR, None, max_rows, want_loglik=True)
    if events:
        events[0].record()
    theta = torch.empty(K, p, dtype=_F64, device=dev)
    status = torch.zeros(K, dtype=torch.int32, device=dev)
    ws = _workspace(lib.dlsa_newton_update_workspace_bytes(K, p), dev)
    check(lib.dlsa_spd_solve_f64(H.data_ptr(), g.data_ptr(), K, p, theta.data_ptr(), status.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _stream()), "dlsa_spd_solve_f64")
    sigma2 = torch.empty(K, dtype=_F64, device=dev)
    check(lib.dlsa_ols_scale_f64(H.data_ptr(), g.data_ptr(), yty.data_ptr(), theta.data_ptr(),
                                 part_offsets.data_ptr(), K, p, sigma2.data_ptr(), _stream()), "dlsa_ols_scale_f64")
    return PartitionFits(coef=theta, Sig_inv=H, Sig_invMcoef=g, status=status,
                         iters=torch.ones(K, dtype=torch.int32, device=dev), loglik=sigma2, passes=1,
                         n_rows=int(X.shape[0]), n_bad=(status != _lib.ST_OK).sum(dtype=torch.int32).view(1))


def reduce_partitions(fits: PartitionFits) -> torch.Tensor:
    """K4: fixed-order sum over this rank's partitions into the packed buffer
    [sum Sig_inv (p*p) | sum Sig_inv coef (p) | sum coef (p) | K_eff | K] — the allreduce payload (parallel.py)."""
    K, p = fits.coef.shape
    packed = torch.empty(parallel.packed_len(p), dtype=_F64, device=fits.coef.device)
    check(lib.dlsa_partition_reduce_f64(fits.Sig_inv.data_ptr(), fits.Sig_invMcoef.data_ptr(), fits.coef.data_ptr(),
                                        fits.status.data_ptr(), K, p, packed.data_ptr(), _stream()),
          "dlsa_partition_reduce_f64")
    return packed


class _PendingCombine:
    """K5 queued on the stream; ``meta`` [3] float64 on the device = (K_eff, K_total, solve status), read by finish()."""

    def __init__(self, packed, p, theta, st, num_partitions):
        self.packed, self.p, self.theta, self.num_partitions = packed, p, theta, num_partitions
        self.meta = torch.cat([packed[p * p + 2 * p:], st.to(_F64)])

    def finish(self, meta_host) -> CombineResult:
        K_eff, K_total, st = int(round(meta_host[0])), int(round(meta_host[1])), int(round(meta_host[2]))
        if K_eff == 0:
            raise Exception("Zero-length grouped pandas DataFrame obtained, check the input.")  # dlsa/dlsa.py:36-39
        p = self.p
        S = self.packed[: p * p].view(p, p)
        v = self.packed[p * p: p * p + p]
        c = self.packed[p * p + p: p * p + 2 * p]
        theta = self.theta
        if st != _lib.ST_OK:
            # dlsa/dlsa.py:48 is np.linalg.lstsq: the minimum-norm solution when the sum is singular.  Rare (collinear
            # dummies, an empty level), off the hot path: SVD pseudo-inverse of the p x p sum through torch.linalg.
            U, sv, Vh = torch.linalg.svd(S)
            cut = torch.finfo(_F64).eps * p * sv.max()   # numpy's default rcond = eps * max(M, N)
            inv = torch.where(sv > cut, 1.0 / sv, torch.zeros_like(sv))
            theta = Vh.T @ (inv * (U.T @ v))
        Kd = K_total if self.num_partitions is None else self.num_partitions
        return CombineResult(beta_byOLS=theta, beta_byONESHOT=c / Kd, Sig_inv_sum=S.clone(), K_eff=K_eff,
                             K_total=K_total)


def _queue_combine(packed: torch.Tensor, p: int, group=None, num_partitions: Optional[int] = None,
                   allreduce: bool = True) -> _PendingCombine:
    dev = packed.device
    if allreduce:
        parallel.allreduce_packed(packed, group)  # the one collective of the whole algorithm
    theta = torch.empty(p, dtype=_F64, device=dev)
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    ws = _workspace(lib.dlsa_newton_update_workspace_bytes(1, p), dev)
    check(lib.dlsa_spd_solve_f64(packed.data_ptr(), packed[p * p:].data_ptr(), 1, p, theta.data_ptr(), st.data_ptr(),
                                 ws.data_ptr(), ws.numel(), _stream()), "dlsa_spd_solve_f64")
    return _PendingCombine(packed, p, theta, st, num_partitions)


def combine_packed(packed: torch.Tensor, p: int, group=None, num_partitions: Optional[int] = None,
                   allreduce: bool = True) -> CombineResult:
    """ONE allreduce of the packed buffer when torch.distributed is initialised, then K5 (Cholesky solve of the SPD
    sum; the reference's lstsq, dlsa/dlsa.py:48, whose minimum-norm answer is reproduced when the sum is singular)."""
    pend = _queue_combine(packed, p, group, num_partitions, allreduce)
    return pend.finish(pend.meta.tolist())


def combine_partitions(fits: PartitionFits, group=None, num_partitions: Optional[int] = None) -> CombineResult:
    """One-round combine (dlsa_mapred, dlsa/dlsa.py:21-61): K4 fixed-order local sum, ONE allreduce of the packed
    [sum Sig_inv | sum Sig_inv coef | sum coef | K_eff | K] buffer when torch.distributed is initialised, then K5.
    beta_byONESHOT = sum coef / K (:51-52), with K the total partition count (skipped partitions count, as
    getNumPartitions() does)."""
    return combine_packed(reduce_partitions(fits), fits.coef.shape[1], group, num_partitions)


class _PendingLars:
    """K6/K7 queued on the stream; ``info`` [4] int32 on the device = (rows, argmin AIC, argmin BIC, error)."""

    def __init__(self, bufs, info):
        self.bufs, self.info = bufs, info

    def finish(self, info_host) -> LarsResult:
        nrows, ia, ib, err = (int(round(v)) for v in info_host)
        if err:
            raise _lib.DlsaLibraryError("lars path: no variable could enter (singular Sigma)")
        beta, beta0, rss, aic, bic, dof = self.bufs
        return LarsResult(AIC=aic[:nrows], BIC=bic[:nrows], beta=beta[:nrows], beta0=beta0[:nrows], RSS=rss[:nrows],
                          dof=dof[:nrows], argmin_AIC=ia, argmin_BIC=ib)


def _queue_lars(Sigma0: torch.Tensor, b0: torch.Tensor, intercept: bool, n: float, type: str = "lar",
                eps: float = 2.220446049250313e-16, max_steps: Optional[int] = None) -> _PendingLars:
    dev = _require_cuda(Sigma0, b0)
    Sigma0 = Sigma0.to(_F64).contiguous()
    b0 = b0.to(_F64).contiguous().view(-1)
    p0 = b0.numel()
    if Sigma0.shape != (p0, p0):
        raise ValueError("Sigma0 must be [p, p] with p = len(b0)")
    if type not in ("lar", "lasso"):
        raise ValueError("type must be 'lar' or 'lasso'")
    m = p0 - (1 if intercept else 0)
    steps = 8 * m if (max_steps is None or int(max_steps) < 0) else int(max_steps)  # 0: the zero row only (dlsa/lsa.py:126)
    rows = steps + 1
    beta = torch.zeros(rows, m, dtype=_F64, device=dev)
    beta0 = torch.zeros(rows, dtype=_F64, device=dev)
    rss = torch.zeros(rows, dtype=_F64, device=dev)
    aic = torch.zeros(rows, dtype=_F64, device=dev)
    bic = torch.zeros(rows, dtype=_F64, device=dev)
    dof = torch.zeros(rows, dtype=torch.int32, device=dev)
    info = torch.zeros(4, dtype=torch.int32, device=dev)
    ws = _workspace(lib.dlsa_lars_path_workspace_bytes(p0, steps), dev)
    check(lib.dlsa_lars_path_f64(Sigma0.data_ptr(), b0.data_ptr(), p0, int(bool(intercept)), float(n),
                                 _lib.LARS_LAR if type == "lar" else _lib.LARS_LASSO, float(eps), steps,
                                 beta.data_ptr(), beta0.data_ptr(), rss.data_ptr(), aic.data_ptr(), bic.data_ptr(),
                                 dof.data_ptr(), info.data_ptr(), ws.data_ptr(), ws.numel(), _stream()),
          "dlsa_lars_path_f64")
    return _PendingLars((beta, beta0, rss, aic, bic, dof), info)


def lars_path(Sigma0: torch.Tensor, b0: torch.Tensor, intercept: bool, n: float, type: str = "lar",
              eps: float = 2.220446049250313e-16, max_steps: Optional[int] = None) -> LarsResult:
    """lars_lsa on the device (K6 + K7): dlsa/lsa.py:90-212; intercept branch per the R original."""
    pend = _queue_lars(Sigma0, b0, intercept, n, type, eps, max_steps)
    return pend.finish(pend.info.tolist())


def logistic_loglik(X, y, part_offsets, par, fit_intercept=False) -> torch.Tensor:
    """Per-partition log-likelihood of each column of ``par`` [p, q] (logistic_model_eval, dlsa/models.py:151-225).
    Returns [K, q]; the global value is the column sum (dlsa/model_eval.py:38-39)."""
    dev, y = _check_inputs(X, y, part_offsets)
    K = part_offsets.numel() - 1
    par = par.to(dev, _F64).contiguous()
    q = par.shape[1]
    out = torch.zeros(K, q, dtype=_F64, device=dev)
    ws = _workspace(lib.dlsa_logistic_loglik_workspace_bytes(K, q), dev)
    check(lib.dlsa_logistic_loglik_f64(X.data_ptr(), X.stride(0), y.data_ptr(), part_offsets.data_ptr(), K, X.shape[1],
                                       int(fit_intercept), par.data_ptr(), q, out.data_ptr(), ws.data_ptr(), ws.numel(),
                                       _stream()), "dlsa_logistic_loglik_f64")
    return out


def column_stats(X: torch.Tensor) -> torch.Tensor:
    """describe() on the device: [5, px] = count, mean, sample stddev, min, max of every column of X [n, px] in one
    streaming pass (Spark's ``data_sdf_i.describe()``, projects/logistic_dlsa.py:279-283; rows 1 and 2 are what
    logistic_model standardises with, dlsa/models.py:99-101)."""
    dev = _require_cuda(X)
    if X.dtype != _F64 or X.dim() != 2 or X.stride(1) != 1 or X.shape[0] < 1 or X.shape[1] < 1:
        raise TypeError("X must be a non-empty float64 [n, px] row-major tensor")
    n, px = X.shape
    stats = torch.empty(5, px, dtype=_F64, device=dev)
    ws = _workspace(lib.dlsa_column_stats_workspace_bytes(px), dev)
    check(lib.dlsa_column_stats_f64(X.data_ptr(), X.stride(0), n, px, stats.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _stream()), "dlsa_column_stats_f64")
    return stats


def _global_rows(n_local: int, dev, group=None) -> float:
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        n = torch.tensor([float(n_local)], dtype=_F64, device=dev)
        dist.all_reduce(n, group=group)
        return float(n.item())
    return float(n_local)


def finish_fit(fits: PartitionFits, fit_intercept: bool, sample_size: float, type: str = "lar", group=None,
               num_partitions: Optional[int] = None, packed: Optional[torch.Tensor] = None):
    """Reduce side of the hot path after the map: K4 -> ONE allreduce -> K5 -> K6/K7, all queued on the stream without
    a host round trip; every scalar the host needs (warm-start failures, K_eff, solve status, path length, argmins) comes
    back in ONE device->host copy.  ``fits.cold_refit`` (set by a deferred warm-started fit) is honoured: if any
    partition did not reach ST_OK the batch is refitted from beta = 0 and the reduce side runs again.  ``packed``: a
    pre-reduced local buffer (chunked maps accumulate it themselves).  Returns (fits, comb, path)."""
    p = fits.coef.shape[1]
    local = reduce_partitions(fits) if packed is None else packed
    pc = _queue_combine(local, p, group, num_partitions)
    pl = _queue_lars(local[: p * p].view(p, p), pc.theta, fit_intercept, sample_size, type=type)
    n_bad = fits.n_bad if fits.n_bad is not None else torch.zeros(1, dtype=torch.int32, device=local.device)
    meta = torch.cat([n_bad.to(_F64), pc.meta, pl.info.to(_F64)]).tolist()   # the step's one host round trip
    if fits.cold_refit is not None and meta[0] != 0 and packed is None:
        cold = fits.cold_refit()
        return finish_fit(cold, fit_intercept, sample_size, type, group, num_partitions)
    fits.cold_refit = None
    comb = pc.finish(meta[1:4])
    if int(round(meta[3])) != _lib.ST_OK:   # singular sum: the path must follow the lstsq estimate, not the failed solve
        path = lars_path(comb.Sig_inv_sum, comb.beta_byOLS, fit_intercept, sample_size, type=type)
    else:
        path = pl.finish(meta[4:8])
    return fits, comb, path


def dlsa_fit(X, y, part_offsets, fit_intercept=False, sample_size=None, type="lar", config=None, group=None,
             family="logistic", max_part_rows=None):
    """Whole hot path on this rank's partitions: map (fit) -> combine (+allreduce) -> LARS + AIC/BIC.
    ``sample_size`` is the global n used by BIC (defaults to the all-rank row count).  After the Newton loop (whose
    convergence counter the host reads one iteration late) the step makes ONE host round trip (``finish_fit``)."""
    if sample_size is None:
        sample_size = _global_rows(int(X.shape[0]), X.device, group)
    if family == "logistic":
        fits = fit_logistic_partitions(X, y, part_offsets, fit_intercept, config, defer_check=True,
                                       max_part_rows=max_part_rows)
    elif family == "linear":
        fits = fit_linear_partitions(X, y, part_offsets)
    elif family == "poisson":
        fits = fit_logistic_partitions(X, y, part_offsets, fit_intercept, config, None, None, _lib.FAMILY_POISSON,
                                       defer_check=True, max_part_rows=max_part_rows)
    else:
        raise ValueError("family must be 'logistic', 'linear' or 'poisson'")
    return finish_fit(fits, fit_intercept, float(sample_size), type, group)


# ---- pageable host memory: staged copies --------------------------------------------------------------------------------
# cudaMemcpyAsync from PAGEABLE memory goes through the driver's own bounce buffer, single-threaded (measured on the bench
# box: 11 GB/s against 55 GB/s from pinned memory).  A caller that holds numpy / pandas data cannot pin 40 GB in place, so
# the host entries stage such inputs themselves: a few pinned slots, filled by a pool of host threads (torch's CPU copy
# releases the GIL) while the previous slots cross PCIe.
_STAGE_SLOTS = {}
_STAGE_SLOT_BYTES = 256 << 20
_STAGE_N_SLOTS = 4


def _stage_slots(dev):
    key = (dev.index, _STAGE_SLOT_BYTES, _STAGE_N_SLOTS)
    if key not in _STAGE_SLOTS:
        _STAGE_SLOTS[key] = [torch.empty(_STAGE_SLOT_BYTES, dtype=torch.uint8).pin_memory() for _ in range(_STAGE_N_SLOTS)]
    return _STAGE_SLOTS[key]


class _StagedCopies:
    """Queues ``dst[r0:r1] <- src[r0:r1]`` for every (dst, src) pair and every row block of ``bounds`` on ``copy_stream``
    from a background thread; ``wait(i, compute_stream)`` makes the compute stream wait for block i."""

    def __init__(self, pairs, bounds, dev, copy_stream, workers=None):
        import concurrent.futures
        import os
        import threading
        self.pairs, self.bounds, self.dev, self.copy = pairs, bounds, dev, copy_stream
        self.ready = [threading.Event() for _ in bounds]
        self.events = [None] * len(bounds)
        self.error = None
        self.workers = workers or max(1, min(16, (os.cpu_count() or 2) - 1))
        self.pool = concurrent.futures.ThreadPoolExecutor(max_workers=self.workers)
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _fill(self, view, src):
        n = src.shape[0]
        step = max(1, -(-n // self.workers))
        futs = [self.pool.submit(lambda a, b: view[a:b].copy_(src[a:b]), a, min(n, a + step)) for a in range(0, n, step)]
        for f in futs:
            f.result()

    def _run(self):
        try:
            torch.cuda.set_device(self.dev)
            slots = _stage_slots(self.dev)
            free = [None] * len(slots)   # the event after which a slot may be overwritten
            turn = 0
            with torch.cuda.stream(self.copy):
                for i, (r0, r1) in enumerate(self.bounds):
                    for dst, src in self.pairs:
                        row_bytes = max(1, src[0:1].numel() * src.element_size())
                        rows_per_slot = max(1, _STAGE_SLOT_BYTES // row_bytes)
                        for a in range(r0, r1, rows_per_slot):
                            b = min(r1, a + rows_per_slot)
                            sl = turn % len(slots)
                            turn += 1
                            if free[sl] is not None:
                                free[sl].synchronize()
                            piece = src[a:b]
                            view = slots[sl][: piece.numel() * piece.element_size()].view(piece.dtype).view(piece.shape)
                            self._fill(view, piece)
                            dst[a:b].copy_(view, non_blocking=True)
                            ev = torch.cuda.Event()
                            ev.record(self.copy)
                            free[sl] = ev
                    ev = torch.cuda.Event()
                    ev.record(self.copy)
                    self.events[i] = ev
                    self.ready[i].set()
        except BaseException as exc:  # surfaced by wait()
            self.error = exc
            for r in self.ready:
                r.set()

    def wait(self, i, compute_stream):
        self.ready[i].wait()
        if self.error is not None:
            raise self.error
        compute_stream.wait_event(self.events[i])

    def close(self):
        self.thread.join()
        self.pool.shutdown(wait=True)
        if self.error is not None:
            raise self.error


def _queue_host_copies(pairs, bounds, dev, compute):
    """Host -> device copies of the row blocks ``bounds`` for every (dst, src) pair, on their own stream.  Returns an
    object with ``wait(i, stream)`` and ``close()``.  Pinned sources are queued directly; pageable ones are staged."""
    copy = torch.cuda.Stream(dev)
    copy.wait_stream(compute)  # the buffers may still be in use by earlier work on the compute stream
    if all(src.is_pinned() for _, src in pairs):
        class _Direct:
            def __init__(self):
                self.events = []
                with torch.cuda.stream(copy):
                    for r0, r1 in bounds:
                        for dst, src in pairs:
                            dst[r0:r1].copy_(src[r0:r1], non_blocking=True)
                        ev = torch.cuda.Event()
                        ev.record(copy)
                        self.events.append(ev)

            def wait(self, i, stream):
                stream.wait_event(self.events[i])

            def close(self):
                pass
        return _Direct()
    return _StagedCopies(pairs, bounds, dev, copy)



def dlsa_fit_host(Xh: torch.Tensor, yh: torch.Tensor, part_offsets_host: torch.Tensor, fit_intercept=False,
                  sample_size=None, type="lar", config=None, group=None, device=None, n_groups: int = 16,
                  buffers=None):
    """The whole hot path for HOST-resident partitions (what a Spark worker hands over: rows in host memory).

    X [n, px] / y [n] float64 host tensors (pinned memory makes the copies asynchronous), ``part_offsets_host`` int64
    [K+1] on the host.  Partitions are independent until the one-round combine, so they are cut into ``n_groups``
    contiguous blocks: every block's host->device copy is queued up front on a copy stream, and block i is fitted on
    the compute stream while blocks i+1.. are still crossing PCIe.  Wall time ~ copy time + one block's fit.
    ``buffers`` = optional (X_dev, y_dev) to reuse.  Returns (fits, comb, path) like ``dlsa_fit``."""
    if Xh.device.type != "cpu" or yh.device.type != "cpu" or part_offsets_host.device.type != "cpu":
        raise ValueError("dlsa_fit_host takes host tensors; use dlsa_fit for device-resident data")
    if Xh.dtype != _F64 or yh.dtype != _F64 or part_offsets_host.dtype != torch.int64:
        raise TypeError("X, y must be float64 and part_offsets int64")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n, px = Xh.shape
    K = part_offsets_host.numel() - 1
    offs = part_offsets_host.tolist()
    if buffers is None:
        Xd = torch.empty(n, px, dtype=_F64, device=dev)
        yd = torch.empty(n, dtype=_F64, device=dev)
    else:
        Xd, yd = buffers
    G = max(1, min(int(n_groups), K))
    bounds = [(K * i) // G for i in range(G + 1)]
    compute = torch.cuda.current_stream(dev)
    copies = _queue_host_copies([(Xd, Xh), (yd, yh)], [(offs[bounds[i]], offs[bounds[i + 1]]) for i in range(G)], dev, compute)
    parts = []
    pooled = None  # mean coefficients of the first fitted block: the warm start of every later block
    try:
        for i in range(G):
            k0, k1 = bounds[i], bounds[i + 1]
            r0, r1 = offs[k0], offs[k1]
            off_g = (part_offsets_host[k0:k1 + 1] - r0).to(dev, non_blocking=True)
            copies.wait(i, compute)
            if k1 > k0:
                parts.append(fit_logistic_partitions(Xd[r0:r1], yd[r0:r1], off_g, fit_intercept, config,
                                                     pooled_start=pooled))
                if pooled is None and K >= (config or EngineConfig()).warm_min_partitions:
                    pooled = _pooled_start(parts[-1])
    finally:
        copies.close()
    fits = PartitionFits(coef=torch.cat([f.coef for f in parts]), Sig_inv=torch.cat([f.Sig_inv for f in parts]),
                         Sig_invMcoef=torch.cat([f.Sig_invMcoef for f in parts]),
                         status=torch.cat([f.status for f in parts]), iters=torch.cat([f.iters for f in parts]),
                         loglik=None if parts[0].loglik is None else torch.cat([f.loglik for f in parts]),
                         passes=max(f.passes for f in parts), n_rows=int(n),
                         pilot_passes=parts[0].pilot_passes, pilot_partitions=parts[0].pilot_partitions,
                         cold_restart=any(f.cold_restart for f in parts),
                         n_bad=torch.stack([f.n_bad for f in parts]).sum(0).to(torch.int32))
    if sample_size is None:
        sample_size = _global_rows(int(n), dev, group)
    return finish_fit(fits, fit_intercept, float(sample_size), type, group)


def dlsa_fit_host_cat(numeric_h: torch.Tensor, codes_h: torch.Tensor, n_levels, yh: torch.Tensor,
                      part_offsets_host: torch.Tensor, fit_intercept=True, mean=None, std=None, sample_size=None,
                      type="lar", config=None, group=None, device=None, n_groups: int = 8, buffers=None):
    """``dlsa_fit_host`` for the COMPACT categorical design: host-resident numeric [n, n_num] float64, codes [n, n_cat]
    int32 (level index inside its block, -1 = baseline) and y [n] — 84 bytes per airline-shaped row instead of the 1456 of
    the one-hot matrix, so the host->device copy shrinks 17x (what a Spark worker would hand over after StringIndexer,
    dlsa/sdummies.py:95-116, or after dummies.select_dummy_factors).  Partition blocks are copied on a copy stream while
    earlier blocks are fitted by the categorical fused pass K1c.  Returns (fits, comb, path)."""
    for t in (numeric_h, codes_h, yh, part_offsets_host):
        if t.device.type != "cpu":
            raise ValueError("dlsa_fit_host_cat takes host tensors; use fit_logistic_partitions_cat for device data")
    if numeric_h.dtype != _F64 or yh.dtype != _F64 or codes_h.dtype != torch.int32 or part_offsets_host.dtype != torch.int64:
        raise TypeError("numeric, y must be float64, codes int32 and part_offsets int64")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    n = yh.numel()
    K = part_offsets_host.numel() - 1
    offs = part_offsets_host.tolist()
    if buffers is None:
        nd = torch.empty(numeric_h.shape, dtype=_F64, device=dev)
        cd = torch.empty(codes_h.shape, dtype=torch.int32, device=dev)
        yd = torch.empty(n, dtype=_F64, device=dev)
    else:
        nd, cd, yd = buffers
    G = max(1, min(int(n_groups), K))
    bounds = [(K * i) // G for i in range(G + 1)]
    compute = torch.cuda.current_stream(dev)
    copies = _queue_host_copies([(nd, numeric_h), (cd, codes_h), (yd, yh)],
                                [(offs[bounds[i]], offs[bounds[i + 1]]) for i in range(G)], dev, compute)
    parts = []
    pooled = None
    try:
        for i in range(G):
            k0, k1 = bounds[i], bounds[i + 1]
            r0, r1 = offs[k0], offs[k1]
            copies.wait(i, compute)
            if k1 == k0:
                continue
            sizes_h = part_offsets_host[k0 + 1:k1 + 1] - part_offsets_host[k0:k1]
            off_g = (part_offsets_host[k0:k1 + 1] - r0).to(dev, non_blocking=True)
            parts.append(fit_logistic_partitions_cat(nd[r0:r1], cd[r0:r1], n_levels, yd[r0:r1], off_g, fit_intercept, mean,
                                                     std, config, pooled_start=pooled, max_part_rows=int(sizes_h.max())))
            if pooled is None and len(parts) == 1 and G > 1:
                pooled = _pooled_start(parts[0])
    finally:
        copies.close()
    fits = PartitionFits(coef=torch.cat([f.coef for f in parts]), Sig_inv=torch.cat([f.Sig_inv for f in parts]),
                         Sig_invMcoef=torch.cat([f.Sig_invMcoef for f in parts]),
                         status=torch.cat([f.status for f in parts]), iters=torch.cat([f.iters for f in parts]),
                         loglik=None, passes=max(f.passes for f in parts), n_rows=int(n),
                         pilot_passes=parts[0].pilot_passes, pilot_partitions=parts[0].pilot_partitions,
                         cold_restart=any(f.cold_restart for f in parts),
                         n_bad=torch.stack([f.n_bad for f in parts]).sum(0).to(torch.int32))
    if sample_size is None:
        sample_size = _global_rows(int(n), dev, group)
    return finish_fit(fits, fit_intercept, float(sample_size), type, group)
