"""TEST INFRASTRUCTURE ONLY — CPU (NumPy float64) restatement of the DLSA hot path.

This file is the *oracle*: the checker the CUDA path is compared with.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs may import it.
The product (``dlsa_b200``) never does; it fails loudly when its CUDA library is missing.

Pinning (see ``oracle/make_golden.py`` and ``tests/test_oracle_pinned.py``): every function below is
checked against the reference's *own* functions run under the compat shims of ``oracle/ref_shim.py``
(in the build container, where ``/root/reference`` exists), and against the frozen outputs of those
runs in ``tests/golden/*.npz``.  The reference ships no tests of its own, so those
frozen reference outputs are the golden vectors.

Citations are relative to the reference tree (``/root/reference``).
"""
from __future__ import annotations

import math
import warnings

import numpy as np

EPS = float(np.finfo(np.float64).eps)

# --------------------------------------------------------------------------------------------
# a1  simulate_logistic                                              dlsa/models.py:6-40
# --------------------------------------------------------------------------------------------


def simulate_logistic(sample_size, p, partition_method="systematic", partition_num=1, seed=None):
    """Same law as dlsa/models.py:6-40: X~U(-.5,.5) (:22), beta=1 on the first int(0.4p) (:12,18-19),
    label~Bernoulli(sigmoid(X beta)) (:23,28-30), partition_id = i % K (:32-33).  Vectorised and
    seeded (the reference loops row by row over an unseeded global RNG)."""
    if partition_method != "systematic":
        raise Exception("No such partition method implemented!")  # dlsa/models.py:34-35
    rng = np.random.default_rng(seed)
    p1 = int(p * 0.4)
    beta = np.zeros(p)
    beta[:p1] = 1.0
    X = rng.random((sample_size, p)) - 0.5
    prob = 1.0 / (1.0 + np.exp(-X @ beta))
    y = (rng.random(sample_size) < prob).astype(np.float64)
    pid = (np.arange(sample_size) % partition_num).astype(np.float64)
    return pid, y, X, beta


# --------------------------------------------------------------------------------------------
# a9 / a3  per-partition MLE: exact Newton                          dlsa/R/wlse_func.R:24-65
# --------------------------------------------------------------------------------------------


def _sigmoid(eta):
    out = np.empty_like(eta)
    pos = eta >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-eta[pos]))
    e = np.exp(eta[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def _log1pexp(eta):
    return np.maximum(eta, 0.0) + np.log1p(np.exp(-np.abs(eta)))


def irls_pass(X, y, beta):
    """One Newton/IRLS pass (the quantity the fused kernel K1 produces for one partition):
    eta = X beta, p = sigmoid(eta), w = p(1-p), g = X'(y-p), H = X' diag(w) X, loglik.
    dlsa/R/wlse_func.R:47-54 (grad, hess) and dlsa/models.py:114,130 (Sig_inv at the final beta)."""
    eta = X @ beta
    pr = _sigmoid(eta)
    w = pr * (1.0 - pr)
    g = X.T @ (y - pr)
    H = X.T @ (w[:, None] * X)
    ll = float(np.sum(y * eta - _log1pexp(eta)))
    return g, H, ll


def logistic_newton(X, y, tol=1e-10, max_iter=50, damp=True):
    """Exact Newton from beta=0 (dlsa/R/wlse_func.R:28-31), step delta = H^-1 g (:55), damping
    delta*0.01 when max|delta| > 5 (:56-60), stop on max|delta| <= tol (:40; the R default is 1e-4,
    the oracle setting is 1e-10 to match sklearn newton-cg at tol=1e-10).
    Returns (beta, n_iter)."""
    p = X.shape[1]
    beta = np.zeros(p)
    for it in range(1, max_iter + 1):
        g, H, _ = irls_pass(X, y, beta)
        delta = np.linalg.solve(H, g)
        md = np.max(np.abs(delta))
        if damp and md > 5:
            delta = delta * 0.01
        beta = beta + delta
        if md <= tol:
            return beta, it
    return beta, max_iter


def poisson_pass(X, y, beta):
    """One Newton pass of the Poisson log-link GLM — what R's glm.fit iterates for family = "poisson"
    (dlsa/R/wlse_func.R:155-168): mu = exp(X beta), g = X'(y - mu), H = X' diag(mu) X (= solve(vcov) at the MLE),
    loglik = sum y*eta - mu (the -log(y!) constant is dropped)."""
    eta = X @ beta
    mu = np.exp(eta)
    return X.T @ (y - mu), (X * mu[:, None]).T @ X, float(np.sum(y * eta - mu))


def poisson_fit(X, y, fit_intercept=False, tol=1e-10, max_iter=50):
    """MLE of glm(y ~ x - 1, family = "poisson") by exact Newton from beta = 0 with the step rule of the logistic fitter
    (delta*0.01 when max|delta| > 5, stop on max|delta| <= tol); returns (coef, Sig_inv = X'diag(mu)X, Sig_inv coef)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    if fit_intercept:
        X = np.concatenate([np.ones((X.shape[0], 1)), X], axis=1)
    beta = np.zeros(X.shape[1])
    for _ in range(max_iter):
        g, H, _ = poisson_pass(X, y, beta)
        delta = np.linalg.solve(H, g)
        if np.max(np.abs(delta)) > 5:
            delta = delta * 0.01
        if np.max(np.abs(delta)) <= tol:
            break
        beta = beta + delta
    _, H, _ = poisson_pass(X, y, beta)
    return beta, H, H @ beta


def logistic_fit(X, y, fit_intercept=False, tol=1e-10, max_iter=50):
    """dlsa/models.py:110-131 with the fit done by exact Newton: returns (coef[p], Sig_inv[p,p],
    Sig_invMcoef[p]); with an intercept the design is [1, X] and the intercept comes first (:116-122).
    Sig_inv is unnormalised and evaluated at the returned coef (:114,130)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    if fit_intercept:
        X = np.concatenate([np.ones((X.shape[0], 1)), X], axis=1)
    coef, _ = logistic_newton(X, y, tol=tol, max_iter=max_iter)
    _, Sig_inv, _ = irls_pass(X, y, coef)
    return coef, Sig_inv, Sig_inv @ coef


# --------------------------------------------------------------------------------------------
# a2 dummy branch on the compact representation                       dlsa/models.py:56-79,99-101
# --------------------------------------------------------------------------------------------


def expand_onehot(numeric, codes, n_levels, mean=None, std=None):
    """The dense design the dummy branch of logistic_model builds (dlsa/models.py:56-79): numeric columns first
    (optionally standardised ``(x-mean)/std``, :99-101), then for every categorical column c its ``n_levels[c]`` 0/1
    level columns (``pd.get_dummies``, :62-65, after the baseline level was dropped, :78).  ``codes[i, c]`` is the
    index of row i's level inside column c's block, or -1 for the baseline level (all zeros)."""
    numeric = np.zeros((codes.shape[0], 0)) if numeric is None else np.asarray(numeric, dtype=np.float64)
    n = codes.shape[0]
    num = numeric if mean is None else (numeric - np.asarray(mean)) / np.asarray(std)
    blocks = [num]
    for c, L in enumerate(n_levels):
        B = np.zeros((n, L))
        rows = np.nonzero(codes[:, c] >= 0)[0]
        B[rows, codes[rows, c]] = 1.0
        blocks.append(B)
    return np.concatenate(blocks, axis=1)


def irls_pass_cat(numeric, codes, n_levels, y, beta, fit_intercept=False, mean=None, std=None):
    """What kernel K1c must produce for one partition: ``irls_pass`` on the expanded design (intercept first)."""
    X = expand_onehot(numeric, codes, n_levels, mean, std)
    if fit_intercept:
        X = np.concatenate([np.ones((X.shape[0], 1)), X], axis=1)
    return irls_pass(X, y, beta)


# --------------------------------------------------------------------------------------------
# a10  linear / OLS block                                             dlsa/R/wlse_func.R:169-182
# --------------------------------------------------------------------------------------------


def ols_pass(X, y):
    """Sufficient statistics of one block (what kernel K8 produces): X'X, X'y, y'y."""
    return X.T @ X, X.T @ y, float(y @ y)


def ols_fit(X, y):
    """R ``lm(y~x-1)``; ``SI = solve(vcov)`` = X'X / sigma2 with sigma2 = RSS/(n-p)
    (dlsa/R/wlse_func.R:169-182).  Returns (theta, Sig_inv, Sig_inv theta, sigma2)."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    n, p = X.shape
    XtX, Xty, yty = ols_pass(X, y)
    theta = np.linalg.solve(XtX, Xty)
    rss = yty - theta @ Xty
    sigma2 = rss / (n - p)
    return theta, XtX / sigma2, Xty / sigma2, sigma2


# --------------------------------------------------------------------------------------------
# a5  one-round combine                                               dlsa/dlsa.py:30-59
# --------------------------------------------------------------------------------------------


def combine(coefs, Sig_invs, Sig_invMcoefs=None, num_partitions=None):
    """S = sum_k Sig_inv_k, v = sum_k Sig_inv_k coef_k, c = sum_k coef_k (dlsa/dlsa.py:30-34);
    beta_byOLS = lstsq(S, v) (:48); beta_byONESHOT = c / K (:51-52).  Returns (theta, oneshot, S)."""
    coefs = np.asarray(coefs, dtype=np.float64)
    Sig_invs = np.asarray(Sig_invs, dtype=np.float64)
    K = coefs.shape[0] if num_partitions is None else num_partitions
    if Sig_invMcoefs is None:
        Sig_invMcoefs = np.einsum("kij,kj->ki", Sig_invs, coefs)
    S = Sig_invs.sum(0)
    v = np.asarray(Sig_invMcoefs).sum(0)
    theta = np.linalg.lstsq(S, v, rcond=None)[0]
    return theta, coefs.sum(0) / K, S


# --------------------------------------------------------------------------------------------
# a7 / a8  LARS on the least-squares approximation       dlsa/lsa.py:90-212, R dlsa_alasso_func.R:107-255
# --------------------------------------------------------------------------------------------


def _chol_append(R, xnew, xold, eps):
    """``updateR`` (dlsa/lsa.py:12-32): append one column to the upper Cholesky factor R of the active
    Gram.  Returns (R_new, rank_increased)."""
    norm_xnew = math.sqrt(xnew)
    if R is None:
        return np.array([[norm_xnew]]), True
    r = np.linalg.solve(R.T, xold)  # backsolvet, dlsa/lsa.py:8-9
    rpp = norm_xnew ** 2 - float(np.sum(r ** 2))
    if rpp <= eps:  # dlsa/lsa.py:24-25: clamp, rank not increased
        rpp = eps
        grew = False
    else:
        rpp = math.sqrt(rpp)
        grew = True
    a = R.shape[0]
    Rn = np.zeros((a + 1, a + 1))
    Rn[:a, :a] = R
    Rn[:a, a] = r
    Rn[a, a] = rpp
    return Rn, grew


def _chol_delete(R, k):
    """``downdateR`` / ``delcol`` (dlsa/lsa.py:35-80): drop column k of upper-triangular R and restore
    triangularity with Givens rotations (same branch rule on |b|>|a| as :46-53)."""
    a_ = R.shape[0]
    if a_ == 1:
        return None
    r = np.delete(R, k, axis=1)  # a_ x (a_-1)
    for i in range(k + 1, a_):
        a = r[i - 1, i - 1]
        b = r[i, i - 1]
        if b != 0:
            if not (abs(b) > abs(a)):
                tau = -b / a
                c = 1 / math.sqrt(1 + tau * tau)
                s = c * tau
            else:
                tau = -a / b
                s = 1 / math.sqrt(1 + tau * tau)
                c = s * tau
            r[i - 1, i - 1] = c * a - s * b
            r[i, i - 1] = s * a + c * b
            for j in range(i, a_ - 1):
                aa = r[i - 1, j]
                bb = r[i, j]
                r[i - 1, j] = c * aa - s * bb
                r[i, j] = s * aa + c * bb
    return r[: a_ - 1, :]


def lars_lsa(Sigma0, b0, intercept, n, type="lar", eps=EPS, max_steps=None):
    """LARS / lasso path on the quadratic form (b-beta)' Sigma (b-beta) with adaptive weights 1/|b|.

    intercept=False follows dlsa/lsa.py:90-212 statement by statement (0-based indices here).
    intercept=True follows the R original, dlsa/R/dlsa_alasso_func.R:115-125 and :238-248 (the
    Python port of that branch indexes by the sample size and cannot run).
    Returns dict(AIC, BIC, beta[(k+1) x m], beta0[k+1]) plus diagnostic 'RSS', 'dof'.
    """
    Sigma0 = np.asarray(Sigma0, dtype=np.float64)
    b0 = np.asarray(b0, dtype=np.float64).ravel()
    if intercept:
        a11 = Sigma0[0, 0]
        a12 = Sigma0[1:, 0].copy()
        Sigma = Sigma0[1:, 1:] - np.outer(a12, a12) / a11
        b = b0[1:].copy()
        beta0_init = float(a12 @ b) / a11
    else:
        Sigma = Sigma0.copy()
        b = b0.copy()
    absb = np.abs(b)
    Sigma = absb[:, None] * Sigma * absb[None, :]  # dlsa/lsa.py:108
    s = np.sign(b)  # :109
    m = Sigma.shape[1]
    Cvec = s @ Sigma  # :114
    if max_steps is None:
        max_steps = 8 * m
    beta = np.zeros((max_steps + 1, m))
    active = []  # ordered, 0-based
    Sign = []
    ignores = []
    drops = None  # boolean array over active, or None
    R = None
    inactive = list(range(m))
    k = 0
    while k < max_steps and len(active) < m:
        k += 1
        C = Cvec[inactive]
        Cmax = np.max(np.abs(C))
        if drops is None or not np.any(drops):
            newmask = np.abs(C) >= Cmax - eps  # :131  (C itself is re-formed over `rest` below, == :132)
            new = [inactive[i] for i in np.nonzero(newmask)[0]]
            for inew in new:
                Rn, grew = _chol_append(R, Sigma[inew, inew], Sigma[inew, active], eps)
                if not grew and R is not None:
                    # singular: back out (R semantics, dlsa_alasso_func.R:168-173); ignore the variable
                    ignores.append(inew)
                else:
                    R = Rn
                    active.append(inew)
                    Sign.append(np.sign(Cvec[inew]))
        Sg = np.asarray(Sign, dtype=np.float64)
        Gi1 = np.linalg.solve(R, np.linalg.solve(R.T, Sg))  # :151
        A = 1.0 / math.sqrt(float(np.sum(Gi1 * Sg)))  # :152
        w = A * Gi1  # :153
        if len(active) >= m:
            gamhat = Cmax / A  # :155
        else:
            excl = set(active) | set(ignores)
            rest = [j for j in range(m) if j not in excl]  # np.delete(Sigma, active+ignores, axis=1)
            a = w @ Sigma[np.ix_(active, rest)]  # :157-160
            # (:132,161) the reference's C is Cvec over `inactive` minus this step's joiners and `a` is over
            # all columns minus active and ignores: the same index set whenever nothing has been ignored.
            # With ignored (machine-singular) variables the reference's two vectors differ in length and it
            # raises; here both are taken over `rest` (the R `lars` convention).
            C = Cvec[rest]
            gam = np.concatenate([(Cmax - C) / (A - a), (Cmax + C) / (A + a)])
            cand = gam[gam > eps]
            gamhat = min(np.min(cand), Cmax / A) if cand.size else Cmax / A  # :162
        if type == "lasso":
            act = np.asarray(active, dtype=int)
            b1 = beta[k - 1, act]
            z1 = -b1 / w
            pos = z1[z1 > eps]
            zmin = min(np.min(pos), gamhat) if pos.size else gamhat
            if zmin < gamhat:
                gamhat = zmin
                drops = z1 == zmin
            else:
                drops = None
        beta[k] = beta[k - 1]
        act = np.asarray(active, dtype=int)
        beta[k, act] += gamhat * w  # :176
        Cvec = Cvec - gamhat * (Sigma[:, act] @ w)  # :177
        if type == "lasso" and drops is not None and np.any(drops):
            dropidx = np.nonzero(drops)[0]
            for idx in dropidx[::-1]:
                R = _chol_delete(R, int(idx))
            for idx in dropidx:
                beta[k, active[idx]] = 0.0
            keep = ~drops
            active = [a_ for a_, kp in zip(active, keep) if kp]
            Sign = [s_ for s_, kp in zip(Sign, keep) if kp]
        inactive = [j for j in range(m) if j not in set(active) and j not in set(ignores)]  # :188
    beta = beta[: k + 1]
    dff = s[None, :] - beta  # (k+1) x m
    RSS = np.einsum("ki,ij,kj->k", dff, Sigma, dff)  # :191-192
    if intercept:
        beta = beta * np.abs(b0[1:])[None, :]
        beta0 = beta0_init - (beta @ a12) / a11  # R :243-245
    else:
        beta = beta * np.abs(b0)[None, :]  # :198-201
        beta0 = np.zeros(k + 1)
    dof = np.sum(np.abs(beta) > eps, axis=1)  # :208
    BIC = RSS + math.log(n) * dof  # :209
    AIC = RSS + 2 * dof  # :210
    return {"AIC": AIC, "BIC": BIC, "beta": beta, "beta0": beta0, "RSS": RSS, "dof": dof}


# --------------------------------------------------------------------------------------------
# a6  dlsa(): path + information-criterion selection                 dlsa/dlsa.py:70-107
# --------------------------------------------------------------------------------------------


def dlsa_select(Sig_inv, beta_ols, sample_size, fit_intercept=False, type="lar"):
    """dlsa/dlsa.py:83-107: first-argmin AIC and BIC over the path (:90-91).  With an intercept the
    returned vectors are [beta0, beta] with ``beta0 = path beta0 + beta_ols[0]`` as in the R original
    (dlsa/R/dlsa_alasso_func.R:54-58; the Python port drops that term).
    Returns (beta_byAIC, beta_byBIC, fitted)."""
    beta_ols = np.asarray(beta_ols, dtype=np.float64).ravel()
    fitted = lars_lsa(np.asarray(Sig_inv), beta_ols, fit_intercept, sample_size, type=type)
    ia = int(np.argmin(fitted["AIC"]))
    ib = int(np.argmin(fitted["BIC"]))
    if fit_intercept:
        b0 = fitted["beta0"] + beta_ols[0]
        return (np.concatenate([[b0[ia]], fitted["beta"][ia]]),
                np.concatenate([[b0[ib]], fitted["beta"][ib]]), fitted)
    return fitted["beta"][ia].copy(), fitted["beta"][ib].copy(), fitted


# --------------------------------------------------------------------------------------------
# f1  log-likelihood evaluation                                       dlsa/models.py:151-225
# --------------------------------------------------------------------------------------------


def logistic_loglik(X, y, par):
    """dlsa/models.py:216-222: for every column beta of ``par`` (p x q):
    sum_i y_i log(p_i) + (1-y_i) log(1-p_i), p = 1/(1+exp(-X beta))."""
    X = np.asarray(X, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64).ravel()
    par = np.asarray(par, dtype=np.float64)
    out = np.empty(par.shape[1])
    for i in range(par.shape[1]):
        with np.errstate(divide="ignore", over="ignore"):
            prob = 1 / (1 + np.exp(-(X @ par[:, i])))
            out[i] = np.sum(y * np.log(prob) + (1 - y) * np.log(1 - prob))
    return out


# --------------------------------------------------------------------------------------------
# a11  dummy layout                                   dlsa/dummies.py:52-108, dlsa/models.py:56-91
# --------------------------------------------------------------------------------------------


def select_dummy_factors(dummy_dict, keep_top, replace_with="000_OTHERS"):
    """dlsa/dummies.py:52-108 without the pickle side effect: levels (in dict insertion order) whose
    cumulative share <= keep_top[i] are selected (:72-81), the rest dropped (:78,83-84); names are
    ``<col>_<level>`` with ``<col>_<replace_with>`` first when something was dropped (:87-96)."""
    factor_set, factor_selected, factor_dropped, factor_selected_names = {}, {}, {}, {}
    for i, col in enumerate(list(dummy_dict)):
        levels = list(dummy_dict[col].keys())
        counts = np.asarray(list(dummy_dict[col].values()), dtype=np.float64)
        cum = np.cumsum(counts)
        share = cum / cum[-1]
        sel = [lv for lv, s_ in zip(levels, share) if s_ <= keep_top[i]]
        drp = [lv for lv, s_ in zip(levels, share) if s_ > keep_top[i]]
        factor_set[col] = levels
        factor_selected[col] = sel
        factor_dropped[col] = drp
        new = ([replace_with] if len(drp) > 0 else []) + sel
        factor_selected_names[col] = [col + "_" + str(x) for x in new]
    return {"factor_set": factor_set, "factor_selected": factor_selected,
            "factor_dropped": factor_dropped, "factor_selected_names": factor_selected_names}


def design_columns(raw_columns, Y_name, dummy_info, dummy_factors_baseline, fit_intercept):
    """Canonical X column order of dlsa/models.py:71-79: sorted non-dummy names, then for each dummy
    column (dict order) its *string-sorted* selected names, minus baselines; intercept first."""
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


def logistic_model_frame(sample_df, Y_name, fit_intercept=False, dummy_info=[], dummy_factors_baseline=[],
                         data_info=[], tol=1e-10):
    """Frame-level restatement of ``logistic_model`` (dlsa/models.py:42-147): p x (3+p) frame
    ``[par_id, coef, Sig_invMcoef] + X names`` (:136-142); a chunk missing an expected dummy level
    gives a warning and an all-zero frame (:84-91); optional standardisation with rows 1,2 of
    ``data_info`` (:99-101)."""
    import pandas as pd

    sample_df = sample_df.reset_index(drop=True)
    has_dummies = dummy_info is not None and len(dummy_info) > 0
    xcols = design_columns(list(sample_df.columns), Y_name, dummy_info if has_dummies else None,
                           dummy_factors_baseline, False)
    icpt = ["intercept"] if fit_intercept else []
    if has_dummies:
        conv = list(dummy_info["factor_selected"].keys())
        df = sample_df.replace({k: v for k, v in dummy_info["factor_dropped"].items() if len(v) > 0}, "000_OTHERS")
        numeric = sorted(set(df.columns) - {"partition_id", Y_name} - set(conv))
        present = set(numeric)
        for c in conv:
            present |= {c + "_" + str(v) for v in pd.unique(df[c])}
        present -= set(dummy_factors_baseline)
        if present != set(xcols):
            warnings.warn("Dummies:" + str(set(xcols) - present) + "missing in this data chunk "
                          + "Skip modeling this part of data.")
            return pd.DataFrame(0, index=np.arange(len(icpt + xcols)),
                                columns=["par_id", "coef", "Sig_invMcoef"] + icpt + xcols)
        X = np.empty((df.shape[0], len(xcols)))
        for j, name in enumerate(xcols):
            if name in numeric:
                X[:, j] = df[name].to_numpy(dtype=np.float64)
            else:
                c = next(c for c in conv if name.startswith(c + "_") and
                         name in dummy_info["factor_selected_names"][c])
                X[:, j] = (df[c].astype(str) == name[len(c) + 1:]).to_numpy(dtype=np.float64)
        std_cols = numeric
    else:
        X = np.array(sample_df[xcols].to_numpy(dtype=np.float64))  # a writable copy (pandas 3 hands out read-only views)
        std_cols = xcols
    if data_info is not None and len(data_info) > 0:
        for name in std_cols:
            j = xcols.index(name)
            X[:, j] = (X[:, j] - float(data_info[name][1])) / float(data_info[name][2])
    y = sample_df[Y_name].to_numpy(dtype=np.float64)
    coef, Sig_inv, Smc = logistic_fit(X, y, fit_intercept=fit_intercept, tol=tol)
    p = coef.size
    out = pd.DataFrame(np.concatenate([coef[:, None], Smc[:, None], Sig_inv], axis=1),
                       columns=["coef", "Sig_invMcoef"] + icpt + xcols)
    out.insert(0, "par_id", np.arange(p))
    return out
