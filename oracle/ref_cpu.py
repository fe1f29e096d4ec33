"""TEST / BASELINE INFRASTRUCTURE ONLY — the reference's OWN CPU implementation of the hot path, timed on host cores.

``bench.py --impl reference`` and the ``cpu_baseline`` leg of the default run call this module.  It runs the
unmodified reference functions (``dlsa.models.logistic_model`` — both its plain and its dummy branch —, the
group-sum + ``lstsq`` of ``dlsa.dlsa.dlsa_mapred`` and ``dlsa.lsa.lars_lsa``) through the in-process API-drift shims
of ``oracle/ref_shim.py``, from ``/root/reference`` in the build container or from the git-ignored install under
``baseline/_ref`` on the GPU box (``kind = "reference"``).  When neither tree exists it falls back to the port in
``oracle/ref_port.py`` (``kind = "port"``).

What is and is not inside the timed region (stated in every JSON line this feeds):
  * inside: ``logistic_model`` per partition exactly as shipped — sklearn ``newton-cg`` with its DEFAULT ``tol``
    (the reference never sets it, dlsa/models.py:110-112), ``predict_proba``, the pandas Gram, the p x (3+p) frame —
    then the pandas group-sum + ``np.linalg.lstsq`` (dlsa/dlsa.py:30-59) and ``lars_lsa`` + argmin (dlsa/dlsa.py:83-91);
  * mapped over one process per host core with BLAS pinned to one thread, mirroring the 1-core Spark executors of
    projects/bash/run_spark_dlsa.sh:15; frames travel back through a multiprocessing queue (the analogue of the
    Arrow hop — cheaper than Spark's);
  * excluded: data generation, Spark / JVM / Arrow / shuffle costs (Spark cannot run here) — which flatters the reference.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time
import warnings

import numpy as np


def reference_available() -> bool:
    from oracle import ref_shim
    return ref_shim.available()


# ---- per-partition synthetic data (generation is never timed) ----------------------------------------------------
def _sim_partition(seed, k, n_k, p):
    """Law of simulate_logistic (dlsa/models.py:6-40): X ~ U(-.5,.5), beta = 1 on the first int(0.4 p), Bernoulli labels."""
    beta = np.zeros(p)
    beta[: int(p * 0.4)] = 1.0
    rng = np.random.default_rng([seed, k])
    X = rng.random((n_k, p)) - 0.5
    y = (rng.random(n_k) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    return X, y


AIRLINE_LEVELS = (12, 7, 8, 76, 76)
AIRLINE_ZIPF = (False, False, True, True, True)
AIRLINE_NAMES = ("Month", "DayOfWeek", "UniqueCarrier", "Origin", "Dest")
AIRLINE_NUMERICS = ("ActualElapsedTime", "CRSArrTime", "CRSDepTime", "DayofMonth", "DepTime", "Distance", "Year")


def _airline_partition(seed, k, n_k):
    """Airline-shaped frame (SURVEY.md section 8d "C3"): 7 numerics + Month / DayOfWeek / UniqueCarrier / Origin / Dest
    with uniform or Zipf(1) level shares; level 1 of every column is the dummy baseline."""
    import pandas as pd
    rng = np.random.default_rng([seed, 7_000_000 + k])
    cols = {"partition_id": np.full(n_k, float(k)), "label": np.zeros(n_k)}
    eta = np.full(n_k, -1.0)
    brng = np.random.default_rng(11)
    for name in AIRLINE_NUMERICS:
        v = rng.standard_normal(n_k)
        cols[name] = v
        eta += v * brng.standard_normal() * 0.3
    for name, L, zipf in zip(AIRLINE_NAMES, AIRLINE_LEVELS, AIRLINE_ZIPF):
        w = 1.0 / np.arange(1, L + 1) if zipf else np.ones(L)
        lvl = rng.choice(L, size=n_k, p=w / w.sum()) + 1
        cols[name] = lvl
        eta += (brng.standard_normal(L) * 0.3)[lvl - 1] * (lvl > 1)
    cols["label"] = (rng.random(n_k) < 1 / (1 + np.exp(-eta))).astype(np.float64)
    return pd.DataFrame(cols)


def _airline_dummy_info():
    """The dummy_info dict select_dummy_factors would produce with every level kept (dlsa/dummies.py:52-108)."""
    sel = {n: list(range(1, L + 1)) for n, L in zip(AIRLINE_NAMES, AIRLINE_LEVELS)}
    return {"factor_set": {n: list(v) for n, v in sel.items()}, "factor_selected": sel,
            "factor_dropped": {n: [] for n in AIRLINE_NAMES},
            "factor_selected_names": {n: [n + "_" + str(v) for v in vals] for n, vals in sel.items()}}


def _worker(rank, nproc, spec, barrier, outq, total, use_ref):
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    warnings.filterwarnings("ignore")
    import pandas as pd
    K, n_k, p, seed, shape = spec["K"], spec["n_k"], spec["p"], spec["seed"], spec["shape"]
    mine = list(range(rank, K, nproc))
    if shape == "airline":
        frames_in = [_airline_partition(seed, k, n_k) for k in mine]
        info = _airline_dummy_info()
        base = [n + "_1" for n in AIRLINE_NAMES]
    else:
        names = ["x" + str(i) for i in range(p)]
        frames_in = []
        for k in mine:
            X, y = _sim_partition(seed, k, n_k, p)
            df = pd.DataFrame(X, columns=names)
            df.insert(0, "label", y)
            df.insert(0, "partition_id", float(k))
            frames_in.append(df)
    if use_ref:
        from oracle import ref_shim
        m = ref_shim.ref_models(tol=1e-4)   # 1e-4 IS sklearn's default: the reference as shipped

        def fit(df):
            if shape == "airline":
                return m.logistic_model(df, "label", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base)
            return m.logistic_model(df, "label")
    else:
        from oracle import ref_port

        def fit(df):
            return ref_port.port_logistic_model(df.iloc[:, 2:].to_numpy(), df["label"].to_numpy(), list(df.columns[2:]))
    for _ in range(total):
        barrier.wait()
        out = [fit(df) for df in frames_in]
        outq.put((rank, out))
        barrier.wait()


def run_reference_cpu(K, n_k, p, seed=1, nproc=None, reps=1, warmup=0, shape="sim"):
    """Time ``reps`` passes of the reference CPU path over K partitions of n_k rows on ``nproc`` processes.
    ``shape``: "sim" (no-dummy branch, p features) or "airline" (dummy branch, 181 + intercept).
    Returns dict(seconds=[...], rows, cores, kind, theta, support_bic)."""
    from oracle import ref_shim
    use_ref = ref_shim.available()
    nproc = min(nproc or os.cpu_count() or 1, K)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(nproc + 1)
    outq = ctx.Queue()
    total = reps + warmup
    spec = {"K": K, "n_k": n_k, "p": p, "seed": seed, "shape": shape}
    procs = [ctx.Process(target=_worker, args=(r, nproc, spec, barrier, outq, total, use_ref)) for r in range(nproc)]
    for pr in procs:
        pr.start()
    secs, theta, fit = [], None, None
    intercept = shape == "airline"
    for it in range(total):
        barrier.wait()  # workers hold their data: start the clock
        t0 = time.perf_counter()
        frames = []
        for _ in range(nproc):
            frames.extend(outq.get()[1])
        if use_ref:
            mr = ref_shim.ref_mapred(frames, K)              # dlsa/dlsa.py:30-59 with the Spark group-sum in pandas
            S = mr.iloc[:, 2:].to_numpy()
            theta = mr["beta_byOLS"].to_numpy()
        else:
            import pandas as pd
            g = pd.concat(frames, axis=0, ignore_index=True).groupby("par_id").sum().sort_index()
            S = g.iloc[:, 2:].to_numpy()
            theta = np.linalg.lstsq(S, g.iloc[:, 1].to_numpy(), rcond=None)[0]
        if use_ref and not intercept:
            fit = ref_shim.ref_lars_lsa(S, theta, False, K * n_k)   # dlsa/lsa.py:90-212 (np.matrix call form)
        else:  # the Python reference's intercept branch raises IndexError: the oracle follows R
            from oracle import numpy_dlsa as onp
            fit = onp.lars_lsa(S, theta, intercept, K * n_k)
        ib = int(np.argmin(fit["BIC"]))
        t1 = time.perf_counter()
        barrier.wait()
        if it >= warmup:
            secs.append(t1 - t0)
    for pr in procs:
        pr.join()
    return {"seconds": secs, "rows": K * n_k, "cores": nproc, "kind": "reference" if use_ref else "port",
            "theta": theta, "support_bic": (fit["beta"][ib] != 0)}


# ---- OLS DLSA (BASELINE configs[3]): the reference's implementation is R lm() per block (dlsa/R/wlse_func.R:169-182) and
# a statsmodels per-group precedent (projects/results/linear_regression_dc.py:27-39); neither R nor statsmodels exists in
# this image, so this leg is a NumPy/LAPACK port ("kind": "port").
def _ols_worker(rank, nproc, spec, barrier, outq, total):
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    K, n_k, p, seed = spec["K"], spec["n_k"], spec["p"], spec["seed"]
    beta = np.tile([3, 1.5, 0, 0, 2, 0, 0, 0, 0, 0], (p + 9) // 10)[:p].astype(np.float64)  # dlsa/R/main.R:26
    data = []
    for k in range(rank, K, nproc):
        rng = np.random.default_rng([seed, 9_000_000 + k])
        X = rng.standard_normal((n_k, p))
        data.append((X, X @ beta + rng.standard_normal(n_k)))
    for _ in range(total):
        barrier.wait()
        out = []
        for X, y in data:
            theta, rss, _, _ = np.linalg.lstsq(X, y, rcond=None)      # lm(y ~ x - 1)
            sigma2 = float(rss[0]) / (X.shape[0] - p) if rss.size else float(np.sum((y - X @ theta) ** 2)) / (X.shape[0] - p)
            XtX = X.T @ X                                              # solve(vcov(fit)) = X'X / sigma2
            out.append((theta, XtX / sigma2))
        outq.put((rank, out))
        barrier.wait()


def run_ols_cpu(K, n_k, p, seed=1, nproc=None, reps=1, warmup=0):
    from oracle import numpy_dlsa as onp
    nproc = min(nproc or os.cpu_count() or 1, K)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(nproc + 1)
    outq = ctx.Queue()
    total = reps + warmup
    spec = {"K": K, "n_k": n_k, "p": p, "seed": seed}
    procs = [ctx.Process(target=_ols_worker, args=(r, nproc, spec, barrier, outq, total)) for r in range(nproc)]
    for pr in procs:
        pr.start()
    secs, theta, fit = [], None, None
    for it in range(total):
        barrier.wait()
        t0 = time.perf_counter()
        res = []
        for _ in range(nproc):
            res.extend(outq.get()[1])
        S = sum(s for _, s in res)
        v = sum(s @ th for th, s in res)
        theta = np.linalg.lstsq(S, v, rcond=None)[0]                   # dlsa/R/wlse_func.R:198-200
        fit = onp.lars_lsa(S, theta, False, K * n_k)
        ib = int(np.argmin(fit["BIC"]))
        t1 = time.perf_counter()
        barrier.wait()
        if it >= warmup:
            secs.append(t1 - t0)
    for pr in procs:
        pr.join()
    return {"seconds": secs, "rows": K * n_k, "cores": nproc, "kind": "port", "theta": theta,
            "support_bic": (fit["beta"][ib] != 0)}
