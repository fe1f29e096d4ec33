"""TEST / BASELINE INFRASTRUCTURE ONLY — the reference's CPU path restated for timing on the GPU box.

``/root/reference`` does not travel to the GPU box and Spark/JVM are absent everywhere, so the CPU baseline
(``bench.py`` cpu_baseline leg and ``--impl reference``) is this port: the same third-party arithmetic the
reference calls — scikit-learn ``LogisticRegression(solver='newton-cg', max_iter=500)`` with its default
``tol`` exactly as shipped (dlsa/models.py:110-113), ``predict_proba`` (:114), the pandas ``X.T.dot(w*X)``
Gram (:130), the p x (3+p) frame (:136-142) — then the group-sum + lstsq of dlsa_mapred (dlsa/dlsa.py:30-59,
pandas stand-in for Spark) and lars_lsa (oracle/numpy_dlsa.py).  Partitions are mapped over one process per
host core with BLAS pinned to one thread, mirroring the 1-core executors of projects/bash/run_spark_dlsa.sh:15.
JVM / Arrow / shuffle costs are excluded, which flatters the reference.
"""
from __future__ import annotations

import multiprocessing as mp
import os
import time

import numpy as np


def port_logistic_model(X, y, names, tol=1e-4):
    """dlsa/models.py:93-147 (no-dummy branch) on one partition -> pandas frame."""
    import pandas as pd
    from sklearn.linear_model import LogisticRegression
    x_train = pd.DataFrame(X, columns=names)
    model = LogisticRegression(solver="newton-cg", C=np.inf, fit_intercept=False, max_iter=500, tol=tol)
    model.fit(x_train, y)
    prob = model.predict_proba(x_train)[:, 0]
    p = model.coef_.size
    coef = model.coef_.reshape(p, 1)
    Sig_inv = x_train.T.dot(np.multiply((prob * (1 - prob))[:, None], x_train))
    Sig_invMcoef = Sig_inv.dot(coef)
    out_np = np.concatenate((coef, Sig_invMcoef, Sig_inv), 1)
    out = pd.DataFrame(out_np, columns=["coef", "Sig_invMcoef"] + list(names))
    out.insert(0, "par_id", np.arange(p))
    return out


def _worker(rank, nproc, K, n_k, p, seed, barrier, q):
    for v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ[v] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass
    import warnings
    warnings.filterwarnings("ignore")
    names = ["x" + str(i) for i in range(p)]
    beta = np.zeros(p)
    beta[: int(p * 0.4)] = 1.0
    mine = list(range(rank, K, nproc))
    data = []
    for k in mine:  # same law as simulate_logistic (dlsa/models.py:6-40); generation is not timed
        rng = np.random.default_rng([seed, k])
        X = rng.random((n_k, p)) - 0.5
        y = (rng.random(n_k) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
        data.append((X, y))
    reps = q["reps"]
    for _ in range(reps):
        barrier.wait()
        frames = [port_logistic_model(X, y, names) for X, y in data]
        q["out"].put((rank, frames))
        barrier.wait()


def run_reference_cpu(K, n_k, p, seed=1, nproc=None, reps=1, warmup=0):
    """Time `reps` passes of the reference CPU path over K partitions of n_k rows x p on `nproc` processes.
    Returns dict(seconds=[...], rows=K*n_k, cores=nproc, theta=..., support_bic=...)."""
    import pandas as pd
    from oracle import numpy_dlsa as onp
    nproc = nproc or os.cpu_count() or 1
    nproc = min(nproc, K)
    ctx = mp.get_context("fork")
    barrier = ctx.Barrier(nproc + 1)
    outq = ctx.Queue()
    total = reps + warmup
    procs = [ctx.Process(target=_worker, args=(r, nproc, K, n_k, p, seed, barrier, {"reps": total, "out": outq}))
             for r in range(nproc)]
    for pr in procs:
        pr.start()
    secs = []
    theta = None
    fit = None
    for it in range(total):
        barrier.wait()  # workers have their data; start the clock
        t0 = time.perf_counter()
        frames = []
        for _ in range(nproc):
            frames.extend(outq.get()[1])
        # dlsa_mapred: group-sum + lstsq (dlsa/dlsa.py:30-59)
        allrows = pd.concat(frames, axis=0, ignore_index=True)
        g = allrows.groupby("par_id").sum().sort_index()
        S = g.iloc[:, 2:].to_numpy()
        theta = np.linalg.lstsq(S, g.iloc[:, 1].to_numpy(), rcond=None)[0]
        fit = onp.lars_lsa(S, theta, False, K * n_k)
        t1 = time.perf_counter()
        barrier.wait()
        if it >= warmup:
            secs.append(t1 - t0)
    for pr in procs:
        pr.join()
    ib = int(np.argmin(fit["BIC"]))
    return {"seconds": secs, "rows": K * n_k, "cores": nproc, "theta": theta,
            "support_bic": (fit["beta"][ib] != 0)}
