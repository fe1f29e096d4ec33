"""Mint ``tests/golden/*.npz`` by running the REFERENCE's own functions (shimmed, oracle/ref_shim.py).

Run in the build container only (needs /root/reference):  ``python -m oracle.make_golden``.
Every array stored under a key starting with ``ref_`` is an output of reference code; keys starting
with ``in_`` are the inputs it was given (or the recipe to regenerate them from a seed).
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from oracle import numpy_dlsa as onp  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def sim_frame(n, p, K, seed):
    pid, y, X, beta = onp.simulate_logistic(n, p, "systematic", K, seed=seed)
    df = pd.DataFrame(np.concatenate([pid[:, None], y[:, None], X], axis=1),
                      columns=["partition_id", "label"] + ["x" + str(i) for i in range(p)])
    return df, X, y


def run_partitions(df, K, **kw):
    frames = []
    for k in range(K):
        part = df[df["partition_id"] == k]
        frames.append(ref_shim.ref_logistic_model(part, "label", **kw))
    return frames


def frames_to_arrays(frames):
    coef = np.stack([f["coef"].to_numpy() for f in frames])
    smc = np.stack([f["Sig_invMcoef"].to_numpy() for f in frames])
    sig = np.stack([f.iloc[:, 3:].to_numpy() for f in frames])
    return coef, smc, sig


def golden_games():
    """projects/results/data/games-expand.csv, K=5 systematic, no intercept."""
    csv = os.path.join(ref_shim.REFERENCE_ROOT, "projects/results/data/games-expand.csv")
    raw = pd.read_csv(csv)
    n = raw.shape[0]
    K = 5
    df = raw.copy()
    df.insert(0, "partition_id", np.arange(n) % K)
    frames = run_partitions(df, K)
    coef, smc, sig = frames_to_arrays(frames)
    mr = ref_shim.ref_mapred(frames, K)
    S = mr.iloc[:, 2:].to_numpy()
    theta = mr["beta_byOLS"].to_numpy()
    out = {"in_bits": np.packbits(raw.to_numpy().astype(np.uint8), axis=None), "in_shape": np.array(raw.shape),
           "in_columns": np.array(list(raw.columns)), "in_K": K,
           "ref_frame_columns": np.array(list(frames[0].columns)),
           "ref_coef": coef, "ref_Sig_invMcoef": smc, "ref_Sig_inv": sig,
           "ref_mapred_columns": np.array(list(mr.columns)),
           "ref_beta_byOLS": theta, "ref_beta_byONESHOT": mr["beta_byONESHOT"].to_numpy(), "ref_Sig_inv_sum": S}
    for typ in ("lar", "lasso"):
        fit = ref_shim.ref_lars_lsa(S, theta, False, n, type=typ)
        for k_ in ("AIC", "BIC", "beta", "beta0"):
            out[f"ref_{typ}_{k_}"] = fit[k_]
    # default-tol (as shipped) run, for the "distance to the as-shipped reference" report
    frames4 = [ref_shim.ref_logistic_model(df[df["partition_id"] == k], "label", tol=1e-4) for k in range(K)]
    mr4 = ref_shim.ref_mapred(frames4, K)
    out["ref_tol1e-4_beta_byOLS"] = mr4["beta_byOLS"].to_numpy()
    out["ref_tol1e-4_Sig_inv_sum"] = mr4.iloc[:, 2:].to_numpy()
    np.savez_compressed(os.path.join(GOLD, "games_expand_k5.npz"), **out)
    print("games: theta", theta[:3], "BIC argmin", int(np.argmin(out["ref_lar_BIC"])))


def golden_sim(name, n, p, K, seed, fit_intercept=False):
    df, X, y = sim_frame(n, p, K, seed)
    frames = run_partitions(df, K, fit_intercept=fit_intercept)
    coef, smc, sig = frames_to_arrays(frames)
    mr = ref_shim.ref_mapred(frames, K)
    S = mr.iloc[:, 2:].to_numpy()
    theta = mr["beta_byOLS"].to_numpy()
    out = {"in_n": n, "in_p": p, "in_K": K, "in_seed": seed, "in_fit_intercept": fit_intercept,
           "in_X_checksum": np.array([X.sum(), (X * X).sum(), y.sum()]),
           "ref_frame_columns": np.array(list(frames[0].columns)),
           "ref_coef": coef, "ref_Sig_invMcoef": smc, "ref_Sig_inv": sig,
           "ref_mapred_columns": np.array(list(mr.columns)),
           "ref_beta_byOLS": theta, "ref_beta_byONESHOT": mr["beta_byONESHOT"].to_numpy(), "ref_Sig_inv_sum": S}
    if not fit_intercept:
        for typ in ("lar", "lasso"):
            fit = ref_shim.ref_lars_lsa(S, theta, False, n, type=typ)
            for k_ in ("AIC", "BIC", "beta", "beta0"):
                out[f"ref_{typ}_{k_}"] = fit[k_]
    np.savez_compressed(os.path.join(GOLD, name + ".npz"), **out)
    print(name, "theta[:3]", theta[:3])


def golden_lars():
    """Stand-alone lars_lsa cases, including ones where the lasso variant drops a variable."""
    rng = np.random.default_rng(7)
    cases = {}
    idx = 0
    for p, rho, trial in [(6, 0.9, 0), (12, 0.95, 1), (30, 0.5, 2), (30, 0.97, 3), (50, 0.9, 4), (8, 0.99, 5)]:
        # SPD Sigma from a correlated design; b = a noisy sparse vector so some signs flip along the path
        n_obs = 400
        Z = rng.standard_normal((n_obs, p))
        Xc = Z.copy()
        for j in range(1, p):
            Xc[:, j] = rho * Xc[:, j - 1] + math_sqrt(1 - rho * rho) * Z[:, j]
        Sigma = Xc.T @ Xc
        b = rng.standard_normal(p) * (rng.random(p) < 0.6) + 0.05 * rng.standard_normal(p)
        for typ in ("lar", "lasso"):
            fit = ref_shim.ref_lars_lsa(Sigma, b, False, n_obs, type=typ)
            key = f"c{idx}_{typ}"
            cases[f"{key}_in_Sigma"] = Sigma
            cases[f"{key}_in_b"] = b
            cases[f"{key}_in_n"] = n_obs
            for k_ in ("AIC", "BIC", "beta", "beta0"):
                cases[f"{key}_ref_{k_}"] = fit[k_]
        lar_steps = cases[f"c{idx}_lar_ref_beta"].shape[0]
        lasso_steps = cases[f"c{idx}_lasso_ref_beta"].shape[0]
        print(f"lars case {idx}: p={p} rho={rho} lar rows={lar_steps} lasso rows={lasso_steps}")
        idx += 1
    # cases where the lasso variant really drops variables (path longer than p+1 rows): scan seeded draws
    rng = np.random.default_rng(123)
    n_drop = 0
    for trial in range(400):
        p = int(rng.integers(5, 16))
        n_obs = 200
        rho = rng.uniform(0.9, 0.999)
        Z = rng.standard_normal((n_obs, p))
        Xc = Z.copy()
        for j in range(1, p):
            Xc[:, j] = rho * Xc[:, j - 1] + math_sqrt(1 - rho * rho) * Z[:, j]
        Sigma = Xc.T @ Xc
        b = rng.standard_normal(p)
        fit = ref_shim.ref_lars_lsa(Sigma, b, False, n_obs, type="lasso")
        if fit["beta"].shape[0] == p + 1:
            continue
        for typ in ("lar", "lasso"):
            fit = ref_shim.ref_lars_lsa(Sigma, b, False, n_obs, type=typ)
            key = f"c{idx}_{typ}"
            cases[f"{key}_in_Sigma"] = Sigma
            cases[f"{key}_in_b"] = b
            cases[f"{key}_in_n"] = n_obs
            for k_ in ("AIC", "BIC", "beta", "beta0"):
                cases[f"{key}_ref_{k_}"] = fit[k_]
        print(f"lars case {idx}: p={p} lasso rows={cases[f'c{idx}_lasso_ref_beta'].shape[0]} (drops)")
        idx += 1
        n_drop += 1
        if n_drop >= 6:
            break
    cases["n_cases"] = idx
    np.savez_compressed(os.path.join(GOLD, "lars_cases.npz"), **cases)


def math_sqrt(x):
    import math
    return math.sqrt(x)


def golden_dummies():
    """Dummy-layout KAT of SURVEY App. C: select_dummy_factors + logistic_model(fit_intercept=True)."""
    rng = np.random.default_rng(11)
    n = 6000
    months = rng.integers(1, 13, size=n)
    carr_levels = np.array(list("ABCDEFGH"))
    probs = np.array([0.35, 0.25, 0.15, 0.08, 0.06, 0.05, 0.04, 0.02])
    carrier = carr_levels[rng.choice(8, size=n, p=probs)]
    znum = rng.standard_normal(n)
    anum = rng.standard_normal(n)
    eta = -0.3 + 0.5 * znum - 0.4 * anum + 0.2 * (months == 3) + 0.3 * (carrier == "B")
    y = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    df = pd.DataFrame({"partition_id": np.zeros(n), "y": y, "znum": znum, "anum": anum, "Month": months,
                       "Carrier": carrier})
    dm = ref_shim.ref_dummies()
    counts = dm.dummy_factors_counts(df, ["Month", "Carrier"])
    import tempfile
    with tempfile.TemporaryDirectory() as td:
        import contextlib, io
        with contextlib.redirect_stdout(io.StringIO()):
            info = dm.select_dummy_factors(counts, [1, 0.8], "000_OTHERS", os.path.join(td, "d.pkl"))
    baselines = ["Month_1", "Carrier_000_OTHERS"]
    frame = ref_shim.ref_logistic_model(df, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=baselines)
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter("always")
        zero = ref_shim.ref_models().logistic_model(df[df["Month"] != 7].reset_index(drop=True), "y", fit_intercept=True,
                                                    dummy_info=info, dummy_factors_baseline=baselines)
    out = {"in_y": y, "in_znum": znum, "in_anum": anum, "in_Month": months, "in_Carrier": carrier,
           "in_keep_top": np.array([1, 0.8]), "in_baselines": np.array(baselines),
           "ref_count_keys_Month": np.array([str(k) for k in counts["Month"].keys()]),
           "ref_count_vals_Month": np.array(list(counts["Month"].values())),
           "ref_count_keys_Carrier": np.array(list(counts["Carrier"].keys())),
           "ref_count_vals_Carrier": np.array(list(counts["Carrier"].values())),
           "ref_selected_Carrier": np.array(info["factor_selected"]["Carrier"]),
           "ref_dropped_Carrier": np.array(info["factor_dropped"]["Carrier"]),
           "ref_selected_names_Carrier": np.array(info["factor_selected_names"]["Carrier"]),
           "ref_selected_names_Month": np.array(info["factor_selected_names"]["Month"]),
           "ref_frame_columns": np.array(list(frame.columns)), "ref_frame": frame.to_numpy(dtype=float),
           "ref_zero_frame_columns": np.array(list(zero.columns)), "ref_zero_frame": zero.to_numpy(dtype=float),
           "ref_zero_warning": np.array([str(w.message) for w in wlist if "Dummies" in str(w.message)][:1])}
    np.savez_compressed(os.path.join(GOLD, "dummy_layout.npz"), **out)
    print("dummy frame", frame.shape, list(frame.columns)[:6], "zero-frame", zero.shape)


def golden_eval():
    """logistic_model_eval (dlsa/models.py:151-225) and logistic_model(..., data_info=...) (:99-101) run by the reference.

    Cases: (a) no dummies, no intercept; (b) no dummies, intercept + standardisation; (c) a dummy column whose
    get_dummies order coincides with the canonical (sorted) column order + intercept + standardisation.  The reference's
    eval does not reindex the dummy design to the coefficient order (it multiplies by position, :216-219), so a design
    whose get_dummies order differs from sorted order is evaluated against permuted coefficients — a reference defect;
    case (c) is built so both orders agree and the value is well defined."""
    out = {}
    m = ref_shim.ref_models()
    # (a), (b): simulated, p = 7, K = 3
    df, X, y = sim_frame(6000, 7, 3, 3)
    rng = np.random.default_rng(21)
    names = ["x" + str(i) for i in range(7)]
    par_a = pd.DataFrame(rng.standard_normal((7, 3)) * 0.5, columns=["beta_byOLS", "beta_byONESHOT", "beta_byBIC"])
    par_b = pd.DataFrame(rng.standard_normal((8, 2)) * 0.5, columns=["beta_byOLS", "beta_byAIC"])
    desc = df[names].describe()
    data_info = {c: desc[c].to_numpy() for c in names}
    ll_a, ll_b = [], []
    for k in range(3):
        part = df[df["partition_id"] == k].reset_index(drop=True)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            ll_a.append(m.logistic_model_eval(part.copy(), "label", par_a).to_numpy().ravel())
            ll_b.append(m.logistic_model_eval(part.copy(), "label", par_b, fit_intercept=True,
                                              data_info=data_info).to_numpy().ravel())
    out.update({"a_in_n": 6000, "a_in_p": 7, "a_in_K": 3, "a_in_seed": 3, "a_in_par": par_a.to_numpy(),
                "a_par_columns": np.array(list(par_a.columns)), "a_ref_loglik": np.stack(ll_a),
                "b_in_par": par_b.to_numpy(), "b_par_columns": np.array(list(par_b.columns)),
                "b_in_data_info": np.stack([data_info[c] for c in names]), "b_ref_loglik": np.stack(ll_b)})
    # standardised fit of the same partitions (logistic_model with data_info, intercept)
    frames = [ref_shim.ref_logistic_model(df[df["partition_id"] == k], "label", fit_intercept=True, data_info=data_info)
              for k in range(3)]
    coef, smc, sig = frames_to_arrays(frames)
    out.update({"b_ref_fit_columns": np.array(list(frames[0].columns)), "b_ref_fit_coef": coef,
                "b_ref_fit_Sig_invMcoef": smc, "b_ref_fit_Sig_inv": sig})
    # (c) dummy branch: numerics already in sorted order, one string-levelled categorical
    rng = np.random.default_rng(22)
    n = 5000
    carr_levels = np.array(list("ABCDEFGH"))
    carrier = carr_levels[rng.choice(8, size=n, p=np.array([0.35, 0.25, 0.15, 0.08, 0.06, 0.05, 0.04, 0.02]))]
    anum, znum = rng.standard_normal(n) * 2 + 1, rng.standard_normal(n) * 0.5 - 3
    eta = -0.3 - 0.4 * (anum - 1) / 2 + 0.5 * (znum + 3) / 0.5 * 0.3 + 0.3 * (carrier == "B")
    yv = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    dfc = pd.DataFrame({"partition_id": np.zeros(n), "y": yv, "anum": anum, "znum": znum, "Carrier": carrier})
    dm = ref_shim.ref_dummies()
    counts = dm.dummy_factors_counts(dfc, ["Carrier"])
    import contextlib, io, tempfile
    with tempfile.TemporaryDirectory() as td, contextlib.redirect_stdout(io.StringIO()):
        info = dm.select_dummy_factors(counts, [0.8], "000_OTHERS", os.path.join(td, "d.pkl"))
    base = ["Carrier_000_OTHERS"]
    descc = dfc[["anum", "znum"]].describe()
    dinfo = {c: descc[c].to_numpy() for c in ["anum", "znum"]}
    frame = ref_shim.ref_logistic_model(dfc, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base,
                                        data_info=dinfo)
    par_c = pd.DataFrame(np.stack([frame["coef"].to_numpy(), rng.standard_normal(frame.shape[0]) * 0.3], axis=1),
                         columns=["beta_byOLS", "beta_rand"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ll_c = m.logistic_model_eval(dfc.copy(), "y", par_c, fit_intercept=True, dummy_info=info,
                                     dummy_factors_baseline=base, data_info=dinfo).to_numpy().ravel()
    out.update({"c_in_y": yv, "c_in_anum": anum, "c_in_znum": znum, "c_in_Carrier": carrier, "c_in_keep_top": np.array([0.8]),
                "c_in_baselines": np.array(base), "c_in_data_info": np.stack([dinfo[c] for c in ["anum", "znum"]]),
                "c_ref_frame_columns": np.array(list(frame.columns)), "c_ref_frame": frame.to_numpy(dtype=float),
                "c_in_par": par_c.to_numpy(), "c_par_columns": np.array(list(par_c.columns)), "c_ref_loglik": ll_c})
    np.savez_compressed(os.path.join(GOLD, "eval_and_datainfo.npz"), **out)
    print("eval golden: loglik a", ll_a[0], "b", ll_b[0], "c", ll_c, "frame c", frame.shape, list(frame.columns))


def main():
    assert ref_shim.available(), "needs /root/reference"
    os.makedirs(GOLD, exist_ok=True)
    golden_games()
    golden_sim("sim_c1_n100000_p20_k16", 100_000, 20, 16, 20210517)
    golden_sim("sim_small_n6000_p7_k3", 6000, 7, 3, 3)
    golden_sim("sim_icpt_n20000_p12_k4", 20_000, 12, 4, 5, fit_intercept=True)
    golden_sim("sim_p50_n40000_k4", 40_000, 50, 4, 9)
    golden_lars()
    golden_dummies()
    golden_eval()


if __name__ == "__main__":
    main()
