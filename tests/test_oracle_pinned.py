"""CPU tests: the oracle (oracle/numpy_dlsa.py) against outputs of the reference's own functions frozen in
tests/golden/ (minted by oracle/make_golden.py), and against the live shimmed reference when present."""
import warnings

import numpy as np
import pandas as pd
import pytest

from conftest import games_data, load_golden, sim_data
from oracle import numpy_dlsa as onp
from oracle import ref_shim

THETA_TOL = 1e-6      # north_star: combined theta within 1e-6 max-abs
SIG_RTOL = 1e-9       # north_star: Sig_inv within 1e-9 relative Frobenius


def relfro(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def fit_partitions(X, y, pid, K, fit_intercept=False):
    coefs, sigs = [], []
    for k in range(K):
        m = pid == k
        c, S, _ = onp.logistic_fit(X[m], y[m], fit_intercept=fit_intercept)
        coefs.append(c)
        sigs.append(S)
    return np.stack(coefs), np.stack(sigs)


def test_games_known_answer_values():
    """SURVEY.md App. C values (live shimmed-reference run) are what the golden file holds."""
    g = load_golden("games_expand_k5.npz")
    assert np.allclose(g["ref_beta_byOLS"][:3], [-1.143567655531, -0.045690361706, -0.915884347200], atol=1e-9)
    assert abs(np.linalg.norm(g["ref_Sig_inv_sum"]) - 1.638318745483e+03) < 1e-5
    assert abs(np.trace(g["ref_Sig_inv_sum"]) - 4.406652096412e+03) < 1e-5
    assert g["ref_lar_BIC"].shape[0] == 11 and int(np.argmin(g["ref_lar_BIC"])) == 9
    assert abs(g["ref_lar_BIC"][9] - 91.073913) < 1e-5
    assert g["ref_lar_beta"][9][1] == 0.0 and np.count_nonzero(g["ref_lar_beta"][9]) == 9


def test_oracle_games_partition_fit_and_combine():
    X, y, K, g = games_data()
    pid = np.arange(X.shape[0]) % K
    coefs, sigs = fit_partitions(X, y, pid, K)
    assert np.max(np.abs(coefs - g["ref_coef"])) < 1e-8
    for k in range(K):
        assert relfro(sigs[k], g["ref_Sig_inv"][k]) < SIG_RTOL
    theta, oneshot, S = onp.combine(coefs, sigs)
    assert np.max(np.abs(theta - g["ref_beta_byOLS"])) < THETA_TOL * 1e-2
    assert np.max(np.abs(oneshot - g["ref_beta_byONESHOT"])) < 1e-8
    assert relfro(S, g["ref_Sig_inv_sum"]) < SIG_RTOL


@pytest.mark.parametrize("name", ["sim_c1_n100000_p20_k16.npz", "sim_small_n6000_p7_k3.npz",
                                  "sim_icpt_n20000_p12_k4.npz", "sim_p50_n40000_k4.npz"])
def test_oracle_sim_fit_and_combine(name):
    g = load_golden(name)
    pid, y, X, K = sim_data(g)
    icpt = bool(g["in_fit_intercept"])
    coefs, sigs = fit_partitions(X, y, pid, K, fit_intercept=icpt)
    assert np.max(np.abs(coefs - g["ref_coef"])) < 1e-8
    assert max(relfro(sigs[k], g["ref_Sig_inv"][k]) for k in range(K)) < SIG_RTOL
    theta, oneshot, S = onp.combine(coefs, sigs)
    assert np.max(np.abs(theta - g["ref_beta_byOLS"])) < THETA_TOL * 1e-2
    assert np.max(np.abs(oneshot - g["ref_beta_byONESHOT"])) < 1e-8
    assert relfro(S, g["ref_Sig_inv_sum"]) < SIG_RTOL


def check_path(fit, g, prefix):
    rb = g[prefix + "beta"]
    assert fit["beta"].shape == rb.shape, (fit["beta"].shape, rb.shape)
    scale = max(1.0, np.max(np.abs(rb)))
    assert np.max(np.abs(fit["beta"] - rb)) < 1e-9 * scale
    assert np.array_equal(fit["beta"] != 0, rb != 0)  # same support on every knot
    assert np.allclose(fit["AIC"], g[prefix + "AIC"], rtol=1e-9, atol=1e-7)
    assert np.allclose(fit["BIC"], g[prefix + "BIC"], rtol=1e-9, atol=1e-7)
    assert int(np.argmin(fit["BIC"])) == int(np.argmin(g[prefix + "BIC"]))
    assert int(np.argmin(fit["AIC"])) == int(np.argmin(g[prefix + "AIC"]))


@pytest.mark.parametrize("typ", ["lar", "lasso"])
def test_oracle_lars_standalone_cases(typ):
    g = load_golden("lars_cases.npz")
    n_cases = int(g["n_cases"])
    assert n_cases >= 10
    n_drop_cases = 0
    for i in range(n_cases):
        key = f"c{i}_{typ}_"
        Sigma, b, n = g[key + "in_Sigma"], g[key + "in_b"], int(g[key + "in_n"])
        fit = onp.lars_lsa(Sigma, b, False, n, type=typ)
        check_path(fit, g, key + "ref_")
        n_drop_cases += int(g[key + "ref_beta"].shape[0] > Sigma.shape[0] + 1)
    if typ == "lasso":
        assert n_drop_cases >= 5  # the goldens do exercise the drop / downdate branch


@pytest.mark.parametrize("typ", ["lar", "lasso"])
def test_oracle_lars_on_dlsa_outputs(typ):
    for name in ["games_expand_k5.npz", "sim_c1_n100000_p20_k16.npz", "sim_p50_n40000_k4.npz"]:
        g = load_golden(name)
        n = int(g["in_shape"][0]) if "in_shape" in g else int(g["in_n"])
        fit = onp.lars_lsa(g["ref_Sig_inv_sum"], g["ref_beta_byOLS"], False, n, type=typ)
        check_path(fit, g, f"ref_{typ}_")


def test_oracle_lars_intercept_matches_profiled_problem():
    """intercept=True follows R (dlsa_alasso_func.R:115-125,243-245): equals the no-intercept path on
    the Schur complement, with beta0_k = (a12'b - a12'beta_k)/a11."""
    g = load_golden("sim_icpt_n20000_p12_k4.npz")
    S, th, n = g["ref_Sig_inv_sum"], g["ref_beta_byOLS"], int(g["in_n"])
    fit = onp.lars_lsa(S, th, True, n)
    a11, a12 = S[0, 0], S[1:, 0]
    sub = onp.lars_lsa(S[1:, 1:] - np.outer(a12, a12) / a11, th[1:], False, n)
    assert np.allclose(fit["beta"], sub["beta"], atol=1e-12)
    assert np.allclose(fit["beta0"], (a12 @ th[1:] - sub["beta"] @ a12) / a11, atol=1e-12)
    bA, bB, _ = onp.dlsa_select(S, th, n, fit_intercept=True)
    # last knot of the path = unpenalised solution: beta0 + theta[0] returns the OLS intercept
    assert abs((fit["beta0"][-1] + th[0]) - th[0]) < 1e-9
    assert bA.shape == (13,) and bB.shape == (13,)


def test_oracle_second_lars_oracle_sklearn():
    """Independent cross-check: sklearn lars_path_gram on the rescaled problem."""
    from sklearn.linear_model import lars_path_gram
    g = load_golden("sim_p50_n40000_k4.npz")
    S, th = g["ref_Sig_inv_sum"], g["ref_beta_byOLS"]
    D = np.abs(th)
    St = D[:, None] * S * D[None, :]
    s = np.sign(th)
    _, _, coefs = lars_path_gram(Xy=St @ s, Gram=St, n_samples=1, method="lar", alpha_min=0)
    fit = onp.lars_lsa(S, th, False, int(g["in_n"]))
    assert np.max(np.abs(coefs.T * D[None, :] - fit["beta"])) < 1e-9


def test_oracle_dummy_layout_and_zero_frame():
    g = load_golden("dummy_layout.npz")
    df = pd.DataFrame({"partition_id": 0.0, "y": g["in_y"], "znum": g["in_znum"], "anum": g["in_anum"],
                       "Month": g["in_Month"], "Carrier": g["in_Carrier"]})
    counts = {"Month": {int(k): int(v) for k, v in zip(g["ref_count_keys_Month"], g["ref_count_vals_Month"])},
              "Carrier": {str(k): int(v) for k, v in zip(g["ref_count_keys_Carrier"], g["ref_count_vals_Carrier"])}}
    info = onp.select_dummy_factors(counts, list(g["in_keep_top"]))
    assert info["factor_selected"]["Carrier"] == list(g["ref_selected_Carrier"])
    assert info["factor_dropped"]["Carrier"] == list(g["ref_dropped_Carrier"])
    assert info["factor_selected_names"]["Carrier"] == list(g["ref_selected_names_Carrier"])
    assert info["factor_selected_names"]["Month"] == list(g["ref_selected_names_Month"])
    base = list(g["in_baselines"])
    frame = onp.logistic_model_frame(df, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base)
    assert list(frame.columns) == list(g["ref_frame_columns"])
    assert list(frame.columns) == ['par_id', 'coef', 'Sig_invMcoef', 'intercept', 'anum', 'znum', 'Month_10', 'Month_11',
                                   'Month_12', 'Month_2', 'Month_3', 'Month_4', 'Month_5', 'Month_6', 'Month_7', 'Month_8',
                                   'Month_9', 'Carrier_A', 'Carrier_B', 'Carrier_C']
    ref = g["ref_frame"]
    got = frame.to_numpy(dtype=float)
    assert got.shape == ref.shape == (17, 20)
    assert np.array_equal(got[:, 0], ref[:, 0])
    assert np.max(np.abs(got[:, 1] - ref[:, 1])) < 1e-7
    assert relfro(got[:, 3:], ref[:, 3:]) < SIG_RTOL
    with pytest.warns(UserWarning, match="Dummies:"):
        zero = onp.logistic_model_frame(df[df["Month"] != 7], "y", fit_intercept=True, dummy_info=info,
                                        dummy_factors_baseline=base)
    assert list(zero.columns) == list(g["ref_zero_frame_columns"])
    assert np.array_equal(zero.to_numpy(dtype=float), g["ref_zero_frame"])
    assert "Month_7" in str(g["ref_zero_warning"][0])


def test_oracle_ols_block_identity():
    """OLS DLSA (R wlse_func.R:169-182): with equal sigma2 weights the combine equals the global OLS;
    Sig_inv_k = X'X / sigma2_k."""
    rng = np.random.default_rng(0)
    n, p, K = 6000, 9, 4
    X = rng.standard_normal((n, p))
    beta = np.array([3, 1.5, 0, 0, 2, 0, 0, 0, 1.0])
    y = X @ beta + rng.standard_normal(n)
    th, Sg, Sgt, s2 = zip(*[onp.ols_fit(X[k::K], y[k::K]) for k in range(K)])
    theta, _, S = onp.combine(np.stack(th), np.stack(Sg), np.stack(Sgt))
    for k in range(K):
        assert np.allclose(Sg[k] * s2[k], X[k::K].T @ X[k::K], rtol=1e-12)
        lst = np.linalg.lstsq(X[k::K], y[k::K], rcond=None)[0]
        assert np.allclose(th[k], lst, atol=1e-10)
    assert np.max(np.abs(theta - beta)) < 0.1


@pytest.mark.skipif(not ref_shim.available(), reason="/root/reference not present (GPU box)")
def test_oracle_against_live_reference():
    """Where the reference tree exists, re-run its functions live on a fresh draw (not just the goldens)."""
    pid, y, X, _ = onp.simulate_logistic(9000, 6, "systematic", 3, seed=77)
    df = pd.DataFrame(np.concatenate([pid[:, None], y[:, None], X], axis=1),
                      columns=["partition_id", "label"] + [f"x{i}" for i in range(6)])
    frames = [ref_shim.ref_logistic_model(df[df.partition_id == k], "label") for k in range(3)]
    mine = [onp.logistic_model_frame(df[df.partition_id == k], "label") for k in range(3)]
    for a, b in zip(frames, mine):
        assert list(a.columns) == list(b.columns)
        assert np.max(np.abs(a["coef"].to_numpy() - b["coef"].to_numpy())) < 1e-8
        assert relfro(b.iloc[:, 3:].to_numpy(), a.iloc[:, 3:].to_numpy()) < SIG_RTOL
    mr = ref_shim.ref_mapred(frames, 3)
    coefs = np.stack([f["coef"].to_numpy() for f in mine])
    sigs = np.stack([f.iloc[:, 3:].to_numpy() for f in mine])
    theta, oneshot, S = onp.combine(coefs, sigs)
    assert np.max(np.abs(theta - mr["beta_byOLS"].to_numpy())) < 1e-8
    for typ in ("lar", "lasso"):
        r = ref_shim.ref_lars_lsa(S, theta, False, 9000, type=typ)
        f = onp.lars_lsa(S, theta, False, 9000, type=typ)
        assert np.max(np.abs(r["beta"] - f["beta"])) < 1e-9


def test_poisson_oracle_against_sklearn_glm():
    """R's glm(y ~ x - 1, family = "poisson") (dlsa/R/wlse_func.R:155-168) cannot run here (no R): the oracle's Poisson
    Newton is cross-checked against scikit-learn's PoissonRegressor (unpenalised: the same MLE) and against the score
    equation.  Parity of the Poisson family is therefore UNPINNED with respect to the reference's own R output."""
    from sklearn.linear_model import PoissonRegressor
    rng = np.random.default_rng(17)
    n, p = 8000, 9
    X = rng.random((n, p)) - 0.5
    beta = np.r_[np.ones(4), np.zeros(p - 4)]
    for icpt in (False, True):
        eta = X @ beta + (0.7 if icpt else 0.0)
        y = rng.poisson(np.exp(eta)).astype(np.float64)
        coef, Sig, Sc = onp.poisson_fit(X, y, fit_intercept=icpt)
        Xd = np.concatenate([np.ones((n, 1)), X], axis=1) if icpt else X
        g, H, ll = onp.poisson_pass(Xd, y, coef)
        assert np.max(np.abs(np.linalg.solve(H, g))) < 1e-10          # score equation
        assert np.array_equal(Sig, H) and np.allclose(Sc, H @ coef)
        sk = PoissonRegressor(alpha=0.0, fit_intercept=icpt, tol=1e-12, max_iter=2000, solver="newton-cholesky").fit(X, y)
        ref = np.r_[sk.intercept_, sk.coef_] if icpt else sk.coef_
        assert np.max(np.abs(coef - ref)) < 1e-7


def _eval_golden_frames():
    """Inputs of tests/golden/eval_and_datainfo.npz rebuilt from their recipe (oracle/make_golden.py::golden_eval)."""
    import pandas as pd
    g = load_golden("eval_and_datainfo.npz")
    pid, y, X, _ = onp.simulate_logistic(int(g["a_in_n"]), int(g["a_in_p"]), "systematic", int(g["a_in_K"]),
                                         seed=int(g["a_in_seed"]))
    names = ["x" + str(i) for i in range(X.shape[1])]
    df = pd.DataFrame(np.concatenate([pid[:, None], y[:, None], X], axis=1), columns=["partition_id", "label"] + names)
    dfc = pd.DataFrame({"partition_id": np.zeros(len(g["c_in_y"])), "y": g["c_in_y"], "anum": g["c_in_anum"],
                        "znum": g["c_in_znum"], "Carrier": g["c_in_Carrier"]})
    return g, df, names, dfc


def test_oracle_loglik_and_standardised_fit_against_reference_goldens():
    """logistic_model_eval (dlsa/models.py:151-225) and logistic_model(..., data_info) (:99-101) as run by the
    reference: the oracle's log-likelihood and its standardised fit reproduce the frozen outputs."""
    g, df, names, dfc = _eval_golden_frames()
    K = int(g["a_in_K"])
    dinfo = {c: g["b_in_data_info"][j] for j, c in enumerate(names)}
    for k in range(K):
        part = df[df["partition_id"] == k]
        Xk, yk = part[names].to_numpy(), part["label"].to_numpy()
        assert np.allclose(onp.logistic_loglik(Xk, yk, g["a_in_par"]), g["a_ref_loglik"][k], rtol=1e-12)
        Xs = (Xk - g["b_in_data_info"][:, 1]) / g["b_in_data_info"][:, 2]
        Xs1 = np.concatenate([np.ones((Xs.shape[0], 1)), Xs], axis=1)
        assert np.allclose(onp.logistic_loglik(Xs1, yk, g["b_in_par"]), g["b_ref_loglik"][k], rtol=1e-12)
        fr = onp.logistic_model_frame(part, "label", fit_intercept=True, data_info=dinfo)
        assert list(fr.columns) == [str(c) for c in g["b_ref_fit_columns"]]
        assert np.max(np.abs(fr["coef"].to_numpy() - g["b_ref_fit_coef"][k])) < 1e-8
        assert relfro(fr.iloc[:, 3:].to_numpy(), g["b_ref_fit_Sig_inv"][k]) < 1e-9
    # dummy branch + intercept + standardisation (case c)
    info = onp.select_dummy_factors({"Carrier": dict(zip(*np.unique(dfc["Carrier"], return_counts=True)))}, [0.8])
    counts = dfc["Carrier"].value_counts()
    info = onp.select_dummy_factors({"Carrier": {k: int(v) for k, v in counts.items()}}, [float(g["c_in_keep_top"][0])])
    dinfo_c = {c: g["c_in_data_info"][j] for j, c in enumerate(["anum", "znum"])}
    fr = onp.logistic_model_frame(dfc, "y", fit_intercept=True, dummy_info=info,
                                  dummy_factors_baseline=[str(b) for b in g["c_in_baselines"]], data_info=dinfo_c)
    assert list(fr.columns) == [str(c) for c in g["c_ref_frame_columns"]]
    assert np.max(np.abs(fr.to_numpy(dtype=float)[:, 1] - g["c_ref_frame"][:, 1])) < 1e-8
    assert relfro(fr.to_numpy(dtype=float)[:, 3:], g["c_ref_frame"][:, 3:]) < 1e-9
