"""GPU tests of the reference-facing façade (dlsa.models / dlsa.dlsa / dlsa.lsa names) against the reference goldens."""
import numpy as np
import pandas as pd
import pytest
import torch

from conftest import games_data, load_golden, sim_data
from oracle import numpy_dlsa as onp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", autouse=True)
def need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def relfro(a, b):
    return np.linalg.norm(a - b) / np.linalg.norm(b)


def games_frame():
    X, y, K, g = games_data()
    df = pd.DataFrame(X, columns=[f"G{i}" for i in range(1, 11)])
    df["label"] = y
    df.insert(0, "partition_id", np.arange(len(df)) % K)
    return df, K, g


def test_reference_names_and_signatures():
    import inspect
    import dlsa
    from dlsa.models import logistic_model, logistic_model_eval, simulate_logistic
    from dlsa.dlsa import dlsa as dlsa_fn, dlsa_mapred, dlsa_mapreduce
    from dlsa.lsa import lars_lsa
    assert list(inspect.signature(logistic_model).parameters)[:6] == ["sample_df", "Y_name", "fit_intercept", "dummy_info",
                                                                      "dummy_factors_baseline", "data_info"]
    assert list(inspect.signature(dlsa_fn).parameters)[:4] == ["Sig_inv_", "beta_", "sample_size", "fit_intercept"]
    assert list(inspect.signature(lars_lsa).parameters) == ["Sigma0", "b0", "intercept", "n", "type", "eps", "max_steps"]
    assert inspect.signature(lars_lsa).parameters["type"].default == "lar"
    assert dlsa_mapreduce is dlsa_mapred
    assert list(inspect.signature(simulate_logistic).parameters)[:4] == ["sample_size", "p", "partition_method", "partition_num"]


def test_logistic_model_frames_mapred_dlsa_games():
    from dlsa.models import logistic_model
    from dlsa.dlsa import dlsa as dlsa_fn, dlsa_mapred
    from dlsa.lsa import lars_lsa
    df, K, g = games_frame()
    frames = [logistic_model(df[df.partition_id == k], "label") for k in range(K)]
    for k, f in enumerate(frames):
        assert list(f.columns) == list(g["ref_frame_columns"])
        assert f.shape == (10, 13) and list(f["par_id"]) == list(range(10))
        assert np.max(np.abs(f["coef"].to_numpy() - g["ref_coef"][k])) < 1e-8
        assert relfro(f.iloc[:, 3:].to_numpy(), g["ref_Sig_inv"][k]) < 1e-9
    mr = dlsa_mapred(pd.concat(frames, ignore_index=True))
    assert list(mr.columns) == list(g["ref_mapred_columns"])
    assert np.max(np.abs(mr["beta_byOLS"].to_numpy() - g["ref_beta_byOLS"])) < 1e-6
    assert np.max(np.abs(mr["beta_byONESHOT"].to_numpy() - g["ref_beta_byONESHOT"])) < 1e-9
    assert relfro(mr.iloc[:, 2:].to_numpy(), g["ref_Sig_inv_sum"]) < 1e-9
    assert np.allclose(dlsa_mapred(frames).to_numpy(), mr.to_numpy())  # list-of-frames input

    class FakeSpark:  # duck-typed Spark DataFrame (dlsa/dlsa.py:30-34,52)
        def __init__(self, pdf, n):
            self._pdf, self._n, self.rdd = pdf, n, self

        def toPandas(self):
            return self._pdf

        def getNumPartitions(self):
            return self._n
    mr2 = dlsa_mapred(FakeSpark(pd.concat(frames, ignore_index=True), 10))
    assert np.allclose(mr2["beta_byONESHOT"].to_numpy() * 2, mr["beta_byONESHOT"].to_numpy())
    out = dlsa_fn(Sig_inv_=mr.iloc[:, 2:], beta_=mr["beta_byOLS"], sample_size=len(df), fit_intercept=False)
    assert list(out.columns) == ["beta_byAIC", "beta_byBIC"]
    ib, ia = int(np.argmin(g["ref_lar_BIC"])), int(np.argmin(g["ref_lar_AIC"]))
    assert np.max(np.abs(out["beta_byBIC"].to_numpy() - g["ref_lar_beta"][ib])) < 1e-8
    assert np.max(np.abs(out["beta_byAIC"].to_numpy() - g["ref_lar_beta"][ia])) < 1e-8
    assert np.array_equal(out["beta_byBIC"].to_numpy() != 0, g["ref_lar_beta"][ib] != 0)
    fit = lars_lsa(np.matrix(mr.iloc[:, 2:].to_numpy()), mr["beta_byOLS"].to_numpy(), False, len(df), type="lasso")
    assert set(fit) == {"AIC", "BIC", "beta", "beta0"}
    assert np.allclose(fit["BIC"], g["ref_lasso_BIC"], rtol=1e-8, atol=1e-6)


def test_partitions_batched_map_equals_per_partition_calls():
    from dlsa.models import logistic_model, logistic_model_partitions
    g = load_golden("sim_small_n6000_p7_k3.npz")
    pid, y, X, K = sim_data(g)
    df = pd.DataFrame(np.concatenate([pid[:, None].astype(float), y[:, None], X], axis=1),
                      columns=["partition_id", "label"] + [f"x{i}" for i in range(X.shape[1])])
    stacked = logistic_model_partitions(df, "label")
    assert list(stacked.columns) == list(g["ref_frame_columns"])
    for k in range(K):
        one = logistic_model(df[df.partition_id == k], "label")
        blk = stacked.iloc[k * 7:(k + 1) * 7]
        assert np.allclose(blk.to_numpy(), one.to_numpy(), rtol=1e-12, atol=1e-12)
        assert np.max(np.abs(blk["coef"].to_numpy() - g["ref_coef"][k])) < 1e-8


def test_dummy_layout_and_zero_frame_match_reference():
    from dlsa.models import logistic_model
    from dlsa import dummies
    g = load_golden("dummy_layout.npz")
    df = pd.DataFrame({"partition_id": 0.0, "y": g["in_y"], "znum": g["in_znum"], "anum": g["in_anum"],
                       "Month": g["in_Month"], "Carrier": g["in_Carrier"]})
    info = dummies.select_dummy_factors(dummies.dummy_factors_counts(df, ["Month", "Carrier"]), list(g["in_keep_top"]))
    base = list(g["in_baselines"])
    frame = logistic_model(df, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base)
    assert list(frame.columns) == list(g["ref_frame_columns"])  # bit-exact layout
    ref = g["ref_frame"]
    got = frame.to_numpy(dtype=float)
    assert got.shape == (17, 20) and np.array_equal(got[:, 0], ref[:, 0])
    assert np.max(np.abs(got[:, 1] - ref[:, 1])) < 1e-7
    assert relfro(got[:, 3:], ref[:, 3:]) < 1e-9
    with pytest.warns(UserWarning, match="Dummies:"):
        zero = logistic_model(df[df["Month"] != 7], "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base)
    assert list(zero.columns) == list(g["ref_zero_frame_columns"])
    assert np.array_equal(zero.to_numpy(dtype=float), g["ref_zero_frame"])
    # standardisation with describe()-style data_info (rows 1, 2 = mean, std; dlsa/models.py:99-101)
    desc = df[["anum", "znum"]].describe()
    data_info = {c: desc[c].to_numpy() for c in ["anum", "znum"]}
    fs = logistic_model(df, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base, data_info=data_info)
    ref_s = onp.logistic_model_frame(df, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base,
                                     data_info=data_info)
    assert np.max(np.abs(fs["coef"].to_numpy() - ref_s["coef"].to_numpy())) < 1e-8
    assert relfro(fs.iloc[:, 3:].to_numpy(), ref_s.iloc[:, 3:].to_numpy()) < 1e-9


def test_intercept_dlsa_follows_r_spec():
    from dlsa.dlsa import dlsa as dlsa_fn
    g = load_golden("sim_icpt_n20000_p12_k4.npz")
    S, th, n = g["ref_Sig_inv_sum"], g["ref_beta_byOLS"], int(g["in_n"])
    out = dlsa_fn(pd.DataFrame(S), pd.Series(th), n, fit_intercept=True)
    bA, bB, _ = onp.dlsa_select(S, th, n, fit_intercept=True)
    assert out.shape == (13, 2)
    assert np.max(np.abs(out["beta_byAIC"].to_numpy() - bA)) < 1e-9
    assert np.max(np.abs(out["beta_byBIC"].to_numpy() - bB)) < 1e-9


def test_model_eval_and_simulate_and_linear():
    from dlsa.models import logistic_model_eval, simulate_logistic, linear_model
    from dlsa.model_eval import logistic_model_eval_sdf
    df, K, g = games_frame()
    par = pd.DataFrame({"beta_byOLS": g["ref_beta_byOLS"], "beta_byONESHOT": g["ref_beta_byONESHOT"],
                        "beta_byBIC": g["ref_lar_beta"][9]})
    X, y = df.iloc[:, 1:11].to_numpy(), df["label"].to_numpy()
    ref = onp.logistic_loglik(X, y, par.to_numpy())
    one = logistic_model_eval(df, "label", par)
    assert list(one.columns) == list(par.columns) and np.allclose(one.to_numpy()[0], ref, rtol=1e-10)
    tot = logistic_model_eval_sdf(df, par, False, "label", [], [], [])
    assert np.allclose(tot.to_numpy()[0], ref, rtol=1e-10)
    sim = simulate_logistic(20000, 10, "systematic", 4, seed=3)
    assert list(sim.columns) == ["partition_id", "label"] + [f"x{i}" for i in range(10)]
    assert sim.shape == (20000, 12) and set(np.unique(sim["label"])) == {0.0, 1.0}
    assert np.array_equal(sim["partition_id"].to_numpy(), np.arange(20000) % 4)
    assert abs(sim.iloc[:, 2:].to_numpy().mean()) < 5e-3 and sim.iloc[:, 2:].to_numpy().min() >= -0.5
    with pytest.raises(Exception, match="No such partition method"):
        simulate_logistic(10, 2, "random", 2)
    rng = np.random.default_rng(0)
    Xl = rng.standard_normal((4000, 6))
    yl = Xl @ np.array([3, 1.5, 0, 0, 2, 0.0]) + rng.standard_normal(4000)
    dfl = pd.DataFrame(Xl, columns=[f"x{i}" for i in range(6)])
    dfl["yy"] = yl
    dfl.insert(0, "partition_id", 0)
    fl = linear_model(dfl, "yy")
    th, Sg, Sgt, _ = onp.ols_fit(Xl, yl)
    assert np.max(np.abs(fl["coef"].to_numpy() - th)) < 1e-10 and relfro(fl.iloc[:, 3:].to_numpy(), Sg) < 1e-11


def test_single_class_partition_raises_like_sklearn():
    from dlsa.models import logistic_model
    df = pd.DataFrame({"partition_id": 0, "label": 1.0, "x0": [1.0, 2.0, 3.0, 4.0], "x1": [0.2, -0.1, 0.3, 0.0]})
    with pytest.raises(ValueError):
        logistic_model(df, "label", config=__import__("dlsa_b200").EngineConfig(max_iter=10))


def test_describe_matches_spark_describe_semantics():
    """projects/logistic_dlsa.py:279-283: data_info = describe() -> rows count/mean/stddev(n-1)/min/max; logistic_model
    reads rows 1 and 2 (dlsa/models.py:99-101).  Columns with a large mean and a small spread (Year) must not lose
    digits; standardising with the device statistics must equal standardising with pandas' statistics."""
    import dlsa_b200
    from dlsa.utils_spark import describe
    rng = np.random.default_rng(4)
    n = 123_457
    df = pd.DataFrame({"Year": 2000.0 + rng.integers(0, 9, n), "Distance": rng.gamma(2.0, 400.0, n),
                       "DepTime": rng.integers(0, 2400, n).astype(float), "z": rng.standard_normal(n) * 1e-3 + 7e5,
                       "const": np.full(n, 3.25), "Carrier": rng.choice(list("ABC"), n)})
    info = describe(df)
    assert list(info["summary"]) == ["count", "mean", "stddev", "min", "max"]
    num = ["Year", "Distance", "DepTime", "z", "const"]
    assert list(info.columns) == ["summary"] + num           # non-numeric columns are left out
    ref = df[num].describe()
    for c in num:
        assert int(info[c][0]) == n
        assert abs(float(info[c][1]) - ref[c]["mean"]) <= 1e-12 * abs(ref[c]["mean"])
        assert abs(float(info[c][2]) - ref[c]["std"]) <= 1e-10 * ref[c]["std"] + 1e-300
        assert float(info[c][3]) == ref[c]["min"] and float(info[c][4]) == ref[c]["max"]
    # tensor entry, wide and ragged shapes, run-to-run reproducible
    for px in (1, 31, 50, 182):
        X = torch.as_tensor(rng.standard_normal((5003, px)) * 3 + 10, device="cuda")
        st = dlsa_b200.column_stats(X)
        assert torch.equal(st, dlsa_b200.column_stats(X))
        Xh = X.cpu().numpy()
        assert np.allclose(st[1].cpu().numpy(), Xh.mean(0), rtol=1e-13)
        assert np.allclose(st[2].cpu().numpy(), Xh.std(0, ddof=1), rtol=1e-12)
        assert np.array_equal(st[3].cpu().numpy(), Xh.min(0)) and np.array_equal(st[4].cpu().numpy(), Xh.max(0))
    # strided view (ldx > px) and a single row
    Xs = torch.as_tensor(rng.standard_normal((777, 12)), device="cuda")[:, :7]
    assert np.allclose(dlsa_b200.column_stats(Xs)[1].cpu().numpy(), Xs.cpu().numpy().mean(0), rtol=1e-12, atol=1e-15)
    one = dlsa_b200.column_stats(torch.ones(1, 3, dtype=torch.float64, device="cuda"))
    assert one[0].tolist() == [1.0, 1.0, 1.0] and torch.isnan(one[2]).all()  # Spark: stddev of one row is NaN/null


def test_poisson_model_frame():
    """models.poisson_model: the frame layout of logistic_model with the estimates of R's glm(family="poisson")."""
    from dlsa.models import poisson_model
    rng = np.random.default_rng(12)
    n, p = 4000, 5
    X = rng.random((n, p)) - 0.5
    y = rng.poisson(np.exp(X @ np.array([1.0, -1.0, 0.5, 0.0, 0.0]) + 0.3)).astype(np.float64)
    df = pd.DataFrame(X, columns=[f"x{i}" for i in range(p)])
    df.insert(0, "label", y)
    df.insert(0, "partition_id", 0)
    out = poisson_model(df, "label", fit_intercept=True)
    c0, S0, Sc0 = onp.poisson_fit(X, y, fit_intercept=True)
    assert list(out.columns) == ["par_id", "coef", "Sig_invMcoef", "intercept"] + [f"x{i}" for i in range(p)]
    assert np.max(np.abs(out["coef"].to_numpy() - c0)) < 1e-9
    assert np.linalg.norm(out.iloc[:, 3:].to_numpy() - S0) / np.linalg.norm(S0) < 1e-9


def test_model_eval_and_data_info_against_reference_goldens():
    """f-1: logistic_model_eval / logistic_model_eval_sdf (dlsa/models.py:151-225, dlsa/model_eval.py:10-42) and the
    standardised fit logistic_model(..., data_info) (:99-101) against outputs of the REFERENCE's own functions
    (tests/golden/eval_and_datainfo.npz, minted by oracle/make_golden.py::golden_eval)."""
    from dlsa.models import logistic_model, logistic_model_eval
    from dlsa.model_eval import logistic_model_eval_sdf
    from dlsa import dummies
    from test_oracle_pinned import _eval_golden_frames
    g, df, names, dfc = _eval_golden_frames()
    K = int(g["a_in_K"])
    par_a = pd.DataFrame(g["a_in_par"], columns=[str(c) for c in g["a_par_columns"]])
    par_b = pd.DataFrame(g["b_in_par"], columns=[str(c) for c in g["b_par_columns"]])
    dinfo = {c: g["b_in_data_info"][j] for j, c in enumerate(names)}
    for k in range(K):
        part = df[df["partition_id"] == k]
        la = logistic_model_eval(part, "label", par_a)
        assert list(la.columns) == list(par_a.columns)
        assert np.allclose(la.to_numpy()[0], g["a_ref_loglik"][k], rtol=1e-11)
        lb = logistic_model_eval(part, "label", par_b, fit_intercept=True, data_info=dinfo)
        assert np.allclose(lb.to_numpy()[0], g["b_ref_loglik"][k], rtol=1e-11)
        fr = logistic_model(part, "label", fit_intercept=True, data_info=dinfo)
        assert list(fr.columns) == [str(c) for c in g["b_ref_fit_columns"]]
        assert np.max(np.abs(fr["coef"].to_numpy() - g["b_ref_fit_coef"][k])) < 1e-8
        assert relfro(fr.iloc[:, 3:].to_numpy(), g["b_ref_fit_Sig_inv"][k]) < 1e-9
        assert np.max(np.abs(fr["Sig_invMcoef"].to_numpy() - g["b_ref_fit_Sig_invMcoef"][k])) < 1e-6
    tot = logistic_model_eval_sdf(df, par_a, False, "label", [], [], [])     # the global value = sum over partitions
    assert np.allclose(tot.to_numpy()[0], g["a_ref_loglik"].sum(0), rtol=1e-11)
    # dummy branch + intercept + standardisation, both device routes behind the facade
    info = dummies.select_dummy_factors(dummies.dummy_factors_counts(dfc, ["Carrier"]), [float(g["c_in_keep_top"][0])])
    base = [str(b) for b in g["c_in_baselines"]]
    dinfo_c = {c: g["c_in_data_info"][j] for j, c in enumerate(["anum", "znum"])}
    fr = logistic_model(dfc, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base, data_info=dinfo_c)
    assert list(fr.columns) == [str(c) for c in g["c_ref_frame_columns"]]
    assert np.max(np.abs(fr["coef"].to_numpy() - g["c_ref_frame"][:, 1])) < 1e-8
    assert relfro(fr.iloc[:, 3:].to_numpy(), g["c_ref_frame"][:, 3:]) < 1e-9
    par_c = pd.DataFrame(g["c_in_par"], columns=[str(c) for c in g["c_par_columns"]])
    lc = logistic_model_eval(dfc, "y", par_c, fit_intercept=True, dummy_info=info, dummy_factors_baseline=base,
                             data_info=dinfo_c)
    assert np.allclose(lc.to_numpy()[0], g["c_ref_loglik"], rtol=1e-10)


def test_ingest_csv_matches_the_pandas_facade(tmp_path):
    """f-3: CSV -> Arrow -> device codes / numerics (ingest.ingest + ingest.fit) gives the frames, the dummy_info and the
    describe() statistics the pandas route gives (dlsa.dummies.* + describe + logistic_model_partitions); the label rule
    (ArrDelay > 0), dropna and partition_id = row % K follow projects/logistic_dlsa.py:218-237."""
    from dlsa_b200 import dummies, ingest, models
    rng = np.random.default_rng(44)
    n = 9000
    df = pd.DataFrame({"Year": rng.integers(1987, 2009, n).astype(float), "Distance": rng.gamma(2.0, 400.0, n),
                       "DepTime": rng.uniform(0, 2400, n),
                       "Month": rng.integers(1, 13, n),
                       "UniqueCarrier": rng.choice(list("ABCDEFGH"), size=n, p=np.array([30, 20, 15, 12, 10, 6, 4, 3]) / 100)})
    eta = -0.2 + 0.002 * (df["Distance"] - 800) / 4 + 0.3 * (df["UniqueCarrier"] == "B") - 0.2 * (df["Month"] == 7)
    df["ArrDelay"] = np.where(rng.random(n) < 1 / (1 + np.exp(-eta)), rng.uniform(1, 90, n), -rng.uniform(0, 30, n))
    df.loc[rng.choice(n, 37, replace=False), "DepTime"] = np.nan          # rows dropna() must remove
    path = tmp_path / "airline_like.csv"
    df.to_csv(path, index=False)
    usecols_x = ["Year", "Distance", "DepTime", "Month", "UniqueCarrier"]
    des = ingest.ingest(str(path), usecols_x, "ArrDelay", dummy_columns=["Month", "UniqueCarrier"], dummy_keep_top=[1, 0.8],
                        dummy_factors_baseline=["Month_1", "UniqueCarrier_000_OTHERS"], fit_intercept=True,
                        partition_num=3)
    clean = df.dropna().reset_index(drop=True)
    assert des.n_rows_read == n and des.y.numel() == clean.shape[0]
    # the pandas route on the same data
    ref = clean.copy()
    ref["ArrDelay"] = (ref["ArrDelay"] > 0).astype(float)
    ref.insert(0, "partition_id", np.arange(ref.shape[0]) % 3)
    info = dummies.select_dummy_factors(dummies.dummy_factors_counts(ref, ["Month", "UniqueCarrier"]), [1, 0.8])
    assert {k: sorted(map(str, v)) for k, v in des.dummy_info["factor_selected"].items()} == \
           {k: sorted(map(str, v)) for k, v in info["factor_selected"].items()}
    assert sorted(des.dummy_info["factor_selected_names"]["UniqueCarrier"]) == sorted(info["factor_selected_names"]["UniqueCarrier"])
    desc = ref[["DepTime", "Distance", "Year"]].describe()
    for c in ["DepTime", "Distance", "Year"]:
        assert np.allclose(des.data_info[c][:3], desc[c].to_numpy()[:3], rtol=1e-12)
    data_info = {c: desc[c].to_numpy() for c in ["DepTime", "Distance", "Year"]}
    want = models.logistic_model_partitions(ref, "ArrDelay", fit_intercept=True, dummy_info=info,
                                            dummy_factors_baseline=["Month_1", "UniqueCarrier_000_OTHERS"],
                                            data_info=data_info)
    got = ingest.fit(des, fit_intercept=True)
    assert list(got.columns) == list(want.columns) and got.shape == want.shape
    assert np.max(np.abs(got["coef"].to_numpy() - want["coef"].to_numpy())) < 1e-8
    assert relfro(got.iloc[:, 3:].to_numpy(), want.iloc[:, 3:].to_numpy()) < 1e-9


def test_spark_adapter_udf_and_schema_on_a_fake_grouped_frame():
    """f-4: the applyInPandas adapter — schema in convert_schema's order, int par_id, one call per partition group —
    driven by a stand-in for ``data_sdf.groupby("partition_id").applyInPandas`` (Spark itself is absent here)."""
    from dlsa_b200 import spark_adapter
    from dlsa.dlsa import dlsa_mapred
    df, K, g = games_frame()
    usecols_x = [f"G{i}" for i in range(1, 11)]
    udf, schema = spark_adapter.make_logistic_model_udf(usecols_x, "label")
    assert schema.startswith("`par_id` int, `coef` double, `Sig_invMcoef` double, `G1` double")

    class FakeGrouped:
        def __init__(self, pdf, key):
            self.pdf, self.key = pdf, key

        def applyInPandas(self, fn, schema):
            names = [f.split("`")[1] for f in schema.split(", ")]
            outs = [fn(part.reset_index(drop=True)) for _, part in self.pdf.groupby(self.key)]
            assert all(list(o.columns) == names for o in outs)
            return pd.concat(outs, ignore_index=True)

    class FakeSdf:
        def __init__(self, pdf):
            self.pdf = pdf

        def groupby(self, key):
            return FakeGrouped(self.pdf, key)

    mapped = spark_adapter.map_on_executors(FakeSdf(df), usecols_x, "label")
    assert mapped["par_id"].dtype == np.int32 and mapped.shape == (K * 10, 13)
    mr = dlsa_mapred(mapped)
    assert np.max(np.abs(mr["beta_byOLS"].to_numpy() - g["ref_beta_byOLS"])) < 1e-8
    batched = spark_adapter.map_on_driver(df, "label")
    assert np.max(np.abs(batched["coef"].to_numpy() - mapped["coef"].to_numpy())) < 1e-10
