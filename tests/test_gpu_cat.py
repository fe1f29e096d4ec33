"""GPU parity of the compact categorical pass K1c (dlsa_irls_pass_cat_f64) — the dummy branch of logistic_model
(dlsa/models.py:56-91) computed without materialising the one-hot columns — against the oracle and against the dense
CUDA pass on the expanded matrix."""
import numpy as np
import pytest
import torch

from oracle import numpy_dlsa as onp

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dlsa_b200 import engine
    return engine


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


def relfro(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_case(rng, sizes, n_num, n_levels, baseline_share=0.15, zipf=True):
    off = np.zeros(len(sizes) + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    n = int(off[-1])
    numeric = rng.standard_normal((n, n_num))
    codes = np.empty((n, len(n_levels)), dtype=np.int32)
    for c, L in enumerate(n_levels):
        w = 1.0 / np.arange(1, L + 1) if zipf else np.ones(L)
        lv = rng.choice(L, size=n, p=w / w.sum())
        lv[rng.random(n) < baseline_share] = -1
        codes[:, c] = lv
    y = (rng.random(n) < 0.4).astype(np.float64)
    return off, numeric, codes, y


CASES = [
    # (n_num, n_levels, intercept, standardise)
    (7, [11, 6, 7, 75, 75], True, False),     # the airline shape: p = 182
    (7, [11, 6, 7, 75, 75], False, True),
    (2, [3, 4], True, False),                 # the dummy-layout KAT shape class
    (0, [5], False, False),                   # pure categorical, one column
    (0, [9, 17, 2], True, False),
    (3, [8, 8, 8, 8, 8, 8, 8, 8], True, True),  # 8 categorical columns, tiles aligned with the blocks
    (5, [1, 1, 30, 100], False, False),       # 132 levels -> 17 tiles
    (6, [40, 60, 90], True, False),           # 190 levels -> 24 tiles
    (7, [20] * 8, True, True),                # 8 scattered categoricals: more 9-column runs than half-warps (dense packing)
    (7, [5, 24, 24, 24, 24, 24, 24, 24], False, False),   # one DMMA tile + 7 scattered
    (4, [16, 9, 16, 16], True, False),        # 6 DMMA tiles used up by the three 16-level columns; the 9-level one scatters
    (1, [1, 8, 2, 17], True, True),           # tiles with 1, 8 and 2 live levels
]


@pytest.mark.parametrize("n_num,n_levels,icpt,standardise", CASES)
def test_cat_pass_matches_oracle(eng, n_num, n_levels, icpt, standardise):
    rng = np.random.default_rng(n_num * 31 + len(n_levels) * 7 + icpt)
    sizes = [0, 1, 31, 32, 33, 257, 1000, 4097, 64, 5]
    off, numeric, codes, y = make_case(rng, sizes, n_num, n_levels)
    p = int(icpt) + n_num + sum(n_levels)
    beta = rng.standard_normal((len(sizes), p)) * 0.4
    mean = std = None
    if standardise:
        mean, std = rng.standard_normal(n_num) * 0.3, 0.5 + rng.random(n_num)
    g, H, ll = eng.irls_pass_cat(dev(numeric) if n_num else None, dev(codes, torch.int32), n_levels, dev(y),
                                 dev(off, torch.int64), dev(beta), fit_intercept=icpt,
                                 mean=None if mean is None else dev(mean), std=None if std is None else dev(std))
    g, H, ll = g.cpu().numpy(), H.cpu().numpy(), ll.cpu().numpy()
    for k, nk in enumerate(sizes):
        sl = slice(off[k], off[k + 1])
        g0, H0, ll0 = onp.irls_pass_cat(numeric[sl], codes[sl], n_levels, y[sl], beta[k], icpt, mean, std)
        assert np.allclose(g[k], g0, rtol=1e-10, atol=1e-9), (k, nk, np.abs(g[k] - g0).max())
        assert np.allclose(H[k], H0, rtol=1e-10, atol=1e-10), (k, nk, np.abs(H[k] - H0).max())
        assert abs(ll[k] - ll0) <= 1e-11 * max(1.0, abs(ll0))
        assert np.array_equal(H[k], H[k].T)


def test_cat_pass_equals_dense_cuda_pass_on_expanded_matrix(eng):
    """K1c against K1 (tiled dense kernel) fed with dlsa_onehot_expand_f64's matrix: big sliced partitions, skipped
    (inactive) partitions left untouched, bitwise run-to-run reproducibility."""
    from dlsa_b200._lib import lib, check
    rng = np.random.default_rng(3)
    n_levels = [11, 6, 7, 75, 75]
    sizes = [150_000, 90_001, 40_000]
    off, numeric, codes, y = make_case(rng, sizes, 7, n_levels, baseline_share=0.1)
    n, K = int(off[-1]), len(sizes)
    px = 7 + sum(n_levels)
    p = px + 1
    beta = rng.standard_normal((K, p)) * 0.3
    num_d, codes_d, y_d, off_d, beta_d = dev(numeric), dev(codes, torch.int32), dev(y), dev(off, torch.int64), dev(beta)
    X = torch.zeros(n, px, dtype=torch.float64, device="cuda")
    base = np.concatenate([[7], 7 + np.cumsum(n_levels)[:-1]]).astype(np.int32)
    check(lib.dlsa_onehot_expand_f64(num_d.data_ptr(), 7, None, None, codes_d.data_ptr(), len(n_levels),
                                     dev(base, torch.int32).data_ptr(), n, X.data_ptr(), X.stride(0),
                                     torch.cuda.current_stream().cuda_stream), "onehot")
    g0, H0, ll0 = eng.irls_pass(X, y_d, off_d, beta_d, fit_intercept=True)
    active = torch.tensor([1, 0, 1], dtype=torch.int32, device="cuda")
    out = (torch.full((K, p), 7.0, dtype=torch.float64, device="cuda"),
           torch.full((K, p, p), 7.0, dtype=torch.float64, device="cuda"),
           torch.full((K,), 7.0, dtype=torch.float64, device="cuda"))
    g1, H1, ll1 = eng.irls_pass_cat(num_d, codes_d, n_levels, y_d, off_d, beta_d, fit_intercept=True, active=active, out=out)
    for k in (0, 2):
        assert relfro(H1[k].cpu().numpy(), H0[k].cpu().numpy()) < 1e-12
        assert np.allclose(g1[k].cpu().numpy(), g0[k].cpu().numpy(), rtol=1e-10, atol=1e-8)
        assert abs(float(ll1[k] - ll0[k])) < 1e-9 * abs(float(ll0[k]))
    assert torch.all(H1[1] == 7.0) and torch.all(g1[1] == 7.0) and float(ll1[1]) == 7.0
    g2, H2, ll2 = eng.irls_pass_cat(num_d, codes_d, n_levels, y_d, off_d, beta_d, fit_intercept=True)
    g3, H3, ll3 = eng.irls_pass_cat(num_d, codes_d, n_levels, y_d, off_d, beta_d, fit_intercept=True)
    assert torch.equal(H2, H3) and torch.equal(g2, g3) and torch.equal(ll2, ll3)
    assert torch.equal(H2[0], H1[0])


def test_airline_generator_is_counter_based_and_matches_host(eng):
    """The compact airline-shaped stream (BASELINE configs[2] input) is keyed by (seed, global row): any sharding sees
    the same rows, and the NumPy restatement in oracle/synth_host.py reproduces it."""
    from dlsa_b200 import synth
    from oracle import synth_host
    n, seed, row0 = 20_000, 5, 987_654_321
    num, codes, n_levels, y = synth.airline_shaped_device(n, seed, "cuda", row0=row0)
    assert n_levels == [11, 6, 7, 75, 75]
    b0, bnum, blev = synth.airline_truth()
    hn, hc, eta, u = synth_host.airline_shaped_host(n, seed, synth.AIRLINE_LEVELS, synth.airline_cdfs(), b0, bnum, blev,
                                                    row0=row0)
    assert np.array_equal(codes.cpu().numpy(), hc)                      # integer work: bit-exact
    assert np.max(np.abs(num.cpu().numpy() - hn)) < 1e-12               # Box-Muller: log / cospi differ by ulps
    yh = (u < 1 / (1 + np.exp(-eta))).astype(np.float64)
    assert np.mean(y.cpu().numpy() != yh) < 1e-3                        # identical up to rounding at the threshold
    share0 = float((codes[:, 3] == -1).double().mean())                # Zipf(1) over 76 levels: baseline share 1/H_76
    assert abs(share0 - 1.0 / np.sum(1.0 / np.arange(1, 77))) < 0.02
    a = synth.airline_shaped_device(7_000, seed, "cuda", row0=row0)
    b = synth.airline_shaped_device(13_000, seed, "cuda", row0=row0 + 7_000)
    assert torch.equal(torch.cat([a[0], b[0]]), num) and torch.equal(torch.cat([a[1], b[1]]), codes)
    assert torch.equal(torch.cat([a[3], b[3]]), y)


def test_cat_fit_matches_oracle_fit_and_selection(eng):
    """Whole map -> combine -> LARS on the compact airline-shaped design vs the oracle run on the expanded matrix."""
    from dlsa_b200 import synth
    n, K = 60_000, 4
    numeric, codes, n_levels, y = synth.airline_shaped_device(n, 5, torch.device("cuda"))
    # keep the level space small enough that every level is populated in every partition
    n_levels = [11, 6, 7, 9, 9]
    codes = codes.clone()
    for c, L in enumerate(n_levels):
        codes[:, c] = torch.where(codes[:, c] >= L, torch.full_like(codes[:, c], -1), codes[:, c])
    off = synth.equal_partition_offsets(n, K, torch.device("cuda"))
    fits = eng.fit_logistic_partitions_cat(numeric, codes, n_levels, y, off, fit_intercept=True)
    assert int(fits.status.abs().sum()) == 0
    comb = eng.combine_partitions(fits)
    path = eng.lars_path(comb.Sig_inv_sum, comb.beta_byOLS, True, float(n))
    Xh = onp.expand_onehot(numeric.cpu().numpy(), codes.cpu().numpy(), n_levels)
    yh, offh = y.cpu().numpy(), off.cpu().numpy()
    coefs, sigs = zip(*[onp.logistic_fit(Xh[offh[k]:offh[k + 1]], yh[offh[k]:offh[k + 1]], fit_intercept=True)[:2]
                        for k in range(K)])
    for k in range(K):
        assert np.max(np.abs(fits.coef[k].cpu().numpy() - coefs[k])) < 1e-8
        assert relfro(fits.Sig_inv[k].cpu().numpy(), sigs[k]) < 1e-9
    theta, oneshot, S = onp.combine(np.stack(coefs), np.stack(sigs))
    assert np.max(np.abs(comb.beta_byOLS.cpu().numpy() - theta)) < 1e-6
    assert relfro(comb.Sig_inv_sum.cpu().numpy(), S) < 1e-9
    ref = onp.lars_lsa(S, theta, True, float(n))
    assert path.argmin_BIC == int(np.argmin(ref["BIC"])) and path.argmin_AIC == int(np.argmin(ref["AIC"]))
    assert np.array_equal(path.beta[path.argmin_BIC].cpu().numpy() != 0, ref["beta"][path.argmin_BIC] != 0)


def test_cat_pass_rejects_unsupported_shapes(eng):
    from dlsa_b200._lib import DlsaLibraryError
    n = 64
    codes = torch.zeros(n, 1, dtype=torch.int32, device="cuda")
    y = torch.zeros(n, dtype=torch.float64, device="cuda")
    off = torch.tensor([0, n], dtype=torch.int64, device="cuda")
    with pytest.raises(DlsaLibraryError):  # 8 numerics: z = [1 | numerics] no longer fits one tile
        eng.irls_pass_cat(torch.zeros(n, 8, dtype=torch.float64, device="cuda"), codes, [3], y, off,
                          torch.zeros(1, 11, dtype=torch.float64, device="cuda"))
    with pytest.raises(DlsaLibraryError):  # 200 levels
        eng.irls_pass_cat(None, codes, [200], y, off, torch.zeros(1, 200, dtype=torch.float64, device="cuda"))
    assert not eng.cat_supported(8, [3]) and not eng.cat_supported(2, [100, 100]) and eng.cat_supported(7, [11, 6, 7, 75, 75])


def test_facade_compact_route_equals_dense_route(eng, monkeypatch):
    """logistic_model / logistic_model_partitions with dummy_info: the compact K1c route and the dense route (one-hot
    expansion + K1) give the same frames — same columns, same numbers within solver tolerance."""
    import pandas as pd
    from dlsa_b200 import dummies, models
    rng = np.random.default_rng(9)
    n = 24_000
    month = rng.integers(1, 13, n)
    carrier = rng.choice(list("ABCDEFGH"), size=n, p=np.array([30, 20, 15, 12, 10, 6, 4, 3]) / 100.0)
    znum, anum = rng.standard_normal(n) * 2 + 1, rng.standard_normal(n)
    eta = -0.3 + 0.5 * anum - 0.2 * (znum - 1) / 2 + 0.2 * (month % 3) + 0.3 * (carrier == "B")
    yv = (rng.random(n) < 1 / (1 + np.exp(-eta))).astype(float)
    df = pd.DataFrame({"partition_id": (np.arange(n) % 3).astype(float), "y": yv, "znum": znum, "anum": anum,
                       "Month": month, "Carrier": carrier})
    info = dummies.select_dummy_factors(dummies.dummy_factors_counts(df, ["Month", "Carrier"]), [1, 0.8])
    base = ["Month_1", "Carrier_000_OTHERS"]
    desc = df[["anum", "znum"]].describe()
    data_info = {c: desc[c].to_numpy() for c in ["anum", "znum"]}
    calls = {"cat": 0}
    real = eng.fit_logistic_partitions_cat

    def spy(*a, **k):
        calls["cat"] += 1
        return real(*a, **k)

    monkeypatch.setattr(eng, "fit_logistic_partitions_cat", spy)
    compact = models.logistic_model_partitions(df, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base,
                                               data_info=data_info)
    one = models.logistic_model(df[df["partition_id"] == 1.0], "y", fit_intercept=True, dummy_info=info,
                                dummy_factors_baseline=base, data_info=data_info)
    assert calls["cat"] == 2
    monkeypatch.setattr(eng, "cat_supported", lambda *a, **k: False)
    dense = models.logistic_model_partitions(df, "y", fit_intercept=True, dummy_info=info, dummy_factors_baseline=base,
                                             data_info=data_info)
    assert calls["cat"] == 2
    assert list(compact.columns) == list(dense.columns) and compact.shape == dense.shape
    assert np.max(np.abs(compact["coef"].to_numpy() - dense["coef"].to_numpy())) < 1e-8
    assert relfro(compact.iloc[:, 3:].to_numpy(), dense.iloc[:, 3:].to_numpy()) < 1e-10
    p = one.shape[0]
    assert np.max(np.abs(one["coef"].to_numpy() - compact["coef"].to_numpy()[p:2 * p])) < 1e-8


def test_host_entry_for_the_compact_design_matches_the_device_entry(eng):
    """dlsa_fit_host_cat: host-resident compact design (pinned or pageable), partition blocks copied on a copy stream while
    earlier blocks are fitted — same estimates as fitting the device-resident design in one batch."""
    from dlsa_b200 import synth
    n, K = 400_000, 80
    num, codes, n_levels, y = synth.airline_shaped_device(n, 3, torch.device("cuda"))
    n_levels = [11, 6, 7, 12, 12]
    codes = codes.clone()
    for c, L in enumerate(n_levels):
        codes[:, c] = torch.where(codes[:, c] >= L, torch.full_like(codes[:, c], -1), codes[:, c])
    off = synth.equal_partition_offsets(n, K, torch.device("cuda"))
    fits = eng.fit_logistic_partitions_cat(num, codes, n_levels, y, off, fit_intercept=True)
    f0, c0, p0 = eng.finish_fit(fits, True, float(n))
    for pinned in (True, False):
        nh, ch, yh = num.cpu(), codes.cpu(), y.cpu()
        if pinned:
            nh, ch, yh = nh.pin_memory(), ch.pin_memory(), yh.pin_memory()
        f1, c1, p1 = eng.dlsa_fit_host_cat(nh, ch, n_levels, yh, off.cpu(), fit_intercept=True, sample_size=float(n), n_groups=4)
        assert int(f1.status.abs().sum()) == 0
        assert (f1.coef - f0.coef).abs().max().item() < 1e-9
        assert float((f1.Sig_inv - f0.Sig_inv).norm() / f0.Sig_inv.norm()) < 1e-10
        assert (c1.beta_byOLS - c0.beta_byOLS).abs().max().item() < 1e-9
        assert p1.argmin_BIC == p0.argmin_BIC and torch.equal(p1.beta[p1.argmin_BIC] != 0, p0.beta[p0.argmin_BIC] != 0)


def test_cat_pass_is_bitwise_reproducible_with_shared_tables(eng):
    """Warp pairs take turns on one set of level tables in a fixed order (mbarrier hand-over): the sums must not depend on
    timing.  Same pass five times on the airline shape, partitions of very different sizes: identical bits."""
    from dlsa_b200 import synth
    n = 3_000_000
    num, codes, n_levels, y = synth.airline_shaped_device(n, 5, torch.device("cuda"))
    sizes = np.array([1, 700_000, 33, 1_200_000, 4096, 1_095_870 - 4096 + 4096], dtype=np.int64)
    sizes[-1] = n - sizes[:-1].sum()
    off = torch.as_tensor(np.concatenate([[0], np.cumsum(sizes)]), device="cuda")
    p = 1 + num.shape[1] + sum(n_levels)
    beta = torch.randn(len(sizes), p, dtype=torch.float64, device="cuda", generator=torch.Generator("cuda").manual_seed(1)) * 0.2
    g0, H0, ll0 = eng.irls_pass_cat(num, codes, n_levels, y, off, beta, fit_intercept=True)
    g0, H0, ll0 = g0.clone(), H0.clone(), ll0.clone()
    for _ in range(4):
        g1, H1, ll1 = eng.irls_pass_cat(num, codes, n_levels, y, off, beta, fit_intercept=True)
        assert torch.equal(g1, g0) and torch.equal(H1, H0) and torch.equal(ll1, ll0)


def test_cat_pass_random_shapes_match_oracle(eng):
    """Thirty seeded random shapes (1..8 categorical columns, 1..60 levels each, 0..7 numerics, with and without intercept
    and standardisation, constant and heavy-tailed level laws, partitions of 0..700 rows): every mix of DMMA tile slots,
    scatter rounds and table sharing the planner can produce, against the oracle on the expanded matrix."""
    rng = np.random.default_rng(2024)
    for trial in range(30):
        n_cat = int(rng.integers(1, 9))
        lev = rng.integers(1, 61, size=n_cat)
        while lev.sum() > 192:
            lev[np.argmax(lev)] //= 2
        n_levels = [int(x) for x in lev]
        n_num = int(rng.integers(0, 8))
        icpt = bool(rng.integers(0, 2))
        sizes = [int(x) for x in rng.integers(0, 700, size=5)] + [65]
        off, numeric, codes, y = make_case(rng, sizes, n_num, n_levels, baseline_share=float(rng.choice([0.0, 0.2, 0.9])),
                                           zipf=bool(trial % 2))
        p = int(icpt) + n_num + sum(n_levels)
        beta = rng.standard_normal((len(sizes), p)) * 0.3
        mean = std = None
        if n_num and trial % 3 == 0:
            mean, std = rng.standard_normal(n_num) * 0.3, 0.5 + rng.random(n_num)
        g, H, ll = eng.irls_pass_cat(dev(numeric) if n_num else None, dev(codes, torch.int32), n_levels, dev(y),
                                     dev(off, torch.int64), dev(beta), fit_intercept=icpt,
                                     mean=None if mean is None else dev(mean), std=None if std is None else dev(std))
        g, H, ll = g.cpu().numpy(), H.cpu().numpy(), ll.cpu().numpy()
        for k, nk in enumerate(sizes):
            sl = slice(off[k], off[k + 1])
            g0, H0, ll0 = onp.irls_pass_cat(numeric[sl], codes[sl], n_levels, y[sl], beta[k], icpt, mean, std)
            tag = (trial, n_num, n_levels, icpt, k, nk)
            assert np.allclose(g[k], g0, rtol=1e-10, atol=1e-9), tag
            assert np.allclose(H[k], H0, rtol=1e-10, atol=1e-10), tag
            assert abs(ll[k] - ll0) <= 1e-11 * max(1.0, abs(ll0)), tag
