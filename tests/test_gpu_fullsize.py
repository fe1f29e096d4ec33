"""GPU tests of the pooled warm start and of the hot path at BASELINE.json's full C2 size (n=1e8, p=50, K=1024).

The oracle cannot fit 1e8 rows in seconds, so the full-size case is checked through size-independent properties:
the score equation at the returned coefficients, additivity of (g, H) under merging partitions, the combine
identity, warm start == cold start, and a direct oracle fit of three sampled partitions (97 656 rows each).
Tolerances (north_star): theta <= 1e-6 max-abs, Sig_inv <= 1e-9 relative Frobenius; the checks below are tighter."""
import numpy as np
import pytest
import torch

from oracle import numpy_dlsa as onp

pytestmark = pytest.mark.gpu


def relfro_t(a, b):
    return float((a - b).norm() / b.norm())


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dlsa_b200 import engine
    return engine


def _sim(n, p, K, seed):
    from dlsa_b200 import synth
    dev = torch.device("cuda", 0)
    X = torch.empty(n, p, dtype=torch.float64, device=dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    chunk = 10_000_000
    for s in range(0, n, chunk):
        m = min(chunk, n - s)
        synth.simulate_logistic_device(m, p, seed, dev, row0=s, out=(X[s:s + m], y[s:s + m]))
    return X, y, synth.equal_partition_offsets(n, K, dev)


def test_warm_start_gives_the_cold_start_fit(eng):
    """EngineConfig.warm_start only changes where Newton starts (R's init.beta, dlsa/R/wlse_func.R:24,28): every
    partition must reach the MLE the cold start (beta = 0, :28-31) reaches, in fewer passes over X."""
    X, y, off = _sim(2_560_000, 20, 128, 3)
    cold = eng.fit_logistic_partitions(X, y, off, config=eng.EngineConfig(warm_start=False))
    warm = eng.fit_logistic_partitions(X, y, off)
    assert cold.pilot_partitions == 0 and warm.pilot_partitions == 8 and not warm.cold_restart
    assert int(cold.status.abs().sum().item()) == 0 and torch.equal(warm.status, cold.status)
    assert (warm.coef - cold.coef).abs().max().item() < 1e-10
    assert relfro_t(warm.Sig_inv, cold.Sig_inv) < 1e-11
    assert warm.passes < cold.passes
    # Sig_inv is still the H of a pass at exactly the returned coefficients (dlsa/models.py:114,130)
    _, H, _ = eng.irls_pass(X, y, off, warm.coef, want_loglik=False)
    assert torch.equal(H, warm.Sig_inv)
    # and against the oracle on two partitions
    offh = off.cpu().numpy()
    for k in (5, 127):
        sl = slice(int(offh[k]), int(offh[k + 1]))
        c0, S0, _ = onp.logistic_fit(X[sl].cpu().numpy(), y[sl].cpu().numpy())
        assert np.max(np.abs(warm.coef[k].cpu().numpy() - c0)) < 1e-9
        assert np.linalg.norm(warm.Sig_inv[k].cpu().numpy() - S0) / np.linalg.norm(S0) < 1e-9
    # the host pipeline chains the pooled start across its partition blocks
    Xh, yh = X.cpu().pin_memory(), y.cpu().pin_memory()
    f1, c1, p1 = eng.dlsa_fit_host(Xh, yh, off.cpu(), sample_size=X.shape[0], n_groups=2)
    assert torch.equal(f1.status, cold.status) and (f1.coef - cold.coef).abs().max().item() < 1e-10
    assert relfro_t(f1.Sig_inv, cold.Sig_inv) < 1e-11


def test_warm_start_falls_back_to_cold_start_when_a_partition_fails(eng):
    """A partition the warm-started fit cannot bring to ST_OK (here: a perfectly separable partition, whose MLE is at
    infinity) sends the batch through the reference's cold start (beta = 0, dlsa/R/wlse_func.R:28-31): statuses and
    estimates must be exactly those of warm_start=False."""
    X, y, off = _sim(160_000, 12, 80, 9)
    offh = off.cpu().numpy()
    y = y.clone()
    sl = slice(int(offh[40]), int(offh[41]))
    y[sl] = (X[sl, 0] > 0).to(torch.float64)
    cfg = dict(max_iter=12)
    cold = eng.fit_logistic_partitions(X, y, off, config=eng.EngineConfig(warm_start=False, **cfg))
    warm = eng.fit_logistic_partitions(X, y, off, config=eng.EngineConfig(**cfg))
    assert int(cold.status[40].item()) != 0 and int((cold.status != 0).sum().item()) == 1
    assert warm.cold_restart and warm.pilot_partitions > 0
    assert torch.equal(warm.status, cold.status)
    assert torch.equal(warm.coef, cold.coef) and torch.equal(warm.Sig_inv, cold.Sig_inv)
    comb = eng.combine_partitions(warm)
    assert comb.K_eff == 79 and comb.K_total == 80


def test_c2_full_size_properties(eng):
    """BASELINE.json configs[1]: n=1e8, p=50, 1024 partitions on one GPU (40.8 GB of float64)."""
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 60e9:
        pytest.skip("needs ~45 GB of free HBM")
    n, p, K = 100_000_000, 50, 1024
    X, y, off = _sim(n, p, K, 20210517)
    fits, comb, path = eng.dlsa_fit(X, y, off, sample_size=n)
    assert int(fits.status.abs().sum().item()) == 0 and fits.pilot_partitions == 32 and not fits.cold_restart
    assert comb.K_eff == K

    # (1) score equation: a fresh pass at the returned coefficients reproduces Sig_inv bit for bit and its Newton
    #     step is below the solver tolerance in every partition
    g, H, _ = eng.irls_pass(X, y, off, fits.coef, want_loglik=False)
    assert torch.equal(H, fits.Sig_inv)
    step = torch.linalg.solve(H, g.unsqueeze(-1)).squeeze(-1)
    assert step.abs().max().item() < 1e-9

    # (2) additivity: merging partitions 2j, 2j+1 adds their (g, H) at a common beta
    common = comb.beta_byOLS.view(1, p)
    g1, H1, l1 = eng.irls_pass(X, y, off, common.expand(K, p).contiguous())
    off2 = off[::2].contiguous()
    g2, H2, l2 = eng.irls_pass(X, y, off2, common.expand(K // 2, p).contiguous())
    assert relfro_t(H1.view(K // 2, 2, p, p).sum(1), H2) < 1e-13
    assert (g1.view(K // 2, 2, p).sum(1) - g2).abs().max().item() < 1e-7 * g2.abs().max().item() + 1e-6
    assert abs(l1.sum().item() - l2.sum().item()) < 1e-12 * abs(l2.sum().item())
    # every weight is in (0, 1/4]: 0 < H_jj <= sum x_j^2 / 4 <= n_k / 16
    d = torch.diagonal(H1, dim1=1, dim2=2)
    assert d.min().item() > 0 and d.max().item() <= (n / K + 1) / 16

    # (3) combine identity: (sum_k H_k) theta = sum_k H_k theta_k, Sig_inv_sum = sum_k H_k, one-shot = mean coef
    S = fits.Sig_inv.sum(0)
    assert relfro_t(comb.Sig_inv_sum, S) < 1e-13
    rhs = fits.Sig_invMcoef.sum(0)
    assert (S @ comb.beta_byOLS - rhs).abs().max().item() < 1e-9 * rhs.abs().max().item()
    assert (comb.beta_byONESHOT - fits.coef.mean(0)).abs().max().item() < 1e-13
    # the estimate recovers the data law (beta = 1 on the first int(0.4p) coordinates, dlsa/models.py:12,18-19)
    truth = torch.zeros(p, dtype=torch.float64, device=X.device)
    truth[:20] = 1.0
    assert (comb.beta_byOLS - truth).abs().max().item() < 5e-3
    sel = path.beta[path.argmin_BIC] != 0
    assert torch.equal(sel, truth != 0)

    # (4) warm start == cold start at full size
    cold = eng.fit_logistic_partitions(X, y, off, config=eng.EngineConfig(warm_start=False))
    assert (cold.coef - fits.coef).abs().max().item() < 1e-10
    assert relfro_t(fits.Sig_inv, cold.Sig_inv) < 1e-11
    assert fits.passes < cold.passes

    # (5) three partitions fitted by the oracle itself
    offh = off.cpu().numpy()
    for k in (0, 517, 1023):
        sl = slice(int(offh[k]), int(offh[k + 1]))
        c0, S0, _ = onp.logistic_fit(X[sl].cpu().numpy(), y[sl].cpu().numpy())
        assert np.max(np.abs(fits.coef[k].cpu().numpy() - c0)) < 1e-9
        assert np.linalg.norm(fits.Sig_inv[k].cpu().numpy() - S0) / np.linalg.norm(S0) < 1e-9


def test_c5_shaped_p500_many_small_partitions(eng):
    """BASELINE.json configs[4] in miniature: p=500 sparse-truth logistic (25 ones), 2 500-row partitions, 'lar' path +
    BIC.  64 of the 8 192 partitions: the oracle fits them and walks the 500-knot path in ~10 s."""
    rng = np.random.default_rng(55)
    K, nk, p = 64, 2500, 500
    n = K * nk
    X = rng.random((n, p)) - 0.5
    beta = np.r_[np.ones(25), np.zeros(p - 25)]
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    off = np.arange(K + 1, dtype=np.int64) * nk
    Xd = torch.as_tensor(X, device="cuda")
    yd = torch.as_tensor(y, device="cuda")
    od = torch.as_tensor(off, device="cuda")
    fits, comb, path = eng.dlsa_fit(Xd, yd, od, sample_size=n, type="lar")
    assert int(fits.status.abs().sum().item()) == 0 and fits.pilot_partitions == 8
    ref = [onp.logistic_fit(X[off[k]:off[k + 1]], y[off[k]:off[k + 1]]) for k in range(K)]
    coef = fits.coef.cpu().numpy()
    Sig = fits.Sig_inv.cpu().numpy()
    for k in range(K):
        assert np.max(np.abs(coef[k] - ref[k][0])) < 1e-8, k
        assert np.linalg.norm(Sig[k] - ref[k][1]) / np.linalg.norm(ref[k][1]) < 1e-9, k
    th, one, S = onp.combine(np.stack([r[0] for r in ref]), np.stack([r[1] for r in ref]))
    assert np.max(np.abs(comb.beta_byOLS.cpu().numpy() - th)) < 1e-6
    assert np.max(np.abs(comb.beta_byONESHOT.cpu().numpy() - one)) < 1e-8
    assert np.linalg.norm(comb.Sig_inv_sum.cpu().numpy() - S) / np.linalg.norm(S) < 1e-9
    rp = onp.lars_lsa(S, th, False, n, type="lar")
    bp = path.beta.cpu().numpy()
    assert bp.shape == rp["beta"].shape
    assert np.array_equal(bp != 0, rp["beta"] != 0)          # same support on every knot
    assert np.max(np.abs(bp - rp["beta"])) < 1e-6
    assert path.argmin_BIC == int(np.argmin(rp["BIC"])) and path.argmin_AIC == int(np.argmin(rp["AIC"]))
    assert np.allclose(path.BIC.cpu().numpy(), rp["BIC"], rtol=1e-9)


def test_c3_compact_full_size_against_oracle_on_sampled_partitions(eng):
    """BASELINE.json configs[2] through the compact categorical pass K1c: the WHOLE airline-shaped data set (n = 1.2e8,
    K = 120 partitions of 1e6 rows, p = 182) on one GPU.  The oracle cannot fit 1e6 x 182 in seconds, but ONE pass of it
    on a sampled partition's expanded design at the returned coefficients pins both halves of the claim: its H is the
    returned Sig_inv (<= 1e-9 rel. Frobenius) and its Newton step H^-1 g is below 1e-8, i.e. the returned coefficients
    are that partition's MLE.  Plus shard invariance of the counter-based generator and the combine identity."""
    from dlsa_b200 import synth
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 40e9:
        pytest.skip("needs ~15 GB of free HBM")
    dev = torch.device("cuda", 0)
    n, K, seed = 120_000_000, 120, 20210517
    num = torch.empty(n, synth.AIRLINE_NUM, dtype=torch.float64, device=dev)
    codes = torch.empty(n, len(synth.AIRLINE_LEVELS), dtype=torch.int32, device=dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    for s in range(0, n, 20_000_000):
        _, _, n_levels, _ = synth.airline_shaped_device(20_000_000, seed, dev, row0=s,
                                                        out=(num[s:s + 20_000_000], codes[s:s + 20_000_000], y[s:s + 20_000_000]))
    off = synth.equal_partition_offsets(n, K, dev)
    fits = eng.fit_logistic_partitions_cat(num, codes, n_levels, y, off, fit_intercept=True)
    assert int(fits.status.abs().sum().item()) == 0
    p = fits.coef.shape[1]
    assert p == 182
    comb = eng.combine_partitions(fits)
    S = fits.Sig_inv.sum(0)
    assert relfro_t(comb.Sig_inv_sum, S) < 1e-13
    rhs = fits.Sig_invMcoef.sum(0)
    assert (S @ comb.beta_byOLS - rhs).abs().max().item() < 1e-9 * rhs.abs().max().item()
    b0, bnum, blev = synth.airline_truth()
    truth = np.r_[b0, bnum, blev]
    assert np.max(np.abs(comb.beta_byOLS.cpu().numpy() - truth)) < 0.02     # 1.2e8 rows recover the data law
    offh = off.cpu().numpy()
    for k in (0, 77):
        lo, hi = int(offh[k]), int(offh[k + 1])
        # the shard [lo, hi) regenerated on its own is the same data (what a rank of an 8-GPU run would hold)
        a = synth.airline_shaped_device(hi - lo, seed, dev, row0=lo)
        assert torch.equal(a[0], num[lo:hi]) and torch.equal(a[1], codes[lo:hi]) and torch.equal(a[3], y[lo:hi])
        Xk = onp.expand_onehot(num[lo:hi].cpu().numpy(), codes[lo:hi].cpu().numpy(), n_levels)
        X1 = np.concatenate([np.ones((Xk.shape[0], 1)), Xk], axis=1)
        g0, H0, _ = onp.irls_pass(X1, y[lo:hi].cpu().numpy(), fits.coef[k].cpu().numpy())
        assert np.linalg.norm(fits.Sig_inv[k].cpu().numpy() - H0) / np.linalg.norm(H0) < 1e-9
        assert np.max(np.abs(np.linalg.solve(H0, g0))) < 1e-8
        del Xk, X1


def test_c5_full_size_k8192_properties(eng):
    """BASELINE.json configs[4] at its full size: p = 500 sparse-truth logistic, K = 8192 partitions of 2500 rows
    (81.9 GB of float64), 'lar' path + BIC.  Properties: every partition converged; the oracle's own fit of three sampled
    partitions; the score equation on a block of partitions; the combine identity; the path has p + 1 knots with
    decreasing RSS and its BIC choice contains the 25 true variables."""
    from dlsa_b200 import synth
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 130e9:
        pytest.skip("needs ~120 GB of free HBM")
    dev = torch.device("cuda", 0)
    K, nk, p = 8192, 2500, 500
    n = K * nk
    beta = np.r_[np.ones(25), np.zeros(p - 25)]
    X = torch.empty(n, p, dtype=torch.float64, device=dev)
    y = torch.empty(n, dtype=torch.float64, device=dev)
    for s in range(0, n, 1_000_000):
        m = min(1_000_000, n - s)
        synth.simulate_logistic_device(m, p, 7, dev, row0=s, beta=beta, out=(X[s:s + m], y[s:s + m]))
    off = synth.equal_partition_offsets(n, K, dev)
    fits, comb, path = eng.dlsa_fit(X, y, off, sample_size=n, type="lar", max_part_rows=nk)
    assert int(fits.status.abs().sum().item()) == 0 and comb.K_eff == K and not fits.cold_restart
    for k in (0, 4097, 8191):
        sl = slice(k * nk, (k + 1) * nk)
        c0, S0, _ = onp.logistic_fit(X[sl].cpu().numpy(), y[sl].cpu().numpy())
        assert np.max(np.abs(fits.coef[k].cpu().numpy() - c0)) < 1e-8
        assert np.linalg.norm(fits.Sig_inv[k].cpu().numpy() - S0) / np.linalg.norm(S0) < 1e-9
    # score equation on partitions 1000..1063: a fresh pass at the returned coefficients reproduces the returned Sig_inv
    # (to rounding: 64 partitions are cut into row slices the 8192-partition launch does not use, so the summation order
    # differs) and leaves a Newton step below the solver tolerance
    k0, k1 = 1000, 1064
    offb = (off[k0:k1 + 1] - off[k0]).contiguous()
    g, H, _ = eng.irls_pass(X[k0 * nk:k1 * nk], y[k0 * nk:k1 * nk], offb, fits.coef[k0:k1].contiguous(), want_loglik=False)
    assert relfro_t(H, fits.Sig_inv[k0:k1]) < 1e-13
    assert torch.linalg.solve(H, g.unsqueeze(-1)).abs().max().item() < 1e-8
    del X, y
    torch.cuda.empty_cache()
    # combine identity in chunks (16 GB of Sig_inv): fixed-order sum == the kernel's sum, theta solves the normal equation
    S = torch.zeros(p, p, dtype=torch.float64, device=dev)
    for c in range(0, K, 512):
        S += fits.Sig_inv[c:c + 512].sum(0)
    assert relfro_t(comb.Sig_inv_sum, S) < 1e-12
    rhs = fits.Sig_invMcoef.sum(0)
    assert (comb.Sig_inv_sum @ comb.beta_byOLS - rhs).abs().max().item() < 1e-9 * rhs.abs().max().item()
    assert (comb.beta_byONESHOT - fits.coef.mean(0)).abs().max().item() < 1e-12
    assert path.beta.shape == (p + 1, p) and int(path.dof[-1].item()) == p
    rss = path.RSS.cpu().numpy()
    assert np.all(np.diff(rss) <= 1e-9 * rss[0])
    sel = (path.beta[path.argmin_BIC] != 0).cpu().numpy()
    assert sel[:25].all() and sel.sum() <= 40
