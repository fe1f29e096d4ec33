"""GPU parity tests: the CUDA path (through the C ABI) against the oracle and the reference goldens.

Tolerances (north_star): combined theta <= 1e-6 max-abs, Sig_inv <= 1e-9 relative Frobenius, same support.
Kernel-level checks use much tighter bounds (float64 summation-order noise only)."""
import numpy as np
import pytest
import torch

from conftest import games_data, load_golden, sim_data
from oracle import numpy_dlsa as onp

pytestmark = pytest.mark.gpu

THETA_TOL = 1e-6
SIG_RTOL = 1e-9


def relfro(a, b):
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


@pytest.fixture(scope="module")
def eng():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import dlsa_b200
    from dlsa_b200 import engine
    return engine


def dev(a, dtype=torch.float64):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device="cuda")


def ragged_offsets(sizes):
    off = np.zeros(len(sizes) + 1, dtype=np.int64)
    off[1:] = np.cumsum(sizes)
    return off


@pytest.mark.parametrize("px,icpt", [(3, False), (4, True), (8, False), (9, False), (20, False), (20, True), (31, True),
                                     (37, False), (40, False), (48, False), (50, False), (50, True), (55, True),
                                     (56, False)])
def test_irls_pass_matches_oracle(eng, px, icpt):
    rng = np.random.default_rng(px * 7 + icpt)
    sizes = [0, 1, 31, 32, 33, 257, 1000, 4097, 64, 5]
    off = ragged_offsets(sizes)
    n = int(off[-1])
    p = px + int(icpt)
    X = rng.random((n, px)) - 0.5
    y = (rng.random(n) < 0.4).astype(np.float64)
    beta = rng.standard_normal((len(sizes), p)) * 0.7
    g, H, ll = eng.irls_pass(dev(X), dev(y), dev(off, torch.int64), dev(beta), fit_intercept=icpt)
    g, H, ll = g.cpu().numpy(), H.cpu().numpy(), ll.cpu().numpy()
    for k, nk in enumerate(sizes):
        Xk = X[off[k]:off[k + 1]]
        if icpt:
            Xk = np.concatenate([np.ones((nk, 1)), Xk], axis=1)
        g0, H0, ll0 = onp.irls_pass(Xk, y[off[k]:off[k + 1]], beta[k])
        assert np.allclose(g[k], g0, rtol=1e-11, atol=1e-11), (k, nk)
        assert np.allclose(H[k], H0, rtol=1e-11, atol=1e-12), (k, nk)
        assert abs(ll[k] - ll0) <= 1e-11 * max(1.0, abs(ll0))
        assert np.array_equal(H[k], H[k].T)  # exactly symmetric by construction


@pytest.mark.parametrize("px,icpt", [(57, False), (64, True), (65, False), (100, True), (128, False), (169, False),
                                     (170, True), (200, False), (256, False), (400, True), (511, True), (512, False)])
def test_irls_pass_tiled_widths_match_oracle(eng, px, icpt):
    """The CTA-tiled pass over the widths where its geometry changes (sibling CTAs from p ~ 170, 16-row chunks, streaming
    link kernel, ragged last tile row, p = 512), partitions shorter than a chunk, of one row, empty, and with tail chunks."""
    rng = np.random.default_rng(px * 7 + icpt)
    sizes = [0, 1, 15, 33, 700, 64]
    off = ragged_offsets(sizes)
    n = int(off[-1])
    p = px + int(icpt)
    X = rng.random((n, px)) - 0.5
    y = (rng.random(n) < 0.4).astype(np.float64)
    beta = rng.standard_normal((len(sizes), p)) * (2.0 / np.sqrt(p))
    g, H, ll = eng.irls_pass(dev(X), dev(y), dev(off, torch.int64), dev(beta), fit_intercept=icpt)
    g, H, ll = g.cpu().numpy(), H.cpu().numpy(), ll.cpu().numpy()
    for k, nk in enumerate(sizes):
        Xk = X[off[k]:off[k + 1]]
        if icpt:
            Xk = np.concatenate([np.ones((nk, 1)), Xk], axis=1)
        g0, H0, ll0 = onp.irls_pass(Xk, y[off[k]:off[k + 1]], beta[k])
        assert np.allclose(g[k], g0, rtol=1e-10, atol=1e-11), (k, nk)
        assert np.allclose(H[k], H0, rtol=1e-10, atol=1e-12), (k, nk)
        assert abs(ll[k] - ll0) <= 1e-11 * max(1.0, abs(ll0))
        assert np.array_equal(H[k], H[k].T)


def test_irls_pass_saturated_eta_and_slices(eng):
    """eta = +-40 rows (weights underflow towards 0) and a few big partitions (sliced across CTAs)."""
    rng = np.random.default_rng(5)
    px, K = 20, 3
    sizes = [70_000, 50_001, 90_017]
    off = ragged_offsets(sizes)
    n = int(off[-1])
    X = rng.random((n, px)) - 0.5
    X[::97] *= 160.0
    y = (rng.random(n) < 0.5).astype(np.float64)
    beta = np.ones((K, px))
    g, H, ll = eng.irls_pass(dev(X), dev(y), dev(off, torch.int64), dev(beta))
    for k in range(K):
        g0, H0, ll0 = onp.irls_pass(X[off[k]:off[k + 1]], y[off[k]:off[k + 1]], beta[k])
        assert np.allclose(g[k].cpu().numpy(), g0, rtol=1e-10, atol=1e-9)
        assert relfro(H[k].cpu().numpy(), H0) < 1e-12
        assert abs(ll[k].item() - ll0) <= 1e-11 * abs(ll0)


def test_irls_pass_unaligned_rows_and_ldx(eng):
    """odd px / odd leading dimension / 8-byte-offset base pointer take the cp.async (non-TMA) staging path."""
    rng = np.random.default_rng(6)
    n, px, ldx = 3000, 13, 17
    big = rng.random((n, ldx)) - 0.5
    y = (rng.random(n) < 0.5).astype(np.float64)
    off = ragged_offsets([1000, 1500, 500])
    beta = rng.standard_normal((3, px))
    Xd = dev(big)[:, 1:1 + px]  # base pointer 8 bytes off, stride 17
    g, H, ll = eng.irls_pass(Xd, dev(y), dev(off, torch.int64), dev(beta))
    for k in range(3):
        g0, H0, ll0 = onp.irls_pass(big[off[k]:off[k + 1], 1:1 + px], y[off[k]:off[k + 1]], beta[k])
        assert np.allclose(g[k].cpu().numpy(), g0, rtol=1e-11, atol=1e-11)
        assert relfro(H[k].cpu().numpy(), H0) < 1e-13


def run_fit(eng, X, y, pid, K, icpt=False):
    from dlsa_b200 import synth
    Xg, yg, off = synth.group_by_partition(dev(X), dev(y), dev(pid, torch.int64), K)
    fits = eng.fit_logistic_partitions(Xg, yg, off, fit_intercept=icpt)
    torch.cuda.synchronize()
    return fits


def check_against_golden(eng, fits, g, n):
    coef = fits.coef.cpu().numpy()
    assert fits.status.cpu().numpy().tolist() == [0] * coef.shape[0]
    assert np.max(np.abs(coef - g["ref_coef"])) < 1e-8
    sig = fits.Sig_inv.cpu().numpy()
    for k in range(coef.shape[0]):
        assert relfro(sig[k], g["ref_Sig_inv"][k]) < SIG_RTOL
    assert np.max(np.abs(fits.Sig_invMcoef.cpu().numpy() - g["ref_Sig_invMcoef"])) < 1e-6 * np.max(np.abs(g["ref_Sig_invMcoef"]))
    comb = eng.combine_partitions(fits)
    assert np.max(np.abs(comb.beta_byOLS.cpu().numpy() - g["ref_beta_byOLS"])) < THETA_TOL
    assert np.max(np.abs(comb.beta_byOLS.cpu().numpy() - g["ref_beta_byOLS"])) < 1e-9  # in fact far tighter
    assert np.max(np.abs(comb.beta_byONESHOT.cpu().numpy() - g["ref_beta_byONESHOT"])) < 1e-9
    assert relfro(comb.Sig_inv_sum.cpu().numpy(), g["ref_Sig_inv_sum"]) < SIG_RTOL
    return comb


def check_path(path, g, prefix):
    rb = g[prefix + "beta"]
    beta = path.beta.cpu().numpy()
    assert beta.shape == rb.shape
    assert np.max(np.abs(beta - rb)) < 1e-8 * max(1.0, np.max(np.abs(rb)))
    assert np.array_equal(beta != 0, rb != 0)
    assert np.allclose(path.AIC.cpu().numpy(), g[prefix + "AIC"], rtol=1e-8, atol=1e-6)
    assert np.allclose(path.BIC.cpu().numpy(), g[prefix + "BIC"], rtol=1e-8, atol=1e-6)
    assert path.argmin_BIC == int(np.argmin(g[prefix + "BIC"]))
    assert path.argmin_AIC == int(np.argmin(g[prefix + "AIC"]))


def test_games_expand_end_to_end(eng):
    X, y, K, g = games_data()
    pid = np.arange(X.shape[0]) % K
    fits = run_fit(eng, X, y, pid, K)
    comb = check_against_golden(eng, fits, g, X.shape[0])
    for typ in ("lar", "lasso"):
        path = eng.lars_path(comb.Sig_inv_sum, comb.beta_byOLS, False, X.shape[0], type=typ)
        check_path(path, g, f"ref_{typ}_")
    # SURVEY App. C known answers
    path = eng.lars_path(comb.Sig_inv_sum, comb.beta_byOLS, False, X.shape[0])
    assert path.argmin_BIC == 9 and abs(path.BIC[9].item() - 91.073913) < 1e-5
    assert path.beta[9][1].item() == 0.0


@pytest.mark.parametrize("name", ["sim_c1_n100000_p20_k16.npz", "sim_small_n6000_p7_k3.npz",
                                  "sim_icpt_n20000_p12_k4.npz", "sim_p50_n40000_k4.npz"])
def test_sim_goldens_end_to_end(eng, name):
    g = load_golden(name)
    pid, y, X, K = sim_data(g)
    icpt = bool(g["in_fit_intercept"])
    fits = run_fit(eng, X, y, pid, K, icpt)
    comb = check_against_golden(eng, fits, g, X.shape[0])
    if not icpt:
        for typ in ("lar", "lasso"):
            path = eng.lars_path(comb.Sig_inv_sum, comb.beta_byOLS, False, X.shape[0], type=typ)
            check_path(path, g, f"ref_{typ}_")
    else:
        S, th = g["ref_Sig_inv_sum"], g["ref_beta_byOLS"]
        ref = onp.lars_lsa(S, th, True, X.shape[0])
        path = eng.lars_path(dev(S), dev(th), True, X.shape[0])
        assert np.max(np.abs(path.beta.cpu().numpy() - ref["beta"])) < 1e-9
        assert np.max(np.abs(path.beta0.cpu().numpy() - ref["beta0"])) < 1e-9
        assert path.argmin_BIC == int(np.argmin(ref["BIC"]))


@pytest.mark.parametrize("typ", ["lar", "lasso"])
def test_lars_standalone_goldens(eng, typ):
    g = load_golden("lars_cases.npz")
    for i in range(int(g["n_cases"])):
        key = f"c{i}_{typ}_"
        path = eng.lars_path(dev(g[key + "in_Sigma"]), dev(g[key + "in_b"]), False, int(g[key + "in_n"]), type=typ)
        check_path(path, g, key + "ref_")


def test_linear_fit_matches_oracle(eng):
    rng = np.random.default_rng(2)
    n, p, K = 30_000, 24, 5
    X = rng.standard_normal((n, p))
    beta = np.tile([3, 1.5, 0, 0, 2, 0, 0, 0], 3).astype(float)
    y = X @ beta + rng.standard_normal(n)
    off = ragged_offsets([5000, 7000, 6000, 6001, 5999])
    fits = eng.fit_linear_partitions(dev(X), dev(y), dev(off, torch.int64))
    for k in range(K):
        th, Sg, Sgt, s2 = onp.ols_fit(X[off[k]:off[k + 1]], y[off[k]:off[k + 1]])
        assert np.max(np.abs(fits.coef[k].cpu().numpy() - th)) < 1e-10
        assert relfro(fits.Sig_inv[k].cpu().numpy(), Sg) < 1e-11
        assert np.allclose(fits.Sig_invMcoef[k].cpu().numpy(), Sgt, rtol=1e-10)
    comb = eng.combine_partitions(fits)
    th_ref, _, S_ref = onp.combine(*zip(*[onp.ols_fit(X[off[k]:off[k + 1]], y[off[k]:off[k + 1]])[:3]
                                           for k in range(K)]))


def test_loglik_eval_matches_oracle(eng):
    rng = np.random.default_rng(3)
    n, px = 20_000, 11
    X = rng.random((n, px)) - 0.5
    y = (rng.random(n) < 0.5).astype(np.float64)
    off = ragged_offsets([3000, 9000, 8000])
    for icpt in (False, True):
        p = px + int(icpt)
        par = rng.standard_normal((p, 4))
        out = eng.logistic_loglik(dev(X), dev(y), dev(off, torch.int64), dev(par), fit_intercept=icpt).cpu().numpy()
        for k in range(3):
            Xk = X[off[k]:off[k + 1]]
            if icpt:
                Xk = np.concatenate([np.ones((Xk.shape[0], 1)), Xk], axis=1)
            ref = onp.logistic_loglik(Xk, y[off[k]:off[k + 1]], par)
            assert np.allclose(out[k], ref, rtol=1e-11)


def test_device_generator_matches_host_restatement(eng):
    from dlsa_b200 import synth
    from oracle import synth_host
    n, p, seed, row0 = 5000, 20, 20210517, 123_456_789
    X, y = synth.simulate_logistic_device(n, p, seed, "cuda", row0=row0)
    rows = (np.arange(n, dtype=np.uint64) + np.uint64(row0))[:, None]
    cols = np.arange(p, dtype=np.uint64)[None, :]
    Xh = synth_host.host_uniform(seed, rows, cols) - 0.5
    assert np.array_equal(X.cpu().numpy(), Xh)  # bit-exact
    u = synth_host.host_uniform(seed, rows[:, 0], np.uint64(synth.LABEL_COL))
    prob = 1 / (1 + np.exp(-(Xh @ synth.true_beta(p))))
    yh = (u < prob).astype(np.float64)
    assert np.mean(y.cpu().numpy() != yh) < 1e-3  # identical up to exp() rounding at the threshold
    # sharding invariance: rows generated as two shards equal the single-shot rows
    Xa, _ = synth.simulate_logistic_device(2000, p, seed, "cuda", row0=row0)
    Xb, _ = synth.simulate_logistic_device(3000, p, seed, "cuda", row0=row0 + 2000)
    assert torch.equal(torch.cat([Xa, Xb]), X)


def test_empty_and_error_paths(eng):
    from dlsa_b200 import _lib
    X = dev(np.zeros((4, 3)))
    with pytest.raises(TypeError):
        eng.irls_pass(X.float(), dev(np.zeros(4)), dev([0, 4], torch.int64), dev(np.zeros((1, 3))))
    with pytest.raises(ValueError):
        eng.irls_pass(torch.zeros(4, 3, dtype=torch.float64), torch.zeros(4, dtype=torch.float64),
                      torch.tensor([0, 4]), torch.zeros(1, 3, dtype=torch.float64))
    # a partition with a single class / separable data diverges: reported through status, not an exception
    Xs = dev(np.array([[1.0, 0.2], [2.0, -0.1], [3.0, 0.3], [4.0, 0.0]]))
    ys = dev(np.array([1.0, 1.0, 1.0, 1.0]))
    fits = eng.fit_logistic_partitions(Xs, ys, dev([0, 4], torch.int64), config=eng.EngineConfig(max_iter=8))
    assert fits.status.item() in (_lib.ST_NOT_CONVERGED, _lib.ST_NOT_SPD)
    with pytest.raises(Exception, match="Zero-length"):
        eng.combine_partitions(fits)


@pytest.mark.parametrize("px,icpt,sizes", [(57, False, [700, 33, 0, 1500]), (100, False, [3000, 1001]),
                                           (181, True, [2500, 999]), (181, False, [1200]), (182, False, [1500, 700]),
                                           (200, False, [900]), (256, True, [1100]), (500, False, [1300, 600])])
def test_irls_pass_large_p_matches_oracle(eng, px, icpt, sizes):
    """p > 56 goes through the CTA-tiled kernel (several tile groups at p >= 200)."""
    rng = np.random.default_rng(px + icpt)
    off = ragged_offsets(sizes)
    n = int(off[-1])
    p = px + int(icpt)
    X = rng.random((n, px)) - 0.5
    y = (rng.random(n) < 0.4).astype(np.float64)
    beta = rng.standard_normal((len(sizes), p)) * (2.0 / np.sqrt(p))
    g, H, ll = eng.irls_pass(dev(X), dev(y), dev(off, torch.int64), dev(beta), fit_intercept=icpt)
    g, H, ll = g.cpu().numpy(), H.cpu().numpy(), ll.cpu().numpy()
    for k, nk in enumerate(sizes):
        Xk = X[off[k]:off[k + 1]]
        if icpt:
            Xk = np.concatenate([np.ones((nk, 1)), Xk], axis=1)
        g0, H0, ll0 = onp.irls_pass(Xk, y[off[k]:off[k + 1]], beta[k])
        assert np.allclose(g[k], g0, rtol=1e-10, atol=1e-10), (k, nk)
        assert relfro(H[k], H0) < 1e-13 if nk else np.all(H[k] == 0)
        assert abs(ll[k] - ll0) <= 1e-11 * max(1.0, abs(ll0))
        assert np.array_equal(H[k], H[k].T)


@pytest.mark.parametrize("px,icpt", [(181, True), (181, False), (99, False), (301, True)])
def test_large_p_odd_px_padded_pitch_takes_tma_path(eng, px, icpt):
    """odd px stored with an even row pitch (what the airline one-hot expand produces): rows are 16-byte aligned, the
    TMA row copy drags the padding column along and the kernel must repair it (the padding holds NaN here)."""
    rng = np.random.default_rng(px)
    sizes = [1700, 0, 811]
    off = ragged_offsets(sizes)
    n = int(off[-1])
    p = px + int(icpt)
    big = np.full((n, px + 1), np.nan)
    big[:, :px] = rng.random((n, px)) - 0.5
    y = (rng.random(n) < 0.4).astype(np.float64)
    beta = rng.standard_normal((len(sizes), p)) * (2.0 / np.sqrt(p))
    Xd = dev(big)[:, :px]
    assert Xd.stride(0) == px + 1
    g, H, ll = eng.irls_pass(Xd, dev(y), dev(off, torch.int64), dev(beta), fit_intercept=icpt)
    g, H, ll = g.cpu().numpy(), H.cpu().numpy(), ll.cpu().numpy()
    for k, nk in enumerate(sizes):
        Xk = big[off[k]:off[k + 1], :px]
        if icpt:
            Xk = np.concatenate([np.ones((nk, 1)), Xk], axis=1)
        g0, H0, ll0 = onp.irls_pass(Xk, y[off[k]:off[k + 1]], beta[k])
        assert np.allclose(g[k], g0, rtol=1e-10, atol=1e-10), (k, nk)
        assert relfro(H[k], H0) < 1e-13 if nk else np.all(H[k] == 0)
        assert abs(ll[k] - ll0) <= 1e-11 * max(1.0, abs(ll0))
        assert np.array_equal(H[k], H[k].T)


def test_large_p_sliced_partitions(eng):
    """few big partitions at p=100: slices + fixed-order slice reduce."""
    rng = np.random.default_rng(9)
    px, sizes = 100, [40_000, 25_003]
    off = ragged_offsets(sizes)
    X = rng.random((int(off[-1]), px)) - 0.5
    y = (rng.random(int(off[-1])) < 0.5).astype(np.float64)
    beta = rng.standard_normal((2, px)) * 0.2
    g, H, ll = eng.irls_pass(dev(X), dev(y), dev(off, torch.int64), dev(beta))
    for k in range(2):
        g0, H0, ll0 = onp.irls_pass(X[off[k]:off[k + 1]], y[off[k]:off[k + 1]], beta[k])
        assert np.allclose(g[k].cpu().numpy(), g0, rtol=1e-10, atol=1e-9)
        assert relfro(H[k].cpu().numpy(), H0) < 1e-13
        assert abs(ll[k].item() - ll0) <= 1e-11 * abs(ll0)


def test_airline_shaped_fit_p182(eng):
    """C3-shaped miniature: intercept + 7 numerics + 174 one-hot dummies (p = 182), 3 partitions."""
    rng = np.random.default_rng(10)
    n = 60_000
    num = rng.standard_normal((n, 7))
    cards = [11, 6, 7, 75, 75]  # levels minus baseline
    cols = [num]
    for c in cards:
        w = 1.0 / np.arange(1, c + 2)
        lv = rng.choice(c + 1, size=n, p=w / w.sum())
        oh = np.zeros((n, c))
        m = lv > 0
        oh[np.nonzero(m)[0], lv[m] - 1] = 1.0
        cols.append(oh)
    X = np.concatenate(cols, axis=1)
    assert X.shape[1] == 181
    beta = rng.standard_normal(182) * 0.3 * (rng.random(182) < 0.4)
    beta[0] = -1.0
    y = (rng.random(n) < 1 / (1 + np.exp(-(beta[0] + X @ beta[1:])))).astype(np.float64)
    off = ragged_offsets([20_000, 20_000, 20_000])
    fits = eng.fit_logistic_partitions(dev(X), dev(y), dev(off, torch.int64), fit_intercept=True)
    assert fits.status.cpu().tolist() == [0, 0, 0]
    for k in range(3):
        c0, S0, _ = onp.logistic_fit(X[off[k]:off[k + 1]], y[off[k]:off[k + 1]], fit_intercept=True)
        assert np.max(np.abs(fits.coef[k].cpu().numpy() - c0)) < 1e-7
        assert relfro(fits.Sig_inv[k].cpu().numpy(), S0) < SIG_RTOL
    comb = eng.combine_partitions(fits)
    path = eng.lars_path(comb.Sig_inv_sum, comb.beta_byOLS, True, n)
    ref = onp.lars_lsa(comb.Sig_inv_sum.cpu().numpy(), comb.beta_byOLS.cpu().numpy(), True, n)
    assert path.beta.shape == ref["beta"].shape
    assert np.max(np.abs(path.beta.cpu().numpy() - ref["beta"])) < 1e-7
    assert path.argmin_BIC == int(np.argmin(ref["BIC"]))
    assert np.array_equal(path.beta[path.argmin_BIC].cpu().numpy() != 0, ref["beta"][path.argmin_BIC] != 0)


def test_linear_p100_matches_oracle(eng):
    rng = np.random.default_rng(12)
    n, p = 24_000, 100
    X = rng.standard_normal((n, p))
    beta = np.tile([3, 1.5, 0, 0, 2, 0, 0, 0, 0, 0], 10).astype(float)
    y = X @ beta + rng.standard_normal(n)
    off = ragged_offsets([8000, 9000, 7000])
    fits = eng.fit_linear_partitions(dev(X), dev(y), dev(off, torch.int64))
    for k in range(3):
        th, Sg, Sgt, s2 = onp.ols_fit(X[off[k]:off[k + 1]], y[off[k]:off[k + 1]])
        assert np.max(np.abs(fits.coef[k].cpu().numpy() - th)) < 1e-10
        assert relfro(fits.Sig_inv[k].cpu().numpy(), Sg) < 1e-11


@pytest.mark.parametrize("sizes", [[9223, 9013, 7777], [32 * 40 + 7] * 5, [551] * 160])
def test_linear_large_p_tail_chunk_in_a_reused_stage(eng, sizes):
    """Linear family, p > 56, partitions whose last (partial) chunk lands in a ring stage an earlier chunk already used:
    the linear Gram body does not scale by w, so the rows past the partition's end must have been blanked (the
    advisor's round-1 finding: X'X picked up 8 stale rows per partition at C4-shaped inputs)."""
    rng = np.random.default_rng(31)
    n, p = int(np.sum(sizes)), 100
    X = rng.standard_normal((n, p))
    beta = np.tile([3, 1.5, 0, 0, 2, 0, 0, 0, 0, 0], 10).astype(float)
    y = X @ beta + rng.standard_normal(n)
    off = ragged_offsets(sizes)
    g, H, yty = eng.irls_pass(dev(X), dev(y), dev(off, torch.int64), None, False, 1, want_loglik=True)
    Hh, gh, yh = H.cpu().numpy(), g.cpu().numpy(), yty.cpu().numpy()
    for k in sorted(set([0, 1, len(sizes) // 2, len(sizes) - 1])):
        Xk, yk = X[off[k]:off[k + 1]], y[off[k]:off[k + 1]]
        assert relfro(Hh[k], Xk.T @ Xk) < 1e-12
        assert np.max(np.abs(gh[k] - Xk.T @ yk)) < 1e-9
        assert abs(yh[k] - yk @ yk) < 1e-9 * (yk @ yk)


def test_linear_p300_sixteen_row_chunks_with_ragged_tails(eng):
    """Linear family at a width that takes 16-row chunks (two lanes per row in the link warps): partitions whose length
    is not a multiple of 16, more chunks per slice than ring stages."""
    rng = np.random.default_rng(33)
    p, sizes = 300, [2500, 2503, 16 * 40 + 5, 1, 17]
    n = int(np.sum(sizes))
    X = rng.standard_normal((n, p))
    y = X[:, :10] @ np.arange(1.0, 11.0) + rng.standard_normal(n)
    off = ragged_offsets(sizes)
    g, H, yty = eng.irls_pass(dev(X), dev(y), dev(off, torch.int64), None, False, 1, want_loglik=True)
    Hh, gh, yh = H.cpu().numpy(), g.cpu().numpy(), yty.cpu().numpy()
    for k in range(len(sizes)):
        Xk, yk = X[off[k]:off[k + 1]], y[off[k]:off[k + 1]]
        assert relfro(Hh[k], Xk.T @ Xk) < 1e-12
        assert np.max(np.abs(gh[k] - Xk.T @ yk)) < 1e-9 * max(1.0, np.max(np.abs(Xk.T @ yk)))
        assert abs(yh[k] - yk @ yk) <= 1e-9 * (yk @ yk)


@pytest.mark.parametrize("px,icpt", [(9, True), (50, True), (55, True), (56, False), (57, False), (63, True), (64, False), (96, False), (103, True), (104, False),
                                     (181, True), (511, False), (511, True), (512, False)])
def test_linear_pass_y_column_edge_widths(eng, px, icpt):
    """Linear family: y rides the Gram as column p of the staged rows (g = row p of the Gram of [X y]).  Widths where
    that column opens a new tile row (p % 8 == 0), sits in the padding of the last one, collides with the column an odd
    px drags along, follows the intercept column, is the last one that fits (p + 1 = 512) — and p = 512, where the kernel
    falls back to the gradient-owner bodies."""
    rng = np.random.default_rng(px * 2 + icpt)
    sizes = [300, 1, 33, 257]
    n = int(np.sum(sizes))
    X = rng.standard_normal((n, px))
    y = X[:, :5] @ np.arange(1.0, 6.0) + 0.5 + rng.standard_normal(n)
    off = ragged_offsets(sizes)
    Xd = dev(X)
    variants = [Xd]
    if px % 2:   # even row pitch: rows go by one bulk copy each and drag the (garbage) padding column along
        Xpad = torch.full((n, px + 1), 7.25, dtype=torch.float64, device="cuda")
        Xpad[:, :px] = Xd
        variants.append(Xpad[:, :px])
    for Xv in variants:
        g, H, yty = eng.irls_pass(Xv, dev(y), dev(off, torch.int64), None, icpt, 1, want_loglik=True)
        Hh, gh, yh = H.cpu().numpy(), g.cpu().numpy(), yty.cpu().numpy()
        for k in range(len(sizes)):
            Xk, yk = X[off[k]:off[k + 1]], y[off[k]:off[k + 1]]
            if icpt:
                Xk = np.hstack([np.ones((Xk.shape[0], 1)), Xk])
            assert relfro(Hh[k], Xk.T @ Xk) < 1e-12
            assert np.max(np.abs(gh[k] - Xk.T @ yk)) < 1e-9 * max(1.0, np.max(np.abs(Xk.T @ yk)))
            assert abs(yh[k] - yk @ yk) <= 1e-9 * (yk @ yk)
            assert np.array_equal(Hh[k], Hh[k].T)


def test_glm_pass_on_a_window_of_the_offsets(eng):
    """A window part_offsets[lo:hi+1] of a longer offsets array with the SHARED X (what parallel.shard_partitions hands
    out): the streaming link kernel's w / r buffers are indexed relative to the window's first row."""
    rng = np.random.default_rng(32)
    p, sizes = 200, [700, 650, 720, 690, 710, 705]
    n = int(np.sum(sizes))
    X = rng.uniform(-0.5, 0.5, (n, p))
    y = (rng.uniform(size=n) < 0.5).astype(float)
    off = ragged_offsets(sizes)
    beta = 0.05 * rng.standard_normal((len(sizes), p))
    Xd, yd = dev(X), dev(y)
    g_all, H_all, ll_all = eng.irls_pass(Xd, yd, dev(off, torch.int64), dev(beta))
    lo, hi = 3, 6
    g_w, H_w, ll_w = eng.irls_pass(Xd, yd, dev(off[lo:hi + 1], torch.int64), dev(beta[lo:hi]))
    assert torch.equal(H_w, H_all[lo:hi]) and torch.equal(g_w, g_all[lo:hi]) and torch.equal(ll_w, ll_all[lo:hi])


def test_returned_sig_inv_is_evaluated_at_returned_coef(eng):
    """dlsa/models.py:114,130: Sig_inv is X' diag(p(1-p)) X at the coefficients the fit RETURNS.  K2 does not apply a
    final step below tol, so the H of the last Newton pass must equal a fresh pass at fits.coef bit for bit."""
    g = load_golden("sim_p50_n40000_k4.npz")
    pid, y, X, K = sim_data(g)
    order = np.argsort(pid, kind="stable")
    Xs, ys = X[order], y[order]
    off = np.zeros(K + 1, dtype=np.int64)
    off[1:] = np.cumsum(np.bincount(pid, minlength=K))
    Xd, yd, od = dev(Xs), dev(ys), dev(off, torch.int64)
    fits = eng.fit_logistic_partitions(Xd, yd, od)
    assert int(fits.status.abs().sum().item()) == 0
    _, H, _ = eng.irls_pass(Xd, yd, od, fits.coef, want_loglik=False)
    assert torch.equal(H, fits.Sig_inv)
    # and the returned coefficients are the MLE to well below the 1e-6 bar
    for k in range(K):
        c0, _, _ = onp.logistic_fit(Xs[off[k]:off[k + 1]], ys[off[k]:off[k + 1]])
        assert np.max(np.abs(fits.coef[k].cpu().numpy() - c0)) < 1e-9
    # optional per-partition log-likelihood at the returned coefficients
    fl = eng.fit_logistic_partitions(Xd, yd, od, config=eng.EngineConfig(want_loglik=True))
    assert torch.equal(fl.coef, fits.coef) and fl.passes == fits.passes + 1
    for k in range(K):
        ref = onp.logistic_loglik(Xs[off[k]:off[k + 1]], ys[off[k]:off[k + 1]], fits.coef[k].cpu().numpy()[:, None])
        assert abs(fl.loglik[k].item() - float(np.ravel(ref)[0])) < 1e-8 * abs(float(np.ravel(ref)[0]))


def test_host_entry_pipelined_blocks_match_device_entry(eng):
    """dlsa_fit_host (host rows in, H2D copies of partition blocks overlapped with the fits of earlier blocks) must
    give what dlsa_fit gives on the same rows already resident on the device."""
    rng = np.random.default_rng(11)
    sizes = rng.integers(300, 900, size=37)
    off = ragged_offsets(sizes)
    n, px = int(off[-1]), 12
    X = rng.random((n, px)) - 0.5
    beta = np.r_[np.ones(4), np.zeros(px - 4)]
    y = (rng.random(n) < 1 / (1 + np.exp(-X @ beta))).astype(np.float64)
    Xh, yh = torch.from_numpy(X).pin_memory(), torch.from_numpy(y).pin_memory()
    offh = torch.from_numpy(off)
    f0, c0, p0 = eng.dlsa_fit(dev(X), dev(y), dev(off, torch.int64), sample_size=n)
    for groups in (1, 5, 64):
        f1, c1, p1 = eng.dlsa_fit_host(Xh, yh, offh, sample_size=n, n_groups=groups)
        assert torch.equal(f1.status, f0.status)
        assert (f1.coef - f0.coef).abs().max().item() < 1e-12
        assert relfro(f1.Sig_inv.cpu().numpy(), f0.Sig_inv.cpu().numpy()) < 1e-13
        assert (c1.beta_byOLS - c0.beta_byOLS).abs().max().item() < 1e-12
        assert p1.argmin_BIC == p0.argmin_BIC and torch.equal(p1.beta != 0, p0.beta != 0)
    with pytest.raises(ValueError):
        eng.dlsa_fit_host(dev(X), yh, offh)


def test_host_entry_stages_pageable_inputs(eng, monkeypatch):
    """Pageable host tensors go through the pinned staging slots (filled by host threads, reused round-robin); tiny
    slots force many pieces per block and every slot to be recycled.  Same results as from the device-resident rows."""
    rng = np.random.default_rng(12)
    sizes = rng.integers(200, 700, size=29)
    off = ragged_offsets(sizes)
    n, px = int(off[-1]), 9
    X = rng.random((n, px)) - 0.5
    y = (rng.random(n) < 1 / (1 + np.exp(-X[:, :3].sum(1)))).astype(np.float64)
    Xh, yh, offh = torch.from_numpy(X), torch.from_numpy(y), torch.from_numpy(off)
    assert not Xh.is_pinned()
    f0, c0, p0 = eng.dlsa_fit(dev(X), dev(y), dev(off, torch.int64), sample_size=n)
    monkeypatch.setattr(eng, "_STAGE_SLOT_BYTES", 40_000)
    monkeypatch.setattr(eng, "_STAGE_N_SLOTS", 3)
    for groups in (1, 4):
        f1, c1, p1 = eng.dlsa_fit_host(Xh, yh, offh, sample_size=n, n_groups=groups)
        assert torch.equal(f1.status, f0.status)
        assert (f1.coef - f0.coef).abs().max().item() < 1e-12
        assert relfro(f1.Sig_inv.cpu().numpy(), f0.Sig_inv.cpu().numpy()) < 1e-13
        assert p1.argmin_BIC == p0.argmin_BIC and torch.equal(p1.beta != 0, p0.beta != 0)
    # a strided (padded-pitch) pageable source
    Xp = torch.zeros(n, px + 3, dtype=torch.float64)
    Xp[:, :px] = Xh
    f2, _, _ = eng.dlsa_fit_host(Xp[:, :px], yh, offh, sample_size=n, n_groups=3)
    assert (f2.coef - f0.coef).abs().max().item() < 1e-12


@pytest.mark.parametrize("px,icpt", [(9, True), (50, False), (100, False), (181, True)])
def test_poisson_pass_and_fit_match_oracle(eng, px, icpt):
    """DLSA_FAMILY_POISSON (R glm(y~x-1, family="poisson") per block, dlsa/R/wlse_func.R:155-168) through the small
    kernel (p <= 56), the tiled kernel with fused link warps (p = 100) and with the streaming link kernel (p = 182)."""
    from dlsa_b200 import _lib
    rng = np.random.default_rng(100 + px)
    sizes = [3000, 1, 2047]
    off = ragged_offsets(sizes)
    n, p = int(off[-1]), px + int(icpt)
    X = rng.random((n, px)) - 0.5
    truth = np.r_[np.ones(min(8, px)), np.zeros(px - min(8, px))] * (0.5 if px > 60 else 1.0)
    y = rng.poisson(np.exp(X @ truth + (0.4 if icpt else 0.0))).astype(np.float64)
    beta = rng.standard_normal((len(sizes), p)) * (1.0 / np.sqrt(p))
    Xd, yd, od = dev(X), dev(y), dev(off, torch.int64)
    g, H, ll = eng.irls_pass(Xd, yd, od, dev(beta), fit_intercept=icpt, family=_lib.FAMILY_POISSON)
    for k, nk in enumerate(sizes):
        Xk = X[off[k]:off[k + 1]]
        if icpt:
            Xk = np.concatenate([np.ones((nk, 1)), Xk], axis=1)
        g0, H0, ll0 = onp.poisson_pass(Xk, y[off[k]:off[k + 1]], beta[k])
        assert np.allclose(g[k].cpu().numpy(), g0, rtol=1e-10, atol=1e-9), k
        assert relfro(H[k].cpu().numpy(), H0) < 1e-12, k
        assert abs(ll[k].item() - ll0) <= 1e-11 * max(1.0, abs(ll0))
    fits = eng.fit_poisson_partitions(Xd, yd, dev(ragged_offsets([3000, 2048]), torch.int64), fit_intercept=icpt)
    assert int(fits.status.abs().sum().item()) == 0
    for k, (lo, hi) in enumerate([(0, 3000), (3000, 5048)]):
        c0, S0, Sc0 = onp.poisson_fit(X[lo:hi], y[lo:hi], fit_intercept=icpt)
        assert np.max(np.abs(fits.coef[k].cpu().numpy() - c0)) < 1e-9
        assert relfro(fits.Sig_inv[k].cpu().numpy(), S0) < SIG_RTOL
        assert np.allclose(fits.Sig_invMcoef[k].cpu().numpy(), Sc0, rtol=1e-8, atol=1e-8)


def test_poisson_dlsa_end_to_end(eng):
    """Poisson DLSA: per-partition glm fits -> one-round combine -> LARS + BIC, against the oracle; the pooled warm
    start (K >= 64) must give the cold-start estimates."""
    rng = np.random.default_rng(8)
    K, nk, p = 64, 1500, 12
    n = K * nk
    X = rng.random((n, p)) - 0.5
    truth = np.r_[np.ones(4), np.zeros(p - 4)]
    y = rng.poisson(np.exp(X @ truth)).astype(np.float64)
    off = np.arange(K + 1, dtype=np.int64) * nk
    fits, comb, path = eng.dlsa_fit(dev(X), dev(y), dev(off, torch.int64), sample_size=n, family="poisson")
    assert int(fits.status.abs().sum().item()) == 0 and fits.pilot_partitions == 8
    ref = [onp.poisson_fit(X[off[k]:off[k + 1]], y[off[k]:off[k + 1]]) for k in range(K)]
    th, one, S = onp.combine(np.stack([r[0] for r in ref]), np.stack([r[1] for r in ref]))
    assert np.max(np.abs(comb.beta_byOLS.cpu().numpy() - th)) < THETA_TOL
    assert relfro(comb.Sig_inv_sum.cpu().numpy(), S) < SIG_RTOL
    rp = onp.lars_lsa(S, th, False, n)
    assert np.array_equal(path.beta.cpu().numpy() != 0, rp["beta"] != 0)
    assert path.argmin_BIC == int(np.argmin(rp["BIC"]))
    assert np.array_equal(path.beta[path.argmin_BIC].cpu().numpy() != 0, truth != 0)


@pytest.mark.parametrize("p", [1, 2, 7, 32, 33, 50, 63, 64, 65, 160, 161, 192, 200, 333, 500, 512])
def test_batched_spd_solve_matches_lapack(eng, p):
    """K5 / the solve inside K2 (np.linalg.lstsq(Sig_inv_sum, ...) of dlsa/dlsa.py:48; solve(hess) of
    dlsa/R/wlse_func.R:55): shared-memory factorisation up to p = 160, blocked DMMA factorisation of the global-memory
    copy above (panels of 32: p = 161, 192, 200, 333 hit ragged last panels and ragged 8x8 tiles).  A matrix that loses
    positive definiteness at a late pivot must come back as ST_NOT_SPD, the others untouched."""
    from dlsa_b200 import _lib
    from dlsa_b200._lib import lib, check
    rng = np.random.default_rng(p)
    K = 5
    A = np.empty((K, p, p))
    for k in range(K):
        M = rng.standard_normal((p + 3, p))
        A[k] = M.T @ M + 0.1 * np.eye(p)
    bad = 3
    if p >= 2:
        A[bad, p - 1, p - 1] = -1.0   # indefinite only in the very last pivot
    b = rng.standard_normal((K, p))
    Ad, bd = dev(A), dev(b)
    x = torch.empty(K, p, dtype=torch.float64, device="cuda")
    st = torch.full((K,), -7, dtype=torch.int32, device="cuda")
    ws = eng._workspace(lib.dlsa_newton_update_workspace_bytes(K, p), Ad.device)
    check(lib.dlsa_spd_solve_f64(Ad.data_ptr(), bd.data_ptr(), K, p, x.data_ptr(), st.data_ptr(), ws.data_ptr(),
                                 ws.numel(), torch.cuda.current_stream().cuda_stream), "dlsa_spd_solve_f64")
    st = st.cpu().numpy()
    for k in range(K):
        if k == bad and p >= 2:
            assert st[k] == _lib.ST_NOT_SPD
            continue
        assert st[k] == _lib.ST_OK
        ref = np.linalg.solve(A[k], b[k])
        assert np.max(np.abs(x[k].cpu().numpy() - ref)) <= 1e-9 * max(1.0, np.max(np.abs(ref))), (k, p)
    assert torch.equal(Ad, dev(A))  # the input matrices are not modified
