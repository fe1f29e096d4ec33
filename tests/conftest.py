import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def games_data():
    """The reference's games-expand fixture (projects/results/data/games-expand.csv) unpacked from
    tests/golden/games_expand_k5.npz: returns (X[n,10], y[n], K, golden)."""
    g = load_golden("games_expand_k5.npz")
    shape = tuple(int(v) for v in g["in_shape"])
    raw = np.unpackbits(g["in_bits"])[: shape[0] * shape[1]].reshape(shape).astype(np.float64)
    return raw[:, :-1].copy(), raw[:, -1].copy(), int(g["in_K"]), g


def sim_data(g):
    """Regenerate the seeded simulated inputs a sim_* golden was minted from; checks the checksum."""
    from oracle import numpy_dlsa as onp
    n, p, K, seed = int(g["in_n"]), int(g["in_p"]), int(g["in_K"]), int(g["in_seed"])
    pid, y, X, _ = onp.simulate_logistic(n, p, "systematic", K, seed=seed)
    chk = np.array([X.sum(), (X * X).sum(), y.sum()])
    assert np.allclose(chk, g["in_X_checksum"], rtol=0, atol=1e-6), "numpy RNG stream drifted from the golden recipe"
    return pid.astype(np.int64), y, X, K
