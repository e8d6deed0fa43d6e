import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def env_golden():
    return np.load(os.path.join(GOLDEN, "env_golden.npz"))


@pytest.fixture(scope="session")
def env27_golden():
    return np.load(os.path.join(GOLDEN, "env27_golden.npz"))


@pytest.fixture(scope="session")
def dqn_golden():
    return np.load(os.path.join(GOLDEN, "dqn_golden.npz"))


def episode(g, i):
    """Unpack episode i of a golden file into a dict."""
    pre = "ep%d_" % i
    return {k[len(pre):]: g[k] for k in g.files if k.startswith(pre)}


def step_batch(g, prefix, ocity, oparams):
    """An OracleBatch holding the pre-states of tests/golden/step_golden.npz (prefix 'c_' continuous / 'd_' discrete-27)."""
    import oracle as O
    k = lambda s: g[prefix + s]                          # noqa: E731
    n = len(k("reward"))
    b = O.OracleBatch(ocity, oparams, n, k("sub").shape[1])
    for f in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len"):
        getattr(b, f)[:] = k(f)
    b.step[:] = k("step"); b.cursor[:] = 0; b.n_sub[:] = k("n_sub"); b.done[:] = 0; b.alias0[:] = k("alias0")
    b.goal[:] = k("goal"); b.sub[:] = k("sub")
    return b, n


def check_step_outputs(g, prefix, b, rew, done, info, coll, obs64=None, obs32=None, exact=True):
    """Compare a stepped batch with the reference's recorded outputs."""
    k = lambda s: g[prefix + s]                          # noqa: E731
    assert np.array_equal(done, k("done_ret")) and np.array_equal(info, k("info")) and np.array_equal(coll, k("collision"))
    assert np.array_equal(b.step, k("o_step")) and np.array_equal(b.cursor, k("o_cursor")) and np.array_equal(b.done, k("o_done"))
    pairs = [(rew, k("reward"))] + [(getattr(b, f), k("o_" + f)) for f in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len")]
    for got, want in pairs:
        if exact:
            assert np.array_equal(got, want)
        else:
            assert (np.abs(got - want) <= 1e-12 * np.maximum(1.0, np.abs(want))).all()
    if obs64 is not None:
        assert np.array_equal(obs64, k("obs"))
    if obs32 is not None:
        want = k("obs")
        np.testing.assert_allclose(obs32, want.astype(np.float32), rtol=0, atol=1e-6)
        assert np.array_equal(obs32[:, 11:86], want[:, 11:86]) and np.array_equal(obs32[:, 90:95], want[:, 90:95])
