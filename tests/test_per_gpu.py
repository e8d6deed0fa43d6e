"""Prioritised replay on the device (csrc/per.cu) against the reference's own SumTree / ReplayTree
(tests/golden/per_golden.npz, produced by executing BaseClass/replay_buffer.py) and against the CPU oracle:
leaf selection exact on the recorded uniform tape, importance weights, beta schedule, push / batch_update rules,
the ring wrap-around and the rotated leaf order of a non-power-of-two capacity; then the integrated update path."""
import os

import numpy as np
import pytest
import torch

import oracle as O
from conftest import ROOT
from gpu_util import city_and_params

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_learner(cap, batch=32, lockstep=0, algo=None, hidden=(64, 64)):
    from uavrl_b200 import engine
    L = engine.Learner(100, list(hidden), 27, False, engine.ALGO_DDQN if algo is None else algo, batch_size=batch,
                       replay_capacity=cap, lockstep_envs=lockstep, seed=5, update_loop=3)
    L.init_params(0)
    L.per_enable()
    return L


def push_n(L, n, rng):
    obs = torch.tensor(rng.standard_normal((n, 100)), dtype=torch.float32, device=DEV)
    nxt = torch.tensor(rng.standard_normal((n, 100)), dtype=torch.float32, device=DEV)
    a = torch.tensor(rng.integers(0, 27, n), dtype=torch.int32, device=DEV)
    r = torch.tensor(rng.standard_normal(n), dtype=torch.float32, device=DEV)
    d = torch.tensor(rng.integers(0, 2, n), dtype=torch.uint8, device=DEV)
    L.push(obs, a, r, nxt, d)
    return obs, a, r, nxt, d


@pytest.mark.parametrize("case", ["p2", "np2", "part"])
def test_per_matches_reference_structures(case):
    g = np.load(os.path.join(ROOT, "tests", "golden", "per_golden.npz"))
    k = lambda s: g["%s_%s" % (case, s)]                       # noqa: E731
    cap, B = int(k("cap")), int(k("B"))
    L = make_learner(cap, batch=B)
    rng = np.random.default_rng(0)
    err = k("push_err")
    # ReplayTree.push in ring order (chunks that never exceed the capacity): priority (|e| + eps)^alpha, float32
    done = 0
    while done < len(err):
        n = min(200, len(err) - done)
        push_n(L, n, rng)
        slots = torch.tensor((np.arange(done, done + n) % cap).astype(np.int32), device=DEV)
        L.per_set_errors(slots, torch.tensor(err[done:done + n], device=DEV), clip=False)
        done += n
    leaves, total, beta = L.per_state(cap)
    np.testing.assert_allclose(leaves, k("leaves_after_push"), rtol=3e-7, atol=0)
    assert L.replay_size() == int(k("n_entries")) and beta == 0.4
    # continue from the reference's exact leaves so that index equality is not blurred by float32 pow ulps
    L.per_set_priorities(torch.arange(cap, dtype=torch.int32, device=DEV), torch.tensor(k("leaves_after_push"), device=DEV))
    leaves, total, _ = L.per_state(cap)
    assert np.array_equal(leaves, k("leaves_after_push"))
    assert abs(total - float(k("total0"))) <= 1e-12 * total
    slots, w = L.per_sample(B, torch.tensor(k("u0"), dtype=torch.float64, device=DEV))
    tree_idx = slots.cpu().numpy().astype(np.int64) + cap - 1
    assert np.array_equal(tree_idx, k("idx0"))                # SumTree.get_leaf on the same draws: same leaves
    np.testing.assert_allclose(w.cpu().numpy(), k("w0"), rtol=2e-6)
    assert L.per_state(cap)[2] == float(k("beta0"))
    L.per_set_errors(slots, torch.tensor(k("abs_err"), device=DEV), clip=True)       # ReplayTree.batch_update
    leaves, _, _ = L.per_state(cap)
    np.testing.assert_allclose(leaves, k("leaves_after_update"), rtol=3e-7, atol=0)
    L.per_set_priorities(torch.arange(cap, dtype=torch.int32, device=DEV), torch.tensor(k("leaves_after_update"), device=DEV))
    slots, w = L.per_sample(B, torch.tensor(k("u1"), dtype=torch.float64, device=DEV))
    assert np.array_equal(slots.cpu().numpy().astype(np.int64) + cap - 1, k("idx1"))
    np.testing.assert_allclose(w.cpu().numpy(), k("w1"), rtol=2e-6)
    assert L.per_state(cap)[2] == float(k("beta1"))
    L.close()


def test_per_large_vs_oracle_and_distribution():
    """100 000-slot ring (non power of two) with random priorities: 4096 stratified draws pick exactly the oracle's
    leaves; without a tape the empirical sampling frequencies follow p / total."""
    cap, B = 100_000, 4096
    L = make_learner(cap, batch=64)
    rng = np.random.default_rng(3)
    for _ in range(cap // 10_000):
        push_n(L, 10_000, rng)
    prio = np.ldexp(rng.integers(1, 4096, cap).astype(np.float64), -10)         # dyadic: every partial sum is exact
    prio[rng.integers(0, cap, 500)] = 0.0
    L.per_set_priorities(torch.arange(cap, dtype=torch.int32, device=DEV), torch.tensor(prio, device=DEV))
    per = O.OraclePer(cap)
    per.add(prio)
    u = rng.random(B)
    idx_o, w_o, beta_o = per.sample(u)
    slots, w = L.per_sample(B, torch.tensor(u, device=DEV))
    assert np.array_equal(slots.cpu().numpy().astype(np.int64) + cap - 1, idx_o)
    np.testing.assert_allclose(w.cpu().numpy(), w_o, rtol=2e-6)
    assert L.per_state(cap)[2] == beta_o
    # no tape: Philox draws; frequencies ~ p / total over 64 x 4096 samples (groups of 1000 slots)
    counts = np.zeros(cap)
    for _ in range(64):
        s, _ = L.per_sample(B)
        counts += np.bincount(s.cpu().numpy(), minlength=cap)
    assert counts[prio == 0].sum() == 0
    got = counts.reshape(100, 1000).sum(1) / counts.sum()
    want = prio.reshape(100, 1000).sum(1) / prio.sum()
    assert np.abs(got - want).max() < 0.1 * want.max()
    L.close()


@pytest.mark.parametrize("tc", [1, 0])
def test_weighted_update_matches_oracle(tc):
    """Trainer.update with 'weights': loss = mean(w (Q - y)^2), |Q - y| returned -- CUDA (tensor-core and fp32 paths)
    vs the CPU oracle's weighted restatement, 3 consecutive updates incl. a hard update."""
    from uavrl_b200 import engine
    B = 96
    L = make_learner(4096, batch=B, algo=engine.ALGO_DDQN)
    L.set_tensor_cores(bool(tc))
    net = O.make_net(100, [64, 64], 27, 0)
    ol = O.OracleLearner(net, O.ALGO_DDQN, L.get_params(0), gamma=0.99, lr=5e-4, update_loop=3)
    ol.target[:] = L.get_params(1)                           # init_params draws q_target independently, like the reference
    rng = np.random.default_rng(11)
    for it in range(3):
        s = rng.standard_normal((B, 100)).astype(np.float32); s2 = rng.standard_normal((B, 100)).astype(np.float32)
        a = rng.integers(0, 27, B).astype(np.int32); r = rng.standard_normal(B).astype(np.float32)
        d = (rng.random(B) < 0.2).astype(np.float32); w = rng.uniform(0.1, 1.0, B).astype(np.float32)
        loss_o, _, ae_o = ol.update(s, a, r, s2, d, is_w=w)
        t = lambda x: torch.tensor(x, device=DEV)            # noqa: E731
        ae = torch.zeros(B, device=DEV); loss = torch.zeros(1, device=DEV)
        L.update_batch_per(t(s), t(a), t(r), t(s2), t(d), t(w), ae, loss)
        np.testing.assert_allclose(float(loss), loss_o, rtol=3e-5)
        np.testing.assert_allclose(ae.cpu().numpy(), ae_o, rtol=1e-4, atol=2e-5)
        np.testing.assert_allclose(L.get_params(0), ol.local, atol=2e-5)
        np.testing.assert_allclose(L.get_params(1), ol.target, atol=2e-5)
    L.close()


def test_lockstep_loop_with_prioritised_replay(env_golden, env27_golden):
    """uavrl_train_run with PER on: every stored transition carries a positive priority, the frame that receives the
    next observations carries none, sampled batches refresh their priorities with the batch_update rule (values in
    [eps^alpha, 1]), the loop trains (finite, changing loss) and beta follows the schedule."""
    from uavrl_b200 import engine
    city, params, _, _ = city_and_params(env_golden, env27_golden)
    N, F = 256, 12
    env = engine.EnvBatch(city, params, N, max_subgoals=64, auto_reset=True)
    env.generate_pool(512, seed=2)
    env.reset(0)
    L = engine.Learner(100, [64, 64], 27, False, engine.ALGO_DDQN, batch_size=N, replay_capacity=N * F, lockstep_envs=N,
                       seed=1, update_loop=3)
    L.init_params(0)
    L.per_enable()
    slots = N * (F + 1)
    st = engine.train_run(env, L, 40, eps=0.3)
    leaves, total, beta = L.per_state(slots)
    assert st.updates == 39 and np.isfinite(st.last_loss)
    assert abs(beta - min(1.0, 0.4 + 0.001 * st.updates)) < 1e-12
    lv = leaves.reshape(F + 1, N)
    empty = np.where((lv == 0).all(1))[0]
    assert len(empty) == 1                                       # exactly the head frame
    filled = np.delete(lv, empty[0], axis=0)
    p0 = float(np.float32(0.01) ** np.float32(0.6))
    assert filled.min() >= p0 * (1 - 1e-6) and filled.max() <= 1.0 + 1e-6
    assert (np.abs(filled - p0) > 1e-9).sum() > N                # sampled transitions were re-prioritised
    assert abs(total - leaves.sum()) <= 1e-9 * total
    p_before = L.get_params(0).copy()
    engine.train_run(env, L, 5, eps=0.3)
    assert not np.array_equal(p_before, L.get_params(0))
    env.close(); L.close()
