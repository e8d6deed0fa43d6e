"""CUDA env step / observation (through the C ABI) against the golden vectors of the Python
reference and against the CPU oracle.  Tolerances (north_star): positions / rewards / real-valued
observation entries within 1e-5 (we assert 1e-9 relative on the fp64 state), every integer output
(collision, done, info, step counters, occupancy bits) bit-exact."""
import numpy as np
import pytest
import torch

import oracle as O
from conftest import episode
from gpu_util import assert_close64, assert_obs, city_and_params

pytestmark = pytest.mark.gpu


def run_golden(g, env_golden, env27_golden, kind_name, apf_v=None):
    from uavrl_b200 import engine
    city, params, _, _ = city_and_params(env_golden, env27_golden)
    ne = int(g["epn_episodes"])
    eps = [episode(g, i) for i in range(ne)]
    K = eps[0]["sub"].shape[0]
    env = engine.EnvBatch(city, params, ne, max_subgoals=K, auto_reset=False)
    if apf_v is not None:
        env.set_extras(obstacle_v=apf_v)
    env.set_pool(np.stack([e["start"] for e in eps]), np.stack([e["goal"] for e in eps]),
                 np.array([e["heading"] for e in eps]), np.stack([e["sub"] for e in eps]),
                 np.array([e["n_sub"] for e in eps]), np.array([e["alias0"] for e in eps]))
    env.reset(0)
    st = env.get_state()
    for i, e in enumerate(eps):        # initial velocity is the reference's bit pattern
        assert st["vx"][i] == e["vx0"] and st["vy"][i] == e["vy0"] and st["V"][i] == e["V0"]
    obs0 = env.observe().cpu().numpy()
    for i, e in enumerate(eps):
        assert_obs(obs0[i], e["obs0"], "obs0 ep%d" % i)
    T = max(len(e["action"]) for e in eps)
    checked = 0
    for t in range(T):
        if kind_name == "continuous":
            a = np.array([e["action"][t] if t < len(e["action"]) else 0.0 for e in eps], np.float64)
            act = torch.tensor(a, dtype=torch.float64, device="cuda")
        else:
            a = np.array([e["action"][t] if t < len(e["action"]) else 13 for e in eps], np.int32)
            act = torch.tensor(a, dtype=torch.int32, device="cuda")
        out = env.step(act)
        st = env.get_state()
        o = {k: v.cpu().numpy() for k, v in out.items()}
        for i, e in enumerate(eps):
            if t >= len(e["action"]):
                continue
            w = "ep%d t%d" % (i, t)
            assert o["done"][i] == e["done_ret"][t] and o["info"][i] == e["info"][t], w
            assert o["collision"][i] == e["collision"][t] and o["ended"][i] == e["done"][t], w
            assert st["step"][i] == e["step"][t] and st["cursor"][i] == e["cursor"][t], w
            assert_close64(st["reward64"][i], e["reward"][t], 1e-9, w + " reward")
            assert abs(o["reward"][i] - e["reward"][t]) <= 1e-5 * max(1.0, abs(e["reward"][t])), w
            for k in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len"):
                assert_close64(st[k][i], e[k][t], 1e-9, w + " " + k)
            assert_obs(o["obs"][i], e["obs"][t], w)
            checked += 1
        if apf_v is not None:                      # Adjust_subgoal: the whole remaining queue, shifted every step
            subs = env.get_subgoals()
            for i, e in enumerate(eps):
                if t >= len(e["action"]):
                    continue
                c, n = int(e["cursor"][t]), int(e["n_sub"])
                assert_close64(subs[i, c:n], e["subq"][t][:n - c], 1e-9, "sub-goal queue ep%d t%d" % (i, t))
    env.close()
    return checked


def test_golden_episodes_continuous(env_golden, env27_golden):
    assert run_golden(env_golden, env_golden, env27_golden, "continuous") > 3000


def test_golden_episodes_discrete27(env_golden, env27_golden):
    assert run_golden(env27_golden, env_golden, env27_golden, "discrete27") > 4000


def test_golden_episodes_apf_moving_obstacles(env_golden, env27_golden):
    """APF_Enabled with moving obstacles (UAV.cal_force / Adjust_subgoal / the reward term, Agents/UAV.py:156-210, 448-453)
    against 1750 steps of the unmodified reference UAV (tests/golden/make_apf_golden.py): every integer output exact, fp64
    state and reward 1e-9, observation 1e-5 / bits exact, and every UAV's shifted sub-goal queue after every step."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "apf_golden.npz"))
    assert np.array_equal(g["buildings"], env_golden["buildings"])
    assert run_golden(g, env_golden, env27_golden, "continuous", apf_v=g["obstacle_v"]) == 1750


def test_apf_against_oracle_with_auto_reset(env_golden, env27_golden):
    """512 UAVs x 120 steps, discrete-27 actions, moving obstacles, in-kernel UAV.reset(): the APF step (per-UAV sub-goal
    queues shifted every step, reloaded from the scenario at a restart) against the oracle, which is pinned bit-exact to
    the reference's APF branch."""
    from uavrl_b200 import engine
    city, params, ocity, oparams = city_and_params(env_golden, env27_golden)
    N, P, T, K = 512, 1024, 120, 64
    rng = np.random.default_rng(31)
    nb = env_golden["buildings"].shape[0]
    vel = np.zeros((nb, 3)); mv = rng.uniform(size=nb) < 0.6
    vel[mv, :2] = rng.normal(0, 1.5, (int(mv.sum()), 2))
    env = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    sc = make_pool(env, P, seed=41)
    env.set_extras(obstacle_v=vel)
    env.reset(0)
    scen = np.arange(N) % P
    ob = O.OracleBatch(ocity, oparams, N, K)
    ob.reset(sc["start"][scen], sc["goal"][scen], sc["heading"][scen], sc["sub"][scen], sc["n_sub"][scen])
    O.set_apf(vel)
    try:
        resets = 0
        for t in range(T):
            a = rng.integers(0, 27, N).astype(np.int32)
            out = env.step(torch.tensor(a, device="cuda"))
            rew, done, info, coll, _ = ob.step_(a.astype(np.float64), O.ACT_DISCRETE27, want_obs=False)
            o = {k: v.cpu().numpy() for k, v in out.items()}
            assert np.array_equal(o["done"], done) and np.array_equal(o["info"], info), t
            assert np.array_equal(o["collision"], coll) and np.array_equal(o["ended"], ob.done), t
            st = env.get_state()
            assert_close64(st["reward64"], rew, 1e-9, "reward t%d" % t)
            resets += oracle_auto_reset(ob, sc, scen, N, P, ocity, oparams, K)
            assert np.array_equal(st["cursor"], ob.cursor) and np.array_equal(st["step"], ob.step), t
            for k in ("px", "py", "pz", "vx", "vy", "V"):
                assert_close64(st[k], getattr(ob, k), 1e-9, "%s t%d" % (k, t))
            assert_obs(o["obs"], ob.state(want64=True)[1], "obs t%d" % t)
            subs = env.get_subgoals()
            for e in range(0, N, 37):
                c, n = int(ob.cursor[e]), int(ob.n_sub[e])
                assert_close64(subs[e, c:n], ob.sub[e, c:n], 1e-9, "queue e%d t%d" % (e, t))
            if t == 0:
                aged = rng.integers(60, 150, N).astype(np.int32)
                env.set_state(step=aged); ob.step[:] = aged
        assert resets > N // 4
    finally:
        O.set_apf(None)
    env.close()


def test_threaten_rate_kat(env_golden, env27_golden):
    from uavrl_b200 import engine
    city, params, _, _ = city_and_params(env_golden, env27_golden)
    env = engine.EnvBatch(city, params, 1, max_subgoals=4)
    got = env.threaten_rate(env_golden["kat_threat_pts"])
    assert np.array_equal(got, env_golden["kat_threat"])      # incl. points within 1 ulp of R / H / bounds


def make_pool(env, P, seed):
    sc = env.make_scenarios(P, seed=seed)
    env.set_pool(sc["start"], sc["goal"], sc["heading"], sc["sub"], sc["n_sub"])
    return sc


@pytest.mark.parametrize("mode", ["continuous_f32", "discrete27"])
def test_against_oracle_1024_envs(env_golden, env27_golden, mode):
    """2048 envs x 170 steps on synthetic scenarios vs the C oracle (which is pinned to the reference)."""
    from uavrl_b200 import engine
    city, params, ocity, oparams = city_and_params(env_golden, env27_golden)
    N, T, K = 1024, 160, 64
    env = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=False)
    sc = make_pool(env, N, seed=11)
    env.reset(0)
    ob = O.OracleBatch(ocity, oparams, N, K)
    ob.reset(sc["start"], sc["goal"], sc["heading"], sc["sub"], sc["n_sub"])
    rng = np.random.default_rng(5)
    n_coll = n_done = 0
    for t in range(T):
        if mode == "continuous_f32":
            a32 = rng.uniform(-1, 1, N).astype(np.float32)
            # half of the envs steer toward the sub-goal so that sub-goal / success branches are hit
            out = env.step(torch.tensor(a32, device="cuda"))
            rew, done, info, coll, oobs = ob.step_(a32.astype(np.float64), O.ACT_CONTINUOUS)
        else:
            a = rng.integers(0, 27, N).astype(np.int32)
            out = env.step(torch.tensor(a, device="cuda"))
            rew, done, info, coll, oobs = ob.step_(a.astype(np.float64), O.ACT_DISCRETE27)
        o = {k: v.cpu().numpy() for k, v in out.items()}
        assert np.array_equal(o["done"], done) and np.array_equal(o["info"], info), t
        assert np.array_equal(o["collision"], coll) and np.array_equal(o["ended"], ob.done), t
        st = env.get_state()
        assert np.array_equal(st["step"], ob.step) and np.array_equal(st["cursor"], ob.cursor), t
        assert_close64(st["reward64"], rew, 1e-9, "reward t%d" % t)
        for k in ("px", "py", "pz", "vx", "vy", "V"):
            assert_close64(st[k], getattr(ob, k), 1e-9, "%s t%d" % (k, t))
        np.testing.assert_allclose(o["reward"], rew, rtol=1e-5, atol=1e-5)
        obs64 = ob.state(want64=True)[1]
        assert_obs(o["obs"], obs64, "obs t%d" % t)
        n_coll += int(coll.sum()); n_done += int(done.sum())
    assert n_coll > 500 and n_done > N      # both branches exercised
    env.close()


def test_auto_reset_and_host_entry_point(env_golden, env27_golden):
    """auto_reset=1: an ended env restarts from scenario (scen + N) mod P inside the step; also drives
    the host-buffer entry point (uavrl_env_step_host) and checks it equals the device-pointer call."""
    from uavrl_b200 import engine
    city, params, ocity, oparams = city_and_params(env_golden, env27_golden)
    N, P, T, K = 256, 1024, 400, 64
    envA = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    envB = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    sc = make_pool(envA, P, seed=3)
    envB.set_pool(sc["start"], sc["goal"], sc["heading"], sc["sub"], sc["n_sub"])
    envA.reset(5); envB.reset(5)
    scen = (5 + np.arange(N)) % P
    ob = O.OracleBatch(ocity, oparams, N, K)
    ob.reset(sc["start"][scen], sc["goal"][scen], sc["heading"][scen], sc["sub"][scen], sc["n_sub"][scen])
    rng = np.random.default_rng(9)
    h_obs = np.zeros((N, 100), np.float32); h_rew = np.zeros(N, np.float32)
    h_done = np.zeros(N, np.uint8); h_info = np.zeros(N, np.uint8); h_coll = np.zeros(N, np.uint8); h_end = np.zeros(N, np.uint8)
    resets = 0
    for t in range(T):
        a = rng.integers(0, 27, N).astype(np.int32)
        out = envA.step(torch.tensor(a, device="cuda"))
        envB.step_host(a, engine.ACT_DISCRETE27, h_obs, h_rew, h_done, h_info, h_coll, h_end)
        o = {k: v.cpu().numpy() for k, v in out.items()}
        assert np.array_equal(o["obs"], h_obs) and np.array_equal(o["reward"], h_rew)
        assert np.array_equal(o["done"], h_done) and np.array_equal(o["ended"], h_end)
        rew, done, info, coll, _ = ob.step_(a.astype(np.float64), O.ACT_DISCRETE27, want_obs=False)
        assert np.array_equal(o["done"], done) and np.array_equal(o["ended"], ob.done)
        np.testing.assert_allclose(o["reward"], rew, rtol=1e-5, atol=1e-5)
        ended = np.nonzero(ob.done)[0]
        if ended.size:                       # UAV.reset() at the episode boundary, oracle side
            scen[ended] = (scen[ended] + N) % P
            sub_all = O.OracleBatch(ocity, oparams, ended.size, K)
            sub_all.reset(sc["start"][scen[ended]], sc["goal"][scen[ended]], sc["heading"][scen[ended]],
                          sc["sub"][scen[ended]], sc["n_sub"][scen[ended]])
            for k in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len", "step", "cursor",
                      "n_sub", "done", "alias0"):
                getattr(ob, k)[ended] = getattr(sub_all, k)
            ob.goal[ended] = sub_all.goal; ob.sub[ended] = sub_all.sub
            resets += ended.size
        st = envA.get_state()
        assert np.array_equal(st["scenario"], scen)
        assert_close64(st["px"], ob.px, 1e-9, "px t%d" % t)
        assert_obs(o["obs"], ob.state(want64=True)[1], "obs t%d" % t)
    assert resets > N
    envA.close(); envB.close()


def oracle_auto_reset(ob, sc, scen, N, P, ocity, oparams, K):
    """UAV.reset() at the episode boundary, oracle side: an ended env restarts from scenario (scen + N) mod P."""
    ended = np.nonzero(ob.done)[0]
    if ended.size:
        scen[ended] = (scen[ended] + N) % P
        fresh = O.OracleBatch(ocity, oparams, ended.size, K)
        fresh.reset(sc["start"][scen[ended]], sc["goal"][scen[ended]], sc["heading"][scen[ended]],
                    sc["sub"][scen[ended]], sc["n_sub"][scen[ended]])
        for k in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len", "step", "cursor",
                  "n_sub", "done", "alias0"):
            getattr(ob, k)[ended] = getattr(fresh, k)
        ob.goal[ended] = fresh.goal; ob.sub[ended] = fresh.sub
    return ended.size


@pytest.mark.parametrize("mode", ["discrete27", "continuous_f32"])
def test_large_batch_kernel_variant_against_oracle(env_golden, env27_golden, mode):
    """N = 20 000 > 16 384 selects env_kernel<true, 32> (32 envs per CTA, one warp of fp64 chains) -- the variant BASELINE
    configs 3-5 (16 384 / 65 536 envs) run.  Values, not invariants: 70 steps incl. in-kernel auto-reset against the C
    oracle, every integer output exact, fp64 state 1e-9, observation bits exact / reals 1e-5.  N is not a multiple of 32."""
    from uavrl_b200 import engine
    city, params, ocity, oparams = city_and_params(env_golden, env27_golden)
    N, P, T, K = 20011, 4096, 70, 64
    env = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    sc = make_pool(env, P, seed=17)
    env.reset(3)
    scen = (3 + np.arange(N)) % P
    ob = O.OracleBatch(ocity, oparams, N, K)
    ob.reset(sc["start"][scen], sc["goal"][scen], sc["heading"][scen], sc["sub"][scen], sc["n_sub"][scen])
    rng = np.random.default_rng(23)
    n_coll = n_done = resets = 0
    for t in range(T):
        if mode == "continuous_f32":
            a32 = rng.uniform(-1, 1, N).astype(np.float32)
            out = env.step(torch.tensor(a32, device="cuda"))
            rew, done, info, coll, _ = ob.step_(a32.astype(np.float64), O.ACT_CONTINUOUS, want_obs=False)
        else:
            a = rng.integers(0, 27, N).astype(np.int32)
            out = env.step(torch.tensor(a, device="cuda"))
            rew, done, info, coll, _ = ob.step_(a.astype(np.float64), O.ACT_DISCRETE27, want_obs=False)
        o = {k: v.cpu().numpy() for k, v in out.items()}
        assert np.array_equal(o["done"], done) and np.array_equal(o["info"], info), t
        assert np.array_equal(o["collision"], coll) and np.array_equal(o["ended"], ob.done), t
        np.testing.assert_allclose(o["reward"], rew, rtol=1e-5, atol=1e-5)
        st = env.get_state()
        assert_close64(st["reward64"], rew, 1e-9, "reward t%d" % t)
        resets += oracle_auto_reset(ob, sc, scen, N, P, ocity, oparams, K)
        assert np.array_equal(st["scenario"], scen), t
        assert np.array_equal(st["step"], ob.step) and np.array_equal(st["cursor"], ob.cursor), t
        for k in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len"):
            assert_close64(st[k], getattr(ob, k), 1e-9, "%s t%d" % (k, t))
        assert_obs(o["obs"], ob.state(want64=True)[1], "obs t%d" % t)
        n_coll += int(coll.sum()); n_done += int(done.sum())
        if t == 0:
            # every episode's first step pops the aliased sub-goal and restarts the segment (Step = 0): age the segments
            # now so that Max_Step ('lose') and the in-kernel UAV.reset() are reached inside this test
            aged = rng.integers(100, 150, N).astype(np.int32)
            env.set_state(step=aged)
            ob.step[:] = aged
    assert n_coll > 1000 and n_done > N and resets > N // 2
    env.close()


def test_full_size_invariants_65536(env_golden, env27_golden):
    """BASELINE config size (65536 envs): size-independent properties.  After any step no UAV sits in
    a threat (collisions revert, UAV.py:425-427), counters stay in range, the centre probes are 0,
    and the run is deterministic."""
    from uavrl_b200 import engine
    city, params, _, _ = city_and_params(env_golden, env27_golden)
    N, P, K = 65536, 4096, 64
    env = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    make_pool(env, P, seed=21)
    g = torch.Generator(device="cuda").manual_seed(1)
    digests = []
    for rep in range(2):
        env.reset(0)
        g.manual_seed(1)
        tot_r = 0.0
        for t in range(60):
            a = torch.randint(0, 27, (N,), generator=g, device="cuda", dtype=torch.int32)
            out = env.step(a)
            tot_r += float(out["reward"].double().sum())
        st = env.get_state()
        pts = np.stack([st["px"], st["py"], st["pz"]], 1)
        assert env.threaten_rate(pts).sum() == 0
        assert (st["step"] >= 0).all() and (st["step"] < params.max_step).all()
        obs = out["obs"]
        assert float(obs[:, [23, 48, 73]].abs().sum()) == 0.0          # probes at offset (0,0)
        assert bool(((obs[:, 11:86] == 0) | (obs[:, 11:86] == 1)).all())
        assert torch.isfinite(obs).all() and np.isfinite(tot_r)
        digests.append((tot_r, float(obs.double().sum()), st["px"].sum()))
    assert digests[0] == digests[1]
    env.close()


def test_device_pool_generator_equals_host_generator(env_golden, env27_golden):
    """uavrl_env_generate_pool (one GPU thread per scenario: reset draws + RRT) writes the SAME pool as the
    host generator + uavrl_env_set_pool for the same seed -- every start / goal / sub-goal bit-for-bit -- and envs
    stepped from it follow the same trajectories."""
    import time
    from uavrl_b200 import engine
    city, params, _, _ = city_and_params(env_golden, env27_golden)
    N, P, K = 512, 1024, 64
    a = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    b = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    sc = a.make_scenarios(P, seed=31)
    a.set_pool(sc["start"], sc["goal"], sc["heading"], sc["sub"], sc["n_sub"])
    b.generate_pool(P, seed=31)
    pa, pb = a.get_pool(), b.get_pool()
    for k in ("start", "goal", "sub", "n_sub"):
        assert np.array_equal(pa[k], pb[k]), k
        assert np.array_equal(pb[k], sc[k]), k
    np.testing.assert_allclose(pb["v0"], pa["v0"], rtol=0, atol=4e-16)      # cos/sin: CUDA libm vs glibc
    assert (pb["n_sub"] >= 2).all() and (pb["n_sub"] <= K).all()
    a.reset(0); b.reset(0)
    g = torch.Generator(device="cuda").manual_seed(3)
    for t in range(200):
        act = torch.randint(0, 27, (N,), generator=g, device="cuda", dtype=torch.int32)
        oa, ob = a.step(act), b.step(act)
        for k in ("done", "info", "collision", "ended"):
            assert torch.equal(oa[k], ob[k]), (k, t)
        assert torch.allclose(oa["obs"], ob["obs"], rtol=0, atol=1e-6)
    # throughput of the generator itself, device vs host threads (printed with -s)
    t0 = time.perf_counter(); b.generate_pool(8192, seed=5); t_dev = time.perf_counter() - t0
    t0 = time.perf_counter(); a.make_scenarios(8192, seed=5); t_host = time.perf_counter() - t0
    print("\npool of 8192 scenarios: device %.1f ms, host threads %.1f ms" % (1e3 * t_dev, 1e3 * t_host))
    assert np.array_equal(b.get_pool()["n_sub"], a.make_scenarios(8192, seed=5)["n_sub"])
    a.close(); b.close()


@pytest.mark.parametrize("prefix", ["c_", "d_"])
def test_single_steps_next_to_every_decision_boundary(env_golden, env27_golden, prefix):
    """The CUDA step from constructed states next to every decision boundary (tests/golden/step_golden.npz, 3000 reference
    steps per action kind): states injected with uavrl_env_set_state, every integer output exact, fp64 state 1e-9."""
    import os
    from conftest import ROOT
    from uavrl_b200 import engine
    g = np.load(os.path.join(ROOT, "tests", "golden", "step_golden.npz"))
    k = lambda s: g[prefix + s]                          # noqa: E731
    city, params, _, _ = city_and_params(env_golden, env27_golden)
    n = len(k("reward"))
    env = engine.EnvBatch(city, params, n, max_subgoals=k("sub").shape[1], auto_reset=False)
    start = np.stack([k("px"), k("py"), k("pz")], 1)
    env.set_pool(start, k("goal"), np.zeros(n), k("sub"), k("n_sub"), k("alias0"))
    env.reset(0)
    env.set_state(px=k("px"), py=k("py"), pz=k("pz"), vx=k("vx"), vy=k("vy"), V=k("V"), step=k("step"), score=k("score"),
                  total_score=k("total_score"), path_len=k("path_len"))
    if prefix == "c_":
        act = torch.tensor(k("action").astype(np.float64), dtype=torch.float64, device="cuda")
    else:
        act = torch.tensor(k("action").astype(np.int32), dtype=torch.int32, device="cuda")
    out = {kk: v.cpu().numpy() for kk, v in env.step(act).items()}
    st = env.get_state()
    for got, want, what in ((out["done"], k("done_ret"), "done"), (out["info"], k("info"), "info"), (out["collision"], k("collision"), "collision"),
                            (out["ended"], k("o_done"), "ended"), (st["step"], k("o_step"), "step"), (st["cursor"], k("o_cursor"), "cursor")):
        assert np.array_equal(got, want), (what, int((got != want).sum()))
    assert_close64(st["reward64"], k("reward"), 1e-9, "reward")
    for f in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len"):
        assert_close64(st[f], k("o_" + f), 1e-9, f)
    assert_obs(out["obs"], k("obs"), "obs")
    env.close()


POWER = dict(P_i=89.0, v_0=4.05, d_0=0.6, rho=1.225, s=0.05, A=0.5, P_b=79.0, F_b=120.0, xi=0.8)     # config/UAV.xml <Fly_power>


def test_energy_model_and_trajectory_recording(env_golden, env27_golden):
    """uavrl_env_set_extras: the energy column is the running sum of Calc_Fly_Power(V) (Agents/UAV.py:239-245, oracle
    ora_fly_power pinned to the reference formula) over the episode's steps, and the recorded trajectory is UAV.path
    (UAV.py:432: the position after every step, post collision revert), double-buffered across the in-kernel reset.
    The step outputs themselves are unchanged by the extras (same kernel body, EXTRAS instantiation)."""
    from uavrl_b200 import engine
    city, params, ocity, oparams = city_and_params(env_golden, env27_golden)
    N, P, K, T, CAP = 200, 512, 64, 260, 512
    env = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    ref = engine.EnvBatch(city, params, N, max_subgoals=K, auto_reset=True)
    sc = make_pool(env, P, seed=29)
    ref.set_pool(sc["start"], sc["goal"], sc["heading"], sc["sub"], sc["n_sub"])
    env.set_extras(power=POWER, track_envs=8, track_capacity=CAP)
    env.reset(0); ref.reset(0)
    assert np.array_equal(env.get_energy(), np.zeros(N))
    lib = O.lib()
    fp = lambda V: lib.ora_fly_power(float(V), *[POWER[k] for k in ("P_i", "v_0", "d_0", "rho", "s", "A", "P_b", "F_b", "xi")])  # noqa: E731
    rng = np.random.default_rng(2)
    energy = np.zeros(N)
    paths = [[] for _ in range(8)]
    last = [None] * 8
    for t in range(T):
        a = rng.integers(0, 27, N).astype(np.int32)
        if t == 1:
            env.set_state(step=np.full(N, 120, np.int32)); ref.set_state(step=np.full(N, 120, np.int32))      # episodes end inside the test
        out, out_ref = env.step(torch.tensor(a, device="cuda")), ref.step(torch.tensor(a, device="cuda"))
        for k in ("obs", "reward", "done", "info", "collision", "ended"):
            assert torch.equal(out[k], out_ref[k]), (k, t)
        st = env.get_state()
        ended = out["ended"].cpu().numpy().astype(bool)
        # energy: + P(V of this step); an ended env restarts with 0.  V of an env that just restarted is the new episode's
        # initial speed, so take this step's speed from |V_vector| before the reset: max_v or one of the speed levels
        l = a % 3
        speed = np.where(l == 0, params.min_v, np.where(l == 1, (params.min_v + params.max_v) / 2, params.max_v))
        V_step = np.where(ended, speed, st["V"])
        energy = np.where(ended, 0.0, energy + np.array([fp(v) for v in V_step]))
        np.testing.assert_allclose(env.get_energy(), energy, rtol=1e-12, atol=1e-9)
        for e in range(8):
            if ended[e]:
                last[e] = paths[e]          # the final position of the episode is appended by the kernel before the reset
                paths[e] = []
            else:
                paths[e].append([st["px"][e], st["py"][e], st["pz"][e]])
    n_checked = 0
    for e in range(8):
        cur = env.get_path(e, 0)
        assert cur.shape[0] == len(paths[e])
        if len(paths[e]):
            np.testing.assert_array_equal(cur, np.array(paths[e]))
        if last[e] is not None:
            prev = env.get_path(e, 1)
            assert prev.shape[0] == len(last[e]) + 1          # + the terminal step's position (recorded before UAV.reset)
            if len(last[e]):
                np.testing.assert_array_equal(prev[:-1], np.array(last[e]))
            n_checked += 1
    assert n_checked >= 4 and energy.max() > 0
    env.close(); ref.close()
