"""Reference-facing plug-ins on the GPU: the trainer classes reproduce the reference trainers'
results through the reference's own method names (get_action / update(transition_dict) / save /
Load_Mod / hard_update), and the env plug-in runs an episode from the XML configs exactly the way
simulator.py drives the reference (env.run_eposide(eps))."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu

BASE = {"w": "100", "hiden_dim": "64", "output": "27", "name": "t0", "LEARNING_RATE": "0.0005", "Batch_Size": "64",
        "gamma": "0.99", "save_loop": "1000000000", "replay_size": "1000", "Update_loop": "3", "Is_Train": "1"}
CASES = [("dueling_vanet2", "DuelingDQN_Trainer_B200", "VAnet2"), ("ddqn_qvalue3", "DDQN_Trainer_B200", "QValueNet_SAC"),
         ("dqn_qnet2", "DQN_Trainer_B200", "Qnet2")]


def make(ttype, net, tmp, **kw):
    import importlib
    mod = importlib.import_module("uavrl_b200.plugins." + ttype)
    p = dict(BASE, Trainer_Type=ttype, NetWork=net, model_path=str(tmp), **kw)
    return getattr(mod, ttype)(p)


@pytest.mark.parametrize("name,ttype,net", CASES)
def test_trainer_plugin_matches_reference_trainer(dqn_golden, tmp_path, name, ttype, net):
    g = dqn_golden
    tr = make(ttype, net, tmp_path)
    tr._learner.set_params(g[name + "_local0"], 0)
    tr._learner.set_params(g[name + "_target0"], 1)
    snap = list(g[name + "_snap"])
    for step in range(10):
        td = {"states": g["batch_s"][step].tolist(), "actions": g["batch_a"][step].tolist(),
              "next_states": g["batch_s2"][step].tolist(), "rewards": g["batch_r"][step].tolist(),
              "dones": g["batch_d"][step].tolist()}
        res = tr.update(td)                                     # the reference's call (PathPlan_City.py:757-776)
        assert res["sum_epoch"] == step + 1 == tr.epoch
        assert np.isclose(float(res["loss"]), g[name + "_loss"][step], rtol=2e-5)
        if step in snap:
            k = snap.index(step)
            flat = np.concatenate([v.numpy().ravel() for v in tr.state_dict(0).values()])
            np.testing.assert_allclose(flat, g[name + "_local"][k], atol=2e-5)
            flat_t = np.concatenate([v.numpy().ravel() for v in tr.state_dict(1).values()])
            np.testing.assert_allclose(flat_t, g[name + "_target"][k], atol=2e-5)
    # get_action: greedy with eps = 0, single state -> python int, batch -> array
    q_ref = g[name + "_q_final"]
    a = tr.get_action(g["batch_s"][0], 0.0)
    assert np.array_equal(a, q_ref.argmax(1))
    assert tr.get_action(g["batch_s"][0][3], 0.0) == int(q_ref[3].argmax())
    assert tr.update({"states": [], "actions": [], "next_states": [], "rewards": [], "dones": []})["sum_epoch"] == 11
    # checkpoint round trip in the reference's .pth format, loadable by torch modules of the reference's shape
    tr.save()
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 2 and all(f.endswith("_t0.pth") for f in files)
    ck = torch.load(os.path.join(tmp_path, [f for f in files if f.startswith("q_local")][0]), weights_only=False)
    assert set(ck) == {"model", "optimizer", "epoch"} and ck["epoch"] == 11
    keys = list(ck["model"])
    assert keys[0] == "fc1.weight" and ck["model"]["fc1.weight"].shape == (64, 100)
    assert ck["optimizer"]["param_groups"][0]["betas"] == (0.9, 0.999)
    tr2 = make(ttype, net, tmp_path)                            # Load_Mod() in the constructor resumes
    assert tr2.epoch == 11
    assert np.array_equal(tr2._learner.get_params(0), tr._learner.get_params(0))
    assert np.array_equal(tr2._learner.get_params(1), tr._learner.get_params(1))
    assert np.array_equal(tr2._learner.get_params(2), tr._learner.get_params(2))
    assert tr2._learner.counters() == tr._learner.counters()
    tr2.hard_update()
    assert np.array_equal(tr2._learner.get_params(1), tr2._learner.get_params(0))


def test_replay_facade_matches_reference_flow(dqn_golden, tmp_path):
    """replay_memory.add / len(buffer) / sample2 / Push_Replay as PathPlan_City.run_thread_OffPolicy uses them."""
    g = dqn_golden
    tr = make("DDQN_Trainer_B200", "QValueNet_SAC", tmp_path)
    s, a, r, s2, d = (g["batch_" + k][0] for k in ("s", "a", "r", "s2", "d"))
    for i in range(70):
        tr.replay_memory.add(s[i % 64], int(a[i % 64]), float(r[i % 64]), s2[i % 64], bool(d[i % 64]))
    assert len(tr.replay_memory.buffer) == 70
    tr.Push_Replay((torch.tensor(s[:1]), torch.tensor([[int(a[0])]]), torch.tensor([[r[0]]]), torch.tensor(s2[:1]),
                    torch.tensor([[d[0]]])))
    assert len(tr.replay_memory.buffer) == 71 > tr.Batch_Size
    b_s, b_a, b_r, b_ns, b_d, idx, w = tr.replay_memory.sample2(tr.Batch_Size)
    assert b_s.shape == (64, 100) and len(b_a) == 64 and idx is None and w is None
    res = tr.update({'states': b_s, 'actions': b_a, 'next_states': b_ns, 'rewards': b_r, 'dones': b_d})
    assert np.isfinite(float(res["loss"]))
    assert tr.learn_off_policy()["sum_epoch"] == 2


def test_env_plugin_runs_an_episode_from_xml(tmp_path):
    """EnvFactory-style construction from the XML files + simulator-style driving."""
    import importlib
    from uavrl_b200.plugins import xmlconfig
    cwd = os.getcwd()
    os.chdir(ROOT)                                   # the reference resolves ./config/... relative to CWD
    try:
        cfg = xmlconfig.XML2Dict(os.path.join(ROOT, "configs", "PathPlan_City_B200.xml"))["simulator"]
        env_dict = cfg["env"]
        env_dict["num_UAV"] = "256"
        env_dict["scenario_pool"] = "512"
        mod = importlib.import_module("uavrl_b200.plugins." + env_dict["Env_Type"])
        env = getattr(mod, env_dict["Env_Type"])(env_dict)           # FactoryClass/EnvFactory.py:17-20
    finally:
        os.chdir(cwd)
    env.Trainer.save_loop = 10 ** 9
    env.Trainer.Batch_Size = 4096
    assert (env.len, env.width, env.h) == (500, 500, 100) and len(env.buildings_table) == 26
    assert env.Threaten_rate(type("L", (), {"x": -1.0, "y": 5.0, "z": 0.0})()) == 1
    eps = xmlconfig.epsilon_annealing(1, float(cfg["min_eps"]), int(cfg["max_eps_episode"]))
    info = env.run_eposide(eps)
    for k in ("average_score", "loss", "score", "success", "lose", "meet_threaten", "normal", "sum_epoch", "eps", "step"):
        assert k in info                                              # the keys simulator.py / record() read
    assert info["episodes"] >= 256 and info["env_steps"] == 256 * info["step"]
    assert info["success"] + info["lose"] + info["normal"] == info["env_steps"]
    assert np.isfinite(info["loss"]) and "%.3f" % info["average_score"]
    assert env.Agents[0].Train_time > 0 and env.Agents[0].Testing_time == 0
    s = env.states()
    ns, r, d, infos = env.Move_Agents(np.full(256, 13, np.int32))
    assert s.shape == ns.shape == (256, 100) and set(infos) <= {"normal", "success", "lose"}


def test_sac_plugin_and_shipped_configuration(tmp_path):
    """The configuration the reference ships (SAC_Trainer + continuous update_PathPlan, config/Trainer.xml + UAV.xml),
    expressed with the B200 plug-ins: trainer API (get_action -> [a0, a1], update(transition_dict), save / Load_Mod in
    the reference's <role>_SAC_<name>.pth files) and an episode of the env plug-in."""
    import importlib
    from uavrl_b200.plugins import xmlconfig
    sac_golden = np.load(os.path.join(ROOT, "tests", "golden", "sac_golden.npz"))
    tdict = xmlconfig.XML2Dict(os.path.join(ROOT, "configs", "Trainer_SAC_B200.xml"))["Trainer"]
    tdict.update(name="UAV_0", model_path=str(tmp_path), Batch_Size="64", replay_size="4096")
    mod = importlib.import_module("uavrl_b200.plugins.SAC_Trainer_B200")
    tr = mod.SAC_Trainer_B200(tdict)
    g = sac_golden
    a = tr.get_action(g["sac_s"][0][0], 0.1)
    assert isinstance(a, list) and len(a) == 2 and all(-1 < x < 1 for x in a)
    td = {"states": g["sac_s"][0].tolist(), "actions": g["sac_a"][0].tolist(), "next_states": g["sac_s2"][0].tolist(),
          "rewards": g["sac_r"][0].tolist(), "dones": g["sac_d"][0].tolist()}
    res = tr.update(td)
    assert res["sum_epoch"] == 1 and np.isfinite(float(res["loss"]))
    assert tr.update({"states": [[]], "actions": [], "next_states": [], "rewards": [], "dones": []})["sum_epoch"] == 2
    tr.save()
    assert sorted(os.listdir(tmp_path)) == ["actor_SAC_UAV_0.pth", "critic_1_SAC_UAV_0.pth", "critic_2_SAC_UAV_0.pth"]
    ck = torch.load(os.path.join(tmp_path, "critic_1_SAC_UAV_0.pth"), weights_only=False)
    assert list(ck["model"]) == ["fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc_out.weight", "fc_out.bias"]
    assert ck["model"]["fc1.weight"].shape == (64, 102) and ck["model"]["fc_out.weight"].shape == (2, 64)   # the 2-wide critic
    tr2 = mod.SAC_Trainer_B200(tdict)
    for role in range(3):
        assert np.array_equal(tr2._learner.get_params(role), tr._learner.get_params(role))
        assert np.array_equal(tr2._learner.get_params(5 + role), tr._learner.get_params(5 + role))
    assert tr2.epoch == 2
    # env plug-in with the continuous step + SAC trainer
    cwd = os.getcwd(); os.chdir(ROOT)
    try:
        env_dict = xmlconfig.XML2Dict(os.path.join(ROOT, "configs", "PathPlan_City_B200.xml"))["simulator"]["env"]
        env_dict.update(num_UAV="128", scenario_pool="256")
        env_dict["Agent"]["xml_path_agent"] = "./configs/UAV_continuous_B200.xml"
        env_dict["Agent"]["Trainer"]["Trainer_path"] = "./configs/Trainer_SAC_B200.xml"
        envmod = importlib.import_module("uavrl_b200.plugins.PathPlan_City_B200")
        env = envmod.PathPlan_City_B200(env_dict)
    finally:
        os.chdir(cwd)
    info = env.run_eposide(0.1)
    assert info["episodes"] >= 128 and info["env_steps"] == 128 * info["step"] and np.isfinite(info["loss"])
    ns, r, d, infos = env.Move_Agents(np.zeros(128, np.float32))
    assert ns.shape == (128, 100)


def test_prioritised_replay_through_the_trainer_plugin():
    """IsPriority_Replay = 1 (Trainer.xml:4): Push_Replay(exp, error), replay_memory.sample2 -> (..., idx, weights),
    update(transition_dict) with them (loss weighted, priorities refreshed like ReplayTree.batch_update) and
    learn_off_policy sampling through the tree."""
    import importlib
    from uavrl_b200.plugins import xmlconfig
    tdict = xmlconfig.XML2Dict(os.path.join(ROOT, "configs", "Trainer_DDQN_B200.xml"))["Trainer"]
    tdict.update(name="UAV_0", IsPriority_Replay="1", Batch_Size="64", replay_size="1000", save_loop="1000000", model_path="/nonexistent")
    mod = importlib.import_module("uavrl_b200.plugins.DDQN_Trainer_B200")
    tr = mod.DDQN_Trainer_B200(tdict)
    rng = np.random.default_rng(0)
    err = rng.gamma(1.5, 0.4, 300).astype(np.float32)
    for i in range(300):
        tr.Push_Replay((rng.standard_normal((1, 100)).astype(np.float32), [int(rng.integers(0, 27))], [[float(rng.standard_normal())]],
                        rng.standard_normal((1, 100)).astype(np.float32), [[False]]), torch.tensor(err[i]))
    leaves, total, beta = tr._learner.per_state(1000)
    np.testing.assert_allclose(leaves[:300], (err + np.float32(0.01)) ** np.float32(0.6), rtol=3e-6)
    assert (leaves[300:] == 0).all() and beta == 0.4
    s, a, r, s2, d, idx, w = tr.replay_memory.sample2(64)
    assert s.shape == (64, 100) and len(idx) == 64 and w.shape == (64,) and abs(w.max() - 1.0) < 1e-6 and (w > 0).all()
    assert min(idx) >= 999 and max(idx) < 999 + 300                          # tree indices of filled leaves
    res = tr.update({"states": s, "actions": a, "next_states": s2, "rewards": r, "dones": d, "idx": idx, "weights": w})
    assert np.isfinite(float(res["loss"])) and res["sum_epoch"] == 1
    after, _, beta2 = tr._learner.per_state(1000)
    touched = np.unique(np.asarray(idx) - 999)
    assert (after[touched] >= 0.01 ** 0.6 * (1 - 1e-5)).all() and (after[touched] <= 1 + 1e-6).all()
    assert (after[touched] != leaves[touched]).any() and abs(beta2 - 0.401) < 1e-12
    untouched = np.setdiff1d(np.arange(300), touched)
    assert np.array_equal(after[untouched], leaves[untouched])
    res = tr.learn_off_policy()
    assert res["sum_epoch"] == 2 and np.isfinite(float(res["loss"]))


def test_env_plugin_writes_csv_path_and_energy(tmp_path):
    """record_csv = 1 + the reference's <Power_param><Fly_power> block: run_eposide leaves logs/<name>_<time>.csv with the
    reference's header (Agents/UAV.py:269-277) and one row per print_loop episodes, path.csv with the finished episode's
    x,y,z rows (UAV.py:461-464), and energy_cost_total = accumulated Calc_Fly_Power; checkpoints appear every save_loop
    epochs although the loop runs on the device (DuelingDQN_Trainer.py:187-188)."""
    import csv
    import importlib
    from uavrl_b200.plugins import xmlconfig
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        cfg = xmlconfig.XML2Dict(os.path.join(ROOT, "configs", "PathPlan_City_B200.xml"))["simulator"]
        ed = cfg["env"]
        ed["num_UAV"], ed["scenario_pool"], ed["record_csv"], ed["print_loop"] = "128", "256", "1", "1"
        ed["Agent"]["xml_path_agent"] = os.path.join(ROOT, "configs", "UAV_energy_B200.xml")
        ed["Obstacles"]["buildings"] = os.path.join(ROOT, "configs", "buildings.xml")
        ed["Agent"]["Trainer"]["Trainer_path"] = os.path.join(ROOT, "configs", "Trainer_DDQN_B200.xml")
        mod = importlib.import_module("uavrl_b200.plugins." + ed["Env_Type"])
        os.chdir(tmp_path)                           # logs/, path.csv and Mod/ are CWD-relative like in the reference
        orig = mod.XML2Dict

        def patched(path):
            d = orig(path)
            if "Trainer" in d and isinstance(d["Trainer"], dict):
                d["Trainer"].update(Batch_Size="128", replay_size="16384", save_loop="64", model_path=str(tmp_path / "Mod"))
            return d
        mod.XML2Dict = patched
        try:
            env = getattr(mod, ed["Env_Type"])(ed)
        finally:
            mod.XML2Dict = orig
        info = env.run_eposide(0.3)
        info = env.run_eposide(0.1)
    finally:
        os.chdir(cwd)
    assert info["episodes"] >= 128
    logs = os.listdir(tmp_path / "logs")
    assert len(logs) == 1 and logs[0].startswith("UAV_batch_") and logs[0].endswith(".csv")
    rows = list(csv.reader(open(tmp_path / "logs" / logs[0])))
    assert rows[0][:9] == ["sum_Episode", "Episode", " Score", " Avg.Score", "eps-greedy", "success", "failed", "meet_threaten", "loss"]
    assert len(rows) == 3 and len(rows[1]) == 21 and float(rows[2][11]) > 0          # energy_cost column
    pts = np.array([[float(x) for x in r] for r in csv.reader(open(tmp_path / "path.csv"))])
    assert pts.ndim == 2 and pts.shape[1] == 3 and len(pts) >= 2
    assert (np.abs(np.diff(pts[:, :2], axis=0)).max(1) <= 1.0 + 1e-9).all()          # one step of at most Max_V per row
    assert env.Agents[0].energy_cost_total > 0
    saved = sorted(os.listdir(tmp_path / "Mod"))
    assert saved == ["q_local_DDQN_UAV_0.pth", "q_target_DDQN_UAV_0.pth"], saved


def test_is_train_0_is_greedy_in_the_device_loop(tmp_path):
    """Trainer.Is_Train = 0 (evaluation): get_action is greedy whatever eps is (DuelingDQN_Trainer.py:90) also inside the
    device-resident loop -- every stored action equals argmax_a q_local(state) -- and update() still counts epochs."""
    from uavrl_b200 import engine
    tr = make("DDQN_Trainer_B200", "QValueNet_SAC", tmp_path, Is_Train="0", lockstep_envs="64", replay_size="4096")
    g = np.load(os.path.join(ROOT, "tests", "golden", "env_golden.npz"))
    city = engine.City(g["dims"][0], g["dims"][1], g["dims"][2], g["buildings"])
    p = g["uav_params"]
    env = engine.EnvBatch(city, engine.UavParams(p[0], p[1], p[2], 1.0, int(p[3])), 64, max_subgoals=64, auto_reset=True)
    sc = env.make_scenarios(128, seed=3)
    env.set_pool(sc["start"], sc["goal"], sc["heading"], sc["sub"], sc["n_sub"])
    env.reset(0)
    st = engine.train_run(env, tr._learner, 20, 0.9, 1, False)
    assert st.env_steps == 64 * 20
    s, a, r, s2, d = tr._learner.gather(np.arange(64 * 20))
    q = tr.get_q(s)
    top2 = np.sort(q, 1)[:, -2:]
    clear = (top2[:, 1] - top2[:, 0]) > 1e-5
    assert clear.mean() > 0.95 and np.array_equal(a[clear], q.argmax(1)[clear])
    e0 = tr.epoch
    tr.update({"states": s[:8].tolist(), "actions": a[:8].tolist(), "next_states": s2[:8].tolist(), "rewards": r[:8].tolist(),
               "dones": d[:8].tolist()})
    assert tr.epoch == e0 + 1
    env.close()
