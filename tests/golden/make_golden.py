#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by EXECUTING THE PYTHON REFERENCE.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py

The reference ships no tests and no golden vectors, so parity is pinned on
outputs of the reference itself, produced here and committed as small .npz fixtures:

  env_golden.npz   city (26 cylinders from the reference's config/buildings.xml), UAV parameters,
                   primitive KATs (calculate_angle, Eu_Loc_distance, Threaten_rate incl. adversarial
                   on-the-boundary points) and full episodes driven through the reference's own
                   env.Move_Agent -> UAV.update_PathPlan / UAV.state_PathPlan
                   (Agents/UAV.py:397-567): per step action, reward, returned done, info,
                   collision, post-step state and the 100-d observation.
  env27_golden.npz the discrete-27 extension.  The reference has no such step function
                   (SURVEY.md section 0); class UAV27 below subclasses the reference's UAV and is
                   built only from the reference's own primitives.  This script ASSERTS that for the
                   three actions with (j,l)=(1,2) UAV27 is bit-identical to the reference's
                   update_PathPlan on every recorded step, then records episodes over all 27.
  sac_golden.npz   SAC continuous update (Trainer/SAC_Trainer.py:317-379) with injected reparameterisation noise:
                   batches, noise, parameter snapshots of actor / critics / targets, log_alpha, actor loss.
  dqn_golden.npz   learner math: the reference's DuelingDQN_Trainer.update, DDQN_Trainer /
                   DQN_Trainer.learn_off_policy (Trainer/*.py) run on fixed batches with torch CPU
                   fp32: initial params, batches, per-step loss, gradients, post-Adam params,
                   target params across hard updates.
"""
import copy
import math
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

sim_mod = ref_harness.load_reference()     # chdir into oracle/_work, seeds everything with 42
import torch  # noqa: E402
from BaseClass.CalMod import Loc, calculate_angle, Eu_Loc_distance  # noqa: E402  (reference)
import Agents.UAV as ref_uav_mod  # noqa: E402

INFO = {"normal": 0, "success": 1, "lose": 2}
KMAX = 64


# --------------------------------------------------------------------------- env helpers
def snapshot(uav):
    return dict(px=uav.position.x, py=uav.position.y, pz=uav.position.z,
                vx=uav.V_vector.x, vy=uav.V_vector.y, V=uav.V, step=uav.Step,
                nleft=len(uav.sub_goals), done=int(uav.done), score=uav.score,
                total_score=uav.total_score, path_len=uav.path_len)


def seek_action(uav, noise):
    """Steer toward the current sub-goal (policy used only to reach more branches of the step)."""
    if len(uav.sub_goals) == 0:
        return 0.0
    sg = uav.sub_goals[0]
    want = math.atan2(sg.y - uav.position.y, sg.x - uav.position.x)
    have = math.atan2(uav.V_vector.y, uav.V_vector.x)
    d = (want - have + math.pi) % (2 * math.pi) - math.pi
    a = d / uav.Steering_angle + noise
    return max(-1.0, min(1.0, a))


def record_episode(env, uav, policy, rng, max_steps, step_fn=None, action_sampler=None):
    """Reset through the reference (UAV.reset -> RRT) and drive it; returns a dict of arrays."""
    env.Scene_Random_Reset()
    nsub = len(uav.sub_goals)
    assert nsub <= KMAX
    alias0 = int(uav.sub_goals[0] is uav.position)
    sub = np.zeros((KMAX, 3))
    for i, sg in enumerate(uav.sub_goals):
        sub[i] = (sg.x, sg.y, sg.z)
    ep = dict(start=np.array([uav.position.x, uav.position.y, uav.position.z]),
              goal=np.array([uav.goal.x, uav.goal.y, uav.goal.z]),
              heading=np.float64(uav.V_dir), vx0=np.float64(uav.V_vector.x),
              vy0=np.float64(uav.V_vector.y), V0=np.float64(uav.V), sub=sub, n_sub=np.int32(nsub),
              alias0=np.uint8(alias0), obs0=uav.state().astype(np.float64))
    keys = ("action", "reward", "done_ret", "info", "collision", "px", "py", "pz", "vx", "vy", "V",
            "step", "cursor", "done", "score", "total_score", "path_len")
    rec = {k: [] for k in keys}
    obs = []
    for _ in range(max_steps):
        if uav.done:
            break
        if action_sampler is not None:
            act = action_sampler(uav, rng)
        elif policy == "random":
            act = rng.uniform(-1, 1)
        else:
            act = seek_action(uav, rng.normal(0, 0.25))
        before = (uav.position.x, uav.position.y, uav.position.z)
        if step_fn is None:
            next_state, reward, done, info = env.Move_Agent(0, [act, rng.uniform(-1, 1)])
        else:
            reward, done, info = step_fn(act)
            next_state = uav.state()
        s = snapshot(uav)
        rec["action"].append(act)
        rec["reward"].append(reward)
        rec["done_ret"].append(int(done))
        rec["info"].append(INFO[info])
        rec["collision"].append(int((uav.position.x, uav.position.y, uav.position.z) == before))
        for k in ("px", "py", "pz", "vx", "vy", "V", "step", "done", "score", "total_score", "path_len"):
            rec[k].append(s[k])
        rec["cursor"].append(nsub - s["nleft"])
        obs.append(np.asarray(next_state, np.float64))
    for k in keys:
        dt = np.float64
        if k in ("done_ret", "info", "collision", "done"):
            dt = np.uint8
        if k in ("step", "cursor"):
            dt = np.int32
        ep[k] = np.asarray(rec[k], dt)
    ep["obs"] = np.stack(obs)
    return ep


def pack_episodes(eps, prefix, out):
    out[prefix + "n_episodes"] = np.int32(len(eps))
    for i, ep in enumerate(eps):
        for k, v in ep.items():
            out["%s%d_%s" % (prefix, i, k)] = v


# --------------------------------------------------------------------------- discrete-27 extension
class UAV27(ref_uav_mod.UAV):
    """update_PathPlan27: the reference step with a discrete 27-action decode in front.

    k -> (i,j,l) = (k//9, (k//3)%3, k%3): steering a0 = i-1 (fed to UAV.py:414), climb
    dz = (j-1)*climb_rate applied with the x/y move (:419-420), speed level l in
    {Min_V,(Min_V+Max_V)/2,Max_V} replacing Max_V at :415-416 only.  Everything else is the
    reference's arithmetic, via the reference's own primitives.
    """
    climb_rate = 1.0

    def update_PathPlan27(self, k):
        import csv  # noqa: F401  (the reference writes path.csv at terminal steps; skipped here)
        i, j, l = k // 9, (k // 3) % 3, k % 3
        a0 = float(i - 1)
        dz = float(j - 1) * self.climb_rate
        min_v = float(self.param.get("Min_V"))
        speed = (min_v, (min_v + self.Max_V) / 2, self.Max_V)[l]
        global_r = 0
        if len(self.sub_goals) == 0:
            self.done = True
            global_r += (self.Max_Step - self.Step)
            self.score += global_r
            return global_r, True, 'success'
        self.Step += 1
        old_position = copy.copy(self.position)
        seta_old = calculate_angle(Loc(0, 0, 0), self.V_vector)
        dis_old = Eu_Loc_distance(self.position, self.sub_goals[0])
        dis2goal_old = Eu_Loc_distance(self.position, self.goal)
        seta_new = seta_old + a0 * self.Steering_angle
        self.V_vector.x = speed * math.cos(seta_new)
        self.V_vector.y = speed * math.sin(seta_new)
        self.V = self.Calc_V()
        self.position.x += self.V_vector.x
        self.position.y += self.V_vector.y
        self.position.z += dz
        tri_goal = calculate_angle(self.position, self.sub_goals[0])
        tri_V = calculate_angle(Loc(0, 0, 0), self.V_vector)
        if self.env.Threaten_rate(self.position) == 1:
            global_r -= 0.3
            self.position = old_position
            tri_V = calculate_angle(self.position, self.sub_goals[0])
        dis_new = Eu_Loc_distance(self.position, self.sub_goals[0])
        dis2goal_new = Eu_Loc_distance(self.position, self.goal)
        global_r -= 0.13 * abs(a0)
        global_r += 0.2 * math.cos(abs(tri_goal - tri_V))
        global_r += 0.4 * (dis_old - dis_new)
        global_r += 0.4 * (dis2goal_old - dis2goal_new)
        global_r -= 0.1
        if len(self.sub_goals) >= 1:
            global_r -= 0.01 * abs(self.position.z - self.sub_goals[0].z)
        self.path_len += self.V
        if self.Step >= self.Max_Step:
            self.done = True
            global_r += (50 - Eu_Loc_distance(self.position, self.sub_goals[0]))
            self.score += global_r
            self.total_score += global_r
            return global_r, True, 'lose'
        elif Eu_Loc_distance(self.position, self.sub_goals[0]) < 7 or \
                (Eu_Loc_distance(self.position, self.goal) < Eu_Loc_distance(self.sub_goals[0], self.goal)):
            global_r += (50 - Eu_Loc_distance(self.position, self.sub_goals[0]))
            self.sub_goals.pop(0)
            if len(self.sub_goals) == 0:
                global_r += 50
                self.done = True
                global_r += (self.Max_Step - self.Step)
                self.score += global_r
                self.total_score += global_r
                return global_r, True, 'success'
            else:
                self.reset("local reset")
                tri_goal = calculate_angle(self.position, self.sub_goals[0])
                tri_V = calculate_angle(Loc(0, 0, 0), self.V_vector)
                global_r += 0.2 * math.cos(abs(tri_goal - tri_V))
                global_r += (self.Max_Step - self.Step)
                self.score += global_r
                self.total_score += global_r
                return global_r, True, 'success'
        elif Eu_Loc_distance(self.position, self.goal) < 7:
            self.done = True
            global_r += 50
            global_r += (self.Max_Step - self.Step)
            self.score += global_r
            self.total_score += global_r
            return global_r, True, 'success'
        else:
            self.score += global_r
            self.total_score += global_r
            return global_r, False, 'normal'


def clone_uav_state(src, dst):
    dst.position = copy.copy(src.position)
    dst.V_vector = copy.copy(src.V_vector)
    dst.V = src.V
    dst.Step = src.Step
    dst.done = src.done
    dst.score = src.score
    dst.total_score = src.total_score
    dst.path_len = src.path_len
    dst.goal = copy.copy(src.goal)
    dst.sub_goals = [dst.position if sg is src.position else copy.copy(sg) for sg in src.sub_goals]


# --------------------------------------------------------------------------- main: env part
def gen_env(out_env, out_env27):
    s = sim_mod.simulator()
    env = s.env
    uav = env.Agents[0]
    rng = np.random.default_rng(20260923)

    b = np.array([[t.position.x, t.position.y, t.position.z, t._R, t._H] for t in env.buildings])
    out_env["buildings"] = b
    out_env["dims"] = np.array([env.len, env.width, env.h], np.float64)
    out_env["uav_params"] = np.array([uav.Max_V, float(uav.param.get("Min_V")), uav.Steering_angle,
                                      uav.Max_Step], np.float64)
    # ---- KATs: angle / distance
    P = rng.uniform(-600, 600, size=(400, 6))
    P[:20, 2:] = 0.0                      # some z = 0 planes
    P[20:30, :] = np.round(P[20:30, :])   # integer-valued
    P[30] = 0.0                           # coincident points: atan2(0,0)
    P[31, 3:] = P[31, :3]
    P[32] = (0, 0, 0, 1, 0, 0); P[33] = (0, 0, 0, -1, 0, 0); P[34] = (0, 0, 0, 0, 1, 0)
    P[35] = (0, 0, 0, 0, -1, 0); P[36] = (0, 0, 0, -1, -1e-300, 0); P[37] = (0, 0, 0, 1, -1e-18, 0)
    out_env["kat_pts"] = P
    out_env["kat_angle"] = np.array([calculate_angle(Loc(*p[:3]), Loc(*p[3:])) for p in P])
    out_env["kat_dist"] = np.array([Eu_Loc_distance(Loc(*p[:3]), Loc(*p[3:])) for p in P])
    # ---- KATs: Threaten_rate, random + adversarial
    pts = [rng.uniform([-20, -20, -5], [520, 520, 110], size=(3000, 3))]
    edge = []
    for (cx, cy, cz, R, H) in b:
        for th in rng.uniform(0, 2 * math.pi, 6):
            for scale in (1.0, 1.0 - 1e-15, 1.0 + 1e-15, 1 - 1e-9, 1 + 1e-9):
                edge.append((cx + scale * R * math.cos(th), cy + scale * R * math.sin(th), 0.0))
        edge.append((cx + R, cy, 0.0)); edge.append((cx - R, cy, 0.0)); edge.append((cx, cy + R, 0.0))
        edge.append((cx, cy, H)); edge.append((cx, cy, np.nextafter(H, 1e9))); edge.append((cx, cy, np.nextafter(H, -1e9)))
    for v in (0.0, -0.0, 500.0, np.nextafter(500.0, 1e9), np.nextafter(0.0, -1.0), 1e-300):
        edge.append((v, 250.0, 10.0)); edge.append((250.0, v, 10.0))
    for v in (0.0, 100.0, np.nextafter(100.0, 1e9), -1e-300):
        edge.append((3.0, 3.0, v))
    pts.append(np.array(edge))
    pts = np.concatenate(pts)
    out_env["kat_threat_pts"] = pts
    out_env["kat_threat"] = np.array([env.Threaten_rate(Loc(*p)) for p in pts], np.uint8)

    # ---- episodes through the reference's own Move_Agent
    eps = []
    for i in range(10):
        policy = "random" if i < 3 else "seek"
        eps.append(record_episode(env, uav, policy, rng, max_steps=3000))
        print("episode", i, policy, "steps", len(eps[-1]["action"]), "collisions",
              int(eps[-1]["collision"].sum()), "final info", eps[-1]["info"][-1], "n_sub", eps[-1]["n_sub"])
    pack_episodes(eps, "ep", out_env)

    # ---- discrete-27: build a UAV27 sharing the env; prove the (j,l)=(1,2) identity, then record
    uav_params = copy.copy(uav.param)
    u27 = UAV27(uav_params, env)
    checked = 0
    for trial in range(6):
        env.Scene_Random_Reset()
        clone_uav_state(uav, u27)
        for t in range(400):
            if uav.done:
                break
            if trial < 3:
                i = int(rng.integers(0, 3))
            else:
                a = seek_action(uav, rng.normal(0, 0.3))
                i = int(np.clip(round(a), -1, 1)) + 1
            k = i * 9 + 1 * 3 + 2
            r_ref = uav.update_PathPlan([float(i - 1), 0.0])
            r_27 = u27.update_PathPlan27(k)
            assert r_ref == r_27, (r_ref, r_27)
            assert snapshot(uav) == snapshot(u27)
            assert np.array_equal(uav.state(), u27.state())
            checked += 1
    print("UAV27 == reference update_PathPlan on", checked, "steps (bit-identical)")
    out_env27["identity_steps_checked"] = np.int32(checked)
    out_env27["climb_rate"] = np.float64(UAV27.climb_rate)

    env.Agents[0] = u27          # so that Scene_Random_Reset / state drive the UAV27 instance
    eps27 = []

    def sampler_random(u, g):
        return int(g.integers(0, 27))

    def sampler_seek(u, g):
        a = seek_action(u, g.normal(0, 0.3))
        i = int(np.clip(round(a), -1, 1)) + 1
        # climb toward the sub-goal height, random speed level
        if len(u.sub_goals):
            dzw = u.sub_goals[0].z - u.position.z
            j = 2 if dzw > 0.5 else (0 if dzw < -0.5 else 1)
        else:
            j = 1
        if g.uniform() < 0.2:
            j = int(g.integers(0, 3))
        l = int(g.integers(0, 3))
        return i * 9 + j * 3 + l

    for i in range(8):
        smp = sampler_random if i < 3 else sampler_seek
        eps27.append(record_episode(env, u27, None, rng, max_steps=3000,
                                    step_fn=u27.update_PathPlan27, action_sampler=smp))
        print("episode27", i, "steps", len(eps27[-1]["action"]), "collisions",
              int(eps27[-1]["collision"].sum()), "final info", eps27[-1]["info"][-1],
              "z range", eps27[-1]["pz"].min(), eps27[-1]["pz"].max())
    pack_episodes(eps27, "ep", out_env27)
    env.Agents[0] = uav


# --------------------------------------------------------------------------- main: learner part
def flat_params(net):
    return np.concatenate([p.detach().numpy().ravel() for p in net.state_dict().values()]).astype(np.float32)


def flat_grads(net):
    return np.concatenate([p.grad.detach().numpy().ravel() for p in net.parameters()]).astype(np.float32)


def synth_batch(rng, B, nA):
    """Observation-shaped synthetic batch: 20 real-valued features + 80 binary probes."""
    def obs():
        o = np.zeros((B, 100), np.float32)
        o[:, :11] = rng.normal(0, 2.0, size=(B, 11))
        o[:, 11:86] = (rng.uniform(size=(B, 75)) < 0.2)
        o[:, 86:90] = rng.normal(0, 10.0, size=(B, 4))
        o[:, 90:95] = (rng.uniform(size=(B, 5)) < 0.5)
        return o
    s, s2 = obs(), obs()
    a = rng.integers(0, nA, size=B).astype(np.int32)
    r = rng.normal(0, 1.0, size=B).astype(np.float32)
    r[rng.uniform(size=B) < 0.1] += 150.0          # sub-goal bonuses
    d = (rng.uniform(size=B) < 0.15).astype(np.float32)
    return s, a, r, s2, d


def gen_dqn(out):
    from FactoryClass.TrainerFactory import TrainerFactory
    rng = np.random.default_rng(7)
    B, nA, K = 64, 27, 10
    SNAP = [0, 2, 3, 9]          # 1st step, 3rd (hard update happens), 4th, 10th
    cases = [("dueling_vanet2", "DuelingDQN_Trainer", "VAnet2", 64),
             ("dueling_vanet3", "DuelingDQN_Trainer", "VAnet3", 64),
             ("ddqn_qvalue3", "DDQN_Trainer", "QValueNet_SAC", 64),
             ("dqn_qvalue3", "DQN_Trainer", "QValueNet_SAC", 64),
             ("dqn_qnet2", "DQN_Trainer", "Qnet2", 64)]
    out["torch_version"] = np.array(torch.__version__)
    # the same K batches feed every case (keeps the fixture small)
    batches = {k: [] for k in "s a r s2 d".split()}
    for step in range(K):
        for k, v in zip("s a r s2 d".split(), synth_batch(rng, B, nA)):
            batches[k].append(v)
    for k in batches:
        out["batch_" + k] = np.stack(batches[k])
    out["cases"] = np.array([c[0] for c in cases])
    for name, trainer_type, net, h in cases:
        torch.manual_seed(1234)
        param = {"Trainer_Type": trainer_type, "NetWork": net, "w": "100", "hiden_dim": str(h),
                 "output": str(nA), "name": "golden_" + name, "LEARNING_RATE": "0.0005",
                 "Batch_Size": str(B), "gamma": "0.99", "save_loop": "1000000000",
                 "replay_size": "10000", "Update_loop": "3", "Is_Train": "1"}
        tr = TrainerFactory().Create_Trainer(param)
        assert tr is not None
        # the reference constructs q_target with independent random weights; record both
        out[name + "_local0"] = flat_params(tr.q_local)
        out[name + "_target0"] = flat_params(tr.q_target)
        losses, locals_, targets, grads = [], [], [], []
        for step in range(K):
            s, a, r, s2, d = (batches[k][step] for k in "s a r s2 d".split())
            if trainer_type == "DuelingDQN_Trainer":
                # lists, not arrays: `transition_dict['states']==[]` (:155) is an elementwise compare on ndarrays
                td = {"states": s.tolist(), "actions": a.tolist(), "next_states": s2.tolist(), "rewards": r.tolist(),
                      "dones": d.tolist()}
                tr.update(td)                                  # DuelingDQN_Trainer.py:150-190
            else:
                tr.replay_memory.memory = [
                    (torch.tensor(s[i:i + 1]), torch.tensor([[int(a[i])]]), torch.tensor([[r[i]]]),
                     torch.tensor(s2[i:i + 1]), torch.tensor([[d[i]]])) for i in range(B)]
                tr.learn_off_policy()                          # DQN_Trainer.py:85-136 / DDQN :72-117
            losses.append(float(tr.loss))
            grads.append(flat_grads(tr.q_local))
            locals_.append(flat_params(tr.q_local))
            targets.append(flat_params(tr.q_target))
        out[name + "_loss"] = np.asarray(losses, np.float32)
        # keep the fixture small: parameter snapshots only at these update indices
        out[name + "_snap"] = np.asarray(SNAP, np.int32)
        out[name + "_grads"] = np.stack([grads[i] for i in SNAP])
        out[name + "_local"] = np.stack([locals_[i] for i in SNAP])
        out[name + "_target"] = np.stack([targets[i] for i in SNAP])
        # forward KAT + greedy actions on the final network
        with torch.no_grad():
            q = tr.q_local(torch.tensor(batches["s"][0])).numpy()
        out[name + "_q_final"] = q.astype(np.float32)
        print(name, "params", out[name + "_local0"].size, "losses", losses[:3], "...")
    # epsilon schedule KAT (simulator.py:141-145)
    s = sim_mod.simulator.__new__(sim_mod.simulator)
    vals = []
    for (mx, mn, ep) in [(1, 0.1, 0), (1, 0.1, 1), (1000, 0.01, 0), (1000, 0.01, 500), (1000, 0.01, 5000), (200, 0.05, 100)]:
        s.min_eps, s.max_eps_episode, s.epoch = mn, mx, ep
        vals.append((mx, mn, ep, s.epsilon_annealing()))
    out["eps_schedule"] = np.asarray(vals, np.float64)


def gen_sac(out):
    """SAC continuous (Trainer/SAC_Trainer.py:317-379, nets BaseClass/BaseCNN.py:459-500) with the
    reparameterisation noise injected: Normal.rsample is patched to consume recorded eps tensors."""
    from FactoryClass.TrainerFactory import TrainerFactory
    rng = np.random.default_rng(11)
    B, K = 64, 6
    SNAP = [0, 2, 5]
    eps_queue = []

    def rsample(self, sample_shape=torch.Size()):
        e = eps_queue.pop(0)
        assert e.shape == self.loc.shape
        return self.loc + self.scale * e
    orig = torch.distributions.Normal.rsample
    torch.distributions.Normal.rsample = rsample
    try:
        torch.manual_seed(4321)
        param = {"Trainer_Type": "SAC_Trainer", "Is_Train": "1", "IsPriority_Replay": "0", "name": "golden_sac",
                 "actor": {"NetWork": "PolicyNetContinuous_SAC", "h": "1", "w": "100", "channel": "1", "action_bound": "1",
                           "hiden_dim": "64", "output": "2", "lr": "0.0001"},
                 "critic": {"NetWork": "QValueNetContinuous_SAC", "h": "1", "w": "100", "channel": "1", "hiden_dim": "64",
                            "action_dim": "2", "lr": "0.001"},
                 "SAC_param": {"IS_Continuous": "1", "alpha_lr": "0.0001", "target_entropy": "1", "gamma": "0.99", "tau": "0.05"},
                 "replay_size": "10000", "LEARNING_RATE": "0.0005", "Batch_Size": str(B), "max_epoch": "100",
                 "save_loop": "1000000000"}
        tr = TrainerFactory().Create_Trainer(param)
        assert tr is not None
        tr.replay_memory.memory = [0] * (B + 1)            # only its length gates training (:327)
        for nm in ("actor", "critic_1", "critic_2", "target_critic_1", "target_critic_2"):
            out["sac_%s0" % nm] = flat_params(getattr(tr, nm))
        out["sac_log_alpha0"] = np.float32(tr.log_alpha.item())
        out["sac_hparams"] = np.array([1e-4, 1e-3, 1e-4, 1.0, 0.99, 0.05], np.float64)  # actor_lr critic_lr alpha_lr target_entropy gamma tau
        keys = ("s", "a", "r", "s2", "d", "eps_next", "eps_cur")
        rec = {k: [] for k in keys}
        snaps = {nm: [] for nm in ("actor", "critic_1", "critic_2", "target_critic_1", "target_critic_2")}
        la, losses = [], []
        for step in range(K):
            s, _, r, s2, d = synth_batch(rng, B, 27)
            a = rng.uniform(-1, 1, size=(B, 2)).astype(np.float32)
            e1 = rng.normal(size=(B, 2)).astype(np.float32); e2 = rng.normal(size=(B, 2)).astype(np.float32)
            eps_queue[:] = [torch.tensor(e1), torch.tensor(e2)]
            for k, v in zip(keys, (s, a, r, s2, d, e1, e2)):
                rec[k].append(v)
            td = {"states": s.tolist(), "actions": a.tolist(), "next_states": s2.tolist(), "rewards": r.tolist(), "dones": d.tolist()}
            res = tr.update(td)
            assert not eps_queue
            losses.append(float(res["loss"]))
            la.append(tr.log_alpha.item())
            if step in SNAP:
                for nm in snaps:
                    snaps[nm].append(flat_params(getattr(tr, nm)))
        for k in keys:
            out["sac_" + k] = np.stack(rec[k])
        out["sac_snap"] = np.asarray(SNAP, np.int32)
        for nm in snaps:
            out["sac_" + nm] = np.stack(snaps[nm])
        out["sac_log_alpha"] = np.asarray(la, np.float32)
        out["sac_actor_loss"] = np.asarray(losses, np.float32)
        # get_action KAT (SAC_Trainer.py:444-448): action list of one state, noise injected
        acts, e_list = [], []
        for i in range(8):
            e = rng.normal(size=(1, 2)).astype(np.float32)
            eps_queue[:] = [torch.tensor(e)]
            acts.append(tr.get_action(rec["s"][0][i], 0.0)); e_list.append(e[0])
        out["sac_act_eps"] = np.stack(e_list); out["sac_act_out"] = np.asarray(acts, np.float32)
        print("sac: actor", out["sac_actor0"].size, "critic", out["sac_critic_10"].size, "actor losses", losses[:3], "log_alpha", la[:2])
    finally:
        torch.distributions.Normal.rsample = orig


if __name__ == "__main__":
    random.seed(42); np.random.seed(42); torch.manual_seed(42)
    env_out, env27_out, dqn_out = {}, {}, {}
    gen_env(env_out, env27_out)
    np.savez_compressed(os.path.join(HERE, "env_golden.npz"), **env_out)
    np.savez_compressed(os.path.join(HERE, "env27_golden.npz"), **env27_out)
    gen_dqn(dqn_out)
    np.savez_compressed(os.path.join(HERE, "dqn_golden.npz"), **dqn_out)
    sac_out = {}
    gen_sac(sac_out)
    np.savez_compressed(os.path.join(HERE, "sac_golden.npz"), **sac_out)
    for f in ("env_golden.npz", "env27_golden.npz", "dqn_golden.npz", "sac_golden.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")
