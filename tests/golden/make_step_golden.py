#!/usr/bin/env python
"""step_golden.npz: single steps of the REFERENCE (Agents/UAV.py:397-567) from constructed states that sit next to every
decision boundary of the step -- cylinder surfaces and roofs, the map edges, the 7 m sub-goal / goal radii, the
`closer to the goal than the sub-goal` test, Step = Max_Step - 1 / Max_Step, one or two sub-goals left, the first step after
a reset (sub_goals[0] IS the position object) -- with offsets from 1e-9 m to metres on either side.  Continuous action
(update_PathPlan) and the discrete-27 extension (UAV27 of make_golden.py, built from the reference's primitives).
Per sample: the pre-state, the action, and the reference's reward, returned done, info, collision, post-state and 100-d
observation.  Run in the build container only:  python tests/golden/make_step_golden.py"""
import copy
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (loads the reference, defines UAV27 / snapshot)
from BaseClass.CalMod import Loc  # noqa: E402  (reference)

KS = 4                      # sub-goal slots stored per sample
OFFS = np.array([1e-9, 1e-6, 1e-3, 0.05, 0.5, 1.5, 4.0])


def build_state(rng, env, b, max_step):
    """Return dict(pos, theta, speed, step, score, total, path_len, sub[list of xyz], alias0)."""
    kind = rng.integers(0, 8)
    theta = rng.uniform(0, 2 * math.pi)
    speed = float(rng.choice([1.0, 1.0, 0.8, 0.6]))
    pos = np.array([rng.uniform(5, 495), rng.uniform(5, 495), float(rng.choice([0.0, 0.0, 3.0, 12.0, 40.0]))])
    n_sub = int(rng.integers(1, KS + 1))
    sub = [np.array([rng.uniform(20, 480), rng.uniform(20, 480), float(rng.choice([0.0, 0.0, 5.0]))]) for _ in range(n_sub)]
    step = int(rng.integers(0, max_step - 2))
    off = float(rng.choice(OFFS)) * float(rng.choice([-1.0, 1.0]))
    if kind == 0:                                   # next to a cylinder wall, heading across it
        cx, cy, cz, R, H = b[rng.integers(0, len(b))]
        phi = rng.uniform(0, 2 * math.pi)
        r = R + 1.0 * rng.uniform(0.0, 1.2) + off       # about one step outside the wall
        pos = np.array([cx + r * math.cos(phi), cy + r * math.sin(phi), float(rng.choice([0.0, 0.0, H + off, H - 1.0]))])
        theta = (phi + math.pi + rng.normal(0, 0.4)) % (2 * math.pi)
    elif kind == 1:                                 # next to the map edge
        edge = rng.integers(0, 4)
        v = 1.0 * rng.uniform(0, 1.1) + off
        if edge == 0: pos[0] = v; theta = math.pi + rng.normal(0, 0.3)
        elif edge == 1: pos[0] = 500.0 - v; theta = rng.normal(0, 0.3)
        elif edge == 2: pos[1] = v; theta = 1.5 * math.pi + rng.normal(0, 0.3)
        else: pos[1] = 500.0 - v; theta = 0.5 * math.pi + rng.normal(0, 0.3)
        theta %= 2 * math.pi
    elif kind == 2:                                 # about one step outside the 7 m radius of the sub-goal
        phi = rng.uniform(0, 2 * math.pi)
        r = 7.0 + speed * rng.uniform(0.0, 1.1) + off
        sub[0] = np.array([pos[0] + r * math.cos(phi), pos[1] + r * math.sin(phi), pos[2] + float(rng.choice([0.0, 0.0, 0.3]))])
        theta = (phi + rng.normal(0, 0.2)) % (2 * math.pi)
    elif kind == 3:                                 # same for the final goal
        phi = rng.uniform(0, 2 * math.pi)
        r = 7.0 + speed * rng.uniform(0.0, 1.1) + off
        sub[-1] = np.array([pos[0] + r * math.cos(phi), pos[1] + r * math.sin(phi), pos[2]])
        theta = (phi + rng.normal(0, 0.2)) % (2 * math.pi)
    elif kind == 4:                                 # Step at the limit
        step = int(rng.choice([max_step - 2, max_step - 1, max_step]))
    elif kind == 5 and n_sub >= 2:                  # |p - goal| vs |sg0 - goal|: the sub-goal is passed
        g = sub[-1]
        d0 = np.linalg.norm(sub[0] - g)
        phi = rng.uniform(0, 2 * math.pi)
        r = d0 + speed * rng.uniform(-0.2, 1.1) + off
        pos = np.array([g[0] + r * math.cos(phi), g[1] + r * math.sin(phi), g[2]])
        theta = (phi + math.pi + rng.normal(0, 0.3)) % (2 * math.pi)
    alias0 = int(kind == 6)                         # the first step after a reset
    if alias0:
        step = 0
    return dict(pos=pos, theta=theta, speed=speed, step=step, score=float(rng.normal(0, 30)), total=float(rng.normal(0, 200)),
                path_len=float(rng.uniform(0, 400)), sub=sub, alias0=alias0)


def install(uav, st):
    uav.position = Loc(float(st["pos"][0]), float(st["pos"][1]), float(st["pos"][2]))
    uav.V_vector = Loc(st["speed"] * math.cos(st["theta"]), st["speed"] * math.sin(st["theta"]), 0.0)
    uav.V = uav.Calc_V()
    uav.Step = st["step"]
    uav.done = False
    uav.score, uav.total_score, uav.path_len = st["score"], st["total"], st["path_len"]
    subs = [Loc(float(s[0]), float(s[1]), float(s[2])) for s in st["sub"]]
    if st["alias0"]:
        subs = [uav.position] + subs[1:] if len(subs) > 1 else [uav.position, subs[0]]
    uav.sub_goals = subs
    uav.goal = copy.copy(subs[-1])
    uav.path = []


def run(rng, env, uav, n, discrete, b, max_step):
    keys = ("px", "py", "pz", "vx", "vy", "V", "step", "score", "total_score", "path_len", "n_sub", "alias0", "action", "reward",
            "done_ret", "info", "collision", "o_px", "o_py", "o_pz", "o_vx", "o_vy", "o_V", "o_step", "o_cursor", "o_done", "o_score",
            "o_total_score", "o_path_len")
    rec = {k: [] for k in keys}
    subs, goals, obs = [], [], []
    for _ in range(n):
        st = build_state(rng, env, b, max_step)
        install(uav, st)
        n_sub = len(uav.sub_goals)
        sub = np.zeros((KS + 1, 3))
        for i, sg in enumerate(uav.sub_goals):
            sub[i] = (sg.x, sg.y, sg.z)
        pre = mg.snapshot(uav)
        before = (uav.position.x, uav.position.y, uav.position.z)
        if discrete:
            act = int(rng.integers(0, 27))
            reward, done, info = uav.update_PathPlan27(act)
        else:
            act = float(rng.uniform(-1, 1))
            reward, done, info = uav.update_PathPlan([act, float(rng.uniform(-1, 1))])
        post = mg.snapshot(uav)
        for k in ("px", "py", "pz", "vx", "vy", "V", "step", "score", "total_score", "path_len"):
            rec[k].append(pre[k]); rec["o_" + k].append(post[k])
        rec["n_sub"].append(n_sub); rec["alias0"].append(st["alias0"]); rec["action"].append(act)
        rec["reward"].append(reward); rec["done_ret"].append(int(done)); rec["info"].append(mg.INFO[info])
        rec["collision"].append(int((uav.position.x, uav.position.y, uav.position.z) == before))
        rec["o_cursor"].append(n_sub - post["nleft"]); rec["o_done"].append(post["done"])
        subs.append(sub); goals.append((uav.goal.x, uav.goal.y, uav.goal.z))
        obs.append(np.asarray(uav.state(), np.float64))
    out = {k: np.asarray(v) for k, v in rec.items()}
    out["sub"] = np.asarray(subs); out["goal"] = np.asarray(goals); out["obs"] = np.asarray(obs)
    return out


if __name__ == "__main__":
    import random
    import torch
    random.seed(42); np.random.seed(42); torch.manual_seed(42)
    s = mg.sim_mod.simulator()
    env = s.env
    uav = env.Agents[0]
    b = np.array([[t.position.x, t.position.y, t.position.z, t._R, t._H] for t in env.buildings])
    rng = np.random.default_rng(20260924)
    res = {}
    for k, v in run(rng, env, uav, 3000, False, b, uav.Max_Step).items():
        res["c_" + k] = v
    u27 = mg.UAV27(copy.copy(uav.param), env)
    env.Agents[0] = u27
    for k, v in run(rng, env, u27, 3000, True, b, u27.Max_Step).items():
        res["d_" + k] = v
    np.savez_compressed(os.path.join(HERE, "step_golden.npz"), **res)
    for p in ("c_", "d_"):
        print(p, "n", len(res[p + "reward"]), "collisions", int(res[p + "collision"].sum()), "done_ret", int(res[p + "done_ret"].sum()),
              "info", np.bincount(res[p + "info"], minlength=3), "ended", int(res[p + "o_done"].sum()))
    print(os.path.getsize(os.path.join(HERE, "step_golden.npz")) // 1024, "KiB")
