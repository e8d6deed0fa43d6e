#!/usr/bin/env python
"""Golden episodes of the reference's APF branch (Agents/UAV.py:156-210 cal_force / Adjust_subgoal, reward term :448-453).

The shipped config has APF_Enabled = 0 and its `building` obstacles carry no velocity (Obstacles/building.py:6-11), so the
branch never runs as shipped and would raise on `threaten.v`.  Here the UNMODIFIED reference UAV is driven with
APF_Enabled = 1 over the shipped 26-cylinder city after every obstacle has been given the `v` attribute cal_force reads
(half of them zero, which cal_force skips): this is the moving-obstacle scenario the code was written for.  Obstacles do not
move in the reference either (building.run() / PathPlan_City.run() are `pass`), `v` only feeds the force.

Run in the build container only (needs /root/reference):   python tests/golden/make_apf_golden.py
Writes tests/golden/apf_golden.npz: obstacle velocities, episodes (same layout as env_golden.npz) plus the remaining
sub-goal queue after every step (`subq`, zero padded) -- Adjust_subgoal shifts every entry every step."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (loads the reference through oracle/ref_harness; seeds everything with 42)
from BaseClass.CalMod import Loc  # noqa: E402  (reference)


def main():
    s = mg.sim_mod.simulator()
    env = s.env
    uav = env.Agents[0]
    rng = np.random.default_rng(20260924)
    nb = len(env.buildings)
    vel = np.zeros((nb, 3))
    moving = np.arange(nb) % 2 == 0
    ang = rng.uniform(0, 2 * np.pi, nb); spd = rng.uniform(0.2, 3.0, nb)
    vel[moving, 0] = (spd * np.cos(ang))[moving]; vel[moving, 1] = (spd * np.sin(ang))[moving]
    vel[4, 2] = 0.5                                         # one obstacle with a vertical component (only |v| sees it)
    for t, v in zip(env.buildings, vel):
        t.v = Loc(float(v[0]), float(v[1]), float(v[2]))
    uav.APF_Enabled = 1
    out = {"obstacle_v": vel, "buildings": np.array([[t.position.x, t.position.y, t.position.z, t._R, t._H] for t in env.buildings]),
           "dims": np.array([env.len, env.width, env.h], np.float64),
           "uav_params": np.array([uav.Max_V, float(uav.param.get("Min_V")), uav.Steering_angle, uav.Max_Step], np.float64)}
    eps = []
    for i in range(6):
        subq = []
        orig_move = env.Move_Agent

        def move(idx, action, _o=orig_move):
            res = _o(idx, action)
            q = np.zeros((mg.KMAX, 3))
            for k, sg in enumerate(uav.sub_goals):
                q[k] = (sg.x, sg.y, sg.z)
            subq.append(q)
            return res
        env.Move_Agent = move
        try:
            ep = mg.record_episode(env, uav, "seek" if i % 2 == 0 else "random", rng, 400)
        finally:
            env.Move_Agent = orig_move
        ep["subq"] = np.stack(subq)
        eps.append(ep)
        print("episode %d: %d steps, %d sub-goals, done=%d, reward sum %.3f" % (i, len(ep["action"]), int(ep["n_sub"]), int(ep["done"][-1]), ep["reward"].sum()))
    mg.pack_episodes(eps, "ep", out)
    np.savez_compressed(os.path.join(HERE, "apf_golden.npz"), **out)
    n = sum(len(e["action"]) for e in eps)
    # the force must actually have acted: queues differ from the scenario's
    moved = max(float(np.abs(e["subq"][0][:int(e["n_sub"])] - e["sub"][:int(e["n_sub"])]).max()) for e in eps)
    print("apf_golden.npz: %d steps, max first-step sub-goal shift %.4f m" % (n, moved))
    assert moved > 1e-3


if __name__ == "__main__":
    main()
