#!/usr/bin/env python
"""rrt_golden.npz: statistics of the REFERENCE's own UAV.reset -> RRTPlanner.getPath (Agents/UAV.py:335-366,
PathPlan/RRT.py:26-105), executed here.  The generator in csrc/rrt_core.cuh cannot reproduce the Python MT19937 stream,
so it is pinned statistically: sub-goal count, chain length and detour ratio distributions of 400 reference resets
(tests/test_abi_cpu.py::test_scenario_generator_matches_reference_statistics).
Run in the build container only (needs /root/reference):   python tests/golden/make_rrt_golden.py"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

sim_mod = ref_harness.load_reference()
import torch  # noqa: E402

if __name__ == "__main__":
    random.seed(42); np.random.seed(42); torch.manual_seed(42)
    s = sim_mod.simulator()
    env = s.env
    uav = env.Agents[0]
    n = 400
    n_sub = np.zeros(n, np.int32); chain = np.zeros(n); straight = np.zeros(n); seg_max = np.zeros(n)
    start = np.zeros((n, 3)); goal = np.zeros((n, 3))
    for i in range(n):
        env.Scene_Random_Reset()
        q = np.array([[g.x, g.y, g.z] for g in uav.sub_goals])
        n_sub[i] = len(q)
        seg = np.linalg.norm(np.diff(q, axis=0), axis=1)
        chain[i] = seg.sum(); seg_max[i] = seg.max()
        straight[i] = np.linalg.norm(q[-1] - q[0])
        start[i] = q[0]; goal[i] = q[-1]
    np.savez_compressed(os.path.join(HERE, "rrt_golden.npz"), n_sub=n_sub, chain=chain, straight=straight, seg_max=seg_max,
                        start=start, goal=goal)
    print("n_sub mean %.2f  median %d  min %d  max %d; chain/straight mean %.4f; seg_max max %.3f"
          % (n_sub.mean(), np.median(n_sub), n_sub.min(), n_sub.max(), (chain / straight).mean(), seg_max.max()))
