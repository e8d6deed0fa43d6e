"""Generate oracle/pool_512.npz: 512 reset scenarios (start, goal, heading, RRT sub-goal queue) of the 26-cylinder city.
The CPU arm of bench.py (`--impl reference`, `cpu_baseline`) loads this committed pool so that it never maps the product
library.  The scenarios come from the repository's host-side generator (a statistical restatement of UAV.reset + RRT.getPath,
pinned against 400 reference resets by tests/test_abi_cpu.py); run once, here:   python oracle/make_pool.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import ctypes as C
    import uavrl_b200  # noqa: F401
    from uavrl_b200 import _lib
    g = np.load(os.path.join(ROOT, "tests", "golden", "env_golden.npz"))
    dims, b, p = g["dims"], np.ascontiguousarray(g["buildings"]), g["uav_params"]
    cfg = _lib.EnvConfig()
    cfg.n_envs, cfg.max_subgoals = 1, 64
    cfg.len, cfg.width, cfg.h = dims
    cfg.max_v, cfg.min_v, cfg.steering_angle, cfg.max_step, cfg.climb_rate = p[0], p[1], p[2], int(p[3]), 1.0
    cfg.n_buildings, cfg.buildings_host = b.shape[0], b.ctypes.data_as(C.POINTER(C.c_double))
    P = 512
    sc = dict(start=np.zeros((P, 3)), goal=np.zeros((P, 3)), heading=np.zeros(P), sub=np.zeros((P, 64, 3)), n_sub=np.zeros(P, np.int32))
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    rc = _lib.lib().uavrl_make_scenarios(C.byref(cfg), 42, P, 30, vp(sc["start"]), vp(sc["goal"]), vp(sc["heading"]), vp(sc["sub"]), vp(sc["n_sub"]))
    assert rc == 0
    kmax = int(sc["n_sub"].max())
    np.savez_compressed(os.path.join(ROOT, "oracle", "pool_512.npz"), start=sc["start"], goal=sc["goal"], heading=sc["heading"],
                        sub=sc["sub"][:, :kmax].astype(np.float64), n_sub=sc["n_sub"], dims=dims, buildings=b, uav_params=p)
    print("pool_512.npz: %d scenarios, sub-goal queues of %d..%d nodes" % (P, sc["n_sub"].min(), kmax))


if __name__ == "__main__":
    main()
