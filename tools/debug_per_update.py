import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import oracle as O
import uavrl_b200
from uavrl_b200 import engine
for B in (96, 128):
  for tc in (1, 0):
    for wmode in ("ones", "rand"):
        L = engine.Learner(100, [64, 64], 27, False, engine.ALGO_DDQN, batch_size=B, replay_capacity=4096, lockstep_envs=0, seed=5, update_loop=3)
        L.init_params(0); L.set_tensor_cores(bool(tc))
        net = O.make_net(100, [64, 64], 27, 0)
        ol = O.OracleLearner(net, O.ALGO_DDQN, L.get_params(0), gamma=0.99, lr=5e-4, update_loop=3)
        rng = np.random.default_rng(11)
        s = rng.standard_normal((B, 100)).astype(np.float32); s2 = rng.standard_normal((B, 100)).astype(np.float32)
        a = rng.integers(0, 27, B).astype(np.int32); r = rng.standard_normal(B).astype(np.float32)
        d = (rng.random(B) < 0.2).astype(np.float32)
        w = np.ones(B, np.float32) if wmode == "ones" else rng.uniform(0.1, 1.0, B).astype(np.float32)
        lo, _, aeo = ol.update(s, a, r, s2, d, is_w=w)
        ts, ta, tr, ts2, td, tw = [torch.tensor(x, device="cuda") for x in (s, a, r, s2, d, w)]
        ae = torch.zeros(B, device="cuda"); loss = torch.zeros(1, device="cuda")
        L.update_batch_per(ts, ta, tr, ts2, td, tw, ae, loss)
        torch.cuda.synchronize()
        aeg = ae.cpu().numpy()
        print("B", B, "tc", tc, wmode, "loss gpu %.6f oracle %.6f" % (float(loss), lo), "abs_err maxdiff %.2e" % np.abs(aeg - aeo).max(),
              "recomputed from gpu abs_err: %.6f" % float((w * aeg ** 2).mean()), "param diff %.2e" % np.abs(L.get_params(0) - ol.local).max())
        L.close()
