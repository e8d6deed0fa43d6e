"""SAC lockstep loop (BASELINE configs[4] shape) on its own: `python tools/sac_probe.py [envs] [iters]`; run under
`ncu --metrics gpu__time_duration.sum` for the per-kernel launch list of one iteration."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
import uavrl_b200  # noqa: F401
from uavrl_b200 import engine
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 20
replay = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
dims, b, p = bench.load_city()
city = engine.City(dims[0], dims[1], dims[2], b)
params = engine.UavParams(p[0], p[1], p[2], 1.0, int(p[3]))
env = engine.EnvBatch(city, params, N, max_subgoals=64, device=0, auto_reset=True)
sc = env.make_scenarios(2048, seed=42)
env.set_pool(sc["start"], sc["goal"], sc["heading"], sc["sub"], sc["n_sub"])
env.reset(0)
L = engine.SacLearner(100, 64, 2, 1.0, 1e-4, 1e-3, 1e-4, 1.0, 0.99, 0.05, batch_size=N, replay_capacity=replay, lockstep_envs=N, seed=7, device=0)
L.init_params(0)
engine.sac_train_run(env, L, (replay + N - 1) // N + 1, False, want_stats=False)
engine.sac_train_run(env, L, 3, True, want_stats=False)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); engine.sac_train_run(env, L, K, True, want_stats=False); e1.record(); torch.cuda.synchronize()
print("sac_probe: %d envs, %d iters: %.1f us/iter, %.2f M env steps/s" % (N, K, e0.elapsed_time(e1) / K * 1e3, N * K / e0.elapsed_time(e1) / 1e3))
