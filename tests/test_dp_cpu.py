"""world_size-2 checks on CPU (gloo): the data-parallel convention of the learner
(uavrl_learner_compute_grads scales the local loss by 1/global_batch; ranks all-reduce-SUM the gradient
vector; every rank applies the same Adam step) reproduces the single-process full-batch update, and
bench.py's reference arm behaves under a multi-rank launch (rank 0 prints, the others exit 0)."""
import json
import os
import subprocess
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle as O
from conftest import GOLDEN, ROOT


def _worker(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(GOLDEN, "dqn_golden.npz"))
    net = O.make_net(100, [64, 64], 27, 0)
    p0, t0 = g["ddqn_qvalue3_local0"], g["ddqn_qvalue3_target0"]
    s, a, r, s2, d = (g["batch_" + k][0] for k in ("s", "a", "r", "s2", "d"))
    B = s.shape[0]
    lo, hi = rank * B // world, (rank + 1) * B // world
    # local gradient of the LOCAL mean loss (ora_dqn_update reports it), rescaled to the global-batch convention
    L = O.OracleLearner(net, O.ALGO_DDQN, p0); L.target[:] = t0
    _, g_local = L.update(s[lo:hi], a[lo:hi], r[lo:hi], s2[lo:hi], d[lo:hi])
    gt = torch.from_numpy(g_local * ((hi - lo) / B))
    dist.all_reduce(gt, op=dist.ReduceOp.SUM)
    # identical Adam step on every rank from the all-reduced gradient
    m = 0.1 * gt.numpy(); v = 0.001 * gt.numpy() ** 2
    p = p0 - (5e-4 / 0.1) * (m / (np.sqrt(v) / np.sqrt(0.001) + 1e-8))
    gathered = [torch.zeros_like(gt) for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(p.astype(np.float32)))
    if rank == 0:
        full = O.OracleLearner(net, O.ALGO_DDQN, p0); full.target[:] = t0
        _, g_full = full.update(s, a, r, s2, d)
        np.save(out_path, np.stack([gt.numpy(), g_full, gathered[0].numpy(), gathered[1].numpy(), full.local]))
    dist.barrier()
    dist.destroy_process_group()


def test_dp_gradient_convention_two_ranks(tmp_path):
    out = str(tmp_path / "dp.npy")
    mp.spawn(_worker, args=(2, 29533, out), nprocs=2, join=True)
    g_dp, g_full, p_r0, p_r1, p_full = np.load(out)
    np.testing.assert_allclose(g_dp, g_full, rtol=1e-5, atol=1e-6)       # sum of rescaled shard gradients == full-batch gradient
    assert np.array_equal(p_r0, p_r1)                                    # replicas stay bit-identical
    np.testing.assert_allclose(p_r0, p_full, atol=2e-6)                  # and equal the single-process step


def test_bench_reference_arm_under_torchrun():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29544", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--envs", "256", "--pool", "64"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1                                               # rank 0 alone prints
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["unit"] == "env_steps/s"
