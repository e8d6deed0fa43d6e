#!/usr/bin/env python
"""bench.py -- the hot path's headline benchmark (BASELINE.json: env steps/sec + DQN updates/sec,
500x500x100 city) on N B200s of one node, with the reference's CPU path timed beside it.

One "step" = one lockstep training iteration of the workload: every env does
    state -> eps-greedy Q-net action -> UAV step (kinematics, collision, reward, termination,
    100-d observation) -> replay add
followed by ONE DQN update (sample B transitions, TD target, MSE, backward, Adam, periodic hard
target update) -- PathPlan_City.run_thread_OffPolicy + update (Envs/PathPlan_City.py:364-385,757-776).

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     (the CPU arm: the oracle port on all host threads)

Prints ONE JSON line (rank 0).  value = whole-job env steps/s with inputs resident in HBM;
e2e = the same iteration driven through host buffers at every plug-in boundary (H2D/D2H inside the
timed region); roofline = dominant kernel, algorithmic bytes / CUDA-event time vs MEASURED_PEAKS.json;
cpu_baseline = the oracle port on the host cores (bounded sample).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np  # noqa: E402

OBS = 100
NETS = {"qvalue3": ([64, 64], 0), "qnet2": ([64], 0), "vanet2": ([64], 1), "vanet3": ([128, 64], 1)}
ALGOS = {"dqn": 0, "ddqn": 1, "dueling": 2}
# algorithmic bytes / flops per unit
ENV_STEP_BYTES = 563            # per env step (discrete action): obs 400 + reward 4 + flags 3 + action 4 + state r/w 128 + sub-goals 24
ACT_BYTES = 404                 # per env: obs read 400 + action write 4
TRANSITION_BYTES = 812          # per sampled transition
FWD_FLOPS = {"qvalue3": 24448, "qnet2": 2 * (100 * 64 + 64 * 27), "vanet2": 2 * (100 * 64 + 64 * 28),
             "vanet3": 2 * (100 * 128 + 128 * 64 + 64 * 28)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=4096, help="envs per GPU (BASELINE configs[1])")
    ap.add_argument("--batch", type=int, default=0, help="DQN batch per GPU per update (0 = envs: one sample per generated transition)")
    ap.add_argument("--net", default="qvalue3", choices=list(NETS))
    ap.add_argument("--algo", default="dqn", choices=list(ALGOS))
    ap.add_argument("--replay", type=int, default=1 << 20, help="replay capacity per GPU (transitions)")
    ap.add_argument("--pool", type=int, default=2048, help="scenario pool size (host RRT)")
    ap.add_argument("--eps", type=float, default=0.1)
    ap.add_argument("--tc", type=int, default=1, help="1 = tcgen05 3xTF32 tensor-core path for the Q-network (default), 0 = fp32 CUDA cores")
    ap.add_argument("--dp", default="fused", choices=["fused", "nccl"],
                    help="N>1 gradient exchange: fused = one-shot NVLink all-reduce inside the Adam kernel, nccl = torch.distributed")
    ap.add_argument("--pdl", type=int, default=1, help="1 = programmatic dependent launch inside the loop (default), 0 = fully serialised kernels")
    ap.add_argument("--fuse", type=int, default=0, help="1 = get_action + env step as one kernel on the tensor-core path, 0 = two PDL-chained kernels (default, faster)")
    ap.add_argument("--per", type=int, default=0, help="1 = prioritised replay (device SumTree equivalent) instead of uniform sampling")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def load_city():
    g = np.load(os.path.join(ROOT, "tests", "golden", "env_golden.npz"))
    return g["dims"], np.ascontiguousarray(g["buildings"]), g["uav_params"]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tf=d["bf16_tflops"], tf_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tf=1590.0, tf_sustained=1400.0, src="fallback")


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.p, self.path = gpu_index, None, "/tmp/uavrl_clocks_%d.csv" % os.getpid()

    def start(self):
        try:
            self.f = open(self.path, "w")
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.p:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.close()
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.remove(self.path)
        except OSError:
            pass
        return out


def config_dict(a, world):
    return {"workload": "%d parallel UAV envs per GPU x %d GPU, PathPlan_City 500x500x100, 26 cylinder buildings "
                        "(reference config/buildings.xml), discrete-27 actions, %s %s MLP 100-%s-27, batch %d/GPU, "
                        "1 update per lockstep step, replay %d transitions/GPU"
                        % (a.envs, world, a.algo.upper(), a.net, "-".join(map(str, NETS[a.net][0])), a.batch, a.replay),
            "envs_per_gpu": a.envs, "global_envs": a.envs * world, "batch_per_gpu": a.batch, "global_batch": a.batch * world,
            "net": a.net, "algo": a.algo, "replay_per_gpu": a.replay, "prioritised_replay": bool(getattr(a, "per", 0)), "eps": a.eps, "scenario_pool": a.pool,
            "parallelism": "dp%d (env shards + replay shards per GPU, %s)" % (world, "one-shot NVLink all-reduce fused into the Adam kernel" if a.dp == "fused" else "NCCL gradient all-reduce"),
            "l2": "replay ring %d MB/GPU > 126 MB L2, fully prefilled before timing; sampled rows come from all of it"
                  % (a.replay * 412 // 1000000)}


# ============================================================================ CPU arm (oracle port)
def make_oracle_loop(a, n_envs, threads):
    import ctypes as C
    import oracle as O
    import uavrl_b200  # noqa: F401  (host-side scenario generator lives in the product library; no GPU needed)
    from uavrl_b200 import _lib
    dims, b, p = load_city()
    nthreads = O.set_threads(threads)
    cfg = _lib.EnvConfig()
    cfg.n_envs, cfg.max_subgoals = 1, 64
    cfg.len, cfg.width, cfg.h = dims
    cfg.max_v, cfg.min_v, cfg.steering_angle, cfg.max_step, cfg.climb_rate = p[0], p[1], p[2], int(p[3]), 1.0
    cfg.n_buildings, cfg.buildings_host = b.shape[0], b.ctypes.data_as(C.POINTER(C.c_double))
    P = min(a.pool, 512)
    sc = dict(start=np.zeros((P, 3)), goal=np.zeros((P, 3)), heading=np.zeros(P), sub=np.zeros((P, 64, 3)),
              n_sub=np.zeros(P, np.int32))
    vp = lambda x: C.c_void_p(x.ctypes.data)  # noqa: E731
    rc = _lib.lib().uavrl_make_scenarios(C.byref(cfg), 42, P, 30, vp(sc["start"]), vp(sc["goal"]), vp(sc["heading"]),
                                         vp(sc["sub"]), vp(sc["n_sub"]))
    assert rc == 0
    hidden, dueling = NETS[a.net]
    net = O.make_net(OBS, hidden, 27, dueling)
    rng = np.random.default_rng(0)
    params = (rng.uniform(-1, 1, O.net_param_count(net)) * 0.1).astype(np.float32)
    city = O.OracleCity(dims[0], dims[1], dims[2], b)
    par = O.UavParams(p[0], p[1], p[2], 1.0, int(p[3]))
    loop = O.OracleTrainLoop(city, par, sc, n_envs, net, ALGOS[a.algo], params, a.batch, max(8 * n_envs, 4 * a.batch))
    return loop, nthreads


def cpu_baseline(a, seconds):
    """The oracle port timed on the host cores: bounded sample of the same workload."""
    loop, nthreads = make_oracle_loop(a, a.envs, 0)
    loop.iteration(a.eps); loop.iteration(a.eps)                   # warm-up (fills the replay past Batch_Size)
    t0 = time.perf_counter(); it = 0
    while it < 3 or time.perf_counter() - t0 < seconds:
        loop.iteration(a.eps); it += 1
    dt = time.perf_counter() - t0
    return {"value": a.envs * it / dt, "unit": "env_steps/s", "updates_per_s": it / dt, "cores": nthreads,
            "kind": "port",
            "sample": "%d lockstep iterations (%d envs each + 1 %s update, batch %d) of the C oracle port "
                      "(oracle/*.c, OpenMP) in %.1f s" % (it, a.envs, a.algo.upper(), a.batch, dt)}


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    loop, nthreads = make_oracle_loop(a, a.envs, 0)
    for _ in range(max(a.warmup, 2)):
        loop.iteration(a.eps)
    t0 = time.perf_counter()
    for _ in range(a.steps):
        loop.iteration(a.eps)
    dt = time.perf_counter() - t0
    v = a.envs * a.steps / dt
    sample = ("each step = one lockstep iteration of %d envs + 1 %s update (batch %d) on the oracle port "
              "(C restatement of the Python reference; the Python reference itself cannot travel to the GPU box)"
              % (a.envs, a.algo.upper(), a.batch))
    out = {"impl": "reference", "metric": "env steps/sec (+ DQN updates/sec), 500x500x100 city", "value": v,
           "unit": "env_steps/s", "updates_per_s": a.steps / dt, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
           "ms_per_step": 1e3 * dt / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f64 env / f32 learner", "data": "synthetic", "config": config_dict(a, 1),
           "cpu_baseline": {"value": v, "unit": "env_steps/s", "cores": nthreads, "kind": "port", "sample": sample},
           "e2e": {"value": v, "unit": "env_steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out, default=float), flush=True)


# ============================================================================ GPU arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    import uavrl_b200  # noqa: F401
    from uavrl_b200 import _lib, engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus and world > 1:
        a.gpus = world
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner on fd 1 at communicator creation: park stdout on stderr meanwhile so that
        # this process's stdout carries the ONE JSON line only
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize(dev)
        finally:
            os.dup2(saved, 1)
            os.close(saved)

    _lib.lib().uavrl_set_pdl(int(a.pdl))
    _lib.lib().uavrl_set_fuse_act_env(int(a.fuse))
    dims, b, p = load_city()
    city = engine.City(dims[0], dims[1], dims[2], b)
    params = engine.UavParams(p[0], p[1], p[2], 1.0, int(p[3]))
    N, B = a.envs, a.batch
    env = engine.EnvBatch(city, params, N, max_subgoals=64, device=local, auto_reset=True)
    sc = env.make_scenarios(a.pool, seed=42 + rank)
    env.set_pool(sc["start"], sc["goal"], sc["heading"], sc["sub"], sc["n_sub"])
    env.reset(0)
    hidden, dueling = NETS[a.net]
    L = engine.Learner(OBS, hidden, 27, dueling, ALGOS[a.algo], lr=5e-4, gamma=0.99, batch_size=B, update_loop=3,
                       replay_capacity=a.replay, lockstep_envs=N, seed=1234 + rank, device=local)
    L.init_params(0)                       # same seed on every rank: replicas start identical
    if a.per:
        L.per_enable()
    tc_on = L.set_tensor_cores(bool(a.tc))
    stream = torch.cuda.current_stream(dev)
    if world > 1 and a.dp == "fused":
        L.connect_peers(dist, rank, world)

    def iterate(k):
        """k lockstep iterations.  1 GPU: the fused C loop.  N GPUs: env/act/ring per rank, local gradient,
        NCCL all-reduce of the gradient vector, identical Adam step on every rank."""
        if world == 1:
            engine.train_run(env, L, k, a.eps, 1, True, want_stats=False)
            return
        if a.dp == "fused":
            engine.train_run_dp(env, L, k, a.eps, B * world)
            return
        gt = L.grad_tensor()
        for _ in range(k):
            engine.train_run(env, L, 1, a.eps, 1, False, want_stats=False)
            L.compute_grads(B * world)
            dist.all_reduce(gt, op=dist.ReduceOp.SUM)
            L.apply_grads()

    # prefill the replay ring so that sampling spans > L2 worth of rows
    ring_frames = (a.replay + N - 1) // N + 1
    engine.train_run(env, L, ring_frames, 1.0, 1, False, want_stats=False)
    iterate(max(a.warmup, 3))
    torch.cuda.synchronize(dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    launches0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    e0.record(stream)
    iterate(a.steps)
    e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    ms = e0.elapsed_time(e1)
    launches = _lib.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = N * world * a.steps / (ms * 1e-3)

    out = None
    if rank == 0:
        pk = measured_peaks()
        out = {"metric": "env steps/sec (+ DQN updates/sec), 500x500x100 city", "value": value, "unit": "env_steps/s",
               "updates_per_s": a.steps / (ms * 1e-3), "samples_per_s": a.steps * B * world / (ms * 1e-3),
               "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps,
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f64 env state / f32 obs+learner", "data": "synthetic",
               "config": dict(config_dict(a, world), qnet_path=("tcgen05 3xTF32 (fp32-grade)" if tc_on else "fp32 CUDA cores"),
                              launch=("programmatic dependent launch" if a.pdl else "serialised")
                              + (", get_action+step fused" if (a.fuse and tc_on) else "")),
               "clocks": clocks, "gpu_launches": int(launches),
               "host_wall_ms_per_step": 1e3 * t_wall / a.steps}

    # ---- roofline pass: per-kernel CUDA-event time (rank 0's GPU; same workload, events between kernels)
    if rank == 0:
        kp = engine.train_profile(env, L, min(a.steps, 200), a.eps) / float(min(a.steps, 200))   # ms per launch
        fused = bool(a.fuse) and tc_on
        names = ("act+env_step_fused" if fused else "act_eps_greedy", "env_step", "td_target", "fwd_bwd", "weight_grad", "reduce_adam")
        fwd = FWD_FLOPS[a.net]
        n_tgt = 1 if a.algo == "dqn" else 2          # target fwd (+ local fwd on s' for double DQN)
        # per sampled transition 4x fwd (DQN) / 5x fwd (DDQN) = target pass(es) + fwd on s + backward (2 fwd)
        if tc_on:
            alg_flops = {"act_eps_greedy": fwd * N, "env_step": 0, "td_target": n_tgt * fwd * B, "fwd_bwd": 2 * fwd * B,
                         "weight_grad": fwd * B, "reduce_adam": 0}
            alg_bytes = {"act_eps_greedy": ACT_BYTES * N, "env_step": ENV_STEP_BYTES * N, "td_target": (OBS * 4 + 9) * B * n_tgt,
                         "fwd_bwd": (OBS * 4 + 8) * B, "weight_grad": OBS * 4 * B, "reduce_adam": 28 * L.P}
        else:
            alg_flops = {"act_eps_greedy": fwd * N, "env_step": 0, "td_target": 0, "fwd_bwd": (n_tgt + 3) * fwd * B,
                         "weight_grad": 0, "reduce_adam": 0}
            alg_bytes = {"act_eps_greedy": ACT_BYTES * N, "env_step": ENV_STEP_BYTES * N, "td_target": 0,
                         "fwd_bwd": TRANSITION_BYTES * B, "weight_grad": 0, "reduce_adam": 28 * L.P}
        if fused:
            for dct in (alg_flops, alg_bytes):
                dct["act+env_step_fused"] = dct["act_eps_greedy"] + dct["env_step"]
        kernels = {}
        for n_, t_ in zip(names, kp):
            if t_ <= 0:
                continue
            kernels[n_] = {"ms": float(t_), "share": float(t_ / kp.sum()), "GBps": alg_bytes[n_] / (t_ * 1e-3) / 1e9,
                           "TFLOPs": alg_flops[n_] / (t_ * 1e-3) / 1e12}
        dom = max(kernels, key=lambda k: kernels[k]["ms"])
        if alg_flops[dom] > 0:
            ach = kernels[dom]["TFLOPs"]
            roof = {"kernel": dom, "bound": "tensor", "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sustained"], "traffic": None,
                    "note": ("3xTF32 tcgen05 path: 3 tensor-core products per algorithmic product, " if tc_on else "fp32 CUDA-core path, ")
                            + "measured against the %s bf16 tensor peak; HBM view %.1f GB/s of %.0f"
                            % (pk["src"], kernels[dom]["GBps"], pk["hbm"])}
        else:
            ach = kernels[dom]["GBps"]
            roof = {"kernel": dom, "bound": "hbm", "achieved": ach, "peak": pk["hbm"], "unit": "GB/s",
                    "frac": ach / pk["hbm"], "traffic": None}
        # dram bytes per launch of the same kernel from the committed `ncu --set full` capture of this exact command
        # (profiles/r01_ncu_traffic.json; only valid for the default workload it was captured on)
        try:
            with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as f:
                tr = json.load(f)
            if tr.get("envs") == N and tr.get("batch") == B and tr.get("net") == a.net and tr.get("algo") == a.algo and bool(tr.get("tc")) == bool(tc_on):
                roof["traffic"] = tr["dram_bytes_per_launch"].get(dom)
                roof["traffic_source"] = tr.get("source")
        except (OSError, ValueError, KeyError):
            pass
        roof["peak_source"] = pk["src"]
        roof["algorithmic_bytes_per_launch"] = alg_bytes[dom]
        out["roofline"] = roof
        out["kernels"] = kernels

    # ---- e2e: the same iteration driven through host buffers at every plug-in boundary
    if not a.no_e2e:
        from uavrl_b200.plugin_loop import HostDrivenLoop
        hl = HostDrivenLoop(env, L, world, dist if world > 1 else None)
        ke = max(10, min(a.steps, 100))
        hl.run(3, a.eps)
        barrier()
        e0.record(stream)
        t0 = time.perf_counter()
        hl.run(ke, a.eps)
        e1.record(stream)
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        if rank == 0:
            out["e2e"] = {"value": N * world * ke / dt, "unit": "env_steps/s", "h2d_bytes_per_step": hl.h2d_bytes,
                          "d2h_bytes_per_step": hl.d2h_bytes, "steps": ke, "ms_per_step": 1e3 * dt / ke,
                          "what": "per step: get_action(host obs)->host actions, Move_Agent(host actions)->host obs/reward/done, "
                                  "replay add from host arrays, update()->host loss; pinned host memory, copies inside the timed region"}

    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(a, a.cpu_seconds)
    if rank == 0:
        print(json.dumps(out, default=float), flush=True)
    sys.stdout.flush()
    if world > 1:
        dist.barrier(device_ids=[local])
    L.close()
    env.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.batch <= 0:
        args.batch = args.envs
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)
