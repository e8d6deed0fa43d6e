"""Host-driven lockstep loop: the data flow a reference-side plug-in has when simulator.py drives the
library (every boundary carries HOST arrays, as in the reference's own Python objects):

    a      = Trainer.get_action(state, eps)          host obs  -> device -> Q-net -> host actions
    s2,r,d = BaseEnv.Move_Agent(action)              host actions -> device step -> host obs/reward/done
    Trainer.replay_memory.add(s, a, r, s2, d)        host arrays -> device replay
    Trainer.update(...)                              device update -> host loss

bench.py times this as the `e2e` number (pinned host memory, all copies inside the timed region).
"""
import torch

from . import engine
from ._lib import ACT_DISCRETE27, OBS_DIM


class HostDrivenLoop:
    def __init__(self, env, lockstep_learner, world=1, dist=None, replay_frames=64):
        self.env, self.world, self.dist = env, world, dist
        src = lockstep_learner
        N = env.n
        dev = env.device
        self.L = engine.Learner(src.in_dim, src.hidden, src.n_actions, src.dueling, src.cfg.algo, lr=src.cfg.lr,
                                gamma=src.cfg.gamma, batch_size=src.cfg.batch_size, update_loop=src.cfg.update_loop,
                                replay_capacity=replay_frames * N, lockstep_envs=0, seed=7, device=dev.index)
        self.L.set_params(src.get_params(0), 0)
        self.L.set_params(src.get_params(1), 1)
        pin = dict(pin_memory=True)
        self.obs_h = torch.empty((N, OBS_DIM), dtype=torch.float32, **pin)
        self.obs2_h = torch.empty((N, OBS_DIM), dtype=torch.float32, **pin)
        self.a_h = torch.empty(N, dtype=torch.int32, **pin)
        self.r_h = torch.empty(N, dtype=torch.float32, **pin)
        self.d_h = torch.empty(N, dtype=torch.uint8, **pin)
        self.info_h = torch.empty(N, dtype=torch.uint8, **pin)
        self.loss_h = torch.empty(1, dtype=torch.float32, **pin)
        self.obs_d = torch.empty((N, OBS_DIM), dtype=torch.float32, device=dev)
        self.obs2_d = torch.empty((N, OBS_DIM), dtype=torch.float32, device=dev)
        self.a_d = torch.empty(N, dtype=torch.int32, device=dev)
        self.r_d = torch.empty(N, dtype=torch.float32, device=dev)
        self.d_d = torch.empty(N, dtype=torch.uint8, device=dev)
        self.loss_d = torch.zeros(1, dtype=torch.float32, device=dev)
        self.obs_h.copy_(env.observe())
        torch.cuda.synchronize(dev)
        ob = N * OBS_DIM * 4
        # per step: get_action (obs in, actions out) + Move_Agent (actions in; obs, reward, done, info out)
        # + replay add (s, a, r, s2, d in) + update (loss out)
        self.h2d_bytes = ob + N * 4 + (2 * ob + N * 4 + N * 4 + N)
        self.d2h_bytes = N * 4 + (ob + N * 4 + N + N) + 4

    def run(self, iters, eps):
        env, L, dev = self.env, self.L, self.env.device
        st = torch.cuda.current_stream(dev)
        B = L.cfg.batch_size
        for _ in range(iters):
            # Trainer.get_action
            self.obs_d.copy_(self.obs_h, non_blocking=True)
            a = L.act(self.obs_d, eps)
            self.a_h.copy_(a, non_blocking=True)
            st.synchronize()
            # BaseEnv.Move_Agent
            env.step_host(self.a_h, ACT_DISCRETE27, self.obs2_h, self.r_h, self.d_h, self.info_h)
            # ReplayMemory.add from the host arrays
            self.obs_d.copy_(self.obs_h, non_blocking=True)
            self.a_d.copy_(self.a_h, non_blocking=True)
            self.r_d.copy_(self.r_h, non_blocking=True)
            self.obs2_d.copy_(self.obs2_h, non_blocking=True)
            self.d_d.copy_(self.d_h, non_blocking=True)
            L.push(self.obs_d, self.a_d, self.r_d, self.obs2_d, self.d_d)
            # Trainer.update
            if self.world == 1:
                L.update(loss=self.loss_d)
            elif L.replay_size() > B:
                L.compute_grads(B * self.world, loss=self.loss_d)
                self.dist.all_reduce(L.grad_tensor(), op=self.dist.ReduceOp.SUM)
                L.apply_grads()
            # Trainer.update returns the loss as a tensor (the reference does too): its device->host copy is enqueued
            # here and is complete at the next boundary that synchronises (get_action of the following step)
            self.loss_h.copy_(self.loss_d, non_blocking=True)
            self.obs_h, self.obs2_h = self.obs2_h, self.obs_h
        st.synchronize()
        return float(self.loss_h[0])
