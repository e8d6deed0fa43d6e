"""Trainer plug-in: SAC with continuous actions (Trainer/SAC_Trainer.py, the trainer config/Trainer.xml ships) on the
B200 library.  Constructor takes the reference's parsed Trainer.xml dict (actor / critic / SAC_param sub-dicts)."""
import collections
import os

import numpy as np
import torch

import uavrl_b200  # noqa: F401  (repository root must be on sys.path)
from uavrl_b200 import engine
from uavrl_b200.plugins.xmlconfig import None2Value


class SAC_Trainer_B200:
    ROLE_FILES = ("actor", "critic_1", "critic_2")          # SAC_Trainer.save (:109-119): <role>_SAC_<name>.pth

    def __init__(self, param: dict) -> None:
        actor, critic, sp = param.get('actor'), param.get('critic'), param.get('SAC_param')
        if int(sp.get('IS_Continuous')) != 1:
            raise ValueError("SAC_Trainer_B200 implements the continuous-action branch (IS_Continuous = 1)")
        if actor.get('NetWork') != 'PolicyNetContinuous_SAC' or critic.get('NetWork') != 'QValueNetContinuous_SAC':
            raise ValueError("SAC_Trainer_B200 needs PolicyNetContinuous_SAC / QValueNetContinuous_SAC")
        self.name = param.get('name')
        self.replay_size = int(None2Value(param.get('replay_size'), 1000))
        self.Batch_Size = int(None2Value(param.get('Batch_Size'), 128))
        self.save_loop = int(None2Value(param.get('save_loop'), 10))
        self.Is_Train = int(None2Value(param.get('Is_Train'), 1))
        self.IsPriority_Replay = int(None2Value(param.get('IsPriority_Replay'), 0))
        if self.IsPriority_Replay:
            raise ValueError("prioritised replay is not implemented (SURVEY.md section 8f-3)")
        self.IS_Continuous = 1
        self.w, self.hidden, self.act_dim = int(actor.get('w')), int(actor.get('hiden_dim')), int(actor.get('output'))
        self.lockstep_envs = int(None2Value(param.get('lockstep_envs'), 0))
        self.device_index = int(None2Value(param.get('device'), 0))
        self._learner = engine.SacLearner(
            self.w, self.hidden, self.act_dim, float(actor.get('action_bound')), float(actor.get('lr')), float(critic.get('lr')),
            float(sp.get('alpha_lr')), float(sp.get('target_entropy')), float(sp.get('gamma')), float(sp.get('tau')),
            batch_size=self.Batch_Size, replay_capacity=self.replay_size, lockstep_envs=self.lockstep_envs,
            seed=int(None2Value(param.get('seed'), 42)), device=self.device_index)
        self._learner.init_params(int(None2Value(param.get('seed'), 42)))
        self._dev = self._learner.device
        self._losses = torch.zeros(4, device=self._dev)
        self.loss = 0
        self.model_dir = None2Value(param.get('model_path'), None)
        self.Load_Mod()

    @property
    def epoch(self):
        return self._learner.scalars()["epoch"]

    @property
    def log_alpha(self):
        return self._learner.scalars()["log_alpha"]

    def get_action(self, state, eps=0.0):
        """SAC_Trainer.get_action (:444-448): [a0, a1] for one state, or an [N, 2] array for a batch."""
        s = np.ascontiguousarray(state, np.float32)
        single = s.ndim == 1
        a = self._learner.act(torch.from_numpy(s.reshape(-1, self.w)).to(self._dev)).cpu().numpy()
        return a[0].tolist() if single else a

    def update(self, transition_dict):
        """SAC_Trainer.update (:317-441), continuous branch, on the batch in transition_dict."""
        states = transition_dict['states']
        if all(len(sub) == 0 for sub in states) if isinstance(states, list) else len(states) == 0:      # :322-324
            sc = self._learner.scalars()
            self._learner.set_scalars(sc["log_alpha"], sc["la_m"], sc["la_v"], sc["epoch"] + 1, sc["adam_step"])
            return {'sum_epoch': self.epoch, 'loss': self.loss}
        dev = self._dev
        f = lambda x, shape: torch.as_tensor(np.asarray(x, np.float32).reshape(shape)).to(dev)  # noqa: E731
        s = f(states, (-1, self.w)); s2 = f(transition_dict['next_states'], (-1, self.w))
        a = f(transition_dict['actions'], (-1, self.act_dim))
        r = f(transition_dict['rewards'], (-1,)); d = f(transition_dict['dones'], (-1,))
        self._learner.update_batch(s, a, r, s2, d, None, None, self._losses)
        self.loss = self._losses[0:1]
        if self.save_loop > 0 and self.epoch % self.save_loop == 0:
            self.save()
        return {'sum_epoch': self.epoch, 'loss': self.loss}

    # ---- checkpoints in the reference's format: {'model', 'optimizer', 'epoch'} per network
    def _state_dict(self, role):
        flat = self._learner.get_params(role)
        o, h, a = self.w, self.hidden, self.act_dim
        if role == 0:
            shapes = [("fc1.weight", (h, o)), ("fc1.bias", (h,)), ("fc_mu.weight", (a, h)), ("fc_mu.bias", (a,)),
                      ("fc_std.weight", (a, h)), ("fc_std.bias", (a,))]
        else:
            shapes = [("fc1.weight", (h, o + a)), ("fc1.bias", (h,)), ("fc2.weight", (h, h)), ("fc2.bias", (h,)),
                      ("fc_out.weight", (a, h)), ("fc_out.bias", (a,))]
        sd, off = collections.OrderedDict(), 0
        for nm, shp in shapes:
            n = int(np.prod(shp))
            sd[nm] = torch.from_numpy(flat[off:off + n].reshape(shp).copy()); off += n
        return sd

    def _optim_dict(self, role, lr):
        sd = self._state_dict(role)
        m, v = self._learner.get_params(5 + role), self._learner.get_params(8 + role)
        t = self._learner.scalars()["adam_step"]
        state, off = {}, 0
        for i, p in enumerate(sd.values()):
            k = p.numel()
            state[i] = {'step': torch.tensor(float(t)), 'exp_avg': torch.from_numpy(m[off:off + k].reshape(p.shape).copy()),
                        'exp_avg_sq': torch.from_numpy(v[off:off + k].reshape(p.shape).copy())}
            off += k
        return {'state': state, 'param_groups': [{'lr': lr, 'betas': (0.9, 0.999), 'eps': 1e-08, 'weight_decay': 0,
                                                  'amsgrad': False, 'params': list(range(len(sd)))}]}

    def save(self, directory=None):
        directory = directory or self.model_dir or os.path.join(os.getcwd(), 'Mod')
        os.makedirs(directory, exist_ok=True)
        lrs = (self._learner.cfg.actor_lr, self._learner.cfg.critic_lr, self._learner.cfg.critic_lr)
        for role, nm in enumerate(self.ROLE_FILES):
            torch.save({'model': self._state_dict(role), 'optimizer': self._optim_dict(role, lrs[role]), 'epoch': self.epoch},
                       os.path.join(directory, '%s_SAC_%s.pth' % (nm, self.name)))

    def Load_Mod(self, Mod=None):
        """SAC_Trainer.Load_Mod (:70-106): resume actor / critics (+ their Adam moments); targets copy the critics."""
        directory = self.model_dir or os.path.join(os.getcwd(), 'Mod')
        paths = [os.path.join(directory, '%s_SAC_%s.pth' % (nm, self.name)) for nm in self.ROLE_FILES]
        if not all(os.path.exists(p) for p in paths):
            return
        try:
            # parse all three checkpoints into host arrays first; only a fully parsed set touches the live learner
            epoch, step, staged = 0, 0, []
            for role, p in enumerate(paths):
                ck = torch.load(p, weights_only=False, map_location='cpu')
                flat = np.concatenate([v.detach().cpu().numpy().ravel() for v in ck['model'].values()]).astype(np.float32)
                if flat.size != self._learner.P[role]:
                    raise ValueError("%s holds %d parameters, the network has %d" % (p, flat.size, self._learner.P[role]))
                m = v = None
                st = ck['optimizer'].get('state', {})
                if st:
                    keys = sorted(st.keys())
                    m = np.concatenate([st[k]['exp_avg'].detach().cpu().numpy().ravel() for k in keys]).astype(np.float32)
                    v = np.concatenate([st[k]['exp_avg_sq'].detach().cpu().numpy().ravel() for k in keys]).astype(np.float32)
                    if m.size != flat.size or v.size != flat.size:
                        raise ValueError("%s: optimizer state does not match the model" % p)
                    step = int(float(st[keys[0]]['step']))
                epoch = int(ck['epoch'])
                staged.append((role, flat, m, v))
        except Exception as e:              # the reference prints and carries on (SAC_Trainer.py:105-106); nothing was applied
            print(e.args)
            return
        for role, flat, m, v in staged:
            self._learner.set_params(role, flat)
            if role > 0:
                self._learner.set_params(role + 2, flat)
            if m is not None:
                self._learner.set_params(5 + role, m)
                self._learner.set_params(8 + role, v)
        sc = self._learner.scalars()
        self._learner.set_scalars(sc["log_alpha"], sc["la_m"], sc["la_v"], epoch, step)

    def hard_update(self):
        pass

    def _learner_reset_lockstep(self):
        pass
