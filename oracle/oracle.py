"""ctypes front-end of the CPU oracle (TEST INFRASTRUCTURE -- see uav_oracle.h / dqn_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does.
"""
import ctypes as C
import math
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def build(force=False):
    srcs = [os.path.join(HERE, f) for f in ("uav_oracle.c", "dqn_oracle.c", "sac_oracle.c", "uav_oracle.h", "dqn_oracle.h", "sac_oracle.h")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return LIB_PATH
    subprocess.check_call(["make", "-s", "-C", HERE, "-B"])
    return LIB_PATH


class Loc(C.Structure):
    _fields_ = [("x", C.c_double), ("y", C.c_double), ("z", C.c_double)]


class City(C.Structure):
    _fields_ = [("len", C.c_double), ("width", C.c_double), ("h", C.c_double),
                ("n_buildings", C.c_int32), ("buildings", C.POINTER(C.c_double))]


class Batch(C.Structure):
    _fields_ = [("n", C.c_int32), ("kmax", C.c_int32)] + \
        [(k, C.POINTER(C.c_double)) for k in ("px", "py", "pz", "vx", "vy", "V")] + \
        [(k, C.POINTER(C.c_int32)) for k in ("step", "cursor", "n_sub")] + \
        [(k, C.POINTER(C.c_uint8)) for k in ("done", "alias0")] + \
        [(k, C.POINTER(C.c_double)) for k in ("score", "total_score", "path_len", "goal", "sub")]


class Net(C.Structure):
    _fields_ = [("in_dim", C.c_int32), ("n_hidden", C.c_int32), ("hidden", C.c_int32 * 4),
                ("n_actions", C.c_int32), ("dueling", C.c_int32)]


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB_PATH)
        _lib.ora_distance.restype = C.c_double
        _lib.ora_distance.argtypes = [Loc, Loc]
        _lib.ora_angle.restype = C.c_double
        _lib.ora_angle.argtypes = [Loc, Loc]
        _lib.ora_threaten_rate.restype = C.c_int32
        _lib.ora_threaten_rate.argtypes = [C.POINTER(City), Loc]
        _lib.ora_fly_power.restype = C.c_double
        _lib.ora_fly_power.argtypes = [C.c_double] * 10
        _lib.ora_net_param_count.restype = C.c_int64
        _lib.ora_dqn_update.restype = C.c_float
        _lib.ora_dqn_update_per.restype = C.c_float
        _lib.ora_sac_update.restype = C.c_float
        _lib.ora_sac_actor_params.restype = C.c_int64
        _lib.ora_sac_critic_params.restype = C.c_int64
        _lib.ora_per_create.restype = C.c_void_p
        _lib.ora_per_create.argtypes = [C.c_int32]
        _lib.ora_per_destroy.argtypes = [C.c_void_p]
        _lib.ora_per_push.argtypes = [C.c_void_p, C.c_float]
        _lib.ora_per_add.argtypes = [C.c_void_p, C.c_double]
        _lib.ora_per_sample.restype = C.c_double
        _lib.ora_per_sample.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ora_per_batch_update.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p]
        _lib.ora_per_leaves.argtypes = [C.c_void_p, C.c_void_p]
        _lib.ora_per_total.restype = C.c_double
        _lib.ora_per_total.argtypes = [C.c_void_p]
        _lib.ora_per_n_entries.restype = C.c_int32
        _lib.ora_per_n_entries.argtypes = [C.c_void_p]
    return _lib


def _p(a, ct):
    return a.ctypes.data_as(C.POINTER(ct))


ACT_CONTINUOUS, ACT_DISCRETE27 = 0, 1
ALGO_DQN, ALGO_DDQN, ALGO_DUELING = 0, 1, 2
INFO_NAMES = ("normal", "success", "lose")


class OracleCity:
    """City = box dims + cylinder table [n,5] (cx, cy, cz, R, H)."""

    def __init__(self, length, width, h, buildings):
        self.buildings = np.ascontiguousarray(buildings, dtype=np.float64).reshape(-1, 5)
        self.c = City(float(length), float(width), float(h), self.buildings.shape[0],
                      _p(self.buildings, C.c_double))

    def threaten_rate(self, pts):
        pts = np.asarray(pts, dtype=np.float64).reshape(-1, 3)
        L = lib()
        return np.array([L.ora_threaten_rate(C.byref(self.c), Loc(*p)) for p in pts], dtype=np.uint8)


class UavParams:
    def __init__(self, max_v=1.0, min_v=0.6, steering=np.pi / 6, climb_rate=1.0, max_step=150):
        self.max_v, self.min_v, self.steering = float(max_v), float(min_v), float(steering)
        self.climb_rate, self.max_step = float(climb_rate), int(max_step)


class OracleBatch:
    """SoA state of n UAVs, stepped by the C oracle (ora_batch_step)."""

    F64 = ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len")

    def __init__(self, city, params, n, kmax):
        self.city, self.p, self.n, self.kmax = city, params, n, kmax
        for k in self.F64:
            setattr(self, k, np.zeros(n, np.float64))
        self.step = np.zeros(n, np.int32)
        self.cursor = np.zeros(n, np.int32)
        self.n_sub = np.zeros(n, np.int32)
        self.done = np.zeros(n, np.uint8)
        self.alias0 = np.zeros(n, np.uint8)
        self.goal = np.zeros((n, 3), np.float64)
        self.sub = np.zeros((n, kmax, 3), np.float64)

    def reset(self, start, goal, heading, sub, n_sub, alias0=None):
        """UAV.reset() (Agents/UAV.py:335-366) with the random draws supplied by the caller."""
        n = self.n
        start = np.asarray(start, np.float64).reshape(n, 3)
        self.px[:], self.py[:], self.pz[:] = start[:, 0], start[:, 1], start[:, 2]
        heading = np.asarray(heading, np.float64).reshape(n)
        # UAV.py:344-348 with python-float (libm) arithmetic, then Calc_V (UAV.py:246-253)
        for e in range(n):
            self.vx[e] = self.p.max_v * math.cos(float(heading[e]))
            self.vy[e] = self.p.max_v * math.sin(float(heading[e]))
            V = math.sqrt(float(self.vx[e]) ** 2 + float(self.vy[e]) ** 2 + 0.0 ** 2)
            if V > self.p.max_v:
                self.vx[e] = self.vx[e] * (self.p.max_v / V)
                self.vy[e] = self.vy[e] * (self.p.max_v / V)
                V = self.p.max_v
            self.V[e] = V
        self.step[:] = 0
        self.cursor[:] = 0
        self.done[:] = 0
        self.score[:] = 0
        self.total_score[:] = 0
        self.path_len[:] = 0
        self.goal[:] = np.asarray(goal, np.float64).reshape(n, 3)
        self.sub[:] = 0
        sub = np.asarray(sub, np.float64)
        self.sub[:, :sub.shape[1], :] = sub
        self.n_sub[:] = np.asarray(n_sub, np.int32)
        self.alias0[:] = 1 if alias0 is None else np.asarray(alias0, np.uint8)

    def _struct(self):
        b = Batch()
        b.n, b.kmax = self.n, self.kmax
        for k in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len"):
            setattr(b, k, _p(getattr(self, k), C.c_double))
        b.goal = _p(self.goal, C.c_double)
        b.sub = _p(self.sub, C.c_double)
        for k in ("step", "cursor", "n_sub"):
            setattr(b, k, _p(getattr(self, k), C.c_int32))
        b.done = _p(self.done, C.c_uint8)
        b.alias0 = _p(self.alias0, C.c_uint8)
        return b

    def step_(self, actions, act_mode=ACT_CONTINUOUS, want_obs=True):
        n = self.n
        actions = np.ascontiguousarray(actions, np.float64).reshape(n)
        rew = np.zeros(n, np.float64)
        done = np.zeros(n, np.uint8)
        info = np.zeros(n, np.uint8)
        coll = np.zeros(n, np.uint8)
        obs = np.zeros((n, 100), np.float32) if want_obs else None
        b = self._struct()
        lib().ora_batch_step(C.byref(self.city.c), C.byref(b), C.c_double(self.p.max_v),
                             C.c_double(self.p.min_v), C.c_double(self.p.steering),
                             C.c_double(self.p.climb_rate), C.c_int32(self.p.max_step),
                             C.c_int32(act_mode), _p(actions, C.c_double), _p(rew, C.c_double),
                             _p(done, C.c_uint8), _p(info, C.c_uint8), _p(coll, C.c_uint8),
                             _p(obs, C.c_float) if want_obs else None)
        return rew, done, info, coll, obs

    def state(self, want64=False):
        obs = np.zeros((self.n, 100), np.float32)
        obs64 = np.zeros((self.n, 100), np.float64) if want64 else None
        b = self._struct()
        lib().ora_batch_state(C.byref(self.city.c), C.byref(b), C.c_double(self.p.max_v),
                              C.c_double(self.p.min_v), C.c_double(self.p.steering),
                              C.c_double(self.p.climb_rate), C.c_int32(self.p.max_step),
                              _p(obs, C.c_float), _p(obs64, C.c_double) if want64 else None)
        return (obs, obs64) if want64 else obs


def angle(p1, p2):
    return lib().ora_angle(Loc(*p1), Loc(*p2))


def distance(p1, p2):
    return lib().ora_distance(Loc(*p1), Loc(*p2))


# ------------------------------------------------------------------ learner math
def make_net(in_dim, hidden, n_actions, dueling):
    n = Net()
    n.in_dim, n.n_hidden, n.n_actions, n.dueling = in_dim, len(hidden), n_actions, int(dueling)
    for i, h in enumerate(hidden):
        n.hidden[i] = h
    return n


def net_param_count(net):
    return int(lib().ora_net_param_count(C.byref(net)))


def net_forward(net, params, x):
    x = np.ascontiguousarray(x, np.float32)
    params = np.ascontiguousarray(params, np.float32)
    q = np.zeros((x.shape[0], net.n_actions), np.float32)
    lib().ora_net_forward(C.byref(net), _p(params, C.c_float), _p(x, C.c_float),
                          C.c_int32(x.shape[0]), _p(q, C.c_float))
    return q


def act(net, params, x, eps, u, rand_action, is_train=1):
    x = np.ascontiguousarray(x, np.float32)
    params = np.ascontiguousarray(params, np.float32)
    u = np.ascontiguousarray(u, np.float32)
    rand_action = np.ascontiguousarray(rand_action, np.int32)
    B = x.shape[0]
    a = np.zeros(B, np.int32)
    q = np.zeros((B, net.n_actions), np.float32)
    lib().ora_act(C.byref(net), _p(params, C.c_float), _p(x, C.c_float), C.c_int32(B),
                  C.c_float(eps), C.c_int32(is_train), _p(u, C.c_float), _p(rand_action, C.c_int32),
                  _p(a, C.c_int32), _p(q, C.c_float))
    return a, q


class OracleLearner:
    """local/target params + Adam state, updated by ora_dqn_update."""

    def __init__(self, net, algo, params, gamma=0.99, lr=5e-4, update_loop=3):
        self.net, self.algo = net, algo
        self.local = np.array(params, np.float32).copy()
        self.target = self.local.copy()
        self.m = np.zeros_like(self.local)
        self.v = np.zeros_like(self.local)
        self.t = C.c_int64(0)
        self.gamma, self.lr, self.update_loop = gamma, lr, update_loop
        self.epoch = 0

    def update(self, s, a, r, s2, d, is_w=None):
        """DuelingDQN_Trainer.update :150-184 (epoch += 1, step, hard update every Update_loop).  With is_w: the
        prioritised-replay form (weighted loss); returns (loss, grads, abs_err) then."""
        self.epoch += 1
        s = np.ascontiguousarray(s, np.float32)
        s2 = np.ascontiguousarray(s2, np.float32)
        a = np.ascontiguousarray(a, np.int32)
        r = np.ascontiguousarray(r, np.float32)
        d = np.ascontiguousarray(d, np.float32)
        grads = np.zeros_like(self.local)
        args = [C.byref(self.net), C.c_int32(self.algo), _p(self.local, C.c_float),
                _p(self.target, C.c_float), _p(self.m, C.c_float), _p(self.v, C.c_float),
                C.byref(self.t), _p(s, C.c_float), _p(a, C.c_int32), _p(r, C.c_float),
                _p(s2, C.c_float), _p(d, C.c_float), C.c_int32(s.shape[0]), C.c_float(self.gamma),
                C.c_float(self.lr), _p(grads, C.c_float)]
        abs_err = None
        if is_w is None:
            loss = lib().ora_dqn_update(*args)
        else:
            w = np.ascontiguousarray(is_w, np.float32)
            abs_err = np.zeros(s.shape[0], np.float32)
            loss = lib().ora_dqn_update_per(*args, _p(w, C.c_float), _p(abs_err, C.c_float))
        if self.epoch % self.update_loop == 0:
            self.target[:] = self.local          # hard_update, :199-202
        return (float(loss), grads) if is_w is None else (float(loss), grads, abs_err)


_apf_keep = None


def set_apf(obstacle_v):
    """APF_Enabled with moving obstacles (UAV.py:156-210, 448-453): obstacle_v [n, 3] = the obstacles' `v`; None = off."""
    global _apf_keep
    if obstacle_v is None:
        _apf_keep = None
        lib().ora_set_apf(None, 0)
        return
    _apf_keep = np.ascontiguousarray(obstacle_v, np.float64).reshape(-1, 3)
    lib().ora_set_apf(_apf_keep.ctypes.data_as(C.POINTER(C.c_double)), C.c_int32(_apf_keep.shape[0]))


def set_loss_kind(kind):
    """0 = MSE (what every reference trainer uses), 1 = Huber / SmoothL1Loss(beta=1) for the learner oracle."""
    lib().ora_set_loss_kind(C.c_int32({"mse": 0, "huber": 1}.get(kind, kind)))


def set_threads(n=0):
    """Set (n > 0) / query the number of host threads the oracle's batched loops use."""
    return int(lib().ora_set_threads(C.c_int32(n)))


class OracleTrainLoop:
    """CPU restatement of one lockstep training iteration (PathPlan_City.run_thread_OffPolicy + update,
    Envs/PathPlan_City.py:364-385,757-776) for N envs sharing one learner -- the workload bench.py times
    on the GPU, run on the host cores: state -> eps-greedy action -> step -> replay add -> sample ->
    DQN update.  Used only for bench.py's cpu_baseline / --impl reference legs."""

    def __init__(self, city, params, scen, n_envs, net, algo, flat_params, batch, capacity, seed=0, kmax=64):
        self.N, self.B = n_envs, batch
        self.batch = OracleBatch(city, params, n_envs, kmax)
        self.scen, self.P = scen, len(scen["n_sub"])
        self.sidx = np.arange(n_envs) % self.P
        s = self.sidx
        self.batch.reset(scen["start"][s], scen["goal"][s], scen["heading"][s], scen["sub"][s], scen["n_sub"][s])
        self.net = net
        self.learner = OracleLearner(net, algo, flat_params)
        self.rng = np.random.default_rng(seed)
        self.cap = capacity
        self.S = np.zeros((capacity, 100), np.float32); self.S2 = np.zeros((capacity, 100), np.float32)
        self.A = np.zeros(capacity, np.int32); self.R = np.zeros(capacity, np.float32); self.D = np.zeros(capacity, np.float32)
        self.head = 0; self.count = 0
        self.obs = self.batch.state()

    def iteration(self, eps, do_update=True):
        N = self.N
        u = self.rng.uniform(size=N).astype(np.float32)
        ra = self.rng.integers(0, self.net.n_actions, N).astype(np.int32)
        a, _ = act(self.net, self.learner.local, self.obs, eps, u, ra)
        rew, done, info, coll, obs2 = self.batch.step_(a.astype(np.float64), ACT_DISCRETE27)
        idx = (self.head + np.arange(N)) % self.cap
        self.S[idx] = self.obs; self.S2[idx] = obs2; self.A[idx] = a; self.R[idx] = rew; self.D[idx] = done
        self.head = (self.head + N) % self.cap; self.count = min(self.count + N, self.cap)
        ended = np.nonzero(self.batch.done)[0]
        if ended.size:                                   # UAV.reset() at the episode boundary
            self.sidx[ended] = (self.sidx[ended] + N) % self.P
            s = self.sidx[ended]
            tmp = OracleBatch(self.batch.city, self.batch.p, ended.size, self.batch.kmax)
            tmp.reset(self.scen["start"][s], self.scen["goal"][s], self.scen["heading"][s], self.scen["sub"][s],
                      self.scen["n_sub"][s])
            for k in ("px", "py", "pz", "vx", "vy", "V", "score", "total_score", "path_len", "step", "cursor",
                      "n_sub", "done", "alias0"):
                getattr(self.batch, k)[ended] = getattr(tmp, k)
            self.batch.goal[ended] = tmp.goal; self.batch.sub[ended] = tmp.sub
            obs2 = self.batch.state()
        self.obs = obs2
        loss = None
        if do_update and self.count > self.B:
            j = self.rng.choice(self.count, self.B, replace=False)
            loss, _ = self.learner.update(self.S[j], self.A[j], self.R[j], self.S2[j], self.D[j])
        return loss


# ------------------------------------------------------------------ SAC continuous (sac_oracle.c)
class SacCfg(C.Structure):
    _fields_ = [("obs_dim", C.c_int32), ("hidden", C.c_int32), ("act_dim", C.c_int32), ("action_bound", C.c_float),
                ("actor_lr", C.c_float), ("critic_lr", C.c_float), ("alpha_lr", C.c_float), ("target_entropy", C.c_float),
                ("gamma", C.c_float), ("tau", C.c_float)]


class SacState(C.Structure):
    _fields_ = [(k, C.POINTER(C.c_float)) for k in ("actor", "c1", "c2", "t1", "t2", "actor_m", "actor_v", "c1_m", "c1_v", "c2_m", "c2_v")] + \
               [("log_alpha", C.c_float), ("la_m", C.c_float), ("la_v", C.c_float), ("step", C.c_int64)]


class OracleSac:
    def __init__(self, actor, c1, c2, t1, t2, log_alpha, actor_lr=1e-4, critic_lr=1e-3, alpha_lr=1e-4, target_entropy=1.0,
                 gamma=0.99, tau=0.05, obs_dim=100, hidden=64, act_dim=2, action_bound=1.0):
        self.cfg = SacCfg(obs_dim, hidden, act_dim, action_bound, actor_lr, critic_lr, alpha_lr, target_entropy, gamma, tau)
        self.arr = {k: np.array(v, np.float32).copy() for k, v in dict(actor=actor, c1=c1, c2=c2, t1=t1, t2=t2).items()}
        for k in ("actor", "c1", "c2"):
            self.arr[k + "_m"] = np.zeros_like(self.arr[k]); self.arr[k + "_v"] = np.zeros_like(self.arr[k])
        self.st = SacState()
        for k, v in self.arr.items():
            setattr(self.st, k, _p(v, C.c_float))
        self.st.log_alpha = float(log_alpha)
        assert lib().ora_sac_actor_params(C.byref(self.cfg)) == self.arr["actor"].size
        assert lib().ora_sac_critic_params(C.byref(self.cfg)) == self.arr["c1"].size

    @property
    def log_alpha(self):
        return float(self.st.log_alpha)

    def actor_forward(self, s, eps):
        s = np.ascontiguousarray(s, np.float32); eps = np.ascontiguousarray(eps, np.float32)
        B = s.shape[0]
        a = np.zeros((B, self.cfg.act_dim), np.float32); lp = np.zeros_like(a)
        lib().ora_sac_actor_forward(C.byref(self.cfg), _p(self.arr["actor"], C.c_float), _p(s, C.c_float), _p(eps, C.c_float),
                                    C.c_int32(B), _p(a, C.c_float), _p(lp, C.c_float))
        return a, lp

    def critic_forward(self, which, s, a):
        s = np.ascontiguousarray(s, np.float32); a = np.ascontiguousarray(a, np.float32)
        q = np.zeros((s.shape[0], self.cfg.act_dim), np.float32)
        lib().ora_sac_critic_forward(C.byref(self.cfg), _p(self.arr[which], C.c_float), _p(s, C.c_float), _p(a, C.c_float),
                                     C.c_int32(s.shape[0]), _p(q, C.c_float))
        return q

    def update(self, s, a, r, s2, d, eps_next, eps_cur):
        f = lambda x: np.ascontiguousarray(x, np.float32)  # noqa: E731
        s, a, r, s2, d, e1, e2 = map(f, (s, a, r, s2, d, eps_next, eps_cur))
        l1, l2, la = C.c_float(), C.c_float(), C.c_float()
        loss = lib().ora_sac_update(C.byref(self.cfg), C.byref(self.st), _p(s, C.c_float), _p(a, C.c_float), _p(r, C.c_float),
                                    _p(s2, C.c_float), _p(d, C.c_float), _p(e1, C.c_float), _p(e2, C.c_float), C.c_int32(s.shape[0]),
                                    C.byref(l1), C.byref(l2), C.byref(la))
        return float(loss), float(l1.value), float(l2.value)


class OraclePer:
    """The reference's SumTree + ReplayTree (BaseClass/replay_buffer.py:57-223), sequential C restatement."""

    def __init__(self, capacity):
        self.cap = int(capacity)
        self.h = lib().ora_per_create(self.cap)

    def __del__(self):
        if getattr(self, "h", None):
            lib().ora_per_destroy(self.h)
            self.h = None

    def push(self, abs_err):
        for e in np.atleast_1d(np.asarray(abs_err, np.float32)):
            lib().ora_per_push(self.h, float(e))

    def add(self, priorities):
        for p in np.atleast_1d(np.asarray(priorities, np.float64)):
            lib().ora_per_add(self.h, float(p))

    def sample(self, u):
        u = np.ascontiguousarray(u, np.float64)
        idx = np.zeros(len(u), np.int64); w = np.zeros(len(u), np.float64)
        beta = lib().ora_per_sample(self.h, len(u), u.ctypes.data, idx.ctypes.data, w.ctypes.data)
        return idx, w, beta

    def batch_update(self, tree_idx, abs_err):
        ti = np.ascontiguousarray(tree_idx, np.int64); ae = np.ascontiguousarray(abs_err, np.float32)
        lib().ora_per_batch_update(self.h, len(ti), ti.ctypes.data, ae.ctypes.data)

    def leaves(self):
        out = np.zeros(self.cap, np.float64)
        lib().ora_per_leaves(self.h, out.ctypes.data)
        return out

    @property
    def total(self):
        return lib().ora_per_total(self.h)

    @property
    def n_entries(self):
        return lib().ora_per_n_entries(self.h)
