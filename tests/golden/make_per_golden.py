#!/usr/bin/env python
"""per_golden.npz: the REFERENCE's own prioritised-replay structures (BaseClass/replay_buffer.py:57-223, SumTree and
ReplayTree) executed here on fixed inputs: pushes with |TD error| tensors (incl. wrap-around of the ring), stratified
sampling with the uniform draws recorded (random.uniform patched to a tape), importance weights, beta schedule,
batch_update, and a second sampling round.  Two capacities: 1024 (complete tree) and 1000 (leaves on two levels, which
rotates the left-to-right leaf order).  Run in the build container only:  python tests/golden/make_per_golden.py"""
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

ref_harness.load_reference()
import torch  # noqa: E402
import BaseClass.replay_buffer as rb  # noqa: E402  (reference)


def run_case(cap, n_push, B, rng):
    tree = rb.ReplayTree(cap)
    out = {}
    err = rng.gamma(1.5, 0.4, size=n_push).astype(np.float32)
    for i in range(n_push):
        tree.push((np.float32(i), 0, 0.0, np.float32(i), 0), torch.tensor(err[i]))
    out["push_err"] = err
    out["leaves_after_push"] = tree.tree.tree[-cap:].copy()
    out["n_entries"] = np.int64(tree.tree.n_entries)
    out["data_pointer"] = np.int64(tree.tree.data_pointer)
    orig = random.uniform
    try:
        for rnd in range(2):
            u = rng.random(B)
            k = [0]

            def tape(a, b):
                v = a + (b - a) * u[k[0]]
                k[0] += 1
                return v
            random.uniform = tape
            s, a_, r_, s2, d_, idxs, w = tree.sample2(B)
            out["u%d" % rnd] = u
            out["idx%d" % rnd] = np.asarray(idxs, np.int64)              # tree indices (leaf = data + cap - 1)
            out["w%d" % rnd] = np.asarray(w, np.float64)
            out["beta%d" % rnd] = np.float64(tree.beta)
            out["total%d" % rnd] = np.float64(tree.tree.tree[0])
            out["data%d" % rnd] = np.asarray(s, np.float32)               # payload: the push index
            if rnd == 0:
                ae = rng.gamma(1.2, 0.5, size=B).astype(np.float32)       # some above the clip at 1
                # the same transition sampled twice carries the same error
                first = {}
                for j, t in enumerate(idxs):
                    ae[j] = ae[first.setdefault(t, j)]
                out["abs_err"] = ae.copy()
                tree.batch_update(idxs, ae.copy())
                out["leaves_after_update"] = tree.tree.tree[-cap:].copy()
    finally:
        random.uniform = orig
    return out


if __name__ == "__main__":
    rng = np.random.default_rng(7)
    res = {}
    for name, cap, n_push, B in (("p2", 1024, 1500, 64), ("np2", 1000, 1300, 50), ("part", 4096, 700, 32)):
        for k, v in run_case(cap, n_push, B, rng).items():
            res["%s_%s" % (name, k)] = v
        res["%s_cap" % name] = np.int64(cap)
        res["%s_B" % name] = np.int64(B)
    res["numpy_version"] = np.array(np.__version__)
    res["torch_version"] = np.array(torch.__version__)
    np.savez_compressed(os.path.join(HERE, "per_golden.npz"), **res)
    print({k: (v.shape if hasattr(v, "shape") else v) for k, v in list(res.items())[:12]})
