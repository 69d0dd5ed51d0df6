#!/usr/bin/env python
"""How far apart can two correct TF32-mode implementations be?  (CPU only; writes tests/golden/tf32_conditioning.json)

The TF32 mode of libe2b200 prescribes WHERE values are rounded to tf32 (oracle/nets.py ``Net.tf32``), not the order in
which the fp32 accumulations inside a convolution happen.  A different order moves an accumulator by ~2^-22 relative;
that is enough to flip the tf32 rounding (10 mantissa bits) of the occasional activation, and the network amplifies the
flips.  This script runs the tf32-operand oracle twice -- once as is, once with every convolution accumulator multiplied
by (1 + u * 2^-22), u uniform in [-1, 1] -- and records the largest relative difference of loss, probabilities and every
parameter gradient.  The GPU test (tests/test_gpu_nets.py::test_tf32_gradients_match_tf32_operand_oracle) uses these
numbers as the tolerance: the GPU may deviate from the tf32-operand oracle by about as much as that oracle deviates from
itself under accumulation-order noise, not more.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import nets as onets, ops  # noqa: E402


def rel(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def run(name, noise_seed):
    o = onets.BUILDERS[name]()
    ish = [1 if s is None else s for s in o.nodes[0].sh.shape]
    x = np.random.RandomState(0).rand(*ish).astype(np.float32)          # the inputs the GPU test uses
    t = np.random.RandomState(1).randint(0, 2, [ish[0], 1] + list(o.nodes[-1].sh.spatial)).astype(np.float32)
    o.tf32 = True
    saved = ops.conv3d, ops.conv3d_dot, ops.conv3d_dgrad, ops.conv3d_wgrad, ops.upconv3d, ops.upconv3d_dgrad
    if noise_seed is not None:
        rs = np.random.RandomState(noise_seed)

        def noisy(f):
            def g(*a, **k):
                y = f(*a, **k)
                return y * (1.0 + rs.uniform(-1, 1, y.shape) * 2.0 ** -22)
            return g
        ops.conv3d, ops.conv3d_dot, ops.conv3d_dgrad = noisy(saved[0]), noisy(saved[1]), noisy(saved[2])
        ops.upconv3d, ops.upconv3d_dgrad = noisy(saved[4]), noisy(saved[5])
    try:
        L, g, p, _ = o.loss_and_grads(x, t)
    finally:
        ops.conv3d, ops.conv3d_dot, ops.conv3d_dgrad, ops.conv3d_wgrad, ops.upconv3d, ops.upconv3d_dgrad = saved
    return L, [g[k] for k in o.param_list()], p


def main():
    out = {}
    for name in ('neuro3d_lite', 'unet3d_litelite'):
        L0, g0, p0 = run(name, None)
        worst = 0.0
        per_seed = []
        for seed in (7, 8):
            L1, g1, p1 = run(name, seed)
            w = max(rel(a, b) for a, b in zip(g1, g0))
            per_seed.append(dict(seed=seed, loss_rel=abs(L1 - L0) / abs(L0), probs_rel=rel(p1, p0), worst_grad_rel=w))
            worst = max(worst, w)
        out[name] = dict(worst_grad_rel=worst, runs=per_seed,
                         note='tf32-operand oracle vs itself with 2^-22 relative noise on every conv accumulator')
        print(name, out[name])
    json.dump(out, open(os.path.join(ROOT, 'tests', 'golden', 'tf32_conditioning.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
