"""Adam / SGD parameter updates exactly as the reference writes them
(TEST INFRASTRUCTURE).  optimiser.py:135-165 (SGD), :273-334 (Adam)."""
import numpy as np

F64 = np.float64


class AdamState(object):
    def __init__(self, params):
        self.t = 0.0
        self.m = [np.zeros_like(np.asarray(p, F64)) for p in params]
        self.s = [np.zeros_like(np.asarray(p, F64)) for p in params]


def adam_step(params, grads, state, apply_reg, lr, mom=0.9, beta2=0.999, wd=0.0):
    """One step.  Note: epsilon=1e-5 sits INSIDE the sqrt, and both bias
    corrections are folded into the scalar ``factor`` (optimiser.py:283, 301-304);
    weight decay is the plain L2 term lr*wd*p on apply_reg params only
    (weights yes, biases no: neural.py:162-163, 203-204)."""
    eps = 1e-5
    state.t = 1 + state.t
    t = state.t
    factor = np.sqrt(1 - beta2 ** t) / (1 - mom ** t)
    out = []
    for i, (p, g) in enumerate(zip(params, grads)):
        p = np.asarray(p, F64)
        g = np.asarray(g, F64)
        state.m[i] = mom * state.m[i] + (1.0 - mom) * g
        state.s[i] = beta2 * state.s[i] + (1.0 - beta2) * g * g
        direction = factor * state.m[i] / np.sqrt(state.s[i] + eps)
        if apply_reg[i]:
            mult = apply_reg[i] if apply_reg[i] > 1 else 1.0
            p = p - lr * (direction + wd * p * mult)
        else:
            p = p - lr * direction
        out.append(p)
    return out


def sgd_step(params, grads, last_dir, apply_reg, lr, mom=0.9, wd=0.0):
    out = []
    for i, (p, g) in enumerate(zip(params, grads)):
        p = np.asarray(p, F64)
        last_dir[i] = np.asarray(g, F64) + mom * last_dir[i]
        if apply_reg[i]:
            mult = apply_reg[i] if apply_reg[i] > 1 else 1.0
            p = p - lr * (last_dir[i] + wd * p * mult)
        else:
            p = p - lr * last_dir[i]
        out.append(p)
    return out
