"""Softmax / MultinoulliNLL / AggregateLoss / Errors (TEST INFRASTRUCTURE).

Restates computations.softmax (computations.py:170-177), MultinoulliNLL with
target_is_sparse=True and no weights/masks (loss.py:261-347), AggregateLoss
(loss.py:1346-1363) and Classification/_Errors (loss.py:730-814), i.e. exactly
the nodes the four BASELINE configs put after the conv stack.
"""
import numpy as np

EPS = 1e-5  # loss.py:30
F64 = np.float64


def softmax(x, axis=1):
    x = np.asarray(x, F64)
    e = np.exp(x - x.max(axis, keepdims=True))
    return e / e.sum(axis, keepdims=True)


def multinoulli_nll(pred, target):
    """pred (b,C,z,x,y) probabilities; target (b,1,z,x,y) class index (float or
    int; values outside [0,C) are unlabelled).  Returns nll (b,1,z,x,y)."""
    pred = np.asarray(pred, F64)
    C = pred.shape[1]
    classes = np.arange(C).reshape(1, C, 1, 1, 1)
    onehot = (np.asarray(target) == classes).astype(F64)            # loss.py:271-276
    nll_up = -np.where(onehot != 0, onehot * np.log(pred + EPS), 0.0)  # xlogy0, :318
    n_tot = onehot.sum()                                             # :319, :343
    nll = nll_up * pred.size / (n_tot + EPS) / 1 / C                 # :344-345 (n_indep=1)
    return nll.sum(axis=1, keepdims=True), n_tot                     # :346


def aggregate_loss(nll):
    """AggregateLoss with a single parent and mixing weight 1: mean of means."""
    return float(np.mean(nll))


def loss_and_dlogits(logits, target):
    """loss = AggregateLoss(MultinoulliNLL(Softmax(logits), target)) and its
    gradient w.r.t. the logits (what T.grad would hand to the last conv)."""
    p = softmax(logits, 1)
    nll, n_tot = multinoulli_nll(p, target)
    loss = aggregate_loss(nll)
    C = p.shape[1]
    classes = np.arange(C).reshape(1, C, 1, 1, 1)
    onehot = (np.asarray(target) == classes).astype(F64)
    n_el = nll.size
    # d loss / d p_c = -onehot_c / (p_c+EPS) * size/(n_tot+EPS)/C / n_el
    dp = -onehot / (p + EPS) * (p.size / (n_tot + EPS) / C) / n_el
    dlogits = p * (dp - (dp * p).sum(axis=1, keepdims=True))
    return loss, dlogits, p


def errors(pred, target):
    """mean(argmax(pred) != int16(target)) (loss.py:736-737, 805-807)."""
    cls = np.argmax(pred, axis=1)[:, None]
    gt = np.asarray(target).astype(np.int16)
    return float(np.mean(gt != cls))
