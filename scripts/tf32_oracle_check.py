#!/usr/bin/env python
"""Diagnostic: GPU TF32 mode vs (a) the float64 oracle, (b) the oracle with tf32 operand roundings (oracle/nets.py tf32)."""
import contextlib, io, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elektronn2_b200 import examples, neuromancer as nm
from oracle import nets as onets

def rel(a, b):
    b = np.asarray(b, np.float64)
    return float(np.abs(np.asarray(a, np.float64) - b).max() / max(np.abs(b).max(), 1e-30))

for name in sys.argv[1:] or ['unet3d_litelite', 'neuro3d_lite']:
    nm.model_manager.reset()
    np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        m = examples.BUILDERS[name]()
    ish = [1 if s is None else s for s in m.input_node.shape.shape]
    tsh = [1 if s is None else s for s in m.target_node.shape.shape]
    x = np.random.RandomState(0).rand(*ish).astype(np.float32)
    t = np.random.RandomState(1).randint(0, 2, tsh).astype(np.float32)
    loss, err, p = m.predict_ext(x, t)
    g = m.gradients(x, t)
    for mode in (False, True):
        o = onets.BUILDERS[name]()
        o.tf32 = mode
        L, grads, probs, _ = o.loss_and_grads(x, t)
        ref = [grads[k] for k in o.param_list()]
        errs = [rel(a, b) for a, b in zip(g, ref)]
        print('%s oracle tf32=%s: loss rel %.2e probs rel %.2e worst grad %.2e' % (name, mode, abs(loss - L) / abs(L), rel(p, probs), max(errs)))
        print('   per param:', ' '.join('%.1e' % e for e in errs))
        l2 = [float(np.linalg.norm(np.asarray(a, np.float64) - b) / max(np.linalg.norm(b), 1e-30)) for a, b in zip(g, ref)]
        print('   per param L2-rel:', ' '.join('%.1e' % e for e in l2), ' worst %.2e' % max(l2))
        # per-layer forward values
        convs = [n for n in m.nodes.values() if type(n).__name__ in ('Conv', 'UpConv')]
        oconvs = [n for n in o.nodes if n.op in ('conv', 'upconv')]
        fw = []
        for a, b in zip(convs, oconvs):
            fw.append(rel(a(x), o.val[b]))
        print('   per layer fwd:', ' '.join('%.1e' % e for e in fw))
