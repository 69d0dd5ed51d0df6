#!/usr/bin/env python
"""Conv + fused max-pool on the three unet3d Conv -> Pool layers: separate launches vs e2_conv3d_fwd_pool
(with / without the argmax, whole unpooled tensor / a small window of it).  E2_ZS_TRACE=1 prints the per-role cycle
counters of CTA 0 for every z-stack launch."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tc_check as t
from elektronn2_b200 import _lib
from elektronn2_b200.devtensor import DevTensor
from elektronn2_b200.ops import ConvOp, PoolOp

LAYERS = {'conv1': (32, (114, 130, 130), 64), 'conv3': (64, (54, 62, 62), 128), 'conv5': (128, (24, 28, 28), 256)}


def main():
    h = _lib.get_handle(0)
    k = (3, 3, 3)
    reps = int(os.environ.get('REPS', '10'))
    for name in sys.argv[1:] or list(LAYERS):
        ci, sp, co = LAYERS[name]
        osp = [s - 2 for s in sp]
        psp = [s // 2 for s in osp]
        xd = t.dev_rand(1, ci, sp, 1)
        g = torch.Generator(device='cuda').manual_seed(2)
        w = torch.randn(co, ci, *k, device='cuda', generator=g) * 0.05
        b = torch.zeros(co, device='cuda')
        yd = DevTensor(1, osp[0], osp[1], osp[2], co)
        fl = 2.0 * np.prod(osp) * co * ci * 27
        op = ConvOp(h, xd, yd, w, b, k, 'relu', 'tf32')
        op.pack()
        res = {}
        for amax in (True, False):
            pd = DevTensor(1, psp[0], psp[1], psp[2], co)
            pop = PoolOp(h, yd, pd, (2, 2, 2), keep_argmax=amax)
            assert op.pool_fusable(pop)
            tag = 'amax' if amax else 'noamax'
            res['conv'] = t.time_ms(op.fwd, reps)
            res['pool_' + tag] = t.time_ms(pop.fwd, reps)
            res['fused_' + tag] = t.time_ms(lambda: op.fwd_pool(pop, True), reps)
            c = [s // 2 - 4 for s in osp]
            win = (c[0], c[0] + 8, c[1], c[1] + 8, c[2], c[2] + 8)
            res['fused_win_' + tag] = t.time_ms(lambda: op.fwd_pool(pop, True, win), reps)
            res['fused_nofull_' + tag] = t.time_ms(lambda: op.fwd_pool(pop, False), reps)
        print(name, ' '.join('%s %.1f us' % (k_, v * 1e3) for k_, v in res.items()),
              '| conv %.0f TF/s fused %.0f TF/s' % (fl / res['conv'] / 1e9, fl / res['fused_amax'] / 1e9), flush=True)


if __name__ == '__main__':
    main()
