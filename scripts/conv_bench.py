#!/usr/bin/env python
"""Micro-benchmark of conv fwd / dgrad on the unet3d layer shapes.  usage: conv_bench.py [layer ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tc_check as t
from wgrad_bench import LAYERS
from elektronn2_b200 import _lib
from elektronn2_b200.devtensor import DevTensor
from elektronn2_b200.ops import ConvOp


def main():
    names = sys.argv[1:] or list(LAYERS)
    h = _lib.get_handle(0)
    k = (3, 3, 3)
    tot = [0.0, 0.0]
    for name in names:
        ci, sp, co = LAYERS[name]
        osp = [s - 2 for s in sp]
        xd = t.dev_rand(1, ci, sp, 1)
        g = torch.Generator(device='cuda').manual_seed(2)
        w = torch.randn(co, ci, *k, device='cuda', generator=g) * 0.05
        b = torch.zeros(co, device='cuda')
        yd = DevTensor(1, osp[0], osp[1], osp[2], co)
        op = ConvOp(h, xd, yd, w, b, k, 'relu', 'tf32')
        op.pack()
        dy = t.dev_rand(1, co, osp, 3, signed=True)
        dx = DevTensor(1, sp[0], sp[1], sp[2], ci)
        gate = t.dev_rand(1, ci, sp, 5, signed=True)
        fl = 2.0 * np.prod(osp) * co * ci * 27
        f = t.time_ms(op.fwd, 10)
        d = t.time_ms(lambda: op.dgrad(dy, dx, relu_gate=gate), 10)
        tot[0] += f
        tot[1] += d
        print('%-7s %4d->%4d %-14s fwd %.3f ms %6.1f TF/s   dgrad(gated) %.3f ms %6.1f TF/s' % (
            name, ci, co, osp, f, fl / f / 1e9, d, fl / d / 1e9), flush=True)
    print('total fwd %.3f ms dgrad %.3f ms' % tuple(tot))


if __name__ == '__main__':
    main()
