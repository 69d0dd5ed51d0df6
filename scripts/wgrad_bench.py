#!/usr/bin/env python
"""Micro-benchmark of conv wgrad on the unet3d layer shapes.  usage: wgrad_bench.py [layer ...]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tc_check as t
from elektronn2_b200 import _lib
from elektronn2_b200.devtensor import DevTensor
from elektronn2_b200.ops import ConvOp

LAYERS = {  # name: (ci, in spatial, co)
    'conv1': (32, (114, 130, 130), 64), 'conv2': (64, (56, 64, 64), 64), 'conv3': (64, (54, 62, 62), 128),
    'conv4': (128, (26, 30, 30), 128), 'conv5': (128, (24, 28, 28), 256), 'conv6': (256, (11, 13, 13), 256),
    'conv7': (256, (9, 11, 11), 512), 'mconv0': (768, (14, 18, 18), 256), 'mconv1': (256, (12, 16, 16), 256),
    'mconv2': (384, (20, 28, 28), 128), 'mconv3': (128, (18, 26, 26), 128), 'mconv4': (192, (32, 48, 48), 64),
    'mconv5': (64, (30, 46, 46), 64),
}


def main():
    names = sys.argv[1:] or list(LAYERS)
    h = _lib.get_handle(0)
    k = (3, 3, 3)
    tot = 0.0
    for name in names:
        ci, sp, co = LAYERS[name]
        osp = [s - 2 for s in sp]
        xd = t.dev_rand(1, ci, sp, 1)
        g = torch.Generator(device='cuda').manual_seed(2)
        w = torch.randn(co, ci, *k, device='cuda', generator=g) * 0.05
        yd = DevTensor(1, osp[0], osp[1], osp[2], co)
        op = ConvOp(h, xd, yd, w, None, k, 'relu', 'tf32')
        op.pack()
        dy = t.dev_rand(1, co, osp, 3, signed=True)
        dw = torch.zeros_like(w)
        db = torch.zeros(co, device='cuda')
        ms = t.time_ms(lambda: op.wgrad(dy, dw, db), 10)
        fl = 2.0 * np.prod(osp) * co * ci * 27
        tot += ms
        print('%-7s %4d->%4d %-14s wgrad %.3f ms  %.1f TF/s' % (name, ci, co, osp, ms, fl / ms / 1e9), flush=True)
    print('total %.3f ms' % tot)


if __name__ == '__main__':
    main()
