#!/usr/bin/env python
"""Per layer of examples/unet3d.py: Adam update + weight re-pack as the two steps (e2_adam_step_dev over the layer's
slice, e2_*_pack_weights) against the one-launch form (e2_*_adam_pack_dev)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tc_check as t
from elektronn2_b200 import _lib
from elektronn2_b200.devtensor import DevTensor
from elektronn2_b200.ops import ConvOp, UpConvOp

C = _lib.C
LAYERS = [('conv0', 'conv', 32, 1), ('conv1', 'conv', 64, 32), ('conv2', 'conv', 64, 64), ('conv3', 'conv', 128, 64),
          ('conv4', 'conv', 128, 128), ('conv5', 'conv', 256, 128), ('conv6', 'conv', 256, 256), ('conv7', 'conv', 512, 256),
          ('upconv0', 'upconv', 512, 512), ('mconv0', 'conv', 256, 768), ('mconv1', 'conv', 256, 256),
          ('upconv1', 'upconv', 256, 256), ('mconv2', 'conv', 128, 384), ('mconv3', 'conv', 128, 128),
          ('upconv2', 'upconv', 128, 128), ('mconv4', 'conv', 64, 192), ('mconv5', 'conv', 64, 64)]


def main():
    h = _lib.get_handle(0)
    hyper = torch.tensor([5e-4, 0.9, 0.999, 0.5e-4, 1.0, 0, 0, 0], dtype=torch.float32, device='cuda')
    tot = [0.0, 0.0, 0.0]
    for name, kind, co, ci in LAYERS:
        k = (3, 3, 3) if kind == 'conv' else (2, 2, 2)
        n = co * ci * int(np.prod(k))
        w, g, m, s = [torch.randn(n, device='cuda') * 0.01 for _ in range(4)]
        s.abs_()
        if kind == 'conv':
            op = ConvOp(h, DevTensor(1, 4, 4, 4, ci), DevTensor(1, 2, 2, 2, co), w.view(co, ci, *k), None, k, 'relu', 'tf32')
            fn = 'e2_conv3d_adam_pack_dev'
        else:
            op = UpConvOp(h, DevTensor(1, 2, 2, 2, ci), DevTensor(1, 4, 4, 4, co), w.view(co, ci, *k), None, k, 'relu', 'tf32')
            fn = 'e2_upconv3d_adam_pack_dev'
        op.pack()
        a = t.time_ms(lambda: h.call('e2_adam_step_dev', _lib.ptr(w), _lib.ptr(g), _lib.ptr(m), _lib.ptr(s), n,
                                     _lib.ptr(hyper), C.c_float(1.0), h.stream()), 20)
        b = t.time_ms(op.pack, 20)
        c = t.time_ms(lambda: h.call(fn, C.byref(op.d), _lib.ptr(w), _lib.ptr(g), _lib.ptr(m), _lib.ptr(s), _lib.ptr(hyper),
                                     C.c_float(1.0), _lib.ptr(op.wf), _lib.ptr(op.wd), h.stream()), 20)
        tot[0] += a; tot[1] += b; tot[2] += c
        print('%-8s %4d x %4d x %2d  %6.2f M  adam %6.1f us  pack %6.1f us  fused %6.1f us  (%.0f GB/s)' % (
            name, co, ci, int(np.prod(k)), n / 1e6, a * 1e3, b * 1e3, c * 1e3, 36.0 * n / c / 1e6), flush=True)
    print('total adam %.1f us pack %.1f us fused %.1f us' % tuple(v * 1e3 for v in tot))


if __name__ == '__main__':
    main()
