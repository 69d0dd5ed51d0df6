#!/usr/bin/env python
"""First-layer (one input channel) conv fwd / wgrad micro-benchmark: unet3d conv0 shape."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tc_check as t
from elektronn2_b200 import _lib
from elektronn2_b200.devtensor import DevTensor
from elektronn2_b200.ops import ConvOp

h = _lib.get_handle(0)
k = (3, 3, 3)
sp = (116, 132, 132)
co = 32
osp = [s - 2 for s in sp]
g = torch.Generator(device='cuda').manual_seed(2)
w = torch.randn(co, 1, *k, device='cuda', generator=g) * 0.2
b = torch.randn(co, device='cuda', generator=g) * 0.1
res = {}
for pitch in (1, 4):
    xd = DevTensor(1, sp[0], sp[1], sp[2], 1, c_pitch=pitch)
    v = torch.rand(sp[0] * sp[1] * sp[2], device='cuda', generator=torch.Generator(device='cuda').manual_seed(1))
    v = (v.view(torch.int32) & ~0x1fff).view(torch.float32)      # tf32-exact values
    xd.buf.view(-1, pitch)[:, 0].copy_(v)
    yd = DevTensor(1, osp[0], osp[1], osp[2], co)
    op = ConvOp(h, xd, yd, w, b, k, 'relu', 'tf32')
    op.pack()
    op.fwd()
    dy = t.dev_rand(1, co, osp, 3, signed=True)
    dw = torch.zeros_like(w)
    db = torch.zeros(co, device='cuda')
    op.wgrad(dy, dw, db)
    torch.cuda.synchronize()
    res[pitch] = (t.view(yd).clone(), dw.clone(), db.clone())
    f = t.time_ms(op.fwd, 10)
    wg = t.time_ms(lambda: op.wgrad(dy, dw, db), 10)
    print('pitch %d: fwd %.3f ms (%.0f GB/s out)  wgrad %.3f ms' % (pitch, f, yd.desc.floats * 4 / f / 1e6, wg), flush=True)
print('checksums: y %.6e  dw %.6e  db %.6e  |y| %.6e' % (float(res[1][0].double().sum()), float(res[1][1].double().sum()), float(res[1][2].double().sum()), float(res[1][0].double().abs().sum())))
print('fwd rel diff pitch4 vs pitch1: %.2e   wgrad %.2e  db %.2e' % (t.rel(res[4][0], res[1][0]), t.rel(res[4][1], res[1][1]),
                                                                    t.rel(res[4][2], res[1][2])))
