#!/usr/bin/env python
"""First-layer conv (unet3d conv0) forward: tensor-core kernel vs CUDA-core kernel, check + time."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tc_check as t
from elektronn2_b200 import _lib
from elektronn2_b200.devtensor import DevTensor
from elektronn2_b200.ops import ConvOp
h = _lib.get_handle(0)
for (sp, co, k) in [((116, 132, 132), 32, (3, 3, 3)), ((23, 185, 185), 20, (1, 6, 6)), ((22, 140, 140), 20, (1, 3, 3))]:
    osp = [s - f + 1 for s, f in zip(sp, k)]
    xd = t.dev_rand(1, 1, sp, 1)
    g = torch.Generator(device='cuda').manual_seed(2)
    w = torch.randn(co, 1, *k, device='cuda', generator=g) * 0.2
    b = torch.randn(co, device='cuda', generator=g) * 0.1
    outs = {}
    for comp in ('f32', 'tf32'):
        yd = DevTensor(1, osp[0], osp[1], osp[2], co)
        op = ConvOp(h, xd, yd, w, b, k, 'relu', comp)
        op.pack()
        op.fwd()
        torch.cuda.synchronize()
        outs[comp] = t.view(yd).clone()
        ms = t.time_ms(op.fwd, 10)
        print(sp, co, k, comp, '%.3f ms' % ms, '%.0f GB/s' % (4 * np.prod(osp) * co / ms / 1e6), flush=True)
    print('  rel err tf32 vs f32: %.2e' % t.rel(outs['tf32'], outs['f32']), flush=True)
