#!/usr/bin/env python
"""Run one upconv op a few times (for ncu).  usage: prof_upconv.py <ci> <z> <x> <y> <co> <op> [reps]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import tc_check as t
from elektronn2_b200 import _lib
from elektronn2_b200.devtensor import DevTensor
from elektronn2_b200.ops import UpConvOp
ci, z, x, y, co = [int(v) for v in sys.argv[1:6]]
opname = sys.argv[6]
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
h = _lib.get_handle(0)
p = (2, 2, 2)
sp = (z, x, y)
osp = [s * 2 for s in sp]
xd = t.dev_rand(1, ci, sp, 1)
g = torch.Generator(device='cuda').manual_seed(2)
w = torch.randn(co, ci, *p, device='cuda', generator=g) * 0.05
b = torch.zeros(co, device='cuda')
yd = DevTensor(1, osp[0], osp[1], osp[2], co)
op = UpConvOp(h, xd, yd, w, b, p, 'relu', 'tf32')
op.pack()
dy = t.dev_rand(1, co, osp, 3, signed=True)
dx = DevTensor(1, sp[0], sp[1], sp[2], ci)
dw = torch.zeros_like(w)
fn = dict(fwd=op.fwd, dgrad=lambda: op.dgrad(dy, dx), wgrad=lambda: op.wgrad(dy, dw, None))[opname]
ms = t.time_ms(fn, reps)
fl = 2.0 * np.prod(sp) * co * ci * 8
print('%s %s: %.3f ms  %.1f TF/s' % (sys.argv[1:6], opname, ms, fl / ms / 1e9))
