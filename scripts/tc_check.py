#!/usr/bin/env python
"""GPU-side check + micro-benchmark of the tcgen05 conv kernels against the CUDA-core
(exact fp32) kernels of the same library.  Usage: python scripts/tc_check.py [--big]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from elektronn2_b200 import _lib  # noqa: E402
from elektronn2_b200.devtensor import DevTensor  # noqa: E402
from elektronn2_b200.ops import ConvOp, UpConvOp  # noqa: E402


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def dev_rand(n, c, sp, seed, signed=False):
    t = DevTensor(n, sp[0], sp[1], sp[2], c)
    g = torch.Generator(device='cuda').manual_seed(seed)
    v = torch.rand(n, sp[0], sp[1], sp[2], t.desc.c_pitch, device='cuda', generator=g)
    if signed:
        v = v * 2 - 1
    v[..., c:] = 0
    t.buf.copy_(v.reshape(-1))
    return t


def view(t):
    d = t.desc
    return t.buf[t.offset:t.offset + d.floats].view(d.n, d.z, d.x, d.y, d.c_pitch)[..., :d.c]


def time_ms(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def check_conv(h, n, ci, sp, co, k, bench=False):
    osp = [s - f + 1 for s, f in zip(sp, k)]
    x = dev_rand(n, ci, sp, 1)
    g = torch.Generator(device='cuda').manual_seed(2)
    w = torch.randn(co, ci, *k, device='cuda', generator=g) * float(np.sqrt(2.0 / (ci * np.prod(k))))
    b = torch.randn(co, device='cuda', generator=g) * 0.1
    res = {}
    outs = {}
    for comp in ('f32', 'tf32'):
        y = DevTensor(n, osp[0], osp[1], osp[2], co)
        op = ConvOp(h, x, y, w, b, k, 'relu', comp)
        op.pack()
        op.fwd()
        dy = dev_rand(n, co, osp, 3, signed=True)
        dx = DevTensor(n, sp[0], sp[1], sp[2], ci)
        op.dgrad(dy, dx)
        dw = torch.zeros_like(w)
        db = torch.zeros_like(b)
        op.wgrad(dy, dw, db)
        # fused ReLU-backward gate + accumulate into a pre-filled gradient
        gate = dev_rand(n, ci, sp, 5, signed=True)
        dx2 = dev_rand(n, ci, sp, 6, signed=True)
        op.dgrad(dy, dx2, accumulate=True, relu_gate=gate)
        dx3 = dev_rand(n, ci, sp, 7, signed=True)
        op.dgrad(dy, dx3, accumulate=True)
        dx4 = DevTensor(n, sp[0], sp[1], sp[2], ci)
        op.dgrad(dy, dx4, relu_gate=gate)
        torch.cuda.synchronize()
        outs[comp] = (view(y).clone(), view(dx).clone(), dw.clone(), view(dx2).clone(), view(dx3).clone(), view(dx4).clone(),
                      db.clone())
        if bench:
            fl = 2.0 * n * np.prod(osp) * co * ci * np.prod(k)
            for name, fn in (('fwd', op.fwd), ('dgrad', lambda: op.dgrad(dy, dx)), ('wgrad', lambda: op.wgrad(dy, dw, None))):
                ms = time_ms(fn)
                res['%s_%s' % (comp, name)] = '%.3f ms %.1f TF/s' % (ms, fl / ms / 1e9)
    e = [rel(a, b_) for a, b_ in zip(outs['tf32'], outs['f32'])]
    print('conv n=%d ci=%d sp=%s co=%d k=%s : fwd %.2e dgrad %.2e wgrad %.2e gate+acc %.2e acc %.2e gate %.2e db %.1e %s %s' % (
        n, ci, sp, co, k, e[0], e[1], e[2], e[3], e[4], e[5], e[6], 'OK' if max(e) < 1e-3 else 'FAIL', res), flush=True)
    return max(e) < 1e-3


def check_upconv(h, n, ci, sp, co, p):
    osp = [s * q for s, q in zip(sp, p)]
    x = dev_rand(n, ci, sp, 1)
    g = torch.Generator(device='cuda').manual_seed(2)
    w = torch.randn(co, ci, *p, device='cuda', generator=g) * 0.2
    b = torch.randn(co, device='cuda', generator=g) * 0.1
    outs = {}
    for comp in ('f32', 'tf32'):
        y = DevTensor(n, osp[0], osp[1], osp[2], co)
        op = UpConvOp(h, x, y, w, b, p, 'relu', comp)
        op.pack()
        op.fwd()
        dy = dev_rand(n, co, osp, 3, signed=True)
        dx = DevTensor(n, sp[0], sp[1], sp[2], ci)
        op.dgrad(dy, dx)
        dw = torch.zeros_like(w)
        op.wgrad(dy, dw, None)
        torch.cuda.synchronize()
        outs[comp] = (view(y).clone(), view(dx).clone(), dw.clone())
        if comp == 'tf32':
            ms = [time_ms(f) for f in (op.fwd, lambda: op.dgrad(dy, dx), lambda: op.wgrad(dy, dw, None))]
    e = [rel(a, b_) for a, b_ in zip(outs['tf32'], outs['f32'])]
    print('upconv n=%d ci=%d sp=%s co=%d p=%s : fwd %.2e dgrad %.2e wgrad %.2e %s  [fwd %.3f dgrad %.3f wgrad %.3f ms]' % (
        n, ci, sp, co, p, e[0], e[1], e[2], 'OK' if max(e) < 1e-3 else 'FAIL', ms[0], ms[1], ms[2]), flush=True)
    return max(e) < 1e-3


def main():
    h = _lib.get_handle(0)
    ok = True
    cases = [(1, 32, (4, 10, 18), 64, (1, 1, 1)), (1, 32, (6, 10, 18), 64, (3, 3, 3)), (2, 20, (5, 12, 11), 40, (3, 3, 3)),
             (1, 40, (4, 9, 9), 150, (2, 4, 4)), (1, 64, (5, 9, 10), 64, (3, 3, 3)), (1, 128, (4, 6, 6), 256, (3, 3, 3)),
             (1, 200, (2, 6, 6), 200, (1, 1, 1)), (1, 256, (5, 7, 7), 512, (3, 3, 3)), (1, 768, (5, 6, 6), 256, (3, 3, 3))]
    # shapes that qualify for the halo-reuse (plane) kernel: ragged edges, several channel blocks,
    # even / asymmetric filters, more than one N tile
    cases += [(1, 32, (6, 30, 18), 48, (3, 3, 3)), (1, 20, (5, 30, 26), 40, (3, 3, 3)), (2, 64, (4, 18, 18), 128, (1, 3, 3)),
              (1, 40, (7, 19, 19), 80, (4, 4, 4)), (1, 64, (3, 18, 18), 300, (3, 3, 3)), (1, 100, (5, 27, 19), 36, (2, 4, 4)),
              (1, 32, (9, 34, 34), 64, (3, 3, 3))]
    # few tiles + long K: split-K plans of the z-stack kernel (partial tiles + reduce/epilogue kernel)
    cases += [(1, 256, (5, 16, 16), 128, (3, 3, 3)), (1, 200, (4, 15, 14), 72, (3, 3, 3)), (1, 768, (6, 18, 18), 256, (3, 3, 3)),
              (1, 384, (5, 28, 28), 128, (3, 3, 3))]
    for c in cases:
        ok &= check_conv(h, *c)
    for c in [(1, 42, (3, 4, 4), 45, (1, 4, 4)), (1, 64, (3, 4, 5), 64, (2, 2, 2)), (1, 512, (7, 9, 9), 512, (2, 2, 2)),
              (1, 256, (10, 14, 14), 256, (2, 2, 2)), (1, 128, (16, 24, 24), 128, (2, 2, 2))]:
        ok &= check_upconv(h, *c)
    if '--big' in sys.argv:
        for c in [(1, 32, (114, 130, 130), 64, (3, 3, 3)), (1, 64, (56, 64, 64), 64, (3, 3, 3)),
                  (1, 64, (54, 62, 62), 128, (3, 3, 3)), (1, 128, (24, 28, 28), 256, (3, 3, 3)),
                  (1, 192, (32, 48, 48), 64, (3, 3, 3)), (1, 768, (14, 18, 18), 256, (3, 3, 3))]:
            ok &= check_conv(h, *c, bench=True)
    print('ALL OK' if ok else 'SOME FAILED', flush=True)
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
