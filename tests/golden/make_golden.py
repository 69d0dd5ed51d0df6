#!/usr/bin/env python
"""Generate the committed golden fixtures in tests/golden/.

Run in the BUILD container only (needs /root/reference; the GPU box has no copy):
    python tests/golden/make_golden.py

Two kinds of fixture come out:

1. ``ref_python.json`` -- outputs of the reference's OWN code, imported from
   /root/reference with a stub ``theano`` module (Theano itself cannot be
   installed here): ``TaggedShape`` (neuromancer/graphutils.py:27-310),
   ``cnncalculator`` (utils/cnncalculator.py) and ``initweights``
   (neuromancer/variables.py:205-266).  They pin the shape algebra, the
   MFP-valid patch sizes and the weight-initialisation rule.
2. ``ops_small.npz`` -- seeded inputs and float64 oracle outputs for every op
   of the hot path at small sizes (oracle/ops.py, cross-checked against
   torch-CPU in tests/test_oracle.py).  These travel to the GPU box, where
   the CUDA path is compared against them as well as against the live oracle.
"""
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/elektronn2'
sys.path.insert(0, ROOT)


def _load_reference_pieces():
    if not hasattr(np, 'int'):
        np.int = int  # removed alias the reference still uses (graphutils.py:60-72)
    th = types.ModuleType('theano')
    th.config = types.SimpleNamespace(floatX='float32')
    th.shared = lambda *a, **k: None
    tt = types.ModuleType('theano.tensor')
    tt.__path__ = []
    tt.TensorConstant = tt.TensorType = object
    sv = types.ModuleType('theano.tensor.sharedvar')
    sv.TensorSharedVariable = object
    th.tensor = tt
    sys.modules.setdefault('theano', th)
    sys.modules.setdefault('theano.tensor', tt)
    sys.modules.setdefault('theano.tensor.sharedvar', sv)

    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod

    cc = load('ref_cnncalculator', os.path.join(REF, 'utils/cnncalculator.py'))
    # graphutils / variables use package-relative imports; exec their source with the
    # relative import lines dropped (nothing of theirs is needed for the pieces we call).
    def load_stripped(name, path, drop):
        src = open(path).read().split('\n')
        src = [l for l in src if not any(l.strip().startswith(d) for d in drop)]
        mod = types.ModuleType(name)
        mod.__dict__['theano'] = th
        mod.__dict__['floatX'] = 'float32'
        mod.__dict__['as_floatX'] = lambda v: np.asarray(v, np.float32)
        exec(compile('\n'.join(src), path, 'exec'), mod.__dict__)
        return mod

    gu = load_stripped('ref_graphutils', os.path.join(REF, 'neuromancer/graphutils.py'),
                       ['from ..', 'from .', 'from builtins'])
    va = load_stripped('ref_variables', os.path.join(REF, 'neuromancer/variables.py'),
                       ['from ..', 'from .', 'from builtins'])
    return cc, gu, va


NETS = {
    'neuro3d_lite': dict(filters=[[1, 4, 4], [3, 3, 3], [2, 4, 4], [1, 3, 3], [1, 3, 3], [1, 1, 1], [1, 1, 1]],
                         pools=[[1, 2, 2], [1, 2, 2], [2, 1, 1], [1, 1, 1], [1, 1, 1], [1, 1, 1], [1, 1, 1]],
                         patch=[11, 155, 155]),
    'neuro3d': dict(filters=[[1, 6, 6], [1, 5, 5], [1, 5, 5], [4, 4, 4], [3, 4, 4], [3, 4, 4], [2, 4, 4],
                             [1, 4, 4], [1, 4, 4], [1, 1, 1], [1, 1, 1]],
                    pools=[[1, 2, 2], [1, 2, 2], [1, 1, 1], [2, 1, 1]] + [[1, 1, 1]] * 7,
                    patch=[23, 185, 185]),
}


def main():
    cc, gu, va = _load_reference_pieces()
    out = {}

    # --- cnncalculator -------------------------------------------------------
    f = [[1, 6, 6], [4, 4, 4], [2, 2, 2], [1, 1, 1]]
    p = [[1, 2, 2], [2, 2, 2], [2, 2, 2], [1, 1, 1]]
    d = cc.cnncalculator(f, p, [8, 211, 211], mfp=[True, True, False, False], force_center=True, ndim=3)
    out['cnncalc_docstring'] = dict(patch_size=[int(v) for v in d.patch_size],
                                    pred_stride=[int(v) for v in d.pred_stride],
                                    offset=[float(v) for v in d.offset],
                                    pool_out=[[int(v) for v in l] for l in d.pool_out],
                                    fields=[[int(v) for v in l] for l in d.fields])
    for name, spec in NETS.items():
        for mfp in (False, True):
            d = cc.cnncalculator(spec['filters'], spec['pools'], spec['patch'], mfp=[mfp] * len(spec['filters']),
                                 ndim=3)
            out['cnncalc_%s_mfp%d' % (name, mfp)] = dict(
                patch_size=[int(v) for v in d.patch_size],
                pred_stride=[int(v) for v in d.pred_stride],
                offset=[float(v) for v in d.offset],
                pool_out=[[int(v) for v in l] for l in d.pool_out],
                fields=[[int(v) for v in l] for l in d.fields],
                valid_head=[[int(v) for v in l[:12]] for l in d.valid_patch_sizes])

    # --- TaggedShape -----------------------------------------------------------
    TS = gu.TaggedShape
    sh = TS([None, 1, 23, 185, 185], 'b,f,z,x,y')
    s2 = sh.updateshape('f', 20).updateshape(3, 90).updatefov(1, 7).updatestrides(np.array([1, 2, 2]))
    s3 = s2.updateshape('b', 4, mode='mult').updateshape('z', 2, mode='mult')
    out['taggedshape'] = dict(
        spatial_axes=sh.spatial_axes, ndim=sh.ndim, spatial_shape=sh.spatial_shape,
        s2_shape=s2.shape, s2_fov=[int(v) for v in s2.fov], s2_offsets=[int(v) for v in s2.offsets],
        s2_strides=[int(v) for v in s2.strides], s2_repr=repr(s2),
        s3_shape=s3.shape, stripnone_prod=int(s2.stripnone_prod), spatial_size=int(s2.spatial_size),
        mfp_offsets=np.asarray(sh.mfp_offsets).tolist(), f_index=sh.tag2index('f'))

    # --- initweights -------------------------------------------------------------
    iw = {}
    np.random.seed(2)
    w = va.initweights((20, 1, 1, 4, 4), scale='glorot', mode='normal', pool=(1, 2, 2), spatial_axes=[2, 3, 4])
    iw['conv_w'] = dict(shape=list(w.shape), std=float(w.std()), head=[float(v) for v in w.ravel()[:8]],
                        dtype=str(w.dtype))
    b = va.initweights((20,), scale=1.0 / 16, mode='const')
    iw['relu_b'] = [float(v) for v in b[:3]]
    b = va.initweights((2,), scale=1e-6, mode='fix-uni')
    iw['lin_b'] = [float(v) for v in b]
    w = va.initweights((45, 42, 1, 4, 4), scale='glorot', mode='normal', pool=(1, 4, 4), spatial_axes=[2, 3, 4])
    iw['upconv_w'] = dict(std=float(w.std()), head=[float(v) for v in w.ravel()[:4]])
    out['initweights_seed2'] = iw

    with open(os.path.join(HERE, 'ref_python.json'), 'w') as fh:
        json.dump(out, fh, indent=1, sort_keys=True)

    # --- oracle vectors for the GPU box ------------------------------------------
    from oracle import ops
    rng = np.random.RandomState(1234)
    g = {}
    x = rng.rand(2, 5, 6, 10, 9).astype(np.float32)
    w = (rng.randn(7, 5, 2, 3, 4) * 0.2).astype(np.float32)
    b = rng.randn(7).astype(np.float32) * 0.1
    g['conv_x'], g['conv_w'], g['conv_b'] = x, w, b
    y, (lin, pre), _ = ops.conv_node_fwd(x, w, b, (1, 2, 2), 'relu')
    g['conv_lin'], g['conv_y_pool122_relu'] = lin, y
    dy = rng.randn(*lin.shape).astype(np.float32)
    g['conv_dy'] = dy
    g['conv_dx'] = ops.conv3d_dgrad(dy, w, x.shape)
    g['conv_dw'] = ops.conv3d_wgrad(dy, x, w.shape)
    g['conv_db'] = ops.bias_grad(dy)
    xu = rng.rand(1, 6, 3, 4, 5).astype(np.float32)
    wu = (rng.randn(4, 6, 1, 2, 2) * 0.3).astype(np.float32)
    g['up_x'], g['up_w'] = xu, wu
    g['up_y'] = ops.upconv3d(xu, wu, (1, 2, 2))
    dyu = rng.randn(*g['up_y'].shape).astype(np.float32)
    g['up_dy'] = dyu
    g['up_dx'] = ops.upconv3d_dgrad(dyu, wu, (1, 2, 2))
    g['up_dw'] = ops.upconv3d_wgrad(dyu, xu, (1, 2, 2))
    xp = rng.rand(2, 3, 4, 6, 8).astype(np.float32)
    xp[0, 0, :2, :2, :2] = 0.5  # a tie block
    g['pool_x'] = xp
    g['pool_y222'] = ops.pooling(xp, (2, 2, 2))
    g['pool_idx222'] = ops.pooling_argmax(xp, (2, 2, 2))
    dyp = rng.randn(*g['pool_y222'].shape).astype(np.float32)
    g['pool_dy'] = dyp
    g['pool_dx_first'] = ops.pooling_bwd(dyp, xp, (2, 2, 2), 'first')
    g['pool_dx_all'] = ops.pooling_bwd(dyp, xp, (2, 2, 2), 'all')
    xm = rng.rand(1, 3, 7, 9, 11).astype(np.float32)
    fr, off, st = ops.fragmentpool(xm, (2, 2, 2), [[0, 0, 0]], [1, 1, 1])
    fr2, off2, st2 = ops.fragmentpool(fr[:, :, :, :4, :5][:, :, :, :3, :3], (1, 2, 2), off, st)
    g['mfp_x'], g['mfp_y'], g['mfp_off'], g['mfp_st'] = xm, fr, off, st
    g['mfp2_y'], g['mfp2_off'], g['mfp2_st'] = fr2, off2, st2
    g['f2d_y'] = ops.fragments2dense(fr, off, st)
    np.savez_compressed(os.path.join(HERE, 'ops_small.npz'), **g)
    print('wrote', os.path.join(HERE, 'ref_python.json'), 'and ops_small.npz')


if __name__ == '__main__':
    main()
