"""numpy in / numpy out calls of the hot-path ops through the C ABI.

This is the granularity of the reference's ``computations.py`` functions (conv / upconv / pooling / fragmentpool /
fragments2dense) and of a Theano ``Op.perform``: arrays in the reference's layouts -- activations ``(b, f, z, x, y)``,
conv weights ``(f_out, f_in, kz, kx, ky)`` -- cross PCIe on every call.  ``theano_ops.py`` wraps these functions as
Theano Ops; the fast face is ``neuromancer.executor`` (everything resident in HBM, CUDA graphs).  Nothing here falls
back to the CPU: without the library / a GPU the first call raises.
"""
import numpy as np
import torch

from . import _lib
from .devtensor import DevTensor
from .ops import ConvOp, UpConvOp, PoolOp, MfpOp, Frag2DenseOp, LossOp


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).cuda()


def _f32(a, ndim=5):
    a = np.ascontiguousarray(a, dtype=np.float32)
    if a.ndim != ndim:
        raise ValueError("expected a %d-d array, got shape %s" % (ndim, a.shape))
    return a


def _conv_op(x_shape, w, b, act, compute):
    h = _lib.get_handle()
    n, ci = x_shape[0], x_shape[1]
    k = tuple(w.shape[2:])
    osp = [x_shape[2 + i] - k[i] + 1 for i in range(3)]
    xd = DevTensor(n, x_shape[2], x_shape[3], x_shape[4], ci)
    yd = DevTensor(n, osp[0], osp[1], osp[2], w.shape[0])
    op = ConvOp(h, xd, yd, _t(w), None if b is None else _t(b), k, act, compute)
    op.pack()
    return h, op, xd, yd


def conv3d(x, w, b=None, act='lin', compute='tf32'):
    """computations.conv, 3-D 'valid' true convolution (computations.py:364-428) [+ bias + activation]."""
    x, w = _f32(x), _f32(w)
    if x.shape[1] != w.shape[1]:
        raise ValueError("conv3d: x has %d channels, w expects %d" % (x.shape[1], w.shape[1]))
    h, op, xd, yd = _conv_op(x.shape, w, b, act, compute)
    DevTensor.from_numpy(x, h, out=xd)
    op.fwd()
    return yd.numpy(h)


def conv3d_grad_input(dy, w, x_shape, compute='tf32'):
    """d/dx of conv3d (full correlation of dy with w): what T.grad asks of the conv op for its first input."""
    dy, w = _f32(dy), _f32(w)
    h, op, xd, yd = _conv_op(tuple(int(v) for v in x_shape), w, None, 'lin', compute)
    DevTensor.from_numpy(dy, h, out=yd)
    op.dgrad(yd, xd)
    return xd.numpy(h)


def conv3d_grad_weights(x, dy, w_shape, compute='tf32'):
    """d/dw of conv3d, in the reference's weight layout."""
    x, dy = _f32(x), _f32(dy)
    w_shape = tuple(int(v) for v in w_shape)
    h, op, xd, yd = _conv_op(x.shape, np.zeros(w_shape, np.float32), None, 'lin', compute)
    DevTensor.from_numpy(x, h, out=xd)
    DevTensor.from_numpy(dy, h, out=yd)
    dw = torch.zeros(w_shape, dtype=torch.float32, device='cuda')
    op.wgrad(yd, dw, None)
    return dw.cpu().numpy()


def _upconv_op(x_shape, w, pool, compute):
    h = _lib.get_handle()
    pool = tuple(int(p) for p in pool)
    n, ci = x_shape[0], x_shape[1]
    xd = DevTensor(n, x_shape[2], x_shape[3], x_shape[4], ci)
    yd = DevTensor(n, x_shape[2] * pool[0], x_shape[3] * pool[1], x_shape[4] * pool[2], w.shape[0])
    op = UpConvOp(h, xd, yd, _t(w), None, pool, 'lin', compute)
    op.pack()
    return h, op, xd, yd


def upconv3d(x, w, pool, compute='tf32'):
    """computations.upconv / unpooling + conv (computations.py:216-255, 749-756; neural.py:1013-1030):
    y[b,o,z*pz+i,x*px+j,y*py+k] = sum_c x[b,c,z,x,y] * w[o,c,i,j,k]."""
    x, w = _f32(x), _f32(w)
    h, op, xd, yd = _upconv_op(x.shape, w, pool, compute)
    DevTensor.from_numpy(x, h, out=xd)
    op.fwd()
    return yd.numpy(h)


def upconv3d_grad_input(dy, w, x_shape, pool, compute='tf32'):
    dy, w = _f32(dy), _f32(w)
    h, op, xd, yd = _upconv_op(tuple(int(v) for v in x_shape), w, pool, compute)
    DevTensor.from_numpy(dy, h, out=yd)
    op.dgrad(yd, xd)
    return xd.numpy(h)


def upconv3d_grad_weights(x, dy, w_shape, pool, compute='tf32'):
    x, dy = _f32(x), _f32(dy)
    w_shape = tuple(int(v) for v in w_shape)
    h, op, xd, yd = _upconv_op(x.shape, np.zeros(w_shape, np.float32), pool, compute)
    DevTensor.from_numpy(x, h, out=xd)
    DevTensor.from_numpy(dy, h, out=yd)
    dw = torch.zeros(w_shape, dtype=torch.float32, device='cuda')
    op.wgrad(yd, dw, None)
    return dw.cpu().numpy()


def _pool_op(x, pool, tie_mode='first', mode='max'):
    h = _lib.get_handle()
    pool = tuple(int(p) for p in pool)
    xd = DevTensor.from_numpy(x, h)
    yd = DevTensor(x.shape[0], x.shape[2] // pool[0], x.shape[3] // pool[1], x.shape[4] // pool[2], x.shape[1])
    return h, PoolOp(h, xd, yd, pool, tie_mode=tie_mode, mode=mode), xd, yd


def pool3d(x, pool, mode='average_inc_pad'):
    """computations.pooling with dnn_pool's other modes (computations.py:556-561, 589-590, 600):
    'average_inc_pad' / 'average_exc_pad' (identical without padding) and 'sum'; stride == pool."""
    x = _f32(x)
    h, op, xd, yd = _pool_op(x, pool, mode=mode)
    op.fwd()
    return yd.numpy(h)


def pool3d_grad(x_shape, dy, pool, mode='average_inc_pad'):
    dy = _f32(dy)
    h, op, xd, yd = _pool_op(np.zeros(tuple(int(v) for v in x_shape), np.float32), pool, mode=mode)
    dyd = DevTensor.from_numpy(dy, h)
    dxd = xd.like()
    op.bwd(dyd, dxd)
    return dxd.numpy(h)


def maxout(x, factor=2, axis=1):
    """computations.maxout (computations.py:455-495): max over groups of ``factor`` consecutive entries of ``axis``."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    factor, axis = int(factor), int(axis)
    if x.shape[axis] % factor:
        raise ValueError("maxout: axis %d of length %d is not divisible by %d" % (axis, x.shape[axis], factor))
    h = _lib.get_handle()
    outer = int(np.prod(x.shape[:axis], dtype=np.int64))
    inner = int(np.prod(x.shape[axis + 1:], dtype=np.int64))
    fo = x.shape[axis] // factor
    xd = _t(x)
    yd = torch.empty(x.shape[:axis] + (fo,) + x.shape[axis + 1:], dtype=torch.float32, device='cuda')
    h.call('e2_maxout_fwd', _lib.ptr(xd), _lib.ptr(yd), outer, fo, inner, factor, h.stream())
    return yd.cpu().numpy()


def maxout_grad(x, dy, factor=2, axis=1):
    x = np.ascontiguousarray(x, dtype=np.float32)
    dy = np.ascontiguousarray(dy, dtype=np.float32)
    factor, axis = int(factor), int(axis)
    h = _lib.get_handle()
    outer = int(np.prod(x.shape[:axis], dtype=np.int64))
    inner = int(np.prod(x.shape[axis + 1:], dtype=np.int64))
    fo = x.shape[axis] // factor
    xd, dyd = _t(x), _t(dy)
    dxd = torch.empty_like(xd)
    h.call('e2_maxout_bwd', _lib.ptr(xd), _lib.ptr(dyd), _lib.ptr(dxd), outer, fo, inner, factor, h.stream())
    return dxd.cpu().numpy()


def affine_act(v, act='lin', b=None, gamma=None, mean=None, std=None, alpha=None):
    """y = act((gamma / std) * v + b - gamma * mean / std) (neural.py:711-712) on a (b,f,z,x,y) array; 'prelu' takes
    its slope in ``alpha`` (= b[:,1] of the reference's (f,2) bias, neural.py:655-657)."""
    from .ops import AffineActOp
    v = _f32(v)
    h = _lib.get_handle()
    vd = DevTensor.from_numpy(v, h)
    yd = vd.like()
    c = v.shape[1]
    bb = np.zeros(c, np.float32) if b is None else np.asarray(b, np.float32)
    if act == 'prelu':
        bb = np.stack([bb, np.asarray(alpha, np.float32)], 1)
    bn = 'predict' if (gamma is not None or mean is not None or std is not None) else False
    one, zero = np.ones(c, np.float32), np.zeros(c, np.float32)
    op = AffineActOp(h, vd, yd, act, _t(bb), bn, _t(one if gamma is None else gamma) if bn else None,
                     _t(zero if mean is None else mean) if bn else None, _t(one if std is None else std) if bn else None)
    op.fwd()
    return yd.numpy(h)


def maxpool3d(x, pool, return_argmax=False):
    """computations.pooling, 3-D 'max' (computations.py:538-649); argmax = first maximum in (z,x,y) scan order."""
    x = _f32(x)
    h, op, xd, yd = _pool_op(x, pool)
    op.fwd()
    y = yd.numpy(h)
    return (y, op.argmax.int_numpy()) if return_argmax else y


def maxpool3d_grad(x, dy, pool, tie_mode='first'):
    """Gradient of maxpool3d w.r.t. x; tie_mode 'first' (cuDNN-like) or 'all' (Theano CPU: every tied element)."""
    x, dy = _f32(x), _f32(dy)
    h, op, xd, yd = _pool_op(x, pool, tie_mode)
    op.fwd()
    dyd = DevTensor.from_numpy(dy, h)
    dxd = xd.like()
    op.bwd(dyd, dxd)
    return dxd.numpy(h)


def fragmentpool(x, pool, return_argmax=False):
    """computations.fragmentpool (computations.py:652-678): all prod(pool) offset fragments in one pass, batch order
    new-offset-major / old-fragment-minor.  The offsets / strides bookkeeping stays with the caller (:668-676)."""
    x = _f32(x)
    h = _lib.get_handle()
    pool = tuple(int(p) for p in pool)
    xd = DevTensor.from_numpy(x, h)
    osp = [(x.shape[2 + i] - pool[i] + 1) // pool[i] for i in range(3)]
    yd = DevTensor(x.shape[0] * int(np.prod(pool)), osp[0], osp[1], osp[2], x.shape[1])
    op = MfpOp(h, xd, yd, pool)
    op.fwd()
    y = yd.numpy(h)
    return (y, op.argmax.int_numpy()) if return_argmax else y


def fragments2dense(fragments, offsets, strides):
    """computations.fragments2dense (computations.py:681-701)."""
    f = _f32(fragments)
    h = _lib.get_handle()
    st = tuple(int(s) for s in strides)
    fd = DevTensor.from_numpy(f, h)
    dd = DevTensor(1, f.shape[2] * st[0], f.shape[3] * st[1], f.shape[4] * st[2], f.shape[1])
    Frag2DenseOp(h, fd, dd, offsets, st).fwd()
    return dd.numpy(h)


def softmax(x):
    """computations.softmax over the feature axis of a (b, f, z, x, y) array (computations.py:137-177): the Softmax
    node alone, i.e. the fused loss head called without a target (include/e2b200.h, e2_softmax_nll_fwd)."""
    x = _f32(x)
    h = _lib.get_handle()
    xd = DevTensor.from_numpy(x, h)
    pd = xd.like()
    LossOp(h, xd, None, pd).fwd()
    return pd.numpy(h)
