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
from .ops import ConvOp, UpConvOp, PoolOp, MfpOp, Frag2DenseOp


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


def _pool_op(x, pool, tie_mode='first'):
    h = _lib.get_handle()
    pool = tuple(int(p) for p in pool)
    xd = DevTensor.from_numpy(x, h)
    yd = DevTensor(x.shape[0], x.shape[2] // pool[0], x.shape[3] // pool[1], x.shape[4] // pool[2], x.shape[1])
    return h, PoolOp(h, xd, yd, pool, tie_mode=tie_mode), xd, yd


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
