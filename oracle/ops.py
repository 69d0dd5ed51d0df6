"""Numpy float64 restatement of the op layer (reference: neuromancer/computations.py).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Layout everywhere is the
reference's: activations ``(b, f, z, x, y)``, conv weights
``(f_out, f_in, kz, kx, ky)``, bias ``(f_out,)``.
"""
from itertools import product

import numpy as np

F64 = np.float64


# --------------------------------------------------------------------------- conv
def flip_w(w):
    """Kernel flipped on the three spatial axes.

    The reference convolves (true convolution, not cross-correlation):
    cuDNN default conv_mode ('conv') at computations.py:391 and
    conv3d2d at :406/:427; pinned by tests/test_conv.py:89-104 (== np.convolve)
    and :117-129 (cuDNN == conv3d2d).
    """
    return w[:, :, ::-1, ::-1, ::-1]


def conv3d(x, w):
    """computations.conv, 3-D 'valid' branch (computations.py:364-428).

    y[b,o,z,x,y] = sum_{c,i,j,k} x[b,c,z+kz-1-i, x+kx-1-j, y+ky-1-k] * w[o,c,i,j,k]
    Direct form: one (positions x c) @ (c x o) product per filter tap.
    """
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    b, c, Z, X, Y = x.shape
    o, c2, kz, kx, ky = w.shape
    assert c == c2, (x.shape, w.shape)
    Zo, Xo, Yo = Z - kz + 1, X - kx + 1, Y - ky + 1
    assert min(Zo, Xo, Yo) >= 1
    wf = flip_w(w)
    y = np.zeros((b, o, Zo, Xo, Yo), F64)
    for i, j, k in product(range(kz), range(kx), range(ky)):
        xs = x[:, :, i:i + Zo, j:j + Xo, k:k + Yo]
        y += np.einsum('bczxy,oc->bozxy', xs, wf[:, :, i, j, k], optimize=True)
    return y


def conv3d_dot(x, w):
    """1x1x1 shortcut: tensordot over the f axis (computations.py:330-335,
    377-384 -> dot :179-213)."""
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    assert w.shape[2:] == (1, 1, 1)
    wm = w[:, :, 0, 0, 0].T  # (f_in, f_out)
    y = np.tensordot(x, wm, axes=[1, 0])  # (b,z,x,y,o)
    return np.moveaxis(y, -1, 1)


def _conv2d_valid(x, w):
    """2-D true convolution, 'valid' (what Theano's conv2d/CorrMM computes with
    filter_flip=True)."""
    n, c, X, Y = x.shape
    o, _, kx, ky = w.shape
    Xo, Yo = X - kx + 1, Y - ky + 1
    wf = w[:, :, ::-1, ::-1]
    y = np.zeros((n, o, Xo, Yo), F64)
    for j, k in product(range(kx), range(ky)):
        y += np.einsum('ncxy,oc->noxy', x[:, :, j:j + Xo, k:k + Yo], wf[:, :, j, k],
                       optimize=True)
    return y


def conv3d_theano_shaped(x, w):
    """The conv3d2d decomposition the CPU reference path runs
    (computations.py:407-428): batched 2-D convs over (b*Z) images with
    (f_out*kz) filters, then the diagonal sum over (z, kz)."""
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    b, c, Z, X, Y = x.shape
    o, _, kz, kx, ky = w.shape
    Zo = Z - kz + 1
    x2 = np.moveaxis(x, 2, 1).reshape(b * Z, c, X, Y)
    w2 = np.moveaxis(w, 2, 1).reshape(o * kz, c, kx, ky)
    y2 = _conv2d_valid(x2, w2)  # (b*Z, o*kz, Xo, Yo)
    Xo, Yo = y2.shape[2:]
    y2 = y2.reshape(b, Z, o, kz, Xo, Yo)
    y = np.zeros((b, o, Zo, Xo, Yo), F64)
    for i in range(kz):  # true convolution along z: x index z + kz-1-i pairs with w tap i
        y += np.moveaxis(y2[:, kz - 1 - i:kz - 1 - i + Zo, :, i], 1, 2)
    return y


def conv3d_dgrad(dy, w, x_shape):
    """d(loss)/dx of conv3d (Theano derives it with T.grad, model.py:182)."""
    dy = np.asarray(dy, F64)
    w = np.asarray(w, F64)
    o, c, kz, kx, ky = w.shape
    _, _, Zo, Xo, Yo = dy.shape
    wf = flip_w(w)
    dx = np.zeros(x_shape, F64)
    for i, j, k in product(range(kz), range(kx), range(ky)):
        dx[:, :, i:i + Zo, j:j + Xo, k:k + Yo] += np.einsum(
            'bozxy,oc->bczxy', dy, wf[:, :, i, j, k], optimize=True)
    return dx


def conv3d_wgrad(dy, x, w_shape):
    """d(loss)/dw and d(loss)/db of conv3d followed by a per-feature bias."""
    dy = np.asarray(dy, F64)
    x = np.asarray(x, F64)
    o, c, kz, kx, ky = w_shape
    _, _, Zo, Xo, Yo = dy.shape
    dwf = np.zeros(w_shape, F64)
    for i, j, k in product(range(kz), range(kx), range(ky)):
        dwf[:, :, i, j, k] = np.einsum('bozxy,bczxy->oc', dy,
                                       x[:, :, i:i + Zo, j:j + Xo, k:k + Yo], optimize=True)
    return np.ascontiguousarray(flip_w(dwf))


def bias_grad(dy):
    return np.asarray(dy, F64).sum(axis=(0, 2, 3, 4))


# ------------------------------------------------------------------------- upconv
def unpooling(x, pool):
    """computations.unpooling_nd (computations.py:749-756): zeros of size
    S*p + p-1 with x at [p-1 :: p]."""
    x = np.asarray(x, F64)
    b, c = x.shape[:2]
    sp = x.shape[2:]
    new = [s * p + (p - 1) for s, p in zip(sp, pool)]
    out = np.zeros((b, c, *new), F64)
    sl = tuple(slice(p - 1, s * p, p) for s, p in zip(sp, pool))
    out[(slice(None), slice(None)) + sl] = x
    return out


def upconv3d_theano_shaped(x, w, pool):
    """UpConv CPU path (neural.py:1013-1020): unpool, then valid true conv with
    the node's w (f_out, f_in, p...)."""
    return conv3d(unpooling(x, pool), w)


def upconv3d(x, w, pool):
    """Both reference paths (unpool+conv; cuDNN GradI with conv_mode='cross',
    computations.py:245-253) reduce to the non-overlapping form
      y[b,o,z*pz+i,x*px+j,y*py+k] = sum_c x[b,c,z,x,y] * w[o,c,i,j,k]
    (kernel == pool == stride, neural.py:968)."""
    x = np.asarray(x, F64)
    w = np.asarray(w, F64)
    b, c, Z, X, Y = x.shape
    o, c2, pz, px, py = w.shape
    assert c == c2 and (pz, px, py) == tuple(pool)
    y = np.einsum('bczxy,ocijk->bozixjyk', x, w, optimize=True)
    return y.reshape(b, o, Z * pz, X * px, Y * py)


def upconv3d_dgrad(dy, w, pool):
    dy = np.asarray(dy, F64)
    w = np.asarray(w, F64)
    o, c, pz, px, py = w.shape
    b, _, Zu, Xu, Yu = dy.shape
    d = dy.reshape(b, o, Zu // pz, pz, Xu // px, px, Yu // py, py)
    return np.einsum('bozixjyk,ocijk->bczxy', d, w, optimize=True)


def upconv3d_wgrad(dy, x, pool):
    dy = np.asarray(dy, F64)
    x = np.asarray(x, F64)
    pz, px, py = pool
    b, o, Zu, Xu, Yu = dy.shape
    d = dy.reshape(b, o, Zu // pz, pz, Xu // px, px, Yu // py, py)
    return np.einsum('bozixjyk,bczxy->ocijk', d, x, optimize=True)


# ------------------------------------------------------------------------ pooling
def pooling(x, pool):
    """computations.pooling, 3-D max, stride == pool, ignore_border
    (computations.py:569-570 short-circuit; direct window form)."""
    x = np.asarray(x)
    pool = tuple(int(p) for p in pool)
    if all(p == 1 for p in pool):
        return x
    b, c, Z, X, Y = x.shape
    pz, px, py = pool
    Zo, Xo, Yo = Z // pz, X // px, Y // py
    v = x[:, :, :Zo * pz, :Xo * px, :Yo * py].reshape(b, c, Zo, pz, Xo, px, Yo, py)
    return v.max(axis=(3, 5, 7))


def pooling_theano_shaped(x, pool):
    """The CPU branch (computations.py:611-631): pool_2d over (x,y) then
    T.maximum over the pz strided z-slices."""
    x = np.asarray(x)
    pz, px, py = pool
    b, c, Z, X, Y = x.shape
    Xo, Yo = X // px, Y // py
    y = x[:, :, :, :Xo * px, :Yo * py].reshape(b, c, Z, Xo, px, Yo, py).max(axis=(4, 6))
    m = y[:, :, 0::pz]
    for z in range(1, pz):
        t = y[:, :, z::pz]
        n = min(m.shape[2], t.shape[2])
        m = np.maximum(t[:, :, :n], m[:, :, :n])
    return m[:, :, :Z // pz]


def pooling_argmax(x, pool):
    """Argmax definition used by the B200 path: the FIRST
    maximum in (z,x,y) row-major scan order of the window, reported as the int32
    linear index z*X*Y + x*Y + y into the (Z,X,Y) volume of that (b,f)."""
    x = np.asarray(x)
    b, c, Z, X, Y = x.shape
    pz, px, py = pool
    Zo, Xo, Yo = Z // pz, X // px, Y // py
    v = x[:, :, :Zo * pz, :Xo * px, :Yo * py].reshape(b, c, Zo, pz, Xo, px, Yo, py)
    v = v.transpose(0, 1, 2, 4, 6, 3, 5, 7).reshape(b, c, Zo, Xo, Yo, pz * px * py)
    a = v.argmax(axis=-1)  # numpy argmax returns the first occurrence
    dz, r = np.divmod(a, px * py)
    dx, dy_ = np.divmod(r, py)
    zz = np.arange(Zo)[:, None, None] * pz + dz
    xx = np.arange(Xo)[None, :, None] * px + dx
    yy = np.arange(Yo)[None, None, :] * py + dy_
    return (zz * (X * Y) + xx * Y + yy).astype(np.int32)


def pooling_bwd(dy, x, pool, tie_mode='first'):
    """Backward of max pooling.

    tie_mode='all'  : Theano-CPU semantics (MaxPoolGrad + grad of T.maximum):
                      every element equal to the window max receives dy.
    tie_mode='first': only the first maximum (scan order) receives dy -- the
                      single-winner rule cuDNN (the reference's GPU path) uses.
    """
    x = np.asarray(x)
    dy = np.asarray(dy, F64)
    b, c, Z, X, Y = x.shape
    pz, px, py = pool
    Zo, Xo, Yo = Z // pz, X // px, Y // py
    dx = np.zeros(x.shape, F64)
    if tie_mode == 'all':
        y = pooling(x, pool)
        up = np.repeat(np.repeat(np.repeat(y, pz, 2), px, 3), py, 4)
        dup = np.repeat(np.repeat(np.repeat(dy, pz, 2), px, 3), py, 4)
        core = (slice(None), slice(None), slice(0, Zo * pz), slice(0, Xo * px), slice(0, Yo * py))
        dx[core] = np.where(x[core] == up, dup, 0.0)
        return dx
    idx = pooling_argmax(x, pool).astype(np.int64)
    flat = dx.reshape(b, c, -1)
    np.put_along_axis(flat, idx.reshape(b, c, -1), dy.reshape(b, c, -1), axis=2)
    return dx


# ---------------------------------------------------------------------------- MFP
def fragmentpool(x, pool, offsets, strides):
    """computations.fragmentpool (computations.py:652-678).

    For every offset ix in product(range(pz),range(px),range(py)) (last axis
    fastest) pool the slice x[..., ix_d : ix_d + S_d - p_d + 1]; concatenate on
    the batch axis (new-offset-major, old-fragment-minor, :674); offsets_new /
    strides_new as at :668-676.
    """
    x = np.asarray(x)
    pool = tuple(int(p) for p in pool)
    offsets = np.atleast_2d(np.array(offsets, np.int64))
    strides = np.array(strides, np.int64)
    if all(p == 1 for p in pool):
        return x, offsets, strides
    sp = x.shape[2:]
    result, offsets_new = [], []
    for ix in product(*[range(p) for p in pool]):
        sl = tuple(slice(i, i + s - p + 1) for i, s, p in zip(ix, sp, pool))
        result.append(pooling(x[(slice(None), slice(None)) + sl], pool))
        for p_ in offsets:
            offsets_new.append(p_ + np.multiply(ix, strides))
    return (np.concatenate(result, axis=0), np.array(offsets_new, np.int64),
            np.multiply(pool, strides))


def fragmentpool_bwd(dy, x, pool, tie_mode='first'):
    x = np.asarray(x)
    pool = tuple(int(p) for p in pool)
    b = x.shape[0]
    sp = x.shape[2:]
    dx = np.zeros(x.shape, F64)
    for n, ix in enumerate(product(*[range(p) for p in pool])):
        sl = (slice(None), slice(None)) + tuple(slice(i, i + s - p + 1) for i, s, p in zip(ix, sp, pool))
        dx[sl] += pooling_bwd(dy[n * b:(n + 1) * b], x[sl], pool, tie_mode)
    return dx


def fragments2dense(frag, offsets, strides):
    """computations.fragments2dense (computations.py:681-701):
    out[0,:, off_k[0]::s0, off_k[1]::s1, off_k[2]::s2] = frag[k]."""
    frag = np.asarray(frag)
    strides = [int(s) for s in strides]
    n = int(np.prod(strides))
    assert frag.shape[0] == n == len(offsets)
    out_sh = [1, frag.shape[1]] + [s * st for s, st in zip(frag.shape[2:], strides)]
    out = np.zeros(out_sh, frag.dtype)
    for i, off in enumerate(offsets):
        out[0, :, off[0]::strides[0], off[1]::strides[1], off[2]::strides[2]] = frag[i]
    return out


def fragments2dense_bwd(dout, offsets, strides):
    strides = [int(s) for s in strides]
    return np.stack([dout[0, :, off[0]::strides[0], off[1]::strides[1], off[2]::strides[2]]
                     for off in offsets], axis=0)


# ----------------------------------------------------------------- crop / concat
def crop(x, c):
    """Crop node (neural.py:1152-1168): x[..., off : S-off] per spatial axis."""
    sl = tuple(slice(o, s - o) for o, s in zip(c, x.shape[2:]))
    return x[(slice(None), slice(None)) + sl]


def crop_bwd(dy, c, x_shape):
    dx = np.zeros(x_shape, F64)
    sl = tuple(slice(o, s - o) for o, s in zip(c, x_shape[2:]))
    dx[(slice(None), slice(None)) + sl] = dy
    return dx


def concat_f(parts):
    """Concat(axis='f') (node_basic.py:1403-1451); AutoMerge order is
    (lo_res/upconv first, hi_res/skip second), neural.py:1399."""
    return np.concatenate(parts, axis=1)


# --------------------------------------------------------------------- epilogue
SELU_ALPHA = 1.6732632423543772848170429916717     # computations.py:96-97
SELU_SCALE = 1.0507009873554804934193349852946


def activation(x, name):
    """apply_activation (computations.py:57-134), the parameter-free ones."""
    x = np.asarray(x, F64)
    if name == 'relu':
        return np.maximum(x, 0.0)  # T.nnet.relu
    if name in ('lin', 'linear'):
        return x
    if name == 'tanh':
        return np.tanh(x)
    if name in ('sig', 'logistic', 'sigmoid'):
        return 1.0 / (1.0 + np.exp(-x))
    if name == 'abs':
        return np.abs(x)
    if name == 'soft+':
        return np.logaddexp(0.0, x)
    if name == 'elu':
        return np.where(x > 0, x, np.expm1(np.minimum(x, 0)))
    if name == 'selu':   # computations.py:89-98: scale * T.nnet.elu(x, alpha)
        return SELU_SCALE * np.where(x > 0, x, SELU_ALPHA * np.expm1(np.minimum(x, 0)))
    raise NotImplementedError(name)


def activation_bwd(dy, pre, name):
    """Derivative w.r.t. the pre-activation ``pre``.  relu'(0) = 0 (Theano's
    relu is 0.5*(x+|x|) whose grad at exactly 0 is 0.5*(1+sgn(0)) = 0.5; exact
    zeros of a biased float conv output do not occur with continuous data, see
    DESIGN.md)."""
    pre = np.asarray(pre, F64)
    if name == 'relu':
        return dy * (pre > 0)
    if name in ('lin', 'linear'):
        return dy
    if name == 'tanh':
        return dy * (1 - np.tanh(pre) ** 2)
    if name in ('sig', 'logistic', 'sigmoid'):
        s = 1.0 / (1.0 + np.exp(-pre))
        return dy * s * (1 - s)
    if name == 'abs':
        return dy * np.sign(pre)
    if name == 'soft+':
        return dy / (1.0 + np.exp(-pre))
    if name == 'elu':
        return dy * np.where(pre > 0, 1.0, np.exp(np.minimum(pre, 0)))
    if name == 'selu':
        return dy * SELU_SCALE * np.where(pre > 0, 1.0, SELU_ALPHA * np.exp(np.minimum(pre, 0)))
    raise NotImplementedError(name)


def conv_node_fwd(x, w, b, pool=(1, 1, 1), act='relu', mfp=False, offsets=None, strides=None):
    """Conv._make_output (neural.py:641-722), BN off / dropout off:
    conv -> pool|MFP -> + bias -> activation (this order; pool BEFORE bias)."""
    if tuple(w.shape[2:]) == (1, 1, 1):
        lin = conv3d_dot(x, w)
    else:
        lin = conv3d(x, w)
    aux = None
    if mfp:
        pooled, off_new, str_new = fragmentpool(lin, pool, offsets, strides)
        aux = (off_new, str_new)
    else:
        pooled = pooling(lin, pool)
    pre = pooled + np.asarray(b, F64).reshape(1, -1, 1, 1, 1)
    return activation(pre, act), (lin, pre), aux


def upconv_node_fwd(x, w, b, pool, act='relu'):
    """UpConv._make_output (neural.py:989-1072), BN/dropout off."""
    lin = upconv3d(x, w, pool)
    pre = lin + np.asarray(b, F64).reshape(1, -1, 1, 1, 1)
    return activation(pre, act), pre


# ----------------------------------------------------- pooling modes, prelu, maxout, batch norm
def pooling_mode(x, pool, mode='max'):
    """computations.pooling with dnn_pool's other modes (computations.py:556-561, 589-590, 600): 'average_inc_pad',
    'average_exc_pad' (identical here: the path never pads) and 'sum'; stride == pool, ignore_border."""
    if mode == 'average':
        mode = 'average_inc_pad'
    if mode == 'max':
        return pooling(x, pool)
    x = np.asarray(x, F64)
    pool = tuple(int(p) for p in pool)
    if all(p == 1 for p in pool):
        return x
    b, c, Z, X, Y = x.shape
    pz, px, py = pool
    Zo, Xo, Yo = Z // pz, X // px, Y // py
    v = x[:, :, :Zo * pz, :Xo * px, :Yo * py].reshape(b, c, Zo, pz, Xo, px, Yo, py)
    if mode in ('average_inc_pad', 'average_exc_pad'):
        return v.mean(axis=(3, 5, 7))
    if mode == 'sum':
        return v.sum(axis=(3, 5, 7))
    raise ValueError(mode)


def pooling_mode_bwd(dy, x_shape, pool, mode):
    """Backward of average / sum pooling: dy spread evenly over (or copied to) the window."""
    dy = np.asarray(dy, F64)
    pz, px, py = pool
    up = np.repeat(np.repeat(np.repeat(dy, pz, 2), px, 3), py, 4)
    if mode in ('average', 'average_inc_pad', 'average_exc_pad'):
        up = up / float(pz * px * py)
    dx = np.zeros(x_shape, F64)
    dx[:, :, :up.shape[2], :up.shape[3], :up.shape[4]] = up
    return dx


def prelu(pre, alpha):
    """T.nnet.relu(x, alpha) as Theano writes it: 0.5*(1+alpha)*x + 0.5*(1-alpha)*|x|
    (apply_activation 'prelu', computations.py:83-85; alpha = b[:,1] broadcast over f, neural.py:655-657)."""
    pre = np.asarray(pre, F64)
    a = np.asarray(alpha, F64).reshape(1, -1, 1, 1, 1)
    return 0.5 * (1 + a) * pre + 0.5 * (1 - a) * np.abs(pre)


def prelu_bwd(dy, pre, alpha):
    """(d/dpre, d/dalpha): slope 1 where pre > 0, alpha where pre < 0; d/dalpha = sum dy * min(pre, 0)."""
    pre = np.asarray(pre, F64)
    a = np.asarray(alpha, F64).reshape(1, -1, 1, 1, 1)
    dpre = dy * np.where(pre > 0, 1.0, a)
    dalpha = (dy * np.minimum(pre, 0.0)).sum(axis=(0, 2, 3, 4))
    return dpre, dalpha


def maxout(x, factor=2, axis=1):
    """computations.maxout (computations.py:455-495): y = max_i x[..., i::factor, ...] along ``axis``.  (The
    reference's default picks axis 2 for every input, :476-477 -- callers here pass the axis explicitly.)"""
    x = np.asarray(x)
    sl = [slice(None)] * x.ndim
    sl[axis] = slice(0, None, factor)
    y = x[tuple(sl)]
    for i in range(1, factor):
        sl[axis] = slice(i, None, factor)
        y = np.maximum(y, x[tuple(sl)])
    return y


BN_EPS = 1e-6


def batchnorm_stats(v):
    """neural.py:681-685: mean and T.std (population) over all axes but f, std + 1e-6."""
    v = np.asarray(v, F64)
    return v.mean(axis=(0, 2, 3, 4)), v.std(axis=(0, 2, 3, 4)) + BN_EPS


def batchnorm_affine(v, gamma, b, mean, std):
    """neural.py:711: (gamma / std) * v + b - gamma * mean / std."""
    r = lambda a: np.asarray(a, F64).reshape(1, -1, 1, 1, 1)
    return (r(gamma) / r(std)) * np.asarray(v, F64) + r(b) - r(gamma) * r(mean) / r(std)


def batchnorm_bwd(dpre, v, gamma, mean, std, batch_stats):
    """Gradients of ``batchnorm_affine`` w.r.t. v, gamma, b.  With ``batch_stats`` mean and std are functions of v
    (std = sigma + eps): dv = (gamma/std) [dpre - mean(dpre) - vhat (std/sigma) mean(dpre * vhat)]."""
    r = lambda a: np.asarray(a, F64).reshape(1, -1, 1, 1, 1)
    v = np.asarray(v, F64)
    dpre = np.asarray(dpre, F64)
    vhat = (v - r(mean)) / r(std)
    ax = (0, 2, 3, 4)
    db = dpre.sum(axis=ax)
    dgamma = (dpre * vhat).sum(axis=ax)
    if batch_stats:
        n = v.size / v.shape[1]
        sigma = np.asarray(std, F64) - BN_EPS
        dv = (r(gamma) / r(std)) * (dpre - r(db) / n - vhat * r(np.asarray(std, F64) / sigma) * r(dgamma) / n)
    else:
        dv = (r(gamma) / r(std)) * dpre
    return dv, dgamma, db
