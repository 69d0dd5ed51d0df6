"""The five functions of the reference's ``neuromancer/computations.py`` that sit on the hot path, with the reference's
signatures, argument meaning and error behaviour, evaluated eagerly on the B200 (numpy in / numpy out through
``elektronn2_b200.functional``).  In the reference these build Theano expressions and are called only from
``Node._make_output``; here the nodes go through ``executor.Plan`` instead, and this module is the call-for-call
equivalent of the seam -- what ``config.backend == 'b200'`` dispatches to inside an existing install,
and what the parity tests read like.

On the B200 path: 1-D / 2-D / 3-D (``axis_order='dnn'``, tags ``b,f[,z],x[,y]``; the lower-dimensional forms,
computations.py:337-362, run as 3-D with leading unit axes), ``'valid'`` borders, stride 1 convs, ``pool == stride``
pooling in every dnn_pool mode, ``apply_activation`` and ``maxout``.  Everything else raises the reference's own
exception type.
"""
from itertools import product

import numpy as np

from .. import functional as F
from ..config import config


def conv(x, w, axis_order=None, conv_dim=None, x_shape=None, w_shape=None, border_mode='valid', stride=None):
    """computations.conv (computations.py:259-452): true convolution (flipped kernel); a 1x1x1 filter is the
    tensordot shortcut (:330-335, 377-384), numerically the same convolution."""
    x, w = np.asarray(x), np.asarray(w)
    assert axis_order in ['dnn', 'theano', None]                                  # :313
    if conv_dim is not None:
        if x.ndim != conv_dim + 2 or w.ndim != conv_dim + 2:                      # :315-318
            raise ValueError("Cannot perform %id conv on input and filter ofdim %i, %i" % (conv_dim, x.ndim, w.ndim))
    else:
        conv_dim = x.ndim - 2
        if w.ndim != conv_dim + 2:                                                # :322-325
            raise ValueError("Dimension mismatch for conv: tried to do %id convon %id input x. This requires %id "
                             "filter, but got%id" % (conv_dim, x.ndim, x.ndim, w.ndim))
        if conv_dim > 3:
            raise ValueError("Input tensor dim to big. No conv for dim>5.")       # :327
    if axis_order not in ('dnn', None):
        raise NotImplementedError("b200 backend: 'dnn' axis order (b,f,z,x,y) only")
    if stride is not None and not np.all(np.equal(stride, 1)):
        if conv_dim == 3:
            raise NotImplementedError("Cannot use strided conv with 3d conv")     # :375-376
        raise NotImplementedError("b200 backend: strided 1-D / 2-D convolution is not on the path")
    if border_mode != 'valid':
        raise NotImplementedError("b200 backend: border_mode 'valid' only (no BASELINE config uses another)")
    if conv_dim < 3:
        # 1-D (:337-349, conv2d on an added axis) and 2-D (:351-362): the same true convolution with kz (= kx) = 1
        lead = (1,) * (3 - conv_dim)
        y = F.conv3d(x.reshape(x.shape[:2] + lead + x.shape[2:]), w.reshape(w.shape[:2] + lead + w.shape[2:]),
                     compute=config.compute)
        return y.reshape(y.shape[:2] + y.shape[2 + len(lead):])
    return F.conv3d(x, w, compute=config.compute)


def upconv(x, w, stride, x_shape=None, w_shape=None, axis_order='dnn'):
    """computations.upconv (computations.py:216-255): ``w`` in the reference's swapped layout ``(f_in, f_out, k...)``
    (neural.py:1024-1030), kernel == stride."""
    assert stride is not None                                                     # :217
    stride = tuple(stride)
    x, w = np.asarray(x), np.asarray(w)
    if len(stride) != 3:
        raise NotImplementedError("b200 backend: 3-D upconv only")
    if axis_order != 'dnn':
        raise ValueError("Need dnn and dnn axis order")                           # :243-244
    if tuple(w.shape[2:]) != stride:
        raise NotImplementedError("b200 backend: upconv kernel size must equal the stride (neural.py:968)")
    return F.upconv3d(x, np.ascontiguousarray(np.swapaxes(w, 0, 1)), stride, compute=config.compute)


def pooling(x, pool, spatial_axes, mode='max', stride=None):
    """computations.pooling (computations.py:538-649), 3-D max, ``stride == pool``."""
    if np.all(np.equal(pool, 1)):                                                 # :569-570 short circuit
        return x
    pool = tuple(int(p) for p in pool)
    x = np.asarray(x)
    if stride is not None and tuple(stride) != pool:
        raise NotImplementedError("pool != stride is not implemented for 3d")     # :612-613
    nd = len(pool)
    if nd not in (1, 2, 3):
        raise NotImplementedError("Only 1/2/3-dim maxpooling with this function.")   # :646-647
    if list(spatial_axes) != list(range(2, 2 + nd)):
        if nd == 3:
            raise ValueError("Axis order not recognised, must be [2,3,4] (dnn) or [1,3,4] (theano).")   # :596-597
        raise NotImplementedError("Can only pool on last axes %s, this input has spatial axes %s"
                                  % (list(range(2, 2 + nd)), list(spatial_axes)))     # :634-635
    if mode == 'average':
        mode = 'average_inc_pad'                                                  # :591-592
    if mode not in ('max', 'average_inc_pad', 'average_exc_pad', 'sum'):
        raise ValueError("unknown pooling mode %r" % (mode,))
    if any(x.shape[2 + i] % pool[i] for i in range(nd)):
        raise ValueError("Cannot pool %s by %s: the spatial axes must be divisible" % (x.shape[2:], pool))
    lead = (1,) * (3 - nd)
    x5 = x.reshape(x.shape[:2] + lead + x.shape[2:])
    y = F.maxpool3d(x5, lead + pool) if mode == 'max' else F.pool3d(x5, lead + pool, mode)
    return y.reshape(y.shape[:2] + y.shape[2 + len(lead):])


def maxout(x, factor=2, axis=None):
    """computations.maxout (computations.py:455-495).  ``axis=None`` follows the reference literally: its default
    expression ``2 if x.ndim==5 else 2`` (:476-477) selects axis 2 for every input."""
    x = np.asarray(x)
    if axis is None:
        axis = 2
    if axis not in [1, 2]:
        raise ValueError("Maxout only permitted on axis 1 or 2")                  # :479-480
    return F.maxout(x, factor, axis)


def apply_activation(x, activation_func, b1=None):
    """computations.apply_activation (computations.py:57-134) on a (b,f,z,x,y) array; ``b1``: prelu slope per feature."""
    x = np.asarray(x)
    if isinstance(activation_func, str) and activation_func.startswith('maxout'):
        import re
        return maxout(x, factor=int(re.findall(r'\d+', activation_func)[0]))
    from .._lib import ACT
    if activation_func not in ACT:
        raise NotImplementedError("%s. Permitted activation_funcs :%s" % (activation_func, sorted(ACT)))   # :125-128
    if activation_func == 'prelu' and b1 is None:
        raise ValueError("prelu needs the slope parameter b1")
    return F.affine_act(x, activation_func, alpha=np.asarray(b1, np.float32).reshape(-1) if b1 is not None else None)


def softmax(x, axis=1, force_builtin=False):
    """computations.softmax (computations.py:137-177) over the feature axis of a ``(b, f[, z], x[, y])`` array:
    ``exp(x - max) / sum`` (:174-176; the cuDNN 'accurate' mode the reference uses on ``axis == 1`` is the same
    max-subtracted form)."""
    x = np.asarray(x)
    if x.ndim == 2 and axis == 1:
        x5 = x.T.reshape((1, x.shape[1], 1, 1, x.shape[0]))                      # T.nnet.softmax on rows (:170-171)
        return F.softmax(x5).reshape(x.shape[1], x.shape[0]).T
    if force_builtin:
        raise NotImplementedError()                                               # :172-173
    if axis != 1 or x.ndim not in (3, 4, 5):
        raise NotImplementedError("b200 backend: softmax over the feature axis (1) of a b,f[,z],x[,y] array")
    lead = (1,) * (5 - x.ndim)
    return F.softmax(x.reshape(x.shape[:2] + lead + x.shape[2:])).reshape(x.shape)


def fragmentpool(conv_out, pool, offsets, strides, spatial_axes, mode='max'):
    """computations.fragmentpool (computations.py:652-678) -> (fragments, offsets_new, strides_new)."""
    if np.all(np.equal(pool, 1)):                                                 # :653-654
        return conv_out, offsets, strides
    conv_out = np.asarray(conv_out)
    pool = tuple(int(p) for p in pool)
    if len(pool) != 3 or list(spatial_axes) != [2, 3, 4] or mode != 'max':
        raise NotImplementedError("b200 backend: 3-D max fragment pooling over axes (2,3,4) only")
    if any((conv_out.shape[2 + i] - pool[i] + 1) % pool[i] for i in range(3)):
        raise ValueError("MFP needs (S - p + 1) %% p == 0 on every axis, got %s for pool %s"
                         % (conv_out.shape[2:], pool))                             # neural.py:739-744
    offsets = np.array(offsets, int)
    strides = np.array(strides, int)
    offsets_new = []
    for ix in product(*[range(p) for p in pool]):                                 # last axis fastest (:665)
        for p in offsets:
            offsets_new.append(p + np.multiply(ix, strides))                       # :668-672
    return F.fragmentpool(conv_out, pool), np.array(offsets_new), np.multiply(pool, strides)   # :674-676


def fragments2dense(fragments, offsets, strides, spatial_axes):
    """computations.fragments2dense (computations.py:681-701)."""
    fragments = np.asarray(fragments)
    if list(spatial_axes) != [2, 3, 4]:
        raise NotImplementedError("b200 backend: fragments over spatial axes (2,3,4) only")
    if fragments.shape[0] != len(offsets) or len(offsets) != int(np.prod(strides)):
        raise ValueError("fragments2dense: need #fragments == #offsets == prod(strides) (%d, %d, %d)"
                         % (fragments.shape[0], len(offsets), int(np.prod(strides))))   # neural.py:873-875
    return F.fragments2dense(fragments, offsets, strides)
