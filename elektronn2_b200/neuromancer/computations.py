"""The five functions of the reference's ``neuromancer/computations.py`` that sit on the hot path, with the reference's
signatures, argument meaning and error behaviour, evaluated eagerly on the B200 (numpy in / numpy out through
``elektronn2_b200.functional``).  In the reference these build Theano expressions and are called only from
``Node._make_output``; here the nodes go through ``executor.Plan`` instead, and this module is the call-for-call
equivalent of the seam -- what ``config.backend == 'b200'`` dispatches to inside an existing install,
and what the parity tests read like.

Only what the BASELINE configs use is on the B200 path: 3-D, ``axis_order='dnn'`` (tags ``b,f,z,x,y``), ``'valid'``
borders, stride 1 convs, ``pool == stride`` max-pooling.  Everything else raises the reference's own exception type.
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
    if conv_dim != 3 or axis_order not in ('dnn', None):
        raise NotImplementedError("b200 backend: 3-D convolution in 'dnn' axis order (b,f,z,x,y) only")
    if stride is not None and not np.all(np.equal(stride, 1)):
        raise NotImplementedError("Cannot use strided conv with 3d conv")         # :375-376
    if border_mode != 'valid':
        raise NotImplementedError("b200 backend: border_mode 'valid' only (no BASELINE config uses another)")
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
    if len(pool) != 3 or list(spatial_axes) != [2, 3, 4]:
        raise NotImplementedError("b200 backend: 3-D pooling over axes (2,3,4) only")
    if mode != 'max':
        raise NotImplementedError("b200 backend: pooling mode 'max' only (the BASELINE configs' mode)")
    if any(x.shape[2 + i] % pool[i] for i in range(3)):
        raise ValueError("Cannot pool %s by %s: the spatial axes must be divisible" % (x.shape[2:], pool))
    return F.maxpool3d(x, pool)


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
