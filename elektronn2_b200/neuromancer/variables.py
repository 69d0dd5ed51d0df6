"""Parameters and their initialisation.

Host-side mirror of neuromancer/variables.py: ``VariableParam`` /
``VariableWeight`` / ``ConstantParam`` keep ``get_value()/set_value()`` returning
arrays in the reference's layout ((f_out,f_in,kz,kx,ky) float32 for conv weights),
which is the weight-exchange contract with the reference's ``.mdl`` files
(node_basic.py:576-627).  Once a model is planned the value lives in the flat
device parameter buffer and these methods copy through it.
"""
import numpy as np

from .graphutils import floatX


class VariableParam(object):
    """Named parameter with ``apply_train`` / ``apply_reg`` flags (variables.py:25-91)."""

    def __init__(self, value=None, name=None, apply_train=True, apply_reg=True, dtype=None, **_ignored):
        self.name = name
        self.apply_train = apply_train
        self.apply_reg = apply_reg
        self.constant = False
        self._host = np.ascontiguousarray(value, dtype=dtype or floatX)
        self._dev = None    # torch view into the flat device buffer (set by the executor)
        self._on_set = []   # callbacks fired after set_value (weights need re-packing)

    @property
    def shape(self):
        return self._host.shape

    @property
    def dtype(self):
        return self._host.dtype

    def get_value(self, borrow=False):
        if self._dev is not None:
            self._host = self._dev.detach().cpu().numpy().reshape(self._host.shape).copy()
        return self._host if borrow else self._host.copy()

    def set_value(self, value, borrow=False):
        value = np.ascontiguousarray(value, dtype=self._host.dtype)
        if value.shape != self._host.shape:
            raise ValueError("Cannot set param %s of shape %s with value of shape %s"
                             % (self.name, self._host.shape, value.shape))
        self._host = value.copy()
        if self._dev is not None:
            import torch
            self._dev.copy_(torch.from_numpy(self._host).reshape(self._dev.shape))
        for cb in self._on_set:
            cb()

    def clone(self):
        return VariableParam(self.get_value(), self.name, self.apply_train, self.apply_reg)

    def __repr__(self):
        return "<%s %s %s>" % (type(self).__name__, self.name, tuple(self.shape))


class VariableWeight(VariableParam):
    """Trainable weight with optional initialisation kwargs (variables.py:94-155)."""

    def __init__(self, shape=None, init_kwargs=None, value=None, name=None, apply_train=True, apply_reg=True,
                 dtype=None):
        if value is None:
            value = initweights(shape, dtype or floatX, **(init_kwargs or {}))
        elif shape is not None and tuple(np.shape(value)) != tuple(shape):
            raise ValueError("Value shape %s does not match shape %s for %s" % (np.shape(value), shape, name))
        super(VariableWeight, self).__init__(value, name, apply_train, apply_reg, dtype)


class ConstantParam(VariableParam):
    """Non-trainable constant (variables.py:157-202)."""

    def __init__(self, value, name=None, dtype=None, make_singletons_broadcastable=True):
        super(ConstantParam, self).__init__(value, name, apply_train=False, apply_reg=False, dtype=dtype)
        self.constant = True

    def set_value(self, new_value, borrow=False):
        raise RuntimeError("ConstantParam %s cannot be changed" % self.name)


def initweights(shape, dtype=floatX, scale='glorot', mode='normal', pool=None, spatial_axes=None):
    """Weight initialisation rules of the reference (variables.py:205-266); draws from
    the global ``np.random`` state in the same order so a fixed seed gives the same
    weights as the reference would."""
    shape = tuple(int(s) for s in shape)
    if mode == 'const':
        w = np.full(shape, scale, dtype=np.float64)
    elif mode == 'prelu':
        w = np.full(shape, scale, dtype=np.float64)
        w[:, 1] = 1.0
    elif mode == 'fix-uni':
        w = np.random.uniform(-scale, scale, shape)
    elif scale == 'glorot':
        if len(shape) == 2:
            n_in, n_out = shape
            denom = n_in + n_out
        else:
            if spatial_axes is None:
                raise ValueError("glorot initialisation of conv weights needs spatial_axes")
            kernel = [s for i, s in enumerate(shape) if i in spatial_axes]
            other = [s for i, s in enumerate(shape) if i not in spatial_axes]
            if len(other) != 2:
                raise ValueError("expected exactly two non-spatial axes (n_out, n_in)")
            n_out, n_in = other
            denom = (n_in + float(n_out) / np.prod(pool)) * np.prod(kernel)
        std = np.sqrt(2.0 / denom)
        if mode == 'normal':
            w = np.random.normal(0, std, shape)
        elif mode == 'uni':
            w = np.random.uniform(-std, std, shape)
        elif mode == 'ortho':
            m = np.random.normal(0, std, size=shape).reshape((n_out, -1))
            n_flat = m.shape[1]
            strip = n_out > n_flat
            if strip:  # more vectors than can be orthogonal in this dimension
                m = np.random.normal(0, std, size=(n_out, n_out))
            _, _, v = np.linalg.svd(m, full_matrices=False)
            w = v / v.std(1)[:, None] * std
            if strip:
                w = w[:, :n_flat]
            w = w.reshape(shape)
        else:
            raise ValueError("Invalid weight initialisation mode %s" % (mode,))
    else:
        raise ValueError("Invalid weigh initialisation parameters")
    return np.ascontiguousarray(w, dtype=dtype)
