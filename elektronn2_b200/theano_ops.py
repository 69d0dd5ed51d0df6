"""Theano Ops over the B200 library: the compatibility face for an existing ELEKTRONN2 install.

Same Op convention as the reference's own plugins (``malis/malisop.py:19-123``, ``neuromancer/various.py:254-305``):
``__props__``, ``make_node -> gof.Apply``, ``perform(node, inputs, output_storage)`` with numpy in / numpy out,
``grad`` returning the sibling Op.  ``computations.py`` selects them with one more backend flag:

    if config.backend == 'b200' and conv_dim == 3 and axis_order == 'dnn':
        return B200Conv3d()(x, w)

Importing this module needs Theano (``theano>=0.8,<0.10``); it is not installable in the image this repository is
built in, so the wiring is exercised against a minimal stand-in for the ``theano`` package (tests/test_host_api.py,
tests/test_gpu_ops.py) and the arithmetic through ``elektronn2_b200.functional``, which is what ``perform`` calls.
A ``perform``-based Op round-trips through host memory on every call -- that is how Theano's Python Ops work; the
executor in ``neuromancer/`` is the fast face.
"""
import theano
import theano.tensor as T
from theano import gof

from . import functional as F


def _as5(v):
    v = T.as_tensor_variable(v)
    if v.ndim != 5:
        raise TypeError("expected a 5-d tensor (b, f, z, x, y), got ndim=%d" % v.ndim)
    return v


class _B200Op(theano.Op):
    def _out(self, like):
        return like.type()


class B200Conv3d(_B200Op):
    """y = conv3d(x, w): 'valid', stride 1, true convolution; replaces dnn.dnn_conv3d / conv3d2d.conv3d
    (computations.py:389-428)."""
    __props__ = ('compute',)

    def __init__(self, compute='tf32'):
        self.compute = compute

    def make_node(self, x, w):
        x, w = _as5(x), _as5(w)
        return gof.Apply(self, [x, w], [self._out(x)])

    def perform(self, node, inputs, output_storage):
        x, w = inputs
        output_storage[0][0] = F.conv3d(x, w, compute=self.compute)

    def infer_shape(self, node, shapes):
        xs, ws = shapes
        return [(xs[0], ws[0]) + tuple(xs[2 + i] - ws[2 + i] + 1 for i in range(3))]

    def grad(self, inputs, output_grads):
        x, w = inputs
        dy, = output_grads
        return [B200Conv3dGradI(self.compute)(dy, w, x.shape), B200Conv3dGradW(self.compute)(x, dy, w.shape)]


class B200Conv3dGradI(_B200Op):
    __props__ = ('compute',)

    def __init__(self, compute='tf32'):
        self.compute = compute

    def make_node(self, dy, w, x_shape):
        dy, w = _as5(dy), _as5(w)
        return gof.Apply(self, [dy, w, T.as_tensor_variable(x_shape)], [self._out(dy)])

    def perform(self, node, inputs, output_storage):
        dy, w, x_shape = inputs
        output_storage[0][0] = F.conv3d_grad_input(dy, w, x_shape, compute=self.compute)


class B200Conv3dGradW(_B200Op):
    __props__ = ('compute',)

    def __init__(self, compute='tf32'):
        self.compute = compute

    def make_node(self, x, dy, w_shape):
        x, dy = _as5(x), _as5(dy)
        return gof.Apply(self, [x, dy, T.as_tensor_variable(w_shape)], [self._out(x)])

    def perform(self, node, inputs, output_storage):
        x, dy, w_shape = inputs
        output_storage[0][0] = F.conv3d_grad_weights(x, dy, w_shape, compute=self.compute)


class B200UpConv3d(_B200Op):
    """computations.upconv 3-D (computations.py:216-255) with kernel == stride == pool."""
    __props__ = ('pool', 'compute')

    def __init__(self, pool, compute='tf32'):
        self.pool, self.compute = tuple(int(p) for p in pool), compute

    def make_node(self, x, w):
        x, w = _as5(x), _as5(w)
        return gof.Apply(self, [x, w], [self._out(x)])

    def perform(self, node, inputs, output_storage):
        x, w = inputs
        output_storage[0][0] = F.upconv3d(x, w, self.pool, compute=self.compute)

    def infer_shape(self, node, shapes):
        xs, ws = shapes
        return [(xs[0], ws[0]) + tuple(xs[2 + i] * self.pool[i] for i in range(3))]

    def grad(self, inputs, output_grads):
        x, w = inputs
        dy, = output_grads
        return [B200UpConv3dGradI(self.pool, self.compute)(dy, w, x.shape),
                B200UpConv3dGradW(self.pool, self.compute)(x, dy, w.shape)]


class B200UpConv3dGradI(_B200Op):
    __props__ = ('pool', 'compute')

    def __init__(self, pool, compute='tf32'):
        self.pool, self.compute = tuple(int(p) for p in pool), compute

    def make_node(self, dy, w, x_shape):
        dy, w = _as5(dy), _as5(w)
        return gof.Apply(self, [dy, w, T.as_tensor_variable(x_shape)], [self._out(dy)])

    def perform(self, node, inputs, output_storage):
        dy, w, x_shape = inputs
        output_storage[0][0] = F.upconv3d_grad_input(dy, w, x_shape, self.pool, compute=self.compute)


class B200UpConv3dGradW(_B200Op):
    __props__ = ('pool', 'compute')

    def __init__(self, pool, compute='tf32'):
        self.pool, self.compute = tuple(int(p) for p in pool), compute

    def make_node(self, x, dy, w_shape):
        x, dy = _as5(x), _as5(dy)
        return gof.Apply(self, [x, dy, T.as_tensor_variable(w_shape)], [self._out(x)])

    def perform(self, node, inputs, output_storage):
        x, dy, w_shape = inputs
        output_storage[0][0] = F.upconv3d_grad_weights(x, dy, w_shape, self.pool, compute=self.compute)


class B200MaxPool3d(_B200Op):
    """computations.pooling 3-D 'max', stride == pool (computations.py:593-631)."""
    __props__ = ('pool', 'tie_mode')

    def __init__(self, pool, tie_mode='first'):
        self.pool, self.tie_mode = tuple(int(p) for p in pool), tie_mode

    def make_node(self, x):
        x = _as5(x)
        return gof.Apply(self, [x], [self._out(x)])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = F.maxpool3d(inputs[0], self.pool)

    def infer_shape(self, node, shapes):
        xs, = shapes
        return [(xs[0], xs[1]) + tuple(xs[2 + i] // self.pool[i] for i in range(3))]

    def grad(self, inputs, output_grads):
        return [B200MaxPool3dGrad(self.pool, self.tie_mode)(inputs[0], output_grads[0])]


class B200MaxPool3dGrad(_B200Op):
    __props__ = ('pool', 'tie_mode')

    def __init__(self, pool, tie_mode='first'):
        self.pool, self.tie_mode = tuple(int(p) for p in pool), tie_mode

    def make_node(self, x, dy):
        x, dy = _as5(x), _as5(dy)
        return gof.Apply(self, [x, dy], [self._out(x)])

    def perform(self, node, inputs, output_storage):
        x, dy = inputs
        output_storage[0][0] = F.maxpool3d_grad(x, dy, self.pool, self.tie_mode)


class B200FragmentPool(_B200Op):
    """computations.fragmentpool (computations.py:652-678); offsets / strides bookkeeping stays in Python."""
    __props__ = ('pool',)

    def __init__(self, pool):
        self.pool = tuple(int(p) for p in pool)

    def make_node(self, x):
        x = _as5(x)
        return gof.Apply(self, [x], [self._out(x)])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = F.fragmentpool(inputs[0], self.pool)


class B200Frag2Dense(_B200Op):
    """computations.fragments2dense (computations.py:681-701)."""
    __props__ = ('offsets', 'strides')

    def __init__(self, offsets, strides):
        self.offsets = tuple(tuple(int(v) for v in o) for o in offsets)
        self.strides = tuple(int(s) for s in strides)

    def make_node(self, fragments):
        f = _as5(fragments)
        return gof.Apply(self, [f], [self._out(f)])

    def perform(self, node, inputs, output_storage):
        output_storage[0][0] = F.fragments2dense(inputs[0], self.offsets, self.strides)
