"""A minimal stand-in for the parts of the ``theano`` package that ``elektronn2_b200.theano_ops`` touches (Theano itself
cannot be installed in this image: python 3.12 / numpy 2).  Only the Op plumbing: symbolic variables carry ``ndim`` and
``shape``, ``Op.__call__`` runs ``make_node`` and hands back the Apply node's output(s)."""
import sys
import types


class Variable(object):
    def __init__(self, ndim, name=None, value=None):
        self.ndim, self.name, self.value, self.owner = ndim, name, value, None

    def type(self):
        return Variable(self.ndim)

    @property
    def shape(self):
        return Variable(1, name='%s.shape' % self.name)


def as_tensor_variable(v):
    if isinstance(v, Variable):
        return v
    import numpy as np
    a = np.asarray(v)
    return Variable(a.ndim, value=a)


class Apply(object):
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for o in self.outputs:
            o.owner = self


class Op(object):
    __props__ = ()

    def __call__(self, *inputs):
        node = self.make_node(*inputs)
        return node.outputs[0] if len(node.outputs) == 1 else node.outputs

    def __eq__(self, other):
        return type(self) is type(other) and all(getattr(self, p) == getattr(other, p) for p in self.__props__)

    def __hash__(self):
        return hash((type(self),) + tuple(getattr(self, p) for p in self.__props__))


def install():
    th = types.ModuleType('theano')
    tensor = types.ModuleType('theano.tensor')
    gof = types.ModuleType('theano.gof')
    tensor.as_tensor_variable = as_tensor_variable
    tensor.Variable = Variable
    gof.Apply = Apply
    th.Op, th.tensor, th.gof = Op, tensor, gof
    sys.modules['theano'], sys.modules['theano.tensor'], sys.modules['theano.gof'] = th, tensor, gof
    return th


def uninstall():
    for k in ('theano', 'theano.tensor', 'theano.gof', 'elektronn2_b200.theano_ops'):
        sys.modules.pop(k, None)
