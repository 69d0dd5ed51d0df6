#!/usr/bin/env python
"""Write ``tests/golden/ref_written_small.mdl`` with the REFERENCE's own serialisation code.

Run in the BUILD container only (needs /root/reference):
    python tests/golden/make_mdl_fixture.py

What runs from /root/reference, unmodified:
  * ``elektronn2/neuromancer/graphmanager.py`` -- imported as the real module ``elektronn2.neuromancer.graphmanager``
    (``GraphManager.register_node`` -> ``NodeDescriptor.__init__`` pointer replacement :63-117,
    ``GraphManager.serialise`` :236-247);
  * ``picklesave`` of ``elektronn2/utils/utils_basic.py`` :602-613 (protocol 2);
  * ``Model.designate_nodes``'s name bookkeeping is restated literally (model.py:106-131: ``_desig_descr[purpose]=name``).
What cannot run: the node classes themselves (their constructors build Theano graphs; Theano is not installable
here).  They are replaced by inert stand-ins registered under the reference's module paths
(``elektronn2.neuromancer.neural.Conv`` ...), carrying exactly what the reference's MetaNode hands to
``register_node`` -- the constructor ``args`` / ``kwargs`` as the user wrote them (examples/neuro3d_lite.py style) and
the parameter arrays ``get_param_values`` returns (node_basic.py:576-598).  The file therefore has the reference's
pickle structure, written by the reference's code; what it does not prove is anything about Theano numerics.

Next to the model the script stores ``ref_written_small.npz``: the parameter arrays (seeded), one input patch and the
float64-oracle prediction for it, so the GPU test can check ``modelload(fixture) -> predict``.
"""
import importlib
import os
import re
import sys
import types
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/elektronn2'
sys.path.insert(0, ROOT)


def install_reference_stubs():
    if not hasattr(np, 'int'):
        np.int = int
    th = types.ModuleType('theano')
    th.config = types.SimpleNamespace(floatX='float32')
    tt = types.ModuleType('theano.tensor')

    class Variable(object):      # graphmanager.py tests ``isinstance(arg, T.Variable)``
        pass
    tt.Variable = Variable
    th.tensor = tt
    sys.modules['theano'] = th
    sys.modules['theano.tensor'] = tt
    # package shells with the reference's paths: sub-modules are then imported from the real files
    pkg = types.ModuleType('elektronn2')
    pkg.__path__ = [REF]
    nmz = types.ModuleType('elektronn2.neuromancer')
    nmz.__path__ = [os.path.join(REF, 'neuromancer')]
    sys.modules['elektronn2'] = pkg
    sys.modules['elektronn2.neuromancer'] = nmz

    # stand-in node classes under the reference's module paths
    def module(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    nb = module('elektronn2.neuromancer.node_basic')
    ne = module('elektronn2.neuromancer.neural')
    lo = module('elektronn2.neuromancer.loss')

    class Param(object):
        constant = False

        def __init__(self, value):
            self.value = value

        def get_value(self):
            return self.value

    class Node(object):
        is_source = False

        def __init__(self, name, params):
            self.name = name
            self.params = OrderedDict((k, Param(v)) for k, v in params.items())
            self.children = OrderedDict()

        def get_param_values(self, skip_const=False):      # node_basic.py:591-598
            p_dict = OrderedDict()
            for k, v in self.params.items():
                if v.constant and skip_const:
                    continue
                p_dict[k] = v.get_value()
            return p_dict
    Node.__module__ = nb.__name__
    nb.Node = Node
    nb.Param = Param

    def choose_name(proposal, names):
        return proposal
    nb.choose_name = choose_name

    def stand_in(mod, name):
        cls = type(name, (Node,), {})
        cls.__module__ = mod.__name__
        cls.__qualname__ = name
        setattr(mod, name, cls)
        return cls
    for n in ('Input', 'Concat', 'FromTensor'):
        stand_in(nb, n)
    for n in ('Conv', 'UpConv', 'Pool', 'Crop', 'FragmentsToDense'):
        stand_in(ne, n)
    for n in ('Softmax', 'MultinoulliNLL', 'AggregateLoss', 'Classification', '_Errors'):
        stand_in(lo, n)
    gm = importlib.import_module('elektronn2.neuromancer.graphmanager')     # the reference's file, as is
    # picklesave: the function's own source out of utils_basic.py (the module imports h5py & co. at the top)
    src = open(os.path.join(REF, 'utils/utils_basic.py')).read()
    m = re.search(r"^def picklesave\(.*?(?=^def )", src, re.S | re.M)
    ns = {'os': os}
    import pickle as pkl
    ns['pkl'] = pkl
    exec(compile(m.group(0), 'utils_basic.py:picklesave', 'exec'), ns)
    return gm, ns['picklesave'], dict(node_basic=nb, neural=ne, loss=lo)


def main():
    gm_mod, picklesave, mods = install_reference_stubs()
    # the graph, written the way examples/neuro3d_lite.py writes one (positional parents, keyword options)
    rs = np.random.RandomState(12)
    spec = [
        ('raw', 'node_basic', 'Input', [(None, 1, 9, 31, 31), 'b,f,z,x,y'], dict(name='raw'), {}),
        ('conv', 'neural', 'Conv', ['@raw', 6, (1, 4, 4), (1, 2, 2)], dict(name='conv'),
         dict(w=(6, 1, 1, 4, 4), b=(6,))),
        ('conv1', 'neural', 'Conv', ['@conv', 8, (2, 3, 3), (2, 1, 1)], dict(name='conv1'),
         dict(w=(8, 6, 2, 3, 3), b=(8,))),
        ('conv2', 'neural', 'Conv', ['@conv1', 9, (1, 3, 3)], dict(name='conv2', activation_func='tanh'),
         dict(w=(9, 8, 1, 3, 3), b=(9,))),
        ('conv3', 'neural', 'Conv', ['@conv2', 2, (1, 1, 1)], dict(name='conv3', activation_func='lin'),
         dict(w=(2, 9, 1, 1, 1), b=(2,))),
        ('softmax', 'loss', 'Softmax', ['@conv3'], dict(name='softmax'), {}),
        ('target', 'node_basic', 'Input', [(None, 1, 4, 10, 10), 'b,f,z,x,y'],
         dict(name='target', strides=np.array([2, 2, 2]), fov=np.array([3, 13, 13]), dtype='float32', hardcoded_shape=False), {}),
        ('nll', 'loss', 'MultinoulliNLL', ['@softmax', '@target'], dict(name='nll', target_is_sparse=True), {}),
        ('loss', 'loss', 'AggregateLoss', ['@nll'], dict(name='loss'),
         dict(mixing_weights=None)),
    ]
    gm = gm_mod.GraphManager(name='fixture')
    nodes, params_out = {}, OrderedDict()
    for name, mod, cls_name, args, kwargs, pshapes in spec:
        cls = getattr(mods[mod], cls_name)
        params = OrderedDict()
        for k, sh in pshapes.items():
            if k == 'mixing_weights':
                params[k] = np.ones(1, np.float32)
            elif k == 'w':
                params[k] = (rs.randn(*sh) * np.sqrt(2.0 / np.prod(sh[1:]))).astype(np.float32)
            else:
                params[k] = (rs.randn(*sh) * 0.1).astype(np.float32)
            params_out['%s_%s' % (name, k)] = params[k]
        node = cls(name, params)
        a = [nodes[x[1:]] if isinstance(x, str) and x.startswith('@') else x for x in args]
        gm.register_node(node, name, a, kwargs)       # -> the reference's NodeDescriptor.__init__
        nodes[name] = node
    descriptors = gm.serialise()                      # the reference's GraphManager.serialise
    desig = {}

    def designate(purpose, name):                     # model.py:106-123, names only
        desig[purpose] = name
    designate('input_node', 'raw')
    designate('target_node', 'target')
    designate('loss_node', 'loss')
    designate('prediction_node', 'softmax')
    designate('error_node', None)
    designate('prediction_ext', ['loss', 'softmax'])
    designate('debug_outputs', [])
    out = os.path.join(HERE, 'ref_written_small.mdl')
    picklesave((descriptors, desig), out)             # the reference's picklesave, protocol 2

    # oracle prediction with these weights (float64), for the GPU test
    from oracle import nets as onets, loss as ol
    o = onets.Net(0)
    n = o.input((1, 1, 9, 31, 31))
    n = o.conv(n, 6, (1, 4, 4), (1, 2, 2))
    n = o.conv(n, 8, (2, 3, 3), (2, 1, 1))
    n = o.conv(n, 9, (1, 3, 3), act='tanh')
    n = o.conv(n, 2, (1, 1, 1), act='lin')
    for (node, k), key in zip(o.param_list(), [k for k in params_out if not k.startswith('loss')]):
        assert node.params[k].shape == params_out[key].shape, (key, node.params[k].shape, params_out[key].shape)
        node.params[k] = params_out[key]
    x = np.random.RandomState(0).rand(1, 1, 9, 31, 31).astype(np.float32)
    probs = ol.softmax(o.forward(x), 1)
    np.savez_compressed(os.path.join(HERE, 'ref_written_small.npz'), x=x, probs=probs,
                        **dict(('p_' + k, v) for k, v in params_out.items()))
    print('wrote', out, os.path.getsize(out), 'bytes; prediction', probs.shape)


if __name__ == '__main__':
    main()
