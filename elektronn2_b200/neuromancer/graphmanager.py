"""The reference's on-disk model format (``.mdl``): descriptors + per-node parameter dicts.

Reference: ``Model.save`` (neuromancer/model.py:229-235) pickles ``(descriptors, desig_descr)`` with protocol 2
(utils_basic.py:602-613); ``descriptors`` is ``GraphManager.serialise()`` (graphmanager.py:236-247): an OrderedDict
``name -> [NodeDescriptor, OrderedDict(param name -> ndarray)]`` where a ``NodeDescriptor`` (graphmanager.py:49-117)
holds the node class and its constructor ``args`` / ``kwargs`` with parent nodes replaced by ``NodePointer(name)``;
``desig_descr`` maps ``input_node`` / ``target_node`` / ... to node names (model.py:95-130).  ``modelload``
(model.py:623-729) restores the nodes in order and re-designates them.

Weights travel as float32 arrays in the reference layout ``(f_out, f_in, kz, kx, ky)`` / ``(f_out,)`` -- that is the
whole weight-exchange contract, so a file written by a Theano install loads here and the other way round:

* writing: module paths in the pickle are the REFERENCE's (``elektronn2.neuromancer.neural Conv``,
  ``numpy.core.multiarray _reconstruct``), whatever this package and the installed numpy call themselves;
* reading: ``elektronn2.*`` globals resolve to this package's mirror classes, anything the B200 path does not have
  (Theano objects, trainer state) becomes an inert placeholder.

No Theano install is available to produce a reference-written file in this environment, so the reader is tested on
files produced by the writer and on a hand-assembled pickle that uses the reference's module paths
(tests/test_host_api.py); parity of the format itself rests on the code citations above.
"""
from collections import OrderedDict
import gzip
import io
import pickle
import sys

import numpy as np

REF_PKG, OUR_PKG = 'elektronn2', 'elektronn2_b200'


class NodePointer(object):                       # graphmanager.py:20-25
    def __init__(self, target_id):
        self.target_id = target_id

    def __repr__(self):
        return "<NodePointer> %s" % self.target_id


class ParamPointer(object):                      # graphmanager.py:27-33
    def __init__(self, target_id, param_name):
        self.target_id = target_id
        self.param_name = param_name

    def __repr__(self):
        return "<ParamPointer> %s of node %s" % (self.param_name, self.target_id)


class SplitDescriptor(object):                   # graphmanager.py:35-45 (split nodes are not on the B200 path)
    def __init__(self, node_id, func, args, kwargs):
        self.node_id, self.func, self.args, self.kwargs = node_id, func, args, kwargs


class NodeDescriptor(object):
    """Constructor record of one node (graphmanager.py:49-117): ``cls``, ``args``, ``kwargs``."""

    def __init__(self, args, kwargs, cls, gm=None):
        from .node_basic import Node

        def conv(a):
            if isinstance(a, Node):
                return NodePointer(a.name)
            if isinstance(a, (list, tuple)) and len(a) and all(isinstance(x, Node) for x in a):
                return [NodePointer(x.name) for x in a]
            return a

        self.args = [conv(a) for a in args]
        # ndarray keyword arguments (initial weights) are dropped like in the reference (:107-108): the values
        # travel in the parameter dict
        self.kwargs = dict((k, conv(v)) for k, v in kwargs.items() if not isinstance(v, np.ndarray))
        self.cls = cls

    def restore(self, param_values, nodes, override_mfp_to_active=False):
        def back(a):
            if isinstance(a, NodePointer):
                return nodes[a.target_id]
            if isinstance(a, (list, tuple)) and len(a) and all(isinstance(x, NodePointer) for x in a):
                return [nodes[x.target_id] for x in a]
            if isinstance(a, ParamPointer):
                return nodes[a.target_id].params[a.param_name]
            return a

        args = [back(a) for a in self.args]
        kwargs = dict((k, back(v)) for k, v in self.kwargs.items())
        kwargs['print_repr'] = False
        if override_mfp_to_active and self.cls.__name__ == 'Conv':      # graphmanager.py:183-185
            kwargs['mfp'] = True
        node = self.cls(*args, **kwargs)
        node.set_param_values(param_values, skip_const=True)             # :188
        return node


class _Placeholder(object):
    """Stands in for pickled objects of classes this package does not have."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        self.state = state


def serialise(model):
    """``GraphManager.serialise`` (graphmanager.py:236-247) for a mirror Model."""
    out = OrderedDict()
    for name, (cls, args, kwargs) in model.node_descriptors.items():
        kw = dict((k, v) for k, v in kwargs.items() if k != 'print_repr')
        out[name] = [NodeDescriptor(args, kw, cls), model.nodes[name].get_param_values()]
    return out


def _desig_names(model):
    def nm(v):
        if v is None:
            return None
        if isinstance(v, (list, tuple)):
            return [nm(x) for x in v]
        return v if isinstance(v, str) else v.name
    return dict((k, nm(v)) for k, v in model._desig_descr.items())


class _RefPickler(pickle._Pickler):
    """Protocol-2 pickler that writes the reference's module paths for globals."""

    @staticmethod
    def _remap(module):
        if module == OUR_PKG or module.startswith(OUR_PKG + '.'):
            return REF_PKG + module[len(OUR_PKG):]
        if module == 'numpy._core' or module.startswith('numpy._core.'):
            return 'numpy.core' + module[len('numpy._core'):]
        return module

    def save_global(self, obj, name=None):
        if name is None:
            name = getattr(obj, '__qualname__', None) or obj.__name__
        module = self._remap(pickle.whichmodule(obj, name))
        if module == 'builtins':
            module = '__builtin__'                      # what a protocol-2 stream calls it (python 2 reference)
        self.write(pickle.GLOBAL + module.encode('utf-8') + b'\n' + name.encode('utf-8') + b'\n')
        self.memoize(obj)


def save_model(model, file_name):
    """``Model.save`` (model.py:229-235): pickle ``(descriptors, desig_descr)`` with protocol 2."""
    buf = io.BytesIO()
    _RefPickler(buf, protocol=2).dump((serialise(model), _desig_names(model)))
    with open(file_name, 'wb') as f:
        f.write(buf.getvalue())


class _RefUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == REF_PKG or module.startswith(REF_PKG + '.'):
            ours = OUR_PKG + module[len(REF_PKG):]
            try:
                __import__(ours)
                mod = sys.modules[ours]
                if hasattr(mod, name):
                    return getattr(mod, name)
            except ImportError:
                pass
            # a node class lives in the mirror's flat namespace whatever sub-module the reference kept it in
            from .. import neuromancer
            if hasattr(neuromancer, name):
                return getattr(neuromancer, name)
            return _Placeholder
        if module.split('.')[0] in ('theano', 'pygpu'):
            return _Placeholder
        if module == '__builtin__':
            module = 'builtins'
        if module == 'copy_reg':
            module = 'copyreg'
        return super(_RefUnpickler, self).find_class(module, name)


def _load_all(file_name):
    def read(opener):
        out = []
        with opener(file_name, 'rb') as f:
            while True:
                try:
                    out.append(_RefUnpickler(f, encoding='latin1').load())   # utils_basic.py:629-631
                except EOFError:
                    break
        return out
    try:
        ret = read(open)
    except pickle.UnpicklingError:
        ret = read(gzip.open)                                                # :641-642
    return ret[0] if len(ret) == 1 else ret


def load_model(file_name, override_mfp_to_active=False, imposed_patch_size=None, imposed_batch_size=None, name=None):
    """``modelload`` (model.py:623-729).  The overrides are applied by re-building the restored model with
    ``rebuild_model`` (same effect as the reference's descriptor surgery: Conv nodes get ``mfp=True``, a
    FragmentsToDense goes in front of the prediction node, the patch size is snapped to a valid one)."""
    from .node_basic import model_manager
    from .model import rebuild_model
    node_descr, desig_descr = _load_all(file_name)
    if 'input_node' not in desig_descr and (override_mfp_to_active or imposed_patch_size is not None or
                                            imposed_batch_size is not None):
        raise ValueError("To use 'override_mfp_to_active' or 'imposed_patch_size', the saved model must have a "
                         "designated 'input_node'")                           # model.py:635-639
    model = model_manager.newmodel(name)
    for nm_, descr in node_descr.items():
        if isinstance(descr, SplitDescriptor):
            raise NotImplementedError("split nodes are not on the B200 path")
        if not isinstance(descr[0], NodeDescriptor):
            raise ValueError("Unknown descriptor: %s." % (descr,))            # graphmanager.py:281-282
        if isinstance(descr[0].cls, type) and issubclass(descr[0].cls, _Placeholder):
            raise NotImplementedError("node '%s' is of a class that is not on the B200 hot path" % nm_)
        descr[0].restore(descr[1], model.nodes)
    if desig_descr:
        model.designate_nodes(**desig_descr)
    if override_mfp_to_active or imposed_patch_size is not None or imposed_batch_size is not None:
        model = rebuild_model(model, override_mfp_to_active=override_mfp_to_active,
                              imposed_patch_size=imposed_patch_size, imposed_batch_size=imposed_batch_size, name=name)
    return model
