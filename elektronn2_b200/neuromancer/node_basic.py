"""Node base class, model registry and the structural nodes (Input, Concat).

Host-side mirror of the part of neuromancer/node_basic.py that the hot path
touches: ``model_manager`` (node_basic.py:51-156), unique node names (:158-195),
automatic registration + ``_finalize_init`` (:200-302, 502-539), ``Node.__call__``
(:464-494), ``predict_dense`` (:805-1012), ``test_run`` (:1014-1063), ``Input`` /
``Input_like`` (:1182-1271) and ``Concat`` (:1403-1451).

Where the reference builds a symbolic Theano graph and compiles it lazily, a node
here only records *what* to compute; ``executor.Plan`` turns a set of nodes into a
static launch sequence over libe2b200 on first call (also lazily, per batch size).
"""
from collections import OrderedDict
import logging
import re
import time
import uuid

import numpy as np

from .graphutils import TaggedShape, floatX

logger = logging.getLogger('elektronn2log')


class ModelContainer(object):
    """Keeps track of the network models (node_basic.py:51-150)."""

    def __init__(self):
        self._default_set = False
        self.models = OrderedDict()
        self.last = None
        self.current = None

    def __getitem__(self, item):
        return self.models[item]

    def __repr__(self):
        return repr(list(self.models.keys()))

    def setdefault(self):
        if self._default_set:
            raise RuntimeError("The default model has already been set.")
        from .model import Model
        self.models["default"] = Model(name="default")
        self.current = self.last = self["default"]
        self._default_set = True

    def newmodel(self, name):
        from .model import Model
        if name in self.models:
            raise ValueError("Model of the same name %s exists already." % (name,))
        if name is None:
            name = str(uuid.uuid4())
        self.models[name] = Model(name=name)
        self.last = self.current
        self.current = self[name]
        return self[name]

    def getmodel(self, *args):
        if len(args) == 1:
            current = self[args[0]]
        elif len(args) == 0:
            current = self["default"]
        else:
            raise ValueError("Either provide name or nothing!")
        self.last = self.current
        self.current = current
        return current

    def togglemodel(self):
        self.last, self.current = self.current, self.last
        return self.current

    def reset(self):
        """Forget all models (the reference needs a fresh interpreter for this)."""
        self.__init__()


model_manager = ModelContainer()


def pretty_string_ops(n):
    """Humanised op count, '25.2 Giga Ops' (utils_basic.py:740-751)."""
    for factor, suffix in [(1000000000000, 'Tera Ops'), (1000000000, 'Giga Ops'), (1000000, 'Mega Ops'), (1000, 'kilo Ops')]:
        if n >= factor:
            break
    return "%.1f %s" % (float(n) / factor, suffix)


def choose_name(proposal, names):
    """Unique node name: 'conv', 'conv1', 'conv2', ... (node_basic.py:158-195)."""
    if proposal in names:
        if not re.findall(r'(\d+)$', proposal):
            proposal = proposal + '1'
        while proposal in names:
            proposal = re.sub(r'(\d+)$', lambda m: str(int(m.group(0)) + 1), proposal)
    return proposal


class _MetaNode(type):
    """Registers every constructed node with the current model under a unique name,
    then runs ``_finalize_init`` exactly once (the job of MetaNode.init_register,
    node_basic.py:229-302)."""

    def __call__(cls, *args, **kwargs):
        import inspect
        if model_manager.current is None:
            model_manager.setdefault()
        model = model_manager.current
        sig = inspect.signature(cls.__init__)
        default_name = ''
        if 'name' in sig.parameters and isinstance(sig.parameters['name'].default, str):
            default_name = sig.parameters['name'].default
        # a positional name is honoured too
        bound = sig.bind_partial(None, *args, **kwargs)
        name = choose_name(bound.arguments.get('name', default_name), model.node_descriptors.keys())
        bound.arguments['name'] = name
        node = cls.__new__(cls)
        model.register_node(node, name, cls, args, dict(kwargs, name=name))
        b_args, b_kwargs = list(bound.args[1:]), bound.kwargs
        try:
            cls.__init__(node, *b_args, **b_kwargs)
            node._finalize_init()
        except Exception:
            # a node that failed to build must not stay registered
            model.nodes.pop(name, None)
            model.node_descriptors.pop(name, None)
            for p in (node.parents if hasattr(node, 'parent') else []):
                p.children.pop(name, None)
            raise
        return node


class Node(object, metaclass=_MetaNode):
    """Basic node: records parents, shape, params and cost."""

    def __init__(self, parent, name="", print_repr=False):
        self.parent = parent
        self.children = OrderedDict()
        self.name = name
        self._features_names = None
        self.params = OrderedDict()
        self.computational_cost = 0
        self.is_source = False
        self.shape = None
        self.dtype = floatX
        self._finalized = False
        self._print_repr = print_repr
        self._plans = {}
        self.last_exec_time = None
        self.model = model_manager.current

    # -- graph structure -------------------------------------------------------
    @property
    def parents(self):
        p = self.parent
        if p is None:
            return []
        return list(p) if isinstance(p, (list, tuple)) else [p]

    def _finalize_init(self):
        if self._finalized:
            return
        for p in self.parents:
            p.children[self.name] = self
        self._calc_shape()
        self._calc_comp_cost()
        self._finalized = True
        if self._print_repr:
            logger.info(repr(self))
            print(repr(self))

    def _calc_shape(self):
        self.shape = self.parents[0].shape.copy()

    def _calc_comp_cost(self):
        """Default: one op per element of the (first) parent's output (node_basic.py:566-573)."""
        ps = self.parents
        self.computational_cost = int(ps[0].shape.stripnone_prod) if ps and ps[0].shape is not None else 0

    @property
    def input_nodes(self):
        """Source nodes this node depends on, in registration order."""
        seen, out = set(), []

        def visit(n):
            if id(n) in seen:
                return
            seen.add(id(n))
            if n.is_source:
                out.append(n)
            for p in n.parents:
                visit(p)
        visit(self)
        order = list(self.model.nodes.values())
        return sorted(out, key=lambda n: order.index(n))

    def ancestors(self):
        """This node and everything it depends on, in registration (topological) order."""
        seen = {}

        def visit(n):
            if id(n) in seen:
                return
            seen[id(n)] = n
            for p in n.parents:
                visit(p)
        visit(self)
        return [n for n in self.model.nodes.values() if id(n) in seen]

    @property
    def all_parents(self):
        """OrderedDict name -> node of this node and everything it depends on (node_basic.py:628-648)."""
        return OrderedDict((n.name, n) for n in self.ancestors())

    @property
    def all_children(self):
        """OrderedDict name -> node of everything computed from this node (node_basic.py:757-765)."""
        out = OrderedDict()
        for child in self.children.values():
            out[child.name] = child
            out.update(child.all_children)
        return out

    @property
    def param_count(self):
        return int(sum(np.prod(p.shape) for p in self.params.values() if p.apply_train))

    @property
    def all_params(self):
        out = OrderedDict()
        for n in self.ancestors():
            for k, p in n.params.items():
                out["%s_%s" % (n.name, k)] = p
        return out

    @property
    def all_trainable_params(self):
        return OrderedDict((k, p) for k, p in self.all_params.items() if p.apply_train)

    @property
    def all_nontrainable_params(self):
        return OrderedDict((k, p) for k, p in self.all_params.items() if not p.apply_train)

    @property
    def all_params_count(self):
        return int(sum(np.prod(p.shape) for p in self.all_trainable_params.values()))

    @property
    def all_computational_cost(self):
        return int(sum(n.computational_cost for n in self.ancestors()))

    @property
    def feature_names(self):
        """Names of the features on the 'f' axis (node_basic.py:787-802)."""
        return self._features_names if self._features_names else None

    @feature_names.setter
    def feature_names(self, value):
        if 'f' not in self.shape.tags:
            raise ValueError("Shape has no feature tag")
        if len(value) != self.shape['f']:
            raise ValueError("Shape has %i features, but %i names were given" % (self.shape['f'], len(value)))
        self._features_names = tuple(value)

    def get_param_values(self, skip_const=False):
        return OrderedDict((k, p.get_value()) for k, p in self.params.items() if not (skip_const and p.constant))

    def set_param_values(self, value_dict, skip_const=False):
        for k, v in value_dict.items():
            if skip_const and self.params[k].constant:
                continue
            self.params[k].set_value(v)

    def __repr__(self):
        s = "<%s-Node> '%s' \n  " % (self.__class__.__name__, self.name)
        if self.param_count > 0:
            s += "#Params={0:,d} ".format(self.param_count)
        if self.computational_cost > 0:
            s += "Comp.Cost=%s, " % (pretty_string_ops(self.computational_cost),)
        s += "Out:%s" % (str(self.shape),)
        if len(self.input_nodes) > 1:                               # node_basic.py:458-460
            s += "\n  Order of sources=%s, " % (str([n.name for n in self.input_nodes]),)
        return s

    # -- execution ---------------------------------------------------------------
    def _plan_for(self, batch):
        from .executor import Plan
        key = int(batch)
        if key not in self._plans:
            self._plans[key] = Plan(self.model, [self], batch)
        return self._plans[key]

    def __call__(self, *args):
        """Compute the output of this node for numpy inputs given in the order of
        ``input_nodes`` (node_basic.py:464-494).  Called without arguments it only
        prepares ("compiles") the launch plan."""
        inputs = self.input_nodes
        if len(args) == 0:
            b = inputs[0].shape['b'] or 1 if inputs else 1
            self._plan_for(b)
            return None
        if len(args) != len(inputs):
            raise ValueError("Node %s needs %d inputs %s, got %d" % (self.name, len(inputs),
                                                                     [n.name for n in inputs], len(args)))
        batch = np.shape(args[0])[0]
        plan = self._plan_for(batch)
        t0 = time.time()
        out = plan.run(dict(zip(inputs, args)))[0]
        self.last_exec_time = time.time() - t0
        return out

    def test_run(self, on_shape_mismatch='warn', debug_outputs=False):
        """node_basic.py:1014-1063."""
        vals = []
        for i in self.input_nodes:
            sh = [1 if s is None else s for s in i.shape.shape]
            vals.append(np.random.rand(*sh).astype(i.dtype))
        self()
        y = self(*vals)
        ok = tuple(y.shape) == tuple(self.shape.shape) or np.prod(y.shape) == self.shape.stripnone_prod
        speed = self.shape.spatial_size * 1.0 / 1000000 / max(self.last_exec_time, 1e-12)
        print("Node '%s' planned and computation tested\nOutshape: %s\nRuntime: %g s\nSpeed: %.3f MB or MPix /s\n%s"
              % (self.name, y.shape, self.last_exec_time, speed,
                 "Shapes agree" if ok else "Shapes do not agree (Outshape should be %s)" % self.shape.shape))
        if not ok and on_shape_mismatch == 'warn':
            logger.warning("Shape of computed output and shape calculated by layer definition not match")
        return y

    def measure_exectime(self, n_samples=5, n_warmup=4, print_info=True, local=True, nonegative=True):
        """Time this node's computation in milliseconds (node_basic.py:1092-1176).  The reference compiles the node's
        function under Theano's profiler, sums the op times and, for ``local=True``, subtracts its parents' totals.
        Here every kernel launch of the node's plan is bracketed by CUDA events on the launching stream
        (``executor.Plan.profile``): the total is the sum over the plan's launches, the local time is the sum over
        the launches that belong to this node (a pool computed in its conv's epilogue costs its Pool node nothing)."""
        inp = self.input_nodes
        if not inp:
            logger.info('Node {} has no inputs -> skipping node, giving it zero execution time.'.format(self.name))
            return 0.0
        vals = []
        for i in inp:
            sh = [1 if s_ is None else s_ for s_ in i.shape.shape]
            vals.append(np.random.rand(*sh).astype(i.dtype))
        plan = self._plan_for(np.shape(vals[0])[0])
        if not plan.fwd_ops:                                           # nothing to launch (e.g. a pure view of an input)
            self._total_exec_time = self._local_exec_time = 0.0
            return 0.0
        plan.feed(dict(zip(inp, vals)))
        for _ in range(max(1, int(n_warmup))):                      # at least once: commits the inputs to the device
            plan.execute()
        samples = np.zeros((max(1, int(n_samples)), 2))
        for k in range(samples.shape[0]):
            rows = plan.profile(repeats=1)
            samples[k, 0] = sum(ms for (_, _, _, _, ms) in rows)
            samples[k, 1] = sum(ms for (label, _, _, _, ms) in rows if self._owns_launch(label))
        self._total_exec_time = float(np.median(samples[:, 0]))
        self._local_exec_time = float(np.median(samples[:, 1]))
        t = self._local_exec_time if local else self._total_exec_time
        if nonegative and t < 0:
            t = 0.0
        if print_info:
            logger.info('{0} samples in ms:\n{1}\n{0}: median execution time: {2} ms\n'
                        .format(self.name, samples[:, 1 if local else 0], t))
        return t

    def _owns_launch(self, label):
        kind, _, name = label.partition(':')
        if name:
            return name == self.name
        return kind.startswith('softmax') and type(self).__name__ == 'Softmax'      # the fused loss head

    @property
    def local_exec_time(self):
        if getattr(self, '_local_exec_time', None) is None:
            self.measure_exectime(print_info=False, local=True)
        return self._local_exec_time

    @property
    def total_exec_time(self):
        if getattr(self, '_total_exec_time', None) is None:
            self.measure_exectime(print_info=False, local=True)
        return self._total_exec_time

    def predict_dense(self, raw_img, as_uint8=False, pad_raw=False):
        """Tiled dense inference (node_basic.py:860-1012); see ``dense.predict_dense``."""
        from .dense import predict_dense
        return predict_dense(self, raw_img, as_uint8=as_uint8, pad_raw=pad_raw)


class Input(Node):
    """Source node (node_basic.py:1182-1243)."""

    def __init__(self, shape, tags, strides=None, fov=None, dtype=floatX, hardcoded_shape=False,
                 name='input', print_repr=True):
        super(Input, self).__init__(None, name, print_repr)
        self._init_shape = TaggedShape(shape, tags, strides, fov=fov)
        self.dtype = dtype
        self.hardcoded_shape = hardcoded_shape
        self.is_source = True

    def _calc_shape(self):
        self.shape = self._init_shape

    def _calc_comp_cost(self):
        self.computational_cost = 0


def Input_like(ref, dtype=None, name='input', print_repr=True, override_f=False, hardcoded_shape=False):
    """node_basic.py:1246-1271."""
    if isinstance(ref, Node):
        shape = list(ref.shape.shape)
        tags, strides, fov = ref.shape.tags, ref.shape.strides, ref.shape.fov
        if override_f:
            shape[ref.shape.tag2index('f')] = override_f
        if dtype is None:
            dtype = ref.dtype
    elif isinstance(ref, TaggedShape):
        shape, tags, strides, fov = ref.shape, ref.tags, ref.strides, ref.fov
        assert dtype is not None
    else:
        raise ValueError("ref must be Node or TaggedShape.")
    return Input(shape, tags, strides, fov=fov, dtype=dtype, name=name, print_repr=print_repr,
                 hardcoded_shape=hardcoded_shape)


class Concat(Node):
    """Concatenate parents along one axis (node_basic.py:1403-1451); only the feature
    axis is on the hot path."""

    def __init__(self, parent_nodes, axis='f', name="concat", print_repr=True):
        super(Concat, self).__init__(list(parent_nodes), name, print_repr)
        self.axis = axis

    def _calc_shape(self):
        ref = self.parents[0].shape
        ax = self.axis if isinstance(self.axis, int) else ref.tag2index(self.axis)
        if ax != ref.tag2index('f'):
            raise NotImplementedError("Concat on the B200 path supports axis='f' only")
        for p in self.parents[1:]:
            for i, (a, b) in enumerate(zip(ref.shape, p.shape.shape)):
                if i != ax and a != b:
                    raise ValueError("Cannot concatenate %s and %s on axis %s" % (ref, p.shape, self.axis))
        size = sum(p.shape[ax] for p in self.parents)
        self.shape = ref.updateshape(ax, size)

    def _calc_comp_cost(self):
        self.computational_cost = 0
