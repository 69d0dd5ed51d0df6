"""Model: the caller of the hot path (reference: neuromancer/model.py).

Mirrors ``Model.designate_nodes`` (model.py:95-227), ``trainingstep`` (:548-600),
``loss / gradients / predict / predict_ext / predict_dense`` (:237-255),
``get/set_param_values``, ``test_run_prediction`` and the MFP re-build done by
``modelload`` / ``rebuild_model`` (:623-729, 848-867).  Instead of compiling Theano
functions, each entry point owns a static launch plan (``executor.Plan``).
"""
from collections import OrderedDict
import logging
import pickle
import time

import numpy as np

from .node_basic import Node, Input, model_manager
from .neural import Conv, FragmentsToDense
from .loss import Softmax
from . import optimiser
from .shapecalc import unet_fov_backfill, closest_valid_patch_size

logger = logging.getLogger('elektronn2log')


class Model(object):
    def __init__(self, name=""):
        self.name = name
        self.nodes = OrderedDict()
        self.node_descriptors = OrderedDict()
        self.input_node = self.target_node = self.loss_node = None
        self.prediction_node = self.error_node = None
        self.prediction_ext, self.debug_outputs = [], []
        self._desig_descr = {}
        self.trainable_params, self.nontrainable_params = [], OrderedDict()
        self.optimisers = {}
        self.batch_size = None
        self.ndim = None
        self.iterations = 0
        self.elapsed_time = 0.0
        from collections import deque
        from ..config import config as _cfg
        n_smooth = int(getattr(_cfg, 'time_per_step_smoothing_length', 50))      # config.py:77 (CircularBuffer, model.py:70-73)
        self._last_exec_times, self._last_losses = deque(maxlen=n_smooth), deque(maxlen=n_smooth)
        self._store = None
        self._train_plans, self._ext_plans = {}, {}
        self.data_parallel = None  # set by parallel.DataParallel

    # -- registration ------------------------------------------------------------
    def register_node(self, node, name, cls, args, kwargs):
        self.nodes[name] = node
        self.node_descriptors[name] = (cls, args, kwargs)

    def __getitem__(self, slice):
        if isinstance(slice, str):
            return self.nodes[slice]
        return list(self.nodes.values())[slice]

    def __repr__(self):
        return repr(list(self.nodes.keys()))                         # graphmanager.py:211-212

    # -- designation (model.py:95-227) ---------------------------------------------
    def designate_nodes(self, input_node='input', target_node=None, loss_node=None, prediction_node=None,
                        prediction_ext=None, error_node=None, debug_outputs=None):
        def resolve(v):
            if v is None:
                return None
            if isinstance(v, (list, tuple)):
                return [resolve(x) for x in v]
            return self.nodes[v if isinstance(v, str) else v.name]

        self.input_node = resolve(input_node)
        self.target_node = resolve(target_node)
        self.loss_node = resolve(loss_node)
        self.prediction_node = resolve(prediction_node)
        self.error_node = resolve(error_node)
        self.prediction_ext = resolve(prediction_ext) or []
        self.debug_outputs = resolve(debug_outputs) or []
        self._desig_descr = dict(input_node=input_node, target_node=target_node, loss_node=loss_node,
                                 prediction_node=prediction_node, prediction_ext=prediction_ext,
                                 error_node=error_node, debug_outputs=debug_outputs)
        if self.prediction_node is not None:
            psh = self.prediction_node.shape
            self.batch_size = psh['b']
            self.ndim = psh.ndim
            if np.any(np.less(psh.fov, 0)):  # UpConvs contained: back-fill the fov (model.py:141-152)
                diff = unet_fov_backfill(self.input_node.shape.spatial_shape, psh.spatial_shape, psh.strides)
                self.prediction_node.shape._fov = diff
                if self.target_node is not None:
                    self.target_node.shape._fov = diff
            elif not psh.fov_all_centered:
                logger.warning("Not all field of views are centered (odd) this might cause problems for many setups")
            self._say("Prediction properties:\n%s" % (self.prediction_node.shape.ext_repr,))       # model.py:157-158
        if self.loss_node is not None:
            self.trainable_params = list(self.loss_node.all_trainable_params.values())
            self.nontrainable_params = self.loss_node.all_nontrainable_params
            self.optimisers = dict(SGD=optimiser.SGD(self), Adam=optimiser.Adam(self))
        # aggregated model stats (model.py:160-169, 200-211)
        top = self.loss_node if self.loss_node is not None else self.prediction_node
        if top is not None:
            from .node_basic import pretty_string_ops
            n_comp = top.all_computational_cost
            ref = self.prediction_node if self.prediction_node is not None else top
            per_pixel = float(n_comp) / max(1, ref.shape.spatial_size)
            self._say("\nTotal Computational Cost of Model: {0:s}\nTotal number of trainable parameters: {1:,d}.\n"
                      "Computational Cost per pixel: {2:s}\n".format(pretty_string_ops(n_comp), top.all_params_count,
                                                                     pretty_string_ops(per_pixel)))

    @staticmethod
    def _say(text):
        logger.info(text)
        print(text)

    # -- training state the interactive shell / trainer read and set (model.py:257-546) -----------
    def _first_optimiser(self):
        if not self.optimisers:
            raise RuntimeError("no optimisers: designate a loss_node first")
        return list(self.optimisers.values())[0]

    def set_opt_meta_params(self, opt_name, value_dict):
        self.optimisers[opt_name].set_opt_meta_params(value_dict)

    lr = property(lambda self: self._first_optimiser().global_lr.get_value(),
                  lambda self, val: self._first_optimiser().setlr(val), doc="learning rate (model.py:281-293)")
    mom = property(lambda self: self._first_optimiser().global_mom.get_value(),
                   lambda self, val: self._first_optimiser().setmom(val), doc="momentum (model.py:295-307)")
    wd = property(lambda self: self._first_optimiser().global_weight_decay.get_value(),
                  lambda self, val: self._first_optimiser().setwd(val), doc="weight decay (model.py:309-321)")

    def _aggregate_loss_node(self):
        from .loss import AggregateLoss
        if isinstance(self.loss_node, AggregateLoss):
            return self.loss_node
        for n in self.nodes.values():
            if isinstance(n, AggregateLoss):
                logger.info("model.loss_node is not of type AggregateLoss and hence has no mixing_weights. "
                            "But '%s' is, so this is used." % n.name)
                return n
        logger.error("no mixing_weights found in model")
        return None

    @property
    def mixing(self):
        """Mixing weights of the AggregateLoss (model.py:323-342)."""
        n = self._aggregate_loss_node()
        return None if n is None else n.mixing_weights.get_value()

    @mixing.setter
    def mixing(self, val):
        n = self._aggregate_loss_node()
        if n is not None:
            n.mixing_weights.set_value(np.asarray(val, np.float32))
            self._train_plans.clear()       # the weight is folded into the planned loss head (executor._plan_loss_head)
            for node in self.nodes.values():
                node._plans.clear()

    def _rates(self, key):
        return [n.params[key] for n in self.nodes.values() if n.params.get(key, None)]

    def _set_rates(self, key, rates):
        for i, r in enumerate(self._rates(key)):
            r.set_value(np.float32(rates[i] if isinstance(rates, (tuple, list, np.ndarray)) else rates))

    # dropout / gradnet nodes are outside the B200 hot path (their constructors raise), so these are empty here
    dropout_rates = property(lambda self: np.array([r.get_value() for r in self._rates('dropout_rate')]),
                             lambda self, rates: self._set_rates('dropout_rate', rates), doc="model.py:358-386")
    gradnet_rates = property(lambda self: [r.get_value() for r in self._rates('gradnet_rate')],
                             lambda self, rates: self._set_rates('gradnet_rate', rates), doc="model.py:388-433")

    @property
    def batch_normalisation_active(self):
        """model.py:435-446."""
        return any(getattr(n, 'batch_normalisation', None) in ('train', 'fadeout') for n in self.nodes.values())

    @property
    def debug_output_names(self):
        return [x.name for x in self.debug_outputs] if self.debug_outputs else None

    @property
    def prediction_feature_names(self):
        return self.prediction_node.feature_names if self.prediction_node is not None else None

    @property
    def loss_input_shapes(self):
        """Shapes of the loss node's input nodes (model.py:511-523)."""
        return [s.shape for s in self.loss_node.input_nodes]

    @property
    def time_per_step(self):
        """Mean run time of the last ``config.time_per_step_smoothing_length`` training steps (model.py:525-534)."""
        return float(np.mean(self._last_exec_times)) + 1e-6 if self._last_exec_times else 1e-6

    @property
    def loss_smooth(self):
        return float(np.mean(self._last_losses)) if self._last_losses else 0.0

    def paramstats(self):
        print("Parameter statistics")
        for k, W in self.loss_node.all_trainable_params.items():
            W = W.get_value()
            print("Param %s:\tshape=%s,\tmean=%f,\tstd=%f,\tmedian(abs)=%f"
                  % (k, W.shape, W.mean(), W.std(), np.median(np.abs(W))))

    def gradstats(self, *args, **kwargs):
        grads = self.gradients(*args, **kwargs)
        print("Gradient statistics")
        for g in grads:
            print("\tshape=%s,\tmean=%f,\tstd=%f,\tmedian(abs)=%f" % (g.shape, np.mean(g), np.std(g), np.median(np.abs(g))))

    # -- parameters ------------------------------------------------------------------
    def _all_named_params(self):
        out = OrderedDict()
        for n in self.nodes.values():
            for k, p in n.params.items():
                if p.apply_train:
                    out["%s_%s" % (n.name, k)] = p
        return out

    def _ensure_param_store(self, device):
        if self._store is None:
            from .executor import ParamStore
            dp = self.data_parallel
            alloc = dp.alloc_gradient_buffer if (dp is not None and getattr(dp, 'use_ce', False)) else None
            self._store = ParamStore(self._all_named_params().items(), device, grad_alloc=alloc)
        return self._store

    def get_param_values(self, skip_const=True, as_list=False):
        if as_list:
            return [n.get_param_values(skip_const) for n in self.nodes.values()]
        return OrderedDict((name, n.get_param_values(skip_const)) for name, n in self.nodes.items())

    def set_param_values(self, value_dict, skip_const=True):
        for k, v in value_dict.items():
            self.nodes[k].set_param_values(v, skip_const)

    # -- execution -------------------------------------------------------------------
    def _train_plan(self, batch):
        from .executor import Plan
        if batch not in self._train_plans:
            outs = [self.loss_node] + [n for n in self.prediction_ext if n is not self.loss_node]
            self._train_plans[batch] = Plan(self, outs, batch, train=True)
        return self._train_plans[batch]

    def _feed_dict(self, plan, args):
        srcs = [n for n in [self.input_node, self.target_node] if n is not None and n in plan.inputs]
        extra = [n for n in plan.inputs if n not in srcs]
        srcs += extra
        if len(args) != len(srcs):
            raise ValueError("expected %d input arrays %s, got %d" % (len(srcs), [n.name for n in srcs], len(args)))
        return dict(zip(srcs, args))

    def trainingstep(self, *args, **kwargs):
        """One optimiser iteration: ``trainingstep(data, target, optimiser='Adam')``
        -> (loss, time, extra) (model.py:548-600).  The returned loss is the one
        computed in the same pass as the gradients (before the update)."""
        opt_name = kwargs.get('optimiser', 'SGD')
        if opt_name not in self.optimisers:
            logger.warning("No optimiser '%s'. Falling back to SGD" % (opt_name,))
            opt_name = 'SGD'
        opt = self.optimisers[opt_name]
        t0 = time.time()
        plan = self._train_plan(np.shape(args[0])[0])
        plan.feed(self._feed_dict(plan, args))
        # [fwd, loss] graph -> loss D2H enqueued -> [bwd (+ all-reduce), update, weight re-pack] graph; the call returns
        # when the loss has arrived, i.e. while the backward pass is still running: the next call's host work and H2D
        # copies overlap with it (stream order keeps everything else sequential)
        plan.train_step(opt, loss_async=True)
        loss = np.float32(plan.loss_op.read_wait()[0])  # the only device->host sync of the step
        if kwargs.get('update_loss', False):
            loss = self.loss(*args)
        t = time.time() - t0
        opt.last_exec_time = t
        self.elapsed_time += t
        self._last_exec_times.append(t + 1e-10)
        self._last_losses.append(loss)
        self.iterations += 1
        return loss, t, None

    def loss(self, *args, **kwargs):
        return self.loss_node(*args)

    def gradients(self, *args, **kwargs):
        """d loss / d trainable_params, in ``self.trainable_params`` order (model.py:182-185)."""
        plan = self._train_plan(np.shape(args[0])[0])
        plan.feed(self._feed_dict(plan, args))
        plan.execute()
        return [p._grad.detach().cpu().numpy().copy() for p in self.trainable_params]

    def predict(self, *args, **kwargs):
        return self.prediction_node(*args)

    def predict_ext(self, *args, **kwargs):
        from .executor import Plan
        b = np.shape(args[0])[0]
        if b not in self._ext_plans:
            self._ext_plans[b] = Plan(self, self.prediction_ext, b)
        plan = self._ext_plans[b]
        return plan.run(self._feed_dict(plan, args))

    def predict_dense(self, raw_img, as_uint8=False, pad_raw=False):
        return self.prediction_node.predict_dense(raw_img, as_uint8=as_uint8, pad_raw=pad_raw)

    def test_run_prediction(self):
        return self.prediction_node.test_run()

    def measure_exectimes(self, n_samples=5, n_warmup=4, print_info=True):
        """OrderedDict node name -> execution time in ms (model.py:608-619); see ``Node.measure_exectime``."""
        return OrderedDict((name, node.measure_exectime(n_samples=n_samples, n_warmup=n_warmup, print_info=print_info))
                           for name, node in self.nodes.items())

    # -- persistence: the reference's .mdl format (model.py:229-235; graphmanager.py) -----
    def save(self, file_name):
        """Pickle ``(descriptors, desig_descr)`` exactly as the reference's ``Model.save`` does, so the file loads
        in a Theano install of ELEKTRONN2 and, through ``modelload``, here."""
        from . import graphmanager
        graphmanager.save_model(self, file_name)

    def serialise(self):
        from . import graphmanager
        return graphmanager.serialise(self)


def modelload(file_name, override_mfp_to_active=False, imposed_patch_size=None, imposed_batch_size=None, name=None,
              **model_load_kwargs):
    """Load a Model from a ``.mdl`` file written by ``Model.save`` -- the reference's or this package's
    (model.py:623-729).  ``model_load_kwargs`` (remove_bn, make_weights_constant, ...) concern features outside the
    B200 hot path and raise if set."""
    if any(model_load_kwargs.values()):
        raise NotImplementedError("modelload options %s are not on the B200 hot path" % sorted(model_load_kwargs))
    from . import graphmanager
    return graphmanager.load_model(file_name, override_mfp_to_active, imposed_patch_size, imposed_batch_size, name)


def params_from_model_file(file_name):
    """OrderedDict node name -> parameter values of a ``.mdl`` file, without building the model (model.py:897-911)."""
    from . import graphmanager
    logger.info("Extracting parameters from %s" % file_name)
    node_descr, _ = graphmanager._load_all(file_name)
    params = OrderedDict()
    for name, descr in node_descr.items():
        if isinstance(descr, (list, tuple)) and len(descr[1]):
            params[name] = descr[1]
    return params


def kernel_lists_from_node_descr(model):
    """(filter_shapes, pool_shapes, mfp) of the Conv nodes in graph order (model.py:871-895)."""
    f, p, m = [], [], []
    for n in model.nodes.values():
        if type(n) is Conv:
            f.append(list(n.filter_shape)), p.append(list(n.pool_shape)), m.append(bool(n.mfp))
    return f, p, m


def rebuild_model(model, override_mfp_to_active=False, imposed_patch_size=None, imposed_batch_size=None, name=None):
    """Re-create ``model`` with the same parameters but MFP switched on and / or another
    patch size: what ``modelload(..., override_mfp_to_active=True, imposed_patch_size=...)``
    does to a saved model (model.py:623-729): every node whose class is exactly ``Conv``
    gets ``mfp=True`` (graphmanager.py:183-185), a ``FragmentsToDense`` is injected in
    front of the prediction (Softmax) node (model.py:668-689), and the patch size is
    snapped to an MFP-valid one."""
    filters, pools, mfps = kernel_lists_from_node_descr(model)
    old_patch = model.input_node.shape.spatial_shape
    if override_mfp_to_active:
        mfps = [True] * len(mfps)
        if imposed_patch_size is None:
            imposed_patch_size = old_patch
    patch = None
    if imposed_patch_size is not None:
        if len(imposed_patch_size) != len(old_patch):
            raise ValueError("Dimensionality of patch size and imposed patchsize do not match.")
        patch = closest_valid_patch_size(filters, pools, imposed_patch_size, mfps)
        if list(patch) != list(imposed_patch_size):
            logger.info("patch_size %s changed to %s (size not possible)" % (list(imposed_patch_size), list(patch)))
    new = model_manager.newmodel(name or ("%s_rebuilt_%d" % (model.name, len(model_manager.models))))
    mapping = {}

    def remap(v):
        if isinstance(v, Node):
            return mapping[id(v)]
        if isinstance(v, (list, tuple)):
            return type(v)(remap(x) for x in v)
        return v

    pred_name = model.prediction_node.name if model.prediction_node is not None else None
    for nm_, (cls, args, kwargs) in model.node_descriptors.items():
        old = model.nodes[nm_]
        try:
            args = [remap(a) for a in args]
            kwargs = {k: remap(v) for k, v in kwargs.items()}
        except KeyError:
            continue  # depends on a node that did not survive the re-build
        kwargs['print_repr'] = False
        if cls is Input and old is model.input_node:
            sh = list(old.shape.shape)
            if patch is not None:
                for ax, s in zip(old.shape.spatial_axes, patch):
                    sh[ax] = int(s)
            if imposed_batch_size is not None:
                sh[old.shape.tag2index('b')] = imposed_batch_size
            elif override_mfp_to_active:
                sh[old.shape.tag2index('b')] = 1  # MFP needs raw batch 1 (neural.py:667-668)
            args = [sh] + list(args[1:])
        if cls is Conv and override_mfp_to_active:
            kwargs['mfp'] = True
            if len(args) > 6:
                args = list(args)
                args[6] = True
        for pk in ('w', 'b'):
            if pk in old.params and issubclass(cls, Conv):
                kwargs[pk] = old.params[pk].get_value()
        if issubclass(cls, Conv) and 'identity_init' in cls.__init__.__code__.co_varnames:
            kwargs['identity_init'] = False
        if override_mfp_to_active and nm_ == pred_name and cls is Softmax:
            args[0] = FragmentsToDense(args[0], print_repr=False)
        try:
            node = cls(*args, **kwargs)
        except ValueError:
            if old in (model.target_node, model.loss_node, model.error_node) or not _needed_for_prediction(model, old):
                continue  # training-only nodes need not survive an inference re-build
            raise
        mapping[id(old)] = node
    d = {}
    for k, v in model._desig_descr.items():
        if v is None:
            d[k] = None
        elif isinstance(v, (list, tuple)):
            vs = [mapping.get(id(model.nodes[x if isinstance(x, str) else x.name])) for x in v]
            d[k] = [x for x in vs if x is not None] or None
        else:
            d[k] = mapping.get(id(model.nodes[v if isinstance(v, str) else v.name]))
    if d.get('loss_node') is None:
        d['loss_node'] = d['target_node'] = d['error_node'] = None
        d['prediction_ext'] = None
    new.designate_nodes(**d)
    model_manager.togglemodel()
    return new


def _needed_for_prediction(model, node):
    return model.prediction_node is not None and any(n is node for n in model.prediction_node.ancestors())
