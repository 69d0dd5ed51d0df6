"""Conv / UpConv / Pool / Crop / FragmentsToDense / AutoMerge node classes.

Host-side mirror of the hot-path part of neuromancer/neural.py: constructor
signatures, shape / stride / fov / MFP arithmetic, parameter creation and error
behaviour follow the reference (cited per method); the numeric work each node stands
for is launched through libe2b200 by ``executor.Plan``.

Supported on the B200 path: tags 'b,f,z,x,y' ('dnn' axis order, neural.py:618-620) and
the 2-D / 1-D forms 'b,f,x,y' / 'b,f,x' (computations.py:337-362; executed as 3-D with
leading unit axes), conv_mode 'valid', every activation of apply_activation
(computations.py:57-134) except 'maxout' (see ``_check_supported``), batch
normalisation 'train' / 'predict' (neural.py:681-709), max / average / sum pooling.
The four BASELINE configs (relu / lin, no BN) run through the fused conv / pool
epilogues; BN, 'prelu' and 'abs' in training go through the unfused epilogue kernels
(csrc/e2_epilogue.cu).  Dropout, gradnet_mode, BN 'fadeout' and the 'theano' axis
order raise NotImplementedError at construction -- never silently compute something
different.
"""
import logging

import numpy as np

from .graphutils import TaggedShape
from .node_basic import Node, Concat
from .variables import VariableWeight, VariableParam
from .shapecalc import mfp_bookkeeping

logger = logging.getLogger('elektronn2log')

_ACTS = ('relu', 'lin', 'linear', 'tanh', 'sig', 'sigmoid', 'logistic', 'abs', 'soft+', 'elu', 'selu', 'prelu')


class NeuralLayer(Node):
    """Parameter handling shared by Conv-like nodes (neural.py:47-256)."""

    def _register_param(self, param, shape, name, init_kwargs=None, apply_train=True, apply_reg=False):
        pname = "<%s%s>" % (self.name, name)
        if param is None:
            p = VariableWeight(shape=shape, init_kwargs=init_kwargs, name=pname, apply_train=apply_train,
                               apply_reg=apply_reg)
        elif isinstance(param, np.ndarray):
            if tuple(param.shape) != tuple(shape):
                raise ValueError("Parameter %s: given array has shape %s, need %s" % (name, param.shape, tuple(shape)))
            p = VariableWeight(shape=shape, value=param, name=pname, apply_train=apply_train, apply_reg=apply_reg)
        elif isinstance(param, VariableParam):
            p = param  # shared parameter
        else:
            raise ValueError("Parameter %s must be either <np.ndarray>, a VariableParam or None "
                             "(to create new param)" % (name,))
        setattr(self, name, p)
        self.params[name] = p

    def _setup_params(self, w_sh, w, b, pool_shape, gamma=None, mean=None, std=None):
        """neural.py:146-239: glorot-normal weights (pool-aware fan), bias by activation ((f,2) for prelu: bias and
        slope, slope initialised to 1), batch-normalisation parameters."""
        w_init = dict(scale='glorot', mode='normal', pool=pool_shape, spatial_axes=self.spatial_axes)
        self._register_param(w, w_sh, 'w', init_kwargs=w_init, apply_train=True, apply_reg=True)
        b_sh = (self.n_f,)
        fov = float(np.prod([w_sh[i] for i in self.spatial_axes]))
        if self.activation_func == 'relu':
            b_init = dict(scale=1.0 / fov, mode='const')
        elif self.activation_func == 'sigmoid':
            b_init = dict(scale=0.5, mode='const')
        elif self.activation_func == 'prelu':
            b_init = dict(scale=1.0 / fov, mode='prelu')
            b_sh = (self.n_f, 2)
        else:
            b_init = dict(scale=1e-6, mode='fix-uni')
        self._register_param(b, b_sh, 'b', init_kwargs=b_init, apply_train=True, apply_reg=False)
        sh = (self.n_f,)
        bn = self.batch_normalisation
        if bn == 'train':   # neural.py:208-228
            self._register_param(gamma, sh, 'gamma', init_kwargs=dict(scale=1.0, mode='const'), apply_train=True,
                                 apply_reg=3.0)
            if mean is not None or std is not None:
                raise ValueError("Cannot pass mean and std for training, they are computed in the theano graph.")
            self._register_param(None, sh, 'mean', init_kwargs=dict(scale=0.0, mode='const'), apply_train=False)
            self._register_param(None, sh, 'std', init_kwargs=dict(scale=1.0, mode='const'), apply_train=False)
        elif bn == 'predict':   # neural.py:230-242
            self._register_param(gamma, sh, 'gamma', init_kwargs=dict(scale=1.0, mode='const'), apply_train=False)
            self._register_param(mean, sh, 'mean', init_kwargs=dict(scale=0.0, mode='const'), apply_train=False)
            self._register_param(std, sh, 'std', init_kwargs=dict(scale=1.0, mode='const'), apply_train=False)
        elif bn is not False:
            raise ValueError("Unknown value %s for batchnormalisation" % (bn,))

    @property
    def unfused_epilogue(self):
        """True when the bias / activation cannot ride in the conv or pool kernel's epilogue (BN scale and batch
        statistics, prelu's slope, abs' sign): the executor then runs conv -> pool -> e2_affine_act."""
        return bool(self.batch_normalisation) or self.activation_func in ('prelu', 'abs')


class Conv(NeuralLayer):
    """Convolutional layer with subsequent pooling (neural.py:500-859).

    Order of operations is the reference's: conv -> pool | MFP -> + bias ->
    activation (neural.py:662-712)."""

    def __init__(self, parent, n_f, filter_shape, pool_shape=None, conv_mode='valid', activation_func='relu',
                 mfp=False, batch_normalisation=False, dropout_rate=0, name="conv", print_repr=True, w=None, b=None,
                 gamma=None, mean=None, std=None, gradnet_mode=None, invalidate_fov=False):
        super(Conv, self).__init__(parent, name, print_repr)
        self._check_supported(conv_mode, activation_func, batch_normalisation, dropout_rate, gradnet_mode)
        self.n_f = n_f
        self.filter_shape = tuple(int(f) for f in filter_shape)
        self.conv_mode = conv_mode
        self.activation_func = activation_func
        self.batch_normalisation = batch_normalisation
        self.dropout_rate = None
        self.mfp = mfp
        self.invalidate_fov = invalidate_fov
        self.strides = parent.shape.strides
        self.mfp_offsets = parent.shape.mfp_offsets
        self.axis = parent.shape.tag2index('f')
        if pool_shape is None:
            pool_shape = tuple(1 for _ in filter_shape)
        self.pool_shape = tuple(int(p) for p in pool_shape)
        self.spatial_axes = parent.shape.spatial_axes
        conv_dim = len(self.spatial_axes)
        if conv_dim != len(self.filter_shape) or len(self.filter_shape) != len(self.pool_shape):
            raise ValueError("The filter_shape dimensionality (%i), the number of spatial dimensions in the input "
                             "(%i) and the dimensionality of pool_shape (%i) differ! Use filter size 1 on axes "
                             "which should not be convolved." % (len(self.filter_shape), conv_dim, len(self.pool_shape)))
        # neural.py:604-633: 1-D 'b,f,x', 2-D 'b,f,x,y' (any two spatial tags on axes 2,3), 3-D 'b,f,z,x,y' ('dnn')
        ok = {1: (3, [2]), 2: (4, [2, 3]), 3: (5, [2, 3, 4])}.get(conv_dim)
        if ok is None or len(parent.shape) != ok[0] or self.spatial_axes != ok[1]:
            raise NotImplementedError("Cannot convolve non-standard shapes / axis orders (got tags %s; the 'theano' "
                                      "order 'b,z,f,x,y' is not on the B200 path). Implement reshaping before conv "
                                      "and re-reshaping after!" % (parent.shape.tags,))
        self.axis_order = 'dnn' if conv_dim == 3 else None
        self.conv_dim = conv_dim
        n_in = parent.shape['f']
        self.w_sh = [n_f, n_in] + list(self.filter_shape)
        self._setup_params(self.w_sh, w, b, self.pool_shape, gamma, mean, std)

    @staticmethod
    def _check_supported(conv_mode, activation_func, batch_normalisation, dropout_rate, gradnet_mode):
        if conv_mode != 'valid':
            raise NotImplementedError("conv_mode '%s': only 'valid' is supported on the B200 path "
                                      "(3-D 'same'/'full' exist only on the reference's cuDNN path)" % conv_mode)
        if isinstance(activation_func, str) and activation_func.startswith('maxout'):
            # the reference cannot build this either: Conv._make_output does ``self.filter_shape /= r``
            # (neural.py:650-653), a TypeError for the tuple / list filter shapes every caller passes, and
            # computations.maxout pools axis 2 (z in the 'b,f,z,x,y' layout, computations.py:476-477) instead of f
            raise NotImplementedError("'%s' as a layer activation is broken in the reference (neural.py:650-653); "
                                      "use computations.maxout(x, factor, axis) on the layer output" % activation_func)
        if activation_func not in _ACTS:
            raise NotImplementedError("%s. Permitted activation_funcs on the B200 path: %s"
                                      % (activation_func, list(_ACTS)))
        if batch_normalisation == 'fadeout':
            raise NotImplementedError("batch_normalisation='fadeout' needs gradnet_mode, which is not on the B200 path")
        if batch_normalisation not in (False, 'train', 'predict'):
            raise ValueError("Unknown value %s for batchnormalisation" % (batch_normalisation,))
        if dropout_rate:
            raise NotImplementedError("dropout is not supported on the B200 path")
        if gradnet_mode:
            raise NotImplementedError("gradnet_mode is not supported on the B200 path")

    def _calc_shape(self):
        """neural.py:725-764."""
        if self.mfp and self.input_nodes[0].shape['b'] not in (1, None):
            raise ValueError("For MFP the batchsize of the raw image input must be 1.")
        sh = self.parent.shape
        for j, (i, f, p) in enumerate(zip(self.spatial_axes, self.filter_shape, self.pool_shape)):
            k = 1 - f
            s = (sh[i] + k) // p
            if self.mfp:
                if (sh[i] + k - p + 1) % p != 0:
                    raise ValueError("Cannot pool spatial axis '%s' of length %i by factor %i after convolving "
                                     "with kernel of size %i and using MFP." % (sh.tags[i], sh[i], p, f))
            elif (sh[i] + k) % p != 0:
                raise ValueError("Cannot pool spatial axis '%s' of length %i by factor %i after convolving with "
                                 "kernel of size %i." % (sh.tags[i], sh[i], p, f))
            if s < 1:
                raise ValueError("Spatial axis '%s' of length %i is too short for kernel %i / pool %i"
                                 % (sh.tags[i], sh[i], f, p))
            sh = sh.updateshape(i, s)
            fov = sh.fov[j] + (f + p - 2) * sh.strides[j] if (sh.fov[j] > 0 and not self.invalidate_fov) else -1
            sh = sh.updatefov(j, fov)
        if self.mfp:
            self.mfp_offsets, self.strides = mfp_bookkeeping(self.pool_shape, self.mfp_offsets, self.strides)
            sh = sh.updatemfp_offsets(self.mfp_offsets)
            if sh['b'] is None:
                sh = sh.updateshape('b', 1)
            sh = sh.updateshape('b', int(np.prod(self.pool_shape)), mode='mult')
        else:
            self.strides = np.multiply(self.pool_shape, self.strides)
        sh = sh.updatestrides(self.strides)
        self.shape = sh.updateshape('f', self.n_f)

    def _calc_comp_cost(self):
        """MACs: prod(w_sh) * n_positions * b (neural.py:767-778)."""
        sh = self.parent.shape
        n_position = 1
        for i, f in zip(self.spatial_axes, self.filter_shape):
            n_position *= sh[i] + 1 - f
        b = 1 if sh['b'] is None else sh['b']
        self.computational_cost = int(np.prod(self.w_sh)) * int(n_position) * int(b)

    def __repr__(self):
        s = super(Conv, self).__repr__() + "\n"
        s += "  n_f=%i, %id conv, kernel=%s, pool=%s, act='%s', " % (
            self.n_f, self.conv_dim, self.filter_shape, self.pool_shape, self.activation_func)
        if self.mfp:
            s += "MFP active, "
        return s


class FragmentsToDense(Node):
    """neural.py:863-901."""

    def __init__(self, parent, name="to_dense", print_repr=True):
        super(FragmentsToDense, self).__init__(parent, name, print_repr)

    def _calc_shape(self):
        sh = self.parent.shape
        if sh['b'] != len(sh.mfp_offsets) or sh['b'] != np.prod(sh.strides):
            raise ValueError("Need %i fragments on the batch axis. Is MFP active at all?" % np.prod(sh.strides))
        for ax, st in zip(sh.spatial_axes, sh.strides):
            sh = sh.updateshape(ax, int(st), mode='mult')
        sh = sh.updateshape('b', 1)
        n = len(sh.spatial_axes)
        self.shape = TaggedShape(sh.shape, sh.tags, np.ones(n, np.int64), np.zeros((1, n), np.int64), sh.fov)

    def _calc_comp_cost(self):
        self.computational_cost = 0                                  # neural.py:895-901


class UpConv(Conv):
    """Transposed convolution with kernel == pool == stride (neural.py:907-1129)."""

    def __init__(self, parent, n_f, pool_shape, activation_func='relu', identity_init=True,
                 batch_normalisation=False, dropout_rate=0, name="upconv", print_repr=True, w=None, b=None,
                 gamma=None, mean=None, std=None, gradnet_mode=None):
        pool_shape = tuple(int(p) for p in pool_shape)
        super(UpConv, self).__init__(parent, n_f, pool_shape, pool_shape, 'valid', activation_func, mfp=False,
                                     batch_normalisation=batch_normalisation, dropout_rate=dropout_rate, name=name,
                                     print_repr=print_repr, w=w, b=b, gamma=gamma, mean=mean, std=std,
                                     gradnet_mode=gradnet_mode)
        if identity_init:  # neural.py:977-986
            w_val = self.w.get_value() * 0.1
            s = np.arange(min(w_val.shape[0], w_val.shape[1]))
            w_val[s, s] = 1.0
            self.w.set_value(w_val)
            self.b.set_value(self.b.get_value() * 0.0)   # prelu: zeroes the slope too, like the reference (:984)

    def _calc_shape(self):
        """neural.py:1074-1097: S*p per axis, strides / p, fov flagged -1."""
        self.strides = np.divide(self.strides, self.pool_shape).astype(np.int64)
        sh = self.parent.shape
        for j, (i, f, p) in enumerate(zip(self.spatial_axes, self.filter_shape, self.pool_shape)):
            sh = sh.updateshape(i, sh[i] * p + p - 1 + (1 - f))
            sh = sh.updatefov(j, -1)
        sh = sh.updateshape('f', self.n_f)
        self.shape = sh.updatestrides(self.strides)

    def _calc_comp_cost(self):
        """neural.py:1100-1111."""
        sh = self.parent.shape
        n_position = 1
        for i, f, p in zip(self.spatial_axes, self.filter_shape, self.pool_shape):
            n_position *= sh[i] * p + 1 - f
        b = 1 if sh['b'] is None else sh['b']
        self.computational_cost = int(np.prod(self.w_sh)) * int(n_position) * int(b)

    def __repr__(self):
        return Node.__repr__(self) + "\n  n_f=%i, %id upconv, kernel=%s, pool=%s, act='%s', " % (
            self.n_f, self.conv_dim, self.filter_shape, self.pool_shape, self.activation_func)


class Crop(Node):
    """Symmetric spatial crop (neural.py:1132-1190)."""

    def __init__(self, parent, crop, name="crop", print_repr=True):
        super(Crop, self).__init__(parent, name, print_repr)
        self.crop = tuple(int(c) for c in crop)

    def _calc_shape(self):
        sh = self.parent.shape.copy()
        for k, i in enumerate(self.parent.shape.spatial_axes):
            s = self.parent.shape[i] - 2 * self.crop[k]
            if s < 1:
                raise ValueError("Crop %s leaves nothing of axis %s" % (self.crop, sh.tags[i]))
            sh = sh.updateshape(i, s)
        self.shape = sh

    def _calc_comp_cost(self):
        self.computational_cost = 0                                  # neural.py:1184-1190


def AutoMerge(parent1, parent2, upconv_n_f=None, merge_mode='concat', disable_upconv=False, upconv_kwargs=None,
              name='merge', print_repr=True):
    """Align a high-res and a low-res branch with UpConv + Crop and concatenate them
    (neural.py:1282-1405).  Concat order: (lo_res, hi_res)."""
    assert len(parent1.shape) == len(parent2.shape)
    assert parent1.shape.spatial_axes == parent2.shape.spatial_axes
    if any(parent2.shape.strides // parent1.shape.strides < 1):
        lo_res, hi_res = parent1, parent2
    else:
        hi_res, lo_res = parent1, parent2
    unpool = lo_res.shape.strides // hi_res.shape.strides
    if np.any(unpool > 1) and not disable_upconv:
        if upconv_n_f is None:
            raise ValueError('AutoMerge is trying to insert an UpConv node, but upconv_n_f is not defined. '
                             'Please set it to the desired number of features to be used for UpConv.')
        lo_res = UpConv(lo_res, upconv_n_f, tuple(int(u) for u in unpool), **(upconv_kwargs or {}))
    sh_hi, sh_lo = hi_res.shape.spatial_shape, lo_res.shape.spatial_shape
    crop_lo, crop_hi = [], []
    for a, b in zip(sh_hi, sh_lo):
        diff = a - b
        if diff % 2 != 0:
            raise ValueError("hi_res and lo_res maps cannot be aligned with shapes:\n%s\n%s" % (sh_hi, sh_lo))
        crop_hi.append(diff // 2 if diff > 0 else 0)
        crop_lo.append(-diff // 2 if diff < 0 else 0)
    if np.any(crop_lo):
        lo_res = Crop(lo_res, crop_lo, print_repr=print_repr)
    if np.any(crop_hi):
        hi_res = Crop(hi_res, crop_hi, print_repr=print_repr)
    if merge_mode == 'concat':
        return Concat((lo_res, hi_res), axis='f', name=name, print_repr=print_repr)
    if merge_mode == 'add':
        raise NotImplementedError("merge_mode='add' is not on the B200 path (no BASELINE config uses it)")
    raise ValueError('Invalid "merge_mode". Should be "add" or "concat".')


UpConvMerge = AutoMerge


class Pool(Node):
    """Max-pooling layer (neural.py:1409-1559)."""

    def __init__(self, parent, pool_shape, stride=None, mfp=False, mode='max', name="pool", print_repr=True):
        super(Pool, self).__init__(parent, name, print_repr)
        if mfp and stride is not None:
            raise ValueError("Cannot use custom stride and MFP together")
        if mfp:
            raise NotImplementedError("Check this first before use")  # neural.py:1537 (dead in the reference too)
        pool_shape = tuple(int(p) for p in pool_shape)
        if stride is not None and tuple(stride) != pool_shape:
            raise NotImplementedError("Stride!=Pool is not supported (neither by the reference's CPU path, "
                                      "computations.py:612-613)")
        if mode == 'average':
            mode = 'average_inc_pad'  # Theano's internal name (neural.py:1451-1452)
        if mode not in ('max', 'average_inc_pad', 'average_exc_pad', 'sum'):
            raise ValueError("Unknown pooling mode '%s' (computations.py:556-561)" % mode)
        self.pool_shape = pool_shape
        self.pool_stride = pool_shape
        self.mfp = mfp
        self.mode = mode
        self.strides = parent.shape.strides
        self.mfp_offsets = parent.shape.mfp_offsets
        self.spatial_axes = parent.shape.spatial_axes
        nd = len(pool_shape)
        if nd not in (1, 2, 3) or len(parent.shape) != nd + 2 or self.spatial_axes != list(range(2, 2 + nd)):
            raise NotImplementedError("The B200 path pools 'b,f,x' / 'b,f,x,y' / 'b,f,z,x,y' tensors over all their "
                                      "spatial axes (got pool %s for tags %s)" % (pool_shape, parent.shape.tags))

    def _calc_shape(self):
        """neural.py:1528-1559."""
        sh = self.parent.shape
        for j, (i, p) in enumerate(zip(self.spatial_axes, self.pool_shape)):
            if sh[i] % p != 0:
                raise ValueError("Cannot downsample spatial axis '%s' of length %i by factor %i with pool %i."
                                 % (sh.tags[i], sh[i], p, p))
            new = sh[i] // p
            fov = sh.fov[j] + (p - 1) * sh.strides[j] if sh.fov[j] > 0 else -1
            sh = sh.updateshape(i, new).updatefov(j, fov)
        self.strides = np.multiply(self.pool_stride, self.strides)
        self.shape = sh.updatestrides(self.strides)
