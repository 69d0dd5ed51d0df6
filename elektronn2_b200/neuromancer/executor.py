"""Static launch plans over libe2b200.

The reference compiles a Theano function per requested output the first time it is
called (graphutils.make_func, graphutils.py:314-387) and derives the backward pass
symbolically (T.grad, model.py:182).  The B200 replacement is a *plan*: all
activation / gradient buffers are allocated once in HBM, every op is bound to its
C-ABI descriptor once, and a step is a fixed list of launches -- captured into a
CUDA graph so the host cost per step is one graph launch instead of ~200 ctypes
calls.

Data layout in HBM:
  activations / gradients : NDHWC float32, channel pitch multiple of 4 floats
  Concat                  : one buffer; UpConv writes its channels in place, the
                            cropped skip branch is copied into the remaining channels
  parameters / gradients  : two flat float32 buffers [weights | biases], laid out in
                            reverse graph order so the backward pass completes them
                            front to back (bucketed all-reduce overlaps with wgrad)
"""
import contextlib
import ctypes as C
import gc
import os

import numpy as np
import torch

from .. import _lib
from .. import ops
from ..config import config
from ..devtensor import DevTensor
from . import loss as loss_nodes
from .neural import Conv, UpConv, Pool, Crop, FragmentsToDense
from .node_basic import Input, Concat


@contextlib.contextmanager
def _capture(graph, stream=None):
    """``torch.cuda.graph`` with the cyclic garbage collector off.  A collection in the middle of a capture can run
    finalisers of objects from earlier plans (CUDA graphs, handles, pinned buffers) whose CUDA calls invalidate the
    capture -- seen as an intermittent "operation failed due to a previous error during capture"."""
    gc.collect()
    was = gc.isenabled()
    gc.disable()
    try:
        kw = dict(capture_error_mode='thread_local')
        if stream is not None:
            kw['stream'] = stream
        with torch.cuda.graph(graph, **kw):
            yield
    finally:
        if was:
            gc.enable()


def _try_capture(fn, device, what):
    """Capture ``fn`` into a CUDA graph; None (with a warning) if the capture is refused or invalidated."""
    g = torch.cuda.CUDAGraph()
    try:
        with _capture(g):
            fn()
        return g
    except Exception as e:          # noqa: BLE001
        import warnings
        warnings.warn('CUDA-graph capture of %s failed (%s); running eagerly' % (what, e))
        try:
            torch.cuda.synchronize(device)
        except Exception:           # noqa: BLE001
            torch.cuda.synchronize(device)
        return None


class ParamStore(object):
    """Flat device buffers for all trainable parameters of a model."""

    def __init__(self, named_params, device, grad_alloc=None):
        # reverse graph order inside each region; weights (regularised) first, then the rest
        plist = list(named_params)[::-1]
        seen = {}
        for k, p in plist:
            if id(p) in seen:
                # the reference shares weights by passing one VariableParam to several nodes; the flat buffer gives
                # every entry its own region and each wgrad kernel OVERWRITES its gradient view, so a shared
                # parameter would silently lose contributions
                raise NotImplementedError("parameter %s is shared by '%s' and '%s': shared parameters are not "
                                          "supported on the B200 path" % (p.name, seen[id(p)], k))
            seen[id(p)] = k
        # grouped by weight-decay multiplier (VariableParam.apply_reg: False/0, True/1, or a factor > 1 such as the
        # 3.0 of batch-norm gammas, neural.py:213; optimiser.py:150-153, 312-315): the plain weights first -- that
        # region [0, n_reg) is what the backward pass completes front to back and the all-reduce buckets cover --
        # then any other multipliers, then the unregularised rest (biases)
        def mult(p):
            r = p.apply_reg
            if isinstance(r, (bool, np.bool_)):
                return 1.0 if r else 0.0
            r = float(r)
            if r < 0 or (0 < r < 1):
                raise ValueError("parameter %s: apply_reg=%r is neither a flag nor a multiplier >= 1" % (p.name, p.apply_reg))
            return r
        mults = sorted({mult(p) for _, p in plist} - {0.0, 1.0})
        groups = [(m, [(k, p) for k, p in plist if mult(p) == m]) for m in [1.0] + mults + [0.0]]
        self.entries = []  # (key, param, offset, size)
        self.regions = []  # (offset, count, weight-decay multiplier)
        off = 0
        self.n_reg = 0
        for m, group in groups:
            start = off
            for k, p in group:
                size = int(np.prod(p.shape))
                self.entries.append((k, p, off, size))
                off += (size + 3) // 4 * 4
            if m == 1.0:
                self.n_reg = off
            if off > start:
                self.regions.append((start, off - start, m))
        self.total = off
        self.device = device
        self.P = torch.zeros(self.total, dtype=torch.float32, device=device)
        # grad_alloc: data parallelism may want the gradient buffer in peer-mapped (symmetric) memory
        self.G = grad_alloc(self.total, device) if grad_alloc is not None else \
            torch.zeros(self.total, dtype=torch.float32, device=device)
        self.version = 0
        for k, p, o, size in self.entries:
            view = self.P[o:o + size].view(*p.shape)
            view.copy_(torch.from_numpy(np.ascontiguousarray(p._host)))
            p._dev = view
            p._grad = self.G[o:o + size].view(*p.shape)
            p._offset = o
            p._on_set.append(self._bump)

    def _bump(self):
        self.version += 1


class Launch(object):
    """One entry of a plan: a bound C-ABI call plus its algorithmic work (for rooflines:
    SURVEY.md 8d "Algorithmic work")."""
    __slots__ = ('label', 'fn', 'flops', 'bytes', 'kind', 'param_end', 'param_start')

    def __init__(self, label, fn, flops=0, nbytes=0, kind='hbm'):
        self.label, self.fn, self.flops, self.bytes, self.kind = label, fn, float(flops), float(nbytes), kind
        self.param_end = None  # wgrad launches: end offset (floats) of the gradient region now complete
        self.param_start = None

    def __call__(self):
        self.fn()


def _nel(t):
    return t.desc.positions * t.desc.c


def _sp3(v, fill=1):
    """Spatial tuple of a 1-D / 2-D node as the 3-D (z,x,y) the library works on: leading unit axes
    (computations.py:337-362 are the kz == 1 / kz == kx == 1 special cases of the 3-D op)."""
    v = [int(a) for a in v]
    return [fill] * (3 - len(v)) + v


def _aux_dev(p, device):
    """Device copy of a parameter that is not part of the trainable flat buffer (batch-norm mean / std, 'predict'
    gamma).  It lives on the parameter, so every plan of the model (training, prediction) sees the same values and
    get_value() / set_value() go through it."""
    if p._dev is None:
        p._dev = torch.from_numpy(np.ascontiguousarray(p._host, dtype=np.float32)).to(device)
    return p._dev


def _nb(node, batch):
    b = node.shape['b']
    return int(batch) if b is None else int(b)


def _tensor_shaped(c_in, c_out):
    """Layers with c_in == 1 or c_out <= 2 are bandwidth-bound and are
    reported against the HBM roofline, everything else against the tensor pipe."""
    return c_in > 1 and c_out > 2


class Plan(object):
    """Forward (and optionally backward) launch plan for a set of output nodes."""

    def __init__(self, model, outputs, batch, train=False, compute=None, use_graph=None):
        if not torch.cuda.is_available():
            raise RuntimeError("elektronn2_b200 needs a CUDA device; there is no CPU fallback")
        self.model, self.outputs, self.batch, self.train = model, list(outputs), int(batch), train
        self.compute = compute or config.compute
        self.use_graph = config.use_cuda_graph if use_graph is None else use_graph
        self.h = _lib.get_handle()
        self.device = torch.device('cuda', self.h.device)
        self.store = model._ensure_param_store(self.device)
        needed = {}
        for o in self.outputs:
            for n in o.ancestors():
                needed[id(n)] = n
        self.nodes = [n for n in model.nodes.values() if id(n) in needed]
        self._in_plan = set(needed.keys())
        self.val, self.aux = {}, {}
        self.fwd_ops, self.bwd_ops, self.pack_ops = [], [], []
        self.conv_ops = {}
        self.bn_ops = []      # batch-norm 'train' epilogues: running averages move only inside an optimiser step
        self.loss_op = None
        self.inputs = {}
        self._graph = None
        self._graph_upd = None
        self._pack_stream = None
        self._stage_in, self._pending, self._fed = {}, {}, False
        self._opt_stream = None
        self._capture_stream = None
        self._opt_graphs = {}
        self._split = None
        self.fuse_optimiser = os.environ.get('E2_FUSE_OPT', '1') != '0'
        self._wgrad_stream = None
        self.wgrad_overlap = self.train and os.environ.get('E2_WGRAD_OVERLAP', '1') != '0'
        self.wgrad_streams = max(1, int(os.environ.get('E2_WGRAD_STREAMS', '2')))
        self._pack_graph = None
        self.overlap_pack = os.environ.get('E2_PACK_OVERLAP', '0') == '1'
        self._copy_stream = None
        self._inputs_free = None
        self._live_sources = {}
        self._packed_version = -1
        self.grad_scale = 1.0
        dp = model.data_parallel
        if train and dp is not None:
            self.grad_scale = dp.grad_scale()
        self._build_forward()
        if train:
            self._build_backward()

    # ------------------------------------------------------------------ helpers
    def _consumers(self, n):
        return [c for c in n.children.values() if id(c) in self._in_plan]

    def _new(self, node, c=None):
        sp = _sp3(node.shape.spatial_shape)
        return DevTensor(_nb(node, self.batch), sp[0], sp[1], sp[2], node.shape['f'] if c is None else c,
                         device=self.device)

    def _pdev(self, p):
        if p._dev is None:
            raise RuntimeError("parameter %s is not part of the model's parameter store" % p.name)
        return p._dev

    def _f(self, label, fn, flops=0, nbytes=0, kind='hbm'):
        self.fwd_ops.append(Launch(label, fn, flops, nbytes, kind))

    def _b(self, label, fn, flops=0, nbytes=0, kind='hbm'):
        l = Launch(label, fn, flops, nbytes, kind)
        self.bwd_ops.append(l)
        return l

    # ------------------------------------------------------------------ forward
    def _build_forward(self):
        h = self.h
        out_ids = set(id(o) for o in self.outputs)
        self.crop_into, self.copy_into, self.alias_of = {}, {}, {}
        self._fused_pools = set()    # Pool nodes whose forward runs in their producer's conv epilogue
        for n in self.nodes:
            if isinstance(n, Concat):
                buf = self._new(n)
                self.val[n] = buf
                c0 = 0
                for p in n.parents:
                    c = p.shape['f']
                    sole = len(self._consumers(p)) == 1 and id(p) not in out_ids
                    if sole and isinstance(p, UpConv):
                        self.val[p] = buf.channel_slice(c0, c)
                        self.alias_of[p] = (n, c0)
                    elif sole and isinstance(p, Crop):
                        self.crop_into[p] = (n, c0)
                    else:
                        self.copy_into.setdefault(n, []).append((p, c0))
                    c0 += c
        loss_group = (loss_nodes.Softmax, loss_nodes.MultinoulliNLL, loss_nodes.AggregateLoss,
                      loss_nodes.Classification, loss_nodes._Errors)
        for n in self.nodes:
            if isinstance(n, loss_group):
                continue
            if isinstance(n, Input):
                self._plan_input(n)
            elif isinstance(n, UpConv):
                x = self.val[n.parent]
                y = self.val[n] if n in self.val else self._new(n)
                self.val[n] = y
                fl = 2.0 * _nel(x) * n.n_f * int(np.prod(n.pool_shape))
                kind = 'tensor' if _tensor_shaped(x.desc.c, n.n_f) else 'hbm'
                if self._unfused(n):
                    d = y.desc
                    lin = DevTensor(d.n, d.z, d.x, d.y, d.c, c_pitch=d.c_pitch, device=self.device)
                    op = ops.UpConvOp(h, x, lin, self._pdev(n.w), None, _sp3(n.pool_shape), 'lin', self.compute)
                    self._f('upconv_fwd:' + n.name, op.fwd, fl, 4 * (_nel(x) + _nel(lin)), kind)
                    self._plan_affine(n, lin, None, lin, y)
                else:
                    op = ops.UpConvOp(h, x, y, self._pdev(n.w), self._pdev(n.b), _sp3(n.pool_shape), n.activation_func,
                                      self.compute)
                    self._f('upconv_fwd:' + n.name, op.fwd, fl, 4 * (_nel(x) + _nel(y)), kind)
                self.conv_ops[n] = op
                self.pack_ops.append(op)
            elif isinstance(n, Conv):
                self._plan_conv(n)
            elif isinstance(n, Pool):
                if n in self._fused_pools:
                    continue      # planned together with its producer (_plan_conv)
                x = self.val[n.parent]
                y = self._new(n)
                self.val[n] = y
                op = ops.PoolOp(h, x, y, _sp3(n.pool_shape), keep_argmax=self.train, tie_mode=config.pool_tie_mode,
                                mode=n.mode)
                self.aux[n] = op
                self._f('pool_fwd:' + n.name, op.fwd, 0, 4 * (_nel(x) + _nel(y)))
            elif isinstance(n, Crop):
                src = self.val[n.parent]
                if n in self.crop_into:
                    cat, c0 = self.crop_into[n]
                    dst = self.val[cat]
                    self.val[n] = dst.channel_slice(c0, n.shape['f'])
                else:
                    dst, c0 = self._new(n), 0
                    self.val[n] = dst
                op = ops.CropConcatOp(h, src, dst, _sp3(n.crop, 0), c0)
                self.aux[n] = op
                self._f('crop_concat_fwd:' + n.name, op.fwd, 0, 8 * _nel(self.val[n]))
            elif isinstance(n, Concat):
                cops = []
                for p, c0 in self.copy_into.get(n, []):
                    op = ops.CropConcatOp(h, self.val[p], self.val[n], (0, 0, 0), c0)
                    cops.append((p, op))
                    self._f('concat_copy:' + n.name, op.fwd, 0, 8 * _nel(self.val[p]))
                self.aux[n] = cops
            elif isinstance(n, FragmentsToDense):
                x = self.val[n.parent]
                y = self._new(n)
                self.val[n] = y
                op = ops.Frag2DenseOp(h, x, y, [_sp3(o, 0) for o in n.parent.shape.mfp_offsets],
                                      _sp3(n.parent.shape.strides))
                self.aux[n] = op
                self._f('frag2dense_fwd:' + n.name, op.fwd, 0, 8 * _nel(x))
            else:
                raise NotImplementedError("Node type %s is not on the B200 hot path" % type(n).__name__)
        self._plan_loss_head()

    def _plan_input(self, n):
        t = self._new(n)
        self.val[n] = t
        shape = (t.desc.n, t.desc.c, t.desc.z, t.desc.x, t.desc.y)
        pinned = torch.empty(shape, dtype=torch.float32).pin_memory()
        if t.desc.c == 1:
            staging = None  # (b,1,z,x,y) is already channels-last
        else:
            staging = torch.empty(shape, dtype=torch.float32, device=self.device)
            self._f('layout_in:' + n.name,
                    lambda t=t, s=staging: self.h.call('e2_ncdhw_to_ndhwc', C.byref(t.desc), _lib.ptr(s), t.ptr(),
                                                       self.h.stream()), 0, 8 * _nel(t))
        self.inputs[n] = (t, pinned, staging)

    def _plan_conv(self, n):
        h = self.h
        x = self.val[n.parent]
        if (x.desc.c == 1 and x.desc.c_pitch == 1 and self.compute == 'tf32' and n.n_f >= 8
                and os.environ.get('E2_C1_TC')):
            # dense single-channel input -> 16-byte rows (tf32-rounded) so the first layer runs on the
            # TMA / tcgen05 path; the copy is ~1 % of the layer's output traffic
            key = ('repitch', n.parent)
            if key not in self.aux:
                d = x.desc
                x4 = DevTensor(d.n, d.z, d.x, d.y, 1, c_pitch=4, device=self.device)
                self.aux[key] = x4
                self._f('repitch:' + n.parent.name,
                        lambda x=x, x4=x4: h.call('e2_repitch', C.byref(x.desc), x.ptr(), C.byref(x4.desc), x4.ptr(), 1,
                                                  h.stream()), 0, 20 * _nel(x))
            x = self.aux[key]
        y = self.val[n] if n in self.val else self._new(n)
        self.val[n] = y
        k3, p3 = _sp3(n.filter_shape), _sp3(n.pool_shape)
        pooled = any(p > 1 for p in p3)
        w, b = self._pdev(n.w), self._pdev(n.b)
        lsp = [s + 1 - f for s, f in zip(_sp3(n.parent.shape.spatial_shape), k3)]
        flops = 2.0 * x.desc.n * int(np.prod(lsp)) * int(np.prod(n.w_sh))
        kind = 'tensor' if _tensor_shaped(x.desc.c, n.n_f) else 'hbm'
        rnd = self.compute == 'tf32'
        unfused = self._unfused(n)
        if not pooled and not unfused:
            op = ops.ConvOp(h, x, y, w, b, k3, n.activation_func, self.compute)
            pn = self._pool_consumer(n) if self._pool_fusion_pays(x, k3) else None
            pop = None
            if pn is not None:
                # Conv -> Pool (examples/unet3d.py:63-74): the window maximum is taken in the conv kernel's epilogue,
                # the unpooled tensor is still written (skip connection / ReLU gate) but not read back
                yp = self._new(pn)
                pop = ops.PoolOp(h, y, yp, _sp3(pn.pool_shape), keep_argmax=self.train, tie_mode=config.pool_tie_mode,
                                 mode='max')
                if op.pool_fusable(pop):
                    self.val[pn], self.aux[pn] = yp, pop
                    self._fused_pools.add(pn)
                else:
                    pop = None
            if pop is not None:
                keep = self._keep_window(n, pn)
                self._f('conv_fwd:' + n.name, lambda op=op, pop=pop, keep=keep: op.fwd_pool(pop, True, keep), flops,
                        4 * (_nel(x) + _nel(y) + 2 * _nel(pop.y)), kind)
            else:
                self._f('conv_fwd:' + n.name, op.fwd, flops, 4 * (_nel(x) + _nel(y)), kind)
        else:
            # conv -> pool|MFP -> +bias -> act  (neural.py:662-712): the conv writes raw
            # accumulators, the pooling kernel carries the bias/activation epilogue -- or, for batch
            # normalisation / prelu / abs, a third kernel does (e2_affine_act, csrc/e2_epilogue.cu)
            lin = DevTensor(x.desc.n, lsp[0], lsp[1], lsp[2], n.n_f, device=self.device)
            op = ops.ConvOp(h, x, lin, w, None, k3, 'lin', self.compute)
            self._f('conv_fwd:' + n.name, op.fwd, flops, 4 * (_nel(x) + _nel(lin)), kind)
            pb, pact = (None, 'lin') if unfused else (b, n.activation_func)
            v = y.like() if (unfused and pooled) else y
            pop = None
            if pooled and n.mfp:
                pop = ops.MfpOp(h, lin, v, p3, bias=pb, act=pact, keep_argmax=self.train, round_tf32=rnd)
                self._f('mfp_fwd:' + n.name, pop.fwd, 0, 4 * (_nel(lin) + _nel(v)))
            elif pooled:
                pop = ops.PoolOp(h, lin, v, p3, bias=pb, act=pact, keep_argmax=self.train,
                                 tie_mode=config.pool_tie_mode, round_tf32=rnd)
                if (not unfused and (not self.train or config.pool_tie_mode == 'first') and self._pool_fusion_pays(x, k3)
                        and op.pool_fusable(pop)):
                    # conv -> pool -> +bias -> act in ONE launch; the raw conv output is never written (the backward
                    # pass routes through the argmax)
                    self.fwd_ops.pop()
                    self._f('conv_fwd:' + n.name, lambda op=op, pop=pop: op.fwd_pool(pop, False), flops,
                            4 * (_nel(x) + 2 * _nel(v)), kind)
                else:
                    self._f('pool_fwd:' + n.name, pop.fwd, 0, 4 * (_nel(lin) + _nel(v)))
            if unfused:
                self._plan_affine(n, lin, pop, v if pooled else lin, y)
            else:
                self.aux[n] = (lin, pop)
        self.conv_ops[n] = op
        self.pack_ops.append(op)

    # reduction length (input channels x filter taps) from which the conv kernel's main loop is long enough to hide the
    # extra epilogue work of a fused max-pool.  Measured: the three Conv -> Pool pairs of unet3d (864 / 1728 / 3456) gain
    # 65 + 20 + 10 us per step; unet3d_litelite (180 ... 315) and the 20 -> 40 layer of neuro3d_lite (540) are epilogue
    # bound already and LOSE 2 % of their step with the pool fused
    POOL_FUSE_MIN_K = int(os.environ.get('E2_POOL_FUSE_MIN_K', '800'))

    def _pool_fusion_pays(self, x, k3):
        return x.desc.c * int(np.prod(k3)) >= self.POOL_FUSE_MIN_K

    def _pool_consumer(self, n):
        """The one max-Pool node (windows of 1 or 2 per axis) among the consumers of Conv ``n``, or None."""
        pools = [c for c in self._consumers(n) if isinstance(c, Pool) and c.parent is n]
        if len(pools) != 1 or pools[0].mode != 'max' or getattr(pools[0], 'mfp', False):
            return None
        p3 = _sp3(pools[0].pool_shape)
        if any(p not in (1, 2) for p in p3) or all(p == 1 for p in p3):
            return None
        return pools[0]

    def _keep_window(self, n, pool_node):
        """Window (z0,z1,x0,x1,y0,y1) of Conv ``n``'s unpooled output that anything but its fused Pool reads, or None
        for all of it.  In the U-Nets the only other reader is the skip connection's Crop (a few percent of the
        tensor): the pool backward gates with the pooled values, the crop backward reads the gate inside its window,
        and a ReLU layer whose gradient contributions are all gated needs no activation backward of its own."""
        if os.environ.get('E2_KEEP_FULL') or n in self.outputs or n.activation_func not in ('relu', 'lin', 'linear'):
            return None
        if self.train and config.pool_tie_mode != 'first':
            return None      # the Theano-CPU tie rule recomputes the window maximum from the unpooled tensor in the backward pass
        sp = _sp3(n.shape.spatial_shape)
        lo, hi = list(sp), [0, 0, 0]
        for c in self._consumers(n):
            if c is pool_node:
                continue
            if not isinstance(c, Crop) or c.parent is not n:
                return None
            cr = _sp3(c.crop, 0)
            for a in range(3):
                lo[a], hi[a] = min(lo[a], cr[a]), max(hi[a], sp[a] - cr[a])
        if any(h <= l for l, h in zip(lo, hi)):
            lo, hi = [0, 0, 0], [0, 0, 0]       # no other reader at all
        return (lo[0], hi[0], lo[1], hi[1], lo[2], hi[2])

    def _unfused(self, n):
        """Does this Conv / UpConv need the separate epilogue kernel?  Batch normalisation and prelu always; 'abs'
        only when a backward pass follows (its derivative needs the sign of the pre-activation, which the fused
        epilogue does not keep)."""
        return bool(n.batch_normalisation) or n.activation_func == 'prelu' or (n.activation_func == 'abs' and self.train)

    def _plan_affine(self, n, lin, pop, v, y):
        bn = n.batch_normalisation
        gamma = mean = std = None
        if bn:
            gamma = self._pdev(n.gamma) if n.gamma.apply_train else _aux_dev(n.gamma, self.device)
            mean, std = _aux_dev(n.mean, self.device), _aux_dev(n.std, self.device)
        aff = ops.AffineActOp(self.h, v, y, n.activation_func, self._pdev(n.b), bn, gamma, mean, std,
                              round_tf32=self.compute == 'tf32')
        self.aux[n] = dict(lin=lin, pop=pop, v=v, aff=aff)
        if bn == 'train':
            self.bn_ops.append(aff)
        self._f('affine_act_fwd:' + n.name, aff.fwd, 0, 4 * (2 * _nel(v) if bn != 'train' else 4 * _nel(v)))

    def _plan_loss_head(self):
        """Softmax (+ MultinoulliNLL + AggregateLoss + Errors) as one fused op."""
        sm = [n for n in self.nodes if isinstance(n, loss_nodes.Softmax)]
        if not sm:
            return
        if len(sm) > 1:
            raise NotImplementedError("more than one Softmax node in a plan")
        sm = sm[0]
        nll = [n for n in self.nodes if isinstance(n, loss_nodes.MultinoulliNLL)]
        logits = self.val[sm.parent]
        probs = self._new(sm)
        self.val[sm] = probs
        target = self.val[nll[0].target] if nll else None
        if nll and nll[0].pred is not sm:
            raise NotImplementedError("MultinoulliNLL must consume the planned Softmax node")
        self.loss_op = ops.LossOp(self.h, logits, target, probs)
        agg = [n for n in self.nodes if isinstance(n, loss_nodes.AggregateLoss)]
        if agg:
            # one component: loss = w * mean(nll) (loss.py:1355-1363).  The weight is read when the plan is built;
            # ``Model.mixing = ...`` drops the training plans so that the next step is planned with the new value
            self.loss_op.weight = float(np.asarray(agg[0].mixing_weights.get_value(), np.float64).reshape(-1)[0])
        self.logits_node = sm.parent
        self._f('softmax_nll_fwd', self.loss_op.fwd, 0, 4 * (2 * _nel(logits) + logits.desc.positions))

    # ----------------------------------------------------------------- backward
    def _build_backward(self):
        if self.loss_op is None or self.loss_op.target is None:
            raise ValueError("a training plan needs a Softmax -> MultinoulliNLL -> AggregateLoss head")
        h = self.h
        self.grad, written = {}, set()
        self._contrib = {}   # node -> [bool]: was each gradient contribution ReLU-gated by its producer?
        for n in self.nodes:
            if n in self.val and not isinstance(n, (Input, loss_nodes.Softmax)):
                if n in self.alias_of or n in self.crop_into:
                    continue
                self.grad[n] = self.val[n].like()
        for n, (cat, c0) in self.alias_of.items():
            self.grad[n] = self.grad[cat].channel_slice(c0, n.shape['f'])
        lg = self.logits_node
        self._b('softmax_nll_bwd', lambda: self.loss_op.bwd(self.grad[lg], self.grad_scale * self.loss_op.weight), 0,
                4 * 3 * _nel(self.grad[lg]))
        written.add(lg)
        # A Crop contributes to only part of its parent's gradient.  Emitted first it would have to
        # zero-fill the whole buffer and every later contribution would read-modify-write it; so crop
        # contributions wait until a full-coverage writer (pool / conv dgrad) has gone, then touch only
        # their own region.  They are flushed at the latest when the parent itself comes up.
        deferred = {}

        def flush(par):
            for emit in deferred.pop(par, []):
                emit(par in written)
                written.add(par)

        self._flush_deferred = flush
        for n in reversed(self.nodes):
            flush(n)
            if n not in written and n not in self.crop_into:
                continue
            if isinstance(n, UpConv):
                dy, op = self.grad[n], self.conv_ops[n]
                db = n.b._grad
                if isinstance(self.aux.get(n), dict):
                    dy, db = self._emit_affine_bwd(n, dy), None
                else:
                    self._emit_act_bwd(n, dy)
                fl = 2.0 * _nel(op.x) * n.n_f * int(np.prod(n.pool_shape))
                kind = 'tensor' if _tensor_shaped(op.x.desc.c, n.n_f) else 'hbm'
                l = self._b('upconv_wgrad:' + n.name, lambda op=op, dy=dy, n=n, db=db: op.wgrad(dy, n.w._grad, db),
                            fl, 4 * (_nel(op.x) + _nel(dy)), kind)
                l.param_end = n.w._offset + int(np.prod(n.w.shape))
                l.param_start = n.w._offset
                self._emit_dgrad(op, dy, n.parent, written, 'upconv_dgrad:' + n.name, fl, kind)
            elif isinstance(n, Conv):
                dy, op = self.grad[n], self.conv_ops[n]
                db = n.b._grad
                if isinstance(self.aux.get(n), dict):
                    a = self.aux[n]
                    dy, db = self._emit_affine_bwd(n, dy), None       # dy is now d loss / d (pooled) conv output
                    lin, pop = a['lin'], a['pop']
                else:
                    self._emit_act_bwd(n, dy)
                    lin, pop = self.aux.get(n, (None, None))
                if pop is not None:
                    dlin = lin.like()
                    self._b(('mfp_bwd:' if n.mfp else 'pool_bwd:') + n.name,
                            lambda pop=pop, dy=dy, dlin=dlin: pop.bwd(dy, dlin), 0, 4 * (2 * _nel(dy) + _nel(dlin)))
                else:
                    dlin = dy
                fl = 2.0 * dlin.desc.positions * int(np.prod(n.w_sh))
                kind = 'tensor' if _tensor_shaped(op.x.desc.c, n.n_f) else 'hbm'
                l = self._b('conv_wgrad:' + n.name, lambda op=op, dlin=dlin, n=n, db=db: op.wgrad(dlin, n.w._grad, db),
                            fl, 4 * (_nel(op.x) + _nel(dlin)), kind)
                l.param_end = n.w._offset + int(np.prod(n.w.shape))
                l.param_start = n.w._offset
                self._emit_dgrad(op, dlin, n.parent, written, 'conv_dgrad:' + n.name, fl, kind)
            elif isinstance(n, Pool):
                par = n.parent
                if isinstance(par, Input):
                    continue
                acc = par in written
                if n.mode == 'max':
                    gate = self._gate_for(par)
                else:   # average / sum pooling: no fused ReLU gate, the producer's activation backward runs on its own
                    gate = None
                    self._contrib.setdefault(par, []).append(False)
                self._b('pool_bwd:' + n.name,
                        lambda op=self.aux[n], dy=self.grad[n], dx=self.grad[par], acc=acc, gate=gate:
                        op.bwd(dy, dx, acc, relu_gate=gate), 0,
                        4 * (2 * _nel(self.grad[n]) + _nel(self.grad[par])))
                written.add(par)
                flush(par)
            elif isinstance(n, Crop):
                par = n.parent
                if isinstance(par, Input):
                    continue
                if n in self.crop_into:
                    cat, _ = self.crop_into[n]
                    if cat not in written:
                        continue
                    ddst = self.grad[cat]
                else:
                    ddst = self.grad[n]
                gate = self._gate_for(par)

                def emit(acc, n=n, par=par, ddst=ddst, gate=gate):
                    self._b('crop_concat_bwd:' + n.name,
                            lambda op=self.aux[n], ddst=ddst, dx=self.grad[par], acc=acc, gate=gate:
                            op.bwd(ddst, dx, acc, relu_gate=gate), 0,
                            4 * (3 * _nel(self.val[n]) if acc else _nel(self.val[n]) + _nel(self.grad[par])))
                if par in written:
                    emit(True)
                else:
                    deferred.setdefault(par, []).append(emit)
            elif isinstance(n, Concat):
                for p in n.parents:
                    if p in self.alias_of:
                        written.add(p)
                for p, op in self.aux.get(n, []):
                    if isinstance(p, Input):
                        continue
                    acc = p in written
                    self._contrib.setdefault(p, []).append(False)
                    self._b('concat_copy_bwd:' + n.name,
                            lambda op=op, ddst=self.grad[n], dx=self.grad[p], acc=acc: op.bwd(ddst, dx, acc), 0,
                            8 * _nel(self.grad[p]))
                    written.add(p)
            elif isinstance(n, FragmentsToDense):
                par = n.parent
                if par in written:
                    raise NotImplementedError("FragmentsToDense parent with several consumers")
                self._contrib.setdefault(par, []).append(False)
                self._b('frag2dense_bwd:' + n.name,
                        lambda op=self.aux[n], dd=self.grad[n], df=self.grad[par]: op.bwd(dd, df), 0,
                        8 * _nel(self.grad[par]))
                written.add(par)

    def _gate_for(self, parent):
        """Tensor to ReLU-gate a gradient contribution into ``grad[parent]`` with (fused ReLU
        backward), or None.  Also records whether the contribution is gated."""
        def relu_node(n):
            return isinstance(n, Conv) and n.activation_func == 'relu'
        gate = None
        if relu_node(parent):
            gate = self.val[parent]
        elif isinstance(parent, Concat) and parent not in self.copy_into and all(
                (q in self.alias_of and relu_node(q)) or (q in self.crop_into and relu_node(q.parent))
                for q in parent.parents):
            gate = self.val[parent]   # every channel of the buffer is a (copied) post-ReLU value
            for q in parent.parents:
                if q in self.alias_of:
                    self._contrib.setdefault(q, []).append(True)
        self._contrib.setdefault(parent, []).append(gate is not None)
        return gate

    def _emit_affine_bwd(self, n, dy):
        """Backward of the unfused epilogue: dy -> d loss / dv (v = the pooled conv output), bias / slope / gamma
        gradients written straight into the flat gradient buffer."""
        a = self.aux[n]
        aff, v = a['aff'], a['v']
        dv = DevTensor(v.desc.n, v.desc.z, v.desc.x, v.desc.y, v.desc.c, c_pitch=v.desc.c_pitch, device=self.device)
        dgamma = n.gamma._grad if (n.batch_normalisation and n.gamma.apply_train) else None
        self._b('affine_act_bwd:' + n.name,
                lambda aff=aff, dy=dy, dv=dv, n=n, dgamma=dgamma: aff.bwd(dy, dv, n.b._grad, dgamma), 0, 4 * 5 * _nel(v))
        return dv

    def _emit_act_bwd(self, n, dy):
        c = self._contrib.get(n)
        if n.activation_func == 'relu' and c and all(c):
            return  # every producer of this gradient already applied the ReLU gate
        if n.activation_func not in ('lin', 'linear'):
            h = self.h
            self._b('act_bwd:' + n.name,
                    lambda n=n, dy=dy: ops.act_bwd(h, dy, n.activation_func, self.val[n], dy, dy), 0, 12 * _nel(dy))

    def _emit_dgrad(self, op, dy, parent, written, label, flops, kind):
        if isinstance(parent, Input):
            return  # gradients are taken w.r.t. parameters only (model.py:182)
        acc = parent in written
        gate = self._gate_for(parent)
        self._b(label, lambda op=op, dy=dy, dx=self.grad[parent], acc=acc, gate=gate:
                op.dgrad(dy, dx, acc, relu_gate=gate), flops, 4 * (_nel(dy) + _nel(self.grad[parent])), kind)
        written.add(parent)
        self._flush_deferred(parent)

    # ---------------------------------------------------------------- execution
    def pack(self, need_dgrad=None):
        nd = self.train if need_dgrad is None else need_dgrad
        for op in self.pack_ops:
            op.pack(nd)
        self._packed_version = self.store.version

    def feed(self, values):
        """H2D copies of the inputs.  A source array that already lives in page-locked memory is copied from where
        it is; anything else goes through the plan's pinned staging buffer first.  The copies run on a side stream
        into device STAGING buffers (the input tensors themselves are still being read by the previous step's
        first-layer wgrad when ``Model.trainingstep`` has returned its loss), so they overlap with the backward pass
        still in flight; ``_commit_inputs`` moves them into place (device-to-device) when the next step is submitted.
        Returns the number of bytes copied."""
        nbytes = 0
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError("Plan.feed() inside a CUDA-graph capture")
        cur = torch.cuda.current_stream(self.device)
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        side = self._copy_stream
        if self._inputs_free is not None:
            side.wait_event(self._inputs_free)     # the staging buffers have been consumed by the previous step
        else:
            side.wait_stream(cur)
        with torch.cuda.stream(side):
            for n, a in values.items():
                t, pinned, staging = self.inputs[n]
                a = np.asarray(a, dtype=np.float32)
                if 3 <= a.ndim < 5 and a.ndim == len(n.shape):
                    a = a.reshape(a.shape[:2] + (1,) * (5 - a.ndim) + a.shape[2:])   # 1-D / 2-D nets: leading unit axes
                if tuple(a.shape) != tuple(pinned.shape):
                    raise ValueError("Input '%s' expects shape %s, got %s" % (n.name, tuple(pinned.shape), a.shape))
                src = None
                if a.flags['C_CONTIGUOUS'] and a.flags['WRITEABLE']:
                    cand = torch.from_numpy(a)
                    if cand.is_pinned():
                        src = cand
                        self._live_sources[n] = cand      # keep the caller's buffer alive until the copy has run
                if src is None:
                    pinned.copy_(torch.from_numpy(np.ascontiguousarray(a)))
                    src = pinned
                if staging is not None:
                    dst = staging                         # multi-channel: the layout kernel reads this buffer
                else:
                    dst = self._stage_in.get(n)
                    if dst is None:
                        dst = self._stage_in[n] = torch.empty(pinned.shape, dtype=torch.float32, device=self.device)
                    self._pending[n] = dst
                dst.copy_(src, non_blocking=True)
                nbytes += pinned.numel() * 4
        self._fed = True
        return nbytes

    def _commit_inputs(self):
        """Submit point of a step: the compute stream waits for the H2D copies and moves single-channel inputs from
        their staging buffers into the input tensors (a few microseconds of device-to-device copy)."""
        if not self._fed:
            return
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self._copy_stream)
        for n, stage in self._pending.items():
            t, pinned, _ = self.inputs[n]
            t.buf[t.offset:t.offset + pinned.numel()].view(pinned.shape).copy_(stage, non_blocking=True)
        self._pending = {}
        self._fed = False

    _WGRAD = ('conv_wgrad:', 'upconv_wgrad:')

    def _launch_all(self, hook=None):
        for f in self.fwd_ops:
            f()
        self._launch_bwd(hook)

    def _launch_bwd(self, hook=None, early=None):
        """Backward launches.  A weight gradient feeds nothing but the optimiser (and the all-reduce), so the wgrad
        kernels go to side streams, each behind an event that marks its place in the sequence, and run underneath
        the dgrad chain: the layers at the bottom of the U-Net have too few tiles to fill 148 SMs on their own.
        They use the handle's extra scratch buffers.  Joined before the step ends.
        ``early`` = (index into bwd_ops, fn): fn() is called on the optimiser stream once everything launched before
        that index is complete (fused optimiser step for the layers whose gradients are final by then)."""
        main = torch.cuda.current_stream(self.device)
        if not self.wgrad_overlap:
            for i, f in enumerate(self.bwd_ops):
                if early is not None and i == early[0]:
                    early[1]()
                f()
                if hook is not None and f.param_end is not None:
                    hook(f.param_end)
            return
        if self._wgrad_stream is None:
            self._wgrad_stream = [torch.cuda.Stream(device=self.device) for _ in range(self.wgrad_streams)]
        # fork every side stream here: an all-reduce bucket waits on events recorded on ALL of them, so they must
        # already be part of the step (and of a CUDA-graph capture) before the first bucket is launched
        ev0 = torch.cuda.Event()
        ev0.record(main)
        for side in self._wgrad_stream:
            side.wait_event(ev0)
        dp = self.model.data_parallel
        if dp is not None:
            dp.producer_streams = list(self._wgrad_stream)
        used = True
        k = 0
        for i, f in enumerate(self.bwd_ops):
            if early is not None and i == early[0]:
                if self._opt_stream is None:
                    self._opt_stream = torch.cuda.Stream(device=self.device)
                ev = torch.cuda.Event()
                ev.record(main)
                self._opt_stream.wait_event(ev)
                if used:
                    for side in self._wgrad_stream:
                        self._opt_stream.wait_stream(side)
                with torch.cuda.stream(self._opt_stream):
                    early[1]()
            if f.label.startswith(self._WGRAD):
                slot = k % len(self._wgrad_stream)
                side = self._wgrad_stream[slot]
                k += 1
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    self.h.ws_slot = 1 + slot
                    try:
                        f()
                        if hook is not None and f.param_end is not None:
                            hook(f.param_end)
                    finally:
                        self.h.ws_slot = 0
                used = True
            else:
                f()
                if hook is not None and f.param_end is not None:
                    hook(f.param_end)
        if used:
            for side in self._wgrad_stream:
                main.wait_stream(side)
        if early is not None and self._opt_stream is not None:
            main.wait_stream(self._opt_stream)

    # ------------------------------------------------------------- training step with the optimiser inside
    # layers (from the input side) whose update waits for the end of the backward pass
    TAIL_LAYERS = int(os.environ.get('E2_TAIL_LAYERS', '4'))

    def _opt_split(self):
        """(S, j, packs_early, packs_late): parameters [0, S) of the flat buffer -- every layer except the first
        TAIL_LAYERS -- have their final gradients once everything before bwd_ops[j] has run.  Those layers hold
        >95 % of the parameters of a U-Net and their update + re-pack is HBM-bound, so it runs on its own stream
        underneath the tensor-bound dgrad / wgrad of the large first layers."""
        if self._split is not None:
            return self._split
        wg = [i for i, f in enumerate(self.bwd_ops) if f.label.startswith(self._WGRAD)]
        S, j = 0, None
        if len(wg) >= 3:
            keep = min(self.TAIL_LAYERS, len(wg) - 1)
            j = wg[-keep]
            S = min(self.bwd_ops[i].param_start for i in wg[-keep:])
            ok = all(self.bwd_ops[i].param_end <= S for i in wg[:-keep]) and S <= self.store.n_reg
            if not ok:
                S, j = 0, None
        if j is not None:
            # E2_EARLY_DELAY: launch the early update that many backward ops later than the first moment it could run
            j = min(j + int(os.environ.get('E2_EARLY_DELAY', '0')), len(self.bwd_ops) - 1)
        off = {op: n.w._offset for n, op in self.conv_ops.items()}
        early = [op for op in self.pack_ops if off.get(op, 1 << 62) < S]
        late = [op for op in self.pack_ops if op not in early]
        self._split = (S, j, early, late)
        return self._split

    def _train_body_fwd(self, opt):
        opt.dev_prepare()
        for f in self.fwd_ops:
            f()

    def _train_body_bwd(self, opt, dp):
        store = self.store
        S, j, packs_early, packs_late = self._opt_split()
        hook = dp.on_gradients_ready if dp is not None else None
        if dp is not None:
            dp.begin_step(store, split=S)

        def early_update():
            if dp is not None:
                dp.on_gradients_ready(S)           # every bucket below S is launched by now
                dp.wait_launched()                 # this stream waits for those collectives
            self._update_and_pack(opt, store, 0, S, packs_early)

        self._launch_bwd(hook, (j, early_update) if j is not None and S > 0 else None)
        if dp is not None:
            dp.finish_step(store)
        for off, cnt, m in store.regions:
            if off >= store.n_reg:
                opt.dev_step(store, off, cnt, m)
        self._update_and_pack(opt, store, S, store.n_reg, packs_late if j is not None and S > 0 else self.pack_ops)

    def _can_fuse_pack(self, opt, store, lo, hi, pack_ops):
        """Is every parameter in [lo, hi) of the flat buffer the weight of one of ``pack_ops`` (and the optimiser one
        whose hyper-parameters live on the device)?  Then each layer's update + re-pack is one launch."""
        if os.environ.get('E2_FUSE_PACK', '0') != '1' or not getattr(opt, 'fusable', False):
            return False
        mine = {op.w.data_ptr(): op for op in pack_ops if getattr(op, 'w', None) is not None}
        need = [p for _, p, off, _ in store.entries if lo <= off < hi]
        return (len(need) == len(mine) == len(pack_ops) and all(p._dev.data_ptr() in mine for p in need)
                and all(op.adam_pack_ok() for op in pack_ops))

    def _update_and_pack(self, opt, store, lo, hi, pack_ops):
        """Adam update of the parameters [lo, hi) of the flat buffer (weights, weight-decay multiplier 1) and the
        re-pack of the layers ``pack_ops`` whose weights they are: one launch per layer that updates and re-packs
        (e2_*_adam_pack_dev: the weights cross HBM once) when the range is exactly those layers' weights, else the flat
        update followed by the separate pack kernels."""
        if not self._can_fuse_pack(opt, store, lo, hi, pack_ops):
            opt.dev_step(store, lo, hi - lo, 1)
            for op in pack_ops:
                op.pack(True)
            return
        for op in pack_ops:
            op.adam_pack(opt, store, 1.0)

    def update_launches(self, opt):
        """Kernel launches of the optimiser + re-pack part of a fused training step (``_train_body_fwd/_bwd``)."""
        store = self.store
        S, j, packs_early, packs_late = self._opt_split()
        parts = [(0, S, packs_early), (S, store.n_reg, packs_late)] if (j is not None and S > 0) else \
            [(0, store.n_reg, self.pack_ops)]
        n = 1 + len([1 for off, cnt, m in store.regions if off >= store.n_reg and cnt])     # adam_prepare + other regions
        for lo, hi, ops_ in parts:
            if self._can_fuse_pack(opt, store, lo, hi, ops_):
                n += len(ops_)
            else:
                before = self.h.launches
                for op in ops_:
                    op.pack(True)
                torch.cuda.synchronize(self.device)
                n += (self.h.launches - before) + (1 if hi > lo else 0)
        return n

    def train_step(self, opt, loss_async=False):
        """Forward + backward + optimiser update + weight re-pack as two CUDA graphs -- [forward, loss] and
        [backward, all-reduce, update, re-pack] -- so that the loss scalars can be copied out in between
        (``loss_async``: enqueue that copy; the caller collects it with ``loss_op.read_wait()`` while the backward
        pass is still running).  Adam only; other optimisers and ``keep_history`` fall back to execute() +
        opt.step().  The caller feeds the inputs first."""
        dp = self.model.data_parallel
        if dp is not None and dp.world <= 1:
            dp = None
        for a in self.bn_ops:
            a.update_running = True            # neural.py:695-698: the running averages are updates of the optimiser step
        try:
            self._train_step(opt, loss_async, dp)
        finally:
            for a in self.bn_ops:
                a.update_running = False

    def _train_step(self, opt, loss_async, dp):
        if not (self.train and self.fuse_optimiser and getattr(opt, 'fusable', False) and not opt.keep_history):
            self.execute()                     # includes the bucketed all-reduce when data parallel
            if loss_async:
                self.loss_op.read_async()
            opt.step(self.store)
            self.repack()
            return
        self._commit_inputs()
        opt.dev_sync(self.store)
        self._ensure_packed_train()            # first step / weights set from outside; afterwards the step re-packs
        key = id(opt)
        graphs = None
        if self.use_graph and (dp is None or dp.graph_ok):
            graphs = self._opt_graphs.get(key)
            if graphs is None:
                # eager warm-up WITHOUT the update (creates streams / communicators, touches every kernel once)
                for a in self.bn_ops:
                    a.update_running = False   # the warm-up pass must not move the running averages a second time
                if dp is not None:
                    dp.begin_step(self.store)
                self._launch_all(dp.on_gradients_ready if dp is not None else None)
                if dp is not None:
                    dp.finish_step(self.store)
                torch.cuda.synchronize(self.device)
                for a in self.bn_ops:
                    a.update_running = True
                try:
                    g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                    # captured on a high-priority stream: kernel nodes inherit it, so whenever a dgrad (critical
                    # path) and a wgrad / optimiser kernel (side streams, default priority) are both ready, the
                    # dgrad's CTAs are scheduled first
                    if self._capture_stream is None:
                        self._capture_stream = torch.cuda.Stream(device=self.device, priority=-1) \
                            if os.environ.get('E2_MAIN_PRIO', '1') != '0' else torch.cuda.Stream(device=self.device)
                    with _capture(g1, self._capture_stream):
                        self._train_body_fwd(opt)
                    with _capture(g2, self._capture_stream):
                        self._train_body_bwd(opt, dp)
                    # the captures ran no kernel: t_dev / parameters are untouched
                    graphs = self._opt_graphs[key] = (g1, g2)
                except Exception as e:      # noqa: BLE001
                    # a refused / invalidated capture must not cost the step: run it eagerly from now on
                    import warnings
                    warnings.warn('CUDA-graph capture of the training step failed (%s); running eagerly' % (e,))
                    if dp is not None:
                        dp.graph_ok = False
                    self.use_graph = False
                    try:
                        torch.cuda.synchronize(self.device)
                    except Exception:       # noqa: BLE001 - the capture error may surface once more here
                        torch.cuda.synchronize(self.device)
                    graphs = None
        if graphs is not None:
            graphs[0].replay()
        else:
            self._train_body_fwd(opt)
        self._mark_inputs_free()               # staging buffers / input layout kernel consumed
        if loss_async:
            self.loss_op.read_async()
        if graphs is not None:
            graphs[1].replay()
        else:
            self._train_body_bwd(opt, dp)
        opt.dev_after(self.store)
        self._packed_version = self.store.version

    def _ensure_packed(self, need_dgrad=None):
        """Re-pack the weights (tf32 rounding, tap flip, K-major layouts) if the parameters changed since the
        last pack (inference plans; training plans re-pack inside every step, see ``_step_with_pack``)."""
        if self._packed_version != self.store.version:
            self.pack(need_dgrad)

    def _ensure_packed_train(self):
        """Training: the re-pack is a small CUDA graph of its own.  ``Model.trainingstep`` replays it right after
        the optimiser, so it overlaps with the host side of the next call and the next step starts with the
        first convolution; a step that finds stale packed weights replays it first."""
        if self._packed_version == self.store.version:
            return
        if self.use_graph:
            if self._pack_graph is None:
                self.pack()                    # warm-up outside capture
                torch.cuda.synchronize(self.device)
                self._pack_graph = _try_capture(self.pack, self.device, 'the weight re-pack') or False
            if self._pack_graph:
                self._pack_graph.replay()
                self._packed_version = self.store.version
            else:
                self.pack()
        else:
            self.pack()

    def repack(self):
        """Refresh the packed weights now (after an optimiser step) instead of at the start of the next step."""
        if not self.train:
            self._ensure_packed(False)
        elif not self.overlap_pack:
            self._ensure_packed_train()

    EARLY_PACKS = 2     # layers whose weights are packed on the compute stream, ahead of the first convolution

    def _step_with_pack(self, hook=None):
        """E2_PACK_OVERLAP=1 variant of a training step: the weight re-pack (fp32 master -> tf32 forward / dgrad
        layouts, ~230 MB of traffic for unet3d) is part of the step's launch list; only the first layers' packs sit
        in front of the forward pass, the rest run on a side stream underneath the first (tensor-bound) convolutions
        and are joined before the third conv layer.  Measured on unet3d (forked at the step start and forked behind
        the first layer): no gain over the sequential order -- the packs slow the convolution they run under by
        what they save -- so the default keeps the re-pack outside the step graph."""
        conv_idx = [i for i, f in enumerate(self.fwd_ops) if f.label.startswith(('conv_fwd:', 'upconv_fwd:'))]
        k = self.EARLY_PACKS
        if len(conv_idx) <= k or len(conv_idx) != len(self.pack_ops):
            self.pack()
            self._launch_all(hook)
            return
        main = torch.cuda.current_stream(self.device)
        if self._pack_stream is None:
            self._pack_stream = torch.cuda.Stream(device=self.device)
        side = self._pack_stream
        fork, join = torch.cuda.Event(), torch.cuda.Event()
        for op in self.pack_ops[:k]:
            op.pack(True)
        self._packed_version = self.store.version
        for i, f in enumerate(self.fwd_ops):
            if i == conv_idx[k]:
                main.wait_event(join)
            f()
            if i == conv_idx[0]:
                # behind the first (HBM-bound) layer: the remaining packs run underneath the second, tensor-bound one
                fork.record(main)
                side.wait_event(fork)
                with torch.cuda.stream(side):
                    for op in self.pack_ops[k:]:
                        op.pack(True)
                    join.record(side)
        self._launch_bwd(hook)

    def execute(self):
        """Run the launch list: eagerly the first time (warm-up), then as a CUDA graph.  With data parallelism
        the bucketed all-reduce is interleaved with the backward launches and captured with them."""
        self._commit_inputs()
        if not self.train:
            self._ensure_packed(False)
            if self.use_graph and self._graph is None:
                self._launch_all()  # warm-up outside capture
                torch.cuda.synchronize(self.device)
                self._graph = _try_capture(self._launch_all, self.device, 'the forward pass')
                if self._graph is None:
                    self.use_graph = False
            if self._graph is not None:
                self._graph.replay()
            else:
                self._launch_all()
            self._mark_inputs_free()
            return
        dp = self.model.data_parallel
        if dp is not None and dp.world <= 1:
            dp = None

        if not self.overlap_pack:
            self._ensure_packed_train()

        def body():
            hook = dp.on_gradients_ready if dp is not None else None
            if dp is not None:
                dp.begin_step(self.store)
            if self.overlap_pack:
                self._step_with_pack(hook)
            else:
                self._launch_all(hook)
            if dp is not None:
                dp.finish_step(self.store)

        # Data parallel: the bucketed NCCL all-reduces (side stream, forked and joined with events) are captured
        # into the same CUDA graph as the kernels: one graph launch per step on every rank.  If this torch/NCCL
        # build refuses the capture the step stays eager (E2_DP_GRAPH=0 forces that).
        # one captured graph per value of the batch-norm running-average switch (graph nodes bake their arguments)
        gkey = '_graph_upd' if (self.bn_ops and self.bn_ops[0].update_running) else '_graph'
        if self.use_graph and (dp is None or dp.graph_ok):
            if getattr(self, gkey) is None:
                body()                      # eager warm-up (also creates the communicators outside the capture)
                torch.cuda.synchronize(self.device)
                setattr(self, gkey, _try_capture(body, self.device, 'the training pass'))
                if getattr(self, gkey) is None:     # refused / invalidated: this plan stays eager
                    self.use_graph = False
                    if dp is not None:
                        dp.graph_ok = False
                self._mark_inputs_free()
                return
            getattr(self, gkey).replay()
        else:
            body()
        self._mark_inputs_free()

    def release_graphs(self):
        """Drop every captured CUDA graph of this plan (they hold references to streams, scratch buffers and -- with
        data parallelism -- the NCCL communicator).  Called before an orderly process teardown; the next step
        re-captures."""
        torch.cuda.synchronize(self.device)
        self._opt_graphs.clear()
        self._graph = None
        self._graph_upd = None
        self._pack_graph = None
        gc.collect()

    def _mark_inputs_free(self):
        if self._inputs_free is None:
            self._inputs_free = torch.cuda.Event()
        self._inputs_free.record()

    def launches_per_step(self, with_pack=True):
        """Kernel launches of one step (counted by the library, eager pass); ``with_pack``: including the separate
        weight re-pack (inference plans after a parameter change; training steps count ``update_launches`` instead)."""
        if with_pack:
            self.pack()
        before = self.h.launches
        if with_pack:
            self.pack()
        self._launch_all()
        torch.cuda.synchronize(self.device)
        return self.h.launches - before

    def profile(self, repeats=3, opt=None):
        """Per-launch device times (ms, CUDA events on the launching stream), eager
        passes after one warm-up; returns [(label, kind, flops, bytes, ms)].  With ``opt`` (training plans) the last
        entry is the optimiser update + weight re-pack as the fused step runs it (it really updates the parameters)."""
        self.pack()
        self._launch_all()
        torch.cuda.synchronize(self.device)
        all_ops = self.fwd_ops + self.bwd_ops
        acc = [0.0] * len(all_ops)
        for _ in range(repeats):
            evs = []
            for f in all_ops:
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                f()
                b.record()
                evs.append((a, b))
            torch.cuda.synchronize(self.device)
            for i, (a, b) in enumerate(evs):
                acc[i] += a.elapsed_time(b)
        out = [(f.label, f.kind, f.flops, f.bytes, acc[i] / repeats) for i, f in enumerate(all_ops)]
        if self.train and opt is not None and getattr(opt, 'fusable', False):
            store = self.store
            opt.dev_sync(store)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(repeats):
                opt.dev_prepare()
                for off, cnt, m in store.regions:
                    if off >= store.n_reg:
                        opt.dev_step(store, off, cnt, m)
                self._update_and_pack(opt, store, 0, store.n_reg, self.pack_ops)
            b.record()
            torch.cuda.synchronize(self.device)
            # p, g, m, s read + p, m, s and the two packed images written
            out.append(('adam_update_pack', 'hbm', 0.0, 36.0 * store.n_reg, a.elapsed_time(b) / repeats))
        elif self.train:
            # the per-step weight re-pack (read the fp32 master once, write the fwd and dgrad layouts)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(repeats):
                self.pack()
            b.record()
            torch.cuda.synchronize(self.device)
            out.append(('pack_weights', 'hbm', 0.0, 12.0 * self.store.n_reg, a.elapsed_time(b) / repeats))
        return out

    def fetch(self, node):
        if isinstance(node, loss_nodes.AggregateLoss):
            return np.float32(self.loss_op.read()[0])
        if isinstance(node, loss_nodes._Errors):
            return np.float32(self.loss_op.read()[1])
        if isinstance(node, loss_nodes.Classification):
            return np.argmax(self.val[node.pred].numpy(self.h), axis=1)[:, None].astype(np.int16)
        if isinstance(node, loss_nodes.MultinoulliNLL):
            raise NotImplementedError("the per-voxel NLL tensor is not materialised on the B200 path")
        out = self.val[node].numpy(self.h)
        nd = len(node.shape)
        if 3 <= nd < 5:
            out = out.reshape(out.shape[:2] + out.shape[2 + (5 - nd):])
        return out

    def run(self, values):
        self.feed(values)
        self.execute()
        return [self.fetch(o) for o in self.outputs]
