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
import ctypes as C

import numpy as np
import torch

from .. import _lib
from .. import ops
from ..config import config
from ..devtensor import DevTensor
from . import loss as loss_nodes
from .neural import Conv, UpConv, Pool, Crop, FragmentsToDense
from .node_basic import Input, Concat


class ParamStore(object):
    """Flat device buffers for all trainable parameters of a model."""

    def __init__(self, named_params, device):
        # reverse graph order inside each region; weights (regularised) first, then the rest
        plist = list(named_params)[::-1]
        reg = [(k, p) for k, p in plist if p.apply_reg]
        noreg = [(k, p) for k, p in plist if not p.apply_reg]
        self.entries = []  # (key, param, offset, size)
        off = 0
        for group in (reg, noreg):
            for k, p in group:
                size = int(np.prod(p.shape))
                self.entries.append((k, p, off, size))
                off += (size + 3) // 4 * 4
            if group is reg:
                self.n_reg = off
        self.total = off
        self.device = device
        self.P = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.G = torch.zeros(self.total, dtype=torch.float32, device=device)
        self.version = 0
        for k, p, o, size in self.entries:
            view = self.P[o:o + size].view(*p.shape)
            view.copy_(torch.from_numpy(np.ascontiguousarray(p._host)))
            p._dev = view
            p._grad = self.G[o:o + size].view(*p.shape)
            p._on_set.append(self._bump)

    def _bump(self):
        self.version += 1

    def grads_numpy(self):
        return [p._grad.detach().cpu().numpy().copy() for _, p, _, _ in self.entries]


def _nb(node, batch):
    b = node.shape['b']
    return int(batch) if b is None else int(b)


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
        self.loss_op = None
        self.inputs = {}
        self._graph = None
        self._packed_version = -1
        self._build_forward()
        if train:
            self._build_backward()

    # ------------------------------------------------------------------ helpers
    def _consumers(self, n):
        return [c for c in n.children.values() if id(c) in self._in_plan]

    def _new(self, node, c=None):
        sp = node.shape.spatial_shape
        return DevTensor(_nb(node, self.batch), sp[0], sp[1], sp[2], node.shape['f'] if c is None else c,
                         device=self.device)

    def _pdev(self, p):
        if p._dev is None:
            raise RuntimeError("parameter %s is not part of the model's parameter store" % p.name)
        return p._dev

    # ------------------------------------------------------------------ forward
    def _build_forward(self):
        h = self.h
        out_ids = set(id(o) for o in self.outputs)
        self.crop_into, self.copy_into, self.alias_of = {}, {}, {}
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
                y = self.val.get(n) or self._new(n)
                self.val[n] = y
                op = ops.UpConvOp(h, x, y, self._pdev(n.w), self._pdev(n.b), n.pool_shape, n.activation_func, self.compute)
                self.conv_ops[n] = op
                self.pack_ops.append(op)
                self.fwd_ops.append(op.fwd)
            elif isinstance(n, Conv):
                self._plan_conv(n)
            elif isinstance(n, Pool):
                x = self.val[n.parent]
                y = self._new(n)
                self.val[n] = y
                op = ops.PoolOp(h, x, y, n.pool_shape, keep_argmax=self.train)
                self.aux[n] = op
                self.fwd_ops.append(op.fwd)
            elif isinstance(n, Crop):
                src = self.val[n.parent]
                if n in self.crop_into:
                    cat, c0 = self.crop_into[n]
                    dst = self.val[cat]
                    self.val[n] = dst.channel_slice(c0, n.shape['f'])
                else:
                    dst, c0 = self._new(n), 0
                    self.val[n] = dst
                op = ops.CropConcatOp(h, src, dst, n.crop, c0)
                self.aux[n] = op
                self.fwd_ops.append(op.fwd)
            elif isinstance(n, Concat):
                cops = []
                for p, c0 in self.copy_into.get(n, []):
                    op = ops.CropConcatOp(h, self.val[p], self.val[n], (0, 0, 0), c0)
                    cops.append((p, op))
                    self.fwd_ops.append(op.fwd)
                self.aux[n] = cops
            elif isinstance(n, FragmentsToDense):
                x = self.val[n.parent]
                y = self._new(n)
                self.val[n] = y
                op = ops.Frag2DenseOp(h, x, y, n.parent.shape.mfp_offsets, n.parent.shape.strides)
                self.aux[n] = op
                self.fwd_ops.append(op.fwd)
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
            self.fwd_ops.append(lambda t=t, s=staging: self.h.call('e2_ncdhw_to_ndhwc', C.byref(t.desc), _lib.ptr(s),
                                                                   t.ptr(), self.h.stream()))
        self.inputs[n] = (t, pinned, staging)

    def _plan_conv(self, n):
        h = self.h
        x = self.val[n.parent]
        y = self.val.get(n) or self._new(n)
        self.val[n] = y
        pooled = any(p > 1 for p in n.pool_shape)
        w, b = self._pdev(n.w), self._pdev(n.b)
        if not pooled:
            op = ops.ConvOp(h, x, y, w, b, n.filter_shape, n.activation_func, self.compute)
            self.fwd_ops.append(op.fwd)
        else:
            # conv -> pool|MFP -> +bias -> act  (neural.py:662-712): the conv writes raw
            # accumulators, the pooling kernel carries the bias/activation epilogue
            lsp = [s + 1 - f for s, f in zip(n.parent.shape.spatial_shape, n.filter_shape)]
            lin = DevTensor(x.desc.n, lsp[0], lsp[1], lsp[2], n.n_f, device=self.device)
            op = ops.ConvOp(h, x, lin, w, None, n.filter_shape, 'lin', self.compute)
            self.fwd_ops.append(op.fwd)
            if n.mfp:
                pop = ops.MfpOp(h, lin, y, n.pool_shape, bias=b, act=n.activation_func, keep_argmax=self.train)
            else:
                pop = ops.PoolOp(h, lin, y, n.pool_shape, bias=b, act=n.activation_func, keep_argmax=self.train)
            self.aux[n] = (lin, pop)
            self.fwd_ops.append(pop.fwd)
        self.conv_ops[n] = op
        self.pack_ops.append(op)

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
        self.logits_node = sm.parent
        self.fwd_ops.append(self.loss_op.fwd)

    # ----------------------------------------------------------------- backward
    def _build_backward(self):
        if self.loss_op is None or self.loss_op.target is None:
            raise ValueError("a training plan needs a Softmax -> MultinoulliNLL -> AggregateLoss head")
        h = self.h
        self.grad, written = {}, set()
        # gradient buffers (aliases follow the forward aliasing)
        for n in self.nodes:
            if n in self.val and not isinstance(n, (Input,)) and isinstance(self.val[n], DevTensor) \
                    and not isinstance(n, loss_nodes.Softmax):
                if n in self.alias_of or n in self.crop_into:
                    continue
                self.grad[n] = self.val[n].like()
        for n, (cat, c0) in self.alias_of.items():
            self.grad[n] = self.grad[cat].channel_slice(c0, n.shape['f'])
        self.grad_scale = 1.0
        lg = self.logits_node
        self.bwd_ops.append(lambda: self.loss_op.bwd(self.grad[lg], self.grad_scale))
        written.add(lg)
        self.wgrad_done = []  # (node, launch index) -- used to place all-reduce buckets
        for n in reversed(self.nodes):
            if n not in written and n not in self.crop_into:
                continue
            if isinstance(n, UpConv):
                dy, op = self.grad[n], self.conv_ops[n]
                if n.activation_func not in ('lin', 'linear'):
                    self.bwd_ops.append(lambda n=n, dy=dy: ops.act_bwd(h, dy, n.activation_func, self.val[n], dy, dy))
                self.bwd_ops.append(lambda op=op, dy=dy, n=n: op.wgrad(dy, n.w._grad, n.b._grad))
                self.wgrad_done.append((n, len(self.bwd_ops)))
                self._emit_dgrad(op, dy, n.parent, written)
            elif isinstance(n, Conv):
                dy, op = self.grad[n], self.conv_ops[n]
                if n.activation_func not in ('lin', 'linear'):
                    self.bwd_ops.append(lambda n=n, dy=dy: ops.act_bwd(h, dy, n.activation_func, self.val[n], dy, dy))
                if n in self.aux:
                    lin, pop = self.aux[n]
                    dlin = lin.like()
                    self.bwd_ops.append(lambda pop=pop, dy=dy, dlin=dlin: pop.bwd(dy, dlin))
                else:
                    dlin = dy
                self.bwd_ops.append(lambda op=op, dlin=dlin, n=n: op.wgrad(dlin, n.w._grad, n.b._grad))
                self.wgrad_done.append((n, len(self.bwd_ops)))
                self._emit_dgrad(op, dlin, n.parent, written)
            elif isinstance(n, Pool):
                par = n.parent
                if isinstance(par, Input):
                    continue
                acc = par in written
                self.bwd_ops.append(lambda op=self.aux[n], dy=self.grad[n], dx=self.grad[par], acc=acc: op.bwd(dy, dx, acc))
                written.add(par)
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
                acc = par in written
                self.bwd_ops.append(lambda op=self.aux[n], ddst=ddst, dx=self.grad[par], acc=acc: op.bwd(ddst, dx, acc))
                written.add(par)
            elif isinstance(n, Concat):
                for p in n.parents:
                    if p in self.alias_of:
                        written.add(p)
                for p, op in self.aux.get(n, []):
                    if isinstance(p, Input):
                        continue
                    acc = p in written
                    self.bwd_ops.append(lambda op=op, ddst=self.grad[n], dx=self.grad[p], acc=acc: op.bwd(ddst, dx, acc))
                    written.add(p)
            elif isinstance(n, FragmentsToDense):
                par = n.parent
                if par in written:
                    raise NotImplementedError("FragmentsToDense parent with several consumers")
                self.bwd_ops.append(lambda op=self.aux[n], dd=self.grad[n], df=self.grad[par]: op.bwd(dd, df))
                written.add(par)

    def _emit_dgrad(self, op, dy, parent, written):
        if isinstance(parent, Input):
            return  # gradients are taken w.r.t. parameters only (model.py:182)
        acc = parent in written
        self.bwd_ops.append(lambda op=op, dy=dy, dx=self.grad[parent], acc=acc: op.dgrad(dy, dx, acc))
        written.add(parent)

    # ---------------------------------------------------------------- execution
    def pack(self, need_dgrad=None):
        nd = self.train if need_dgrad is None else need_dgrad
        for op in self.pack_ops:
            op.pack(nd)
        self._packed_version = self.store.version

    def feed(self, values):
        """H2D copies of the inputs from pinned host memory (async on the current stream)."""
        nbytes = 0
        for n, a in values.items():
            t, pinned, staging = self.inputs[n]
            a = np.asarray(a, dtype=np.float32)
            if tuple(a.shape) != tuple(pinned.shape):
                raise ValueError("Input '%s' expects shape %s, got %s" % (n.name, tuple(pinned.shape), a.shape))
            pinned.copy_(torch.from_numpy(np.ascontiguousarray(a)))
            dst = staging if staging is not None else t.buf[t.offset:t.offset + pinned.numel()].view(pinned.shape)
            dst.copy_(pinned, non_blocking=True)
            nbytes += pinned.numel() * 4
        return nbytes

    def _launch_all(self):
        for f in self.fwd_ops:
            f()
        for f in self.bwd_ops:
            f()

    def execute(self):
        """Run the launch list: eagerly the first time (warm-up), then as a CUDA graph."""
        if self.train:
            # weights change every step: the re-pack is part of the step itself
            if self.use_graph:
                if self._graph is None:
                    self.pack()
                    self._launch_all()  # warm-up outside capture
                    torch.cuda.synchronize(self.device)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self.pack()
                        self._launch_all()
                    self._graph = g
                self._graph.replay()
            else:
                self.pack()
                self._launch_all()
            return
        if self._packed_version != self.store.version:
            self.pack()
        if self.use_graph:
            if self._graph is None:
                self._launch_all()
                torch.cuda.synchronize(self.device)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_all()
                self._graph = g
            self._graph.replay()
        else:
            self._launch_all()

    def launches_per_step(self):
        before = self.h.launches
        self.pack()
        self._launch_all()
        return self.h.launches - before

    def fetch(self, node):
        if isinstance(node, loss_nodes.AggregateLoss):
            return np.float32(self.loss_op.read()[0])
        if isinstance(node, loss_nodes._Errors):
            return np.float32(self.loss_op.read()[1])
        if isinstance(node, loss_nodes.Classification):
            return np.argmax(self.val[node.pred].numpy(self.h), axis=1)[:, None].astype(np.int16)
        if isinstance(node, loss_nodes.MultinoulliNLL):
            raise NotImplementedError("the per-voxel NLL tensor is not materialised on the B200 path")
        return self.val[node].numpy(self.h)

    def run(self, values):
        self.feed(values)
        self.execute()
        return [self.fetch(o) for o in self.outputs]
