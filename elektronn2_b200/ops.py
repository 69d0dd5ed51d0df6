"""Op objects: one per hot-path primitive, each binding caller-owned device buffers to a
C-ABI descriptor once and then launching with no further host work.

These are the Python face of include/e2b200.h.  ``elektronn2_b200.computations`` wraps
them with the reference's function signatures; the graph executor
(``neuromancer/executor.py``) strings them into static plans.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT, COMPUTE, TIE, POOL_MODE
from .devtensor import DevTensor


def _dev_f32(n, device):
    return torch.zeros(int(n), dtype=torch.float32, device=device)


class ConvOp(object):
    """computations.conv 3-D valid branch + bias/activation epilogue
    (reference: computations.py:364-428, neural.py:711-712)."""

    def __init__(self, h, x, y, w, b, k, act='relu', compute='tf32'):
        self.h, self.x, self.y, self.w, self.b = h, x, y, w, b
        self.k = tuple(int(v) for v in k)
        self.d = _lib.ConvDesc(x.desc, y.desc, self.k[0], self.k[1], self.k[2], ACT[act], 1 if b is not None else 0,
                               COMPUTE[compute], 0)
        f, g = C.c_size_t(), C.c_size_t()
        _lib.lib.e2_conv3d_packed_floats(C.byref(self.d), C.byref(f), C.byref(g))
        self.wf = _dev_f32(f.value, x.buf.device)
        self.wd = _dev_f32(g.value, x.buf.device)
        ws = C.c_size_t()
        _lib.lib.e2_conv3d_workspace_size(C.byref(self.d), C.byref(ws))
        h.reserve_workspace(ws.value)

    def _desc(self, x=None, y=None, accumulate=0, act=None, has_bias=None):
        d = _lib.ConvDesc()
        C.memmove(C.byref(d), C.byref(self.d), C.sizeof(d))
        if x is not None:
            d.x = x.desc
        if y is not None:
            d.y = y.desc
        d.accumulate = int(accumulate)
        if act is not None:
            d.act = ACT[act]
        if has_bias is not None:
            d.has_bias = int(has_bias)
        return d

    def pack(self, need_dgrad=True):
        self.h.call('e2_conv3d_pack_weights', C.byref(self.d), _lib.ptr(self.w), _lib.ptr(self.wf),
                    _lib.ptr(self.wd) if need_dgrad else None, self.h.stream())

    def fwd(self, act=None, has_bias=None, y=None):
        d = self.d if (act is None and has_bias is None and y is None) else self._desc(act=act, has_bias=has_bias, y=y)
        yy = y if y is not None else self.y
        ws, ws_bytes = self.h.workspace()
        self.h.call('e2_conv3d_fwd', C.byref(d), self.x.ptr(), _lib.ptr(self.wf),
                    _lib.ptr(self.b) if d.has_bias else None, yy.ptr(), ws, ws_bytes, self.h.stream())

    def adam_pack(self, opt, store, wd_mult=1.0):
        """Adam update of this layer's weights (a view into ``store.P``) and the re-pack of the updated weights in one
        launch (e2_conv3d_adam_pack_dev); ``opt`` holds the moments and the device-resident hyper-parameters."""
        off = (self.w.data_ptr() - store.P.data_ptr()) // 4
        sl = lambda t: C.c_void_p(t.data_ptr() + 4 * off)      # noqa: E731
        self.h.call('e2_conv3d_adam_pack_dev', C.byref(self.d), sl(store.P), sl(store.G), sl(opt._state[0]),
                    sl(opt._state[1]), _lib.ptr(opt._hyper_dev), C.c_float(wd_mult), _lib.ptr(self.wf), _lib.ptr(self.wd),
                    self.h.stream())

    def adam_pack_ok(self):
        return self.k[0] * self.k[1] * self.k[2] <= 125

    def pool_fusable(self, pop):
        """Does libe2b200 run this conv and the max-pool ``pop`` (a PoolOp whose input is this conv's output) as one
        launch (e2_conv3d_fwd_pool_supported)?"""
        if not pop.is_max or pop.x.desc.dims() != self.y.desc.dims() or pop.x.desc.c_pitch != self.y.desc.c_pitch:
            return False
        return self.h.query('e2_conv3d_fwd_pool_supported', C.byref(self.d), C.byref(pop.d)) == 1

    def fwd_pool(self, pop, store_full=True, keep=None):
        """conv -> max-pool in one launch (e2_conv3d_fwd_pool): the same bits as ``self.fwd(); pop.fwd()``;
        with store_full=False the unpooled tensor is not written at all, with keep=(z0,z1,x0,x1,y0,y1) only the
        part of it inside that window is guaranteed to be."""
        ws, ws_bytes = self.h.workspace()
        win = C.byref(_lib.Window(*[int(v) for v in keep])) if keep is not None else None
        self.h.call('e2_conv3d_fwd_pool', C.byref(self.d), C.byref(pop.d), self.x.ptr(), _lib.ptr(self.wf),
                    _lib.ptr(self.b) if self.d.has_bias else None, _lib.ptr(pop.bias),
                    self.y.ptr() if store_full else None, win, pop.y.ptr(),
                    pop.argmax.ptr() if pop.argmax is not None else None, ws, ws_bytes, self.h.stream())

    def dgrad(self, dy, dx, accumulate=False, relu_gate=None):
        """relu_gate: post-ReLU output of the layer that produced x (fused ReLU backward)."""
        d = self._desc(x=dx, y=dy, accumulate=accumulate)
        ws, ws_bytes = self.h.workspace()
        self.h.call('e2_conv3d_dgrad', C.byref(d), dy.ptr(), _lib.ptr(self.wd), dx.ptr(),
                    relu_gate.ptr() if relu_gate is not None else None, ws, ws_bytes, self.h.stream())

    def wgrad(self, dy, dw, db=None):
        d = self._desc(y=dy)
        ws, ws_bytes = self.h.workspace()
        self.h.call('e2_conv3d_wgrad', C.byref(d), self.x.ptr(), dy.ptr(), _lib.ptr(dw), _lib.ptr(db), ws, ws_bytes,
                    self.h.stream())


class UpConvOp(object):
    """UpConv (neural.py:989-1072; computations.py:216-255, 749-756)."""

    def __init__(self, h, x, y, w, b, pool, act='relu', compute='tf32'):
        self.h, self.x, self.y, self.w, self.b = h, x, y, w, b
        self.pool = tuple(int(v) for v in pool)
        self.d = _lib.UpConvDesc(x.desc, y.desc, self.pool[0], self.pool[1], self.pool[2], ACT[act],
                                 1 if b is not None else 0, COMPUTE[compute], 0)
        f, g = C.c_size_t(), C.c_size_t()
        _lib.lib.e2_upconv3d_packed_floats(C.byref(self.d), C.byref(f), C.byref(g))
        self.wf = _dev_f32(f.value, x.buf.device)
        self.wd = _dev_f32(g.value, x.buf.device)
        ws = C.c_size_t()
        _lib.lib.e2_upconv3d_workspace_size(C.byref(self.d), C.byref(ws))
        h.reserve_workspace(ws.value)

    def _desc(self, x=None, y=None, accumulate=0):
        d = _lib.UpConvDesc()
        C.memmove(C.byref(d), C.byref(self.d), C.sizeof(d))
        if x is not None:
            d.x = x.desc
        if y is not None:
            d.y = y.desc
        d.accumulate = int(accumulate)
        return d

    def pack(self, need_dgrad=True):
        self.h.call('e2_upconv3d_pack_weights', C.byref(self.d), _lib.ptr(self.w), _lib.ptr(self.wf),
                    _lib.ptr(self.wd) if need_dgrad else None, self.h.stream())

    def adam_pack(self, opt, store, wd_mult=1.0):
        """See ConvOp.adam_pack (e2_upconv3d_adam_pack_dev)."""
        off = (self.w.data_ptr() - store.P.data_ptr()) // 4
        sl = lambda t: C.c_void_p(t.data_ptr() + 4 * off)      # noqa: E731
        self.h.call('e2_upconv3d_adam_pack_dev', C.byref(self.d), sl(store.P), sl(store.G), sl(opt._state[0]),
                    sl(opt._state[1]), _lib.ptr(opt._hyper_dev), C.c_float(wd_mult), _lib.ptr(self.wf), _lib.ptr(self.wd),
                    self.h.stream())

    def adam_pack_ok(self):
        return self.pool[0] * self.pool[1] * self.pool[2] <= 125

    def fwd(self):
        ws, ws_bytes = self.h.workspace()
        self.h.call('e2_upconv3d_fwd', C.byref(self.d), self.x.ptr(), _lib.ptr(self.wf),
                    _lib.ptr(self.b) if self.d.has_bias else None, self.y.ptr(), ws, ws_bytes, self.h.stream())

    def dgrad(self, dy, dx, accumulate=False, relu_gate=None):
        d = self._desc(x=dx, y=dy, accumulate=accumulate)
        ws, ws_bytes = self.h.workspace()
        self.h.call('e2_upconv3d_dgrad', C.byref(d), dy.ptr(), _lib.ptr(self.wd), dx.ptr(),
                    relu_gate.ptr() if relu_gate is not None else None, ws, ws_bytes, self.h.stream())

    def wgrad(self, dy, dw, db=None):
        d = self._desc(y=dy)
        ws, ws_bytes = self.h.workspace()
        self.h.call('e2_upconv3d_wgrad', C.byref(d), self.x.ptr(), dy.ptr(), _lib.ptr(dw), _lib.ptr(db), ws, ws_bytes,
                    self.h.stream())


class PoolOp(object):
    """computations.pooling 3-D max (computations.py:538-649) with optional fused
    +bias -> act (Conv nodes that carry a pool, neural.py:678, 711-712)."""

    def __init__(self, h, x, y, pool, bias=None, act='lin', keep_argmax=True, tie_mode='first', round_tf32=False,
                 mode='max'):
        self.h, self.x, self.y, self.bias = h, x, y, bias
        self.pool = tuple(int(v) for v in pool)
        if mode not in POOL_MODE:
            raise ValueError("unknown pooling mode %r (computations.py:556-561)" % (mode,))
        self.d = _lib.PoolDesc(x.desc, y.desc, self.pool[0], self.pool[1], self.pool[2], ACT[act],
                               1 if bias is not None else 0, TIE[tie_mode], 0, int(bool(round_tf32)), 0, POOL_MODE[mode])
        self.is_max = POOL_MODE[mode] == 0
        self.argmax = y.like(dtype=torch.int32) if (keep_argmax and self.is_max) else None

    def fwd(self):
        self.h.call('e2_maxpool3d_fwd', C.byref(self.d), self.x.ptr(), _lib.ptr(self.bias), self.y.ptr(),
                    self.argmax.ptr() if self.argmax is not None else None, self.h.stream())

    def bwd(self, dy, dx, accumulate=False, relu_gate=None):
        d = _lib.PoolDesc()
        C.memmove(C.byref(d), C.byref(self.d), C.sizeof(d))
        d.x, d.y, d.accumulate = dx.desc, dy.desc, int(accumulate)
        gate = relu_gate
        if not self.is_max:
            # average / sum pooling: no argmax, no fused gate (the executor emits the activation backward separately)
            assert relu_gate is None
            self.h.call('e2_maxpool3d_bwd', C.byref(d), dy.ptr(), None, None, dx.ptr(), None, self.h.stream())
            return
        if (relu_gate is self.x and self.argmax is not None and not d.has_bias and d.act == ACT['lin']
                and d.tie_mode == TIE['first']):
            # the gate is this pool's own (post-ReLU) input: gate[argmax] == pooled value, read y instead of x
            gate, d.gate_pooled = self.y, 1
        self.h.call('e2_maxpool3d_bwd', C.byref(d), dy.ptr(), self.argmax.ptr() if self.argmax is not None else None,
                    self.x.ptr(), dx.ptr(), gate.ptr() if gate is not None else None, self.h.stream())


class AffineActOp(object):
    """The unfused epilogue y = act((gamma/std) * v + b - gamma*mean/std) of Conv / UpConv (neural.py:681-712) for
    batch normalisation, 'prelu' and activations whose derivative needs the pre-activation.

    bn: False | 'train' | 'predict'.  Parameters are device views: b (f,) or (f,2) for prelu; gamma / mean / std (f,).
    In 'train' mode ``fwd`` computes the batch statistics and, when ``update_running`` is set, the running averages
    mean <- 0.9995 mean + 0.0005 batch_mean (same for std), which the reference applies as Theano updates of the
    optimiser step (neural.py:695-698)."""

    KEEP = 0.9995

    def __init__(self, h, v, y, act, b, bn=False, gamma=None, mean=None, std=None, round_tf32=False):
        self.h, self.v, self.y, self.act, self.b, self.bn = h, v, y, act, b, bn
        self.gamma, self.mean, self.std = gamma, mean, std
        self.prelu = act == 'prelu'
        stride = 2 if self.prelu else 1
        self.d = _lib.AffineDesc(v.desc, ACT[act], 1 if bn == 'train' else 0, int(bool(round_tf32)), stride)
        c = v.desc.c
        dev = v.buf.device
        self.scale = _dev_f32(c, dev)
        self.shift = _dev_f32(c, dev)
        self.bmean = _dev_f32(c, dev)          # batch statistics ('train')
        self.bstd = _dev_f32(c, dev)
        nb = C.c_size_t()
        _lib.lib.e2_affine_scratch_bytes(c, C.byref(nb))
        self.scratch = torch.zeros((nb.value + 7) // 8, dtype=torch.float64, device=dev)
        self.update_running = False
        self.alpha = None
        if self.prelu:
            self.alpha = C.c_void_p(b.data_ptr() + 4)       # b[:, 1]

    def fwd(self):
        h, s = self.h, self.h.stream()
        if self.bn == 'train':
            upd = self.update_running
            h.call('e2_bn_batch_stats', C.byref(self.v.desc), self.v.ptr(), _lib.ptr(self.gamma), _lib.ptr(self.b),
                   self.d.param_stride, _lib.ptr(self.bmean), _lib.ptr(self.bstd), _lib.ptr(self.scale),
                   _lib.ptr(self.shift), _lib.ptr(self.mean) if upd else None, _lib.ptr(self.std) if upd else None,
                   C.c_float(self.KEEP), _lib.ptr(self.scratch), s)
        else:
            h.call('e2_bn_fold', self.v.desc.c, _lib.ptr(self.gamma) if self.bn else None, _lib.ptr(self.b),
                   self.d.param_stride, _lib.ptr(self.mean) if self.bn else None, _lib.ptr(self.std) if self.bn else None,
                   _lib.ptr(self.scale), _lib.ptr(self.shift), s)
        h.call('e2_affine_act_fwd', C.byref(self.d), self.v.ptr(), _lib.ptr(self.scale), _lib.ptr(self.shift), self.alpha,
               self.y.ptr(), s)

    def bwd(self, dy, dv, db, dgamma=None):
        """dv, db (and dgamma, and for prelu the slope gradient interleaved in db) from dy."""
        m = self.bmean if self.bn == 'train' else (self.mean if self.bn else None)
        sd = self.bstd if self.bn == 'train' else (self.std if self.bn else None)
        dalpha = C.c_void_p(db.data_ptr() + 4) if self.prelu else None
        self.h.call('e2_affine_act_bwd', C.byref(self.d), self.v.ptr(), dy.ptr(), _lib.ptr(self.scale), _lib.ptr(self.shift),
                    self.alpha, _lib.ptr(m), _lib.ptr(sd), dv.ptr(), _lib.ptr(dgamma), _lib.ptr(db), dalpha,
                    _lib.ptr(self.scratch), self.h.stream())


class MfpOp(object):
    """computations.fragmentpool (computations.py:652-678)."""

    def __init__(self, h, x, y, pool, bias=None, act='lin', keep_argmax=True, round_tf32=False):
        self.h, self.x, self.y, self.bias = h, x, y, bias
        self.pool = tuple(int(v) for v in pool)
        self.d = _lib.MfpDesc(x.desc, y.desc, self.pool[0], self.pool[1], self.pool[2], ACT[act],
                              1 if bias is not None else 0, int(bool(round_tf32)))
        self.argmax = y.like(dtype=torch.int32) if keep_argmax else None

    def fwd(self):
        self.h.call('e2_mfp_fwd', C.byref(self.d), self.x.ptr(), _lib.ptr(self.bias), self.y.ptr(),
                    self.argmax.ptr() if self.argmax is not None else None, self.h.stream())

    def bwd(self, dy, dx):
        self.h.call('e2_mfp_bwd', C.byref(self.d), dy.ptr(), self.argmax.ptr(), dx.ptr(), self.h.stream())


class Frag2DenseOp(object):
    """computations.fragments2dense (computations.py:681-701)."""

    def __init__(self, h, frag, dense, offsets, strides):
        self.h, self.frag, self.dense = h, frag, dense
        st = tuple(int(v) for v in strides)
        self.d = _lib.F2DDesc(frag.desc, dense.desc, st[0], st[1], st[2])
        self.offsets = torch.tensor([[int(v) for v in o] for o in offsets], dtype=torch.int32,
                                    device=frag.buf.device).contiguous()

    def fwd(self):
        self.h.call('e2_frag2dense_fwd', C.byref(self.d), self.frag.ptr(), _lib.ptr(self.offsets), self.dense.ptr(),
                    self.h.stream())

    def bwd(self, ddense, dfrag):
        self.h.call('e2_frag2dense_bwd', C.byref(self.d), ddense.ptr(), _lib.ptr(self.offsets), dfrag.ptr(),
                    self.h.stream())


class CropConcatOp(object):
    """Crop (neural.py:1152-1168) writing straight into a channel slice of a Concat
    buffer (node_basic.py:1403-1451)."""

    def __init__(self, h, src, dst, crop, dst_c0=0):
        self.h, self.src, self.dst = h, src, dst
        c = tuple(int(v) for v in crop)
        self.d = _lib.CropDesc(src.desc, dst.desc, c[0], c[1], c[2], int(dst_c0), 0)

    def fwd(self):
        self.h.call('e2_crop_concat_fwd', C.byref(self.d), self.src.ptr(), self.dst.ptr(), self.h.stream())

    def bwd(self, ddst, dsrc, accumulate=False, relu_gate=None):
        d = _lib.CropDesc()
        C.memmove(C.byref(d), C.byref(self.d), C.sizeof(d))
        d.accumulate = int(accumulate)
        self.h.call('e2_crop_concat_bwd', C.byref(d), ddst.ptr(), dsrc.ptr(),
                    relu_gate.ptr() if relu_gate is not None else None, self.h.stream())


class LossOp(object):
    """Softmax -> MultinoulliNLL(sparse) -> AggregateLoss mean, and Errors
    (computations.py:170-177; loss.py:261-347, 1346-1363, 730-814)."""

    EPS = 1e-5

    def __init__(self, h, logits, target, probs):
        self.h, self.logits, self.target, self.probs = h, logits, target, probs
        self.scalars = _dev_f32(4, logits.buf.device)
        self.weight = 1.0     # AggregateLoss mixing weight of this (single) component: loss = w * mean(nll) (loss.py:1355-1363)

    def fwd(self):
        self.h.call('e2_softmax_nll_fwd', C.byref(self.logits.desc), self.logits.ptr(),
                    self.target.ptr() if self.target is not None else None,
                    self.probs.ptr(), _lib.ptr(self.scalars), self.h.stream())

    def bwd(self, dlogits, grad_scale=1.0):
        self.h.call('e2_softmax_nll_bwd', C.byref(self.logits.desc), self.probs.ptr(), self.target.ptr(),
                    _lib.ptr(self.scalars), C.c_float(grad_scale), dlogits.ptr(), self.h.stream())

    def read(self):
        """(loss, error_rate) -- one D2H copy of 4 floats."""
        s = self.scalars.cpu().numpy()
        n_pos = self.logits.desc.positions
        return float(self.weight * s[0] / (s[1] + self.EPS)), float(s[2] / n_pos)

    def read_async(self):
        """Enqueue the D2H copy of the 4 scalars behind the work submitted so far; ``read_wait`` blocks on that
        copy only, so kernels enqueued afterwards (optimiser, weight re-pack) overlap with the host."""
        if getattr(self, '_host', None) is None:
            self._host = torch.empty(4, dtype=torch.float32).pin_memory()
            self._ev = torch.cuda.Event()
        self._host.copy_(self.scalars, non_blocking=True)
        self._ev.record()

    def read_wait(self):
        self._ev.synchronize()
        s = self._host.numpy()
        n_pos = self.logits.desc.positions
        return float(self.weight * s[0] / (s[1] + self.EPS)), float(s[2] / n_pos)


def act_bwd(h, t, act, y, dy, dpre):
    h.call('e2_act_bwd', C.byref(t.desc), _lib.i32(ACT[act]), y.ptr(), dy.ptr(), dpre.ptr(), h.stream())
