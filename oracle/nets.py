"""Float64 reference networks for end-to-end parity (TEST INFRASTRUCTURE).

A tiny define-by-run graph whose node constructors restate what the reference's
node classes do at graph-build time (shape checks, AutoMerge role assignment and
cropping, parameter initialisation) and whose forward/backward call oracle.ops.
The four builders restate ``create_model()`` of the BASELINE configs:
examples/neuro3d_lite.py:46-76, neuro3d.py:46-80, unet3d_litelite.py:55-108,
unet3d.py:58-110.
"""
import numpy as np

from . import ops, shapes, loss as oloss

F64 = np.float64


def rna_tf32(v):
    """float64 -> float32 (nearest even) -> tf32 with ``cvt.rna`` (nearest, ties away from zero), in a float64
    container: what a TF32-mode kernel of libe2b200 stores (csrc/e2_common.cuh e2_round_tf32)."""
    a = np.ascontiguousarray(np.asarray(v, np.float32)).view(np.uint32)
    return ((a + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32).astype(F64)


def trunc_tf32(v):
    """What ``tcgen05.mma.kind::tf32`` does to an fp32 operand that was not rounded by its producer: the low 13
    mantissa bits are ignored."""
    a = np.ascontiguousarray(np.asarray(v, np.float32)).view(np.uint32)
    return (a & np.uint32(0xFFFFE000)).view(np.float32).astype(F64)


def glorot_normal(w_sh, pool, rng):
    """initweights(scale='glorot', mode='normal') for conv weights
    (variables.py:219-240): std = sqrt(2 / ((n_in + n_out/prod(pool)) * prod(k)))."""
    n_out, n_in = w_sh[0], w_sh[1]
    fov = np.prod(w_sh[2:])
    s = (n_in + float(n_out) / np.prod(pool)) * fov
    return rng.normal(0, np.sqrt(2.0 / s), w_sh).astype(np.float32)


def bias_init(n_f, k, act, rng):
    """_setup_params (neural.py:174-204): relu -> const 1/prod(k); other -> U(-1e-6,1e-6)."""
    if act == 'relu':
        return (np.ones(n_f) * (1.0 / np.prod(k))).astype(np.float32)
    return rng.uniform(-1e-6, 1e-6, n_f).astype(np.float32)


class N(object):
    def __init__(self, net, op, parents, sh, **kw):
        self.net, self.op, self.parents, self.sh = net, op, parents, sh
        self.kw = kw
        self.params = {}
        self.name = kw.get('name') or '%s%d' % (op, len(net.nodes))
        net.nodes.append(self)


class Net(object):
    def __init__(self, seed=2):
        self.nodes = []
        self.rng = np.random.RandomState(seed)
        # tf32=True: emulate the operand roundings of libe2b200's TF32 mode (tests pin "the end-to-end tf32
        # deviation is the network's conditioning, not kernel error"): stored activations / gradients are
        # rounded with cvt.rna by their producer, packed weights likewise, tensor-core operands that were not
        # rounded by their producer (sums of two gradient contributions) are truncated by the MMA; the
        # CUDA-core kernels (F_in == 1 wgrad, 1x1x1 layers with <= 4 outputs) multiply exact fp32 values.
        self.tf32 = False

    def _q(self, v):
        return rna_tf32(v) if self.tf32 else v

    def _use(self, v, exact=False):
        return trunc_tf32(v) if (self.tf32 and not exact) else v

    @staticmethod
    def _c1_fwd_on_tensor_cores(w_sh, x_sh):
        """Which first-layer forward kernel libe2b200 picks (csrc/e2_conv_ffma.cu e2_conv3d_fwd): the warp-specialised
        one (9..32 taps, 8..32 channels, y extent % 4 == 0) or the older tcgen05 one (16..64 taps) round x to tf32;
        otherwise the CUDA-core kernel multiplies exact fp32 values."""
        taps, n = int(np.prod(w_sh[2:])), int(w_sh[0])
        out_pos = int(np.prod([x_sh[2 + i] - w_sh[2 + i] + 1 for i in range(3)]))
        if out_pos < 4096:
            return False
        ws = 9 <= taps <= 32 and 8 <= n <= 32 and x_sh[4] % 4 == 0
        old = 16 <= taps <= 64 and 8 <= n <= 64 and n % 4 == 0
        return ws or old

    @staticmethod
    def _c1_wgrad_on_tensor_cores(w_sh, x_sh):
        """First-layer wgrad (csrc/e2_wgrad_c1_tc.cu): <= 32 taps, <= 32 channels, y extent % 4 == 0."""
        return int(np.prod(w_sh[2:])) <= 32 and int(w_sh[0]) <= 32 and x_sh[4] % 4 == 0

    @staticmethod
    def _cuda_core_layer(n):
        """Layers libe2b200 runs on CUDA cores in fp32 even in TF32 mode (csrc/e2_conv_pw.cu)."""
        w = n.params['w']
        return n.op == 'conv' and tuple(w.shape[2:]) == (1, 1, 1) and w.shape[0] <= 4

    # -- node constructors ---------------------------------------------------
    def input(self, shape, name='raw'):
        b, f = shape[0], shape[1]
        return N(self, 'input', [], shapes.Sh(1 if b is None else b, f, shape[2:]), name=name)

    def conv(self, parent, n_f, k, pool=(1, 1, 1), act='relu', mfp=False, name=None, bn=False):
        """bn: False | 'train' | 'predict' (neural.py:206-242: gamma=1, running mean=0 / std=1)."""
        sh = shapes.conv_shape(parent.sh, n_f, k, pool, mfp)
        n = N(self, 'conv', [parent], sh, k=tuple(k), pool=tuple(pool), act=act, mfp=mfp, name=name, bn=bn)
        w_sh = (n_f, parent.sh.f) + tuple(k)
        n.params['w'] = glorot_normal(w_sh, pool, self.rng)
        if act == 'prelu':      # neural.py:186-200, variables.py:209-213: (f,2): bias 1/prod(k), slope 1
            b = np.ones((n_f, 2), np.float32) / np.prod(k)
            b[:, 1] = 1.0
            n.params['b'] = b
        else:
            n.params['b'] = bias_init(n_f, k, act, self.rng)
        if bn:
            n.params['gamma'] = np.ones(n_f, np.float32)
            n.state = dict(mean=np.zeros(n_f, F64), std=np.ones(n_f, F64))
        return n

    def pool(self, parent, pool, name=None, mode='max'):
        return N(self, 'pool', [parent], shapes.pool_shape(parent.sh, pool), pool=tuple(pool), name=name, mode=mode)

    def upconv(self, parent, n_f, pool, act='relu', name=None):
        sh = shapes.upconv_shape(parent.sh, n_f, pool)
        n = N(self, 'upconv', [parent], sh, pool=tuple(pool), act=act, name=name)
        w_sh = (n_f, parent.sh.f) + tuple(pool)
        w = glorot_normal(w_sh, pool, self.rng)
        b = bias_init(n_f, pool, act, self.rng)
        # identity_init (neural.py:977-984)
        w = w * 0.1
        s = np.arange(min(w.shape[0], w.shape[1]))
        w[s, s] = 1.0
        n.params['w'] = w.astype(np.float32)
        n.params['b'] = (b * 0.0).astype(np.float32)
        return n

    def crop(self, parent, c):
        return N(self, 'crop', [parent], shapes.crop_shape(parent.sh, c), crop=tuple(c))

    def concat(self, parents, name=None):
        sh = parents[0].sh.copy()
        sh.f = sum(p.sh.f for p in parents)
        return N(self, 'concat', list(parents), sh, name=name)

    def frag2dense(self, parent):
        return N(self, 'frag2dense', [parent], shapes.frag2dense_shape(parent.sh))

    def upconv_merge(self, parent1, parent2, upconv_n_f, name='merge'):
        """AutoMerge (neural.py:1346-1405)."""
        st1, st2 = np.array(parent1.sh.strides), np.array(parent2.sh.strides)
        if np.any(st2 // st1 < 1):
            lo, hi = parent1, parent2
        else:
            hi, lo = parent1, parent2
        unpool = np.array(lo.sh.strides) // np.array(hi.sh.strides)
        if np.any(unpool > 1):
            lo = self.upconv(lo, upconv_n_f, tuple(int(u) for u in unpool))
        crop_lo, crop_hi = [], []
        for a, b in zip(hi.sh.spatial, lo.sh.spatial):
            diff = a - b
            if diff % 2 != 0:
                raise ValueError("hi_res and lo_res maps cannot be aligned")
            if diff > 0:
                crop_hi.append(diff // 2), crop_lo.append(0)
            else:
                crop_lo.append(-diff // 2), crop_hi.append(0)
        if np.any(crop_lo):
            lo = self.crop(lo, crop_lo)
        if np.any(crop_hi):
            hi = self.crop(hi, crop_hi)
        return self.concat((lo, hi), name=name)

    # -- execution -----------------------------------------------------------
    def param_list(self):
        """(node, key) in graph order: w then b per layer."""
        out = []
        for n in self.nodes:
            for k in ('w', 'b', 'gamma'):
                if k in n.params and not (k == 'gamma' and n.kw.get('bn') != 'train'):
                    out.append((n, k))
        return out

    def forward(self, x, upto=None):
        self.val, self.aux = {}, {}
        for n in self.nodes:
            p = [self.val[q] for q in n.parents]
            if n.op == 'input':
                v = np.asarray(x, F64)
            elif n.op == 'conv' and self.tf32:
                w = n.params['w']
                xin = p[0]
                if w.shape[1] == 1:
                    # first layer: the tcgen05 kernels round their input halo, the CUDA-core one does not
                    xin = rna_tf32(xin) if self._c1_fwd_on_tensor_cores(w.shape, xin.shape) else \
                        np.asarray(xin, np.float32).astype(F64)
                else:
                    xin = self._use(xin, self._cuda_core_layer(n))
                lin = (ops.conv3d_dot if tuple(w.shape[2:]) == (1, 1, 1) else ops.conv3d)(xin, rna_tf32(w))
                if n.kw['mfp'] or any(q > 1 for q in n.kw['pool']):
                    lin = rna_tf32(lin)            # the conv kernel stores the raw accumulators rounded, the pool kernel
                    if n.kw['mfp']:                # applies max -> +bias -> act and rounds again
                        pooled = ops.fragmentpool(lin, n.kw['pool'], n.parents[0].sh.mfp_offsets, n.parents[0].sh.strides)[0]
                    else:
                        pooled = ops.pooling(lin, n.kw['pool'])
                else:
                    pooled = lin
                pre = pooled + np.asarray(n.params['b'], F64).reshape(1, -1, 1, 1, 1)
                v = rna_tf32(ops.activation(pre, n.kw['act']))
                self.aux[n] = (lin, pre)
            elif n.op == 'conv' and (n.kw.get('bn') or n.kw['act'] == 'prelu'):
                # Conv._make_output with batch normalisation / prelu (neural.py:655-712)
                w = n.params['w']
                lin = (ops.conv3d_dot if tuple(w.shape[2:]) == (1, 1, 1) else ops.conv3d)(p[0], w)
                pooled = ops.pooling(lin, n.kw['pool'])
                b = np.asarray(n.params['b'], F64)
                b0, b1 = (b[:, 0], b[:, 1]) if n.kw['act'] == 'prelu' else (b, None)
                bn = n.kw.get('bn')
                if bn == 'train':
                    mean, std = ops.batchnorm_stats(pooled)
                    n.batch = dict(mean=mean, std=std)
                elif bn == 'predict':
                    mean, std = n.state['mean'], n.state['std']
                else:
                    mean, std = np.zeros(len(b0)), np.ones(len(b0))
                gamma = n.params['gamma'] if bn else np.ones(len(b0))
                pre = ops.batchnorm_affine(pooled, gamma, b0, mean, std)
                v = ops.prelu(pre, b1) if b1 is not None else ops.activation(pre, n.kw['act'])
                self.aux[n] = (lin, pre, pooled, mean, std)
            elif n.op == 'conv':
                v, (lin, pre), _ = ops.conv_node_fwd(p[0], n.params['w'], n.params['b'], n.kw['pool'],
                                                     n.kw['act'], n.kw['mfp'], n.parents[0].sh.mfp_offsets,
                                                     n.parents[0].sh.strides)
                self.aux[n] = (lin, pre)
            elif n.op == 'pool':
                v = ops.pooling_mode(p[0], n.kw['pool'], n.kw.get('mode', 'max'))
            elif n.op == 'upconv':
                v, pre = ops.upconv_node_fwd(self._use(p[0]), self._q(n.params['w']), n.params['b'], n.kw['pool'],
                                             n.kw['act'])
                v = self._q(v)
                self.aux[n] = pre
            elif n.op == 'crop':
                v = ops.crop(p[0], n.kw['crop'])
            elif n.op == 'concat':
                v = ops.concat_f(p)
            elif n.op == 'frag2dense':
                v = ops.fragments2dense(p[0], n.parents[0].sh.mfp_offsets, n.parents[0].sh.strides)
            else:
                raise NotImplementedError(n.op)
            self.val[n] = v
            if n is upto:
                break
        return self.val[upto if upto is not None else self.nodes[-1]]

    def loss_and_grads(self, x, target, tie_mode='first'):
        """Returns loss, {(node,key): grad}, probs, and d(loss)/d(activation) per node."""
        logits = self.forward(x)
        loss, dlogits, probs = oloss.loss_and_dlogits(logits, target)
        g = {self.nodes[-1]: dlogits}
        grads = {}
        for n in reversed(self.nodes):
            if n not in g or n.op == 'input':
                continue
            dy = g[n]
            par = n.parents
            if n.op == 'conv' and len(self.aux[n]) == 5:
                lin, pre, pooled, mean, std = self.aux[n]
                bn = n.kw.get('bn')
                if n.kw['act'] == 'prelu':
                    dpre, dalpha = ops.prelu_bwd(dy, pre, np.asarray(n.params['b'], F64)[:, 1])
                else:
                    dpre, dalpha = ops.activation_bwd(dy, pre, n.kw['act']), None
                gamma = n.params['gamma'] if bn else np.ones(pre.shape[1])
                dpooled, dgamma, db = ops.batchnorm_bwd(dpre, pooled, gamma, mean, std, bn == 'train')
                grads[(n, 'b')] = np.stack([db, dalpha], 1) if dalpha is not None else db
                if bn == 'train':
                    grads[(n, 'gamma')] = dgamma
                xin = self.val[par[0]]
                dlin = ops.pooling_bwd(dpooled, lin, n.kw['pool'], tie_mode) if any(q > 1 for q in n.kw['pool']) else dpooled
                grads[(n, 'w')] = ops.conv3d_wgrad(dlin, xin, n.params['w'].shape)
                if par[0].op != 'input':
                    self._acc(g, par[0], ops.conv3d_dgrad(dlin, n.params['w'], xin.shape))
            elif n.op == 'conv':
                lin, pre = self.aux[n]
                dpre = ops.activation_bwd(dy, pre, n.kw['act'])
                grads[(n, 'b')] = ops.bias_grad(dpre)
                xin = self.val[par[0]]
                if n.kw['mfp']:
                    dlin = ops.fragmentpool_bwd(dpre, lin, n.kw['pool'], tie_mode)
                elif any(p > 1 for p in n.kw['pool']):
                    dlin = ops.pooling_bwd(dpre, lin, n.kw['pool'], tie_mode)
                else:
                    dlin = dpre
                exact = self._cuda_core_layer(n) or n.params['w'].shape[1] == 1
                xw = xin
                if self.tf32 and n.params['w'].shape[1] == 1 and self._c1_wgrad_on_tensor_cores(n.params['w'].shape, xin.shape):
                    exact, xw = False, rna_tf32(xin)     # the tensor-core first-layer wgrad rounds x while building im2col
                grads[(n, 'w')] = ops.conv3d_wgrad(self._use(dlin, exact), self._use(xw, exact), n.params['w'].shape)
                if par[0].op != 'input':
                    self._acc(g, par[0], ops.conv3d_dgrad(self._use(dlin, self._cuda_core_layer(n)),
                                                          self._q(n.params['w']), xin.shape), rounded=True)
            elif n.op == 'pool' and n.kw.get('mode', 'max') != 'max':
                self._acc(g, par[0], ops.pooling_mode_bwd(dy, self.val[par[0]].shape, n.kw['pool'], n.kw['mode']))
            elif n.op == 'pool':
                self._acc(g, par[0], ops.pooling_bwd(dy, self.val[par[0]], n.kw['pool'], tie_mode))
            elif n.op == 'upconv':
                dpre = ops.activation_bwd(dy, self.aux[n], n.kw['act'])
                grads[(n, 'b')] = ops.bias_grad(dpre)
                grads[(n, 'w')] = ops.upconv3d_wgrad(self._use(dpre), self._use(self.val[par[0]]), n.kw['pool'])
                self._acc(g, par[0], ops.upconv3d_dgrad(self._use(dpre), self._q(n.params['w']), n.kw['pool']),
                          rounded=True)
            elif n.op == 'crop':
                self._acc(g, par[0], ops.crop_bwd(dy, n.kw['crop'], self.val[par[0]].shape))
            elif n.op == 'concat':
                o = 0
                for q in par:
                    self._acc(g, q, dy[:, o:o + q.sh.f])
                    o += q.sh.f
            elif n.op == 'frag2dense':
                self._acc(g, par[0], ops.fragments2dense_bwd(dy, par[0].sh.mfp_offsets, par[0].sh.strides))
        return loss, grads, probs, g

    def _acc(self, g, node, val, rounded=False):
        """``rounded``: the contribution comes from a dgrad kernel, whose epilogue rounds (accumulated sum included)
        in TF32 mode; pool / crop / concat backward kernels add without rounding."""
        v = val if node not in g else g[node] + val
        g[node] = self._q(v) if rounded else v


# ---------------------------------------------------------------- the four configs
def neuro3d_lite(in_sp=(11, 155, 155), seed=2, mfp=False):
    n = Net(seed)
    o = n.input((None, 1) + tuple(in_sp))
    o = n.conv(o, 20, (1, 4, 4), (1, 2, 2), mfp=mfp)
    o = n.conv(o, 40, (3, 3, 3), (1, 2, 2), mfp=mfp)
    o = n.conv(o, 150, (2, 4, 4), (2, 1, 1), mfp=mfp)
    o = n.conv(o, 200, (1, 3, 3), mfp=mfp)
    o = n.conv(o, 200, (1, 3, 3), mfp=mfp)
    o = n.conv(o, 200, (1, 1, 1), mfp=mfp)
    if mfp:
        # modelload injects FragmentsToDense before the *prediction* (softmax) node,
        # i.e. after the last conv (model.py:668-689)
        o = n.conv(o, 2, (1, 1, 1), act='lin', mfp=mfp)
        o = n.frag2dense(o)
    else:
        o = n.conv(o, 2, (1, 1, 1), act='lin')
    return n


def neuro3d(in_sp=(23, 185, 185), seed=2, mfp=False):
    n = Net(seed)
    o = n.input((None, 1) + tuple(in_sp))
    o = n.conv(o, 20, (1, 6, 6), (1, 2, 2), mfp=mfp)
    o = n.conv(o, 30, (1, 5, 5), (1, 2, 2), mfp=mfp)
    o = n.conv(o, 40, (1, 5, 5), mfp=mfp)
    o = n.conv(o, 80, (4, 4, 4), (2, 1, 1), mfp=mfp)
    o = n.conv(o, 100, (3, 4, 4), mfp=mfp)
    o = n.conv(o, 100, (3, 4, 4), mfp=mfp)
    o = n.conv(o, 150, (2, 4, 4), mfp=mfp)
    o = n.conv(o, 200, (1, 4, 4), mfp=mfp)
    o = n.conv(o, 200, (1, 4, 4), mfp=mfp)
    o = n.conv(o, 200, (1, 1, 1), mfp=mfp)
    o = n.conv(o, 2, (1, 1, 1), act='lin', mfp=mfp)
    if mfp:
        o = n.frag2dense(o)
    return n


def unet3d_litelite(in_sp=(22, 140, 140), seed=2):
    n = Net(seed)
    inp = n.input((None, 1) + tuple(in_sp))
    conv0 = n.conv(inp, 20, (1, 3, 3))
    conv1 = n.conv(conv0, 20, (1, 3, 3))
    down0 = n.pool(conv1, (1, 2, 2))
    conv2 = n.conv(down0, 30, (1, 3, 3))
    conv3 = n.conv(conv2, 30, (1, 3, 3))
    down1 = n.pool(conv3, (1, 2, 2))
    conv4 = n.conv(down1, 35, (1, 3, 3))
    conv5 = n.conv(conv4, 35, (1, 3, 3))
    down2 = n.pool(conv5, (1, 2, 2))
    conv6 = n.conv(down2, 42, (3, 3, 3))
    down2b = n.pool(conv6, (1, 2, 2))
    conv7 = n.conv(down2b, 42, (3, 3, 3))
    mrg0 = n.upconv_merge(conv5, conv7, 45)
    mconv0 = n.conv(mrg0, 42, (1, 3, 3))
    mconv1 = n.conv(mconv0, 42, (1, 3, 3))
    mrg1 = n.upconv_merge(conv3, mconv1, 42)
    mconv2 = n.conv(mrg1, 35, (3, 3, 3))
    mconv3 = n.conv(mconv2, 35, (3, 3, 3))
    mrg2 = n.upconv_merge(conv1, mconv3, 30)
    mconv4 = n.conv(mrg2, 20, (3, 3, 3))
    mconv5 = n.conv(mconv4, 20, (3, 3, 3))
    n.conv(mconv5, 2, (1, 1, 1), act='lin', name='barr')
    return n


def unet3d(in_sp=(116, 132, 132), seed=2, width=1.0):
    """``width`` scales the channel counts (tests use a narrow copy; 1.0 is the
    shipped config)."""
    c = lambda v: max(2, int(round(v * width)))
    n = Net(seed)
    inp = n.input((None, 1) + tuple(in_sp))
    conv0 = n.conv(inp, c(32), (3, 3, 3))
    conv1 = n.conv(conv0, c(64), (3, 3, 3))
    down0 = n.pool(conv1, (2, 2, 2))
    conv2 = n.conv(down0, c(64), (3, 3, 3))
    conv3 = n.conv(conv2, c(128), (3, 3, 3))
    down1 = n.pool(conv3, (2, 2, 2))
    conv4 = n.conv(down1, c(128), (3, 3, 3))
    conv5 = n.conv(conv4, c(256), (3, 3, 3))
    down2 = n.pool(conv5, (2, 2, 2))
    conv6 = n.conv(down2, c(256), (3, 3, 3))
    conv7 = n.conv(conv6, c(512), (3, 3, 3))
    mrg0 = n.upconv_merge(conv5, conv7, c(512))
    mconv0 = n.conv(mrg0, c(256), (3, 3, 3))
    mconv1 = n.conv(mconv0, c(256), (3, 3, 3))
    mrg1 = n.upconv_merge(conv3, mconv1, c(256))
    mconv2 = n.conv(mrg1, c(128), (3, 3, 3))
    mconv3 = n.conv(mconv2, c(128), (3, 3, 3))
    mrg2 = n.upconv_merge(conv1, mconv3, c(128))
    mconv4 = n.conv(mrg2, c(64), (3, 3, 3))
    mconv5 = n.conv(mconv4, c(64), (3, 3, 3))
    n.conv(mconv5, 2, (1, 1, 1), act='lin', name='barr')
    return n


BUILDERS = dict(neuro3d_lite=neuro3d_lite, neuro3d=neuro3d,
                unet3d_litelite=unet3d_litelite, unet3d=unet3d)
