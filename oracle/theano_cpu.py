"""CPU baseline: the algorithm Theano's CPU backend runs for this path, restated on
torch-CPU float32 (TEST / BENCH INFRASTRUCTURE -- see oracle/__init__.py).

Theano itself cannot be installed here, so ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs time this "Theano-equivalent CPU
restatement, not Theano"  on the GPU host's own cores:

  * 3-D conv  = ``conv3d2d.conv3d`` decomposition (computations.py:406-428): (b*Z) 2-D
    convolutions against (f_out*kz) flipped filters (MKL/oneDNN sgemm standing in for
    Theano's CorrMM), then the diagonal sum over (z, kz) -- it computes Z*kz 2-D convs
    of which only (Z-kz+1)*kz are used, exactly like the reference;
  * 1x1x1 conv = tensordot over the feature axis (computations.py:377-384);
  * 3-D max-pool = 2-D pool over (x,y) + maximum over strided z slices (:617-631);
  * UpConv     = zero-stuffing unpool + valid conv (neural.py:1013-1020);
  * backward   = autograd through exactly these ops (the reference uses T.grad);
  * Adam       = optimiser.py:301-324.
``direct=True`` swaps the conv for ``F.conv3d`` (the friendlier second baseline).
"""
import time

import numpy as np
import torch
import torch.nn.functional as F


def conv3d2d(x, w):
    b, c, Z, X, Y = x.shape
    o, _, kz, kx, ky = w.shape
    if (kz, kx, ky) == (1, 1, 1):
        y = torch.tensordot(x, w[:, :, 0, 0, 0].t(), dims=([1], [0]))  # (b,z,x,y,o)
        return y.permute(0, 4, 1, 2, 3)
    Zo = Z - kz + 1
    x2 = x.permute(0, 2, 1, 3, 4).reshape(b * Z, c, X, Y)
    w2 = w.permute(0, 2, 1, 3, 4).reshape(o * kz, c, kx, ky).flip(2, 3)
    y2 = F.conv2d(x2, w2)
    Xo, Yo = y2.shape[2:]
    y2 = y2.view(b, Z, o, kz, Xo, Yo)
    y = None
    for i in range(kz):
        t = y2[:, kz - 1 - i:kz - 1 - i + Zo, :, i]
        y = t if y is None else y + t
    return y.permute(0, 2, 1, 3, 4)


def pool3d(x, pool):
    pz, px, py = pool
    if (pz, px, py) == (1, 1, 1):
        return x
    b, c, Z, X, Y = x.shape
    y = F.max_pool2d(x.reshape(b, c * Z, X, Y), (px, py)).view(b, c, Z, X // px, Y // py)
    m = y[:, :, 0::pz]
    for z in range(1, pz):
        m = torch.maximum(y[:, :, z::pz], m)
    return m


def unpool(x, pool):
    b, c, Z, X, Y = x.shape
    pz, px, py = pool
    out = x.new_zeros(b, c, Z * pz + pz - 1, X * px + px - 1, Y * py + py - 1)
    out[:, :, pz - 1:Z * pz:pz, px - 1:X * px:px, py - 1:Y * py:py] = x
    return out


class TorchNet(object):
    """Runs an ``oracle.nets.Net`` graph with the ops above."""

    def __init__(self, net, direct=False):
        self.net, self.direct = net, direct
        self.params = {}
        for n in net.nodes:
            for k, v in n.params.items():
                self.params[(n, k)] = torch.tensor(np.asarray(v, np.float32), requires_grad=True)
        self.m = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.s = {k: torch.zeros_like(v) for k, v in self.params.items()}
        self.t = 0

    def conv(self, x, w):
        if self.direct:
            return F.conv3d(x, w.flip(2, 3, 4))
        return conv3d2d(x, w)

    def forward(self, x):
        val = {}
        for n in self.net.nodes:
            p = [val[q] for q in n.parents]
            if n.op == 'input':
                v = x
            elif n.op == 'conv':
                v = pool3d(self.conv(p[0], self.params[(n, 'w')]), n.kw['pool'])
                v = v + self.params[(n, 'b')].view(1, -1, 1, 1, 1)
                v = torch.relu(v) if n.kw['act'] == 'relu' else v
            elif n.op == 'pool':
                v = pool3d(p[0], n.kw['pool'])
            elif n.op == 'upconv':
                v = self.conv(unpool(p[0], n.kw['pool']), self.params[(n, 'w')])
                v = v + self.params[(n, 'b')].view(1, -1, 1, 1, 1)
                v = torch.relu(v) if n.kw['act'] == 'relu' else v
            elif n.op == 'crop':
                c = n.kw['crop']
                v = p[0][:, :, c[0]:p[0].shape[2] - c[0], c[1]:p[0].shape[3] - c[1], c[2]:p[0].shape[4] - c[2]]
            elif n.op == 'concat':
                v = torch.cat(p, 1)
            else:
                raise NotImplementedError(n.op)
            val[n] = v
        return val[self.net.nodes[-1]]

    def loss(self, logits, target):
        eps = 1e-5
        p = torch.softmax(logits, 1)
        C = p.shape[1]
        onehot = (target == torch.arange(C).view(1, C, 1, 1, 1)).float()
        nll = -(onehot * torch.log(p + eps)) * p.numel() / (onehot.sum() + eps) / C
        return nll.sum(1, keepdim=True).mean()

    def train_step(self, x, target, lr=5e-4, mom=0.9, beta2=0.999, wd=0.5e-4):
        for v in self.params.values():
            v.grad = None
        loss = self.loss(self.forward(x), target)
        loss.backward()
        self.t += 1
        factor = float(np.sqrt(1 - beta2 ** self.t) / (1 - mom ** self.t))
        with torch.no_grad():
            for k, p in self.params.items():
                g = p.grad
                self.m[k].mul_(mom).add_(g, alpha=1 - mom)
                self.s[k].mul_(beta2).addcmul_(g, g, value=1 - beta2)
                d = factor * self.m[k] / torch.sqrt(self.s[k] + 1e-5)
                if k[1] == 'w':
                    d = d + wd * p
                p.sub_(lr * d)
        return float(loss.detach())


def time_training(net_builder, in_sp, steps=1, warmup=0, threads=None, direct=False, seed=0):
    """Seconds per fwd+bwd+Adam step of ``net_builder(in_sp)`` on the host cores.
    Returns dict(seconds_per_step, voxels_per_s, cores, loss)."""
    if threads:
        torch.set_num_threads(int(threads))
    net = net_builder(tuple(in_sp))
    tn = TorchNet(net, direct=direct)
    ish = net.nodes[0].sh.shape
    osh = net.nodes[-1].sh.shape
    r = np.random.RandomState(seed)
    x = torch.tensor(r.rand(*ish).astype(np.float32))
    t = torch.tensor(r.randint(0, 2, [ish[0], 1] + osh[2:]).astype(np.float32))
    loss = None
    for _ in range(warmup):
        loss = tn.train_step(x, t)
    t0 = time.perf_counter()
    for _ in range(steps):
        loss = tn.train_step(x, t)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dict(seconds_per_step=dt, voxels_per_s=float(np.prod(ish)) / dt, cores=torch.get_num_threads(),
                loss=loss, input_shape=ish)


def time_dense_tile(net_builder, in_sp, strides, threads=None, seed=0):
    """Seconds for ONE tile of ``Node._predict_densetile`` on the host cores, the way the reference does it without
    MFP (node_basic.py:832-856): prod(strides) forward passes of the strided net on shifted crops of a
    (patch + strides - 1) tile, interleaved into the dense output.  Returns dict(seconds, voxels_per_s, cores)."""
    if threads:
        torch.set_num_threads(int(threads))
    net = net_builder(tuple(in_sp))
    tn = TorchNet(net)
    osh = net.nodes[-1].sh
    st = [int(s) for s in strides]
    tile = [p + s - 1 for p, s in zip(in_sp, st)]
    r = np.random.RandomState(seed)
    raw = torch.tensor(r.rand(1, 1, *tile).astype(np.float32))
    out_sp = [int(o) * s for o, s in zip(osh.spatial, st)]
    prob = torch.zeros([2] + out_sp)
    t0 = time.perf_counter()
    with torch.no_grad():
        for xo in range(st[1]):
            for yo in range(st[2]):
                for zo in range(st[0]):
                    cut = raw[:, :, zo:zo + in_sp[0], xo:xo + in_sp[1], yo:yo + in_sp[2]]
                    p = torch.softmax(tn.forward(cut), 1)[0]
                    prob[:, zo::st[0], xo::st[1], yo::st[2]] = p
    dt = time.perf_counter() - t0
    return dict(seconds=dt, voxels_per_s=float(np.prod(out_sp)) / dt, cores=torch.get_num_threads(),
                out_spatial=out_sp, passes=int(np.prod(st)))
