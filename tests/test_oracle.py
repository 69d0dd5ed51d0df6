"""CPU tests of the oracle itself: the direct restatement vs the Theano-shaped
decomposition vs an independent torch-CPU float64 implementation, the known answers
the reference holds (np.convolve, docs shape printouts), and the committed goldens."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ops, shapes, nets, loss as ol, adam as oadam, tiling

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, 'golden')
rng = np.random.RandomState(7)


def test_conv_three_way():
    x = rng.rand(2, 3, 6, 9, 8)
    w = rng.randn(4, 3, 2, 3, 4)
    y = ops.conv3d(x, w)
    assert np.allclose(y, ops.conv3d_theano_shaped(x, w), atol=1e-12)
    yt = F.conv3d(torch.tensor(x), torch.tensor(w).flip(2, 3, 4)).numpy()
    assert np.allclose(y, yt, atol=1e-12)


def test_conv_equals_np_convolve():
    """The reference's only non-Theano known answer: tests/test_conv.py:89-104."""
    x = rng.rand(1, 1, 1, 1, 60)
    w = rng.randn(1, 1, 1, 1, 7)
    assert np.allclose(ops.conv3d(x, w)[0, 0, 0, 0], np.convolve(x[0, 0, 0, 0], w[0, 0, 0, 0], 'valid'), atol=1e-13)


def test_conv_1x1_dot_shortcut():
    x = rng.rand(2, 5, 3, 4, 4)
    w = rng.randn(6, 5, 1, 1, 1)
    assert np.allclose(ops.conv3d_dot(x, w), ops.conv3d(x, w), atol=1e-12)


def test_conv_grads_vs_autograd():
    x = rng.rand(2, 3, 5, 7, 6)
    w = rng.randn(4, 3, 2, 3, 3)
    xt, wt = torch.tensor(x, requires_grad=True), torch.tensor(w, requires_grad=True)
    y = F.conv3d(xt, wt.flip(2, 3, 4))
    dy = rng.randn(*y.shape)
    y.backward(torch.tensor(dy))
    assert np.allclose(ops.conv3d_dgrad(dy, w, x.shape), xt.grad.numpy(), atol=1e-11)
    assert np.allclose(ops.conv3d_wgrad(dy, x, w.shape), wt.grad.numpy(), atol=1e-11)
    assert np.allclose(ops.bias_grad(dy), dy.sum((0, 2, 3, 4)))


@pytest.mark.parametrize('pool', [(1, 2, 2), (2, 2, 2), (1, 4, 4)])
def test_upconv_three_way(pool):
    x = rng.rand(2, 3, 3, 4, 5)
    w = rng.randn(4, 3, *pool)
    a = ops.upconv3d(x, w, pool)
    assert np.allclose(a, ops.upconv3d_theano_shaped(x, w, pool), atol=1e-12)
    c = F.conv_transpose3d(torch.tensor(x), torch.tensor(w).transpose(0, 1), stride=pool).numpy()
    assert np.allclose(a, c, atol=1e-12)
    xt, wt = torch.tensor(x, requires_grad=True), torch.tensor(w, requires_grad=True)
    y = F.conv_transpose3d(xt, wt.transpose(0, 1), stride=pool)
    dy = rng.randn(*y.shape)
    y.backward(torch.tensor(dy))
    assert np.allclose(ops.upconv3d_dgrad(dy, w, pool), xt.grad.numpy(), atol=1e-11)
    assert np.allclose(ops.upconv3d_wgrad(dy, x, pool), wt.grad.numpy(), atol=1e-11)


@pytest.mark.parametrize('pool', [(1, 2, 2), (2, 2, 2), (2, 1, 1), (2, 3, 2)])
def test_pool_forms_and_argmax(pool):
    x = rng.rand(2, 3, 4, 6, 8).astype(np.float32)
    y = ops.pooling(x, pool)
    assert np.array_equal(y, ops.pooling_theano_shaped(x, pool))
    yt, idx = F.max_pool3d(torch.tensor(x), pool, return_indices=True)
    assert np.array_equal(y, yt.numpy())
    assert np.array_equal(ops.pooling_argmax(x, pool), idx.numpy().astype(np.int32))


def test_pool_bwd_tie_modes():
    x = rng.rand(1, 2, 4, 4, 4).astype(np.float32)
    x[0, 0, :2, :2, :2] = 0.75  # a full tie window
    dy = rng.randn(1, 2, 2, 2, 2)
    first = ops.pooling_bwd(dy, x, (2, 2, 2), 'first')
    allm = ops.pooling_bwd(dy, x, (2, 2, 2), 'all')
    assert np.isclose(first.sum(), dy.sum())
    assert np.count_nonzero(first[0, 0, :2, :2, :2]) == 1 and first[0, 0, 0, 0, 0] == dy[0, 0, 0, 0, 0]
    assert np.all(allm[0, 0, :2, :2, :2] == dy[0, 0, 0, 0, 0])
    # away from ties both rules agree with autograd
    xt = torch.tensor(x[:, 1:].astype(np.float64), requires_grad=True)
    F.max_pool3d(xt, (2, 2, 2)).backward(torch.tensor(dy[:, 1:]))
    assert np.allclose(first[:, 1:], xt.grad.numpy()) and np.allclose(allm[:, 1:], xt.grad.numpy())


def test_mfp_is_dense_pool_deinterleaved():
    """MFP == stride-1 max-pool, de-interleaved; fragments2dense restores it."""
    x = rng.rand(1, 2, 7, 9, 11)
    fr, off, st = ops.fragmentpool(x, (2, 2, 2), [[0, 0, 0]], [1, 1, 1])
    assert fr.shape == (8, 2, 3, 4, 5) and list(st) == [2, 2, 2]
    assert off.tolist() == [[i, j, k] for i in range(2) for j in range(2) for k in range(2)]
    dense = F.max_pool3d(torch.tensor(x), (2, 2, 2), stride=1).numpy()
    d2 = ops.fragments2dense(fr, off, st)
    assert np.array_equal(d2, dense[:, :, :6, :8, :10])
    # second layer: new-offset-major, old-fragment-minor
    fr2, off2, st2 = ops.fragmentpool(fr[:, :, :, :3, :3], (1, 2, 2), off, st)
    assert fr2.shape[0] == 32 and list(st2) == [2, 4, 4]
    assert off2[:8].tolist() == off.tolist() and off2[8].tolist() == [0, 0, 2]
    back = ops.fragments2dense_bwd(d2, off, st)
    assert np.array_equal(back, fr)


def test_mfp_conv_is_dilated_conv():
    """conv on fragments == dilated conv on the dense map."""
    x = rng.rand(1, 2, 9, 9, 9)
    w = rng.randn(3, 2, 2, 2, 2)
    fr, off, st = ops.fragmentpool(x, (2, 2, 2), [[0, 0, 0]], [1, 1, 1])
    yf = ops.conv3d(fr, w)
    dense = ops.fragments2dense(yf, off, st)
    pooled = F.max_pool3d(torch.tensor(x), (2, 2, 2), stride=1)
    yd = F.conv3d(pooled, torch.tensor(w).flip(2, 3, 4), dilation=2).numpy()
    assert np.allclose(dense, yd[:, :, :dense.shape[2], :dense.shape[3], :dense.shape[4]], atol=1e-12)


def test_loss_matches_autograd():
    logits = rng.randn(2, 3, 2, 4, 4)
    target = rng.randint(0, 3, (2, 1, 2, 4, 4)).astype(np.float32)
    target[0, 0, 0, 0, 0] = -1  # unlabelled
    L, dl, p = ol.loss_and_dlogits(logits, target)
    lt = torch.tensor(logits, requires_grad=True)
    pt = torch.softmax(lt, 1)
    onehot = torch.tensor((target == np.arange(3).reshape(1, 3, 1, 1, 1)).astype(np.float64))
    nll = -(onehot * torch.log(pt + 1e-5)) * pt.numel() / (onehot.sum() + 1e-5) / 3
    loss = nll.sum(1, keepdim=True).mean()
    loss.backward()
    assert np.isclose(L, loss.item()) and np.allclose(dl, lt.grad.numpy(), atol=1e-12)


def test_adam_reference_formula():
    p, g = [rng.randn(5)], [rng.randn(5)]
    st = oadam.AdamState(p)
    out = oadam.adam_step(p, g, st, [True], lr=1e-3, mom=0.9, beta2=0.999, wd=0.5e-4)
    m = 0.1 * g[0]
    s = 0.001 * g[0] ** 2
    factor = np.sqrt(1 - 0.999) / (1 - 0.9)
    exp = p[0] - 1e-3 * (factor * m / np.sqrt(s + 1e-5) + 0.5e-4 * p[0])
    assert np.allclose(out[0], exp)


def test_golden_reference_python_pieces():
    """Fixtures produced by the reference's own TaggedShape / cnncalculator / initweights
    (tests/golden/make_golden.py)."""
    g = json.load(open(os.path.join(GOLD, 'ref_python.json')))
    d = g['cnncalc_docstring']  # utils/cnncalculator.py:241-260
    assert d['patch_size'] == [10, 210, 210] and d['pred_stride'] == [4, 8, 8] and d['offset'] == [4.5, 11.5, 11.5]
    # non-MFP nets: reference cnncalculator output sizes / strides == oracle shape algebra
    for name in ('neuro3d_lite', 'neuro3d'):
        c = g['cnncalc_%s_mfp0' % name]
        net = nets.BUILDERS[name]()
        convs = [n for n in net.nodes if n.op == 'conv']
        assert [list(n.sh.spatial) for n in convs] == [list(v) for v in zip(*c['pool_out'])]
        assert convs[-1].sh.strides == c['pred_stride']
        assert convs[-1].sh.fov == [f[-1] for f in c['fields']]
    # MFP patch snapping (docs/examples.rst:201-209): (23,185,185) -> (22,184,184)
    assert g['cnncalc_neuro3d_mfp1']['patch_size'] == [22, 184, 184]
    net = nets.neuro3d((22, 184, 184), mfp=True)
    assert net.nodes[-1].sh.shape == [1, 2, 8, 80, 80]
    # initweights: same seed, same rule -> same numbers
    iw = g['initweights_seed2']
    r = np.random.RandomState(2)
    w = nets.glorot_normal((20, 1, 1, 4, 4), (1, 2, 2), r)
    assert np.allclose(w.ravel()[:8], iw['conv_w']['head'], rtol=1e-6)
    assert np.allclose(nets.bias_init(20, (1, 4, 4), 'relu', r)[:3], iw['relu_b'])
    assert np.allclose(nets.bias_init(2, (1, 1, 1), 'lin', r), iw['lin_b'])


def test_docs_known_answers():
    """docs/examples.rst:100-140: Conv(10x3x23x183x183, 32,(1,6,6),(1,2,2)) -> (10,32,23,89,89),
    3,488 params, 25.2 GOps; second layer (10,64,10,42,42), 294,976 params, 416.2 GOps."""
    s0 = shapes.Sh(10, 3, (23, 183, 183))
    s1 = shapes.conv_shape(s0, 32, (1, 6, 6), (1, 2, 2))
    assert s1.shape == [10, 32, 23, 89, 89]
    assert 32 * 3 * 36 + 32 == 3488
    assert round(shapes.conv_macs(s0, 32, (1, 6, 6), 10) / 1e9, 1) == 25.2
    s2 = shapes.conv_shape(s1, 64, (4, 6, 6), (2, 2, 2))
    assert s2.shape == [10, 64, 10, 42, 42]
    assert 64 * 32 * 4 * 36 + 64 == 294976
    assert round(shapes.conv_macs(s1, 64, (4, 6, 6), 10) / 1e9, 1) == 416.2
    # docs/predictions.rst:62-79: neuro3d on (1,32,160,160) -> (2,18,56,56)
    n3 = nets.neuro3d()
    o = n3.nodes[-1].sh
    _, prob_sh, pred_sh, nt = tiling.tile_grid((32, 160, 160), (23, 185, 185), o.spatial, o.strides, o.offsets)
    assert list(pred_sh) == [18, 56, 56]


def test_golden_ops_fixture_is_current():
    """ops_small.npz must equal what the oracle computes today."""
    g = np.load(os.path.join(GOLD, 'ops_small.npz'))
    y, (lin, _), _ = ops.conv_node_fwd(g['conv_x'], g['conv_w'], g['conv_b'], (1, 2, 2), 'relu')
    assert np.array_equal(lin, g['conv_lin']) and np.array_equal(y, g['conv_y_pool122_relu'])
    assert np.array_equal(ops.pooling_argmax(g['pool_x'], (2, 2, 2)), g['pool_idx222'])
    fr, off, st = ops.fragmentpool(g['mfp_x'], (2, 2, 2), [[0, 0, 0]], [1, 1, 1])
    assert np.array_equal(fr, g['mfp_y']) and np.array_equal(off, g['mfp_off'])


def test_nets_param_counts_and_shapes():
    """parameter counts and output shapes of the four configs."""
    expect = dict(neuro3d_lite=(885132, [1, 2, 4, 30, 30]), neuro3d=(2756042, [1, 2, 5, 21, 21]),
                  unet3d_litelite=(352207, [1, 2, 10, 36, 36]), unet3d=(19069058, [1, 2, 28, 44, 44]))
    for name, (npar, osh) in expect.items():
        net = nets.BUILDERS[name]()
        assert sum(v.size for n in net.nodes for v in n.params.values()) == npar
        assert net.nodes[-1].sh.shape == osh


def test_tiling_oracle_identity_net():
    """predict_dense with a 'network' that returns its centre crop reproduces the input."""
    raw = rng.randint(0, 256, (1, 20, 30, 30)).astype(np.uint8)
    patch, out_sp, strides, offsets = (9, 13, 13), (5, 9, 9), (1, 1, 1), (2, 2, 2)

    def fwd(p):
        return p[:, :, 2:-2, 2:-2, 2:-2]
    pred = tiling.predict_dense(fwd, raw, patch, out_sp, strides, offsets, 1)
    assert np.allclose(pred, raw[:, 2:-2, 2:-2, 2:-2].astype(np.float32) / 255)
    pred8 = tiling.predict_dense(fwd, raw, patch, out_sp, strides, offsets, 1, as_uint8=True)
    assert np.array_equal(pred8, (raw[:, 2:-2, 2:-2, 2:-2].astype(np.float32) / 255 * 255).astype(np.uint8))


# ---------------------------------------------------------------- SURVEY 8f-4 restatements vs torch-CPU float64
def test_f4_batchnorm_prelu_pool_modes_maxout_match_torch():
    torch = pytest.importorskip("torch")
    import torch.nn.functional as TF
    rs = np.random.RandomState(0)
    v = rs.randn(2, 5, 4, 6, 6)
    gamma, b, dy = rs.rand(5) + 0.5, rs.randn(5), rs.randn(2, 5, 4, 6, 6)
    r = lambda a: a.reshape(1, -1, 1, 1, 1)
    tv = torch.tensor(v, requires_grad=True)
    tg = torch.tensor(gamma, requires_grad=True)
    tb = torch.tensor(b, requires_grad=True)
    mean = tv.mean(dim=(0, 2, 3, 4))
    std = tv.std(dim=(0, 2, 3, 4), unbiased=False) + 1e-6            # T.std + 1e-6, neural.py:683-684
    pre = (r(tg) / r(std)) * tv + r(tb) - r(tg) * r(mean) / r(std)      # neural.py:711
    (pre * torch.tensor(dy)).sum().backward()
    m, s = ops.batchnorm_stats(v)
    assert np.allclose(ops.batchnorm_affine(v, gamma, b, m, s), pre.detach().numpy(), atol=1e-12)
    dv, dg, db = ops.batchnorm_bwd(dy, v, gamma, m, s, True)
    assert np.allclose(dv, tv.grad.numpy(), atol=1e-12) and np.allclose(dg, tg.grad.numpy(), atol=1e-11)
    assert np.allclose(db, tb.grad.numpy(), atol=1e-11)
    # prelu == T.nnet.relu(x, alpha) == leaky relu with a per-feature slope
    alpha = rs.uniform(-0.5, 0.9, 5)
    ta = torch.tensor(alpha, requires_grad=True)
    tp = torch.tensor(v, requires_grad=True)
    y = torch.where(tp > 0, tp, r(ta) * tp)
    (y * torch.tensor(dy)).sum().backward()
    assert np.allclose(ops.prelu(v, alpha), y.detach().numpy(), atol=1e-12)
    dpre, dalpha = ops.prelu_bwd(dy, v, alpha)
    assert np.allclose(dpre, tp.grad.numpy(), atol=1e-12) and np.allclose(dalpha, ta.grad.numpy(), atol=1e-11)
    # pooling modes
    x = torch.tensor(v)
    assert np.allclose(ops.pooling_mode(v, (2, 3, 2), 'average_inc_pad'), TF.avg_pool3d(x, (2, 3, 2)).numpy(), atol=1e-12)
    assert np.allclose(ops.pooling_mode(v, (2, 3, 2), 'sum'), TF.avg_pool3d(x, (2, 3, 2)).numpy() * 12, atol=1e-12)
    assert np.array_equal(ops.pooling_mode(v, (2, 3, 2), 'max'), ops.pooling(v, (2, 3, 2)))
    d = rs.randn(2, 5, 2, 2, 3)
    xg = torch.tensor(v, requires_grad=True)
    (TF.avg_pool3d(xg, (2, 3, 2)) * torch.tensor(d)).sum().backward()
    assert np.allclose(ops.pooling_mode_bwd(d, v.shape, (2, 3, 2), 'average'), xg.grad.numpy(), atol=1e-12)
    # maxout: max over strided feature slices (computations.py:481-493)
    z = rs.randn(2, 6, 3, 4, 4)
    assert np.array_equal(ops.maxout(z, 3, 1), z.reshape(2, 2, 3, 3, 4, 4).max(axis=2))
    assert np.array_equal(ops.maxout(z, 3, 2), np.maximum(np.maximum(z[:, :, 0::3], z[:, :, 1::3]), z[:, :, 2::3]))


def test_f4_oracle_net_gradients_by_finite_differences():
    """The oracle network with batch norm / prelu / abs / average pooling: analytic gradients == central differences."""
    o = nets.Net(3)
    n = o.input((2, 1, 4, 10, 10))
    a = o.conv(n, 3, (1, 3, 3), (1, 2, 2), bn='train')
    a = o.conv(a, 4, (2, 3, 3), act='prelu')
    a = o.pool(a, (1, 2, 2), mode='average_inc_pad')
    a = o.conv(a, 3, (1, 1, 1), act='abs', bn='train')
    o.conv(a, 2, (1, 1, 1), act='lin')
    rs = np.random.RandomState(1)
    for node, k in o.param_list():
        node.params[k] = (node.params[k].astype(np.float64) + 0.1 * rs.randn(*node.params[k].shape))
    x = rs.rand(2, 1, 4, 10, 10)
    t = rs.randint(0, 2, (2, 1, 3, 1, 1)).astype(np.float64)
    L, grads, _, _ = o.loss_and_grads(x, t)
    for node, k in o.param_list():
        p = node.params[k]
        idx = tuple(rs.randint(0, s) for s in p.shape)
        old = p[idx]
        p[idx] = old + 1e-6
        Lp = o.loss_and_grads(x, t)[0]
        p[idx] = old - 1e-6
        Lm = o.loss_and_grads(x, t)[0]
        p[idx] = old
        num = (Lp - Lm) / 2e-6
        assert abs(num - grads[(node, k)][idx]) <= 1e-5 * max(1.0, abs(num)), (node.name, k, num, grads[(node, k)][idx])
