"""SURVEY.md 8f-4 on the GPU: batch normalisation ('train' with batch statistics and running averages, 'predict'),
'prelu' ((f,2) bias), 'abs' with its backward, average / sum pooling, the 2-D / 1-D entry points and maxout --
each against the float64 oracle functions that restate the cited reference lines (oracle/ops.py)."""
import contextlib
import io

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")

from oracle import nets as onets, ops as oops, loss as ol  # noqa: E402


def _cuda():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")


def rel(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def build_f4(bn):
    """conv+pool+BN(relu) -> prelu conv -> average pool -> abs conv with BN -> 1x1x1 lin -> softmax / NLL."""
    from elektronn2_b200 import neuromancer as nm
    np.random.seed(7)
    with contextlib.redirect_stdout(io.StringIO()):
        inp = nm.Input((2, 1, 8, 20, 20), 'b,f,z,x,y', name='raw')
        c0 = nm.Conv(inp, 8, (1, 3, 3), (1, 2, 2), batch_normalisation=bn)
        c1 = nm.Conv(c0, 12, (3, 3, 3), activation_func='prelu')
        p1 = nm.Pool(c1, (2, 1, 1), mode='average')
        c2 = nm.Conv(p1, 10, (1, 3, 3), activation_func='abs', batch_normalisation=bn)
        out = nm.Conv(c2, 2, (1, 1, 1), activation_func='lin')
        probs = nm.Softmax(out)
        target = nm.Input_like(probs, override_f=1, name='target')
        loss = nm.AggregateLoss(nm.MultinoulliNLL(probs, target, target_is_sparse=True), name='loss')
        errors = nm.Errors(probs, target, target_is_sparse=True)
        m = nm.model_manager.getmodel()
        m.designate_nodes(input_node=inp, target_node=target, loss_node=loss, prediction_node=probs,
                          prediction_ext=[loss, errors, probs])
    return m, (c0, c1, c2)


def oracle_f4(bn):
    o = onets.Net(7)
    n = o.input((2, 1, 8, 20, 20))
    a0 = o.conv(n, 8, (1, 3, 3), (1, 2, 2), bn=bn)
    a1 = o.conv(a0, 12, (3, 3, 3), act='prelu')
    q1 = o.pool(a1, (2, 1, 1), mode='average_inc_pad')
    a2 = o.conv(q1, 10, (1, 3, 3), act='abs', bn=bn)
    o.conv(a2, 2, (1, 1, 1), act='lin')
    return o, (a0, a1, a2)


def test_bn_train_prelu_abs_avgpool_net_matches_oracle():
    _cuda()
    from elektronn2_b200.config import config
    from elektronn2_b200.neuromancer import optimiser
    config.compute = 'f32'
    try:
        m, (c0, c1, c2) = build_f4('train')
        o, (a0, a1, a2) = oracle_f4('train')
        # identical parameters (and non-trivial gammas / slopes)
        rs = np.random.RandomState(11)
        for p in m.trainable_params:
            if p.name.endswith('gamma>'):
                p.set_value(rs.uniform(0.5, 1.5, p.shape).astype(np.float32))
        b1 = c1.b.get_value()
        assert b1.shape == (12, 2) and np.all(b1[:, 1] == 1.0) and np.allclose(b1[:, 0], 1.0 / 27)   # variables.py:209-213
        b1[:, 1] = rs.uniform(-0.3, 0.6, 12)
        c1.b.set_value(b1)
        plist = o.param_list()
        assert len(plist) == len(m.trainable_params)
        for (node, k), p in zip(plist, m.trainable_params):
            assert tuple(node.params[k].shape) == tuple(p.shape), (node.name, k, p.name)
            node.params[k] = p.get_value()
        x = np.random.RandomState(0).rand(2, 1, 8, 20, 20).astype(np.float32)
        t = np.random.RandomState(1).randint(0, 2, (2, 1, 3, 5, 5)).astype(np.float32)
        L, grads, probs, _ = o.loss_and_grads(x, t)
        loss, err, p = m.predict_ext(x, t)
        assert abs(loss - L) <= 1e-4 * abs(L)
        assert rel(p, probs) <= 1e-4
        g = m.gradients(x, t)
        for a, (node, k) in zip(g, plist):
            assert rel(a, grads[(node, k)]) <= 1e-3, (node.name, k, rel(a, grads[(node, k)]))
        # the running averages move only in an optimiser step: 0.9995 / 0.0005 (neural.py:695-698)
        assert np.all(c0.mean.get_value() == 0) and np.all(c0.std.get_value() == 1)
        optimiser.Optimiser.setlr(1e-3), optimiser.Optimiser.setwd(1e-2), optimiser.Optimiser.setmom(0.9)
        for opt in ('SGD', 'Adam'):
            before = (c0.mean.get_value().astype(np.float64), c0.std.get_value().astype(np.float64))
            o.forward(x)                                  # batch statistics with the current parameters
            for (node, k), p in zip(plist, m.trainable_params):
                node.params[k] = p.get_value()
            o.forward(x)
            m.trainingstep(x, t, optimiser=opt)
            assert np.allclose(c0.mean.get_value(), 0.9995 * before[0] + 0.0005 * a0.batch['mean'], rtol=1e-4, atol=1e-7)
            assert np.allclose(c0.std.get_value(), 0.9995 * before[1] + 0.0005 * a0.batch['std'], rtol=1e-5, atol=1e-7)
        # weight decay on gamma carries the multiplier 3 (apply_reg=3.0, neural.py:213; optimiser.py:150-153): SGD step 1
        # from zero momentum moved gamma by -lr * (g + wd * 3 * gamma)
        assert c0.gamma.apply_reg == 3.0
        from elektronn2_b200.neuromancer import model_manager
        model_manager.reset()
        m2, (d0, _, _) = build_f4('train')
        g0 = d0.gamma.get_value().astype(np.float64)
        gg = m2.gradients(x, t)
        idx = [i for i, p in enumerate(m2.trainable_params) if p is d0.gamma][0]
        m2.trainingstep(x, t, optimiser='SGD')
        expect = g0 - 1e-3 * (gg[idx].astype(np.float64) + 1e-2 * 3.0 * g0)
        assert np.allclose(d0.gamma.get_value(), expect, rtol=0, atol=2e-7)
    finally:
        config.compute = 'tf32'
        optimiser.Optimiser.setlr(1), optimiser.Optimiser.setwd(0)


@pytest.mark.parametrize('compute', ['f32', 'tf32'])
def test_bn_predict_mode_matches_oracle(compute):
    _cuda()
    from elektronn2_b200 import neuromancer as nm
    from elektronn2_b200.config import config
    config.compute = compute
    try:
        rs = np.random.RandomState(5)
        gamma, mean, std = (rs.uniform(0.5, 1.5, 6).astype(np.float32), rs.randn(6).astype(np.float32) * 0.2,
                            rs.uniform(0.5, 2.0, 6).astype(np.float32))
        np.random.seed(3)
        with contextlib.redirect_stdout(io.StringIO()):
            inp = nm.Input((1, 3, 6, 14, 14), 'b,f,z,x,y', name='raw')
            c0 = nm.Conv(inp, 6, (2, 3, 3), (1, 2, 2), activation_func='tanh', batch_normalisation='predict', gamma=gamma,
                         mean=mean, std=std)
        assert not c0.gamma.apply_train and set(c0.params) == {'w', 'b', 'gamma', 'mean', 'std'}
        x = rs.rand(1, 3, 6, 14, 14).astype(np.float32)
        got = c0(x)
        w, b = c0.w.get_value(), c0.b.get_value()
        pooled = oops.pooling(oops.conv3d(x, w), (1, 2, 2))
        ref = np.tanh(oops.batchnorm_affine(pooled, gamma, b, mean, std))
        assert got.shape == ref.shape
        assert rel(got, ref) <= (2e-5 if compute == 'f32' else 1e-3)
    finally:
        config.compute = 'tf32'


def test_2d_and_1d_entry_points():
    """computations.conv / pooling 2-D (:351-362, 633-639) and 1-D (:337-349, 641-645) forms and a 2-D Conv node
    (examples/mnist.py:35-37 shape): executed as 3-D with leading unit axes."""
    _cuda()
    from elektronn2_b200 import neuromancer as nm
    from elektronn2_b200.neuromancer import computations as comp
    from elektronn2_b200.config import config
    config.compute = 'f32'
    try:
        rs = np.random.RandomState(2)
        x2, w2 = rs.randn(2, 3, 17, 15).astype(np.float32), rs.randn(5, 3, 3, 4).astype(np.float32)
        y2 = comp.conv(x2, w2)
        ref = oops.conv3d(x2[:, :, None], w2[:, :, None])[:, :, 0]
        assert y2.shape == ref.shape == (2, 5, 15, 12) and rel(y2, ref) <= 2e-5
        x1, w1 = rs.randn(1, 1, 40).astype(np.float32), rs.randn(1, 1, 7).astype(np.float32)
        y1 = comp.conv(x1, w1)
        assert np.allclose(y1[0, 0], np.convolve(x1[0, 0], w1[0, 0], mode='valid'), atol=1e-5)   # tests/test_conv.py:89-104
        p2 = comp.pooling(x2[:, :, :16, :14], (2, 2), [2, 3])
        assert np.array_equal(p2, oops.pooling(x2[:, :, None, :16, :14], (1, 2, 2))[:, :, 0])
        for mode in ('average', 'average_exc_pad', 'sum'):
            pm = comp.pooling(x2[:, :, :16, :14], (2, 2), [2, 3], mode=mode)
            assert rel(pm, oops.pooling_mode(x2[:, :, None, :16, :14], (1, 2, 2), mode)[:, :, 0]) <= 1e-6
        with pytest.raises(NotImplementedError):
            comp.pooling(x2, (2, 2), [1, 3])
        # the mnist.py conv stack (2-D, 'b,f,y,x' tags, BN 'train')
        np.random.seed(4)
        with contextlib.redirect_stdout(io.StringIO()):
            inp = nm.Input((4, 1, 26, 26), 'b,f,y,x', name='raw')
            c = nm.Conv(inp, 12, (3, 3), (2, 2), batch_normalisation='train')
            c = nm.Conv(c, 36, (3, 3), (2, 2), batch_normalisation='train')
            c = nm.Conv(c, 64, (3, 3), (1, 1), batch_normalisation='train')
        assert c.shape.shape == [4, 64, 3, 3] and c.conv_dim == 2
        x = rs.rand(4, 1, 26, 26).astype(np.float32)
        got = c(x)
        assert got.shape == (4, 64, 3, 3)
        v = x[:, :, None].astype(np.float64)
        for node in [n for n in c.model.nodes.values() if type(n).__name__ == 'Conv']:
            w = node.w.get_value()[:, :, None]
            pooled = oops.pooling(oops.conv3d(v, w), (1,) + tuple(node.pool_shape))
            mu, sd = oops.batchnorm_stats(pooled)
            v = np.maximum(oops.batchnorm_affine(pooled, node.gamma.get_value(), node.b.get_value(), mu, sd), 0)
        assert rel(got, v[:, :, 0]) <= 1e-4
    finally:
        config.compute = 'tf32'


def test_maxout_and_apply_activation():
    _cuda()
    from elektronn2_b200 import functional as F
    from elektronn2_b200.neuromancer import computations as comp
    rs = np.random.RandomState(9)
    x = rs.randn(2, 12, 3, 5, 4).astype(np.float32)
    for axis, factor in ((1, 2), (1, 3), (2, 3)):
        assert np.array_equal(F.maxout(x, factor, axis), oops.maxout(x, factor, axis))
    assert np.array_equal(comp.maxout(x, 3), oops.maxout(x, 3, 2))          # the reference's default: axis 2
    assert np.array_equal(comp.apply_activation(x, 'maxout 3'), oops.maxout(x, 3, 2))
    with pytest.raises(ValueError):
        comp.maxout(x, 2, axis=3)
    dy = rs.randn(2, 6, 3, 5, 4).astype(np.float32)
    dx = F.maxout_grad(x, dy, 2, 1)
    pick = x[:, 0::2] >= x[:, 1::2]
    assert np.array_equal(dx[:, 0::2], np.where(pick, dy, 0)) and np.array_equal(dx[:, 1::2], np.where(pick, 0, dy))
    slope = rs.uniform(-0.5, 0.5, 12).astype(np.float32)
    assert rel(comp.apply_activation(x, 'prelu', slope), oops.prelu(x, slope)) <= 1e-6
    assert np.array_equal(comp.apply_activation(x, 'abs'), np.abs(x))
    for a in ('tanh', 'soft+', 'elu', 'selu', 'sigmoid', 'relu', 'lin'):
        assert rel(comp.apply_activation(x, a), oops.activation(x, a)) <= 2e-6
    # a layer with maxout cannot be built -- in the reference it raises TypeError (neural.py:650-653)
    from elektronn2_b200 import neuromancer as nm
    with contextlib.redirect_stdout(io.StringIO()):
        inp = nm.Input((1, 1, 6, 10, 10), 'b,f,z,x,y', name='raw')
        with pytest.raises(NotImplementedError, match='neural.py:650'):
            nm.Conv(inp, 8, (1, 3, 3), activation_func='maxout 2')


def test_avg_sum_pool_backward():
    _cuda()
    from elektronn2_b200 import functional as F
    rs = np.random.RandomState(4)
    dy = rs.randn(2, 7, 2, 3, 5).astype(np.float32)
    for mode in ('average_inc_pad', 'sum'):
        dx = F.pool3d_grad((2, 7, 4, 6, 5), dy, (2, 2, 1), mode)
        assert rel(dx, oops.pooling_mode_bwd(dy, (2, 7, 4, 6, 5), (2, 2, 1), mode)) <= 1e-6
