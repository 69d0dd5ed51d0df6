"""GPU parity tests: every op of the hot path, called through the C ABI
(elektronn2_b200.ops -> ctypes -> libe2b200.so), against the CPU oracle on the same
seeded inputs, plus the committed golden vectors.

Tolerances (north star): pooling / MFP / fragments->dense / crop-concat and all index
outputs are BIT-EXACT; conv / upconv forward, dgrad and wgrad must satisfy
max|got-ref| / max|ref| <= 1e-3 in TF32 mode (tcgen05 kind::tf32, fp32 accumulate) and
<= 2e-5 in F32 mode (CUDA-core FFMA; differences are summation order only).
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import ops as oo, loss as ol, adam as oadam  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
TOL = {'f32': 2e-5, 'tf32': 1e-3}


@pytest.fixture(scope='module')
def h():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from elektronn2_b200 import _lib
    return _lib.get_handle(0)


def dev(a):
    from elektronn2_b200.devtensor import DevTensor
    return DevTensor.from_numpy(np.asarray(a, np.float32))


def empty(n, c, sp):
    from elektronn2_b200.devtensor import DevTensor
    return DevTensor(n, sp[0], sp[1], sp[2], c)


def rel(got, ref):
    ref = np.asarray(ref, np.float64)
    return float(np.abs(np.asarray(got, np.float64) - ref).max() / max(np.abs(ref).max(), 1e-30))


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a, np.float32)).cuda()


# ------------------------------------------------------------------------ layout
@pytest.mark.parametrize('shape', [(1, 1, 3, 5, 7), (2, 20, 4, 9, 33), (1, 45, 2, 3, 4), (3, 64, 5, 6, 7)])
def test_layout_roundtrip(h, shape):
    a = np.random.RandomState(0).rand(*shape).astype(np.float32)
    d = dev(a)
    assert np.array_equal(d.numpy(), a)
    # the device image really is channels-last with the advertised pitch
    raw = d.buf.cpu().numpy().reshape(shape[0], shape[2], shape[3], shape[4], d.desc.c_pitch)
    assert np.array_equal(raw[..., :shape[1]], a.transpose(0, 2, 3, 4, 1))


def test_u8_scaling_bit_exact(h):
    from elektronn2_b200 import _lib
    raw = np.arange(256, dtype=np.uint8).repeat(5)[:1277]
    src = torch.from_numpy(raw).cuda()
    dst = torch.empty(raw.size, dtype=torch.float32, device='cuda')
    h.call('e2_u8_to_f32', _lib.ptr(src), _lib.ptr(dst), raw.size, _lib.C.c_float(255.0), h.stream())
    assert np.array_equal(dst.cpu().numpy(), raw.astype(np.float32) / 255)   # node_basic.py:910
    p = np.random.RandomState(1).rand(1000).astype(np.float32)
    out = torch.empty(1000, dtype=torch.uint8, device='cuda')
    h.call('e2_f32_to_u8', _lib.ptr(t(p)), _lib.ptr(out), 1000, _lib.C.c_float(255.0), h.stream())
    assert np.array_equal(out.cpu().numpy(), (p * 255).astype(np.uint8))     # node_basic.py:990-996


# -------------------------------------------------------------------------- conv
CONV_CASES = [
    # (n, c_in, spatial, c_out, k)
    (1, 1, (5, 20, 21), 20, (1, 4, 4)),      # first layer of neuro3d_lite (c_in == 1 kernel)
    (1, 1, (6, 12, 12), 32, (3, 3, 3)),      # first layer of unet3d
    (1, 1, (6, 40, 40), 32, (3, 3, 3)),      # same, large enough for the tensor-core first-layer kernel (ragged tiles)
    (2, 1, (5, 40, 45), 20, (1, 6, 6)),      # neuro3d first layer: 36 taps = two K blocks, 20 channels padded to 32
    (2, 20, (5, 12, 11), 40, (3, 3, 3)),     # odd channel counts
    (1, 40, (4, 9, 9), 150, (2, 4, 4)),      # even, asymmetric kernel
    (1, 35, (3, 8, 8), 42, (1, 3, 3)),       # litelite channels
    (1, 32, (6, 10, 18), 64, (3, 3, 3)),     # unet3d conv1 shape family
    (1, 64, (5, 9, 10), 64, (3, 3, 3)),
    (1, 200, (2, 6, 6), 200, (1, 1, 1)),     # 1x1x1 (the reference's tensordot shortcut)
    (1, 200, (2, 6, 6), 2, (1, 1, 1)),       # last layer, c_out == 2
    (1, 128, (4, 6, 6), 256, (3, 3, 3)),
    (1, 256, (5, 16, 16), 128, (3, 3, 3)),   # few tiles, long K: the z-stack kernel splits K over channel blocks
    (1, 200, (4, 15, 14), 72, (3, 3, 3)),    # same with ragged channel counts and edges
    # the tensor-core first-layer wgrad (e2_wgrad_c1_tc.cu: <= 32 taps, <= 32 output channels, y extent % 4 == 0)
    (1, 1, (7, 24, 44), 20, (3, 3, 3)),      # fewer than 32 output channels, two ragged y tiles
    (2, 1, (4, 36, 72), 32, (1, 3, 3)),      # batch 2, 9 taps, three y tiles, several x tiles
    (1, 1, (5, 19, 36), 16, (1, 4, 4)),      # 16 taps
    # the z-stacked wgrad kernel (e2_wgrad_zs_tc.cu: <= 64 output channels, kz >= 2, batch 1)
    (1, 40, (7, 11, 13), 32, (3, 3, 3)),     # one r block (N = 96), second s block partial, ragged tiles
    (1, 24, (6, 9, 9), 20, (2, 3, 3)),       # fewer than 32 output channels (clipped r block), kz = 2
    (1, 32, (9, 12, 12), 64, (4, 4, 4)),     # kz = 4: N = 256, ky = 4 fills all four M chunks
    (1, 48, (5, 10, 21), 64, (2, 4, 4)),     # kz = 2: N = 128
    (1, 32, (20, 19, 25), 64, (3, 3, 3)),    # several z / x / y tiles per split, z halo across tiles
]


@pytest.mark.parametrize('compute', ['f32', 'tf32'])
@pytest.mark.parametrize('case', CONV_CASES)
def test_conv_fwd_dgrad_wgrad(h, case, compute):
    from elektronn2_b200.ops import ConvOp
    n, ci, sp, co, k = case
    r = np.random.RandomState(hash(case) % 2**31)
    x = r.rand(n, ci, *sp).astype(np.float32)
    w = (r.randn(co, ci, *k) * np.sqrt(2.0 / (ci * np.prod(k)))).astype(np.float32)
    b = (r.randn(co) * 0.1).astype(np.float32)
    osp = [s - f + 1 for s, f in zip(sp, k)]
    xd, yd = dev(x), empty(n, co, osp)
    op = ConvOp(h, xd, yd, t(w), t(b), k, 'relu', compute)
    op.pack()
    op.fwd()
    ref = oo.activation(oo.conv3d(x, w) + b.reshape(1, -1, 1, 1, 1), 'relu')
    assert rel(yd.numpy(), ref) <= TOL[compute]
    # linear epilogue, no bias
    op.fwd(act='lin', has_bias=0)
    assert rel(yd.numpy(), oo.conv3d(x, w)) <= TOL[compute]
    dy = r.randn(n, co, *osp).astype(np.float32)
    dyd = dev(dy)
    if ci > 1:
        dxd = empty(n, ci, sp)
        op.dgrad(dyd, dxd)
        ref_dx = oo.conv3d_dgrad(dy, w, x.shape)
        assert rel(dxd.numpy(), ref_dx) <= TOL[compute]
        op.dgrad(dyd, dxd, accumulate=True)          # second consumer of the same tensor
        assert rel(dxd.numpy(), 2 * ref_dx) <= TOL[compute]
    dw = torch.zeros(w.shape, device='cuda')
    db = torch.zeros(co, device='cuda')
    op.wgrad(dyd, dw, db)
    assert rel(dw.cpu().numpy(), oo.conv3d_wgrad(dy, x, w.shape)) <= TOL[compute]
    assert rel(db.cpu().numpy(), oo.bias_grad(dy)) <= 1e-5


def test_conv_matches_np_convolve(h):
    """The reference's own known answer (tests/test_conv.py:89-104): true convolution."""
    from elektronn2_b200.ops import ConvOp
    r = np.random.RandomState(3)
    x = r.rand(1, 1, 1, 1, 300).astype(np.float32)
    w = r.randn(1, 1, 1, 1, 9).astype(np.float32)
    xd, yd = dev(x), empty(1, 1, (1, 1, 292))
    op = ConvOp(h, xd, yd, t(w), None, (1, 1, 9), 'lin', 'f32')
    op.pack()
    op.fwd()
    assert np.allclose(yd.numpy()[0, 0, 0, 0], np.convolve(x[0, 0, 0, 0], w[0, 0, 0, 0], 'valid'), atol=1e-5)


def test_conv_golden(h):
    from elektronn2_b200.ops import ConvOp, PoolOp
    g = np.load(os.path.join(HERE, 'golden', 'ops_small.npz'))
    x, w, b = g['conv_x'], g['conv_w'], g['conv_b']
    xd, lin = dev(x), empty(2, 7, g['conv_lin'].shape[2:])
    op = ConvOp(h, xd, lin, t(w), None, (2, 3, 4), 'lin', 'f32')
    op.pack()
    op.fwd()
    assert rel(lin.numpy(), g['conv_lin']) <= 2e-5
    y = empty(2, 7, g['conv_y_pool122_relu'].shape[2:])
    PoolOp(h, lin, y, (1, 2, 2), bias=t(b), act='relu').fwd()         # conv -> pool -> +bias -> relu
    assert rel(y.numpy(), g['conv_y_pool122_relu']) <= 2e-5
    dyd, dxd = dev(g['conv_dy']), empty(2, 5, x.shape[2:])
    op.dgrad(dyd, dxd)
    assert rel(dxd.numpy(), g['conv_dx']) <= 2e-5
    dw, db = torch.zeros(w.shape, device='cuda'), torch.zeros(7, device='cuda')
    op.wgrad(dyd, dw, db)
    assert rel(dw.cpu().numpy(), g['conv_dw']) <= 2e-5 and rel(db.cpu().numpy(), g['conv_db']) <= 1e-5


def test_conv_invalid_descriptor_raises(h):
    from elektronn2_b200.ops import ConvOp
    from elektronn2_b200._lib import E2InvalidError
    xd, yd = empty(1, 4, (5, 5, 5)), empty(1, 4, (4, 4, 4))   # wrong output extent for k=3
    op = ConvOp(h, xd, yd, torch.zeros(4, 4, 3, 3, 3, device='cuda'), None, (3, 3, 3), 'lin', 'f32')
    with pytest.raises(ValueError):
        op.pack()
    with pytest.raises(E2InvalidError, match="output extents"):
        op.fwd()


# ------------------------------------------------------------------------ upconv
@pytest.mark.parametrize('compute', ['f32', 'tf32'])
@pytest.mark.parametrize('case', [(1, 42, (3, 4, 4), 45, (1, 4, 4)), (2, 6, (3, 4, 5), 4, (1, 2, 2)),
                                  (1, 64, (3, 4, 5), 64, (2, 2, 2)), (1, 35, (2, 5, 5), 30, (1, 2, 2))])
def test_upconv(h, case, compute):
    from elektronn2_b200.ops import UpConvOp
    n, ci, sp, co, p = case
    r = np.random.RandomState(11)
    x = r.rand(n, ci, *sp).astype(np.float32)
    w = (r.randn(co, ci, *p) * 0.2).astype(np.float32)
    b = (r.randn(co) * 0.1).astype(np.float32)
    osp = [s * q for s, q in zip(sp, p)]
    xd, yd = dev(x), empty(n, co, osp)
    op = UpConvOp(h, xd, yd, t(w), t(b), p, 'relu', compute)
    op.pack()
    op.fwd()
    ref, _ = oo.upconv_node_fwd(x, w, b, p, 'relu')
    assert rel(yd.numpy(), ref) <= TOL[compute]
    dy = r.randn(n, co, *osp).astype(np.float32)
    dyd, dxd = dev(dy), empty(n, ci, sp)
    op.dgrad(dyd, dxd)
    assert rel(dxd.numpy(), oo.upconv3d_dgrad(dy, w, p)) <= TOL[compute]
    dw, db = torch.zeros(w.shape, device='cuda'), torch.zeros(co, device='cuda')
    op.wgrad(dyd, dw, db)
    assert rel(dw.cpu().numpy(), oo.upconv3d_wgrad(dy, x, p)) <= TOL[compute]
    assert rel(db.cpu().numpy(), oo.bias_grad(dy)) <= 1e-5


def test_upconv_writes_into_concat_slice(h):
    """UpConv output lands directly in channels [0,F) of a wider Concat buffer."""
    from elektronn2_b200.ops import UpConvOp, CropConcatOp
    r = np.random.RandomState(5)
    x = r.rand(1, 6, 2, 3, 3).astype(np.float32)
    w = (r.randn(5, 6, 1, 2, 2) * 0.3).astype(np.float32)
    skip = r.rand(1, 3, 4, 10, 10).astype(np.float32)
    cat = empty(1, 8, (2, 6, 6))
    op = UpConvOp(h, dev(x), cat.channel_slice(0, 5), t(w), None, (1, 2, 2), 'lin', 'f32')
    op.pack()
    op.fwd()
    CropConcatOp(h, dev(skip), cat, (1, 2, 2), 5).fwd()
    ref = oo.concat_f([oo.upconv3d(x, w, (1, 2, 2)), oo.crop(skip, (1, 2, 2))])   # (lo_res, hi_res) order
    assert rel(cat.numpy(), ref) <= 2e-5


# -------------------------------------------------------------------------- pool
@pytest.mark.parametrize('case', [(2, 3, (4, 6, 8), (2, 2, 2)), (1, 20, (5, 12, 10), (1, 2, 2)),
                                  (1, 64, (4, 8, 8), (2, 2, 2)), (1, 150, (4, 5, 5), (2, 1, 1)),
                                  (1, 7, (6, 9, 4), (2, 3, 2))])
def test_maxpool_bit_exact(h, case):
    from elektronn2_b200.ops import PoolOp
    n, c, sp, p = case
    r = np.random.RandomState(21)
    x = r.rand(n, c, *sp).astype(np.float32)
    x[0, 0, :p[0], :p[1], :p[2]] = 0.5          # a window of ties
    x = np.maximum(x - 0.3, 0)                  # plenty of exact zeros (post-ReLU-like ties)
    osp = [s // q for s, q in zip(sp, p)]
    xd, yd = dev(x), empty(n, c, osp)
    op = PoolOp(h, xd, yd, p)
    op.fwd()
    assert np.array_equal(yd.numpy(), oo.pooling(x, p))
    assert np.array_equal(yd.numpy(), oo.pooling_theano_shaped(x, p))
    assert np.array_equal(op.argmax.int_numpy(), oo.pooling_argmax(x, p))       # first max in (z,x,y) scan order
    dy = r.randn(n, c, *osp).astype(np.float32)
    dyd, dxd = dev(dy), empty(n, c, sp)
    op.bwd(dyd, dxd)
    assert np.array_equal(dxd.numpy(), oo.pooling_bwd(dy, x, p, 'first').astype(np.float32))
    op.bwd(dyd, dxd, accumulate=True)
    assert np.array_equal(dxd.numpy(), (2 * oo.pooling_bwd(dy, x, p, 'first')).astype(np.float32))
    op_all = PoolOp(h, xd, yd, p, tie_mode='all')
    op_all.fwd()
    op_all.bwd(dyd, dxd)
    assert np.array_equal(dxd.numpy(), oo.pooling_bwd(dy, x, p, 'all').astype(np.float32))   # Theano-CPU rule


def test_maxpool_golden_and_fused_epilogue(h):
    from elektronn2_b200.ops import PoolOp
    g = np.load(os.path.join(HERE, 'golden', 'ops_small.npz'))
    x = g['pool_x']
    xd, yd = dev(x), empty(2, 3, (2, 3, 4))
    op = PoolOp(h, xd, yd, (2, 2, 2))
    op.fwd()
    assert np.array_equal(yd.numpy(), g['pool_y222']) and np.array_equal(op.argmax.int_numpy(), g['pool_idx222'])
    dxd = empty(2, 3, (4, 6, 8))
    op.bwd(dev(g['pool_dy']), dxd)
    assert np.array_equal(dxd.numpy(), g['pool_dx_first'].astype(np.float32))
    b = np.array([0.25, -0.5, 0.125], np.float32)
    PoolOp(h, xd, yd, (2, 2, 2), bias=t(b), act='relu').fwd()
    assert np.array_equal(yd.numpy(), np.maximum(g['pool_y222'] + b.reshape(1, 3, 1, 1, 1), 0))


# conv with the max-pool in its epilogue (e2_conv3d_fwd_pool): must equal the two separate launches BIT FOR BIT --
# values, argmax (first maximum in (z,x,y) scan order, ties included) and the unpooled tensor when it is kept
POOLFUSE_CASES = [
    # (n, c_in, spatial, c_out, k, pool, own_pool, quantised)
    (1, 32, (8, 20, 26), 64, (3, 3, 3), (2, 2, 2), False, False),    # unet3d conv1 -> Pool family, ragged x / y tiles
    (1, 64, (10, 36, 20), 128, (3, 3, 3), (2, 2, 2), False, False),  # several x and y tiles, four channel chunks
    (1, 20, (3, 30, 30), 20, (1, 3, 3), (1, 2, 2), False, False),    # unet3d_litelite: flat filter, in-plane pool
    (2, 20, (5, 22, 22), 30, (1, 5, 5), (1, 2, 2), True, False),     # neuro3d: Conv carrying its pool (bias/act after), batch 2
    (1, 40, (13, 32, 31), 80, (4, 4, 4), (2, 1, 1), True, False),    # neuro3d: z-only pool, kz = 4
    (1, 32, (6, 16, 14), 40, (3, 3, 3), (2, 2, 2), True, True),      # integer data: exact ties among raw accumulators
    (1, 32, (6, 16, 14), 40, (3, 3, 3), (2, 2, 2), False, True),
    (1, 24, (4, 14, 22), 32, (1, 3, 3), (2, 2, 1), False, True),     # windows over z and x only
]


@pytest.mark.parametrize('train', [True, False])
@pytest.mark.parametrize('case', POOLFUSE_CASES)
def test_conv_pool_fused_equals_separate_launches(h, case, train):
    from elektronn2_b200.ops import ConvOp, PoolOp
    n, ci, sp, co, k, pool, own, quant = case
    r = np.random.RandomState(abs(hash(case)) % 2**31)
    if quant:
        x = r.randint(0, 2, (n, ci) + sp).astype(np.float32)
        w = r.randint(-1, 2, (co, ci) + k).astype(np.float32)
        b = r.randint(-2, 3, co).astype(np.float32)
    else:
        x = r.rand(n, ci, *sp).astype(np.float32)
        w = (r.randn(co, ci, *k) * np.sqrt(2.0 / (ci * np.prod(k)))).astype(np.float32)
        b = (r.randn(co) * 0.1).astype(np.float32)
    osp = [s - f + 1 for s, f in zip(sp, k)]
    psp = [s // q for s, q in zip(osp, pool)]
    xd = dev(x)

    def run(fused):
        yd, pd = empty(n, co, osp), empty(n, co, psp)
        if own:     # conv -> pool -> +bias -> relu (neural.py:662-712)
            op = ConvOp(h, xd, yd, t(w), None, k, 'lin', 'tf32')
            pop = PoolOp(h, yd, pd, pool, bias=t(b), act='relu', keep_argmax=train, round_tf32=True)
        else:       # Conv(relu) -> Pool
            op = ConvOp(h, xd, yd, t(w), t(b), k, 'relu', 'tf32')
            pop = PoolOp(h, yd, pd, pool, keep_argmax=train)
        op.pack()
        if fused:
            assert op.pool_fusable(pop)
            op.fwd_pool(pop, store_full=not own)
        else:
            op.fwd()
            pop.fwd()
        torch.cuda.synchronize()
        return yd.numpy(), pd.numpy(), (pop.argmax.int_numpy() if train else None), op, pop

    y0, p0, a0, _, _ = run(False)
    y1, p1, a1, op, pop = run(True)
    assert np.array_equal(p0, p1)
    if train:
        assert np.array_equal(a0, a1)
        assert np.array_equal(a0, oo.pooling_argmax(y0, pool))     # and both are the oracle's argmax of the stored tensor
    if own:
        assert not y1.any()                                        # the unpooled tensor was not written
    else:
        assert np.array_equal(y0, y1)
    ref = oo.conv3d(x, w)
    ref = (np.maximum(oo.pooling(ref, pool) + b.reshape(1, -1, 1, 1, 1), 0) if own else
           oo.pooling(np.maximum(ref + b.reshape(1, -1, 1, 1, 1), 0), pool))
    assert rel(p1, ref) <= TOL['tf32']


def test_conv_pool_fused_keeps_the_requested_window(h):
    """y_keep of e2_conv3d_fwd_pool: the window of the unpooled tensor the caller reads (a skip connection's Crop) is
    written exactly; whole tiles outside it are skipped (the buffer keeps its zeros there); pooled outputs unaffected."""
    from elektronn2_b200.ops import ConvOp, PoolOp
    r = np.random.RandomState(11)
    x = r.rand(1, 32, 10, 40, 30).astype(np.float32)
    w = (r.randn(64, 32, 3, 3, 3) * 0.05).astype(np.float32)
    b = (r.randn(64) * 0.1).astype(np.float32)
    osp, psp = (8, 38, 28), (4, 19, 14)
    xd = dev(x)
    outs = []
    for keep in (None, (2, 6, 9, 30, 6, 20), (0, 0, 0, 0, 0, 0)):
        yd, pd = empty(1, 64, osp), empty(1, 64, psp)
        op = ConvOp(h, xd, yd, t(w), t(b), (3, 3, 3), 'relu', 'tf32')
        pop = PoolOp(h, yd, pd, (2, 2, 2))
        op.pack()
        assert op.pool_fusable(pop)
        op.fwd_pool(pop, True, keep)
        outs.append((yd.numpy(), pd.numpy(), pop.argmax.int_numpy()))
    (y0, p0, a0), (y1, p1, a1), (y2, p2, a2) = outs
    assert np.array_equal(p0, p1) and np.array_equal(a0, a1) and np.array_equal(p0, p2) and np.array_equal(a0, a2)
    assert np.array_equal(y1[:, :, 2:6, 9:30, 6:20], y0[:, :, 2:6, 9:30, 6:20])
    written = y1 != 0
    assert not written[:, :, :2].any() and not written[:, :, 6:].any()       # planes outside the window: untouched
    assert not written[:, :, :, :8].any() and not written[:, :, :, 32:].any()   # x boxes of 4: [8,32) covers [9,30)
    assert np.array_equal(y1[written], y0[written])
    assert not y2.any()


def test_conv_pool_unfusable_pairs_run_as_two_launches(h):
    from elektronn2_b200.ops import ConvOp, PoolOp
    r = np.random.RandomState(5)
    x = r.rand(1, 16, 6, 14, 14).astype(np.float32)
    w = (r.randn(24, 16, 3, 3, 3) * 0.1).astype(np.float32)
    b = (r.randn(24) * 0.1).astype(np.float32)
    xd, yd, pd = dev(x), empty(1, 24, (4, 12, 12)), empty(1, 24, (2, 6, 6))
    op = ConvOp(h, xd, yd, t(w), t(b), (3, 3, 3), 'relu', 'f32')       # exact-fp32 mode never fuses
    pop = PoolOp(h, yd, pd, (2, 2, 2))
    op.pack()
    assert not op.pool_fusable(pop)
    n0 = h.launches
    op.fwd_pool(pop, store_full=True)
    assert h.launches - n0 == 2
    y = yd.numpy()
    assert np.array_equal(pd.numpy(), oo.pooling(y, (2, 2, 2))) and np.array_equal(pop.argmax.int_numpy(),
                                                                                   oo.pooling_argmax(y, (2, 2, 2)))
    with pytest.raises(ValueError):
        op.fwd_pool(pop, store_full=False)      # nothing to pool from without the unpooled tensor


def test_maxpool_rejects_non_dividing_axes(h):
    from elektronn2_b200.ops import PoolOp
    with pytest.raises(ValueError, match="cannot downsample"):
        PoolOp(h, empty(1, 4, (5, 6, 6)), empty(1, 4, (2, 3, 3)), (2, 2, 2)).fwd()


# --------------------------------------------------------------------------- MFP
@pytest.mark.parametrize('case', [(1, 3, (7, 9, 11), (2, 2, 2)), (1, 20, (5, 13, 13), (1, 2, 2)),
                                  (4, 8, (7, 7, 7), (2, 1, 1)), (1, 5, (9, 8, 11), (2, 3, 2)),
                                  # the sliding kernel with several row chunks and several segments of the walking axis
                                  (1, 30, (6, 41, 43), (1, 2, 2)), (1, 80, (37, 9, 10), (2, 1, 1)),
                                  (1, 12, (19, 21, 23), (2, 2, 2)), (1, 8, (4, 9, 6), (1, 2, 1))])
def test_mfp_bit_exact_fragment_order(h, case):
    from elektronn2_b200.ops import MfpOp, Frag2DenseOp
    n, c, sp, p = case
    r = np.random.RandomState(31)
    x = np.maximum(r.rand(n, c, *sp).astype(np.float32) - 0.2, 0)
    old_off = [[0, 0, 0]] if n == 1 else [[i, j, 0] for i in range(2) for j in range(2)]
    old_st = [1, 1, 1] if n == 1 else [2, 2, 1]
    ref, off, st = oo.fragmentpool(x, p, old_off, old_st)
    xd, yd = dev(x), empty(ref.shape[0], c, ref.shape[2:])
    op = MfpOp(h, xd, yd, p)
    op.fwd()
    assert np.array_equal(yd.numpy(), ref)           # values AND fragment order (new-offset-major)
    dy = r.randn(*ref.shape).astype(np.float32)
    dxd = empty(n, c, sp)
    op.bwd(dev(dy), dxd)
    assert rel(dxd.numpy(), oo.fragmentpool_bwd(dy, x, p, 'first')) <= 1e-6   # sums of <= prod(p) terms
    if n == 1:
        dd = empty(1, c, [s * q for s, q in zip(ref.shape[2:], st)])
        f2d = Frag2DenseOp(h, yd, dd, off, st)
        f2d.fwd()
        assert np.array_equal(dd.numpy(), oo.fragments2dense(ref, off, st))
        back = empty(ref.shape[0], c, ref.shape[2:])
        f2d.bwd(dd, back)
        assert np.array_equal(back.numpy(), ref)     # round trip


def test_mfp_two_layers_golden(h):
    from elektronn2_b200.ops import MfpOp, Frag2DenseOp
    g = np.load(os.path.join(HERE, 'golden', 'ops_small.npz'))
    xd = dev(g['mfp_x'])
    y1 = empty(8, 3, g['mfp_y'].shape[2:])
    MfpOp(h, xd, y1, (2, 2, 2)).fwd()
    assert np.array_equal(y1.numpy(), g['mfp_y'])
    x2 = dev(g['mfp_y'][:, :, :, :3, :3])
    y2 = empty(32, 3, g['mfp2_y'].shape[2:])
    MfpOp(h, x2, y2, (1, 2, 2)).fwd()
    assert np.array_equal(y2.numpy(), g['mfp2_y'])
    dd = empty(1, 3, g['f2d_y'].shape[2:])
    Frag2DenseOp(h, y1, dd, g['mfp_off'], g['mfp_st']).fwd()
    assert np.array_equal(dd.numpy(), g['f2d_y'])


def test_mfp_rule_and_f2d_count_errors(h):
    from elektronn2_b200.ops import MfpOp, Frag2DenseOp
    with pytest.raises(ValueError, match="using MFP"):                           # neural.py:739-744
        MfpOp(h, empty(1, 4, (6, 6, 6)), empty(8, 4, (2, 2, 2)), (2, 2, 2)).fwd()
    with pytest.raises(ValueError, match="fragments on the batch axis"):         # neural.py:873-875
        Frag2DenseOp(h, empty(4, 2, (2, 2, 2)), empty(1, 2, (4, 4, 4)), [[0, 0, 0]] * 4, (2, 2, 2)).fwd()


# ------------------------------------------------------------------ crop + concat
def test_crop_concat(h):
    from elektronn2_b200.ops import CropConcatOp
    r = np.random.RandomState(41)
    a = r.rand(2, 5, 6, 10, 12).astype(np.float32)
    dst = empty(2, 9, (4, 6, 6))
    op = CropConcatOp(h, dev(a), dst, (1, 2, 3), 4)
    op.fwd()
    got = dst.numpy()
    assert np.array_equal(got[:, 4:], oo.crop(a, (1, 2, 3))) and np.all(got[:, :4] == 0)
    dd = r.randn(2, 9, 4, 6, 6).astype(np.float32)
    ds = empty(2, 5, (6, 10, 12))
    op.bwd(dev(dd), ds)
    ref = oo.crop_bwd(dd[:, 4:], (1, 2, 3), a.shape).astype(np.float32)
    assert np.array_equal(ds.numpy(), ref)
    op.bwd(dev(dd), ds, accumulate=True)
    assert np.array_equal(ds.numpy(), 2 * ref)


# --------------------------------------------------------------------- loss head
def test_softmax_nll_errors(h):
    from elektronn2_b200.ops import LossOp
    r = np.random.RandomState(51)
    logits = r.randn(2, 2, 4, 6, 6).astype(np.float32) * 2
    target = r.randint(0, 2, (2, 1, 4, 6, 6)).astype(np.float32)
    target[0, 0, 0, :2, :2] = -1                       # unlabelled voxels
    ld, td, pd = dev(logits), dev(target), empty(2, 2, (4, 6, 6))
    op = LossOp(h, ld, td, pd)
    op.fwd()
    L, dl, p = ol.loss_and_dlogits(logits, target)
    loss, err = op.read()
    assert abs(loss - L) <= 1e-5 * abs(L)
    assert rel(pd.numpy(), p) <= 1e-6
    assert abs(err - ol.errors(p, target)) < 1e-7
    gd = empty(2, 2, (4, 6, 6))
    op.bwd(gd)
    assert rel(gd.numpy(), dl) <= 1e-5


# --------------------------------------------------------------------- optimisers
def test_adam_and_sgd_match_reference_formulas(h):
    from elektronn2_b200 import _lib
    r = np.random.RandomState(61)
    p0, g0 = r.randn(1001).astype(np.float32), r.randn(1001).astype(np.float32)
    p, g = t(p0), t(g0)
    m, s = torch.zeros_like(p), torch.zeros_like(p)
    st = oadam.AdamState([p0])
    ref = [p0]
    for step in (1, 2, 3):
        h.call('e2_adam_step', _lib.ptr(p), _lib.ptr(g), _lib.ptr(m), _lib.ptr(s), 1001, 5e-4, 0.9, 0.999, 0.5e-4, 1,
               step, h.stream())
        ref = oadam.adam_step(ref, [g0], st, [True], lr=5e-4, mom=0.9, beta2=0.999, wd=0.5e-4)
        assert rel(p.cpu().numpy(), ref[0]) <= 1e-6
    p, d = t(p0), torch.zeros(1001, device='cuda')
    last = [np.zeros(1001)]
    h.call('e2_sgd_step', _lib.ptr(p), _lib.ptr(g), _lib.ptr(d), 1001, 1e-2, 0.9, 1e-3, 0, h.stream())
    ref = oadam.sgd_step([p0], [g0], last, [False], lr=1e-2, mom=0.9, wd=1e-3)
    assert rel(p.cpu().numpy(), ref[0]) <= 1e-6


ADAM_PACK_CASES = [
    # (kind, f_out, f_in, filter / pool)
    ('conv', 64, 32, (3, 3, 3)), ('conv', 30, 20, (1, 5, 5)), ('conv', 20, 1, (1, 6, 6)), ('conv', 2, 200, (1, 1, 1)),
    ('conv', 150, 40, (2, 4, 4)), ('conv', 35, 42, (3, 3, 3)), ('conv', 80, 40, (4, 4, 4)),
    ('upconv', 24, 32, (2, 2, 2)), ('upconv', 45, 35, (1, 2, 2)), ('upconv', 256, 512, (2, 2, 2)),
]


@pytest.mark.parametrize('case', ADAM_PACK_CASES)
def test_adam_pack_fused_equals_update_then_pack(h, case):
    """e2_*_adam_pack_dev (one launch per layer: Adam update + both packed images) against the two steps it replaces,
    e2_adam_step_dev over the layer's slice and e2_*_pack_weights: parameters, both moments and both packed images
    must be the same bits (pad lanes included: they keep the zeros of the first pack)."""
    from elektronn2_b200 import _lib
    from elektronn2_b200.ops import ConvOp, UpConvOp
    kind, co, ci, k = case
    r = np.random.RandomState(abs(hash(case)) % 2**31)
    n = co * ci * int(np.prod(k))
    w0 = (r.randn(n) * 0.1).astype(np.float32)
    g0 = (r.randn(n) * 0.01).astype(np.float32)
    m0 = (r.randn(n) * 0.01).astype(np.float32)
    s0 = (r.rand(n) * 1e-4).astype(np.float32)
    hyper = torch.tensor([5e-4, 0.9, 0.999, 0.5e-4, 0.0, 0, 0, 0], dtype=torch.float32, device='cuda')
    tdev = torch.tensor([6], dtype=torch.int32, device='cuda')
    h.call('e2_adam_prepare', _lib.ptr(hyper), _lib.ptr(tdev), h.stream())       # factor(t = 7)

    def make(w):
        if kind == 'conv':
            sp = [f + 1 for f in k]
            return ConvOp(h, empty(1, ci, sp), empty(1, co, (2, 2, 2)), w.view(co, ci, *k), None, k, 'relu', 'tf32')
        return UpConvOp(h, empty(1, ci, (2, 2, 2)), empty(1, co, [2 * q for q in k]), w.view(co, ci, *k), None, k, 'relu',
                        'tf32')

    fn = 'e2_conv3d_adam_pack_dev' if kind == 'conv' else 'e2_upconv3d_adam_pack_dev'
    res = []
    for fused in (False, True):
        w, g, m, s = t(w0), t(g0), t(m0), t(s0)
        op = make(w)
        op.pack()                                  # first pack: zeroes the pad lanes
        if fused:
            h.call(fn, _lib.C.byref(op.d), _lib.ptr(w), _lib.ptr(g), _lib.ptr(m), _lib.ptr(s), _lib.ptr(hyper),
                   _lib.C.c_float(1.0), _lib.ptr(op.wf), _lib.ptr(op.wd), h.stream())
        else:
            h.call('e2_adam_step_dev', _lib.ptr(w), _lib.ptr(g), _lib.ptr(m), _lib.ptr(s), n, _lib.ptr(hyper),
                   _lib.C.c_float(1.0), h.stream())
            op.pack()
        torch.cuda.synchronize()
        res.append([a.cpu().numpy() for a in (w, m, s, op.wf, op.wd)])
    for a, b in zip(*res):
        assert np.array_equal(a, b)
    assert not np.array_equal(res[0][0], w0)       # the update happened


# -------------------------------------------- size-independent properties at full size
def test_full_size_properties_unet3d_conv1(h):
    """unet3d conv1 (1,32,114,130,130) -> (1,64,112,128,128), the largest layer of the
    BASELINE workload: linearity in x and agreement on a random sample of outputs with
    the oracle formula (the full oracle conv would take minutes)."""
    from elektronn2_b200.ops import ConvOp
    from elektronn2_b200.config import config
    r = np.random.RandomState(71)
    sp, k = (114, 130, 130), (3, 3, 3)
    x1 = r.rand(1, 32, *sp).astype(np.float32)
    x2 = r.rand(1, 32, *sp).astype(np.float32)
    w = (r.randn(64, 32, *k) * np.sqrt(2.0 / (32 * 27))).astype(np.float32)
    osp = (112, 128, 128)
    ys = []
    for x in (x1, x2, x1 + x2):
        xd, yd = dev(x), empty(1, 64, osp)
        op = ConvOp(h, xd, yd, t(w), None, k, 'lin', config.compute)
        op.pack()
        op.fwd()
        ys.append(yd.numpy())
    tol = TOL[config.compute]
    assert rel(ys[2], ys[0] + ys[1]) <= 2 * tol
    wf = w[:, :, ::-1, ::-1, ::-1].astype(np.float64)
    for _ in range(64):
        o, z, xx, yy = r.randint(64), r.randint(112), r.randint(128), r.randint(128)
        ref = float((x1[0, :, z:z + 3, xx:xx + 3, yy:yy + 3].astype(np.float64) * wf[o]).sum())
        assert abs(ys[0][0, o, z, xx, yy] - ref) <= tol * np.abs(ys[0]).max()


# ------------------------------------------------------------ out-of-bounds canaries
CANARY = 12345.678
GUARD = 4096          # floats on either side (multiple of 4: the tensor keeps its 16-byte alignment)


def guarded(n, c, sp, dtype=torch.float32):
    """A DevTensor placed between two canary zones of one larger buffer."""
    from elektronn2_b200.devtensor import DevTensor, pitch_for
    floats = n * sp[0] * sp[1] * sp[2] * pitch_for(c)
    fill = CANARY if dtype == torch.float32 else 123456789
    buf = torch.full((floats + 2 * GUARD,), fill, dtype=dtype, device='cuda')
    buf[GUARD:GUARD + floats] = 0
    return DevTensor(n, sp[0], sp[1], sp[2], c, buf=buf, offset=GUARD)


def guards_intact(tn):
    b, f = tn.buf, tn.desc.floats
    fill = CANARY if b.dtype == torch.float32 else 123456789
    ref = torch.full((GUARD,), fill, dtype=b.dtype, device=b.device)
    return bool(torch.equal(b[:GUARD], ref)) and bool(torch.equal(b[GUARD + f:], ref))


def guarded_flat(shape):
    numel = int(np.prod(shape))
    pad = (numel + 3) // 4 * 4
    buf = torch.full((pad + 2 * GUARD,), CANARY, dtype=torch.float32, device='cuda')
    view = buf[GUARD:GUARD + numel].view(*shape)
    view.zero_()
    return buf, view, numel


@pytest.mark.parametrize('case', [
    (1, 1, (6, 40, 40), 32, (3, 3, 3)),      # first-layer tensor-core kernel, ragged tiles
    (2, 1, (5, 40, 45), 20, (1, 6, 6)),      # ... two K blocks, 20 channels
    (1, 256, (5, 16, 16), 128, (3, 3, 3)),   # z-stack kernel with a K split (workspace partial tiles)
    (1, 200, (4, 15, 14), 72, (3, 3, 3)),    # ragged channels and edges
    (2, 20, (5, 12, 11), 40, (3, 3, 3)),     # tap kernel, scalar epilogue path
    (1, 200, (2, 6, 6), 2, (1, 1, 1)),       # few-output-channel kernels
    (1, 128, (4, 6, 6), 256, (3, 3, 3)),     # tap kernel with split-K
])
def test_conv_kernels_stay_inside_their_buffers(h, case):
    """compute-sanitizer is not available on the GPU pool: every output tensor sits between canary zones and the
    zones must be untouched after fwd / dgrad / wgrad (TMA stores clip at the tensor-map extents, the hand-written
    epilogues bound-check rows and channel tails)."""
    from elektronn2_b200.ops import ConvOp
    n, ci, sp, co, k = case
    r = np.random.RandomState(7)
    x = r.rand(n, ci, *sp).astype(np.float32)
    w = (r.randn(co, ci, *k) * 0.1).astype(np.float32)
    b = (r.randn(co) * 0.1).astype(np.float32)
    osp = [s - f + 1 for s, f in zip(sp, k)]
    xd, yd = dev(x), guarded(n, co, osp)
    op = ConvOp(h, xd, yd, t(w), t(b), k, 'relu', 'tf32')
    op.pack()
    op.fwd()
    torch.cuda.synchronize()
    assert guards_intact(yd)
    assert rel(yd.numpy(), oo.activation(oo.conv3d(x, w) + b.reshape(1, -1, 1, 1, 1), 'relu')) <= TOL['tf32']
    dyd = dev(r.randn(n, co, *osp).astype(np.float32))
    if ci > 1:
        dxd = guarded(n, ci, sp)
        op.dgrad(dyd, dxd)
        op.dgrad(dyd, dxd, accumulate=True)
        torch.cuda.synchronize()
        assert guards_intact(dxd)
    wbuf, dw, _ = guarded_flat(w.shape)
    bbuf, db, _ = guarded_flat(b.shape)
    op.wgrad(dyd, dw, db)
    torch.cuda.synchronize()
    for buf, numel in ((wbuf, w.size), (bbuf, b.size)):
        ref = torch.full((GUARD,), CANARY, device='cuda')
        assert torch.equal(buf[:GUARD], ref) and torch.equal(buf[GUARD + (numel + 3) // 4 * 4:], ref)


@pytest.mark.parametrize('case', [(1, 42, (3, 4, 4), 45, (1, 4, 4)), (1, 64, (3, 4, 5), 64, (2, 2, 2)),
                                  (1, 128, (6, 10, 10), 96, (2, 2, 2))])
def test_upconv_kernels_stay_inside_their_buffers(h, case):
    from elektronn2_b200.ops import UpConvOp
    n, ci, sp, co, p = case
    r = np.random.RandomState(8)
    x = r.rand(n, ci, *sp).astype(np.float32)
    w = (r.randn(co, ci, *p) * 0.2).astype(np.float32)
    b = (r.randn(co) * 0.1).astype(np.float32)
    osp = [s * q for s, q in zip(sp, p)]
    xd, yd = dev(x), guarded(n, co, osp)
    op = UpConvOp(h, xd, yd, t(w), t(b), p, 'relu', 'tf32')
    op.pack()
    op.fwd()
    dyd = dev(r.randn(n, co, *osp).astype(np.float32))
    dxd = guarded(n, ci, sp)
    op.dgrad(dyd, dxd)
    wbuf, dw, _ = guarded_flat(w.shape)
    op.wgrad(dyd, dw, None)
    torch.cuda.synchronize()
    assert guards_intact(yd) and guards_intact(dxd)
    ref = torch.full((GUARD,), CANARY, device='cuda')
    assert torch.equal(wbuf[:GUARD], ref) and torch.equal(wbuf[GUARD + (w.size + 3) // 4 * 4:], ref)


@pytest.mark.parametrize('c', [3, 30, 64])
def test_pool_and_mfp_kernels_stay_inside_their_buffers(h, c):
    """Channel counts that are not a multiple of 4 take the float4 paths through the tensors' pad lanes."""
    from elektronn2_b200.ops import PoolOp, MfpOp
    r = np.random.RandomState(9)
    x = r.rand(1, c, 6, 10, 12).astype(np.float32)
    xd, yd = dev(x), guarded(1, c, (3, 5, 6))
    op = PoolOp(h, xd, yd, (2, 2, 2))
    op.argmax = guarded(1, c, (3, 5, 6), dtype=torch.int32)
    op.fwd()
    dxd = guarded(1, c, (6, 10, 12))
    op.bwd(dev(r.randn(1, c, 3, 5, 6).astype(np.float32)), dxd)
    torch.cuda.synchronize()
    assert guards_intact(yd) and guards_intact(op.argmax) and guards_intact(dxd)
    assert np.array_equal(yd.numpy(), oo.pooling(x, (2, 2, 2)))
    xm = r.rand(1, c, 7, 9, 11).astype(np.float32)
    ym = guarded(8, c, (3, 4, 5))
    mop = MfpOp(h, dev(xm), ym, (2, 2, 2))
    mop.argmax = guarded(8, c, (3, 4, 5), dtype=torch.int32)
    mop.fwd()
    torch.cuda.synchronize()
    assert guards_intact(ym) and guards_intact(mop.argmax)
    assert np.array_equal(ym.numpy(), oo.fragmentpool(xm, (2, 2, 2), np.zeros((1, 3), int), [1, 1, 1])[0])


@pytest.mark.parametrize('act', ['tanh', 'sigmoid', 'soft+', 'elu', 'selu'])
def test_epilogue_activations_and_their_backward(h, act):
    """apply_activation (computations.py:57-134) in the fused conv epilogue (z-stack kernel shape and tap-kernel
    shape) and its derivative, which the B200 path evaluates from the OUTPUT y (the pre-activation is not kept)."""
    from elektronn2_b200.ops import ConvOp, act_bwd
    r = np.random.RandomState(11)
    for (ci, sp, co) in [(32, (5, 16, 12), 48), (20, (3, 7, 9), 24)]:
        x = (r.rand(1, ci, *sp) - 0.5).astype(np.float32)
        w = (r.randn(co, ci, 3, 3, 3) * np.sqrt(2.0 / (ci * 27))).astype(np.float32)
        b = (r.randn(co) * 0.1).astype(np.float32)
        osp = [s - 2 for s in sp]
        pre = oo.conv3d(x, w) + b.reshape(1, -1, 1, 1, 1)
        for compute in ('f32', 'tf32'):
            xd, yd = dev(x), empty(1, co, osp)
            op = ConvOp(h, xd, yd, t(w), t(b), (3, 3, 3), act, compute)
            op.pack()
            op.fwd()
            assert rel(yd.numpy(), oo.activation(pre, act)) <= TOL[compute]
        dy = r.randn(1, co, *osp).astype(np.float32)
        yref = dev(oo.activation(pre, act).astype(np.float32))
        dyd = dev(dy)
        act_bwd(h, dyd, act, yref, dyd, dyd)           # in place, as the executor does
        assert rel(dyd.numpy(), oo.activation_bwd(dy, pre, act)) <= 2e-5


def _round_tf32(a):
    """Round-to-nearest-even to 10 explicit mantissa bits (what the producers of tensor-core operands do)."""
    u = np.ascontiguousarray(a, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0xFFF + ((u >> 13) & 1)) & 0xFFFFE000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize('case', [
    (768, (14, 18, 18), 256),    # unet3d mconv0: z-stack kernel with K split (fwd), tap kernel split-K (dgrad)
    (256, (11, 13, 13), 256),    # unet3d conv6: tiny planes -> tap kernel with split-K, wgrad with many splits
    (64, (56, 64, 64), 64),      # unet3d conv2: z-stack / halo-wgrad kernels, several waves
    (512, (7, 9, 9), 512),       # upconv mrg0 (as UpConv below)
])
def test_full_size_adjoint_identities(h, case):
    """Size-independent property at the BASELINE layer sizes (the float64 oracle would take minutes here):
    forward, dgrad and wgrad are the three faces of one bilinear form,
        <dy, conv(x, w)> == <dgrad(dy, w), x> == <wgrad(dy, x), w>.
    With tf32-representable x, w, dy the tensor-core products are exact, so the three numbers differ only by fp32
    accumulation order / output rounding."""
    from elektronn2_b200.ops import ConvOp, UpConvOp
    ci, sp, co = case
    r = np.random.RandomState(5)
    up = ci == 512
    k = (2, 2, 2) if up else (3, 3, 3)
    osp = [s * 2 for s in sp] if up else [s - 2 for s in sp]
    x = _round_tf32(r.rand(1, ci, *sp).astype(np.float32) - 0.5)
    w = _round_tf32((r.randn(co, ci, *k) * np.sqrt(2.0 / (ci * np.prod(k)))).astype(np.float32))
    dy = _round_tf32(r.randn(1, co, *osp).astype(np.float32))
    xd, yd, dyd, dxd = dev(x), empty(1, co, osp), dev(dy), empty(1, ci, sp)
    op = (UpConvOp if up else ConvOp)(h, xd, yd, t(w), None, k, 'lin', 'tf32')
    op.pack()
    d = op.d
    d.has_bias = 0
    op.fwd() if up else op.fwd(act='lin', has_bias=0)
    op.dgrad(dyd, dxd)
    dw = torch.zeros(w.shape, device='cuda')
    op.wgrad(dyd, dw, None)
    y = yd.numpy().astype(np.float64)
    a = float((dy.astype(np.float64) * y).sum())
    b = float((dxd.numpy().astype(np.float64) * x.astype(np.float64)).sum())
    c = float((dw.cpu().numpy().astype(np.float64) * w.astype(np.float64)).sum())
    # y and dx are rounded to tf32 on store (relative 2^-11 per element, random sign), dw is plain fp32: with
    # nat = the natural magnitude of such an inner product (||dy|| ||y|| / sqrt(#elements)) the three numbers agree
    # to ~3e-4 * nat; a dropped filter tap (1/27) or a lost tile would show up at >= 1e-2 * nat
    nat = float(np.sqrt((dy.astype(np.float64) ** 2).sum() * (y ** 2).sum() / y.size))
    assert abs(a - b) <= 2e-3 * nat and abs(a - c) <= 2e-3 * nat and abs(b - c) <= 2e-3 * nat, (a, b, c, nat)


def test_functional_layer_and_theano_perform(h):
    """numpy in / numpy out face (elektronn2_b200.functional: the granularity of computations.py and of a Theano
    Op.perform) against the oracle, and B200Conv3d / B200MaxPool3d .perform() driven through the stand-in theano."""
    import sys
    from elektronn2_b200 import functional as F
    r = np.random.RandomState(13)
    x = r.rand(2, 6, 5, 9, 10).astype(np.float32)
    w = (r.randn(8, 6, 3, 3, 3) * 0.2).astype(np.float32)
    y = F.conv3d(x, w, compute='f32')
    assert rel(y, oo.conv3d(x, w)) <= TOL['f32']
    dy = r.randn(*y.shape).astype(np.float32)
    assert rel(F.conv3d_grad_input(dy, w, x.shape, compute='f32'), oo.conv3d_dgrad(dy, w, x.shape)) <= TOL['f32']
    assert rel(F.conv3d_grad_weights(x, dy, w.shape, compute='f32'), oo.conv3d_wgrad(dy, x, w.shape)) <= TOL['f32']
    assert rel(F.conv3d(x, w), oo.conv3d(x, w)) <= TOL['tf32']
    wu = (r.randn(4, 6, 1, 2, 2) * 0.3).astype(np.float32)
    yu = F.upconv3d(x, wu, (1, 2, 2), compute='f32')
    assert rel(yu, oo.upconv3d(x, wu, (1, 2, 2))) <= TOL['f32']
    dyu = r.randn(*yu.shape).astype(np.float32)
    assert rel(F.upconv3d_grad_input(dyu, wu, x.shape, (1, 2, 2), compute='f32'), oo.upconv3d_dgrad(dyu, wu, (1, 2, 2))) <= TOL['f32']
    assert rel(F.upconv3d_grad_weights(x, dyu, wu.shape, (1, 2, 2), compute='f32'), oo.upconv3d_wgrad(dyu, x, (1, 2, 2))) <= TOL['f32']
    xp = np.maximum(r.rand(1, 5, 4, 6, 8).astype(np.float32) - 0.3, 0)
    yp, idx = F.maxpool3d(xp, (2, 2, 2), return_argmax=True)
    assert np.array_equal(yp, oo.pooling(xp, (2, 2, 2))) and np.array_equal(idx, oo.pooling_argmax(xp, (2, 2, 2)))
    dyp = r.randn(*yp.shape).astype(np.float32)
    for tie in ('first', 'all'):
        assert np.array_equal(F.maxpool3d_grad(xp, dyp, (2, 2, 2), tie), oo.pooling_bwd(dyp, xp, (2, 2, 2), tie).astype(np.float32))
    xm = r.rand(1, 3, 7, 9, 11).astype(np.float32)
    ref, off, st = oo.fragmentpool(xm, (2, 2, 2), [[0, 0, 0]], [1, 1, 1])
    fr = F.fragmentpool(xm, (2, 2, 2))
    assert np.array_equal(fr, ref)
    assert np.array_equal(F.fragments2dense(fr, off, st), oo.fragments2dense(ref, off, st))
    # Theano Op.perform() through the stand-in package
    sys.path.insert(0, HERE)
    import theano_stub
    theano_stub.install()
    try:
        import importlib
        tops = importlib.import_module('elektronn2_b200.theano_ops')
        out = [[None]]
        tops.B200Conv3d('f32').perform(None, [x, w], out)
        assert out[0][0].shape == y.shape and rel(out[0][0], oo.conv3d(x, w)) <= TOL['f32']
        tops.B200MaxPool3d((2, 2, 2)).perform(None, [xp], out)
        assert np.array_equal(out[0][0], yp)
        tops.B200Conv3dGradW('f32').perform(None, [x, dy, np.array(w.shape)], out)
        assert rel(out[0][0], oo.conv3d_wgrad(dy, x, w.shape)) <= TOL['f32']
    finally:
        theano_stub.uninstall()


def test_computations_seam_values(h):
    """neuromancer.computations (the reference's five functions, same signatures) on the GPU: conv == np.convolve along
    every axis (the reference's own known answer, tests/test_conv.py:89-104), upconv with the reference's swapped
    weight layout, pooling, fragmentpool (+ offsets / strides bookkeeping) and fragments2dense against the oracle."""
    from elektronn2_b200.neuromancer import computations as cp
    from elektronn2_b200.config import config
    r = np.random.RandomState(17)
    old = config.compute
    config.compute = 'f32'
    try:
        sig, ker = r.rand(24).astype(np.float32), r.rand(5).astype(np.float32)
        for ax in range(3):
            xs, ks = [1, 1, 1, 1, 1], [1, 1, 1, 1, 1]
            xs[2 + ax], ks[2 + ax] = 24, 5
            y = cp.conv(sig.reshape(xs), ker.reshape(ks), axis_order='dnn', conv_dim=3)
            assert np.allclose(y.ravel(), np.convolve(sig, ker, mode='valid'), rtol=1e-5, atol=1e-6)
        x = r.rand(1, 6, 4, 5, 6).astype(np.float32)
        w = (r.randn(4, 6, 2, 2, 2) * 0.3).astype(np.float32)                # node layout (f_out, f_in, p...)
        yu = cp.upconv(x, np.swapaxes(w, 0, 1), (2, 2, 2))                    # the reference passes (f_in, f_out, p...)
        assert rel(yu, oo.upconv3d(x, w, (2, 2, 2))) <= TOL['f32']
        xp = r.rand(1, 3, 7, 9, 11).astype(np.float32)
        assert np.array_equal(cp.pooling(xp[:, :, :6, :8, :10], (2, 2, 2), [2, 3, 4]), oo.pooling(xp[:, :, :6, :8, :10], (2, 2, 2)))
        fr, off, st = cp.fragmentpool(xp, (2, 2, 2), [[0, 0, 0]], [1, 1, 1], [2, 3, 4])
        ref, roff, rst = oo.fragmentpool(xp, (2, 2, 2), [[0, 0, 0]], [1, 1, 1])
        assert np.array_equal(fr, ref) and np.array_equal(off, roff) and np.array_equal(st, rst)
        assert np.array_equal(cp.fragments2dense(fr, off, st, [2, 3, 4]), oo.fragments2dense(ref, roff, rst))
    finally:
        config.compute = old


def test_reference_test_conv_at_its_own_size(h):
    """The reference's only enabled hot-path test (tests/test_conv.py:106-163), at its own size and with its own
    tolerances: conv of x (1,5,100,300,200) with W (7,5,3,3,3), uniform[0,1) float32; then conv + (2,2,2) max-pool;
    then + softmax over the features (atol 1e-5).  The reference compares its cuDNN path with its conv3d2d path; here the
    B200 path (through the computations.py seam) stands where cuDNN stood and the oracle's restatement of the conv3d2d /
    pool_2d path stands on the other side.  F32 mode carries the reference's np.allclose defaults; TF32 mode the
    north star's 1e-3."""
    from elektronn2_b200.neuromancer import computations as cp
    from elektronn2_b200.config import config
    r = np.random.RandomState(1234)
    x_val = r.rand(1, 5, 100, 300, 200).astype(np.float32)
    W_val = r.rand(7, 5, 3, 3, 3).astype(np.float32)                             # (nof, ch, zf, xf, yf)
    r4 = oo.conv3d(x_val, W_val)
    r4p = oo.pooling(r4, (2, 2, 2))
    r4s = ol.softmax(r4p, 1)
    old = config.compute
    try:
        config.compute = 'f32'
        r3 = cp.conv(x_val, W_val, axis_order='dnn', conv_dim=3)
        assert r3.shape == (1, 7, 98, 298, 198)
        assert np.allclose(r3, r4)                                               # test_conv.py:127-128
        r3p = cp.pooling(r3, (2, 2, 2), [2, 3, 4])
        assert np.allclose(r3p, r4p)                                             # :141-142
        assert np.array_equal(r3p, oo.pooling(r3, (2, 2, 2)))                    # the pool itself: bit-exact
        r3s = cp.softmax(r3p, axis=1)
        assert np.allclose(r3s, r4s, atol=1e-5)                                  # :161-162
        assert np.allclose(r3s.sum(1), 1.0, atol=1e-6)
        config.compute = 'tf32'
        r5 = cp.conv(x_val, W_val, axis_order='dnn', conv_dim=3)
        assert rel(r5, r4) <= TOL['tf32']
    finally:
        config.compute = old


def test_reference_demo_pooling_3d_shapes(h):
    """tests/test_pooling.py:31-60 (demo_pooling_3d) only runs pooling and its gradient on (1,4,30,200,200) with pool
    (2,2,2); here the same call is also checked: values, first-maximum argmax and both gradient tie rules bit-exact."""
    from elektronn2_b200 import functional as F
    r = np.random.RandomState(5)
    x = r.rand(1, 4, 30, 200, 200).astype(np.float32)
    x[0, :, :4, :6, :6] = 0.5                                                     # windows full of ties
    y, am = F.maxpool3d(x, (2, 2, 2), return_argmax=True)
    assert np.array_equal(y, oo.pooling(x, (2, 2, 2)))
    assert np.array_equal(am, oo.pooling_argmax(x, (2, 2, 2)))
    dy = r.randn(*y.shape).astype(np.float32)
    for tie in ('first', 'all'):
        assert np.array_equal(F.maxpool3d_grad(x, dy, (2, 2, 2), tie), oo.pooling_bwd(dy, x, (2, 2, 2), tie).astype(np.float32))
