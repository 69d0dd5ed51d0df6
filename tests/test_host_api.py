"""CPU tests of the host side: the C-ABI library loads and exports every symbol the
header declares, and the neuromancer mirror builds the reference's graphs with the
same shapes / strides / fovs / parameter counts as the oracle and the goldens."""
import ctypes
import io
import contextlib
import json
import os
import re
import runpy
import sys

import numpy as np
import pytest

from oracle import nets as onets

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF_EXAMPLES = '/root/reference/examples'


def _header_symbols():
    src = open(os.path.join(ROOT, 'include', 'e2b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(e2_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from elektronn2_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(_lib.lib, s), "libe2b200.so does not export %s" % s
    # and the ctypes signature table covers the whole header
    assert sorted(_lib.SIGNATURES.keys()) == syms
    assert _lib.lib.e2_version() == 100


def test_descriptor_structs_match_header_sizes(tmp_path):
    """sizeof of every POD descriptor as gcc sees include/e2b200.h == the ctypes mirror in _lib.py."""
    import subprocess
    from elektronn2_b200 import _lib
    pairs = [('e2_tensor', _lib.Tensor), ('e2_conv_desc', _lib.ConvDesc), ('e2_upconv_desc', _lib.UpConvDesc),
             ('e2_pool_desc', _lib.PoolDesc), ('e2_mfp_desc', _lib.MfpDesc), ('e2_f2d_desc', _lib.F2DDesc),
             ('e2_crop_desc', _lib.CropDesc), ('e2_affine_desc', _lib.AffineDesc)]
    src = tmp_path / 'sz.c'
    src.write_text('#include <stdio.h>\n#include "e2b200.h"\nint main(void){%s return 0;}\n'
                   % ''.join('printf("%%zu\\n", sizeof(%s));' % n for n, _ in pairs))
    exe = tmp_path / 'sz'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(c) for _, c in pairs]
    assert ctypes.sizeof(_lib.Tensor) == 24 and ctypes.sizeof(_lib.PoolDesc) == 48 + 10 * 4


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    from elektronn2_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.Handle(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, 'elektronn2_b200')
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dp, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt, f


def _build(name):
    from elektronn2_b200 import neuromancer as nm
    from elektronn2_b200 import examples
    np.random.seed(2)
    with contextlib.redirect_stdout(io.StringIO()):
        return getattr(examples, name)()


@pytest.mark.parametrize('name', ['neuro3d_lite', 'neuro3d', 'unet3d_litelite', 'unet3d'])
def test_graph_matches_oracle(name):
    m = _build(name)
    o = onets.BUILDERS[name]()
    pred = m.prediction_node
    osh = o.nodes[-1].sh
    assert [1 if s is None else s for s in pred.shape.shape] == osh.shape
    assert [int(s) for s in pred.shape.strides] == osh.strides
    assert m.loss_node.all_params_count == sum(v.size for n in o.nodes for v in n.params.values())
    # same seed, same draw order -> identical initial weights (reference init rule)
    mine = [p.get_value() for p in m.trainable_params]
    theirs = [n.params[k] for n, k in o.param_list()]
    assert len(mine) == len(theirs)
    for a, b in zip(mine, theirs):
        assert a.shape == b.shape and np.array_equal(a, b)
    if name.startswith('neuro3d'):
        assert [int(v) for v in pred.shape.fov] == osh.fov
    else:
        # U-Net fov back-fill (model.py:141-152)
        from oracle.shapes import unet_fov_backfill
        assert [int(v) for v in pred.shape.fov] == unet_fov_backfill(m.input_node.shape.spatial_shape, osh)


@pytest.mark.skipif(not os.path.isdir(REF_EXAMPLES), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize('name', ['neuro3d_lite', 'neuro3d', 'unet3d_litelite', 'unet3d_lite', 'unet3d'])
def test_unedited_reference_model_files_build(name):
    """Drop-in check: the reference's own example files execute unedited against the
    mirror (``from elektronn2 import neuromancer as nm``)."""
    import elektronn2_b200
    elektronn2_b200.install_as_elektronn2()
    np.random.seed(2)
    ns = runpy.run_path(os.path.join(REF_EXAMPLES, name + '.py'), run_name='not_main')
    with contextlib.redirect_stdout(io.StringIO()):
        m = ns['create_model']()
    known = dict(neuro3d_lite=885132, neuro3d=2756042, unet3d_litelite=352207, unet3d=19069058)
    if name in known:
        assert m.loss_node.all_params_count == known[name]
    assert m.prediction_node.shape['f'] == 2
    if name in dict(neuro3d_lite=1, neuro3d=1):
        # Conv cost == MACs (neural.py:767-778); SURVEY 8d: 5.469 / 20.016 GMAC
        macs = sum(n.computational_cost for n in m.nodes.values() if type(n).__name__ == 'Conv')
        assert round(macs / 1e9, 3) == dict(neuro3d_lite=5.469, neuro3d=20.016)[name]


def test_taggedshape_matches_reference_golden():
    from elektronn2_b200.neuromancer import TaggedShape
    g = json.load(open(os.path.join(HERE, 'golden', 'ref_python.json')))['taggedshape']
    sh = TaggedShape([None, 1, 23, 185, 185], 'b,f,z,x,y')
    s2 = sh.updateshape('f', 20).updateshape(3, 90).updatefov(1, 7).updatestrides(np.array([1, 2, 2]))
    s3 = s2.updateshape('b', 4, mode='mult').updateshape('z', 2, mode='mult')
    assert sh.spatial_axes == g['spatial_axes'] and sh.ndim == g['ndim'] and sh.spatial_shape == g['spatial_shape']
    assert s2.shape == g['s2_shape'] and [int(v) for v in s2.fov] == g['s2_fov'] and s2.offsets == g['s2_offsets']
    assert [int(v) for v in s2.strides] == g['s2_strides'] and repr(s2) == g['s2_repr']
    assert s3.shape == g['s3_shape'] and int(s2.stripnone_prod) == g['stripnone_prod']
    assert s2.spatial_size == g['spatial_size'] and sh.tag2index('f') == g['f_index']
    assert np.asarray(sh.mfp_offsets).tolist() == g['mfp_offsets']


def test_initweights_matches_reference_golden():
    from elektronn2_b200.neuromancer import initweights
    iw = json.load(open(os.path.join(HERE, 'golden', 'ref_python.json')))['initweights_seed2']
    np.random.seed(2)
    w = initweights((20, 1, 1, 4, 4), scale='glorot', mode='normal', pool=(1, 2, 2), spatial_axes=[2, 3, 4])
    assert w.dtype == np.float32 and np.allclose(w.ravel()[:8], iw['conv_w']['head'], rtol=1e-6)
    assert np.isclose(float(w.std()), iw['conv_w']['std'], rtol=1e-6)
    assert np.allclose(initweights((20,), scale=1.0 / 16, mode='const')[:3], iw['relu_b'])
    assert np.allclose(initweights((2,), scale=1e-6, mode='fix-uni'), iw['lin_b'])
    w = initweights((45, 42, 1, 4, 4), scale='glorot', mode='normal', pool=(1, 4, 4), spatial_axes=[2, 3, 4])
    assert np.allclose(w.ravel()[:4], iw['upconv_w']['head'], rtol=1e-6)


def test_error_behaviour_mirrors_reference():
    from elektronn2_b200 import neuromancer as nm
    inp = nm.Input((None, 1, 11, 155, 155), 'b,f,z,x,y', name='raw', print_repr=False)
    with pytest.raises(ValueError, match="Cannot pool spatial axis"):      # neural.py:746-750
        nm.Conv(inp, 4, (1, 3, 3), (1, 2, 2), print_repr=False)
    with pytest.raises(ValueError, match="dimensionality"):                # neural.py:594-601
        nm.Conv(inp, 4, (3, 3), print_repr=False)
    c = nm.Conv(inp, 4, (1, 4, 4), (1, 2, 2), print_repr=False)
    with pytest.raises(ValueError, match="linear activation"):             # loss.py:56-59
        nm.Softmax(c, print_repr=False)
    with pytest.raises(ValueError, match="Need .* fragments"):             # neural.py:873-875
        nm.FragmentsToDense(c, print_repr=False)
    with pytest.raises(ValueError, match="upconv_n_f"):                    # neural.py:1357-1360
        c2 = nm.Conv(c, 4, (1, 3, 3), (1, 2, 2), print_repr=False)
        nm.UpConvMerge(c, c2)
    assert [n for n in nm.model_manager.current.nodes] == ['raw', 'conv', 'conv1']


def test_mfp_rebuild_shapes():
    """modelload(override_mfp_to_active=True): patch snaps to (22,184,184), fragments
    (32,2,4,20,20) -> dense (1,2,8,80,80) (docs/examples.rst:201-209)."""
    from elektronn2_b200 import neuromancer as nm
    m = _build('neuro3d')
    m2 = nm.rebuild_model(m, override_mfp_to_active=True, imposed_patch_size=(23, 185, 185))
    assert m2.input_node.shape.spatial_shape == [22, 184, 184]
    f2d = [n for n in m2.nodes.values() if isinstance(n, nm.FragmentsToDense)][0]
    assert f2d.parent.shape.shape == [32, 2, 4, 20, 20]
    assert m2.prediction_node.shape.shape == [1, 2, 8, 80, 80]
    assert [int(s) for s in m2.prediction_node.shape.strides] == [1, 1, 1]
    batches = [n.shape['b'] for n in m2.nodes.values() if type(n) is nm.Conv]
    assert batches == [4, 16, 16, 32, 32, 32, 32, 32, 32, 32, 32]
    # weights carried over
    a = [p.get_value() for p in m.prediction_node.all_trainable_params.values()]
    b = [p.get_value() for p in m2.prediction_node.all_trainable_params.values()]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    # fragment offsets follow the reference's nesting: newest layer most significant
    o = onets.neuro3d((22, 184, 184), mfp=True)
    assert np.array_equal(f2d.parent.shape.mfp_offsets, o.nodes[-2].sh.mfp_offsets)


def test_tile_geometry_config4():
    """512^3 with neuro3d+MFP at patch (22,184,184): pred (2,498,408,408), 2268 tiles."""
    from elektronn2_b200 import neuromancer as nm
    from elektronn2_b200.neuromancer.dense import tile_geometry, tile_list, shard_tiles
    m2 = nm.rebuild_model(_build('neuro3d'), override_mfp_to_active=True, imposed_patch_size=(23, 185, 185))
    tile_sh, prob_sh, pred_sh, n_tiles = tile_geometry(m2.prediction_node, (512, 512, 512))
    assert list(pred_sh) == [498, 408, 408] and list(prob_sh) == [8, 80, 80] and int(np.prod(n_tiles)) == 2268
    tiles = tile_list(n_tiles)
    parts = [shard_tiles(tiles, r, 8) for r in range(8)]
    assert sum(len(p) for p in parts) == 2268 and sorted(sum(parts, [])) == sorted(tiles)


def test_zstack_planner_invariants():
    """Host-side tile planner of the z-stack conv kernel (no GPU): resource limits hold for every
    BASELINE layer shape, the K split only appears when a workspace is allowed, and layers with few
    tiles and a long reduction (unet3d mconv0: 768 -> 256 on 12x16x16) are split over channel blocks."""
    import ctypes as C
    from elektronn2_b200 import _lib
    fn = _lib.lib.e2_debug_zstack_plan
    layers = [(32, 64, 112, 128, 128), (64, 64, 54, 62, 62), (64, 128, 52, 60, 60), (128, 128, 24, 28, 28),
              (128, 256, 22, 26, 26), (768, 256, 12, 16, 16), (256, 256, 10, 14, 14), (384, 128, 18, 26, 26),
              (128, 128, 16, 24, 24), (192, 64, 30, 46, 46), (64, 64, 28, 44, 44), (64, 32, 114, 130, 130)]
    out = (C.c_int * 8)()
    for K, N, z, x, y in layers:
        for may_split in (0, 1):
            assert fn(148, K, N, z, x, y, 3, 3, 3, may_split, out) == 1, (K, N, z, x, y)
            bn, tz, ks, cb_per, wsl, nslot, units, ntn = list(out)
            assert bn % 16 == 0 and 16 <= bn <= 256 and min(tz, 3) * bn <= 256      # widest stacked MMA
            assert 2 * tz * bn <= 512                                               # double-buffered TMEM accumulators
            assert ntn * bn >= N and (ntn == 1 or bn % 32 == 0)
            assert wsl >= 2 and nslot in (tz + 2, 2 * (tz + 2))
            cbn = (K + 31) // 32
            assert 1 <= ks <= cbn and (ks - 1) * cb_per < cbn <= ks * cb_per       # no empty split
            if not may_split:
                assert ks == 1
            ntz = (z + tz - 1) // tz
            assert units == ntz * ((x + 15) // 16) * ((y + 7) // 8) * ntn * ks
    assert fn(148, 768, 256, 12, 16, 16, 3, 3, 3, 1, out) == 1 and out[2] > 1
    assert fn(148, 32, 64, 112, 128, 128, 3, 3, 3, 1, out) == 1 and out[2] == 1      # 3584 tiles: never split
    assert fn(148, 256, 512, 7, 9, 9, 3, 3, 3, 1, out) == 0                          # 9x9 planes: tap kernel
    # neuro3d's N = 100 layers (dense-prediction tile): one N tile of 112 with two output planes per tile fits since the
    # epilogue warps share one bias tile when there is a single N tile (two N tiles of 64 used 78 % of the columns)
    for K in (80, 100):
        assert fn(148, K, 100, 21, 83, 83, 3, 4, 4, 0, out) == 1
        assert (out[0], out[1], out[7]) == (112, 2, 1), list(out)


def test_zstack_pool_fusion_plans():
    """Conv + max-pool in one launch (e2_conv3d_fwd_pool): a tile must hold whole pool windows (TZ even for a z window
    of 2), the pair is not fused where the conv would split K (its epilogue then lives in the reduce kernel), and the
    three Conv -> Pool pairs of examples/unet3d.py keep the plan they have without the pool."""
    import ctypes as C
    from elektronn2_b200 import _lib
    plain, fused = _lib.lib.e2_debug_zstack_plan, _lib.lib.e2_debug_zstack_pool_plan
    a, b = (C.c_int * 8)(), (C.c_int * 8)()
    for K, N, z, x, y in [(32, 64, 112, 128, 128), (64, 128, 52, 60, 60), (128, 256, 22, 26, 26)]:
        assert plain(148, K, N, z, x, y, 3, 3, 3, 1, a) == 1 and fused(148, K, N, z, x, y, 3, 3, 3, 2, 2, 2, b) == 1
        assert list(a) == list(b) and b[1] % 2 == 0 and b[2] == 1
    # small layers: the plain plan has one output plane per tile, the fused one two
    assert plain(148, 32, 64, 6, 18, 24, 3, 3, 3, 1, a) == 1 and a[1] == 1
    assert fused(148, 32, 64, 6, 18, 24, 3, 3, 3, 2, 2, 2, b) == 1 and b[1] == 2
    assert fused(148, 32, 64, 6, 18, 24, 3, 3, 3, 1, 2, 2, b) == 1 and b[1] == 1        # in-plane window: any TZ
    assert fused(148, 768, 256, 12, 16, 16, 3, 3, 3, 2, 2, 2, b) == 0                    # K split wins
    assert fused(148, 32, 64, 6, 18, 24, 3, 3, 3, 3, 2, 2, b) == 0                       # windows of 1 or 2 only
    assert fused(148, 32, 64, 7, 18, 24, 3, 3, 3, 2, 2, 2, b) == 0                       # extents must divide


def test_mdl_round_trip_uses_reference_module_paths(tmp_path):
    """``Model.save`` writes the reference's .mdl pickle (model.py:229-235, graphmanager.py:236-247): every global in
    the stream carries the REFERENCE's module path (so a Theano install can load it), and ``modelload`` brings back
    the same graph and weights; with ``override_mfp_to_active`` the strided net comes back as an MFP net."""
    import contextlib
    import io
    import pickletools
    from elektronn2_b200 import examples, neuromancer as nm
    np.random.seed(2)
    nm.model_manager.reset()
    with contextlib.redirect_stdout(io.StringIO()):
        m = examples.unet3d_litelite()
    fn = str(tmp_path / 'unet.mdl')
    m.save(fn)
    globs = sorted(set(arg for op, arg, _ in pickletools.genops(open(fn, 'rb').read()) if op.name == 'GLOBAL'))
    assert 'elektronn2.neuromancer.graphmanager NodeDescriptor' in globs
    assert 'elektronn2.neuromancer.neural Conv' in globs and 'numpy.core.multiarray _reconstruct' in globs
    assert not [g for g in globs if 'elektronn2_b200' in g or 'numpy._core' in g]
    with contextlib.redirect_stdout(io.StringIO()):
        m2 = nm.modelload(fn)
    assert list(m2.nodes) == list(m.nodes)
    assert [tuple(n.shape.shape) for n in m2.nodes.values()] == [tuple(n.shape.shape) for n in m.nodes.values()]
    for a, b in zip(m.trainable_params, m2.trainable_params):
        assert a.get_value().dtype == np.float32 and np.array_equal(a.get_value(), b.get_value())
    assert m2.input_node.name == m.input_node.name and m2.loss_node.name == m.loss_node.name
    # descriptor content follows graphmanager.py:49-117: parents as NodePointers, no ndarray kwargs
    d = m.serialise()['conv1'][0]
    assert isinstance(d.args[0], nm.graphmanager.NodePointer) and d.args[0].target_id == 'conv'
    assert not any(isinstance(v, np.ndarray) for v in d.kwargs.values())
    # a strided net re-loaded for dense prediction (model.py:655-689)
    nm.model_manager.reset()
    with contextlib.redirect_stdout(io.StringIO()):
        s = examples.neuro3d_lite()
    fn2 = str(tmp_path / 'n3d.mdl')
    s.save(fn2)
    with contextlib.redirect_stdout(io.StringIO()):
        s2 = nm.modelload(fn2, override_mfp_to_active=True, imposed_patch_size=(12, 156, 156))
    assert all(int(v) == 1 for v in s2.prediction_node.shape.strides)
    assert any(type(n).__name__ == 'FragmentsToDense' for n in s2.nodes.values())
    for a, b in zip(s.trainable_params, s2.trainable_params):
        assert np.array_equal(a.get_value(), b.get_value())
    with pytest.raises(NotImplementedError):
        nm.modelload(fn2, make_weights_constant=True)


def test_theano_op_wiring_against_a_stub_theano():
    """The Theano Ops of INTEGRATION.md §4 follow the reference's Op convention (malis/malisop.py:19-123):
    __props__, make_node -> Apply, grad -> sibling Ops, infer_shape.  Theano is not installable here, so the wiring
    runs against tests/theano_stub.py (perform() is exercised on the GPU in tests/test_gpu_ops.py)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import theano_stub
    theano_stub.install()
    try:
        import importlib
        ops = importlib.import_module('elektronn2_b200.theano_ops')
        x, w = theano_stub.Variable(5, 'x'), theano_stub.Variable(5, 'w')
        y = ops.B200Conv3d()(x, w)
        assert isinstance(y.owner.op, ops.B200Conv3d) and y.owner.inputs == [x, w]
        gx, gw = y.owner.op.grad([x, w], [theano_stub.Variable(5, 'dy')])
        assert isinstance(gx.owner.op, ops.B200Conv3dGradI) and isinstance(gw.owner.op, ops.B200Conv3dGradW)
        assert ops.B200Conv3d().infer_shape(None, [(1, 32, 10, 20, 20), (64, 32, 3, 3, 3)]) == [(1, 64, 8, 18, 18)]
        assert ops.B200Conv3d('tf32') == ops.B200Conv3d('tf32') and ops.B200Conv3d('f32') != ops.B200Conv3d('tf32')
        u = ops.B200UpConv3d((2, 2, 2))
        assert u.infer_shape(None, [(1, 8, 3, 4, 5), (6, 8, 2, 2, 2)]) == [(1, 6, 6, 8, 10)]
        gi, gwt = u.grad([x, w], [theano_stub.Variable(5)])
        assert isinstance(gi.owner.op, ops.B200UpConv3dGradI) and gi.owner.op.pool == (2, 2, 2)
        assert isinstance(gwt.owner.op, ops.B200UpConv3dGradW)
        p = ops.B200MaxPool3d((1, 2, 2))
        assert p.infer_shape(None, [(2, 4, 6, 8, 10)]) == [(2, 4, 6, 4, 5)]
        g, = p.grad([x], [theano_stub.Variable(5)])
        assert isinstance(g.owner.op, ops.B200MaxPool3dGrad) and g.owner.op.tie_mode == 'first'
        assert hash(ops.B200Frag2Dense([[0, 0, 0], [0, 0, 1]], (1, 1, 2))) == hash(ops.B200Frag2Dense([[0, 0, 0], [0, 0, 1]], (1, 1, 2)))
        with pytest.raises(TypeError):
            ops.B200Conv3d()(theano_stub.Variable(4), w)
    finally:
        theano_stub.uninstall()


def test_computations_seam_error_behaviour():
    """neuromancer.computations keeps the reference's signatures and exception types (computations.py:259-260, 216,
    538, 652, 681); the value checks run on the GPU (tests/test_gpu_ops.py)."""
    from elektronn2_b200.neuromancer import computations as cp
    x, w = np.zeros((1, 2, 4, 6, 6), np.float32), np.zeros((3, 2, 3, 3, 3), np.float32)
    with pytest.raises(ValueError):
        cp.conv(x, w[..., 0], conv_dim=3)                        # computations.py:315-318
    with pytest.raises(ValueError):
        cp.conv(x, w[..., 0])                                    # :322-325
    with pytest.raises(NotImplementedError):
        cp.conv(x, w, axis_order='dnn', stride=(2, 2, 2))        # :375-376
    with pytest.raises(NotImplementedError):
        cp.conv(x, w, axis_order='dnn', border_mode='full')
    with pytest.raises(ValueError):
        cp.upconv(x, np.zeros((2, 3, 2, 2, 2), np.float32), (2, 2, 2), axis_order='theano')   # :243-244
    with pytest.raises(NotImplementedError):
        cp.pooling(x, (2, 2, 2), [2, 3, 4], stride=(1, 1, 1))    # :612-613
    with pytest.raises(ValueError):
        cp.pooling(x, (3, 2, 2), [2, 3, 4])
    assert cp.pooling(x, (1, 1, 1), [2, 3, 4]) is x             # :569-570 short circuit
    out = cp.fragmentpool(x, (1, 1, 1), [[0, 0, 0]], [1, 1, 1], [2, 3, 4])
    assert out[0] is x                                           # :653-654
    with pytest.raises(ValueError):
        cp.fragmentpool(x, (2, 2, 2), [[0, 0, 0]], [1, 1, 1], [2, 3, 4])    # (4-2+1) % 2 != 0
    with pytest.raises(ValueError):
        cp.fragments2dense(np.zeros((3, 2, 2, 2, 2), np.float32), [[0, 0, 0]] * 3, (1, 2, 2), [2, 3, 4])
    with pytest.raises(NotImplementedError):
        cp.softmax(x, axis=1, force_builtin=True)                # :172-173 (not a 2-d input)
    with pytest.raises(NotImplementedError):
        cp.softmax(x, axis=2)


def test_modelload_reads_a_file_written_by_the_reference_serialiser():
    """tests/golden/ref_written_small.mdl was written by the REFERENCE's own graphmanager.py (NodeDescriptor pointer
    replacement, GraphManager.serialise) and picklesave (tests/golden/make_mdl_fixture.py) -- not by this package's
    writer.  ``modelload`` (model.py:623-729) must rebuild the graph, the designations and every weight from it."""
    from elektronn2_b200 import neuromancer as nm
    golden = os.path.join(ROOT, 'tests', 'golden')
    z = np.load(os.path.join(golden, 'ref_written_small.npz'))
    with contextlib.redirect_stdout(io.StringIO()):
        m = nm.modelload(os.path.join(golden, 'ref_written_small.mdl'))
    assert [type(n).__name__ for n in m.nodes.values()] == ['Input', 'Conv', 'Conv', 'Conv', 'Conv', 'Softmax', 'Input',
                                                            'MultinoulliNLL', 'AggregateLoss']
    assert (m.input_node.name, m.target_node.name, m.loss_node.name, m.prediction_node.name) == ('raw', 'target', 'loss', 'softmax')
    assert [n.name for n in m.prediction_ext] == ['loss', 'softmax'] and m.error_node is None
    assert m.prediction_node.shape.shape == [None, 2, 4, 10, 10]
    assert list(m.prediction_node.shape.strides) == [2, 2, 2]
    assert m.nodes['conv2'].activation_func == 'tanh' and m.nodes['conv1'].pool_shape == (2, 1, 1)
    ps = m.prediction_node.all_trainable_params
    assert len(ps) == 8
    for k, p in ps.items():
        assert np.array_equal(p.get_value(), z['p_' + k]), k
    # the stream carries the reference's module paths, none of this package's
    raw = open(os.path.join(golden, 'ref_written_small.mdl'), 'rb').read()
    assert b'elektronn2.neuromancer.graphmanager' in raw and b'elektronn2_b200' not in raw
    # params_from_model_file (model.py:897-911): the weights without building the model
    pf = nm.params_from_model_file(os.path.join(golden, 'ref_written_small.mdl'))
    assert list(pf.keys()) == ['conv', 'conv1', 'conv2', 'conv3', 'loss']
    for node_name in ['conv', 'conv1', 'conv2', 'conv3']:
        for k, v in pf[node_name].items():
            assert np.array_equal(v, m.nodes[node_name].params[k].get_value()), (node_name, k)


def test_f4_nodes_parameters_and_shapes():
    """SURVEY 8f-4 on the host: parameter registration of batch norm / prelu (neural.py:146-242), pooling modes
    (:1448-1452), the 2-D form of examples/mnist.py:35-37, and the weight-decay regions of the flat buffer."""
    import torch
    from elektronn2_b200 import neuromancer as nm
    from elektronn2_b200.neuromancer.executor import ParamStore
    np.random.seed(1)
    with contextlib.redirect_stdout(io.StringIO()):
        inp = nm.Input((None, 1, 8, 20, 20), 'b,f,z,x,y', name='raw')
        c0 = nm.Conv(inp, 8, (1, 3, 3), (1, 2, 2), batch_normalisation='train')
        c1 = nm.Conv(c0, 12, (3, 3, 3), activation_func='prelu')
        p1 = nm.Pool(c1, (2, 1, 1), mode='average')
        c2 = nm.Conv(p1, 6, (1, 3, 3), batch_normalisation='predict', gamma=np.full(6, 2, np.float32),
                     mean=np.zeros(6, np.float32), std=np.ones(6, np.float32))
    assert list(c0.params) == ['w', 'b', 'gamma', 'mean', 'std']
    assert c0.gamma.apply_train and c0.gamma.apply_reg == 3.0 and not c0.mean.apply_train and not c0.std.apply_train
    assert np.all(c0.gamma.get_value() == 1) and np.all(c0.mean.get_value() == 0) and np.all(c0.std.get_value() == 1)
    assert c1.b.shape == (12, 2) and np.all(c1.b.get_value()[:, 1] == 1) and np.allclose(c1.b.get_value()[:, 0], 1 / 27.)
    assert p1.mode == 'average_inc_pad' and p1.shape.spatial_shape == [3, 7, 7]
    assert not c2.gamma.apply_train and np.all(c2.gamma.get_value() == 2)
    assert c0.unfused_epilogue and c1.unfused_epilogue and c2.unfused_epilogue
    with pytest.raises(ValueError):
        nm.Pool(c1, (2, 1, 1), mode='median')
    with pytest.raises(ValueError):
        nm.Conv(inp, 4, (1, 3, 3), batch_normalisation='train', mean=np.zeros(4, np.float32))
    with pytest.raises(NotImplementedError):
        nm.Conv(inp, 4, (1, 3, 3), batch_normalisation='fadeout')
    with pytest.raises(NotImplementedError, match='neural.py:650'):
        nm.Conv(inp, 4, (1, 3, 3), activation_func='maxout 2')
    # flat parameter buffer: plain weights first (decay x1), then the gamma region (x3), then the undecayed rest
    named = [('%s_%s' % (n.name, k), p) for n in (c0, c1, c2) for k, p in n.params.items() if p.apply_train]
    store = ParamStore(named, torch.device('cpu'))
    assert [m for _, _, m in store.regions] == [1.0, 3.0, 0.0]
    assert store.regions[0][:2] == (0, store.n_reg) and store.regions[1][1] == 8
    assert sum(c for _, c, _ in store.regions) == store.total
    with pytest.raises(NotImplementedError, match='shared'):
        ParamStore(named + [('again', c1.w)], torch.device('cpu'))
    # 2-D: the conv stack of examples/mnist.py:35-37
    nm.model_manager.reset()
    with contextlib.redirect_stdout(io.StringIO()):
        inp2 = nm.Input((None, 1, 26, 26), 'b,f,y,x', name='raw')
        o = nm.Conv(inp2, 12, (3, 3), (2, 2), batch_normalisation='train')
        o = nm.Conv(o, 36, (3, 3), (2, 2), batch_normalisation='train')
        o = nm.Conv(o, 64, (3, 3), (1, 1), batch_normalisation='train')
    assert o.shape.shape == [None, 64, 3, 3] and o.conv_dim == 2 and o.w_sh == [64, 36, 3, 3]
    with pytest.raises(ValueError):
        nm.Conv(inp2, 4, (3, 3, 3))


def test_docs_walkthrough_prints_what_the_reference_prints(capsys):
    """docs/examples.rst:55-218, the reference's own model walkthrough, replayed through the product's host API: the node
    printouts (#Params / Comp.Cost / Out), the prediction properties, the model totals, the dict interface, the MFP
    rebuild's patch snapping (23,183,183) -> (22,182,182) and the tile count of predict_dense on a (3,58,326,326) image."""
    from elektronn2_b200 import neuromancer as nm
    from elektronn2_b200.neuromancer import dense
    nm.model_manager.reset()
    try:
        image = nm.Input((10, 3, 23, 183, 183), 'b,f,z,x,y', name='image')
        conv0 = nm.Conv(image, 32, (1, 6, 6), (1, 2, 2))
        conv1 = nm.Conv(conv0, 64, (4, 6, 6), (2, 2, 2))
        conv2 = nm.Conv(conv1, 5, (3, 3, 3), (1, 1, 1), activation_func='lin')
        class_probs = nm.Softmax(conv2)
        target = nm.Input_like(class_probs, override_f=1, name='target', dtype='int16')
        voxel_loss = nm.MultinoulliNLL(class_probs, target, target_is_sparse=True)
        scalar_loss = nm.AggregateLoss(voxel_loss, name='loss')
        errors = nm.Errors(class_probs, target, target_is_sparse=True)
        model = nm.model_manager.getmodel()
        model.designate_nodes(input_node=image, target_node=target, loss_node=scalar_loss, prediction_node=class_probs,
                              prediction_ext=[scalar_loss, errors, class_probs])
        out = capsys.readouterr().out
        for line in [                                                       # docs/examples.rst:100-140, verbatim
            "<Input-Node> 'image'", "Out:[(10,b), (3,f), (23,z), (183,x), (183,y)]",
            "#Params=3,488 Comp.Cost=25.2 Giga Ops, Out:[(10,b), (32,f), (23,z), (89,x), (89,y)]",
            "n_f=32, 3d conv, kernel=(1, 6, 6), pool=(1, 2, 2), act='relu',",
            "#Params=294,976 Comp.Cost=416.2 Giga Ops, Out:[(10,b), (64,f), (10,z), (42,x), (42,y)]",
            "n_f=64, 3d conv, kernel=(4, 6, 6), pool=(2, 2, 2), act='relu',",
            "#Params=8,645 Comp.Cost=1.1 Giga Ops, Out:[(10,b), (5,f), (8,z), (40,x), (40,y)]",
            "n_f=5, 3d conv, kernel=(3, 3, 3), pool=(1, 1, 1), act='lin',",
            "<Softmax-Node> 'softmax'", "Comp.Cost=640.0 kilo Ops, Out:[(10,b), (5,f), (8,z), (40,x), (40,y)]",
            "<MultinoulliNLL-Node> 'nll'", "Comp.Cost=640.0 kilo Ops, Out:[(10,b), (1,f), (8,z), (40,x), (40,y)]",
            "Order of sources=['image', 'target'],",
            "<AggregateLoss-Node> 'loss'", "Comp.Cost=128.0 kilo Ops, Out:[(1,f)]", "<_Errors-Node> 'errors'",
            "Prediction properties:", "[(10,b), (5,f), (8,z), (40,x), (40,y)]",
            "fov=[9, 27, 27], offsets=[4, 13, 13], strides=[2 4 4], spatial shape=[8, 40, 40]",
            "Total Computational Cost of Model: 442.5 Giga Ops", "Total number of trainable parameters: 307,109.",
            "Computational Cost per pixel: 34.6 Mega Ops",
        ]:
            assert line in out, line
        # docs/examples.rst:176-183: the dict interface
        assert repr(model) == "['image', 'conv', 'conv1', 'conv2', 'softmax', 'target', 'nll', 'loss', 'cls for errors', 'errors']"
        assert list(conv1.all_parents.keys()) == ['image', 'conv', 'conv1']              # node_basic.py:628-648
        assert list(conv2.all_children.keys()) == ['softmax', 'nll', 'loss', 'cls for errors', 'errors']
        assert model['nll'] == voxel_loss
        assert conv2.shape.ext_repr == ('[(10,b), (5,f), (8,z), (40,x), (40,y)]\nfov=[9, 27, 27], offsets=[4, 13, 13], '
                                        'strides=[2 4 4], spatial shape=[8, 40, 40]')
        # docs/examples.rst:197-209: MFP needs a different patch size, the closest possible one is selected
        mp = nm.rebuild_model(model, imposed_batch_size=1, override_mfp_to_active=True)
        assert mp.input_node.shape.shape == [1, 3, 22, 182, 182]
        # docs/examples.rst:214-218: "Predicting img (3, 58, 326, 326) in 16 Blocks: (4, 2, 2)" -- the shape in that message
        # is the image after pad_raw added the offsets (4,13,13) on both sides (node_basic.py:930-937, 958-959)
        _, prob_sh, pred_sh, n_tiles = dense.tile_geometry(mp.prediction_node, (58, 326, 326))
        assert list(pred_sh) == [50, 300, 300] and list(prob_sh) == [14, 156, 156]
        assert list(n_tiles) == [4, 2, 2] and int(np.prod(n_tiles)) == 16
    finally:
        nm.model_manager.reset()


def test_model_training_state_properties():
    """Model.lr / mom / wd / mixing / time_per_step / loss_smooth / ... (model.py:257-546): what the reference's trainer and
    interactive shell read and set between trainingstep calls."""
    from elektronn2_b200 import neuromancer as nm
    from elektronn2_b200.neuromancer import optimiser
    nm.model_manager.reset()
    old = (optimiser.Optimiser.global_lr.get_value(), optimiser.Optimiser.global_mom.get_value(),
           optimiser.Optimiser.global_weight_decay.get_value())
    try:
        inp = nm.Input((1, 1, 8, 20, 20), 'b,f,z,x,y', name='raw', print_repr=False)
        c = nm.Conv(inp, 4, (1, 3, 3), (1, 1, 1), print_repr=False)
        out = nm.Conv(c, 2, (1, 1, 1), (1, 1, 1), activation_func='lin', print_repr=False)
        probs = nm.Softmax(out, print_repr=False)
        target = nm.Input_like(probs, override_f=1, name='target', print_repr=False)
        nll = nm.MultinoulliNLL(probs, target, target_is_sparse=True, print_repr=False)
        loss = nm.AggregateLoss(nll, name='loss', print_repr=False)
        model = nm.model_manager.getmodel()
        model.designate_nodes(input_node=inp, target_node=target, loss_node=loss, prediction_node=probs)
        model.lr, model.mom, model.wd = 1e-3, 0.8, 5e-4
        assert np.isclose(model.lr, 1e-3) and np.isclose(model.mom, 0.8) and np.isclose(model.wd, 5e-4)
        assert np.isclose(model.optimisers['Adam'].global_lr.get_value(), 1e-3)        # class-level, optimiser.py:19-55
        model.set_opt_meta_params('Adam', dict(lr=2e-3))
        assert np.isclose(model.lr, 2e-3)
        assert np.allclose(model.mixing, [1.0])
        model._train_plans[1] = object()
        model.mixing = [0.5]
        assert np.allclose(loss.mixing_weights.get_value(), [0.5]) and not model._train_plans    # re-planned with the new weight
        assert model.dropout_rates.size == 0 and model.gradnet_rates == []
        assert model.batch_normalisation_active is False
        assert model.debug_output_names is None and model.prediction_feature_names is None
        probs.feature_names = ['bg', 'fg']
        assert model.prediction_feature_names == ('bg', 'fg')
        with pytest.raises(ValueError):
            probs.feature_names = ['one']
        assert [list(s.shape) for s in model.loss_input_shapes] == [[1, 1, 8, 20, 20], [1, 1, 8, 18, 18]]
        for i in range(120):                                                            # bounded smoothing window
            model._last_exec_times.append(1.0 if i >= 70 else 100.0)
            model._last_losses.append(float(i))
        assert len(model._last_exec_times) == 50 and np.isclose(model.time_per_step, 1.0 + 1e-6)
        assert np.isclose(model.loss_smooth, np.mean(np.arange(70, 120)))
    finally:
        optimiser.Optimiser.setlr(old[0]), optimiser.Optimiser.setmom(old[1]), optimiser.Optimiser.setwd(old[2])
        nm.model_manager.reset()
