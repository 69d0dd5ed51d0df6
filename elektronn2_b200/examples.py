"""The four BASELINE model graphs, written against this package's neuromancer face.

They define the same graphs as the reference's examples/neuro3d_lite.py:46-76,
neuro3d.py:46-80, unet3d_litelite.py:55-108 and unet3d.py:58-110 (those files also run
unedited through ``elektronn2_b200.install_as_elektronn2()``, see
tests/test_host_api.py; they are restated here because the reference checkout does not
travel to the GPU box).
"""
from . import neuromancer as nm


def _head(inp, feat, target_ref):
    probs = nm.Softmax(feat)
    target = nm.Input_like(target_ref if target_ref is not None else probs, override_f=1, name='target')
    loss_pix = nm.MultinoulliNLL(probs, target, target_is_sparse=True,
                                 name='nll_barr' if target_ref is not None else 'nll')
    loss = nm.AggregateLoss(loss_pix, name='loss')
    errors = nm.Errors(probs, target, target_is_sparse=True)
    model = nm.model_manager.getmodel()
    model.designate_nodes(input_node=inp, target_node=target, loss_node=loss, prediction_node=probs,
                          prediction_ext=[loss, errors, probs])
    return model


def _seq(in_sh, layers):
    inp = nm.Input(in_sh, 'b,f,z,x,y', name='raw')
    out = inp
    for n_f, k, pool in layers:
        out = nm.Conv(out, n_f, k, pool)
    out = nm.Conv(out, 2, (1, 1, 1), activation_func='lin')
    return _head(inp, out, None)


def neuro3d_lite(in_sh=(None, 1, 11, 155, 155)):
    return _seq(in_sh, [(20, (1, 4, 4), (1, 2, 2)), (40, (3, 3, 3), (1, 2, 2)), (150, (2, 4, 4), (2, 1, 1)),
                        (200, (1, 3, 3), None), (200, (1, 3, 3), None), (200, (1, 1, 1), None)])


def neuro3d(in_sh=(None, 1, 23, 185, 185)):
    return _seq(in_sh, [(20, (1, 6, 6), (1, 2, 2)), (30, (1, 5, 5), (1, 2, 2)), (40, (1, 5, 5), None),
                        (80, (4, 4, 4), (2, 1, 1)), (100, (3, 4, 4), None), (100, (3, 4, 4), None),
                        (150, (2, 4, 4), None), (200, (1, 4, 4), None), (200, (1, 4, 4), None),
                        (200, (1, 1, 1), None)])


def _unet_down(in_sh, ch, flat, pools):
    """Contracting path shared by the two U-Nets: three (conv, conv, pool) levels."""
    inp = nm.Input(in_sh, 'b,f,z,x,y', name='raw')
    conv0 = nm.Conv(inp, ch[0], flat)
    conv1 = nm.Conv(conv0, ch[1], flat)
    down0 = nm.Pool(conv1, pools, mode='max')
    conv2 = nm.Conv(down0, ch[2], flat)
    conv3 = nm.Conv(conv2, ch[3], flat)
    down1 = nm.Pool(conv3, pools, mode='max')
    conv4 = nm.Conv(down1, ch[4], flat)
    conv5 = nm.Conv(conv4, ch[5], flat)
    down2 = nm.Pool(conv5, pools, mode='max')
    return inp, conv1, conv3, conv5, down2


def unet3d_litelite(in_sh=(None, 1, 22, 140, 140)):
    k3 = (3, 3, 3)
    inp, conv1, conv3, conv5, down2 = _unet_down(in_sh, (20, 20, 30, 30, 35, 35), (1, 3, 3), (1, 2, 2))
    conv6 = nm.Conv(down2, 42, k3)
    down2b = nm.Pool(conv6, (1, 2, 2), mode='max')
    conv7 = nm.Conv(down2b, 42, k3)
    mrg0 = nm.UpConvMerge(conv5, conv7, 45)
    mconv0 = nm.Conv(mrg0, 42, (1, 3, 3))
    mconv1 = nm.Conv(mconv0, 42, (1, 3, 3))
    mrg1 = nm.UpConvMerge(conv3, mconv1, 42)
    mconv2 = nm.Conv(mrg1, 35, k3)
    mconv3 = nm.Conv(mconv2, 35, k3)
    mrg2 = nm.UpConvMerge(conv1, mconv3, 30)
    mconv4 = nm.Conv(mrg2, 20, k3)
    mconv5 = nm.Conv(mconv4, 20, k3)
    barr = nm.Conv(mconv5, 2, (1, 1, 1), activation_func='lin', name='barr')
    return _head(inp, barr, mconv5)


def unet3d(in_sh=(None, 1, 116, 132, 132), width=1.0):
    """``width`` scales channel counts (tests use a narrow copy; 1.0 is the shipped config)."""
    def c(v):
        return max(2, int(round(v * width)))
    k3 = (3, 3, 3)
    inp, conv1, conv3, conv5, down2 = _unet_down(in_sh, tuple(c(v) for v in (32, 64, 64, 128, 128, 256)), k3,
                                                 (2, 2, 2))
    conv6 = nm.Conv(down2, c(256), k3)
    conv7 = nm.Conv(conv6, c(512), k3)
    mrg0 = nm.UpConvMerge(conv5, conv7, c(512))
    mconv0 = nm.Conv(mrg0, c(256), k3)
    mconv1 = nm.Conv(mconv0, c(256), k3)
    mrg1 = nm.UpConvMerge(conv3, mconv1, c(256))
    mconv2 = nm.Conv(mrg1, c(128), k3)
    mconv3 = nm.Conv(mconv2, c(128), k3)
    mrg2 = nm.UpConvMerge(conv1, mconv3, c(128))
    mconv4 = nm.Conv(mrg2, c(64), k3)
    mconv5 = nm.Conv(mconv4, c(64), k3)
    barr = nm.Conv(mconv5, 2, (1, 1, 1), activation_func='lin', name='barr')
    return _head(inp, barr, mconv5)


BUILDERS = dict(neuro3d_lite=neuro3d_lite, neuro3d=neuro3d, unet3d_litelite=unet3d_litelite, unet3d=unet3d)
