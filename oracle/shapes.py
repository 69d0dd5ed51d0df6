"""Shape / stride / fov algebra of the hot-path nodes (TEST INFRASTRUCTURE).

Restates Conv._calc_shape (neural.py:725-764), UpConv._calc_shape (:1074-1097),
Pool._calc_shape (:1528-1559), FragmentsToDense._calc_shape (:880-892),
Crop._calc_shape (:1170-1182), TaggedShape.offsets (graphutils.py:146-147) and
the U-Net fov back-fill in Model.designate_nodes (model.py:141-152).
"""
import numpy as np


class Sh(object):
    """Spatial bookkeeping for a (b,f,z,x,y) tensor."""

    def __init__(self, b, f, spatial, strides=(1, 1, 1), fov=(1, 1, 1), mfp_offsets=None):
        self.b, self.f = b, f
        self.spatial = [int(s) for s in spatial]
        self.strides = [int(s) for s in strides]
        self.fov = [int(v) for v in fov]
        self.mfp_offsets = (np.zeros((1, 3), np.int64) if mfp_offsets is None
                            else np.atleast_2d(np.array(mfp_offsets, np.int64)))

    @property
    def shape(self):
        return [self.b, self.f] + self.spatial

    @property
    def offsets(self):
        return [v // 2 for v in self.fov]

    def copy(self):
        return Sh(self.b, self.f, self.spatial, self.strides, self.fov, self.mfp_offsets)


def mfp_bookkeeping(pool, offsets, strides):
    """Offsets/strides side of computations.fragmentpool (computations.py:665-676):
    new-offset-major, old-fragment-minor; last spatial axis fastest."""
    from itertools import product
    offsets = np.atleast_2d(np.array(offsets, np.int64))
    strides = np.array(strides, np.int64)
    if all(int(p) == 1 for p in pool):
        return offsets, strides
    new = []
    for ix in product(*[range(int(p)) for p in pool]):
        for p_ in offsets:
            new.append(p_ + np.multiply(ix, strides))
    return np.array(new, np.int64), np.multiply(pool, strides)


def conv_shape(sh, n_f, k, pool=(1, 1, 1), mfp=False):
    out = sh.copy()
    for j, (f, p) in enumerate(zip(k, pool)):
        s_in = sh.spatial[j]
        if mfp:
            if (s_in + 1 - f - p + 1) % p != 0:
                raise ValueError("MFP: axis %d len %d pool %d kernel %d" % (j, s_in, p, f))
        elif (s_in + 1 - f) % p != 0:
            raise ValueError("pool: axis %d len %d pool %d kernel %d" % (j, s_in, p, f))
        out.spatial[j] = (s_in + 1 - f) // p
        if out.spatial[j] < 1:
            raise ValueError("axis %d of length %d is too short for kernel %d / pool %d" % (j, s_in, f, p))
        out.fov[j] = sh.fov[j] + (f + p - 2) * sh.strides[j] if sh.fov[j] > 0 else -1
    if mfp:
        b = 1 if sh.b is None else sh.b
        out.mfp_offsets, st = mfp_bookkeeping(pool, sh.mfp_offsets, sh.strides)
        out.strides = [int(s) for s in st]
        out.b = b * int(np.prod(pool))
    else:
        out.strides = [int(p * s) for p, s in zip(pool, sh.strides)]
    out.f = n_f
    return out


def pool_shape(sh, pool):
    out = sh.copy()
    for j, p in enumerate(pool):
        if sh.spatial[j] % p != 0:
            raise ValueError("Cannot downsample axis %d of length %d by %d" % (j, sh.spatial[j], p))
        out.spatial[j] = sh.spatial[j] // p
        out.fov[j] = sh.fov[j] + (p - 1) * sh.strides[j] if sh.fov[j] > 0 else -1
    out.strides = [int(p * s) for p, s in zip(pool, sh.strides)]
    return out


def upconv_shape(sh, n_f, pool):
    out = sh.copy()
    out.spatial = [s * p for s, p in zip(sh.spatial, pool)]
    out.strides = [int(s // p) for s, p in zip(sh.strides, pool)]
    out.fov = [-1, -1, -1]
    out.f = n_f
    return out


def crop_shape(sh, c):
    out = sh.copy()
    out.spatial = [s - 2 * o for s, o in zip(sh.spatial, c)]
    return out


def frag2dense_shape(sh):
    out = sh.copy()
    out.spatial = [s * st for s, st in zip(sh.spatial, sh.strides)]
    out.b = 1
    out.strides = [1, 1, 1]
    out.mfp_offsets = np.zeros((1, 3), np.int64)
    return out


def unet_fov_backfill(in_spatial, pred_sh):
    """model.py:141-152."""
    out = np.array(pred_sh.strides) * (np.array(pred_sh.spatial) - 1) + 1
    diff = np.subtract(in_spatial, out)
    if np.any(np.mod(diff, 2)):
        raise ValueError("FOV is not centered")
    return [int(d) for d in diff]


def conv_macs(sh_in, n_f, k, b=1):
    """Conv._calc_comp_cost (neural.py:767-778): prod(w_sh) * n_positions * b."""
    npos = int(np.prod([s + 1 - f for s, f in zip(sh_in.spatial, k)]))
    return int(n_f * sh_in.f * int(np.prod(k)) * npos * b)


def upconv_macs(sh_in, n_f, pool, b=1):
    """UpConv._calc_comp_cost (neural.py:1100-1111) -- note the reference counts
    positions of the *output* grid minus (k-1): (S*p + 1 - p)."""
    npos = int(np.prod([s * p + 1 - p for s, p in zip(sh_in.spatial, pool)]))
    return int(n_f * sh_in.f * int(np.prod(pool)) * npos * b)
