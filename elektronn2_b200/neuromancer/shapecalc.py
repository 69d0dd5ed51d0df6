"""Small integer helpers shared by the node classes (host side)."""
from itertools import product

import numpy as np


def mfp_bookkeeping(pool, offsets, strides):
    """Offset / stride side of max-fragment-pooling (computations.py:665-676):
    for every window offset (itertools.product order, last spatial axis fastest) and
    every existing fragment offset, ``new = old + ix * strides``; strides multiply by
    the pool factors.  Fragment order: new-offset-major, old-fragment-minor."""
    offsets = np.atleast_2d(np.array(offsets, np.int64))
    strides = np.array(strides, np.int64)
    pool = [int(p) for p in pool]
    if all(p == 1 for p in pool):
        return offsets, strides
    new = [old + np.multiply(ix, strides) for ix in product(*[range(p) for p in pool]) for old in offsets]
    return np.array(new, np.int64), np.multiply(pool, strides)


def unet_fov_backfill(in_spatial, out_spatial, strides):
    """model.py:141-152: with UpConvs the fov is unknown (-1); Model.designate_nodes
    back-fills it from the in/out size difference."""
    out = np.array(strides) * (np.array(out_spatial) - 1) + 1
    diff = np.subtract(in_spatial, out)
    if np.any(np.mod(diff, 2)):
        raise ValueError("FOV is not centered. In_sh=%s, out_sh*strides=%s, diff=%s" % (in_spatial, out, diff))
    return diff.astype(np.int64)


def _axis_valid(size, filters, pools, mfps):
    s = int(size)
    for f, p, m in zip(filters, pools, mfps):
        out = s - f + 1
        if out < 1:
            return False
        if p > 1:
            if m:
                if (out - p + 1) % p != 0:   # Conv._calc_shape MFP rule, neural.py:739-744
                    return False
            elif out % p != 0:               # neural.py:746-750
                return False
        s = out // p
        if s < 1:
            return False
    return True


def closest_valid_patch_size(filters, pools, desired, mfps):
    """Largest valid patch size <= desired per axis (smallest valid one if desired is
    too small) for a sequential Conv stack -- the job utils/cnncalculator.py does for
    modelload (model.py:703-707).  Uses the divisibility rule that actually gates graph
    construction (neural.py:739-750); cnncalculator's own MFP rule (out % p == 1,
    cnncalculator.py:31-36) coincides with it for p == 2, i.e. for every shipped config."""
    out = []
    for ax, d in enumerate(desired):
        f = [fl[ax] for fl in filters]
        p = [pl[ax] for pl in pools]
        m = [bool(mm[ax]) if hasattr(mm, '__len__') else bool(mm) for mm in mfps]
        cand = [s for s in range(int(d), 0, -1) if _axis_valid(s, f, p, m)]
        if cand:
            out.append(cand[0])
            continue
        up = [s for s in range(int(d) + 1, int(d) + 5000) if _axis_valid(s, f, p, m)]
        if not up:
            raise ValueError("no valid patch size near %s on axis %d" % (d, ax))
        out.append(up[0])
    return out
