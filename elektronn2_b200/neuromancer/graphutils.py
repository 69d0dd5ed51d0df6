"""Shape bookkeeping for graph nodes.

Mirrors the public surface of the reference's ``TaggedShape``
(neuromancer/graphutils.py:27-310) -- same attribute and method names, same
semantics (tags b,f,z,x,y,r,s; per-spatial-axis strides / fov / mfp_offsets;
``offsets = fov // 2``) -- re-implemented on plain Python lists.
"""
import numpy as np

floatX = 'float32'

_ALLOWED_TAGS = ('r', 'b', 'z', 'f', 'x', 'y', 's')
_SPATIAL_TAGS = ('z', 'x', 'y')


def as_floatX(x):
    return np.asarray(x, dtype=floatX)


def _parse_tags(tags):
    if tags is None:
        return None
    if isinstance(tags, str):
        tags = [t.strip() for t in tags.split(',')]
    elif not isinstance(tags, (list, tuple)):
        raise ValueError("Tags must be either list/tuple of comma-separated string, not %s" % (tags,))
    tags = list(tags)
    for t in tags:
        if t not in _ALLOWED_TAGS:
            raise ValueError("Unknown tag %s" % (t,))
    return tags


class TaggedShape(object):
    """Shape + axis tags + (strides, mfp_offsets, fov) of the spatial axes."""

    def __init__(self, shape, tags, strides=None, mfp_offsets=None, fov=None):
        self._shape = list(shape)
        self._tags = _parse_tags(tags)
        if len(self._shape) != len(self._tags):
            raise ValueError("Shape %s and tags %s must have same length" % (self._shape, self._tags))
        nsp = len(self.spatial_axes)
        self._strides = np.ones(nsp, np.int64) if strides is None else np.array(strides, np.int64)
        self._mfp_offsets = (np.zeros((1, nsp), np.int64) if mfp_offsets is None
                             else np.atleast_2d(np.array(mfp_offsets, np.int64)))
        self._fov = np.ones(nsp, np.int64) if fov is None else np.array(fov, np.int64)

    # -- representation
    def __repr__(self):
        return "[" + ", ".join("(%s,%s)" % (s, t) for s, t in zip(self._shape, self._tags)) + "]"

    @property
    def ext_repr(self):
        return "%r\nfov=%s, offsets=%s, strides=%s, spatial shape=%s" % (
            self, self.fov, self.offsets, self.strides, self.spatial_shape)

    # -- access
    def __getitem__(self, key):
        if isinstance(key, str):
            return self._shape[self.tag2index(key)]
        return self._shape[key]

    def __len__(self):
        return len(self._shape)

    def __iter__(self):
        return iter(self._shape)

    shape = property(lambda self: self._shape)
    tags = property(lambda self: self._tags)
    strides = property(lambda self: self._strides)
    mfp_offsets = property(lambda self: self._mfp_offsets)
    fov = property(lambda self: [int(v) for v in self._fov])

    @property
    def fov_all_centered(self):
        return bool(np.all(np.mod(self.fov, 2) == 1))

    @property
    def offsets(self):
        return [int(v) // 2 for v in self._fov]

    @property
    def spatial_axes(self):
        return sorted(self._tags.index(t) for t in _SPATIAL_TAGS if t in self._tags)

    @property
    def ndim(self):
        return len(self.spatial_axes)

    @property
    def spatial_shape(self):
        return [self._shape[i] for i in self.spatial_axes]

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    @property
    def stripnone(self):
        return [s for s in self._shape if s is not None]

    @property
    def stripbatch_prod(self):
        return np.prod([s for s, t in zip(self._shape, self._tags) if t != 'b'])

    @property
    def stripnone_prod(self):
        return np.prod(self.stripnone)

    def tag2index(self, target_tag):
        if target_tag not in self._tags:
            raise ValueError("Shape does not have tag %s, only tags %s" % (target_tag, self._tags))
        return self._tags.index(target_tag)

    def hastag(self, tag):
        return tag in self._tags

    # -- functional updates (each returns a new object)
    def copy(self):
        return TaggedShape(self._shape, self._tags, self._strides, self._mfp_offsets, self.fov)

    def updateshape(self, axis, new_size, mode=None):
        i = axis if isinstance(axis, (int, np.integer)) else self.tag2index(axis)
        ret = self.copy()
        cur = ret._shape[i]
        if mode is None:
            ret._shape[i] = new_size
        elif cur is not None:  # None (unspecified batch) stays None
            if mode == 'add':
                ret._shape[i] = cur + new_size
            elif mode == 'mult':
                ret._shape[i] = cur * new_size
        return ret

    def updatefov(self, axis, new_fov):
        ret = self.copy()
        ret._fov[axis] = new_fov
        return ret

    def updatestrides(self, strides):
        ret = self.copy()
        ret._strides = np.array(strides, np.int64)
        return ret

    def updatemfp_offsets(self, mfp_offsets):
        ret = self.copy()
        ret._mfp_offsets = np.atleast_2d(np.array(mfp_offsets, np.int64))
        return ret

    def addaxis(self, axis, size, tag):
        i = axis if isinstance(axis, (int, np.integer)) else self.tag2index(axis) + 1
        sh, tags = list(self._shape), list(self._tags)
        sh.insert(i, size)
        tags.insert(i, tag)
        return TaggedShape(sh, tags, self._strides, self._mfp_offsets, self._fov)

    def delaxis(self, axis):
        i = axis if isinstance(axis, (int, np.integer)) else self.tag2index(axis) + 1
        sh, tags = list(self._shape), list(self._tags)
        sh.pop(i)
        tags.pop(i)
        return TaggedShape(sh, tags, self._strides, self._mfp_offsets, self._fov)
