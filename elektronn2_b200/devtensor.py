"""Device-resident channels-last activation tensors.

The reference moves (b,f,z,x,y) numpy arrays in and out of Theano on every call
(node_basic.py:464-494).  Here activations stay in HBM as NDHWC float32 with a
channel pitch (a multiple of 4 floats so rows are 16-byte aligned for float4 and
TMA access); numpy only appears at the boundary (``from_numpy`` / ``numpy``).
"""
import numpy as np
import torch

from . import _lib


def pitch_for(c):
    """Channel pitch: multiple of 4 floats (16 B rows); single-channel tensors (raw input,
    targets) stay dense."""
    return 1 if int(c) == 1 else (int(c) + 3) // 4 * 4


class DevTensor(object):
    """A (n,z,x,y,c) view into a flat float32 torch buffer."""

    def __init__(self, n, z, x, y, c, buf=None, c_pitch=None, offset=0, device=None, dtype=torch.float32):
        c_pitch = pitch_for(c) if c_pitch is None else int(c_pitch)
        self.desc = _lib.Tensor(n, z, x, y, c, c_pitch)
        if buf is None:
            dev = torch.device('cuda', torch.cuda.current_device()) if device is None else device
            # zero-filled: pad lanes must hold finite values (they meet zero weights in the GEMMs)
            buf = torch.zeros(self.desc.floats, dtype=dtype, device=dev)
        self.buf = buf
        self.offset = int(offset)

    # -- geometry
    @property
    def shape(self):
        """Logical shape in the reference's order (b,f,z,x,y)."""
        d = self.desc
        return (d.n, d.c, d.z, d.x, d.y)

    @property
    def spatial(self):
        return (self.desc.z, self.desc.x, self.desc.y)

    def ptr(self):
        return _lib.C.c_void_p(self.buf.data_ptr() + self.buf.element_size() * self.offset)

    def channel_slice(self, c0, c):
        """View of channels [c0, c0+c) sharing this tensor's storage and pitch."""
        d = self.desc
        assert 0 <= c0 and c0 + c <= d.c
        return DevTensor(d.n, d.z, d.x, d.y, c, buf=self.buf, c_pitch=d.c_pitch, offset=self.offset + c0)

    def like(self, dtype=torch.float32):
        d = self.desc
        return DevTensor(d.n, d.z, d.x, d.y, d.c, c_pitch=d.c_pitch, device=self.buf.device, dtype=dtype)

    # -- boundary
    @staticmethod
    def from_numpy(a, handle=None, out=None):
        """(b,f,z,x,y) float32 numpy -> device NDHWC (one H2D copy + one transpose kernel)."""
        h = handle or _lib.get_handle()
        a = np.ascontiguousarray(a, dtype=np.float32)
        assert a.ndim == 5, "expected (b,f,z,x,y)"
        n, c, z, x, y = a.shape
        t = out if out is not None else DevTensor(n, z, x, y, c)
        assert t.shape == tuple(a.shape)
        src = torch.from_numpy(a).to(t.buf.device, non_blocking=False)
        if c == 1 and t.desc.c_pitch == 1:
            t.buf[t.offset:t.offset + src.numel()].copy_(src.reshape(-1))
        else:
            h.call('e2_ncdhw_to_ndhwc', _lib.C.byref(t.desc), _lib.ptr(src), t.ptr(), h.stream())
        return t

    def numpy(self, handle=None):
        """device NDHWC -> (b,f,z,x,y) float32 numpy."""
        h = handle or _lib.get_handle()
        d = self.desc
        dst = torch.empty((d.n, d.c, d.z, d.x, d.y), dtype=torch.float32, device=self.buf.device)
        h.call('e2_ndhwc_to_ncdhw', _lib.C.byref(d), self.ptr(), _lib.ptr(dst), h.stream())
        return dst.cpu().numpy()

    def int_numpy(self):
        """For int32 side tensors (argmax) stored with the same geometry."""
        d = self.desc
        v = self.buf[self.offset:self.offset + d.floats].view(d.n, d.z, d.x, d.y, d.c_pitch)[..., :d.c]
        return v.permute(0, 4, 1, 2, 3).contiguous().cpu().numpy()
