"""ctypes binding of libe2b200.so (the C ABI declared in include/e2b200.h).

There is no CPU fallback: importing this module without the built library raises,
and every op raises ``E2Error`` when the CUDA call fails.  torch is used only to
own device memory and streams -- all arithmetic happens inside the library.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'lib', 'libe2b200.so')

if not os.path.exists(LIB_PATH):
    raise ImportError("libe2b200.so is not built (%s missing). Run `python -c 'import __graft_entry__ as g; "
                      "g.build()'` or elektronn2_b200/csrc/build.sh; there is no CPU fallback." % LIB_PATH)

lib = C.CDLL(LIB_PATH)

OK, ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA, ERR_WORKSPACE = 0, -1, -2, -3, -4
ACT = {'lin': 0, 'linear': 0, 'relu': 1, 'tanh': 2, 'sig': 3, 'sigmoid': 3, 'logistic': 3, 'abs': 4, 'soft+': 5, 'elu': 6,
       'selu': 7, 'prelu': 8}
POOL_MODE = {'max': 0, 'average': 1, 'average_inc_pad': 1, 'average_exc_pad': 1, 'sum': 2}
COMPUTE = {'f32': 0, 'tf32': 1, 'bf16': 2}
TIE = {'first': 0, 'all': 1}


class E2Error(RuntimeError):
    def __init__(self, code, msg):
        super(E2Error, self).__init__("libe2b200 error %d: %s" % (code, msg))
        self.code = code


class E2InvalidError(E2Error, ValueError):
    """E2_ERR_INVALID -- the reference raises ValueError for the same conditions."""


class E2UnsupportedError(E2Error, NotImplementedError):
    """E2_ERR_UNSUPPORTED -- NotImplementedError in the reference."""


i32 = C.c_int32


class Tensor(C.Structure):
    _fields_ = [('n', i32), ('z', i32), ('x', i32), ('y', i32), ('c', i32), ('c_pitch', i32)]

    def __init__(self, n, z, x, y, c, c_pitch=None):
        super(Tensor, self).__init__(int(n), int(z), int(x), int(y), int(c), int(c if c_pitch is None else c_pitch))

    @property
    def positions(self):
        return self.n * self.z * self.x * self.y

    @property
    def floats(self):
        return self.positions * self.c_pitch

    def dims(self):
        return (self.n, self.z, self.x, self.y, self.c)


class ConvDesc(C.Structure):
    _fields_ = [('x', Tensor), ('y', Tensor), ('kz', i32), ('kx', i32), ('ky', i32), ('act', i32),
                ('has_bias', i32), ('compute', i32), ('accumulate', i32)]


class UpConvDesc(C.Structure):
    _fields_ = [('x', Tensor), ('y', Tensor), ('pz', i32), ('px', i32), ('py', i32), ('act', i32),
                ('has_bias', i32), ('compute', i32), ('accumulate', i32)]


class PoolDesc(C.Structure):
    _fields_ = [('x', Tensor), ('y', Tensor), ('pz', i32), ('px', i32), ('py', i32), ('act', i32),
                ('has_bias', i32), ('tie_mode', i32), ('accumulate', i32), ('round_tf32', i32), ('gate_pooled', i32),
                ('mode', i32)]


class MfpDesc(C.Structure):
    _fields_ = [('x', Tensor), ('y', Tensor), ('pz', i32), ('px', i32), ('py', i32), ('act', i32),
                ('has_bias', i32), ('round_tf32', i32)]


class AffineDesc(C.Structure):
    _fields_ = [('t', Tensor), ('act', i32), ('batch_stats', i32), ('round_tf32', i32), ('param_stride', i32)]


class F2DDesc(C.Structure):
    _fields_ = [('frag', Tensor), ('dense', Tensor), ('sz', i32), ('sx', i32), ('sy', i32)]


class Window(C.Structure):
    _fields_ = [('z0', i32), ('z1', i32), ('x0', i32), ('x1', i32), ('y0', i32), ('y1', i32)]


class CropDesc(C.Structure):
    _fields_ = [('src', Tensor), ('dst', Tensor), ('oz', i32), ('ox', i32), ('oy', i32), ('dst_c0', i32),
                ('accumulate', i32)]


vp, sz = C.c_void_p, C.c_size_t
P = C.POINTER

# name -> (restype, argtypes); mirrors include/e2b200.h one to one
SIGNATURES = {
    'e2_version': (C.c_int, []),
    'e2_create': (C.c_int, [P(vp), C.c_int]),
    'e2_destroy': (C.c_int, [vp]),
    'e2_last_error': (C.c_char_p, [vp]),
    'e2_launch_count': (C.c_int64, [vp]),
    'e2_ncdhw_to_ndhwc': (C.c_int, [vp, P(Tensor), vp, vp, vp]),
    'e2_repitch': (C.c_int, [vp, P(Tensor), vp, P(Tensor), vp, i32, vp]),
    'e2_ndhwc_to_ncdhw': (C.c_int, [vp, P(Tensor), vp, vp, vp]),
    'e2_u8_to_f32': (C.c_int, [vp, vp, vp, C.c_int64, C.c_float, vp]),
    'e2_f32_to_u8': (C.c_int, [vp, vp, vp, C.c_int64, C.c_float, vp]),
    'e2_conv3d_packed_floats': (C.c_int, [P(ConvDesc), P(sz), P(sz)]),
    'e2_conv3d_pack_weights': (C.c_int, [vp, P(ConvDesc), vp, vp, vp, vp]),
    'e2_conv3d_workspace_size': (C.c_int, [P(ConvDesc), P(sz)]),
    'e2_conv3d_fwd': (C.c_int, [vp, P(ConvDesc), vp, vp, vp, vp, vp, sz, vp]),
    'e2_conv3d_fwd_pool_supported': (C.c_int, [vp, P(ConvDesc), P(PoolDesc)]),
    'e2_conv3d_fwd_pool': (C.c_int, [vp, P(ConvDesc), P(PoolDesc), vp, vp, vp, vp, vp, P(Window), vp, vp, vp, sz, vp]),
    'e2_conv3d_dgrad': (C.c_int, [vp, P(ConvDesc), vp, vp, vp, vp, vp, sz, vp]),
    'e2_conv3d_wgrad': (C.c_int, [vp, P(ConvDesc), vp, vp, vp, vp, vp, sz, vp]),
    'e2_upconv3d_packed_floats': (C.c_int, [P(UpConvDesc), P(sz), P(sz)]),
    'e2_upconv3d_workspace_size': (C.c_int, [P(UpConvDesc), P(sz)]),
    'e2_upconv3d_pack_weights': (C.c_int, [vp, P(UpConvDesc), vp, vp, vp, vp]),
    'e2_upconv3d_fwd': (C.c_int, [vp, P(UpConvDesc), vp, vp, vp, vp, vp, sz, vp]),
    'e2_upconv3d_dgrad': (C.c_int, [vp, P(UpConvDesc), vp, vp, vp, vp, vp, sz, vp]),
    'e2_upconv3d_wgrad': (C.c_int, [vp, P(UpConvDesc), vp, vp, vp, vp, vp, sz, vp]),
    'e2_act_bwd': (C.c_int, [vp, P(Tensor), i32, vp, vp, vp, vp]),
    'e2_maxpool3d_fwd': (C.c_int, [vp, P(PoolDesc), vp, vp, vp, vp, vp]),
    'e2_maxpool3d_bwd': (C.c_int, [vp, P(PoolDesc), vp, vp, vp, vp, vp, vp]),
    'e2_mfp_fwd': (C.c_int, [vp, P(MfpDesc), vp, vp, vp, vp, vp]),
    'e2_mfp_bwd': (C.c_int, [vp, P(MfpDesc), vp, vp, vp, vp]),
    'e2_frag2dense_fwd': (C.c_int, [vp, P(F2DDesc), vp, vp, vp, vp]),
    'e2_frag2dense_bwd': (C.c_int, [vp, P(F2DDesc), vp, vp, vp, vp]),
    'e2_crop_concat_fwd': (C.c_int, [vp, P(CropDesc), vp, vp, vp]),
    'e2_crop_concat_bwd': (C.c_int, [vp, P(CropDesc), vp, vp, vp, vp]),
    'e2_softmax_nll_fwd': (C.c_int, [vp, P(Tensor), vp, vp, vp, vp, vp]),
    'e2_softmax_nll_bwd': (C.c_int, [vp, P(Tensor), vp, vp, vp, C.c_float, vp, vp]),
    'e2_adam_step': (C.c_int, [vp, vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, i32, vp]),
    'e2_sgd_step': (C.c_int, [vp, vp, vp, vp, C.c_int64, C.c_float, C.c_float, C.c_float, C.c_float, vp]),
    'e2_adam_prepare': (C.c_int, [vp, vp, vp, vp]),
    'e2_adam_step_dev': (C.c_int, [vp, vp, vp, vp, vp, C.c_int64, vp, C.c_float, vp]),
    'e2_conv3d_adam_pack_dev': (C.c_int, [vp, P(ConvDesc), vp, vp, vp, vp, vp, C.c_float, vp, vp, vp]),
    'e2_upconv3d_adam_pack_dev': (C.c_int, [vp, P(UpConvDesc), vp, vp, vp, vp, vp, C.c_float, vp, vp, vp]),
    'e2_bn_batch_stats': (C.c_int, [vp, P(Tensor), vp, vp, vp, i32, vp, vp, vp, vp, vp, vp, C.c_float, vp, vp]),
    'e2_bn_fold': (C.c_int, [vp, i32, vp, vp, i32, vp, vp, vp, vp, vp]),
    'e2_affine_act_fwd': (C.c_int, [vp, P(AffineDesc), vp, vp, vp, vp, vp, vp]),
    'e2_affine_act_bwd': (C.c_int, [vp, P(AffineDesc), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]),
    'e2_affine_scratch_bytes': (C.c_int, [i32, P(sz)]),
    'e2_maxout_fwd': (C.c_int, [vp, vp, vp, C.c_int64, i32, C.c_int64, i32, vp]),
    'e2_maxout_bwd': (C.c_int, [vp, vp, vp, vp, C.c_int64, i32, C.c_int64, i32, vp]),
    'e2_debug_zstack_plan': (C.c_int, [C.c_int] * 10 + [P(C.c_int)]),
    'e2_debug_zstack_pool_plan': (C.c_int, [C.c_int] * 12 + [P(C.c_int)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == a declared symbol is not exported
    _fn.restype = _res
    _fn.argtypes = _args


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    return C.c_void_p(t.data_ptr())


class Handle(object):
    """One per device.  ``call`` raises on any non-zero status."""

    def __init__(self, device=0):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("elektronn2_b200 needs a CUDA device (B200, sm_100a); none is visible "
                               "and there is no CPU fallback.")
        self.device = int(device)
        h = vp()
        rc = lib.e2_create(C.byref(h), self.device)
        if rc != OK:
            raise E2Error(rc, "e2_create(device=%d) failed" % self.device)
        self._h = h
        self._ws, self._ws_old = None, []     # caller-owned scratch shared by all ops of this handle's stream
        self._ws_side = {}                    # further scratch buffers for ops launched on side streams (ws_slot >= 1)
        self.ws_slot = 0
        msg = lib.e2_last_error(self._h)
        if msg:
            raise E2Error(ERR_UNSUPPORTED, msg.decode())

    def stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def reserve_workspace(self, nbytes):
        """Grow the shared scratch buffer (e2_*_workspace_size).  Old buffers stay alive: a captured
        CUDA graph may still reference them."""
        import torch
        nbytes = int(nbytes)
        if nbytes and (self._ws is None or self._ws.numel() < nbytes):
            if self._ws is not None:
                self._ws_old.append(self._ws)
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device='cuda:%d' % self.device)

    def workspace(self):
        """(void* ws, size_t ws_bytes) for the entry points that take a workspace.  ``ws_slot`` >= 1 selects another
        buffer of the same size: kernels that run concurrently on a side stream must not share scratch."""
        if self._ws is None:
            return None, 0
        if self.ws_slot >= 1:
            import torch
            cur = self._ws_side.get(self.ws_slot)
            if cur is None or cur.numel() < self._ws.numel():
                if cur is not None:
                    self._ws_old.append(cur)
                cur = torch.empty(self._ws.numel(), dtype=torch.uint8, device='cuda:%d' % self.device)
                self._ws_side[self.ws_slot] = cur
            return C.c_void_p(cur.data_ptr()), C.c_size_t(cur.numel())
        return C.c_void_p(self._ws.data_ptr()), C.c_size_t(self._ws.numel())

    def call(self, name, *args):
        rc = getattr(lib, name)(self._h, *args)
        if rc != OK:
            msg = lib.e2_last_error(self._h).decode()
            cls = {ERR_INVALID: E2InvalidError, ERR_UNSUPPORTED: E2UnsupportedError}.get(rc, E2Error)
            raise cls(rc, "%s: %s" % (name, msg))

    def query(self, name, *args):
        """For entry points whose non-negative return value is an answer, not a status."""
        rc = getattr(lib, name)(self._h, *args)
        if rc < 0:
            msg = lib.e2_last_error(self._h).decode()
            cls = {ERR_INVALID: E2InvalidError, ERR_UNSUPPORTED: E2UnsupportedError}.get(rc, E2Error)
            raise cls(rc, "%s: %s" % (name, msg))
        return rc

    @property
    def launches(self):
        return int(lib.e2_launch_count(self._h))

    def __del__(self):
        try:
            if getattr(self, '_h', None):
                lib.e2_destroy(self._h)
                self._h = None
        except Exception:
            pass


_handles = {}


def get_handle(device=None):
    import torch
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    if device not in _handles:
        _handles[device] = Handle(device)
    return _handles[device]
