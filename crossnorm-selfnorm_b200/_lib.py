"""ctypes binding of libcnsn_b200.so (the C ABI in include/cnsn_b200.h) and the tensor-level backend.

There is NO fallback here: if the shared library is missing this module raises, and every
operator raises on non-CUDA tensors.  ``set_backend_for_tests`` exists so that the CPU test-suite
can exercise the host logic (RNG order, ``.active`` protocol, autograd wiring) against a stand-in;
the package itself never installs one.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_size_t, c_ulonglong, c_void_p

import torch

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "libcnsn_b200.so")
EXT_PATH = os.path.join(_PKG, "_cnsn_torch.so")

CNSN_F32, CNSN_BF16, CNSN_F16 = 0, 1, 2
CNSN_E_BATCH1 = -3
CNSN_E_TIMEOUT = -6
ABI_VERSION = 7

_DTYPES = {torch.float32: CNSN_F32, torch.bfloat16: CNSN_BF16, torch.float16: CNSN_F16}


class GateParams(Structure):       # struct cnsn_gate_params
    _fields_ = [("w", c_void_p), ("gamma", c_void_p), ("beta", c_void_p),
                ("run_mean", c_void_p), ("run_var", c_void_p), ("nbt", c_void_p)]


class GateGrads(Structure):        # struct cnsn_gate_grads
    _fields_ = [("dw", c_void_p), ("dgamma", c_void_p), ("dbeta", c_void_p)]


class IbnParams(Structure):        # struct cnsn_ibn_params
    _fields_ = [("in_w", c_void_p), ("in_b", c_void_p), ("bn_w", c_void_p), ("bn_b", c_void_p),
                ("run_mean", c_void_p), ("run_var", c_void_p), ("nbt", c_void_p)]


_I4 = c_int * 4
_DIMS = [c_int, c_int, c_int, c_int]

# name -> (restype, argtypes); must list every symbol include/cnsn_b200.h declares
SIGNATURES = {
    "cnsn_version": (c_int, []),
    "cnsn_error_string": (c_char_p, [c_int]),
    "cnsn_launch_count": (c_ulonglong, []),
    "cnsn_async_error": (c_int, [c_int]),
    "cnsn_tune": (c_int, [c_char_p, c_int]),
    "cnsn_instance_stats": (c_int, [c_void_p, c_int, *_DIMS, c_int, c_int, c_int, c_int, c_float,
                                    c_void_p, c_void_p, c_void_p]),
    "cnsn_instance_stats_strided": (c_int, [c_void_p, c_int, *_DIMS, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong,
                                            ctypes.c_longlong, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "cnsn_instance_stats_bwd": (c_int, [c_void_p, c_void_p, c_int, *_DIMS, c_int, c_int, c_int, c_int,
                                        c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnsn_instance_affine": (c_int, [c_void_p, c_void_p, c_int, *_DIMS, c_void_p, c_void_p, c_void_p]),
    "cnsn_instance_dot": (c_int, [c_void_p, c_void_p, c_int, *_DIMS, c_void_p, c_void_p, c_void_p]),
    "cnsn_selfnorm_save_floats": (c_size_t, [c_int, c_int, c_int]),
    "cnsn_selfnorm_workspace_floats": (c_size_t, [c_int, c_int, c_int]),
    "cnsn_selfnorm_fwd": (c_int, [c_void_p, c_void_p, c_int, *_DIMS, POINTER(GateParams), POINTER(GateParams),
                                  c_int, c_float, c_float, c_float, c_void_p, c_void_p]),
    "cnsn_selfnorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, *_DIMS,
                                  POINTER(GateParams), POINTER(GateParams), c_int, c_void_p,
                                  POINTER(GateGrads), POINTER(GateGrads), c_void_p, c_void_p]),
    "cnsn_selfnorm_block_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, *_DIMS, POINTER(GateParams),
                                        c_int, c_float, c_float, c_float, c_void_p, c_void_p]),
    "cnsn_selfnorm_block_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, *_DIMS, POINTER(GateParams),
                                        c_int, c_void_p, POINTER(GateGrads), c_void_p, c_void_p]),
    "cnsn_selfnorm_nhwc_supported": (c_int, [c_int, *_DIMS]),
    "cnsn_selfnorm_nhwc_save_floats": (c_size_t, [c_int, *_DIMS]),
    "cnsn_selfnorm_nhwc_workspace_floats": (c_size_t, [c_int, *_DIMS]),
    "cnsn_selfnorm_block_fwd_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, *_DIMS, POINTER(GateParams),
                                             c_int, c_float, c_float, c_float, c_void_p, c_void_p]),
    "cnsn_selfnorm_block_bwd_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, *_DIMS, POINTER(GateParams),
                                             c_int, c_void_p, POINTER(GateGrads), c_void_p, c_void_p]),
    "cnsn_bn_nhwc_supported": (c_int, [c_int, *_DIMS]),
    "cnsn_bn_nhwc_save_floats": (c_size_t, [c_int, *_DIMS]),
    "cnsn_bn_nhwc_workspace_floats": (c_size_t, [c_int, *_DIMS]),
    "cnsn_bn_nhwc_fwd": (c_int, [c_void_p, c_void_p, c_int, *_DIMS, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                 c_int, c_int, c_float, c_float, c_void_p, c_void_p]),
    "cnsn_bn_nhwc_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, *_DIMS, c_void_p, c_int, c_int, c_void_p,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnsn_maxpool_nhwc_out": (c_int, [c_int, c_int, c_int, c_int, c_int, POINTER(c_int), POINTER(c_int)]),
    "cnsn_maxpool_nhwc_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, *_DIMS, c_int, c_int, c_int, c_void_p]),
    "cnsn_maxpool_nhwc_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, *_DIMS, c_int, c_int, c_int, c_void_p]),
    "cnsn_bn_selfnorm_tail_supported": (c_int, [c_int, *_DIMS]),
    "cnsn_bn_selfnorm_tail_fwd_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, *_DIMS,
                                               c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p,
                                               POINTER(GateParams), c_int, c_float, c_float, c_float, c_void_p, c_void_p]),
    "cnsn_bn_selfnorm_tail_bwd_nhwc": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, *_DIMS,
                                               c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                               POINTER(GateParams), c_int, c_void_p, POINTER(GateGrads), c_void_p, c_void_p]),
    "cnsn_ibn_save_floats": (c_size_t, [c_int, c_int, c_int]),
    "cnsn_ibn_workspace_floats": (c_size_t, [c_int, c_int]),
    "cnsn_ibn_resident": (c_int, [c_int, *_DIMS, c_int, c_int]),
    "cnsn_ibn_fwd": (c_int, [c_void_p, c_void_p, c_int, *_DIMS, c_int, POINTER(IbnParams), c_int, c_int, c_float, c_float, c_float,
                             c_void_p, c_void_p]),
    "cnsn_ibn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, *_DIMS, c_int, POINTER(IbnParams), c_int, c_int, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnsn_jsd_fwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "cnsn_jsd_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "cnsn_site_save_floats": (c_size_t, [c_int, c_int]),
    "cnsn_site_workspace_floats": (c_size_t, [c_int, c_int]),
    "cnsn_site_supported": (c_int, [c_int, *_DIMS]),
    "cnsn_site_fwd": (c_int, [c_void_p, c_void_p, c_int, *_DIMS, c_void_p, POINTER(c_int), POINTER(c_int), c_float, c_float,
                              POINTER(GateParams), c_float, c_float, c_float, c_int, c_void_p, c_void_p]),
    "cnsn_site_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, *_DIMS, c_void_p, POINTER(c_int), POINTER(c_int), c_float, c_float, c_int,
                              POINTER(GateParams), c_void_p, POINTER(GateGrads), c_void_p, c_void_p]),
    "cnsn_crossnorm_save_floats": (c_size_t, [c_int, c_int]),
    "cnsn_crossnorm_workspace_floats": (c_size_t, [c_int, c_int]),
    "cnsn_crossnorm_fwd": (c_int, [c_void_p, c_void_p, c_int, *_DIMS, c_void_p, c_void_p,
                                   POINTER(c_int), POINTER(c_int), c_float, c_float, c_void_p, c_void_p]),
    "cnsn_crossnorm_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_int, *_DIMS, c_void_p, c_void_p,
                                   POINTER(c_int), POINTER(c_int), c_float, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def lib():
    """Load (once) and return the ctypes handle.  Raises if the library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.isfile(LIB_PATH):
            raise RuntimeError(
                "cnsn_b200: %s is missing. Build it with `python crossnorm-selfnorm_b200/build.py` "
                "(or __graft_entry__.build()). There is no CPU / PyTorch fallback." % LIB_PATH)
        h = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        if h.cnsn_version() != ABI_VERSION:
            raise RuntimeError("cnsn_b200: ABI mismatch (library %d, binding %d); rebuild" %
                               (h.cnsn_version(), ABI_VERSION))
        _lib = h
    return _lib


def launch_count():
    return int(lib().cnsn_launch_count())


def async_error(clear=True):
    """CNSN_E_TIMEOUT state of the library (a kernel gave up a bounded wait; see include/cnsn_b200.h).  Raises
    RuntimeError when set; ``clear`` resets it so that the process can carry on."""
    rc = lib().cnsn_async_error(int(clear))
    if rc:
        raise RuntimeError("cnsn_b200 error %d: %s" % (rc, lib().cnsn_error_string(rc).decode()))


# knob name -> {symbolic value -> int}; plain ints pass through (struct Knobs, csrc/flow_common.cuh)
_KNOB_WORDS = {"selfnorm_impl": {"auto": 0, "v1": 1, "flow": 3}, "crossnorm_impl": {"auto": 0, "v1": 1},
               "flow_mode": {"auto": 0, "res": 1, "l2": 2}, "flow_bwd": {"auto": 0, "res": 1, "dyg": 2, "l2": 3}}


def tune(**knobs):
    """MEASUREMENT / TEST HOOK: set process-wide tuning knobs of the library (``tune(reset=1)`` restores the
    defaults).  The library never reads the environment; tools that take CNSN_* variables call ``tune_from_env``."""
    for k, v in knobs.items():
        if isinstance(v, str):
            v = _KNOB_WORDS[k][v] if k in _KNOB_WORDS and v in _KNOB_WORDS[k] else int(v)
        rc = lib().cnsn_tune(k.encode(), int(v))
        if rc:
            raise KeyError("cnsn_b200: unknown tuning knob %r" % k)


class tuned:
    """``with tuned(flow_mode="res"): ...`` -- knobs set inside, defaults restored on exit (tests, A/B tools)."""

    def __init__(self, **knobs):
        self.knobs = knobs

    def __enter__(self):
        tune(**self.knobs)
        return self

    def __exit__(self, *a):
        tune(reset=1)
        return False


def tune_from_env(environ=None):
    """For the measurement tools only: CNSN_TUNE_<KNOB>=value variables -> tune().  Returns what was set."""
    environ = os.environ if environ is None else environ
    got = {k[len("CNSN_TUNE_"):].lower(): v for k, v in environ.items() if k.startswith("CNSN_TUNE_")}
    tune(**got)
    return got


_SIZES = {}


def _size(fn_name, *dims):
    """Workspace / save sizes are pure functions of the dims: one ctypes call per distinct shape."""
    key = (fn_name,) + dims
    v = _SIZES.get(key)
    if v is None:
        v = _SIZES[key] = int(getattr(lib(), fn_name)(*dims))
    return v


def _check(rc):
    if rc == 0:
        return
    msg = lib().cnsn_error_string(rc).decode()
    if rc == CNSN_E_BATCH1:          # what nn.BatchNorm1d raises inside the reference SelfNorm
        raise ValueError(msg)
    raise RuntimeError("cnsn_b200 error %d: %s" % (rc, msg))


def _dtype_code(t):
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise TypeError("cnsn_b200 supports float32 / bfloat16 / float16 tensors, got %s" % t.dtype)


def _require_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("cnsn_b200 operators run only on CUDA tensors (B200, sm_100a); got a %s "
                               "tensor. There is no CPU fallback." % t.device)


class _NoCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


_NOCTX = _NoCtx()


def _on(device):
    """Device guard that costs nothing in the common case (tensor already on the current device)."""
    if torch.cuda.current_device() == device.index:
        return _NOCTX
    return torch.cuda.device(device)


def _stream(t):
    return c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _p(t):
    return c_void_p(t.data_ptr()) if t is not None else c_void_p(0)


def _f32(t):
    """fp32 contiguous view of a parameter / buffer (copy only when needed)."""
    if t.dtype == torch.float32 and t.is_contiguous():
        return t
    return t.detach().to(torch.float32).contiguous()


class GateTensors:
    """One SelfNorm gate branch: Conv1d weight (C,1,2), BatchNorm1d affine and buffers."""
    __slots__ = ("w", "gamma", "beta", "run_mean", "run_var", "nbt")

    def __init__(self, w, gamma, beta, run_mean, run_var, nbt):
        self.w, self.gamma, self.beta = w, gamma, beta
        self.run_mean, self.run_var, self.nbt = run_mean, run_var, nbt


class CudaBackend:
    """Tensor-level calls into the C ABI.  Inputs are dense NCHW CUDA tensors."""

    name = "cuda"

    # -- statistics ---------------------------------------------------------------------
    def instance_stats(self, x, window, eps):
        """x: any strides (a sliced / transposed view, a crop, channels_last): reduced in place, no dense copy."""
        _require_cuda(x)
        N, C, H, W = x.shape
        mean = torch.empty((N, C), dtype=torch.float32, device=x.device)
        std = torch.empty_like(mean)
        with _on(x.device):
            if x.is_contiguous():
                _check(lib().cnsn_instance_stats(_p(x), _dtype_code(x), N, C, H, W, *window, eps,
                                                 _p(mean), _p(std), _stream(x)))
            else:
                _check(lib().cnsn_instance_stats_strided(_p(x), _dtype_code(x), N, C, H, W, *x.stride(), *window, eps,
                                                         _p(mean), _p(std), _stream(x)))
        return mean, std

    def instance_stats_bwd(self, x, window, mean, std, dmean, dstd):
        _require_cuda(x)
        N, C, H, W = x.shape
        dx = torch.empty_like(x)
        with _on(x.device):
            _check(lib().cnsn_instance_stats_bwd(_p(x), _p(dx), _dtype_code(x), N, C, H, W, *window,
                                                 _p(mean), _p(std), _p(dmean), _p(dstd), _stream(x)))
        return dx

    def instance_affine(self, x, scale, shift):
        _require_cuda(x)
        N, C, H, W = x.shape
        out = torch.empty_like(x)
        with _on(x.device):
            _check(lib().cnsn_instance_affine(_p(x), _p(out), _dtype_code(x), N, C, H, W,
                                              _p(scale), _p(shift), _stream(x)))
        return out

    def instance_dot(self, x, dy):
        _require_cuda(x, dy)
        N, C, H, W = x.shape
        sxy = torch.empty((N, C), dtype=torch.float32, device=x.device)
        st = torch.empty_like(sxy)
        with _on(x.device):
            _check(lib().cnsn_instance_dot(_p(x), _p(dy), _dtype_code(x), N, C, H, W, _p(sxy), _p(st), _stream(x)))
        return sxy, st

    # -- SelfNorm -------------------------------------------------------------------------
    @staticmethod
    def _gate_struct(g, keep):
        if g is None:
            return None
        w, ga, be = _f32(g.w), _f32(g.gamma), _f32(g.beta)
        rm = _f32(g.run_mean) if g.run_mean is not None else None
        rv = _f32(g.run_var) if g.run_var is not None else None
        keep.extend([w, ga, be, rm, rv])
        return GateParams(_p(w).value, _p(ga).value, _p(be).value, _p(rm).value, _p(rv).value,
                          _p(g.nbt).value), rm, rv

    def selfnorm_fwd(self, x, g, f, training, momentum, bn_eps, eps):
        _require_cuda(x)
        N, C, H, W = x.shape
        two = f is not None
        keep = []
        gs, g_rm, g_rv = self._gate_struct(g, keep)
        fs, f_rm, f_rv = self._gate_struct(f, keep) if two else (None, None, None)
        save = torch.empty(_size("cnsn_selfnorm_save_floats", N, C, int(two)), dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        with _on(x.device):
            _check(lib().cnsn_selfnorm_fwd(_p(x), _p(y), _dtype_code(x), N, C, H, W,
                                           ctypes.byref(gs), ctypes.byref(fs) if two else None,
                                           int(training), momentum, bn_eps, eps, _p(save), _stream(x)))
        if training:                     # buffers that were not fp32-contiguous get written back
            for gate, rm, rv in ((g, g_rm, g_rv), (f, f_rm, f_rv)):
                if gate is not None:
                    if rm is not gate.run_mean:
                        gate.run_mean.copy_(rm)
                    if rv is not gate.run_var:
                        gate.run_var.copy_(rv)
        return y, save

    def selfnorm_bwd(self, x, dy, g, f, training, save):
        _require_cuda(x, dy)
        N, C, H, W = x.shape
        two = f is not None
        keep = []
        gs, _, _ = self._gate_struct(g, keep)
        fs = self._gate_struct(f, keep)[0] if two else None
        dev = x.device

        def grad_block():                # one allocation per gate: dw (C,2) | dgamma (C) | dbeta (C)
            buf = torch.empty(4 * C, dtype=torch.float32, device=dev)
            return buf[:2 * C].view(C, 2), buf[2 * C:3 * C], buf[3 * C:]

        out_g = grad_block()
        out_f = grad_block() if two else None
        gg = GateGrads(*[_p(t).value for t in out_g])
        gf = GateGrads(*[_p(t).value for t in out_f]) if two else None
        ws = torch.empty(_size("cnsn_selfnorm_workspace_floats", N, C, int(two)), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x)
        with _on(dev):
            _check(lib().cnsn_selfnorm_bwd(_p(x), _p(dy), _p(dx), _dtype_code(x), N, C, H, W,
                                           ctypes.byref(gs), ctypes.byref(fs) if two else None,
                                           int(training), _p(save),
                                           ctypes.byref(gg), ctypes.byref(gf) if two else None,
                                           _p(ws), _stream(x)))
        return dx, out_g, out_f

    # -- fused block: [x + res ->] SelfNorm [-> ReLU] ------------------------------------------
    def selfnorm_block_fwd(self, x, res, relu, g, training, momentum, bn_eps, eps):
        """Returns (y, z, save): z is x + res (what backward needs), or x itself when res is None."""
        _require_cuda(x, res)
        N, C, H, W = x.shape
        keep = []
        gs, g_rm, g_rv = self._gate_struct(g, keep)
        cl = self.is_nhwc(x)                 # channels_last tensors stay in their layout (y, z come out channels_last too)
        if cl and res is not None and not self.is_nhwc(res):
            res = res.contiguous(memory_format=torch.channels_last)
        floats = (_size("cnsn_selfnorm_nhwc_save_floats", _dtype_code(x), N, C, H, W) if cl
                  else _size("cnsn_selfnorm_save_floats", N, C, 0))
        save = torch.empty(floats, dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        z = torch.empty_like(x) if res is not None else x
        fn = lib().cnsn_selfnorm_block_fwd_nhwc if cl else lib().cnsn_selfnorm_block_fwd
        with _on(x.device):
            _check(fn(_p(x), _p(res), _p(z) if res is not None else None, _p(y), int(relu),
                      _dtype_code(x), N, C, H, W, ctypes.byref(gs), int(training),
                      momentum, bn_eps, eps, _p(save), _stream(x)))
        if training:
            if g_rm is not g.run_mean:
                g.run_mean.copy_(g_rm)
            if g_rv is not g.run_var:
                g.run_var.copy_(g_rv)
        return y, z, save

    def selfnorm_block_bwd(self, z, dy, relu, g, training, save):
        _require_cuda(z, dy)
        N, C, H, W = z.shape
        keep = []
        gs, _, _ = self._gate_struct(g, keep)
        dev = z.device
        buf = torch.empty(4 * C, dtype=torch.float32, device=dev)
        out_g = (buf[:2 * C].view(C, 2), buf[2 * C:3 * C], buf[3 * C:])
        gg = GateGrads(*[_p(t).value for t in out_g])
        cl = self.is_nhwc(z)
        if cl and not self.is_nhwc(dy):
            dy = dy.contiguous(memory_format=torch.channels_last)
        floats = (_size("cnsn_selfnorm_nhwc_workspace_floats", _dtype_code(z), N, C, H, W) if cl
                  else _size("cnsn_selfnorm_workspace_floats", N, C, 0))
        ws = torch.empty(floats, dtype=torch.float32, device=dev)
        dz = torch.empty_like(z)
        fn = lib().cnsn_selfnorm_block_bwd_nhwc if cl else lib().cnsn_selfnorm_block_bwd
        with _on(dev):
            _check(fn(_p(z), _p(dy), _p(dz), int(relu), _dtype_code(z), N, C, H, W,
                      ctypes.byref(gs), int(training), _p(save), ctypes.byref(gg), _p(ws), _stream(z)))
        return dz, out_g

    # -- channels-last MaxPool2d ------------------------------------------------------------------
    @staticmethod
    def maxpool_nhwc_ok(x, k, stride, pad):
        if x.dim() != 4 or x.is_contiguous() or not x.is_contiguous(memory_format=torch.channels_last):
            return False
        return (x.data_ptr() % 16 == 0 and (x.size(1) * x.element_size()) % 16 == 0 and 1 <= k <= 15 and stride >= 1
                and 0 <= 2 * pad <= k and x.dtype in (torch.float32, torch.bfloat16, torch.float16))

    def maxpool_nhwc_fwd(self, x, k, stride, pad):
        _require_cuda(x)
        N, C, H, W = x.shape
        OH, OW = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        y = torch.empty((N, C, OH, OW), dtype=x.dtype, device=x.device, memory_format=torch.channels_last)
        code = torch.empty((N, OH, OW, C), dtype=torch.uint8, device=x.device)
        with _on(x.device):
            _check(lib().cnsn_maxpool_nhwc_fwd(_p(x), _p(y), _p(code), _dtype_code(x), N, C, H, W, k, stride, pad, _stream(x)))
        return y, code

    def maxpool_nhwc_bwd(self, dy, code, shape, k, stride, pad):
        _require_cuda(dy)
        N, C, H, W = shape
        dx = torch.empty(shape, dtype=dy.dtype, device=dy.device, memory_format=torch.channels_last)
        with _on(dy.device):
            _check(lib().cnsn_maxpool_nhwc_bwd(_p(dy), _p(code), _p(dx), _dtype_code(dy), N, C, H, W, k, stride, pad, _stream(dy)))
        return dx

    # -- channels-last BatchNorm2d [+ ReLU] -------------------------------------------------------
    @staticmethod
    def bn_nhwc_ok(x):
        """x is a dense channels_last tensor the NHWC batch-norm kernels take."""
        if x.dim() != 4 or x.is_contiguous() or not x.is_contiguous(memory_format=torch.channels_last):
            return False
        N, C, H, W = x.shape
        return x.data_ptr() % 16 == 0 and bool(_size("cnsn_bn_nhwc_supported", _dtype_code(x), N, C, H, W))

    def bn_nhwc_fwd(self, x, weight, bias, run_mean, run_var, nbt, training, relu, momentum, eps):
        _require_cuda(x)
        N, C, H, W = x.shape
        save = torch.empty(_size("cnsn_bn_nhwc_save_floats", _dtype_code(x), N, C, H, W), dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        with _on(x.device):
            _check(lib().cnsn_bn_nhwc_fwd(_p(x), _p(y), _dtype_code(x), N, C, H, W, _p(weight), _p(bias), _p(run_mean), _p(run_var),
                                          _p(nbt) if nbt is not None else None, int(training), int(relu), momentum, eps,
                                          _p(save), _stream(x)))
        return y, save

    def bn_nhwc_bwd(self, x, dy, weight, training, relu, save):
        _require_cuda(x, dy)
        N, C, H, W = x.shape
        dev = x.device
        pg = torch.empty(2 * C, dtype=torch.float32, device=dev)
        ws = torch.empty(_size("cnsn_bn_nhwc_workspace_floats", _dtype_code(x), N, C, H, W), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x)
        with _on(dev):
            _check(lib().cnsn_bn_nhwc_bwd(_p(x), _p(dy), _p(dx), _dtype_code(x), N, C, H, W, _p(weight), int(training), int(relu),
                                          _p(save), _p(pg[:C]), _p(pg[C:]), _p(ws), _stream(x)))
        return dx, pg[:C], pg[C:]

    @staticmethod
    def is_nhwc(x):
        """x is a dense channels_last tensor (and not at the same time dense NCHW) that the NHWC kernels take."""
        if x.dim() != 4 or x.is_contiguous() or not x.is_contiguous(memory_format=torch.channels_last):
            return False
        N, C, H, W = x.shape
        return x.data_ptr() % 16 == 0 and bool(_size("cnsn_selfnorm_nhwc_supported", _dtype_code(x), N, C, H, W))

    # -- IBN ----------------------------------------------------------------------------------
    @staticmethod
    def _ibn_struct(p, keep):
        vals = []
        for name in ("in_w", "in_b", "bn_w", "bn_b", "run_mean", "run_var"):
            t = p.get(name)
            t = _f32(t) if t is not None else None
            keep.append(t)
            vals.append(_p(t).value)
        vals.append(_p(p.get("nbt")).value)
        return IbnParams(*vals)

    def ibn_resident(self, x, half, training):
        """True when both directions of this shape run as the one-launch resident kernel (cached per shape)."""
        if not x.is_cuda or x.data_ptr() % 16:
            return False
        N, C, H, W = x.shape
        with _on(x.device):
            return bool(_size("cnsn_ibn_resident", _dtype_code(x), N, C, H, W, int(half), int(training)))

    def ibn_fwd(self, x, half, p, training, momentum, eps_in, eps_bn, relu=False):
        _require_cuda(x)
        N, C, H, W = x.shape
        keep = []
        ps = self._ibn_struct(p, keep)
        save = torch.empty(_size("cnsn_ibn_save_floats", N, C, half), dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        with _on(x.device):
            _check(lib().cnsn_ibn_fwd(_p(x), _p(y), _dtype_code(x), N, C, H, W, half, ctypes.byref(ps), int(training),
                                      int(relu), momentum, eps_in, eps_bn, _p(save), _stream(x)))
        if training:                     # buffers that were not fp32-contiguous were updated in a temporary: write back
            for name, tmp in (("run_mean", keep[4]), ("run_var", keep[5])):
                buf = p.get(name)
                if buf is not None and tmp is not buf:
                    buf.copy_(tmp)
        return y, save

    def ibn_bwd(self, x, dy, half, p, training, save, relu=False):
        _require_cuda(x, dy)
        N, C, H, W = x.shape
        keep = []
        ps = self._ibn_struct(p, keep)
        dev = x.device
        g = torch.empty(2 * C, dtype=torch.float32, device=dev)
        d_in_w, d_in_b, d_bn_w, d_bn_b = g[:half], g[half:2 * half], g[2 * half:C + half], g[C + half:]
        ws = torch.empty(_size("cnsn_ibn_workspace_floats", N, C), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x)
        with _on(dev):
            _check(lib().cnsn_ibn_bwd(_p(x), _p(dy), _p(dx), _dtype_code(x), N, C, H, W, half, ctypes.byref(ps), int(training),
                                      int(relu), _p(save), _p(d_in_w), _p(d_in_b), _p(d_bn_w), _p(d_bn_b), _p(ws), _stream(x)))
        return dx, (d_in_w, d_in_b, d_bn_w, d_bn_b)

    # -- JSD consistency --------------------------------------------------------------------
    def jsd_fwd(self, z0, z1, z2):
        _require_cuda(z0, z1, z2)
        B, K = z0.shape
        row = torch.empty(B, dtype=torch.float32, device=z0.device)
        loss = torch.empty((), dtype=torch.float32, device=z0.device)
        with _on(z0.device):
            _check(lib().cnsn_jsd_fwd(_p(z0), _p(z1), _p(z2), _dtype_code(z0), B, K, _p(row), _p(loss), _stream(z0)))
        return loss

    def jsd_bwd(self, z0, z1, z2, gout):
        _require_cuda(z0, z1, z2, gout)
        B, K = z0.shape
        d = [torch.empty_like(z) for z in (z0, z1, z2)]
        with _on(z0.device):
            _check(lib().cnsn_jsd_bwd(_p(z0), _p(z1), _p(z2), _dtype_code(z0), B, K, _p(gout), _p(d[0]), _p(d[1]), _p(d[2]),
                                      _stream(z0)))
        return d

    # -- CrossNorm ------------------------------------------------------------------------
    def crossnorm_fwd(self, x, perm, chan_perm, cwin, swin, lam, eps):
        _require_cuda(x, perm, chan_perm)
        N, C, H, W = x.shape
        save = torch.empty(_size("cnsn_crossnorm_save_floats", N, C), dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        with _on(x.device):
            _check(lib().cnsn_crossnorm_fwd(_p(x), _p(y), _dtype_code(x), N, C, H, W, _p(perm), _p(chan_perm),
                                            _I4(*cwin), _I4(*swin), lam, eps, _p(save), _stream(x)))
        return y, save

    def crossnorm_bwd(self, x, dy, perm, chan_perm, cwin, swin, lam, save):
        _require_cuda(x, dy)
        N, C, H, W = x.shape
        ws = torch.empty(_size("cnsn_crossnorm_workspace_floats", N, C), dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        with _on(x.device):
            _check(lib().cnsn_crossnorm_bwd(_p(x), _p(dy), _p(dx), _dtype_code(x), N, C, H, W,
                                            _p(perm), _p(chan_perm), _I4(*cwin), _I4(*swin), lam,
                                            _p(save), _p(ws), _stream(x)))
        return dx


    # -- fused site: CrossNorm -> SelfNorm ---------------------------------------------------
    def site_supported(self, x):
        """True when cnsn_site_fwd/_bwd can run this shape on x's device (cached per shape)."""
        if not x.is_cuda or x.data_ptr() % 16:          # a misaligned slice goes through the two-operator sequence
            return False
        N, C, H, W = x.shape
        with _on(x.device):
            return bool(_size("cnsn_site_supported", _dtype_code(x), N, C, H, W))

    def site_fwd(self, x, perm, cwin, swin, lam, cn_eps, g, momentum, bn_eps, sn_eps, relu=False):
        _require_cuda(x, perm)
        N, C, H, W = x.shape
        keep = []
        gs, g_rm, g_rv = self._gate_struct(g, keep)
        save = torch.empty(_size("cnsn_site_save_floats", N, C), dtype=torch.float32, device=x.device)
        y = torch.empty_like(x)
        with _on(x.device):
            _check(lib().cnsn_site_fwd(_p(x), _p(y), _dtype_code(x), N, C, H, W, _p(perm), _I4(*cwin), _I4(*swin),
                                       lam, cn_eps, ctypes.byref(gs), momentum, bn_eps, sn_eps, int(relu), _p(save),
                                       _stream(x)))
        if g_rm is not g.run_mean:
            g.run_mean.copy_(g_rm)
        if g_rv is not g.run_var:
            g.run_var.copy_(g_rv)
        return y, save

    def site_bwd(self, x, dy, perm, cwin, swin, lam, cn_eps, g, save, relu=False):
        _require_cuda(x, dy)
        N, C, H, W = x.shape
        keep = []
        gs, _, _ = self._gate_struct(g, keep)
        dev = x.device
        buf = torch.empty(4 * C, dtype=torch.float32, device=dev)
        out_g = (buf[:2 * C].view(C, 2), buf[2 * C:3 * C], buf[3 * C:])
        gg = GateGrads(*[_p(t).value for t in out_g])
        ws = torch.empty(_size("cnsn_site_workspace_floats", N, C), dtype=torch.float32, device=dev)
        dx = torch.empty_like(x)
        with _on(dev):
            _check(lib().cnsn_site_bwd(_p(x), _p(dy), _p(dx), _dtype_code(x), N, C, H, W, _p(perm), _I4(*cwin), _I4(*swin),
                                       lam, cn_eps, int(relu), ctypes.byref(gs), _p(save), ctypes.byref(gg), _p(ws),
                                       _stream(x)))
        return dx, out_g


_backend = None
_ext = None
_ext_error = None
_binding = "ext"          # "ext": C++ autograd nodes (_cnsn_torch.so) where they apply; "ctypes": always the ctypes binding


def ext():
    """The C++ host binding (_cnsn_torch.so: one autograd node per operator above the same C ABI), loaded once; None
    when it has not been built or cannot be imported -- the ctypes binding then serves every call (same kernels)."""
    global _ext, _ext_error
    if _ext is None and _ext_error is None:
        try:
            lib()                                    # libcnsn_b200.so first (the binding links it)
            import importlib.machinery
            import importlib.util
            loader = importlib.machinery.ExtensionFileLoader("_cnsn_torch", EXT_PATH)
            spec = importlib.util.spec_from_file_location("_cnsn_torch", EXT_PATH, loader=loader)
            mod = importlib.util.module_from_spec(spec)
            loader.exec_module(mod)
            if mod.abi_version() != ABI_VERSION:
                raise RuntimeError("ABI mismatch (binding %d, library %d)" % (ABI_VERSION, mod.abi_version()))
            _ext = mod
        except Exception as e:      # noqa: BLE001
            _ext_error = e
    return _ext


def set_binding(name):
    """"ext" (default) or "ctypes" -- tests and A/B measurements."""
    global _binding
    assert name in ("ext", "ctypes")
    old, _binding = _binding, name
    return old


def fast_binding():
    """The C++ binding when it is to be used: built, selected, and the tensor-level backend is the real one (the CPU
    test-suite installs a stand-in backend behind the Python autograd functions)."""
    if _binding != "ext" or (_backend is not None and not isinstance(_backend, CudaBackend)):
        return None
    return ext()


def backend():
    global _backend
    if _backend is None:
        lib()                        # fail loudly if the CUDA library is not built
        _backend = CudaBackend()
    return _backend


def set_backend_for_tests(b):
    """TEST HOOK ONLY: swap the tensor-level backend (pass None to restore the CUDA one)."""
    global _backend
    old = _backend
    _backend = b
    return old
