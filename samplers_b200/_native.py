"""ctypes binding of libpsx.so (include/psx.h) -- the only way the host package
reaches the GPU for the hot path.  There is deliberately no fallback: if the
library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_lib", "libpsx.so")

PSX_OK, PSX_ERR_INVALID, PSX_ERR_CUDA, PSX_ERR_UNSUPPORTED = 0, 1, 2, 3
OP_IDENTITY, OP_MASK, OP_BOX, OP_SEPBLUR, OP_CONV2D = range(5)
ABI_VERSION = 5

# name -> (restype, argtypes); must list every prototype of include/psx.h
_f32p, _i64, _vp, _f = C.c_void_p, C.c_int64, C.c_void_p, C.c_float
_opp = C.c_void_p
PROTOTYPES = {
    "psx_abi_version": (C.c_int, []),
    "psx_last_error": (C.c_char_p, []),
    "psx_reload_env": (None, []),
    "psx_kernel_launches": (C.c_longlong, []),
    "psx_op_create_identity": (C.c_int, [_i64, C.POINTER(_opp)]),
    "psx_op_create_mask": (C.c_int, [_i64, _vp, C.POINTER(_opp)]),
    "psx_op_create_box": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_opp)]),
    "psx_op_create_box_masked": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(_opp)]),
    "psx_op_create_sepblur": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int,
                                        C.POINTER(C.c_float), C.c_int, C.POINTER(_opp)]),
    "psx_op_create_conv2d": (C.c_int, [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int,
                                       C.POINTER(_opp)]),
    "psx_op_destroy": (C.c_int, [_opp]),
    "psx_op_kind": (C.c_int, [_opp]),
    "psx_op_x_numel": (_i64, [_opp]),
    "psx_op_y_numel": (_i64, [_opp]),
    "psx_op_err_parts": (C.c_int, [_opp]),
    "psx_op_workspace_bytes": (C.c_size_t, [_opp, _i64]),
    "psx_op_apply": (C.c_int, [_opp, _f32p, _f32p, _i64, _vp, C.c_size_t, _vp]),
    "psx_op_adjoint": (C.c_int, [_opp, _f32p, _f32p, _i64, _vp, C.c_size_t, _vp]),
    "psx_observe": (C.c_int, [_opp, _f32p, _f32p, _f, _f, _f32p, _i64, _vp, C.c_size_t, _vp]),
    "psx_add_noise": (C.c_int, [_f32p, _f32p, _i64, _f, _f, _vp]),
    "psx_image_to_u8": (C.c_int, [_f32p, _vp, _i64, C.c_int, C.c_int, C.c_int, _vp]),
    "psx_image_from_u8": (C.c_int, [_vp, _f32p, _i64, C.c_int, C.c_int, C.c_int, _vp]),
    "psx_gather": (C.c_int, [_f32p, _vp, _f32p, _i64, _i64, _i64, _vp]),
    "psx_scatter": (C.c_int, [_f32p, _vp, _f32p, _i64, _i64, _i64, _vp]),
    "psx_dps_pre": (C.c_int, [_opp, _f32p, _f32p, _f32p, _i64, _i64, _f, _f, _f, _f32p, _f32p, _f32p,
                              _vp, C.c_size_t, _vp]),
    "psx_dps_post": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _i64, _i64,
                               _f, _f, _f, _f, _f, _f, _f32p, _f32p, _vp]),
    "psx_dps_pre_dev": (C.c_int, [_opp, _f32p, _f32p, _f32p, _i64, _i64, _f32p, _f32p, _f32p, _f32p, _vp,
                                  C.c_size_t, _vp]),
    "psx_dps_post_dev": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _i64, _i64, _f32p, _f32p,
                                   _f32p, _vp]),
    "psx_op_fuses_mean": (C.c_int, [_opp, _i64]),
    "psx_dps_pre_mean": (C.c_int, [_opp, _f32p, _f32p, _f32p, _i64, _i64, _f, _f, _f, _f, _f, _f32p, _f, _f32p, _f32p,
                                   _f32p, _vp, C.c_size_t, _vp]),
    "psx_dps_pre_mean_dev": (C.c_int, [_opp, _f32p, _f32p, _f32p, _i64, _i64, _f32p, _f32p, _f32p, _f32p, _f32p, _vp,
                                       C.c_size_t, _vp]),
    "psx_dps_post_mean": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _i64, _i64, _f, _f, _f, _f32p,
                                    _f32p, _vp]),
    "psx_dps_post_mean_dev": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _i64, _i64, _f32p, _f32p,
                                        _f32p, _vp]),
    "psx_dps_post_philox": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _i64, _i64, _f, _f, _f, _f, _f, _f,
                                      C.c_uint64, C.c_uint64, _f32p, _f32p, _vp]),
    "psx_dps_post_philox_dev": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _f32p, C.c_int, _i64, _i64, _f32p, _vp, _f32p,
                                          _f32p, _vp]),
    "psx_dps_pre_bf16": (C.c_int, [_opp, _vp, _vp, _f32p, _i64, _i64, _f, _f, _f, _f32p, _vp, _f32p, _vp, C.c_size_t,
                                   _vp]),
    "psx_dps_post_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _f32p, C.c_int, _i64, _i64, _f, _f, _f, _f, _f, _f, _f32p,
                                    C.c_int, C.c_uint64, C.c_uint64, _vp, _vp, _f32p, _vp]),
    "psx_philox_normal": (C.c_int, [_f32p, _i64, C.c_uint64, C.c_uint64, _vp]),
    "psx_tweedie": (C.c_int, [_f32p, _f32p, _i64, _i64, _f, _f, _f32p, _f32p, _f32p, _vp]),
    "psx_bridge_update": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _i64, _f, _f, _f, _f, _f, _f, _f32p, _vp]),
    "psx_lincomb3": (C.c_int, [_f32p, _f, _f32p, _f, _f32p, _f, _f32p, _i64, _vp]),
    "psx_lincomb3_dev": (C.c_int, [_f32p, _f, _f32p, _f, _f32p, _f, _f32p, _f32p, _f32p, _i64, _vp]),
    "psx_ddim_eps_step": (C.c_int, [_f32p, _f32p, _f32p, _i64, _f, _f, _f, _f, _f, _f, _f32p, _f32p, _f32p, _vp]),
    "psx_stochastic_resample": (C.c_int, [_f32p, _f32p, _f32p, _i64, _f, _f, _f, _f, _f32p, _vp]),
    "psx_adamw_step": (C.c_int, [_f32p, _f32p, _f32p, _f32p, _i64, _f, _f, _f, _f, _f, C.c_int, _vp, C.c_int,
                                 _f32p, _i64, _f, _f, _vp]),
}


class PsxError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libpsx error {code}: {message}")
        self.code = code


_lib = None
_lock = threading.Lock()
launch_count = 0  # kernel-launching calls made through the ABI (bench.py reports it)
# CUDA kernels launched by one ABI call (psx_dps_pre depends on the operator kind)
KERNELS_PER_CALL = {"pre_identity": 1, "pre_mask": 1, "pre_box": 1, "pre_sepblur": 3, "pre_conv2d": 2,
                    "post": 1, "tweedie": 1}


def load() -> C.CDLL:
    """Loads libpsx.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"{LIB_PATH} is missing: build it with `python -m samplers_b200.build` "
                    "(samplers_b200 has no non-CUDA implementation of the sampling step)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in PROTOTYPES.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            if lib.psx_abi_version() != ABI_VERSION:
                raise RuntimeError("libpsx.so ABI version mismatch; rebuild with `python -m samplers_b200.build --force`")
            _lib = lib
    return _lib


def kernel_launches() -> int:
    """CUDA kernels launched by libpsx so far (graph captures count once)."""
    return int(load().psx_kernel_launches())


def reload_env() -> None:
    """Have libpsx read its PSX_* kernel-selection switches again (they are cached at the first launch)."""
    load().psx_reload_env()


def check(code: int) -> None:
    if code == PSX_OK:
        return
    msg = load().psx_last_error().decode(errors="replace")
    if code == PSX_ERR_INVALID:
        raise ValueError(f"libpsx: {msg}")
    if code == PSX_ERR_UNSUPPORTED:
        raise NotImplementedError(f"libpsx: {msg}")
    raise PsxError(code, msg)


def require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{name} must be a CUDA tensor: samplers_b200 runs the sampling step in sm_100a kernels only "
            "(no CPU path)")
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype} (bf16 state is not implemented yet)")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def ptr(t: torch.Tensor | None):
    return None if t is None else t.data_ptr()


def stream_ptr(device=None):
    return torch.cuda.current_stream(device).cuda_stream


class NativeOp:
    """Owns one psx_op descriptor (host struct) and the device buffers it points to."""

    def __init__(self, handle: int, keepalive=()):
        self.handle = C.c_void_p(handle)
        self._keepalive = keepalive
        lib = load()
        self.kind = lib.psx_op_kind(self.handle)
        self.n = lib.psx_op_x_numel(self.handle)
        self.n_y = lib.psx_op_y_numel(self.handle)
        self.err_parts = lib.psx_op_err_parts(self.handle)

    def fuses_mean(self, L: int) -> bool:
        """True when this operator's K1 can also write the bridge mean for L samples (psx_dps_pre_mean): the
        tensor-core blur at batches that leave SMs idle."""
        return bool(load().psx_op_fuses_mean(self.handle, L))

    def workspace_bytes(self, L: int) -> int:
        return load().psx_op_workspace_bytes(self.handle, L)

    def __del__(self):
        try:
            if self.handle and _lib is not None:
                _lib.psx_op_destroy(self.handle)
        except Exception:
            pass

    # -- constructors ------------------------------------------------------
    @staticmethod
    def identity(n: int) -> "NativeOp":
        h = _opp()
        check(load().psx_op_create_identity(n, C.byref(h)))
        return NativeOp(h.value)

    @staticmethod
    def mask(keep_u8: torch.Tensor) -> "NativeOp":
        assert keep_u8.is_cuda and keep_u8.dtype == torch.uint8 and keep_u8.is_contiguous()
        h = _opp()
        check(load().psx_op_create_mask(keep_u8.numel(), keep_u8.data_ptr(), C.byref(h)))
        return NativeOp(h.value, keepalive=(keep_u8,))

    @staticmethod
    def box(c: int, hh: int, w: int, factor: int) -> "NativeOp":
        h = _opp()
        check(load().psx_op_create_box(c, hh, w, factor, C.byref(h)))
        return NativeOp(h.value)

    @staticmethod
    def box_masked(c: int, hh: int, w: int, factor: int, keep_u8: torch.Tensor) -> "NativeOp":
        assert keep_u8.is_cuda and keep_u8.dtype == torch.uint8 and keep_u8.is_contiguous()
        assert keep_u8.numel() == c * (hh // factor) * (w // factor)
        h = _opp()
        check(load().psx_op_create_box_masked(c, hh, w, factor, keep_u8.data_ptr(), C.byref(h)))
        return NativeOp(h.value, keepalive=(keep_u8,))

    @staticmethod
    def sepblur(c: int, hh: int, w: int, taps_h, taps_v) -> "NativeOp":
        th = (C.c_float * len(taps_h))(*[float(v) for v in taps_h])
        tv = (C.c_float * len(taps_v))(*[float(v) for v in taps_v])
        h = _opp()
        check(load().psx_op_create_sepblur(c, hh, w, th, len(taps_h), tv, len(taps_v), C.byref(h)))
        return NativeOp(h.value)

    @staticmethod
    def conv2d(c: int, hh: int, w: int, kernel2d: torch.Tensor) -> "NativeOp":
        kh, kw = kernel2d.shape
        flat = [float(v) for v in kernel2d.flatten().tolist()]
        arr = (C.c_float * len(flat))(*flat)
        h = _opp()
        check(load().psx_op_create_conv2d(c, hh, w, arr, kh, kw, C.byref(h)))
        return NativeOp(h.value)

    # -- stand-alone A / A^T ------------------------------------------------
    def _run(self, fn, src: torch.Tensor, dst_numel: int) -> torch.Tensor:
        global launch_count
        require_cuda(src, "operator input")
        L = src.shape[0]
        dst = torch.empty((L, dst_numel), device=src.device, dtype=torch.float32)
        wsb = self.workspace_bytes(L)
        ws = torch.empty(wsb // 4, device=src.device, dtype=torch.float32) if wsb else None
        with torch.cuda.device(src.device):
            check(fn(self.handle, src.data_ptr(), dst.data_ptr(), L, ptr(ws), wsb, stream_ptr(src.device)))
        launch_count += 1
        return dst

    def apply(self, x_flat: torch.Tensor) -> torch.Tensor:  # (L, n) -> (L, n_y)
        return self._run(load().psx_op_apply, x_flat, self.n_y)

    def adjoint(self, y_flat: torch.Tensor) -> torch.Tensor:  # (L, n_y) -> (L, n)
        return self._run(load().psx_op_adjoint, y_flat, self.n)


def gather(src: torch.Tensor, idx: torch.Tensor, scatter: bool, n: int) -> torch.Tensor:
    """(L, n) -> (L, m) gather of kept pixels, or its transpose (scatter into zeros)."""
    global launch_count
    require_cuda(src, "gather input")
    assert idx.is_cuda and idx.dtype == torch.int64 and idx.is_contiguous()
    L, m = src.shape[0], idx.numel()
    out = torch.empty((L, n if scatter else m), device=src.device, dtype=torch.float32)
    fn = load().psx_scatter if scatter else load().psx_gather
    with torch.cuda.device(src.device):
        check(fn(src.data_ptr(), idx.data_ptr(), out.data_ptr(), L, n, m, stream_ptr(src.device)))
    launch_count += 1
    return out


def dps_pre(op: NativeOp, x_t, eps, y, obs_repeat: int, sa: float, s1: float, weight: float,
            cot, err_part, ws, x0_out=None) -> None:
    global launch_count
    L = x_t.shape[0]
    with torch.cuda.device(x_t.device):
        check(load().psx_dps_pre(op.handle, x_t.data_ptr(), eps.data_ptr(), y.data_ptr(), L, obs_repeat,
                                 sa, s1, weight, cot.data_ptr(), err_part.data_ptr(), ptr(x0_out),
                                 ptr(ws), 0 if ws is None else ws.numel() * 4, stream_ptr(x_t.device)))
    launch_count += 1


def dps_post(x_t, eps, cot, vjp, z, err_part, err_parts: int, n: int, sa: float, s1: float,
             c_ell: float, c_s: float, std: float, gamma: float, x_next, err_out=None) -> None:
    global launch_count
    L = x_t.shape[0]
    with torch.cuda.device(x_t.device):
        check(load().psx_dps_post(x_t.data_ptr(), eps.data_ptr(), cot.data_ptr(), vjp.data_ptr(), ptr(z),
                                  ptr(err_part), err_parts if err_part is not None else 0, L, n, sa, s1, c_ell,
                                  c_s, std, gamma,
                                  x_next.data_ptr(), ptr(err_out), stream_ptr(x_t.device)))
    launch_count += 1


STEP_ROW = 8  # floats per device step row (PSX_STEP_ROW): sa, s1, w / sa, c_ell, c_s, std, gamma, unused


def dps_pre_dev(op: NativeOp, x_t, eps, y, obs_repeat: int, step_row, cot, err_part, ws, x0_out=None) -> None:
    """psx_dps_pre with the step scalars read from the device row ``step_row`` (graph-replayable)."""
    global launch_count
    L = x_t.shape[0]
    with torch.cuda.device(x_t.device):
        check(load().psx_dps_pre_dev(op.handle, x_t.data_ptr(), eps.data_ptr(), y.data_ptr(), L, obs_repeat,
                                     step_row.data_ptr(), cot.data_ptr(), err_part.data_ptr(), ptr(x0_out),
                                     ptr(ws), 0 if ws is None else ws.numel() * 4, stream_ptr(x_t.device)))
    launch_count += 1


def dps_post_dev(x_t, eps, cot, vjp, z, err_part, err_parts: int, n: int, step_row, x_next, err_out=None) -> None:
    """psx_dps_post with the step scalars read from the device row ``step_row`` (graph-replayable)."""
    global launch_count
    L = x_t.shape[0]
    with torch.cuda.device(x_t.device):
        check(load().psx_dps_post_dev(x_t.data_ptr(), eps.data_ptr(), cot.data_ptr(), vjp.data_ptr(), z.data_ptr(),
                                      ptr(err_part), err_parts if err_part is not None else 0, L, n,
                                      step_row.data_ptr(), x_next.data_ptr(), ptr(err_out),
                                      stream_ptr(x_t.device)))
    launch_count += 1


def dps_pre_mean(op: NativeOp, x_t, eps, y, obs_repeat: int, sa: float, s1: float, weight: float, c_ell: float,
                 c_s: float, cot, err_part, mean, ws, step_row=None, z=None, std: float = 0.0) -> None:
    """psx_dps_pre that also writes the bridge mean c_ell x_t + c_s x0 (``op.fuses_mean`` operators only) -- plus
    std * z when the step's noise ``z`` is given; with ``step_row`` the scalars come from the device row
    (psx_dps_pre_mean_dev)."""
    global launch_count
    L = x_t.shape[0]
    wsb = 0 if ws is None else ws.numel() * 4
    with torch.cuda.device(x_t.device):
        if step_row is None:
            check(load().psx_dps_pre_mean(op.handle, x_t.data_ptr(), eps.data_ptr(), y.data_ptr(), L, obs_repeat, sa,
                                          s1, weight, c_ell, c_s, ptr(z), std, cot.data_ptr(), err_part.data_ptr(),
                                          mean.data_ptr(), ptr(ws), wsb, stream_ptr(x_t.device)))
        else:
            check(load().psx_dps_pre_mean_dev(op.handle, x_t.data_ptr(), eps.data_ptr(), y.data_ptr(), L, obs_repeat,
                                              step_row.data_ptr(), ptr(z), cot.data_ptr(), err_part.data_ptr(),
                                              mean.data_ptr(), ptr(ws), wsb, stream_ptr(x_t.device)))
    launch_count += 1


def dps_post_mean(mean, cot, vjp, z, err_part, err_parts: int, n: int, s1: float, std: float, gamma: float, x_next,
                  err_out=None, step_row=None) -> None:
    """psx_dps_post behind psx_dps_pre_mean: reads the bridge mean instead of x_t and eps."""
    global launch_count
    L = mean.shape[0]
    with torch.cuda.device(mean.device):
        if step_row is None:
            check(load().psx_dps_post_mean(mean.data_ptr(), cot.data_ptr(), vjp.data_ptr(), ptr(z), ptr(err_part),
                                           err_parts if err_part is not None else 0, L, n, s1, std, gamma,
                                           x_next.data_ptr(), ptr(err_out), stream_ptr(mean.device)))
        else:
            check(load().psx_dps_post_mean_dev(mean.data_ptr(), cot.data_ptr(), vjp.data_ptr(), ptr(z),
                                               ptr(err_part), err_parts if err_part is not None else 0, L, n,
                                               step_row.data_ptr(), x_next.data_ptr(), ptr(err_out),
                                               stream_ptr(mean.device)))
    launch_count += 1


def dps_post_philox(x_t, eps, cot, vjp, err_part, err_parts: int, n: int, sa: float, s1: float, c_ell: float,
                    c_s: float, std: float, gamma: float, seed: int, step: int, x_next, err_out=None) -> None:
    """psx_dps_post with the N(0,1) field drawn inside the kernel (Philox4x32-10 keyed by seed, counter = step)."""
    global launch_count
    L = x_t.shape[0]
    with torch.cuda.device(x_t.device):
        check(load().psx_dps_post_philox(x_t.data_ptr(), eps.data_ptr(), cot.data_ptr(), vjp.data_ptr(),
                                         ptr(err_part), err_parts if err_part is not None else 0, L, n, sa, s1,
                                         c_ell, c_s, std, gamma, seed & (2 ** 64 - 1), step, x_next.data_ptr(),
                                         ptr(err_out), stream_ptr(x_t.device)))
    launch_count += 1


def dps_post_philox_dev(x_t, eps, cot, vjp, err_part, err_parts: int, n: int, step_row, seed_step, x_next,
                        err_out=None) -> None:
    """Graph-replayable form: scalars from ``step_row``, {seed, step} from the int64 device pair ``seed_step``."""
    global launch_count
    L = x_t.shape[0]
    with torch.cuda.device(x_t.device):
        check(load().psx_dps_post_philox_dev(x_t.data_ptr(), eps.data_ptr(), cot.data_ptr(), vjp.data_ptr(),
                                             ptr(err_part), err_parts if err_part is not None else 0, L, n,
                                             step_row.data_ptr(), seed_step.data_ptr(), x_next.data_ptr(),
                                             ptr(err_out), stream_ptr(x_t.device)))
    launch_count += 1


def dps_pre_bf16(op: NativeOp, x_t, eps, y, obs_repeat: int, sa: float, s1: float, weight: float, cot, err_part,
                 step_row=None, ws=None) -> None:
    """K1 on a bf16 state (x_t, eps, cot bf16; y, err_part fp32); ``step_row`` overrides the by-value scalars."""
    global launch_count
    L = x_t.shape[0]
    with torch.cuda.device(x_t.device):
        check(load().psx_dps_pre_bf16(op.handle, x_t.data_ptr(), eps.data_ptr(), y.data_ptr(), L, obs_repeat, sa, s1,
                                      weight, ptr(step_row), cot.data_ptr(), err_part.data_ptr(), ptr(ws),
                                      0 if ws is None else ws.numel() * 4, stream_ptr(x_t.device)))
    launch_count += 1


def dps_post_bf16(x_t, eps, cot, vjp, z, err_part, err_parts: int, n: int, sa: float, s1: float, c_ell: float,
                  c_s: float, std: float, gamma: float, x_next, err_out=None, step_row=None, philox=None,
                  seed_step=None) -> None:
    """K2 on a bf16 state.  Noise: ``z`` (bf16 tensor), or ``philox=(seed, step)`` / ``seed_step`` (device pair)."""
    global launch_count
    L = x_t.shape[0]
    use_philox = philox is not None or seed_step is not None
    seed, step = philox if philox is not None else (0, 0)
    with torch.cuda.device(x_t.device):
        check(load().psx_dps_post_bf16(x_t.data_ptr(), eps.data_ptr(), cot.data_ptr(), vjp.data_ptr(), ptr(z),
                                       ptr(err_part), err_parts if err_part is not None else 0, L, n, sa, s1, c_ell,
                                       c_s, std, gamma, ptr(step_row), int(use_philox), seed & (2 ** 64 - 1), step,
                                       ptr(seed_step), x_next.data_ptr(), ptr(err_out), stream_ptr(x_t.device)))
    launch_count += 1


def philox_normal(out, seed: int, step: int) -> None:
    """out[i] = the N(0,1) field psx_dps_post_philox adds at (seed, step), i in flat element order."""
    global launch_count
    with torch.cuda.device(out.device):
        check(load().psx_philox_normal(out.data_ptr(), out.numel(), seed & (2 ** 64 - 1), step,
                                       stream_ptr(out.device)))
    launch_count += 1


def tweedie(x_t, eps, sa: float, s1: float, x0, total=None, total_sq=None) -> None:
    global launch_count
    L = x_t.shape[0]
    n = x_t.numel() // L
    with torch.cuda.device(x_t.device):
        check(load().psx_tweedie(x_t.data_ptr(), eps.data_ptr(), L, n, sa, s1, x0.data_ptr(),
                                 ptr(total), ptr(total_sq), stream_ptr(x_t.device)))
    launch_count += 1


def bridge_update(x, eps, z, grad, sa: float, s1: float, c_ell: float, c_s: float, std: float,
                  grad_scale: float, x_next) -> None:
    global launch_count
    with torch.cuda.device(x.device):
        check(load().psx_bridge_update(x.data_ptr(), eps.data_ptr(), ptr(z), ptr(grad), x.numel(), sa, s1, c_ell,
                                       c_s, std, grad_scale, x_next.data_ptr(), stream_ptr(x.device)))
    launch_count += 1


def lincomb3(a, ca: float, b, cb: float, c, cc: float, out) -> None:
    global launch_count
    with torch.cuda.device(a.device):
        check(load().psx_lincomb3(a.data_ptr(), ca, b.data_ptr(), cb, ptr(c), cc, out.data_ptr(), a.numel(),
                                  stream_ptr(a.device)))
    launch_count += 1


def lincomb3_dev(a, ca: float, b, cb: float, c, cc: float, num, den, out) -> None:
    """out = ca*a + cb*b + (cc * num / den) * c with `num`, `den` 0-dim CUDA tensors (den may be None)."""
    global launch_count
    with torch.cuda.device(a.device):
        check(load().psx_lincomb3_dev(a.data_ptr(), ca, b.data_ptr(), cb, c.data_ptr(), cc, num.data_ptr(), ptr(den),
                                      out.data_ptr(), a.numel(), stream_ptr(a.device)))
    launch_count += 1


def ddim_eps_step(x, eps, z, sc: dict, x_prev, pred_x0=None, pseudo_x0=None) -> None:
    global launch_count
    with torch.cuda.device(x.device):
        check(load().psx_ddim_eps_step(x.data_ptr(), eps.data_ptr(), ptr(z), x.numel(), sc["sqrt_a_t"], sc["sqrt_oma"],
                                       sc["oma"], sc["sqrt_a_p"], sc["dir"], sc["sigma_t"], x_prev.data_ptr(),
                                       ptr(pred_x0), ptr(pseudo_x0), stream_ptr(x.device)))
    launch_count += 1


def stochastic_resample(pseudo_x0, x_t, noise, c_p: float, c_x: float, den: float, k_n: float, out) -> None:
    global launch_count
    with torch.cuda.device(x_t.device):
        check(load().psx_stochastic_resample(pseudo_x0.data_ptr(), x_t.data_ptr(), noise.data_ptr(), x_t.numel(),
                                             c_p, c_x, den, k_n, out.data_ptr(), stream_ptr(x_t.device)))
    launch_count += 1


def adamw_step(param, grad, m, v, lr: float, step: int, *, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=1e-2,
               flags=None, flag_in=0, loss_parts=None, loss_scale=0.0, loss_threshold=0.0) -> None:
    global launch_count
    with torch.cuda.device(param.device):
        check(load().psx_adamw_step(param.data_ptr(), grad.data_ptr(), m.data_ptr(), v.data_ptr(), param.numel(), lr,
                                    beta1, beta2, eps, weight_decay, step, ptr(flags), flag_in, ptr(loss_parts),
                                    0 if loss_parts is None else loss_parts.numel(), loss_scale, loss_threshold,
                                    stream_ptr(param.device)))
    launch_count += 1


def observe(op: NativeOp, x, noise, scale: float, shift: float, y, ws=None) -> None:
    """y = A x + (scale * noise + shift); ``noise`` may be None."""
    global launch_count
    L = x.shape[0]
    with torch.cuda.device(x.device):
        check(load().psx_observe(op.handle, x.data_ptr(), ptr(noise), scale, shift, y.data_ptr(), L, ptr(ws),
                                 0 if ws is None else ws.numel() * 4, stream_ptr(x.device)))
    launch_count += 1


def add_noise(y, noise, scale: float, shift: float) -> None:
    """In place: y += scale * noise + shift (torch's two roundings)."""
    global launch_count
    with torch.cuda.device(y.device):
        check(load().psx_add_noise(y.data_ptr(), noise.data_ptr(), y.numel(), scale, shift, stream_ptr(y.device)))
    launch_count += 1


def image_to_u8(chw, hwc, images: int, c: int, h: int, w: int) -> None:
    global launch_count
    with torch.cuda.device(chw.device):
        check(load().psx_image_to_u8(chw.data_ptr(), hwc.data_ptr(), images, c, h, w, stream_ptr(chw.device)))
    launch_count += 1


def image_from_u8(hwc, chw, images: int, c: int, h: int, w: int) -> None:
    global launch_count
    with torch.cuda.device(hwc.device):
        check(load().psx_image_from_u8(hwc.data_ptr(), chw.data_ptr(), images, c, h, w, stream_ptr(hwc.device)))
    launch_count += 1
