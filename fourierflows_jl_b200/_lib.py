"""ctypes binding of libfourierflows_b200.so -- the Python stand-in for Julia's `ccall` layer.

Every call goes through `check()`, which turns a non-zero status into the exception the reference would raise
(`DomainError` for odd grid sizes, src/domains.jl:66,179,316; `error(...)` otherwise).  There is NO CPU fallback:
if the shared library is missing, or no CUDA device is present, the compute entry points fail loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FFB_LIB_PATH: another build of the same library (A/B timing of two builds on one box); the default is the in-tree build
LIB_PATH = os.environ.get("FFB_LIB_PATH") or os.path.join(_HERE, "lib", "libfourierflows_b200.so")

FFB_OK, FFB_EINVAL, FFB_EDOMAIN, FFB_ENOMEM, FFB_ECUDA, FFB_ENCCL, FFB_EUNSUPPORTED, FFB_ESTEPPER = 0, -1, -2, -3, -4, -5, -6, -7
FFB_F32, FFB_F64 = 0, 1
FFB_R2C, FFB_C2C = 0, 1
FFB_COEF_SCALAR, FFB_COEF_REAL, FFB_COEF_COMPLEX = 0, 1, 2
FFB_PLAN_DEFAULT, FFB_PLAN_FORCE_GENERIC = 0, 1
FFB_CALCN_CALLBACK, FFB_CALCN_ZERO, FFB_CALCN_DIFFUSION, FFB_CALCN_VORTICITY2D, FFB_CALCN_BURGERS3D = 0, 1, 2, 3, 4
FFB_FORWARD_EULER, FFB_RK4, FFB_LSRK54, FFB_ETDRK4, FFB_AB3 = 0, 1, 2, 3, 4


class DomainError(ValueError):
    """Julia `DomainError` (odd grid size)."""


class FFBError(RuntimeError):
    """Julia `error(...)` raised from a failing library call."""

    def __init__(self, code, msg):
        super().__init__(f"libfourierflows_b200 error {code}: {msg}")
        self.code = code


class ffb_coef(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("kind", C.c_int), ("dtype", C.c_int), ("re", C.c_double), ("im", C.c_double)]


class ffb_desc(C.Structure):
    _fields_ = [("ndim", C.c_int), ("dims", C.c_int64 * 4), ("dtype", C.c_int), ("alias_lo", C.c_int32 * 3),
                ("alias_hi", C.c_int32 * 3)]


class ffb_fuse(C.Structure):
    _fields_ = [("cr", C.c_double), ("ci", C.c_double), ("kx", C.c_void_p), ("l", C.c_void_p), ("m", C.c_void_p), ("w", C.c_void_p),
                ("acc", C.c_void_p), ("ar", C.c_double), ("ai", C.c_double), ("akx", C.c_void_p), ("al", C.c_void_p), ("am", C.c_void_p),
                ("dealias", C.c_int), ("alias_lo", C.c_int32 * 3), ("alias_hi", C.c_int32 * 3), ("mul", C.c_void_p), ("square_input", C.c_int),
                ("galias_lo", C.c_int32 * 3), ("galias_hi", C.c_int32 * 3)]


CALCN_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_void_p)


class ffb_problem_config(C.Structure):
    _fields_ = [("ndim", C.c_int), ("n", C.c_int64 * 3), ("L", C.c_double * 3), ("dtype", C.c_int),
                ("aliased_fraction", C.c_double), ("stepper", C.c_int), ("filtered", C.c_int),
                ("filter_order", C.c_double), ("filter_innerK", C.c_double), ("filter_outerK", C.c_double),
                ("filter_tol", C.c_double), ("dt", C.c_double), ("calcN", C.c_int), ("callback", CALCN_FN),
                ("user", C.c_void_p), ("nu", C.c_double), ("scalar_zero_L", C.c_int), ("kappa", C.c_void_p),
                ("coef_dtype", C.c_int), ("fused", C.c_int), ("dist", C.c_void_p)]


# name -> argtypes; every function returns int except the two listed in _SPECIAL
_vp, _i, _i64, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
_P = C.POINTER
SIGNATURES = {
    "ffb_version": [],
    "ffb_device_count": [_P(_i)],
    "ffb_set_device": [_i],
    "ffb_set_stream": [_vp],
    "ffb_get_stream": [_P(_vp)],
    "ffb_sync": [],
    "ffb_launch_count": [_P(C.c_uint64)],
    "ffb_prof_enable": [_i],
    "ffb_prof_report": [C.c_char_p, _sz],
    "ffb_malloc": [_P(_vp), _sz],
    "ffb_free": [_vp],
    "ffb_memset_zero": [_vp, _sz],
    "ffb_h2d": [_vp, _vp, _sz],
    "ffb_d2h": [_vp, _vp, _sz],
    "ffb_d2d": [_vp, _vp, _sz],
    "ffb_host_alloc_pinned": [_P(_vp), _sz],
    "ffb_host_free_pinned": [_vp],
    "ffb_mem_info": [_P(_sz), _P(_sz)],
    "ffb_plan_create": [_P(_vp), _i, _P(_i64), _i, _i, _i, _i],
    "ffb_plan_destroy": [_vp],
    "ffb_plan_workspace_bytes": [_vp, _P(_sz)],
    "ffb_plan_describe": [_vp, C.c_char_p, _sz],
    "ffb_fft_forward": [_vp, _vp, _vp],
    "ffb_fft_inverse": [_vp, _vp, _vp],
    "ffb_fft_forward_ex": [_vp, _vp, _vp, _P(ffb_fuse)],
    "ffb_fft_inverse_ex": [_vp, _vp, _vp, _P(ffb_fuse)],
    "ffb_fft_inverse_multi": [_vp, _vp, _i, _P(_vp), _P(ffb_fuse)],
    "ffb_dist_unique_id": [_vp],
    "ffb_dist_init": [_P(_vp), _i, _i, _vp],
    "ffb_dist_destroy": [_vp],
    "ffb_dist_info": [_vp, _P(_i), _P(_i)],
    "ffb_dist_alltoall": [_vp, _vp, _vp, _sz],
    "ffb_plan_create_dist": [_P(_vp), _i, _P(_i64), _i, _vp, _i],
    "ffb_plan_dist_recv_buffers": [_vp, _P(_vp), _P(_vp), _P(_sz)],
    "ffb_plan_dist_set_peers": [_vp, _P(_vp), _P(_vp)],
    "ffb_plan_dist_set_exchange": [_vp, _i],
    "ffb_plan_dist_get_exchange": [_vp, _P(_i)],
    "ffb_dist_ipc_export": [_vp, _vp, _P(_sz)],
    "ffb_dist_ipc_open": [_vp, _sz, _P(_vp)],
    "ffb_dist_ipc_close": [_vp],
    "ffb_dist_barrier": [_vp],
    "ffb_wavenumbers": [_vp, _i64, _d, _i, _i],
    "ffb_ksq": [_vp, _vp, _vp, _vp, _vp, _P(ffb_desc)],
    "ffb_dealias": [_vp, _P(ffb_desc)],
    "ffb_make_filter": [_vp, _vp, _vp, _vp, _d, _d, _d, _d, _d, _d, _d, _P(ffb_desc)],
    "ffb_etd_coeffs": [_d, _P(ffb_coef), _i, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _P(_d)],
    "ffb_stage_fe": [_vp, _vp, _P(ffb_coef), _d, _vp, _i, _i64],
    "ffb_stage_rk4_substep": [_vp, _vp, _vp, _vp, _P(ffb_coef), _d, _i, _i64],
    "ffb_stage_rk4_final": [_vp, _vp, _vp, _vp, _vp, _vp, _P(ffb_coef), _d, _vp, _i, _i, _i64],
    "ffb_stage_lsrk54": [_vp, _vp, _vp, _P(ffb_coef), _d, _d, _d, _i, _vp, _i, _i64],
    "ffb_stage_etdrk4_substep12": [_vp, _P(ffb_coef), _vp, _P(ffb_coef), _vp, _i, _i64],
    "ffb_stage_etdrk4_substep3": [_vp, _P(ffb_coef), _vp, _P(ffb_coef), _vp, _vp, _i, _i64],
    "ffb_stage_etdrk4_update": [_vp, _P(ffb_coef), _P(ffb_coef), _P(ffb_coef), _P(ffb_coef), _vp, _vp, _vp, _vp, _vp, _i, _i64],
    "ffb_stage_ab3": [_vp, _vp, _vp, _vp, _P(ffb_coef), _d, _i64, _vp, _i, _i64],
    "ffb_ew_axpby": [_vp, _d, _vp, _d, _vp, _i, _i, _i64],
    "ffb_ew_mul_real": [_vp, _vp, _vp, _i, _i64],
    "ffb_ew_spectral_mul": [_vp, _vp, _d, _d, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _P(ffb_desc)],
    "ffb_parseval_sum": [_P(_d), _vp, _i, _i, _P(ffb_desc)],
    "ffb_ew_mul": [_vp, _vp, _i, _vp, _i, _i, _i64],
    "ffb_jacobianh": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "ffb_snapshot_create": [_P(_vp), _sz, _i],
    "ffb_snapshot_destroy": [_vp],
    "ffb_snapshot_begin": [_vp, _vp, _sz, _P(_i)],
    "ffb_snapshot_wait": [_vp, _i, _P(_vp)],
    "ffb_snapshot_ready": [_vp, _i, _P(_i)],
    "ffb_snapshot_release": [_vp, _i],
    "ffb_problem_create": [_P(_vp), _P(ffb_problem_config)],
    "ffb_problem_destroy": [_vp],
    "ffb_problem_sol": [_vp, _P(_vp), _P(_i64)],
    "ffb_problem_plan": [_vp, _P(_vp)],
    "ffb_problem_clock": [_vp, _P(_d), _P(_i64), _P(_d)],
    "ffb_problem_set_dt": [_vp, _d],
    "ffb_problem_bytes": [_vp, _P(_sz)],
    "ffb_problem_set_physical": [_vp, _vp],
    "ffb_problem_get_physical": [_vp, _vp],
    "ffb_step": [_vp, _i64],
    "ffb_step_until": [_vp, _d],
    "ffb_pipeline_create": [_P(_vp), _vp, _i],
    "ffb_pipeline_destroy": [_vp],
    "ffb_pipeline_submit": [_vp, _vp, _vp, _i64, _P(_i)],
    "ffb_pipeline_wait": [_vp, _i],
}

_lib = None


def load():
    """Load the shared library (once).  Raises if it has not been built: the product has no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FFBError(FFB_ECUDA, f"{LIB_PATH} not found: build it with `make -C fourierflows.jl_b200/csrc` "
                                  "(or `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
    lib = C.CDLL(LIB_PATH)
    lib.ffb_last_error.restype = C.c_char_p
    lib.ffb_last_error.argtypes = []
    for name, args in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == FFB_OK:
        return
    msg = load().ffb_last_error().decode(errors="replace")
    if rc == FFB_EDOMAIN:
        raise DomainError(msg)
    if rc == FFB_ENOMEM:
        raise MemoryError(msg)
    raise FFBError(rc, msg)


def call(name, *args):
    check(getattr(load(), name)(*args))


def have_device() -> bool:
    n = C.c_int(0)
    try:
        return load().ffb_device_count(C.byref(n)) == 0 and n.value > 0
    except (OSError, FFBError):
        return False


def prof_enable(on: bool) -> None:
    call("ffb_prof_enable", 1 if on else 0)


def prof_report():
    import json
    buf = C.create_string_buffer(1 << 16)
    call("ffb_prof_report", buf, 1 << 16)
    return json.loads(buf.value.decode())


def launch_count() -> int:
    n = C.c_uint64(0)
    call("ffb_launch_count", C.byref(n))
    return n.value
