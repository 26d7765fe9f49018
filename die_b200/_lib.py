"""ctypes binding of libdie_sm100a.so (include/die_b200.h).  There is no fallback: if the
library is missing this module raises, and so does everything that imports it."""
import ctypes as C
import os

from ._build import LIB_PATH

DIE_MAX_RADIUS = 8
BOUNDARY_WRAP, BOUNDARY_LIMIT, BOUNDARY_NONE = 0, 1, 2
DIFFUSE_MODES = {'wrap': 0, 'reflect': 1, 'nearest': 2, 'mirror': 3, 'constant': 4}
FIELD_F64, FIELD_F32 = 0, 1
FWD_USE_GRADIENT, FWD_USE_CELLS, FWD_SPECULATE_MOVE, FWD_STEP_ON_DEVICE, FWD_COMMIT_MOVE, FWD_WRITE_COST = 1, 2, 4, 8, 16, 32
STEP_ADOPT_MOVE, STEP_ALIVE_BITS, STEP_USE_COST = 1, 2, 4
HOST_KEEP_ALIVE_CHANNEL = 1


class DieDynamics(C.Structure):
    _fields_ = [
        ("rate_feed", C.c_double),
        ("rate_decay_chem", C.c_double),
        ("cost_w_deposit", C.c_double),
        ("cost_w_dist", C.c_double),
        ("blur_w", C.c_double * (2 * DIE_MAX_RADIUS + 1)),
        ("blur_radius", C.c_int32),
        ("boundary", C.c_int32),
        ("food_infinite", C.c_int32),
        ("diffuse_mode", C.c_int32),
        ("agents_die", C.c_int32),
    ]


class DieGradientParams(C.Structure):
    _fields_ = [
        ("scale", C.c_double),
        ("deposit", C.c_double),
        ("inertia", C.c_double),
        ("sense_offset", C.c_double),
        ("noise_scale", C.c_double),
        ("grad_clip", C.c_double),
        ("turn_radians", C.c_double),
        ("sense_radians", C.c_double),
        ("turn_tolerance", C.c_double),
        ("normalized_grad", C.c_int32),
        ("use_grad_clip", C.c_int32),
        ("discrete_turn", C.c_int32),
        ("reserved", C.c_int32),
    ]


class DieJonesParams(C.Structure):
    _fields_ = [("scale", C.c_double), ("deposit", C.c_double), ("sense_offset", C.c_double),
                ("sense_radians", C.c_double), ("turn_radians", C.c_double)]


DIE_MAX_RANKS = 8


class DieSlabGeom(C.Structure):
    _fields_ = [
        ("G", C.c_int32), ("rank", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("M", C.c_int64),
        ("s0", C.c_int64 * DIE_MAX_RANKS), ("n0", C.c_int64 * DIE_MAX_RANKS),
        ("s1", C.c_int64 * DIE_MAX_RANKS), ("n1", C.c_int64 * DIE_MAX_RANKS),
    ]


class DieError(RuntimeError):
    pass


# every symbol include/die_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "die_version": (C.c_char_p, []),
    "die_last_error": (C.c_char_p, []),
    "die_env_create": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.POINTER(DieDynamics), C.POINTER(_P)]),
    "die_env_destroy": (C.c_int, [_P]),
    "die_env_set_dynamics": (C.c_int, [_P, C.POINTER(DieDynamics)]),
    "die_env_set_food_flow": (C.c_int, [_P, _P, _P, _P, _P, C.c_int64, C.c_int64, C.c_double, C.c_double]),
    "die_env_set_food_frames": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_double, C.c_double]),
    "die_env_step": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P]),
    "die_env_cells": (_P, [_P]),
    "die_env_publish_gradient": (C.c_int, [_P, C.c_int32]),
    "die_env_gradient": (_P, [_P]),
    "die_env_gradient_kind": (C.c_int, [_P]),
    "die_env_set_profiling": (C.c_int, [_P, C.c_int32]),
    "die_env_kernel_times": (C.c_int, [_P, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "die_env_read_stats": (C.c_int, [_P, _P, _P, _P, _P, _P]),
    "die_env_step_host": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "die_env_step_host_dev": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "die_env_step_host_flags": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "die_sense_mask": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, C.c_int32, _P, _P, _P]),
    "die_render_frames": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P, _P, _P, C.c_double, _P, _P, _P, _P]),
    "die_brownian_forward": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_double, C.c_double, _P,
                                       C.c_uint64, C.c_uint64, _P]),
    "die_brownian_forward_dev": (C.c_int, [_P, _P, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_uint64, _P, _P]),
    "die_conv_policy_forward": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32,
                                          C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "die_conv_policy_forward_population": (C.c_int, [C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_int32, _P, C.c_int32,
                                                     C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int64,
                                                     _P, _P, _P, _P, _P, _P, _P, _P]),
    "die_device_l2_fetch_granularity": (C.c_int, [C.c_int32, _P]),
    "die_const_forward": (C.c_int, [_P, C.c_int64, C.c_int32, C.c_double, C.c_double, C.c_double, _P]),
    "die_jones_forward": (C.c_int, [C.POINTER(DieJonesParams), C.c_int32, C.c_int32, C.c_int64, C.c_int32, _P, _P, C.c_int32,
                                    _P, _P, _P, C.c_uint64, C.c_uint64, C.c_int32, _P]),
    "die_env_set_field_dtype": (C.c_int, [_P, C.c_int32]),
    "die_env_field_dtype": (C.c_int, [_P]),
    "die_gradient_forward_f32": (C.c_int, [C.POINTER(DieGradientParams), C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                           _P, _P, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_uint64, _P]),
    "die_gradient_forward": (C.c_int, [C.POINTER(DieGradientParams), C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                       _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_uint64, _P]),
    "die_host_ctx_create": (C.c_int, [C.POINTER(_P)]),
    "die_host_ctx_destroy": (C.c_int, [_P]),
    "die_gradient_forward_host": (C.c_int, [_P, C.POINTER(DieGradientParams), C.c_int32, C.c_int32, C.c_int64, C.c_int32,
                                            _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, C.c_uint64, C.c_uint64, _P]),
    "die_env_forward_gradient": (C.c_int, [_P, C.POINTER(DieGradientParams), _P, _P, _P, _P, _P, _P, _P, _P,
                                           C.c_int32, C.c_uint64, C.c_uint64, _P]),
    "die_env_step_flags": (C.c_int, [_P, _P, _P, _P, _P, _P, _P, C.c_int32, _P]),
    "die_env_discard_move": (C.c_int, [_P, _P]),
    "die_env_pending_move": (C.c_int, [_P]),
    "die_env_refresh_alive": (C.c_int, [_P, _P, _P]),
    "die_set_turn_quick": (C.c_int, [C.c_int32]),
    "die_set_tuning": (C.c_int, [C.c_char_p, C.c_int32]),
    "die_get_counter": (C.c_int64, [C.c_char_p]),
    "die_slab_create": (C.c_int, [C.POINTER(DieSlabGeom), C.POINTER(DieDynamics), C.POINTER(_P)]),
    "die_slab_destroy": (C.c_int, [_P]),
    "die_slab_bind": (C.c_int, [_P, _P, _P, _P, _P, _P, _P]),
    "die_slab_forward": (C.c_int, [_P, C.POINTER(DieGradientParams), C.c_int32, _P, _P, _P, _P, C.c_int32,
                                   C.c_uint64, C.c_uint64, _P]),
    "die_slab_move_claim": (C.c_int, [_P, _P, _P, _P]),
    "die_slab_field": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "die_slab_feed": (C.c_int, [_P, _P, _P, _P, _P]),
    "die_slab_cells": (_P, [_P]),
    "die_slab_set_corner_mirror": (C.c_int, [_P, C.c_int32]),
    "die_slab_corner_refresh": (C.c_int, [_P, C.c_int32, C.c_int32, _P]),
    "die_set_step_impl": (C.c_int, [C.c_int32]),
    "die_math_sincos": (C.c_int, [_P, _P, _P, C.c_int64, _P]),
    "die_math_atan2": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, _P]),
}

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing. die_b200 has no CPU / PyTorch fallback: build the CUDA library "
            f"with `python die_b200/_build.py` (needs nvcc) before use.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class _NoGuard:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def on_device(device):
    """``torch.cuda.device(device)`` only when `device` is not already current (the context manager costs a few
    microseconds per call, which is most of the host time of a step on a small environment)."""
    import torch
    if torch.cuda.current_device() == device.index:
        return _NO_GUARD
    return torch.cuda.device(device)


def check(rc: int) -> None:
    if rc != 0:
        raise DieError(f"libdie_sm100a error {rc}: {load().die_last_error().decode()}")
