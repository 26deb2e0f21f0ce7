"""ctypes binding of librdg_b200.so (C ABI declared in include/rdg_b200.h).

The library is the product: if it is missing or no sm_100 GPU is present the ops raise --
there is no CPU fallback (the CPU restatement lives in oracle/ and is test infrastructure).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RDG_LIB") or os.path.join(_HERE, "librdg_b200.so")     # RDG_LIB: another build of the same library (A/B timing)

MODE_FP32, MODE_BF16, MODE_FP16 = 0, 1, 2
OUT_FRACTION, OUT_MM = 0, 1
E_BADARG, E_NOWEIGHT, E_NONFINITE, E_NODEVICE, E_NOMEM = -1, -2, -3, -4, -5

MODES = {"fp32": MODE_FP32, "bf16": MODE_BF16, "fp16": MODE_FP16}

_c_float_p = C.POINTER(C.c_float)
_c_float_pp = C.POINTER(_c_float_p)
_c_size_p = C.POINTER(C.c_size_t)
_c_int_p = C.POINTER(C.c_int)

# name -> (restype, argtypes); every symbol include/rdg_b200.h declares
SIGNATURES = {
    "rdg_last_error": (C.c_char_p, []),
    "rdg_version": (C.c_int, []),
    "rdg_ctx_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.c_int, C.c_int, C.c_int]),
    "rdg_ctx_destroy": (None, [C.c_void_p]),
    "rdg_ctx_info": (C.c_int, [C.c_void_p, _c_int_p, _c_int_p, _c_int_p, _c_int_p]),
    "rdg_ctx_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "rdg_generator_set_weights": (C.c_int, [C.c_void_p, _c_float_pp, _c_size_p, C.c_int]),
    "rdg_generator_get_weights": (C.c_int, [C.c_void_p, _c_float_pp, _c_size_p, C.c_int]),
    "rdg_generator_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "rdg_generate_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong, C.c_int,
                                    C.c_int, C.c_float]),
    "rdg_ensemble_stats": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]),
    "rdg_generate_stats_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_longlong,
                                          C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_void_p]),
    "rdg_sample_windows": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "rdg_fill_normal": (C.c_int, [C.c_void_p, C.c_longlong, C.c_uint64, C.c_uint64, C.c_void_p]),
    "rdg_critic_set_weights": (C.c_int, [C.c_void_p, _c_float_pp, _c_size_p, C.c_int]),
    "rdg_critic_get_weights": (C.c_int, [C.c_void_p, _c_float_pp, _c_size_p, C.c_int]),
    "rdg_critic_forward": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_int,
                                     C.c_void_p]),
    "rdg_critic_forward_tc": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rdg_critic_step_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_int,
                                        C.c_int, C.c_void_p, C.c_void_p]),
    "rdg_generator_step_grads": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int,
                                           C.c_void_p, C.c_void_p]),
    "rdg_critic_input_grad": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "rdg_grad_buffer": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), _c_size_p]),
    "rdg_param_buffer": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), _c_size_p]),
    "rdg_adam_apply": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_longlong,
                                 C.c_float, C.c_void_p]),
    "rdg_adam_reset": (C.c_int, [C.c_void_p, C.c_int]),
    "rdg_set_train_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "rdg_critic_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_ulonglong, C.c_int,
                                      C.c_void_p, C.c_int, C.c_void_p]),
    "rdg_generator_step_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_ulonglong, C.c_int, C.c_void_p, C.c_int,
                                         C.c_void_p]),
    "rdg_adam_apply_dev": (C.c_int, [C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float, C.c_int,
                                     C.c_void_p]),
    "rdg_train_state": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_ulonglong)]),
    "rdg_adam_buffers": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), _c_size_p]),
    "rdg_peer_handle_bytes": (C.c_int, []),
    "rdg_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rdg_peer_connect": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "rdg_peer_disconnect": (C.c_int, [C.c_void_p]),
    "rdg_peer_allreduce": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "rdg_peer_status": (C.c_int, [C.c_void_p, _c_int_p, _c_int_p]),
    "rdg_pixelnorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_int, C.c_void_p]),
    "rdg_softmax_hours": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]),
    "rdg_conv3d": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                             C.c_int, C.c_void_p]),
    "rdg_profile_enable": (C.c_int, [C.c_void_p, C.c_int]),
    "rdg_profile_collect": (C.c_int, [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_longlong),
                                      C.POINTER(C.c_longlong)]),
    "rdg_launch_count": (C.c_longlong, [C.c_void_p]),
    "rdg_tc_layer": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
}

_lib = None


class RdgError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"rdg_b200 error {code}: {msg}")
        self.code = code


class NonFiniteError(FloatingPointError):
    """Python face of tf.debugging.check_numerics (gan_train_cwgangp_pixelnorm.py:349-350)."""


def load():
    """Load the shared library (no CUDA call is made by loading)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C pr-disagg-radar-gan_b200/csrc`. There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc):
    if rc == 0:
        return
    msg = load().rdg_last_error().decode(errors="replace")
    if rc == E_NONFINITE:
        raise NonFiniteError(msg or "found nan in output of per_gridpoint_softmax")
    raise RdgError(rc, msg)


def float_ptr_array(arrays):
    """(float**) and (size_t*) views over a list of contiguous float32 numpy arrays."""
    n = len(arrays)
    ptrs = (_c_float_p * n)(*[a.ctypes.data_as(_c_float_p) for a in arrays])
    sizes = (C.c_size_t * n)(*[a.size for a in arrays])
    return ptrs, sizes
