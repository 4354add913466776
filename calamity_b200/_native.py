"""ctypes binding of libcalamity_b200.so (include/calamity_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` (or ``python -m calamity_b200.build``).
There is NO CPU fallback: if the shared object is missing or no CUDA device is usable, calls raise.
"""
import ctypes as C
import os

import numpy as np

_LIB_PATH = os.environ.get("CALB2_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "_lib", "libcalamity_b200.so")
_lib = None


class NativeError(RuntimeError):
    pass


class PlanDesc(C.Structure):
    _fields_ = [
        ("device", C.c_int32),
        ("nants", C.c_int32),
        ("nfreqs", C.c_int32),
        ("ngroups", C.c_int32),
        ("group_ncomp", C.POINTER(C.c_int32)),
        ("group_nslots", C.POINTER(C.c_int32)),
        ("slot_nbls", C.POINTER(C.c_int32)),
        ("bl_ant0", C.POINTER(C.c_int32)),
        ("bl_ant1", C.POINTER(C.c_int32)),
        ("tile_freqs", C.c_int32),
        ("dtype", C.c_int32),
        ("group_class", C.POINTER(C.c_int32)),
        ("shared_basis", C.c_int32),
    ]


class FitOptions(C.Structure):
    _fields_ = [
        ("optimizer", C.c_int32),
        ("maxsteps", C.c_int32),
        ("learning_rate", C.c_double),
        ("beta_1", C.c_double),
        ("beta_2", C.c_double),
        ("epsilon", C.c_double),
        ("tol", C.c_double),
        ("use_min", C.c_int32),
        ("freeze_model", C.c_int32),
        ("regularization", C.c_int32),
        ("prior_r_sum", C.c_double),
        ("prior_i_sum", C.c_double),
        ("n_profile_steps", C.c_int32),
        ("steps_per_sync", C.c_int32),
        ("use_graph", C.c_int32),
        ("fuse_tail_update", C.c_int32),
        ("rho", C.c_double),
        ("momentum", C.c_double),
        ("initial_accumulator_value", C.c_double),
        ("l1_regularization_strength", C.c_double),
        ("l2_regularization_strength", C.c_double),
        ("learning_rate_power", C.c_double),
        ("nesterov", C.c_int32),
        ("weight_decay", C.c_double),
    ]


class FitResult(C.Structure):
    _fields_ = [
        ("nsteps_recorded", C.c_int32),
        ("nsteps_total", C.c_int32),
        ("final_loss", C.c_float),
        ("loop_ms", C.c_float),
        ("heavy_ms", C.c_float),
        ("heavy_launches", C.c_int64),
        ("kernel_launches", C.c_int64),
    ]


class PlanInfo(C.Structure):
    _fields_ = [
        ("n_d", C.c_int64),
        ("n_a_nz", C.c_int64),
        ("n_a_stored", C.c_int64),
        ("n_c_nz", C.c_int64),
        ("nbls_total", C.c_int64),
        ("nslots_total", C.c_int64),
        ("nitems", C.c_int64),
        ("tile_freqs", C.c_int32),
        ("rows_per_item_max", C.c_int32),
        ("device_bytes", C.c_int64),
        ("generic", C.c_int32),
        ("dtype", C.c_int32),
        ("n_classes", C.c_int64),
        ("n_class_slots", C.c_int64),
        ("n_class_ctas", C.c_int64),
        ("n_a_class", C.c_int64),
        ("n_a_class_nz", C.c_int64),
        ("class_fma", C.c_int64),
        ("n_tc_ctas", C.c_int64),
        ("n_tc_slots", C.c_int64),
    ]


# every symbol include/calamity_b200.h declares, with its argument types
_FP = C.c_void_p  # float32 or float64 elements, per the plan's dtype
_DP = C.POINTER(C.c_double)
_SIGNATURES = {
    "calb2_last_error": (C.c_char_p, []),
    "calb2_version": (C.c_char_p, []),
    "calb2_device_count": (C.c_int, [C.POINTER(C.c_int32)]),
    "calb2_debug_check_guards": (C.c_int, [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "calb2_debug_tc_record": (C.c_int, [C.POINTER(C.c_uint32)]),
    "calb2_plan_create": (C.c_int, [C.POINTER(PlanDesc), C.POINTER(C.c_void_p)]),
    "calb2_plan_destroy": (C.c_int, [C.c_void_p]),
    "calb2_plan_get_info": (C.c_int, [C.c_void_p, C.POINTER(PlanInfo)]),
    "calb2_plan_set_basis": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "calb2_debug_tc_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.c_int32]),
    "calb2_plan_set_variables": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_int64)]),
    "calb2_set_integration": (C.c_int, [C.c_void_p, _FP, _FP, _FP]),
    "calb2_set_gains": (C.c_int, [C.c_void_p, _FP, _FP]),
    "calb2_set_coeffs": (C.c_int, [C.c_void_p, _FP, _FP]),
    "calb2_init_coeffs": (C.c_int, [C.c_void_p, _FP, _FP]),
    "calb2_prior_sums": (C.c_int, [C.c_void_p, _FP, _FP, _DP, _DP]),
    "calb2_apply_model_snr_weights": (C.c_int, [C.c_void_p]),
    "calb2_fit": (C.c_int, [C.c_void_p, C.POINTER(FitOptions), _FP, C.POINTER(FitResult)]),
    "calb2_loss_and_grads": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_double, _DP, _FP, _FP, _FP, _FP]),
    "calb2_get_gains": (C.c_int, [C.c_void_p, _FP, _FP]),
    "calb2_get_coeffs": (C.c_int, [C.c_void_p, _FP, _FP]),
    "calb2_get_model": (C.c_int, [C.c_void_p, _FP, _FP]),
    "calb2_get_weights": (C.c_int, [C.c_void_p, _FP]),
    "calb2_comm_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_char_p]),
    "calb2_comm_unique_id": (C.c_int, [C.c_void_p, C.c_char_p]),
    "calb2_comm_peer_export": (C.c_int, [C.c_void_p, C.c_void_p]),
    "calb2_comm_peer_import": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    "calb2_comm_peer_close": (C.c_int, [C.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib_path():
    return _LIB_PATH


def load():
    """dlopen the library and declare every prototype; raises NativeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        raise NativeError(
            f"{_LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
            "calamity_b200 has no CPU fallback."
        )
    lib = C.CDLL(_LIB_PATH)
    for name, (restype, argtypes) in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        msg = load().calb2_last_error().decode("utf-8", "replace")
        raise NativeError(f"calamity_b200 native call failed ({rc}): {msg}")


DTYPE_IDS = {np.dtype(np.float32): 0, np.dtype(np.float64): 1}


def fptr(arr, dtype=np.float32):
    """C-contiguous ndarray of the plan's dtype -> void*; the caller keeps `arr` alive for the duration of the call."""
    assert arr.dtype == np.dtype(dtype) and arr.flags["C_CONTIGUOUS"], (arr.dtype, dtype, arr.flags)
    return arr.ctypes.data_as(C.c_void_p)


def iptr(arr):
    assert arr.dtype == np.int32 and arr.flags["C_CONTIGUOUS"]
    return arr.ctypes.data_as(C.POINTER(C.c_int32))


def find_nccl():
    """Path of the NCCL shared object torch itself uses (so both sides agree on the version)."""
    try:
        import torch

        base = os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib")
        cand = os.path.join(base, "libnccl.so.2")
        if os.path.exists(cand):
            return cand
    except Exception:
        pass
    return "libnccl.so.2"
