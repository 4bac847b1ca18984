"""ctypes binding of include/rgpot_b200.h (the engine's C ABI).

Loading never falls back to anything: if librgpot_b200.so is missing the import of the
binding raises, and if no B200 is usable every call returns a CUDA status which the host
layer turns into an exception.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
# RGPOT_B200_ENGINE selects another build of the same ABI (A/B timing of kernel variants)
LIB_PATH = os.environ.get("RGPOT_B200_ENGINE") or os.path.join(HERE, "librgpot_b200.so")
CSRC = os.path.join(HERE, "csrc")

KIND_LJ, KIND_LJ_CLUSTER, KIND_CUH2, KIND_MORSE, KIND_ZBL, KIND_EAM_AL, KIND_FEHE = 0, 1, 2, 3, 4, 5, 6


class Config(C.Structure):
    """RgpotB200Config"""

    _fields_ = [
        ("kind", C.c_int),
        ("device", C.c_int),
        ("u0", C.c_double),
        ("cutoff", C.c_double),
        ("psi", C.c_double),
        ("lj_all_images", C.c_int),
        ("n_devices", C.c_int),
        ("host_threads", C.c_int),
        ("reserved", C.c_int * 5),
        ("morse_De", C.c_double),
        ("morse_a", C.c_double),
        ("morse_re", C.c_double),
        ("zbl_cut_inner", C.c_double),
        ("devices", C.c_int * 16),
    ]


# ---- include/rgpot_b200_core.h: DLPack 1.0 subset and the rgpot-core carriers (rgpot.h:76-117) ----
class DLPackVersion(C.Structure):
    _fields_ = [("major", C.c_uint32), ("minor", C.c_uint32)]


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int32)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int32), ("dtype", DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)),
                ("byte_offset", C.c_uint64)]


class DLManagedTensorVersioned(C.Structure):
    pass


DLDeleter = C.CFUNCTYPE(None, C.POINTER(DLManagedTensorVersioned))
DLManagedTensorVersioned._fields_ = [("version", DLPackVersion), ("manager_ctx", C.c_void_p),
                                     ("deleter", DLDeleter), ("flags", C.c_uint64), ("dl_tensor", DLTensor)]


class NeighborList(C.Structure):
    """RgpotB200NeighborList"""
    _fields_ = [("length", C.c_size_t), ("pairs", C.POINTER(C.c_size_t)), ("shifts", C.POINTER(C.c_int32)),
                ("distances", C.POINTER(C.c_double)), ("vectors", C.POINTER(C.c_double))]


class ForceInput(C.Structure):
    """rgpot_force_input_t"""
    _fields_ = [("positions", C.POINTER(DLManagedTensorVersioned)),
                ("atomic_numbers", C.POINTER(DLManagedTensorVersioned)),
                ("box_matrix", C.POINTER(DLManagedTensorVersioned))]


class ForceOut(C.Structure):
    """rgpot_force_out_t"""
    _fields_ = [("forces", C.POINTER(DLManagedTensorVersioned)), ("energy", C.c_double),
                ("variance", C.c_double)]


# every symbol include/*.h declares: (name, restype, argtypes)
_vp, _sz, _i, _l, _d = C.c_void_p, C.c_size_t, C.c_int, C.c_long, C.c_double
SYMBOLS = [
    ("rgpot_b200_abi_version", _i, []),
    ("rgpot_b200_available", _i, []),
    ("rgpot_b200_default_config", None, [_i, C.POINTER(Config)]),
    ("rgpot_b200_create", _vp, [C.POINTER(Config), C.c_char_p, _sz]),
    ("rgpot_b200_destroy", None, [_vp]),
    ("rgpot_b200_force", _i, [_vp, _l, _vp, _vp, _vp, C.POINTER(_d), C.POINTER(_d), _vp]),
    ("rgpot_b200_force_batch", _i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("rgpot_b200_force_batch_packed", _i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("rgpot_b200_force_device", _i, [_vp, _l, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("rgpot_b200_force_batch_device", _i, [_vp, _sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    ("rgpot_b200_sync_status", _i, [_vp]),
    ("rgpot_b200_neighbors", _i, [_vp, _l, _vp, _vp, _vp, _i, _l, _vp, _vp, C.POINTER(_l)]),
    ("rgpot_b200_neighbor_list", _i, [_vp, _sz, _vp, _vp, _vp, _d, _i, _i, _vp]),
    ("rgpot_b200_neighbor_list_free", None, [_vp]),
    ("rgpot_b200_cache_keys", _i, [_vp, _sz, _vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp]),
    ("rgpot_b200_cache_keys_device", _i, [_vp, _sz, _vp, _vp, _vp, _vp, C.c_uint64, C.c_uint64, _vp, _vp]),
    ("rgpot_b200_hash64", C.c_uint64, [C.c_char_p, _sz]),
    ("rgpot_b200_last_error", _i, [_vp, C.c_char_p, _i]),
    ("rgpot_b200_launch_count", _l, [_vp]),
    ("rgpot_b200_stat", _l, [_vp, C.c_char_p]),
    ("rgpot_b200_set_option", _i, [_vp, C.c_char_p, _l]),
    ("rgpot_b200_host_alloc", _vp, [_sz]),
    ("rgpot_b200_host_free", None, [_vp]),
    ("rgpot_b200_measure_fp64_tflops", _d, [_i, _i]),
    ("rgpot_b200_measure_copy_gbs", _d, [_vp, _i, _sz, _sz, _i]),
    ("rgpot_b200_measure_host_copy_gbs", _d, [_i, _sz, _i]),
    ("rgpot_b200_core_callback", _i, [_vp, _vp, _vp]),
    ("rgpot_b200_core_free", None, [_vp]),
    ("rgpot_b200_core_last_error", C.c_char_p, []),
    ("rgpot_cuh2_force", _i, [C.c_int32, _vp, _vp, _vp, _vp, C.POINTER(_d)]),
    ("rgpot_eam_al_force", _i, [C.c_int32, _vp, _vp, _vp, C.POINTER(_d)]),
    ("rgpot_fehe_force", _i, [C.c_int32, _vp, _vp, _vp, _vp, C.POINTER(_d)]),
    ("rgpot_fortran_last_error", _i, [C.c_char_p, _i]),
]


def build(verbose: bool = False) -> str:
    """Compile the CUDA engine for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-s", "-C", CSRC], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_LIB = None


def lib():
    """The loaded engine.  Raises if the shared library is absent -- there is no other path."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C rgpot_b200/csrc` "
                "(or __graft_entry__.build()); rgpot_b200 has no CPU fallback")
        handle = C.CDLL(LIB_PATH)
        for name, restype, argtypes in SYMBOLS:
            fn = getattr(handle, name)  # AttributeError if the ABI lost a symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _LIB = handle
    return _LIB
