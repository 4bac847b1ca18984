"""The rgpot-core boundary (include/rgpot_b200_core.h) driven from Python.

`calculate(pot, positions, atomic_numbers, box)` does what `rgpot_potential_calculate` does with a
potential registered as `rgpot_potential_new(rgpot_b200_core_callback, handle, ...)`
(rgpot-core/include/rgpot.h:357-378): it wraps the arrays in borrowed `DLManagedTensorVersioned`
tensors, calls the callback, and takes ownership of the callee-allocated forces tensor (released
through its own deleter, like `rgpot_tensor_free`, rgpot-core/src/tensor.rs:393-401).

numpy arrays travel as kDLCPU tensors and come back as a numpy array; torch CUDA tensors travel as
kDLCUDA tensors (no host copies on the data path) and come back as a torch CUDA tensor.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib

K_DL_CPU, K_DL_CUDA = 1, 2
K_DL_INT, K_DL_FLOAT = 0, 2

STATUS = {0: "RGPOT_SUCCESS", 1: "RGPOT_INVALID_PARAMETER", 2: "RGPOT_INTERNAL_ERROR"}


class CoreError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


def _borrowed(ptr: int, shape, code: int, bits: int, device_type: int, device_id: int, strides=None):
    """A DLManagedTensorVersioned that borrows `ptr` (no deleter); keeps its shape array alive."""
    t = _lib.DLManagedTensorVersioned()
    shp = (C.c_int64 * len(shape))(*shape)
    t.version = _lib.DLPackVersion(1, 0)
    t.manager_ctx = None
    t.deleter = C.cast(None, _lib.DLDeleter)
    t.flags = 0
    t.dl_tensor.data = ptr
    t.dl_tensor.device = _lib.DLDevice(device_type, device_id)
    t.dl_tensor.ndim = len(shape)
    t.dl_tensor.dtype = _lib.DLDataType(code, bits, 1)
    t.dl_tensor.shape = shp
    if strides is not None:
        st = (C.c_int64 * len(strides))(*strides)
        t.dl_tensor.strides = st
        t._strides = st
    else:
        t.dl_tensor.strides = C.cast(None, C.POINTER(C.c_int64))
    t.dl_tensor.byte_offset = 0
    t._shape = shp
    return t


def calculate(pot, positions, atomic_numbers=None, box=None):
    """(energy, forces, variance) through rgpot_b200_core_callback; `pot` is a host-side potential
    object of rgpot_b200.potentials (its engine handle is the callback's user_data)."""
    lib = _lib.lib()
    on_device = hasattr(positions, "is_cuda") and positions.is_cuda
    keep = []
    if on_device:
        import torch
        dev = positions.device.index or 0
        pos = positions.contiguous()
        tp = _borrowed(pos.data_ptr(), tuple(pos.shape), K_DL_FLOAT, 64, K_DL_CUDA, dev)
        tz = None
        if atomic_numbers is not None:
            z = atomic_numbers.to(device=positions.device, dtype=torch.int32).contiguous()
            tz = _borrowed(z.data_ptr(), tuple(z.shape), K_DL_INT, 32, K_DL_CUDA, dev)
            keep.append(z)
        keep.append(pos)
    else:
        pos = np.ascontiguousarray(positions, dtype=np.float64)
        tp = _borrowed(pos.ctypes.data, pos.shape, K_DL_FLOAT, 64, K_DL_CPU, 0)
        tz = None
        if atomic_numbers is not None:
            z = np.ascontiguousarray(atomic_numbers, dtype=np.int32)
            tz = _borrowed(z.ctypes.data, z.shape, K_DL_INT, 32, K_DL_CPU, 0)
            keep.append(z)
        keep.append(pos)
    tb = None
    if box is not None:
        b = np.ascontiguousarray(np.asarray(box.cpu() if hasattr(box, "cpu") else box, dtype=np.float64).reshape(3, 3))
        tb = _borrowed(b.ctypes.data, (3, 3), K_DL_FLOAT, 64, K_DL_CPU, 0)
        keep.append(b)
    fin = _lib.ForceInput(C.pointer(tp), C.pointer(tz) if tz is not None else None,
                          C.pointer(tb) if tb is not None else None)
    out = _lib.ForceOut()
    status = lib.rgpot_b200_core_callback(pot._h, C.byref(fin), C.byref(out))
    if status != 0:
        raise CoreError(status, lib.rgpot_b200_core_last_error().decode())
    F = out.forces.contents
    try:
        dl = F.dl_tensor
        assert (F.version.major, F.version.minor) == (1, 0) and F.flags & 2  # IS_COPIED
        n = dl.shape[0]
        assert dl.ndim == 2 and dl.shape[1] == 3 and (dl.strides[0], dl.strides[1]) == (3, 1)
        if dl.device.device_type == K_DL_CUDA:
            import torch
            forces = torch.empty((n, 3), dtype=torch.float64, device=positions.device)
            if n:
                C.CDLL("libcudart.so.12").cudaMemcpy(C.c_void_p(forces.data_ptr()), C.c_void_p(dl.data),
                                                     C.c_size_t(24 * n), 3)  # device to device
        else:
            forces = np.ctypeslib.as_array(C.cast(dl.data, C.POINTER(C.c_double)), shape=(n, 3)).copy() if n \
                else np.zeros((0, 3))
    finally:
        F.deleter(out.forces)  # rgpot_tensor_free
    return out.energy, forces, out.variance
