"""Host-side mirror of the reference's potential interface, on top of the engine's C ABI.

Names, argument meaning and error behaviour follow the reference:

* ``LJPot`` / ``LJClusterPot`` / ``CuH2Pot`` -- ``Potential<D>::operator()(positions, atmtypes,
  box) -> (energy, forces, variance)`` (CppCore/rgpot/Potential.hpp:165-211), constructed from
  the same plain config aggregates (LJPot.hpp:30-34, LJClusterPot.hpp:34-38).
* ``forceBatch`` -- ``PotentialBase::forceBatch(const ForceBatch&)`` (Potential.hpp:89-91,
  251-312); here it is native (``caps().batched`` is True).
* ``evaluate_lj`` / ``LJPot.__call__`` -- the nanobind module surface
  (python/bind/bind_module.cpp:84-96, 160-173): shape errors raise ``ValueError``, outputs are
  fresh numpy arrays.
* failures of the CuH2 kernel raise ``RuntimeError("CuH2 potential failed (status k): msg")``
  like CppCore/rgpot/fortran/FortranPots.cc:44-56.

The compute is entirely in the CUDA engine; this module only marshals numpy / torch buffers.
"""
from __future__ import annotations

import ctypes as C
import struct
from dataclasses import dataclass
from enum import Enum

import numpy as np

from . import _lib


# ----------------------------------------------------------------------------- capabilities
class Reentrancy(Enum):
    """CppCore/rgpot/pot_caps.hpp:20-29"""

    SharedInstance = 0
    PerInstance = 1
    ProcessSerial = 2


@dataclass(frozen=True)
class PotCaps:
    """CppCore/rgpot/pot_caps.hpp:34-45"""

    reentrancy: Reentrancy = Reentrancy.SharedInstance
    perImageInstances: bool = False
    batched: bool = False
    periodic: bool = True


class PotType(Enum):
    """The three members of CppCore/rgpot/pot_types.hpp:26-45 this engine serves."""

    CuH2 = 1
    LJ = 2
    Morse = 8
    LJCluster = 9
    ZBL = 10
    EAMAl = 15
    FeHe = 16


@dataclass
class LJConfig:
    """LJPot.hpp:30-34"""

    u0: float = 1.0
    cutoff: float = 15.0
    psi: float = 1.0


@dataclass
class LJClusterConfig:
    """LJClusterPot.hpp:34-38"""

    u0: float = 1.0
    cutoff: float = 15.0
    psi: float = 1.0


@dataclass
class MorseConfig:
    """Morse/MorsePot.hpp:31-36 (platinum defaults)"""

    De: float = 0.7102
    a: float = 1.6047
    re: float = 2.8970
    cutoff: float = 9.5


@dataclass
class ZBLConfig:
    """ZBL/ZBLPot.hpp:29-32"""

    cut_inner: float = 2.0
    cut_global: float = 2.5


class PotentialError(RuntimeError):
    def __init__(self, label: str, status: int, message: str):
        text = f"{label} potential failed (status {status})"
        if message:
            text += ": " + message
        super().__init__(text)
        self.status = status
        self.message = message


# kernel-version salt mixed into paramsKey (ParamHash.hpp:19-51, LJPot.hpp:54-59): a different
# value from the reference's CPU kernels (1) so a shared result cache never mixes the two.
_KERNEL_VERSION = 0xB2000001


def _fnv1a(chunks) -> int:
    h = 0xCBF29CE484222325
    for chunk in chunks:
        for b in chunk:
            h ^= b
            h = (h * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def _as_positions(positions) -> np.ndarray:
    arr = np.asarray(positions)
    if arr.ndim != 2 or arr.shape[1] != 3:
        raise ValueError("positions must have shape (n_atoms, 3)")
    return np.ascontiguousarray(arr, dtype=np.float64)


def _as_types(atom_types, n: int) -> np.ndarray:
    if atom_types is None:  # potentials that do not read atomic numbers (LJ, Morse, EAM-Al)
        return np.zeros(n, dtype=np.int32)
    arr = np.asarray(atom_types)
    if arr.ndim != 1 or arr.shape[0] != n:
        raise ValueError("atom_types length must match n_atoms")
    return np.ascontiguousarray(arr, dtype=np.int32)


def _as_box(box) -> np.ndarray:
    arr = np.asarray(box)
    if arr.ndim != 2 or arr.shape != (3, 3):
        raise ValueError("box must have shape (3, 3)")
    return np.ascontiguousarray(arr, dtype=np.float64)


class _EnginePot:
    """One engine handle (one CUDA device, one stream, private scratch)."""

    _kind = _lib.KIND_LJ
    _label = "LJ"
    _type = PotType.LJ

    def __init__(self, u0=1.0, cutoff=15.0, psi=1.0, device: int = -1, lj_all_images: bool = False,
                 morse=None, zbl=None, devices=None, host_threads: int = 0):
        self._lib = _lib.lib()
        cfg = _lib.Config()
        self._lib.rgpot_b200_default_config(self._kind, C.byref(cfg))
        cfg.device = int(device)
        cfg.host_threads = int(host_threads)
        # batches on several GPUs of the box through ONE call (RgpotB200Config.devices): a list of
        # ordinals, or "all"
        if isinstance(devices, str):
            if devices != "all":
                raise ValueError('devices must be a list of CUDA ordinals or "all"')
            cfg.device = -2
        elif devices is not None:
            devices = [int(d) for d in devices]
            if not 1 <= len(devices) <= 16:
                raise ValueError("devices must name 1 to 16 CUDA ordinals")
            cfg.device = devices[0]
            cfg.n_devices = len(devices)
            for k, d in enumerate(devices):
                cfg.devices[k] = d
        if self._kind == _lib.KIND_MORSE:
            cfg.morse_De, cfg.morse_a, cfg.morse_re = float(morse.De), float(morse.a), float(morse.re)
            cfg.cutoff = float(morse.cutoff)
        elif self._kind == _lib.KIND_ZBL:
            cfg.zbl_cut_inner, cfg.cutoff = float(zbl.cut_inner), float(zbl.cut_global)
        elif self._kind not in (_lib.KIND_CUH2, _lib.KIND_EAM_AL, _lib.KIND_FEHE):
            cfg.u0, cfg.cutoff, cfg.psi = float(u0), float(cutoff), float(psi)
            cfg.lj_all_images = int(bool(lj_all_images))
        self._cfg = cfg
        err = C.create_string_buffer(512)
        self._h = self._lib.rgpot_b200_create(C.byref(cfg), err, 512)
        if not self._h:
            msg = err.value.decode() or "rgpot_b200_create failed"
            if "invalid cutoffs" in msg:
                raise ValueError(msg)  # std::invalid_argument upstream (ZBLPot.cc:76-79)
            raise RuntimeError(msg)
        self.force_calls = 0  # registry<T>::forceCalls (PotHelpers.hpp:31-34)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._lib.rgpot_b200_destroy(h)
            self._h = None

    # -- interface mirror ---------------------------------------------------------------------
    def get_type(self) -> PotType:
        return self._type

    def caps(self) -> PotCaps:
        return PotCaps(reentrancy=Reentrancy.PerInstance, batched=True,
                       periodic=self._kind not in (_lib.KIND_LJ_CLUSTER, _lib.KIND_ZBL))

    def paramsKey(self) -> int:
        if self._kind == _lib.KIND_CUH2:
            return _fnv1a([struct.pack("<Q", _KERNEL_VERSION)])
        if self._kind in (_lib.KIND_EAM_AL, _lib.KIND_FEHE):  # parameter-free like CuH2; the kind keeps them apart
            return _fnv1a([struct.pack("<Q", _KERNEL_VERSION), struct.pack("<i", self._kind)])
        c = self._cfg
        if self._kind == _lib.KIND_ZBL:  # ZBLPot.cc:80-84
            return _fnv1a([struct.pack("<Q", _KERNEL_VERSION), struct.pack("<d", c.zbl_cut_inner),
                           struct.pack("<d", c.cutoff)])
        if self._kind == _lib.KIND_MORSE:  # MorsePot.hpp:62-68
            return _fnv1a([struct.pack("<Q", _KERNEL_VERSION), struct.pack("<d", c.morse_De),
                           struct.pack("<d", c.morse_a), struct.pack("<d", c.morse_re),
                           struct.pack("<d", c.cutoff)])
        return _fnv1a([struct.pack("<Q", _KERNEL_VERSION), struct.pack("<d", c.u0),
                       struct.pack("<d", c.cutoff), struct.pack("<d", c.psi)])

    def _last_error(self) -> str:
        buf = C.create_string_buffer(512)
        self._lib.rgpot_b200_last_error(self._h, buf, 512)
        return buf.value.decode()

    def _raise(self, status: int):
        raise PotentialError(self._label, status, self._last_error())

    def set_path(self, path: int) -> None:
        """0 automatic, 1 cell-list kernels, 2 all-pairs kernels (test hook)."""
        self._lib.rgpot_b200_set_option(self._h, b"path", int(path))

    def set_option(self, key: str, value: int) -> None:
        """Tuning / test hooks of the engine (rgpot_b200_set_option): "path", "brick" (cells per
        brick as decimal xyz, 0 = automatic), "brick_tpa", "brick_threads"."""
        if self._lib.rgpot_b200_set_option(self._h, key.encode(), int(value)) != 0:
            raise ValueError(f"unknown engine option {key!r}")

    def launch_count(self) -> int:
        return int(self._lib.rgpot_b200_launch_count(self._h))

    def stat(self, key: str) -> int:
        """Engine diagnostics (rgpot_b200_stat): "launches", "slow_atoms"."""
        return int(self._lib.rgpot_b200_stat(self._h, key.encode()))

    def __call__(self, positions, atom_types, box, out=None):
        """(energy, forces, variance).  `out`, if given, is a C-contiguous float64 (n, 3) array that
        receives the forces (e.g. a `pinned_empty` buffer reused across calls)."""
        pos = _as_positions(positions)
        types = _as_types(atom_types, pos.shape[0])
        cell = _as_box(box)
        if out is None:
            forces = np.empty_like(pos)  # the engine overwrites (or zeroes) every element
        else:
            forces = out
            if forces.shape != pos.shape or forces.dtype != np.float64 or not forces.flags.c_contiguous:
                raise ValueError("out must be a C-contiguous float64 array of shape (n_atoms, 3)")
        e, v = C.c_double(0.0), C.c_double(0.0)
        st = self._lib.rgpot_b200_force(self._h, pos.shape[0], pos.ctypes.data, types.ctypes.data,
                                        forces.ctypes.data, C.byref(e), C.byref(v), cell.ctypes.data)
        if st != 0:
            self._raise(st)
        self.force_calls += 1
        return e.value, forces, v.value

    def forceImpl(self, positions, atom_types, box):
        e, f, _ = self(positions, atom_types, box)
        return e, f

    def forceBatch(self, systems, return_statuses: bool = False):
        """systems: sequence of (positions, atom_types, box).  Returns [(energy, forces, variance)].

        Native batch: one engine call, independent systems of any sizes (ForceStructs.hpp:39-54).
        """
        n = len(systems)
        if n == 0:
            return []
        pos = [_as_positions(s[0]) for s in systems]
        typ = [_as_types(s[1], p.shape[0]) for s, p in zip(systems, pos)]
        box = [_as_box(s[2]) for s in systems]
        frc = [np.zeros_like(p) for p in pos]
        natoms = (C.c_size_t * n)(*[p.shape[0] for p in pos])
        pp = (C.c_void_p * n)(*[p.ctypes.data for p in pos])
        zz = (C.c_void_p * n)(*[t.ctypes.data for t in typ])
        bb = (C.c_void_p * n)(*[b.ctypes.data for b in box])
        ff = (C.c_void_p * n)(*[f.ctypes.data for f in frc])
        energies = np.zeros(n)
        statuses = np.zeros(n, dtype=np.int32)
        st = self._lib.rgpot_b200_force_batch(self._h, n, natoms, pp, zz, bb, ff, energies.ctypes.data,
                                              statuses.ctypes.data)
        if st != 0 and not return_statuses:
            self._raise(st)
        self.force_calls += n
        out = [(float(energies[i]), frc[i], 0.0) for i in range(n)]
        return (out, statuses) if return_statuses else out

    def prepareForceBatch(self, systems, forces=None):
        """The ``ForceBatch`` a C++ caller hands to ``Potential::forceBatch`` (ForceStructs.hpp:50-54),
        built once: parallel arrays of pointers to every system's own (separately allocated,
        pageable) buffers.  ``forceBatchPrepared`` then is one rgpot_b200_force_batch call with no
        Python marshalling in it."""
        n = len(systems)
        pos = [_as_positions(s[0]) for s in systems]
        typ = [_as_types(s[1], p.shape[0]) for s, p in zip(systems, pos)]
        box = [_as_box(s[2]) for s in systems]
        frc = forces if forces is not None else [np.zeros_like(p) for p in pos]
        fb = {"n": n, "pos": pos, "typ": typ, "box": box, "forces": frc,
              "natoms": (C.c_size_t * n)(*[p.shape[0] for p in pos]),
              "pp": (C.c_void_p * n)(*[p.ctypes.data for p in pos]),
              "zz": (C.c_void_p * n)(*[t.ctypes.data for t in typ]),
              "bb": (C.c_void_p * n)(*[b.ctypes.data for b in box]),
              "ff": (C.c_void_p * n)(*[f.ctypes.data for f in frc]),
              "energies": np.zeros(n), "statuses": np.zeros(n, dtype=np.int32)}
        return fb

    def forceBatchPrepared(self, fb, check: bool = True) -> int:
        """Evaluate a prepared ForceBatch in place (forces in fb["forces"], energies in fb["energies"])."""
        st = self._lib.rgpot_b200_force_batch(self._h, fb["n"], fb["natoms"], fb["pp"], fb["zz"], fb["bb"],
                                              fb["ff"], fb["energies"].ctypes.data, fb["statuses"].ctypes.data)
        if st != 0 and check:
            self._raise(st)
        self.force_calls += fb["n"]
        return int(st)

    def forceBatchPacked(self, offsets, positions, atom_types, boxes, forces=None, energies=None,
                         statuses=None):
        """Back-to-back layout (rgpot_b200_force_batch_packed); buffers may be pinned arrays from
        ``pinned_empty``.  Returns (energies, forces)."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        nsys = len(offsets) - 1
        total = int(offsets[-1])
        positions = np.ascontiguousarray(positions, dtype=np.float64)
        boxes = np.ascontiguousarray(boxes, dtype=np.float64)
        if atom_types is not None:
            atom_types = np.ascontiguousarray(atom_types, dtype=np.int32)
        if forces is None:
            forces = np.zeros((total, 3))
        if energies is None:
            energies = np.zeros(nsys)
        st = self._lib.rgpot_b200_force_batch_packed(
            self._h, nsys, offsets.ctypes.data, positions.ctypes.data,
            atom_types.ctypes.data if atom_types is not None else None, boxes.ctypes.data,
            forces.ctypes.data, energies.ctypes.data,
            statuses.ctypes.data if statuses is not None else None)
        if st != 0 and statuses is None:
            self._raise(st)
        self.force_calls += nsys
        return energies, forces

    # -- device-resident entry points (torch CUDA tensors carry the memory) ---------------------
    def force_device(self, d_pos, d_Z, box, d_F, d_E, stream=None, sync: bool = True):
        cell = _as_box(box)
        s = self._lib.rgpot_b200_force_device(
            self._h, int(d_pos.shape[0]), d_pos.data_ptr(), d_Z.data_ptr() if d_Z is not None else None,
            d_F.data_ptr(), d_E.data_ptr(), cell.ctypes.data, stream)
        if s != 0:
            self._raise(s)
        if sync:
            s = self._lib.rgpot_b200_sync_status(self._h)
            if s != 0:
                self._raise(s)

    def force_batch_device(self, offsets, d_pos, d_Z, d_boxes, d_F, d_E, stream=None):
        """Enqueue only; call ``sync_status`` to collect the status."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        s = self._lib.rgpot_b200_force_batch_device(
            self._h, len(offsets) - 1, offsets.ctypes.data, d_pos.data_ptr(),
            d_Z.data_ptr() if d_Z is not None else None, d_boxes.data_ptr(), d_F.data_ptr(),
            d_E.data_ptr(), stream)
        if s != 0:
            self._raise(s)

    def sync_status(self) -> int:
        st = int(self._lib.rgpot_b200_sync_status(self._h))
        self._inflight = None
        return st

    def force_torch(self, positions, atom_types, box, sync: bool = True):
        """torch interop: `positions` a CUDA float64 (n, 3) tensor, `atom_types` a
        CUDA int32 (n,) tensor or None; nothing crosses PCIe.  Runs on torch's current stream of that
        device and returns (energy, forces): forces a new CUDA tensor; energy a float when `sync`, else
        a one-element CUDA tensor (collect the status later with ``sync_status``)."""
        import torch

        if not (positions.is_cuda and positions.dtype == torch.float64 and positions.dim() == 2 and
                positions.shape[1] == 3):
            raise ValueError("positions must be a CUDA float64 tensor of shape (n_atoms, 3)")
        positions = positions.contiguous()
        if atom_types is not None:
            if not (atom_types.is_cuda and atom_types.shape == (positions.shape[0],)):
                raise ValueError("atom_types must be a CUDA tensor of shape (n_atoms,)")
            atom_types = atom_types.to(torch.int32).contiguous()
        forces = torch.empty_like(positions)
        energy = torch.empty(1, dtype=torch.float64, device=positions.device)
        stream = torch.cuda.current_stream(positions.device).cuda_stream
        self.force_device(positions, atom_types, box, forces, energy, stream, sync=sync)
        # the engine may re-read its inputs in sync_status (include/rgpot_b200.h, *_device lifetime
        # rule): the converted tensors must outlive this call when it only enqueues
        self._inflight = None if sync else (positions, atom_types, forces, energy)
        self.force_calls += 1
        return (float(energy.item()) if sync else energy), forces

    # -- result cache (Potential.hpp:251-336) ------------------------------------------------------
    def cacheKeys(self, systems) -> np.ndarray:
        """uint64 cache keys of independent systems [(positions, atom_types, box), ...], each equal to
        the reference's ``Potential::cacheKey`` (XXH3 of geometry, species and cell, XOR the hashes of
        the potential type and ``paramsKey()``): one launch for the whole batch."""
        n = len(systems)
        if n == 0:
            return np.zeros(0, dtype=np.uint64)
        pos = [_as_positions(s[0]) for s in systems]
        typ = [_as_types(s[1], p.shape[0]) for s, p in zip(systems, pos)]
        offsets = np.zeros(n + 1, dtype=np.uint64)
        offsets[1:] = np.cumsum([p.shape[0] for p in pos])
        P = np.ascontiguousarray(np.concatenate(pos)) if offsets[-1] else np.zeros((1, 3))
        T = np.ascontiguousarray(np.concatenate(typ)) if offsets[-1] else np.zeros(1, np.int32)
        B = np.ascontiguousarray(np.stack([_as_box(s[2]) for s in systems]))
        keys = np.zeros(n, dtype=np.uint64)
        st = self._lib.rgpot_b200_cache_keys(self._h, n, offsets.ctypes.data, P.ctypes.data, T.ctypes.data,
                                             B.ctypes.data, int(self.get_type().value), int(self.paramsKey()),
                                             keys.ctypes.data)
        if st != 0:
            self._raise(st)
        return keys

    def cache_keys_device(self, offsets, d_pos, d_Z, d_boxes, d_keys, stream=None):
        """Device-resident form (torch CUDA tensors; `d_keys` int64/uint64 of nSystems): enqueue only."""
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        s = self._lib.rgpot_b200_cache_keys_device(
            self._h, len(offsets) - 1, offsets.ctypes.data, d_pos.data_ptr(), d_Z.data_ptr(), d_boxes.data_ptr(),
            int(self.get_type().value), int(self.paramsKey()), d_keys.data_ptr(), stream)
        if s != 0:
            self._raise(s)

    def forceBatchCached(self, systems, cache: dict):
        """``Potential::forceBatch`` with a result cache (Potential.hpp:251-306): every system is keyed
        on its own, hits are served from `cache` (a mapping key -> (energy, forces)), only the misses
        reach the kernel as one compact batch, and their results are written back.  Returns
        ([(energy, forces, variance)], number of misses)."""
        keys = self.cacheKeys(systems)
        out = [None] * len(systems)
        misses = []
        for i, k in enumerate(keys.tolist()):
            hit = cache.get(k)
            if hit is not None:
                out[i] = (hit[0], hit[1].copy(), 0.0)
            else:
                misses.append(i)
        if misses:
            res = self.forceBatch([systems[i] for i in misses])
            for i, r in zip(misses, res):
                out[i] = r
                cache[int(keys[i])] = (r[0], r[1].copy())
        return out, len(misses)

    # -- neighbor table --------------------------------------------------------------------------
    def neighbors(self, positions, atom_types, box, path: int = 0):
        """(pairs (m, 2), shifts (m, 3)) of the table the kernels act on, unspecified order."""
        pos = _as_positions(positions)
        cell = _as_box(box)
        types = None if atom_types is None else _as_types(atom_types, pos.shape[0])
        n_out = C.c_long(0)
        cap = max(1024, 160 * pos.shape[0])
        while True:
            pairs = np.zeros((cap, 2), dtype=np.int32)
            shifts = np.zeros((cap, 3), dtype=np.int32)
            st = self._lib.rgpot_b200_neighbors(
                self._h, pos.shape[0], pos.ctypes.data, types.ctypes.data if types is not None else None,
                cell.ctypes.data, int(path), cap, pairs.ctypes.data, shifts.ctypes.data, C.byref(n_out))
            if st != 0:
                self._raise(st)
            if n_out.value <= cap:
                return pairs[: n_out.value], shifts[: n_out.value]
            cap = int(n_out.value)


class LJPot(_EnginePot):
    """Shifted 12-6 Lennard-Jones, periodic, orthorhombic nearest image (LJPot.hpp:40-86)."""

    _kind, _label, _type = _lib.KIND_LJ, "LJ", PotType.LJ

    def __init__(self, config: LJConfig | None = None, device: int = -1, all_images: bool = False,
                 devices=None, host_threads: int = 0):
        c = config or LJConfig()
        self._config = c
        super().__init__(c.u0, c.cutoff, c.psi, device, all_images, devices=devices, host_threads=host_threads)

    def config(self) -> LJConfig:
        return self._config


class LJClusterPot(_EnginePot):
    """Free-boundary shifted 12-6 Lennard-Jones (LJClusterPot.hpp:47-100)."""

    _kind, _label, _type = _lib.KIND_LJ_CLUSTER, "LJCluster", PotType.LJCluster

    def __init__(self, config: LJClusterConfig | None = None, device: int = -1):
        c = config or LJClusterConfig()
        self._config = c
        super().__init__(c.u0, c.cutoff, c.psi, device)

    def config(self) -> LJClusterConfig:
        return self._config


class MorsePot(_EnginePot):
    """Shifted Morse pair potential, periodic orthorhombic nearest image (Morse/MorsePot.hpp:44-101)."""

    _kind, _label, _type = _lib.KIND_MORSE, "Morse", PotType.Morse

    def __init__(self, config: MorseConfig | None = None, device: int = -1):
        c = config or MorseConfig()
        self._config = c
        super().__init__(device=device, morse=c)

    def config(self) -> MorseConfig:
        return self._config

    def energyShift(self) -> float:
        """MorsePot.hpp:72-74: the unshifted pair energy at the cutoff."""
        import math
        c = self._config
        d = 1.0 - math.exp(-c.a * (c.cutoff - c.re))
        return c.De * d * d - c.De


class ZBLPot(_EnginePot):
    """Ziegler-Biersack-Littmark screened nuclear repulsion with switching, free boundaries
    (ZBL/ZBLPot.hpp:60-130)."""

    _kind, _label, _type = _lib.KIND_ZBL, "ZBL", PotType.ZBL

    def __init__(self, config: ZBLConfig | None = None, device: int = -1):
        c = config or ZBLConfig()
        self._config = c
        super().__init__(device=device, zbl=c)

    def config(self) -> ZBLConfig:
        return self._config


class CuH2Pot(_EnginePot):
    """Cu-H embedded-atom potential (fortranpots::CuH2Pot, FortranPots.hpp:29-45)."""

    _kind, _label, _type = _lib.KIND_CUH2, "CuH2", PotType.CuH2

    def __init__(self, device: int = -1, devices=None, host_threads: int = 0):
        super().__init__(device=device, devices=devices, host_threads=host_threads)


class EAMAlPot(_EnginePot):
    """Double-exponential embedded-atom aluminium (fortranpots::EAMAlPot, FortranPots.hpp:43,
    rgpot_eam_al.f90).  Single species: atomic numbers are accepted and not read."""

    _kind, _label, _type = _lib.KIND_EAM_AL, "EAM-Al", PotType.EAMAl

    def __init__(self, device: int = -1):
        super().__init__(device=device)


class FeHePot(_EnginePot):
    """Fe-He embedded-atom potential (fortranpots::FeHePot, FortranPots.hpp:44, rgpot_fehe.f90):
    atomic numbers 26 and 2, anything else raises (status 2)."""

    _kind, _label, _type = _lib.KIND_FEHE, "FeHe", PotType.FeHe

    def __init__(self, device: int = -1):
        super().__init__(device=device)


# -------------------------------------------------------------------- module-level functions
_DEFAULT = {}


def _default(cls):
    if cls not in _DEFAULT:
        _DEFAULT[cls] = cls()
    return _DEFAULT[cls]


def evaluate_lj(positions, atom_types, box):
    """python/bind/bind_module.cpp:84-96 -- (energy, forces (n, 3), variance), default LJConfig."""
    pos = np.asarray(positions)
    typ = np.asarray(atom_types)
    cell = np.asarray(box)
    if pos.ndim != 2 or typ.ndim != 1 or cell.ndim != 2:
        raise ValueError("evaluate_lj expects positions (n,3), atom_types (n,), box (3,3)")
    return _default(LJPot)(pos, typ, cell)


def evaluate_cuh2(positions, atom_types, box):
    """Same call shape for the CuH2 EAM (the reference module exposes LJ only)."""
    return _default(CuH2Pot)(positions, atom_types, box)


_BY_NAME = {"LJ": LJPot, "LJCluster": LJClusterPot, "Morse": MorsePot, "ZBL": ZBLPot, "CuH2": CuH2Pot,
            "EAMAl": EAMAlPot, "FeHe": FeHePot}  # the server's selector strings (server.cpp:208-220) and kin


def evaluate_batch(systems, potential="CuH2"):
    """[(energy, forces, variance)] for independent systems [(positions, atom_types, box), ...] in ONE
    engine call (ForceBatch semantics, ForceStructs.hpp:39-54; the reference module has no batch
    function).  `potential`: a selector string or a potential object."""
    pot = _default(_BY_NAME[potential]) if isinstance(potential, str) else potential
    return pot.forceBatch(systems)


def neighbor_list(points, box, cutoff: float, full: bool = True, sorted: bool = True, periodic: bool = True,
                  pot=None) -> dict:
    """vesin-shaped neighbor list from the GPU cell list (rgpot_b200_neighbor_list): dict of
    `pairs` (m, 2) int64, `shifts` (m, 3) int32, `distances` (m,), `vectors` (m, 3).  Pair set,
    vectors and distances are bit-identical to vesin's (third_party/vesin) for the same options."""
    pot = pot or _default(LJPot)
    pos = _as_positions(points)
    cell = _as_box(box) if box is not None else np.eye(3)
    per = (C.c_int * 3)(*([1, 1, 1] if periodic else [0, 0, 0]))
    nl = _lib.NeighborList()
    st = pot._lib.rgpot_b200_neighbor_list(pot._h, pos.shape[0], pos.ctypes.data, cell.ctypes.data, per,
                                           float(cutoff), int(full), int(sorted), C.byref(nl))
    if st != 0:
        pot._raise(st)
    m = nl.length
    try:
        if m == 0:
            return {"pairs": np.zeros((0, 2), np.int64), "shifts": np.zeros((0, 3), np.int32),
                    "distances": np.zeros(0), "vectors": np.zeros((0, 3))}
        return {
            "pairs": np.ctypeslib.as_array(nl.pairs, shape=(m, 2)).astype(np.int64),
            "shifts": np.ctypeslib.as_array(nl.shifts, shape=(m, 3)).copy(),
            "distances": np.ctypeslib.as_array(nl.distances, shape=(m,)).copy(),
            "vectors": np.ctypeslib.as_array(nl.vectors, shape=(m, 3)).copy(),
        }
    finally:
        pot._lib.rgpot_b200_neighbor_list_free(C.byref(nl))


def pinned_empty(shape, dtype=np.float64) -> np.ndarray:
    """numpy array on page-locked memory from rgpot_b200_host_alloc (kept alive by the array)."""
    lib = _lib.lib()
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    nbytes = max(count * dtype.itemsize, 1)
    ptr = lib.rgpot_b200_host_alloc(nbytes)
    if not ptr:
        raise MemoryError("rgpot_b200_host_alloc failed")

    class _Owner:
        def __init__(self, p):
            self.p = p

        def __del__(self):
            lib.rgpot_b200_host_free(self.p)

    buf = (C.c_char * nbytes).from_address(ptr)
    buf._owner = _Owner(ptr)
    arr = np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)
    return arr


def measure_fp64_tflops(device: int = -1, iters: int = 20000) -> float:
    return float(_lib.lib().rgpot_b200_measure_fp64_tflops(device, iters))


def measure_copy_gbs(devices, h2d_bytes: int, d2h_bytes: int, reps: int = 3) -> float:
    """Aggregate pinned host<->device copy rate of `devices` (GB/s), both directions and all devices at
    once (rgpot_b200_measure_copy_gbs): what bounds every end-to-end batch number on this box."""
    devs = (C.c_int * len(devices))(*[int(d) for d in devices])
    return float(_lib.lib().rgpot_b200_measure_copy_gbs(devs, len(devices), int(h2d_bytes), int(d2h_bytes), int(reps)))


def measure_host_copy_gbs(threads: int, nbytes: int, reps: int = 3) -> float:
    """Payload GB/s of `threads` host threads packing `nbytes` into page-locked memory and unpacking them
    again (rgpot_b200_measure_host_copy_gbs): the host-side ceiling of the pageable ForceBatch entry."""
    return float(_lib.lib().rgpot_b200_measure_host_copy_gbs(int(threads), int(nbytes), int(reps)))
