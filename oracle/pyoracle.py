"""ctypes doorway to the CPU oracle.  TEST INFRASTRUCTURE ONLY (see oracle/oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module; nothing under rgpot_b200/ does.

Two libraries:
  * oracle/liboracle.so      -- the C99 restatement (always buildable: `make -C oracle`)
  * oracle/_ref/libref.so    -- the reference's own LJPot / LJClusterPot / vesin objects,
                                compiled in place from /root/reference (`make -C oracle ref`);
                                travels to the GPU box prebuilt, may be absent.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


class NList(C.Structure):
    _fields_ = [
        ("n_pairs", C.c_long),
        ("row", C.POINTER(C.c_int32)),
        ("idx", C.POINTER(C.c_int32)),
        ("shift", C.POINTER(C.c_int32)),
        ("vec", C.POINTER(C.c_double)),
        ("dist", C.POINTER(C.c_double)),
        ("n_self_images", C.c_long),
    ]


def build(ref: bool = True) -> None:
    """Compile the checker (building the checker is not using it)."""
    subprocess.run(["make", "-s", "-C", HERE], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", HERE, "ref"], check=True)


def _declare_common(lib):
    lib.oracle_lj_force.restype = C.c_long
    lib.oracle_lj_force.argtypes = [C.c_long, _f64p, _f64p, C.c_double, C.c_double, C.c_double,
                                    C.c_int, _f64p, C.POINTER(C.c_double)]
    lib.oracle_morse_force.restype = C.c_long
    lib.oracle_morse_force.argtypes = [C.c_long, _f64p, _f64p, C.c_double, C.c_double, C.c_double,
                                       C.c_double, _f64p, C.POINTER(C.c_double)]
    lib.oracle_zbl_force.restype = C.c_long
    lib.oracle_zbl_force.argtypes = [C.c_long, _f64p, _i32p, C.c_double, C.c_double, _f64p,
                                     C.POINTER(C.c_double)]
    lib.oracle_lj_pairs.restype = C.c_long
    lib.oracle_lj_pairs.argtypes = [C.c_long, _f64p, _f64p, C.c_double, C.c_int, _i32p, C.c_long]
    lib.oracle_nlist_build.restype = C.c_int
    lib.oracle_nlist_build.argtypes = [C.c_long, _f64p, _f64p, C.c_double, C.c_double,
                                       C.POINTER(NList), C.c_char_p, C.c_size_t]
    lib.oracle_nlist_free.restype = None
    lib.oracle_nlist_free.argtypes = [C.POINTER(NList)]
    lib.oracle_eam_al_force.restype = C.c_int
    lib.oracle_eam_al_force.argtypes = [C.c_int32, _f64p, _f64p, _f64p, C.POINTER(C.c_double), C.c_char_p,
                                        C.c_size_t]
    lib.oracle_eam_al_scalar.restype = C.c_double
    lib.oracle_eam_al_scalar.argtypes = [C.c_int, C.c_double]
    lib.oracle_fehe_force.restype = C.c_int
    lib.oracle_fehe_force.argtypes = [C.c_int32, _f64p, _i32p, _f64p, _f64p, C.POINTER(C.c_double), C.c_char_p,
                                      C.c_size_t]
    lib.oracle_xxh3_64.restype = C.c_uint64
    lib.oracle_xxh3_64.argtypes = [C.c_char_p, C.c_size_t]
    lib.oracle_cache_key.restype = C.c_uint64
    lib.oracle_cache_key.argtypes = [C.c_size_t, _f64p, _i32p, _f64p, C.c_size_t, C.c_uint64]
    lib.oracle_fehe_scalar.restype = C.c_double
    lib.oracle_fehe_scalar.argtypes = [C.c_int, C.c_double]
    lib.oracle_cuh2_force.restype = C.c_int
    lib.oracle_cuh2_force.argtypes = [C.c_int32, _f64p, _i32p, _f64p, _f64p, C.POINTER(C.c_double),
                                      C.c_char_p, C.c_size_t, C.POINTER(NList)]
    for name in ("phi", "dphi", "rho", "drho", "embed", "dembed"):
        fn = getattr(lib, "oracle_cuh2_" + name)
        fn.restype = C.c_double
        fn.argtypes = [C.c_int, C.c_double]
    lib.oracle_lj_force_allimages.restype = C.c_int
    lib.oracle_lj_force_allimages.argtypes = [C.c_long, _f64p, _f64p, C.c_double, C.c_double,
                                              C.c_double, _f64p, C.POINTER(C.c_double),
                                              C.c_char_p, C.c_size_t]


_ORACLE = None
_REF = None


def oracle_lib():
    global _ORACLE
    if _ORACLE is None:
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        _ORACLE = C.CDLL(path)
        _declare_common(_ORACLE)
    return _ORACLE


def ref_available() -> bool:
    return os.path.exists(os.path.join(HERE, "_ref", "libref.so"))


def ref_lib():
    global _REF
    if _REF is None:
        path = os.path.join(HERE, "_ref", "libref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (run `make -C oracle ref` where /root/reference exists)")
        _REF = C.CDLL(path)
        _declare_common(_REF)
        _REF.ref_lj_force.restype = C.c_int
        _REF.ref_lj_force.argtypes = [C.c_long, _f64p, _i32p, _f64p, C.c_double, C.c_double,
                                      C.c_double, C.c_int, _f64p, C.POINTER(C.c_double)]
        _REF.ref_morse_force.restype = C.c_int
        _REF.ref_morse_force.argtypes = [C.c_long, _f64p, _i32p, _f64p, C.c_double, C.c_double,
                                         C.c_double, C.c_double, _f64p, C.POINTER(C.c_double)]
        _REF.ref_zbl_force.restype = C.c_int
        _REF.ref_zbl_force.argtypes = [C.c_long, _f64p, _i32p, _f64p, C.c_double, C.c_double, _f64p,
                                       C.POINTER(C.c_double)]
        _REF.ref_lj_params_key.restype = C.c_uint64
        _REF.ref_lj_params_key.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int]
        _REF.ref_table_new.restype = C.c_void_p
        _REF.ref_table_delete.argtypes = [C.c_void_p]
        _REF.ref_vesin_table.restype = C.c_int
        _REF.ref_vesin_table.argtypes = [C.c_void_p, C.c_long, _f64p, _f64p, C.c_double, C.c_double,
                                         C.c_int, C.POINTER(NList), C.c_char_p, C.c_size_t]
        _REF.ref_cuh2_force.restype = C.c_int
        _REF.ref_cuh2_force.argtypes = [C.c_void_p, C.c_int32, _f64p, _i32p, _f64p, _f64p,
                                        C.POINTER(C.c_double), C.c_char_p, C.c_size_t]
    return _REF


def _prep(pos, box):
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(9)
    return pos, box


# ------------------------------------------------------------------ LJ
def lj_force(pos, box, u0=1.0, cutoff=15.0, psi=1.0, cluster=False):
    pos, box = _prep(pos, box)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    npairs = oracle_lib().oracle_lj_force(len(pos), pos, box, u0, cutoff, psi, int(cluster), F,
                                          C.byref(e))
    return e.value, F, npairs


def morse_force(pos, box, De=0.7102, a=1.6047, re=2.8970, cutoff=9.5):
    pos, box = _prep(pos, box)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    n = oracle_lib().oracle_morse_force(len(pos), pos, box, De, a, re, cutoff, F, C.byref(e))
    return e.value, F, n


def ref_morse_force(pos, box, De=0.7102, a=1.6047, re=2.8970, cutoff=9.5):
    pos, box = _prep(pos, box)
    Z = np.full(len(pos), 78, dtype=np.int32)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    if ref_lib().ref_morse_force(len(pos), pos, Z, box, De, a, re, cutoff, F, C.byref(e)) != 0:
        raise RuntimeError("reference Morse threw")
    return e.value, F


def zbl_force(pos, Z, cut_inner=2.0, cut_global=2.5):
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    n = oracle_lib().oracle_zbl_force(len(pos), pos, Z, cut_inner, cut_global, F, C.byref(e))
    return e.value, F, n


def ref_zbl_force(pos, Z, box, cut_inner=2.0, cut_global=2.5):
    pos, box = _prep(pos, box)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    if ref_lib().ref_zbl_force(len(pos), pos, Z, box, cut_inner, cut_global, F, C.byref(e)) != 0:
        raise RuntimeError("reference ZBL threw")
    return e.value, F


def lj_pairs(pos, box, cutoff, cluster=False):
    pos, box = _prep(pos, box)
    lib = oracle_lib()
    dummy = np.zeros(2, dtype=np.int32)
    n = lib.oracle_lj_pairs(len(pos), pos, box, cutoff, int(cluster), dummy, 0)
    out = np.zeros(2 * max(n, 1), dtype=np.int32)
    lib.oracle_lj_pairs(len(pos), pos, box, cutoff, int(cluster), out, n)
    return out[: 2 * n].reshape(-1, 2)


def ref_lj_force(pos, box, u0=1.0, cutoff=15.0, psi=1.0, cluster=False, Z=None):
    pos, box = _prep(pos, box)
    Z = np.ones(len(pos), dtype=np.int32) if Z is None else np.ascontiguousarray(Z, dtype=np.int32)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    st = ref_lib().ref_lj_force(len(pos), pos, Z, box, u0, cutoff, psi, int(cluster), F, C.byref(e))
    if st != 0:
        raise RuntimeError("reference LJ threw")
    return e.value, F


def lj_force_allimages(pos, box, u0=1.0, cutoff=2.5, psi=1.0):
    pos, box = _prep(pos, box)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    err = C.create_string_buffer(512)
    st = oracle_lib().oracle_lj_force_allimages(len(pos), pos, box, u0, cutoff, psi, F,
                                                C.byref(e), err, 512)
    if st != 0:
        raise RuntimeError(err.value.decode())
    return e.value, F


# ------------------------------------------------------------------ neighbor tables
def _table_to_numpy(t: NList, n: int):
    m = t.n_pairs
    row = np.ctypeslib.as_array(t.row, shape=(n + 1,)).copy()
    if m > 0:
        idx = np.ctypeslib.as_array(t.idx, shape=(m,)).copy()
        shift = np.ctypeslib.as_array(t.shift, shape=(m, 3)).copy()
        vec = np.ctypeslib.as_array(t.vec, shape=(m, 3)).copy()
        dist = np.ctypeslib.as_array(t.dist, shape=(m,)).copy()
    else:
        idx = np.zeros(0, np.int32)
        shift = np.zeros((0, 3), np.int32)
        vec = np.zeros((0, 3))
        dist = np.zeros(0)
    return {"row": row, "idx": idx, "shift": shift, "vec": vec, "dist": dist,
            "n_self_images": int(t.n_self_images)}


def nlist(pos, box, cutoff=6.1, skin=None, use_ref=False, n_threads=1):
    """Full neighbor table (CSR).  use_ref=True asks the vendored vesin itself."""
    pos, box = _prep(pos, box)
    skin = 0.1 * cutoff if skin is None else skin
    t = NList()
    err = C.create_string_buffer(512)
    if use_ref:
        lib = ref_lib()
        st = lib.ref_vesin_table(None, len(pos), pos, box, cutoff, skin, n_threads, C.byref(t), err, 512)
    else:
        lib = oracle_lib()
        st = lib.oracle_nlist_build(len(pos), pos, box, cutoff, skin, C.byref(t), err, 512)
    if st != 0:
        raise RuntimeError(err.value.decode())
    out = _table_to_numpy(t, len(pos))
    lib.oracle_nlist_free(C.byref(t))
    return out


class RawList(C.Structure):
    _fields_ = [("length", C.c_long), ("pairs", C.POINTER(C.c_longlong)), ("shifts", C.POINTER(C.c_int32)),
                ("dists", C.POINTER(C.c_double)), ("vecs", C.POINTER(C.c_double))]


def vesin_raw(pos, box, cutoff, full=True, periodic=True):
    """The vendored vesin's own list (oracle/_ref only): dict of pairs (m,2), shifts (m,3), distances (m,),
    vectors (m,3), sorted by (i, j, shift).  Self images are kept, like vesin returns them."""
    pos, box = _prep(pos, box)
    lib = ref_lib()
    lib.ref_vesin_raw.restype = C.c_int
    lib.ref_vesin_raw.argtypes = [C.c_long, _f64p, _f64p, C.c_int, C.c_double, C.c_int, C.POINTER(RawList),
                                  C.c_char_p, C.c_size_t]
    lib.ref_vesin_raw_free.restype = None
    lib.ref_vesin_raw_free.argtypes = [C.POINTER(RawList)]
    r = RawList()
    err = C.create_string_buffer(512)
    st = lib.ref_vesin_raw(len(pos), pos, box, int(periodic), float(cutoff), int(full), C.byref(r), err, 512)
    if st != 0:
        raise RuntimeError(err.value.decode())
    m = r.length
    out = {
        "pairs": np.ctypeslib.as_array(r.pairs, shape=(max(m, 1), 2))[:m].astype(np.int64),
        "shifts": np.ctypeslib.as_array(r.shifts, shape=(max(m, 1), 3))[:m].copy(),
        "distances": np.ctypeslib.as_array(r.dists, shape=(max(m, 1),))[:m].copy(),
        "vectors": np.ctypeslib.as_array(r.vecs, shape=(max(m, 1), 3))[:m].copy(),
    }
    lib.ref_vesin_raw_free(C.byref(r))
    return out


def lj_vesin_force(pos, box, u0=1.0, cutoff=2.5, psi=1.0, n_threads=0):
    """(energy, forces, n_pairs) of the linear-scaling LJ comparator (oracle/_ref only): vendored vesin half
    list + the pair lambda of LJPot.cc:61-81.  Not the reference algorithm (that is the O(N^2) scan)."""
    pos, box = _prep(pos, box)
    lib = ref_lib()
    lib.ref_lj_vesin_force.restype = C.c_int
    lib.ref_lj_vesin_force.argtypes = [C.c_long, _f64p, _f64p, C.c_double, C.c_double, C.c_double, C.c_int, _f64p,
                                       C.POINTER(C.c_double), C.POINTER(C.c_long), C.c_char_p, C.c_size_t]
    F = np.zeros_like(pos)
    e, m = C.c_double(0.0), C.c_long(0)
    err = C.create_string_buffer(512)
    st = lib.ref_lj_vesin_force(len(pos), pos, box, u0, cutoff, psi, int(n_threads), F, C.byref(e), C.byref(m), err, 512)
    if st != 0:
        raise RuntimeError(err.value.decode())
    return e.value, F, m.value


def canonical_pairs(table) -> np.ndarray:
    """(i, j, s0, s1, s2) rows in lexicographic order: the pair SET of a table."""
    row = table["row"]
    n = len(row) - 1
    i = np.repeat(np.arange(n, dtype=np.int64), np.diff(row))
    rec = np.column_stack([i, table["idx"].astype(np.int64), table["shift"].astype(np.int64)])
    if len(rec) == 0:
        return rec.reshape(0, 5)
    order = np.lexsort((rec[:, 4], rec[:, 3], rec[:, 2], rec[:, 1], rec[:, 0]))
    return rec[order]


# ------------------------------------------------------------------ CuH2
class CuH2Error(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"CuH2 potential failed (status {status}): {message}")
        self.status = status
        self.message = message


def cuh2_force(pos, Z, box, use_ref=False, return_table=False):
    pos, box = _prep(pos, box)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    err = C.create_string_buffer(512)
    if use_ref:
        st = ref_lib().ref_cuh2_force(None, len(pos), pos, Z, box, F, C.byref(e), err, 512)
        table = None
    else:
        t = NList()
        st = oracle_lib().oracle_cuh2_force(len(pos), pos, Z, box, F, C.byref(e), err, 512,
                                            C.byref(t) if return_table else None)
        table = None
        if return_table and st == 0:
            table = _table_to_numpy(t, len(pos))
            oracle_lib().oracle_nlist_free(C.byref(t))
    if st != 0:
        raise CuH2Error(st, err.value.decode())
    if return_table:
        return e.value, F, table
    return e.value, F


# ------------------------------------------------------------------ EAM-Al
def eam_al_force(pos, box):
    """(energy, forces) of rgpot_eam_al_force (rgpot_eam_al_capi.f90:31-58); atomic numbers are not an input."""
    pos, box = _prep(pos, box)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    err = C.create_string_buffer(512)
    st = oracle_lib().oracle_eam_al_force(len(pos), pos, box, F, C.byref(e), err, 512)
    if st != 0:
        raise RuntimeError(f"EAM-Al potential failed (status {st}): {err.value.decode()}")
    return e.value, F


def eam_al_scalar(which, x):
    """0 pair_energy, 1 pair_deriv, 2 density, 3 density_deriv (r); 4 embedding, 5 embedding_deriv (rho)."""
    return oracle_lib().oracle_eam_al_scalar(int(which), float(x))


# ------------------------------------------------------------------ result-cache key
def xxh3_64(data: bytes) -> int:
    """XXH3_64bits (xxHash 0.8.3, seed 0) -- oracle/xxh3_oracle.c."""
    return int(oracle_lib().oracle_xxh3_64(data, len(data)))


def cache_key(pos, Z, box, pot_type: int, params_key: int) -> int:
    """Potential<D>::cacheKey (Potential.hpp:324-336)."""
    pos = np.ascontiguousarray(pos, dtype=np.float64).reshape(-1, 3)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    box = np.ascontiguousarray(box, dtype=np.float64).reshape(3, 3)
    return int(oracle_lib().oracle_cache_key(len(pos), pos, Z, box, int(pot_type), int(params_key)))


# ------------------------------------------------------------------ FeHe
class FeHeError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"FeHe potential failed (status {status}): {message}")
        self.status = status
        self.message = message


def fehe_force(pos, Z, box):
    """(energy, forces) of rgpot_fehe_force (rgpot_fehe_capi.f90:36-63)."""
    pos, box = _prep(pos, box)
    Z = np.ascontiguousarray(Z, dtype=np.int32)
    F = np.zeros_like(pos)
    e = C.c_double(0.0)
    err = C.create_string_buffer(512)
    st = oracle_lib().oracle_fehe_force(len(pos), pos, Z, box, F, C.byref(e), err, 512)
    if st != 0:
        raise FeHeError(st, err.value.decode())
    return e.value, F


def fehe_scalar(which, x):
    return oracle_lib().oracle_fehe_scalar(int(which), float(x))


def cuh2_scalar(name, kind, x):
    return getattr(oracle_lib(), "oracle_cuh2_" + name)(int(kind), float(x))
