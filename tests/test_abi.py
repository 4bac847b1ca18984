"""The C-ABI library loads and exports every symbol include/*.h declares (no compute)."""
import ctypes as C
import os
import re

import pytest

from rgpot_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def engine():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def declared_symbols():
    names = set()
    inc = os.path.join(ROOT, "include")
    for header in sorted(os.listdir(inc)):
        if not header.endswith(".h"):
            continue
        text = open(os.path.join(inc, header)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names.update(re.findall(r"RGPOT_B200_API\s+[^;(]*?\b(rgpot_\w+)\s*\(", text))
    return sorted(names)


def test_header_symbols_are_exported(engine):
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(engine, name), f"{name} declared in include/*.h but not exported"
    assert set(names) == {s[0] for s in _lib.SYMBOLS}, "python binding table out of sync with the header"


def test_abi_version_and_defaults(engine):
    assert engine.rgpot_b200_abi_version() == 3
    assert C.sizeof(_lib.Config) == 160  # RgpotB200Config incl. the device list (ABI 3)
    cfg = _lib.Config()
    engine.rgpot_b200_default_config(_lib.KIND_LJ, cfg)
    assert (cfg.u0, cfg.cutoff, cfg.psi, cfg.device) == (1.0, 15.0, 1.0, -1)  # LJPot.hpp:30-34
    engine.rgpot_b200_default_config(_lib.KIND_CUH2, cfg)
    assert cfg.cutoff == 6.1  # rgpot_cuh2.f90:100
    engine.rgpot_b200_default_config(_lib.KIND_MORSE, cfg)
    assert (cfg.morse_De, cfg.morse_a, cfg.morse_re, cfg.cutoff) == (0.7102, 1.6047, 2.8970, 9.5)  # MorsePot.hpp:31-36
    engine.rgpot_b200_default_config(_lib.KIND_ZBL, cfg)
    assert (cfg.zbl_cut_inner, cfg.cutoff) == (2.0, 2.5)  # ZBLPot.hpp:29-32


def test_core_callback_rejects_bad_input_without_a_gpu(engine):
    """rgpot_b200_core_callback (the rgpot-core PotentialCallback) validates before it computes."""
    import ctypes as C
    engine.rgpot_b200_core_last_error.restype = C.c_char_p
    out = _lib.ForceOut()
    assert engine.rgpot_b200_core_callback(None, None, C.byref(out)) == 1  # RGPOT_INVALID_PARAMETER
    assert b"NULL" in engine.rgpot_b200_core_last_error()


def test_fortran_dropin_symbols_present(engine):
    # the exact names FortranPots.cc:33-40 binds
    assert hasattr(engine, "rgpot_cuh2_force") and hasattr(engine, "rgpot_fortran_last_error")


def test_no_cpu_fallback_without_gpu(engine):
    """Without a usable device the engine refuses to create a handle (it must not compute)."""
    if engine.rgpot_b200_available() > 0:
        pytest.skip("a GPU is present")
    import ctypes as C
    cfg = _lib.Config()
    engine.rgpot_b200_default_config(_lib.KIND_LJ, cfg)
    err = C.create_string_buffer(256)
    assert not engine.rgpot_b200_create(cfg, err, 256)
    assert b"no CPU fallback" in err.value
    import rgpot_b200
    with pytest.raises(RuntimeError):
        rgpot_b200.LJPot()


def test_engine_sources_do_not_touch_the_oracle():
    """Product code must never import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "rgpot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cc", "Makefile")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "pyoracle" not in text and "liboracle" not in text and "oracle/" not in text, f


def test_host_hash_matches_xxhash_vectors():
    """rgpot_b200_hash64 -- the host build of the source the kernel compiles for its short inputs
    (hash.cuh) -- against the committed libxxhash digests, every length class.  No GPU involved."""
    import json

    from rgpot_b200 import _lib

    lib = _lib.lib()
    with open(os.path.join(ROOT, "tests", "golden", "xxh3_vectors.json")) as fh:
        g = json.load(fh)

    def pattern(n):
        return bytes(((i * 131 + n * 31 + (i >> 8) * 7) & 0xFF) for i in range(n))

    for n, digest in g["digests"]:
        assert lib.rgpot_b200_hash64(pattern(n), n) == int(digest), n
