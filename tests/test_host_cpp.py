"""The C++ plugin classes (rgpot_b200/host) and the reference's own CuH2Pot over the engine."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "rgpot_b200", "host")


def test_host_classes_build_standalone_and_against_reference_headers():
    """Compiles cuda_pots.cc against the stand-in interface AND (where the checkout exists) the
    reference's real Potential.hpp: the classes are source-compatible plugin subclasses."""
    from rgpot_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    subprocess.run(["make", "-s", "-C", HOST], check=True)
    out = subprocess.run(["make", "-s", "-C", HOST, "refcheck"], check=True, capture_output=True, text=True).stdout
    assert "refcheck:" in out
    if os.path.exists("/root/reference/CppCore/rgpot/Potential.hpp"):
        assert "compile against" in out


@pytest.mark.gpu
def test_cpp_plugin_classes_reproduce_reference_tests():
    """CuH2PotTest.cc / LJClusterPotTest.cc through operator() and forceBatch of the C++ classes."""
    exe = os.path.join(HOST, "test_host")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", HOST], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "test_host: ok" in res.stdout


@pytest.mark.gpu
def test_core_callback_from_plain_c():
    """include/rgpot_b200_core.h compiled as C11: the rgpot-core PotentialCallback over DLPack
    tensors reproduces CuH2PotTest.cc:11-31 and rejects a malformed tensor."""
    exe = os.path.join(HOST, "test_core")
    if not os.path.exists(exe):
        subprocess.run(["make", "-s", "-C", HOST], check=True)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "test_core: ok" in res.stdout


@pytest.mark.gpu
def test_unmodified_reference_cuh2pot_runs_on_the_engine():
    """oracle/_ref/dropin_cuh2pot = the reference's FortranPots.cc linked against librgpot_b200.so
    (built only where /root/reference exists; travels to the GPU box prebuilt)."""
    exe = os.path.join(ROOT, "oracle", "_ref", "dropin_cuh2pot")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/dropin_cuh2pot not built (needs /root/reference)")
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "ok" in res.stdout
