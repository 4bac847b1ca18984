"""`rgpot_b200._core`: the compiled (pybind11) module with the reference's nanobind surface
(python/bind/bind_module.cpp:142-174).  The GPU tests re-run the assertions of the reference's own
tests/test_rgpot_pip_core.py:20-56 against it."""
import math

import numpy as np
import pytest


def test_module_imports_and_exports_the_reference_names():
    from rgpot_b200 import _core
    assert _core.__version__
    assert callable(_core.evaluate_lj) and _core.LJPot is not None
    assert _core.has_metatomic_dlopen is False  # the Metatomic dlopen frontend is out of scope here
    assert callable(_core.evaluate_cuh2) and _core.CuH2Pot is not None


@pytest.mark.gpu
def test_evaluate_lj_finite_forces():
    from rgpot_b200 import _core
    positions = np.array([[0.0, 0.0, 0.0], [1.5, 0.0, 0.0]], dtype=np.float64)
    atom_types = np.array([0, 0], dtype=np.int32)
    box = np.eye(3, dtype=np.float64) * 20.0
    energy, forces, variance = _core.evaluate_lj(positions, atom_types, box)
    assert math.isfinite(float(energy))
    assert forces.shape == (2, 3) and np.all(np.isfinite(forces))
    assert math.isfinite(float(variance))
    assert np.allclose(forces[0] + forces[1], 0.0, atol=1e-9)


@pytest.mark.gpu
def test_ljpot_class_matches_evaluate_lj_and_the_ctypes_surface(oracle):
    import rgpot_b200
    from rgpot_b200 import _core
    positions = np.array([[0.0, 0.0, 0.0], [1.2, 0.0, 0.0], [0.0, 1.2, 0.0]], dtype=np.float64)
    atom_types = np.zeros(3, dtype=np.int32)
    box = np.eye(3, dtype=np.float64) * 30.0
    e1, f1, v1 = _core.evaluate_lj(positions, atom_types, box)
    e2, f2, v2 = _core.LJPot()(positions, atom_types, box)
    assert abs(e1 - e2) < 1e-12 and np.allclose(f1, f2) and abs(v1 - v2) < 1e-12
    e3, f3, _ = rgpot_b200.evaluate_lj(positions, atom_types, box)
    assert e3 == e1 and np.array_equal(f3, f1)
    e_o, f_o, _ = oracle.lj_force(positions, box, 1.0, 15.0, 1.0)
    assert abs(e1 - e_o) <= 1e-10 * abs(e_o) and np.abs(f1 - f_o).max() <= 1e-9
    assert f1.flags.owndata or f1.base is not None  # a fresh array, not a view of the input


@pytest.mark.gpu
def test_shape_errors_are_value_errors():
    from rgpot_b200 import _core
    box = np.eye(3) * 20.0
    with pytest.raises(ValueError, match=r"positions must have shape \(n_atoms, 3\)"):
        _core.LJPot()(np.zeros((2, 2)), np.zeros(2, np.int32), box)
    with pytest.raises(ValueError, match="atom_types length must match n_atoms"):
        _core.evaluate_lj(np.zeros((2, 3)), np.zeros(3, np.int32), box)
    with pytest.raises(ValueError, match=r"box must have shape \(3, 3\)"):
        _core.evaluate_lj(np.zeros((2, 3)), np.zeros(2, np.int32), np.eye(2))
    with pytest.raises(ValueError, match="evaluate_lj expects"):
        _core.evaluate_lj(np.zeros(6), np.zeros(2, np.int32), box)


@pytest.mark.gpu
def test_cuh2_and_batch_through_the_compiled_module(oracle):
    from rgpot_b200 import _core, workloads
    p, z, b = workloads.cuh2_slab218()
    e, f, v = _core.evaluate_cuh2(p, z, b)
    e_o, f_o = oracle.cuh2_force(p, z, b)
    assert abs(e - e_o) <= 1e-10 * abs(e_o) and np.abs(f - f_o).max() <= 1e-9 and v == 0.0
    offsets, pos, Z, boxes = workloads.cuh2_batch(6, seed=3)
    systems = [(pos[int(offsets[s]):int(offsets[s + 1])], Z[int(offsets[s]):int(offsets[s + 1])], boxes[s].reshape(3, 3))
               for s in range(6)]
    out = _core.CuH2Pot().forceBatch(systems)
    for (ei, fi, vi), (ps, zs, bs) in zip(out, systems):
        e_o, f_o = oracle.cuh2_force(ps, zs, bs)
        assert abs(ei - e_o) <= 1e-10 * abs(e_o) and np.abs(fi - f_o).max() <= 1e-9
    with pytest.raises(RuntimeError, match=r"CuH2 potential failed \(status 2\)"):
        zb = z.copy()
        zb[0] = 5
        _core.evaluate_cuh2(p, zb, b)
