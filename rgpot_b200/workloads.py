"""Deterministic synthetic geometries of the BASELINE configs.

C1  13-atom LJ cluster (the reference's "lj38" fixture is this 13-atom cluster)
C2  CuH2 4-atom pin and the 218-atom Cu slab + H2 (tests/golden fixtures)
C3  LJ fluid, 1 000 000 atoms, rho* = 0.8442, rc = 2.5 sigma, cubic periodic box
C4  Cu slab + H2, 256 512 atoms, triclinic box commensurate with the fcc lattice
C5  batch of perturbed copies of the 218-atom slab (NEB / dimer images)

Pure numpy, no reference files needed at run time.
"""
from __future__ import annotations

import json
import os

import numpy as np

_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _load(name):
    with open(os.path.join(_GOLDEN, name)) as fh:
        return json.load(fh)


def lj13_cluster():
    g = _load("lj13_cluster.json")
    return (np.array(g["positions"]), np.ones(13, dtype=np.int32), np.array(g["box"]).reshape(3, 3))


def cuh2_pin4():
    g = _load("cuh2_pin4.json")
    return (np.array(g["positions"]), np.array(g["Z"], dtype=np.int32), np.array(g["box"]).reshape(3, 3))


def cuh2_slab218():
    g = _load("cuh2_slab218.json")
    return (np.array(g["positions"]), np.array(g["Z"], dtype=np.int32), np.array(g["box"]).reshape(3, 3))


_FCC_BASIS = np.array([[0.0, 0.0, 0.0], [0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]])


def _fcc(cells, a):
    """fcc sites of cells^3 conventional cells, loop order ix, iy, iz, basis (PotBench.cc:61-83)."""
    g = np.arange(cells, dtype=np.float64)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    sites = (corner[:, None, :] + _FCC_BASIS[None, :, :]).reshape(-1, 3)
    return a * sites


def _sin_jitter(n, amp):
    k = np.arange(n, dtype=np.float64)
    return amp * np.stack([np.sin(1.7 * k), np.sin(2.3 * k + 1.0), np.sin(3.1 * k + 2.0)], axis=1)


def wrap_into_cell(pos, box):
    """Move every atom to its image inside the cell (fractional coordinates in [0, 1))."""
    frac = pos @ np.linalg.inv(box)
    frac -= np.floor(frac)
    frac[frac >= 1.0] = 0.0
    return np.ascontiguousarray(frac @ box)


def lj_fluid(n_atoms=1_000_000, rho=0.8442, seed=20261019, sigma_d=0.1):
    """C3 recipe: fcc sites at density rho, sin jitter 0.05, Gaussian displacements sigma_d.

    Returns positions, Z (18), box, and the LJConfig numbers (u0, cutoff, psi) = (1, 2.5, 1).
    """
    cells = int(np.ceil((n_atoms / 4.0) ** (1.0 / 3.0)))
    while 4 * cells ** 3 < n_atoms:
        cells += 1
    n_sites = 4 * cells ** 3
    L = (n_sites / rho) ** (1.0 / 3.0)
    pos = _fcc(cells, L / cells)[:n_atoms]
    pos = pos + _sin_jitter(n_atoms, 0.05)
    rng = np.random.default_rng(seed)
    pos = pos + rng.normal(0.0, sigma_d, size=pos.shape)
    pos = np.mod(pos, L)  # keep every atom inside [0, L): the fluid has no preferred image
    box = np.diag([L, L, L])
    return np.ascontiguousarray(pos), np.full(n_atoms, 18, dtype=np.int32), box


def cu_slab_triclinic(cells=40, a=3.615, n_h2_side=16, vacuum=20.0):
    """C4 recipe: cells^3 fcc Cu + (n_h2_side^2) H2 above the top layer, sheared periodic box."""
    cu = _fcc(cells, a)
    n_cu = len(cu)
    span = cells * a
    top = span - 0.5 * a
    gx = (np.arange(n_h2_side) + 0.5) * span / n_h2_side
    hx, hy = np.meshgrid(gx, gx, indexing="ij")
    centre = np.stack([hx.ravel(), hy.ravel(), np.full(hx.size, top + 2.08)], axis=1)
    h1 = centre + np.array([-0.37, 0.0, 0.0])
    h2 = centre + np.array([+0.37, 0.0, 0.0])
    pos = np.concatenate([cu, np.stack([h1, h2], axis=1).reshape(-1, 3)], axis=0)
    Z = np.concatenate([np.full(n_cu, 29, dtype=np.int32), np.ones(2 * hx.size, dtype=np.int32)])
    pos = pos + _sin_jitter(len(pos), 0.05)
    tilt = cells // 4
    box = np.array([[span, 0.0, 0.0],
                    [tilt * a, span, 0.0],
                    [(tilt // 2) * a, (tilt // 2) * a, span + vacuum]])
    # the cubic block of sites is re-expressed inside the sheared cell (same crystal: the tilt is
    # a lattice vector); the 0.1 A offset keeps jittered atoms off the cell faces
    return wrap_into_cell(pos + 0.1, box), Z, box


def cuh2_batch(n_configs, seed=1000, sigma=0.05):
    """C5 recipe: copies of the 218-atom slab with N(0, sigma^2) displacements.

    Returns packed arrays: offsets (n+1), positions (n*218, 3), Z (n*218), boxes (n, 9).
    (One vectorised generator seeded with `seed` instead of one generator per configuration.)
    """
    base, Z, box = cuh2_slab218()
    n = len(base)
    rng = np.random.default_rng(seed)
    pos = base[None, :, :] + rng.normal(0.0, sigma, size=(n_configs, n, 3))
    offsets = np.arange(n_configs + 1, dtype=np.uint64) * n
    return (offsets, np.ascontiguousarray(pos.reshape(-1, 3)), np.tile(Z, n_configs),
            np.tile(box.reshape(9), (n_configs, 1)))


def cu_slab_with_h2(cells=2, a=3.615, gap=8.0, height=2.08, bond=0.74):
    """The geometry of the reference's pure-Fortran physics test
    (CppCore/rgpot/fortran/test_cuh2.f90:146-199): 4*cells^3 Cu + H2, box span x span x span+gap."""
    g = np.arange(cells, dtype=np.float64)
    ix, iy, iz = np.meshgrid(g, g, g, indexing="ij")
    corner = np.stack([ix.ravel(), iy.ravel(), iz.ravel()], axis=1)
    cu = a * (corner[:, None, :] + _FCC_BASIS[None, :, :]).reshape(-1, 3)
    span = a * cells
    top = span - 0.5 * a
    centre = 0.5 * span
    h = np.array([[centre - 0.5 * bond, centre, top + height], [centre + 0.5 * bond, centre, top + height]])
    pos = np.concatenate([cu, h])
    Z = np.concatenate([np.full(len(cu), 29, dtype=np.int32), np.ones(2, dtype=np.int32)])
    box = np.diag([span, span, span + gap])
    return pos, Z, box
