"""EAM-Al and Fe-He: the C oracles against an independent 40-digit evaluation written by formula from
the Fortran sources (oracle/second_opinion.py).  The reference pins these two potentials only to
1e-4 / 1e-3 (CppCore/tests/FortranPotsTest.cc:135-155) and its Fortran cannot be compiled here, so
the restatements are otherwise a single reader's port; this is the second reader.  Geometries cover
every branch: the Al truncation radius from both sides, multi-image cells, a sheared cell; for Fe-He the
screened-Coulomb core (r < 1), the exponential bridge (1 .. 2.05), every knot interval of the three
splines, the damped and undamped helium dimer, each of the five cutoffs from both sides, the density
floor (an isolated atom) and helium in a bcc crystal.
"""
import numpy as np
import pytest

from oracle import second_opinion as so


def _close(e, f, e_ref, f_ref, what):
    scale_f = max(1.0, float(np.abs(f_ref).max()))
    assert abs(e - e_ref) <= 1e-12 * max(1.0, abs(e_ref)), (what, e, e_ref)
    assert float(np.abs(f - f_ref).max()) <= 1e-12 * scale_f, (what, float(np.abs(f - f_ref).max()))


def _al_cases():
    big = np.eye(3) * 40.0
    rt = np.sqrt(0.9999) * 7.1
    cases = []
    for r in (2.3, 2.86, 4.05, 6.5, rt - 1e-6, rt + 1e-6, 7.0999, 7.2):
        cases.append((f"dimer r={r}", np.array([[1.0, 2.0, 3.0], [1.0 + r, 2.0, 3.0]]), big))
    rng = np.random.default_rng(1)
    a = 4.05
    fcc = a * np.array([[0, 0, 0], [0, .5, .5], [.5, 0, .5], [.5, .5, 0]])
    cases.append(("fcc cell 4.05 A (every pair has many images)", fcc + rng.normal(0, 0.05, fcc.shape), np.eye(3) * a))
    two = np.concatenate([fcc, fcc + [a, 0, 0]]) + rng.normal(0, 0.08, (8, 3))
    cases.append(("2 x 1 x 1 cells, sheared", two, np.array([[2 * a, 0, 0], [0.9, a, 0], [0.4, -0.7, a]])))
    cl = rng.normal(0, 2.2, (9, 3)) + 20.0
    cases.append(("9-atom cluster in a wide cell", cl, big))
    cases.append(("atoms outside the cell", two + rng.integers(-2, 3, size=(8, 3)) @ np.array([[2 * a, 0, 0], [0, a, 0], [0, 0, a]]),
                  np.array([[2 * a, 0, 0], [0, a, 0], [0, 0, a]])))
    return cases


@pytest.mark.parametrize("case", _al_cases(), ids=lambda c: c[0])
def test_eam_al_oracle_agrees_with_the_second_reader(oracle, case):
    name, pos, box = case
    e, f = oracle.eam_al_force(pos, box)
    e2, f2 = so.eam_al(pos, box)
    _close(e, f, e2, f2, name)


def _fehe_cases():
    big = np.eye(3) * 40.0
    cases = []

    def dimer(z1, z2, r):
        return (f"dimer {z1}-{z2} r={r}", np.array([[3.0, 3.0, 3.0], [3.0 + r * 0.6, 3.0 + r * 0.8, 3.0]]),
                np.array([z1, z2], np.int32), big)

    # Fe-Fe: core, bridge, every knot interval of the pair spline and the d-band spline, both cutoffs
    for r in (0.6, 0.999, 1.001, 1.5, 2.049, 2.051, 2.25, 2.35, 2.45, 2.55, 2.65, 2.75, 2.9, 3.1, 3.25, 3.5, 3.9,
              4.1999, 4.2001, 4.5, 5.0, 5.2999, 5.3001, 5.39):
        cases.append(dimer(26, 26, r))
    # Fe-He: every knot interval, pair cutoff 3.9028 and s-band cutoff 4.10 from both sides
    for r in (1.2, 1.58, 1.65, 1.75, 1.9, 2.2, 3.0, 3.6, 3.9027, 3.9029, 4.0999, 4.1001, 4.6):
        cases.append(dimer(26, 2, r))
        cases.append(dimer(2, 26, r))
    # He-He: damped (x < 1.438 <=> r < 4.2684) and undamped, cutoff 5.4
    for r in (1.8, 2.6, 2.9683, 4.26, 4.27, 5.0, 5.3999, 5.4001):
        cases.append(dimer(2, 2, r))
    rng = np.random.default_rng(3)
    a = 2.87
    bcc = a * np.array([[0, 0, 0], [.5, .5, .5]])
    cell2 = np.concatenate([bcc + [i * a, j * a, 0] for i in range(2) for j in range(2)])
    Z = np.full(len(cell2), 26, np.int32)
    pos = cell2 + rng.normal(0, 0.06, cell2.shape)
    cases.append(("bcc Fe 2 x 2 x 1 cells (multi-image)", pos, Z, np.diag([2 * a, 2 * a, a])))
    posh = np.concatenate([pos, [[0.25 * a, 0.5 * a, 0.1 * a], [1.3 * a, 1.1 * a, 0.6 * a]]])
    Zh = np.concatenate([Z, [2, 2]]).astype(np.int32)
    cases.append(("bcc Fe + two He interstitials, sheared cell", posh, Zh,
                  np.array([[2 * a, 0, 0], [0.6, 2 * a, 0], [-0.3, 0.5, 1.4 * a]])))
    cases.append(("helium only (no d band, no s band)", rng.random((5, 3)) * 4.0 + 10.0, np.full(5, 2, np.int32), big))
    cases.append(("an isolated iron atom next to a far one (density floor)", np.array([[1.0, 1, 1], [30.0, 30, 30]]),
                  np.array([26, 26], np.int32), np.eye(3) * 80.0))
    return cases


@pytest.mark.parametrize("case", _fehe_cases(), ids=lambda c: c[0])
def test_fehe_oracle_agrees_with_the_second_reader(oracle, case):
    name, pos, Z, box = case
    e, f = oracle.fehe_force(pos, Z, box)
    e2, f2 = so.fehe(pos, Z, box)
    _close(e, f, e2, f2, name)
