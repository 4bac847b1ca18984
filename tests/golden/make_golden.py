"""Generate the committed golden fixtures.  Run HERE (the container that has /root/reference):

    make -C oracle && make -C oracle ref && python tests/golden/make_golden.py

Sources of truth, in order of authority:
  1. literal values pinned by the reference's own tests (copied as DATA, with file:line);
  2. outputs of the reference's own objects compiled into oracle/_ref/libref.so
     (LJPot / LJClusterPot / vendored vesin, unmodified);
  3. the C restatement (oracle/liboracle.so) where the reference cannot be built here (the
     CuH2 Fortran kernel) -- stored with "source": "oracle" and checked against 1.

The fixtures travel to the GPU box, /root/reference does not.
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyoracle as po  # noqa: E402

REF = "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def dump(name, obj):
    def conv(o):
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, (np.floating, np.integer)):
            return o.item()
        raise TypeError(type(o))

    with open(os.path.join(OUT, name), "w") as fh:
        json.dump(obj, fh, default=conv)
    print("wrote", name)


def read_con(path):
    """eOn .con reader: box lengths, per-type counts, coordinates (CppCore/rgpot/CuH2/tmp.con)."""
    with open(path) as fh:
        lines = fh.read().splitlines()
    lengths = [float(x) for x in lines[2].split()[:3]]
    ntypes = int(lines[6].split()[0])
    counts = [int(x) for x in lines[7].split()[:ntypes]]
    pos, symbols = [], []
    cur = 9
    for t in range(ntypes):
        sym = lines[cur].strip()
        cur += 2
        for _ in range(counts[t]):
            p = lines[cur].split()
            pos.append([float(p[0]), float(p[1]), float(p[2])])
            symbols.append(sym)
            cur += 1
    Z = [29 if s == "Cu" else 1 for s in symbols]
    return np.array(pos), np.array(Z, dtype=np.int32), np.diag(lengths)


def main():
    assert po.ref_available(), "build oracle/_ref first (make -C oracle ref)"

    # ---- LJ 13-atom cluster: CppCore/tests/LJClusterPotTest.cc:24-41, pins :55-69
    cluster = np.array([
        [50.594227, 52.017165, 52.277482], [51.578540, 52.642148, 51.195222],
        [51.243758, 52.919086, 52.218383], [50.490127, 52.736378, 51.417663],
        [51.522643, 50.884535, 51.656125], [50.927263, 51.743087, 51.272643],
        [51.240676, 50.960436, 50.573042], [51.275727, 52.061877, 50.284160],
        [50.434110, 50.976423, 51.879062], [51.695007, 51.919472, 52.038680],
        [51.363596, 52.199101, 53.064923], [52.017483, 51.639074, 51.008389],
        [51.187843, 51.165217, 52.678225]])
    cell = np.diag([101.942400, 103.142600, 102.605500])
    e_c, f_c = po.ref_lj_force(cluster, cell, cluster=True)
    e_p, f_p = po.ref_lj_force(cluster, cell, cluster=False)
    dump("lj13_cluster.json", {
        "source": "CppCore/tests/LJClusterPotTest.cc:24-41 (geometry), :55-69 (pins); forces from "
                  "the reference LJClusterPot / LJPot objects in oracle/_ref",
        "positions": cluster, "box": cell.reshape(9),
        "pin_energy": -39.9653513626179, "pin_max_force": 0.00470419517301762,
        "pin_eon_energy": -39.965379,
        "ref_cluster_energy": e_c, "ref_cluster_forces": f_c,
        "ref_periodic_energy": e_p, "ref_periodic_forces": f_p})

    # ---- CuH2 4-atom pin: CppCore/tests/CuH2PotTest.cc:11-31
    pos4 = np.array([[0.63940268750835, 0.90484742551374, 6.97516498544584],
                     [3.19652040936288, 0.90417430354811, 6.97547796369474],
                     [8.98363230369760, 9.94703496017833, 7.83556854923689],
                     [7.64080177576300, 9.94703114803832, 7.83556986121272]])
    box4 = np.diag([15.345599999999999, 21.702000000000002, 100.0])
    dump("cuh2_pin4.json", {
        "source": "CppCore/tests/CuH2PotTest.cc:11-31",
        "positions": pos4, "Z": [29, 29, 1, 1], "box": box4.reshape(9),
        "pin_energy": -2.7114096242662238,
        "pin_forces": [[1.4919411183978113, -3.9273058476626193E-004, 1.8260603127768336E-004],
                       [-1.4919411183978113, 3.9273058476626193E-004, -1.8260603127768336E-004],
                       [-4.9118653085630006, -1.3944215503304855E-005, 4.7990036210569753E-006],
                       [4.9118653085630006, 1.3944215503304855E-005, -4.7990036210569753E-006]]})

    # ---- CuH2 dimer over RPC: tests/rpc_integ.py:52-78
    dump("cuh2_dimer.json", {
        "source": "tests/rpc_integ.py:52-78",
        "positions": [[0.0, 0.0, 0.0], [1.5, 0.0, 0.0]], "Z": [29, 1], "box": (np.eye(3) * 10).reshape(9),
        "pin_energy": -0.678808, "pin_f0x": -7.5565221, "pin_f1x": 7.5565221, "pin_rel_tol": 1e-6})

    # ---- 218-atom slab: CppCore/rgpot/CuH2/tmp.con (geometry), values from vesin + restatement
    pos, Z, box = read_con(os.path.join(REF, "CppCore/rgpot/CuH2/tmp.con"))
    e_o, f_o, table = po.cuh2_force(pos, Z, box, return_table=True)
    e_r, f_r = po.cuh2_force(pos, Z, box, use_ref=True)
    assert abs(e_o - e_r) < 1e-9 and np.abs(f_o - f_r).max() < 1e-10
    dump("cuh2_slab218.json", {
        "source": "geometry CppCore/rgpot/CuH2/tmp.con; energy/forces: the reference's vendored "
                  "vesin (oracle/_ref) feeding the C restatement of rgpot_cuh2.f90 -- derived, "
                  "not reference-pinned (SURVEY 4.1)",
        "positions": pos, "Z": Z, "box": box.reshape(9),
        "energy": e_r, "forces": f_r, "n_directed_pairs": int(table["row"][-1]),
        "survey_energy": -696.6121457947647})

    # ---- pair sets from the REAL vesin for small periodic cells incl. triclinic + multi-image
    rng = np.random.default_rng(7)
    cases = {}
    boxes = {
        "ortho_small": np.diag([7.23, 7.23, 15.23]),
        "triclinic": np.array([[9.0, 0.0, 0.0], [2.5, 8.0, 0.0], [1.0, -1.5, 10.0]]),
        "tiny_multi_image": np.array([[4.0, 0.0, 0.0], [0.7, 4.2, 0.0], [0.3, 0.4, 5.0]]),
    }
    for name, b in boxes.items():
        n = 40 if name != "tiny_multi_image" else 6
        frac = rng.random((n, 3)) * 1.6 - 0.3  # some atoms outside the cell on purpose
        p = frac @ b
        t = po.nlist(p, b, cutoff=6.1, use_ref=True)
        canon = po.canonical_pairs(t)
        t2 = po.nlist(p, b, cutoff=6.1, use_ref=False)
        assert np.array_equal(canon, po.canonical_pairs(t2)), name
        cases[name] = {"positions": p, "box": b.reshape(9), "cutoff": 6.1,
                       "pairs": canon, "n_self_images": t["n_self_images"]}
    dump("vesin_pairsets.json", {"source": "third_party/vesin (oracle/_ref), full list, skin 0.61, "
                                           "self images dropped as rgpot_neighbors.f90:129-134",
                                 "cases": cases})

    # ---- LJ periodic reference outputs (the reference LJPot itself) incl. rc > L/2
    lj_cases = {}
    for name, (n, L, rc) in {"fluid_rc2.5": (150, 6.0, 2.5), "rc_gt_half_box": (30, 4.0, 3.1),
                             "default_rc15": (20, 10.0, 15.0)}.items():
        p = rng.random((n, 3)) * L * 1.2 - 0.1 * L
        # push apart overlapping atoms
        for _ in range(50):
            d = p[:, None, :] - p[None, :, :]
            d -= L * np.round(d / L)
            r = np.sqrt((d ** 2).sum(-1)) + np.eye(n) * 10
            close = np.argwhere(r < 0.85)
            if len(close) == 0:
                break
            for i, j in close:
                p[i] += 0.2 * d[i, j] / r[i, j]
        b = np.diag([L, L * 1.1, L * 0.95])
        e, f = po.ref_lj_force(p, b, 1.0, rc, 1.0, cluster=False)
        pairs = po.lj_pairs(p, b, rc)
        lj_cases[name] = {"positions": p, "box": b.reshape(9), "u0": 1.0, "cutoff": rc, "psi": 1.0,
                          "energy": e, "forces": f, "pairs": pairs}
    dump("lj_reference_cases.json", {"source": "CppCore/rgpot/LennardJones/LJPot.cc compiled unmodified "
                                               "(oracle/_ref); pair sets from the restated scan",
                                     "cases": lj_cases})


if __name__ == "__main__":
    main()
