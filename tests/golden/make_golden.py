"""Generates the committed golden fixtures in tests/golden/ from the CPU oracle (oracle/md_oracle.c) and, for the
potential known-answer table, from the reference formulas evaluated independently in Python floats
(src/potentials.jl:11-29, 66-77; README.md:89-145).  The reference itself cannot run here (no Julia), so
these vectors pin the ORACLE BUILD (regression) and the independent formula table pins the potentials.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import mdoracle as orc  # noqa: E402
from mdjl_b200 import workloads  # noqa: E402


def kat_table():
    a, b = 134.5526623421209, 1.0204081632653061
    rows = []
    for r, s1, s2 in [(1.0, 1, 1), (1.01, 1, 1), (1.02, 1, 1), (0.99, 1, 1), (0.97, 1, 1), (0.95, 0.9, 1.0), (1.0204081632653061, 1, 1), (1.3, 1, 1)]:
        s = (s1 + s2) / 2.0
        if r < b:
            u = a * ((s / r) ** 50.0 - (s / r) ** 49.0) + 1.0
            f = a * (50.0 * (s / r) ** 51.0 - 49.0 * (s / r) ** 50.0)
        else:
            u = f = 0.0
        rows.append((0, 0, 0, 0, r, s1, s2, u, f))
    for eps, rc, r, s1, s2 in [(1, 2.5, 1.0, 1, 1), (1, 2.5, 2 ** (1 / 6), 1, 1), (1, 2.5, 1.5, 1, 1), (1, 2.5, 2.4999, 1, 1), (1, 2.5, 2.5, 1, 1),
                               (0.7, 2.2, 1.1, 0.9, 1.2)]:
        s = (s1 + s2) / 2.0
        if r >= rc:
            u = f = 0.0
        else:
            sr6 = (s / r) ** 6
            u = 4.0 * eps * (sr6 * sr6 - sr6)
            f = 24.0 * eps * (2.0 * sr6 * sr6 - sr6) / r
        rows.append((1, eps, rc, 0, r, s1, s2, u, f))
    for r, s1, s2 in [(1.0, 1.0, 1.0), (1.1, 1.0, 1.2), (0.9, 0.8, 1.1), (1.6, 1.0, 1.0), (2.0, 1.62, 1.62)]:
        rc, na = 1.25, 0.2
        s = 0.5 * (s1 + s2) * (1.0 - na * abs(s1 - s2))
        if r < rc * s:
            c0, c2, c4 = -28.0 / rc ** 12, 48.0 / rc ** 14, -21.0 / rc ** 16
            u = (s / r) ** 12 + c0 + c2 * (r / s) ** 2 + c4 * (r / s) ** 4
            f = 12.0 * s ** 12 / r ** 13 - 2.0 * c2 * r / s ** 2 - 4.0 * c4 * r ** 3 / s ** 4
        else:
            u = f = 0.0
        rows.append((3, rc, na, 0, r, s1, s2, u, f))
    return np.array(rows)


def main():
    np.savez(os.path.join(HERE, "potential_kat.npz"), table=kat_table(),
             columns="tag p0 p1 p2 r sigma1 sigma2 u f (independent Python-float evaluation of the reference formulas)")
    # C1: README example, melted by 2000 oracle NVT steps so that pairs actually interact
    cfg = workloads.phs_fluid(1024)
    v0 = workloads.velocities(1024, 3, workloads.KT_README)
    x, v, f, img, th = orc.run(orc.NVT, cfg["x"], v0, np.zeros_like(v0), np.zeros((1024, 3), np.int32), cfg["diam"], cfg["box"],
                               1.5, orc.POT_PHS, (), 1e-3, 2000, ktemp=workloads.KT_README, tau=0.1, seed=workloads.BASE_SEED)
    ref = orc.forces(x, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, brute=True, counts=True)
    np.savez(os.path.join(HERE, "c1_phs_n1024.npz"), x=x, v=v, f=f, img=img, box=cfg["box"], diam=cfg["diam"], F=ref["F"], E=ref["E"],
             W=ref["W"], n_cut=ref["n_cut"], n_int=ref["n_int"], nbr=ref["nbr"], thermo_tail=th[-5:])
    # C2: 2-D polydisperse plugin, static lattice start
    p = workloads.poly2d(1200)
    for cut in (1.5, 2.03):
        ref = orc.forces(p["x"], p["diam"], p["box"], cut, orc.POT_POLY, (1.25, 0.2), brute=True, counts=True)
        np.savez(os.path.join(HERE, "c2_poly_n1200_cut%s.npz" % cut), x=p["x"], diam=p["diam"], box=p["box"], F=ref["F"], E=ref["E"],
                 W=ref["W"], n_cut=ref["n_cut"], n_int=ref["n_int"], nbr=ref["nbr"])
    # thermostat / Brownian streams
    np.savez(os.path.join(HERE, "rng_streams.npz"),
             philox=np.array([orc.philox((0, 0, 0, 0), (0, 0)), orc.philox((0xffffffff,) * 4, (0xffffffff,) * 2),
                              orc.philox((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0))], dtype=np.uint64),
             bussi=np.array([orc.bussi_noises(1234, s, nf) for s in (0, 1, 99) for nf in (3069.0, 3068.0)]),
             brownian=np.array([orc.brownian_noise(1234, s, i, 3) for s in (0, 7) for i in (0, 1, 1023)]))
    setup_fixtures()
    print("golden fixtures written to", HERE)


CELL = np.array([[11.0, 2.5, -1.5], [0.7, 10.0, 2.0], [-0.4, 0.9, 12.0]])


def setup_fixtures():
    """set-up streams and the general-cell path: velocity normals, uniform positions, wrap_to_box with a
    full matrix and nearest-image forces for a small fixed configuration"""
    x = orc.random_positions(3, 200, (1.0, 1.0, 1.0), 7, 3) @ CELL.T          # uniform fractional coordinates -> cell
    x[::7] += CELL[:, 0] * 3 - CELL[:, 2]                                       # some points outside the cell
    xw, img = orc.wrap_tri(x, np.zeros((200, 3), np.int32), CELL)
    ref = orc.forces_tri(x, np.ones(200), CELL, 1.6, orc.POT_SOFT, (1.0, 1.6))
    np.savez(os.path.join(HERE, "setup_streams.npz"), cell=CELL, x=x, xw=xw, img=img, F=ref["F"], E=ref["E"], W=ref["W"],
             n_cut=ref["n_cut"], n_int=ref["n_int"],
             velocities3=orc.init_velocities(3, 64, 1.4737, 1234, 5), velocities2=orc.init_velocities(2, 64, 0.11, 1234, 5),
             positions3=orc.random_positions(3, 64, (9.0, 11.0, 13.0), 1234, 5),
             positions2=orc.random_positions(2, 64, (9.0, 11.0, 1.0), 1234, 5))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "setup":
        setup_fixtures()
    else:
        main()
