"""Parity pin against the REFERENCE ITSELF (missing #2).

tests/golden/ref_c1.npz / ref_c2.npz hold outputs of the stock MolecularDynamics.jl package on the committed snapshots,
produced by julia/make_reference_golden.jl (+ tests/golden/make_reference_inputs.py, ref_to_npz.py).  Julia is not in the
build image, so the files cannot be generated here; until a maintainer commits them these tests SKIP and DESIGN.md keeps
saying "parity unpinned".  With the files present they pin, at the tolerance north_star states (pair counts bit-exact;
forces, energy, virial within 1e-12 relative):
  * the C oracle (CPU, runs everywhere) against the reference;
  * the CUDA path (-m gpu) against the reference, forces and a 50-step NVE trajectory."""
import os

import numpy as np
import pytest

from conftest import force_error, relerr

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"c1": dict(src="c1_phs_n1024.npz", dim=3, pot="phs", pp=(), dt=1e-3),
         "c2": dict(src="c2_poly_n1200_cut1.5.npz", dim=2, pot="poly", pp=(1.25, 0.2), dt=5e-3)}


def _load(name):
    path = os.path.join(GOLD, "ref_%s.npz" % name)
    if not os.path.exists(path):
        pytest.skip("parity unpinned: %s absent -- run julia/make_reference_golden.jl where Julia is available" % os.path.basename(path))
    from mdjl_b200 import workloads
    c = CASES[name]
    g = np.load(os.path.join(GOLD, c["src"]))
    n, d = g["x"].shape
    v = np.ascontiguousarray(g["v"]) if "v" in g.files else workloads.velocities(n, d, 0.11)
    return np.load(path), c, g["x"], v, g["diam"], np.asarray(g["box"], dtype=np.float64).ravel()[:d]


@pytest.mark.parametrize("name", ["c1", "c2"])
def test_oracle_matches_the_reference(orc, name):
    ref, c, x, v, diam, box = _load(name)
    tag = {"phs": orc.POT_PHS, "poly": orc.POT_POLY}[c["pot"]]
    o = orc.forces(x, diam, box, 1.5, tag, c["pp"])
    assert o["n_cut"] == int(ref["n_cut"]) and o["n_int"] == int(ref["n_int"])          # bit-exact pair sets
    assert relerr(o["E"], float(ref["E"])) <= 1e-12 and relerr(o["W"], float(ref["W"])) <= 1e-12
    assert force_error(o["F"], ref["F"]) <= 1e-12
    n, d = x.shape
    ox, ov, of, oi, ot = orc.run(orc.NVE, x, v, np.zeros_like(x), np.zeros((n, d), np.int32), diam, box, 1.5, tag, c["pp"], c["dt"], 50)
    assert np.array_equal(oi, ref["nve_img"])
    assert np.max(np.abs(ox - ref["nve_x"])) < 1e-10 and np.max(np.abs(ov - ref["nve_v"])) < 1e-9
    assert np.allclose(ot[:, :3], ref["nve_thermo"], rtol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["c1", "c2"])
@pytest.mark.parametrize("mode", ["auto", "list", "cells"])
def test_gpu_matches_the_reference(md, name, mode):
    ref, c, x, v, diam, box = _load(name)
    n, d = x.shape
    tag = {"phs": md._capi.POT_PSEUDOHS, "poly": md._capi.POT_POLY}[c["pot"]]
    m = {"auto": md._capi.MODE_AUTO, "list": md._capi.MODE_LIST, "cells": md._capi.MODE_CELLS}[mode]
    e = md.Engine(d, n, box, 1.5, tag, c["pp"], seed=1, mode=m)
    e.upload(x, diam, velocities=v)
    E, W, npairs = e.compute_forces()
    assert npairs == int(ref["n_int"]) and e.count_pairs(1.5) == int(ref["n_cut"])
    assert relerr(E, float(ref["E"])) <= 1e-12 and relerr(W, float(ref["W"])) <= 1e-12
    assert force_error(e.download()[2], ref["F"]) <= 1e-12
    e.upload(x, diam, velocities=v)
    t = e.run_nve(50, c["dt"])
    x1, v1, _, i1 = e.download()
    assert np.array_equal(i1, ref["nve_img"])
    assert np.max(np.abs(x1 - ref["nve_x"])) < 1e-10 and np.max(np.abs(v1 - ref["nve_v"])) < 1e-9
    assert np.allclose(t[:, :3], ref["nve_thermo"], rtol=1e-10)
    e.close()
