"""General (triclinic) unit cells, SURVEY 8f row 4: wrap_to_box with a full matrix (src/boundary.jl:7-17), pair enumeration
under the nearest periodic image, and the step loops -- against the oracle's all-pairs restatement (orc_forces_tri,
orc_run_tri).  A diagonal matrix pushed through the general code path must agree with the per-axis path."""
import numpy as np
import pytest

from conftest import force_error, relerr

pytestmark = pytest.mark.gpu

CELL3 = np.array([[11.0, 2.5, -1.5], [0.0, 10.0, 2.0], [0.0, 0.0, 12.0]])     # LAMMPS-style upper triangular
CELL3G = np.array([[11.0, 2.5, -1.5], [0.7, 10.0, 2.0], [-0.4, 0.9, 12.0]])   # general (no zero entries below the diagonal)
CELL2 = np.array([[30.0, 9.0], [0.0, 28.0]])


def _fluid(md, orc, cell, n, dim, seed=7, tol=0.95):
    """random points of the cell with overlaps removed by the packer (itself running under the general cell)"""
    x = md.initialize_random(cell, n, np.random.default_rng(seed), dim, tol=tol)
    return x


@pytest.mark.parametrize("cell,dim,n", [(CELL3, 3, 1100), (CELL3G, 3, 1100), (CELL2, 2, 600)])
@pytest.mark.parametrize("mode", ["cells", "list"])
def test_forces_match_oracle(md, orc, cell, dim, n, mode):
    x = _fluid(md, orc, cell, n, dim)
    diam = np.ones(n)
    modes = {"cells": md._capi.MODE_CELLS, "list": md._capi.MODE_LIST}
    e = md.Engine(dim, n, cell, 1.5, md._capi.POT_LJ, (1.0, 1.5), seed=1, mode=modes[mode])
    e.upload(x, diam)
    E, W, npairs = e.compute_forces()
    F = e.download()[2]
    ref = orc.forces_tri(x, diam, cell, 1.5, orc.POT_LJ, (1.0, 1.5))
    assert npairs == ref["n_int"] > n
    assert e.count_pairs(1.5) == ref["n_cut"]                               # d2 <= cutoff^2 decisions: exact
    assert relerr(E, ref["E"]) <= 1e-12 and relerr(W, ref["W"]) <= 1e-12 and force_error(F, ref["F"]) <= 1e-12
    assert np.max(np.abs(F.sum(axis=0))) < 1e-9 * np.abs(F).sum()
    e.close()


def test_upload_wraps_like_the_reference(md, orc):
    rng = np.random.default_rng(3)
    n = 500
    x = rng.uniform(-40, 60, (n, 3))
    e = md.Engine(3, n, CELL3G, 1.5, md._capi.POT_LJ, (1.0, 1.5))
    e.upload(x, np.ones(n))
    xw, _, _, img = e.download()
    ox, oimg = orc.wrap_tri(x, np.zeros((n, 3), np.int32), CELL3G)
    assert np.array_equal(img, oimg) and np.array_equal(xw, ox)             # same arithmetic: same bits
    assert np.max(np.abs(xw + img @ CELL3G.T - x)) < 1e-12                  # x + U*img is preserved
    fr = np.linalg.solve(CELL3G, xw.T).T
    assert fr.min() > -1e-12 and fr.max() < 1 + 1e-12
    e.close()


@pytest.mark.parametrize("ensemble", ["nve", "nvt", "brownian"])
@pytest.mark.parametrize("cell,dim,n", [(CELL3G, 3, 1100), (CELL2, 2, 600)])
def test_dynamics_match_oracle(md, orc, cell, dim, n, ensemble):
    from mdjl_b200 import workloads
    x = _fluid(md, orc, cell, n, dim, tol=1.0)
    diam = np.ones(n)
    v0 = workloads.velocities(n, dim, 1.0)
    e = md.Engine(dim, n, cell, 1.5, md._capi.POT_PSEUDOHS, seed=99, mode=md._capi.MODE_LIST)
    e.upload(x, diam, velocities=v0)
    nsteps, dt = 40, 1e-3
    z, zi = np.zeros_like(x), np.zeros((n, dim), np.int32)
    if ensemble == "nve":
        t = e.run_nve(nsteps, dt)
        ref = orc.run_tri(orc.NVE, x, v0, z, zi, diam, cell, 1.5, orc.POT_PHS, (), dt, nsteps, seed=99)
    elif ensemble == "nvt":
        t = e.run_nvt(nsteps, dt, 1.2, 0.1)
        ref = orc.run_tri(orc.NVT, x, v0, z, zi, diam, cell, 1.5, orc.POT_PHS, (), dt, nsteps, ktemp=1.2, tau=0.1, seed=99)
    else:
        dt = 1e-5
        t = e.run_brownian(nsteps, dt, 1.2)
        ref = orc.run_tri(orc.BROWNIAN, x, v0, z, zi, diam, cell, 1.5, orc.POT_PHS, (), dt, nsteps, ktemp=1.2, seed=99)
    xo, vo, fo, io_, to = ref
    xg, vg, fg, ig = e.download()
    assert np.array_equal(ig, io_)
    assert np.max(np.abs(xg - xo)) < 1e-10
    if ensemble != "brownian":
        assert np.max(np.abs(vg - vo)) < 1e-9
    assert np.array_equal(t[:, 3], to[:, 3])                                # interacting pairs every step: exact
    assert np.allclose(t[:, :3], to[:, :3], rtol=1e-10, atol=1e-12)
    e.close()


def test_diagonal_matrix_through_the_general_path(md, orc):
    """an orthorhombic cell with a tiny tilt (1e-300) takes the general code path and reproduces the per-axis engine"""
    from mdjl_b200 import workloads
    n = 4096
    cfg = workloads.phs_fluid(n)
    L = float(np.asarray(cfg["box"]).ravel()[0])
    v0 = workloads.velocities(n, 3, 1.4737)
    tilted = np.diag([L, L, L]).astype(np.float64)
    tilted[0, 1] = 1e-300
    out = []
    for box in (cfg["box"], tilted):
        e = md.Engine(3, n, box, 1.5, md._capi.POT_PSEUDOHS, seed=5, mode=md._capi.MODE_LIST)
        e.upload(cfg["x"], cfg["diam"], velocities=v0)
        out.append((e.run_nve(200, 1e-3), e.download(), e.stats()))
        e.close()
    (ta, sa, _), (tb, sb, stb) = out
    assert np.array_equal(ta[:, 3], tb[:, 3])
    assert np.allclose(ta[:, :3], tb[:, :3], rtol=1e-9)
    assert np.max(np.abs(sa[0] - sb[0])) < 1e-8 and np.array_equal(sa[3], sb[3])
    assert stb["rebuilds"] > 1


def test_frames_and_lammps_header_carry_the_tilt(md, tmp_path):
    from mdjl_b200 import api, workloads
    n = 1100
    x = md.initialize_random(CELL3, n, np.random.default_rng(2), 3, tol=1.0)
    e = md.Engine(3, n, CELL3, 1.5, md._capi.POT_PSEUDOHS, seed=4)
    e.upload(x, np.ones(n), velocities=workloads.velocities(n, 3, 2.0))
    e.run_nve(600, 2e-3, thermo=False)
    e.frame_capture(0)
    ours, ref = str(tmp_path / "a"), str(tmp_path / "b")
    e.frame_write_lammps(0, ours, 600, append=False)
    fr = e.frame_wait(0).copy()
    xg, _, _, img = e.download(velocities=False, forces=False)
    assert np.any(img != 0)
    assert np.array_equal(fr[:, 1:4], xg)
    assert np.max(np.abs(fr[:, 4:7] - (xg + img @ CELL3.T))) < 1e-12
    e.frame_flush()
    api.write_to_file_lammps(ref, 600, CELL3, n, xg, img, np.ones(n), 3, mode="w")
    assert open(ours).read().splitlines()[:9] == open(ref).read().splitlines()[:9]      # header incl. xy xz yz
    a = np.loadtxt(ours, skiprows=9)
    b = np.loadtxt(ref, skiprows=9)
    assert np.max(np.abs(a - b)) <= 1.0000001e-6                                # columns agree to the printed digit
    e.close()


def test_errors(md):
    with pytest.raises(md.MdbError) as ei:
        md.Engine(3, 100, np.array([[5.0, 5.0, 0.0], [5.0, 5.0, 0.0], [0.0, 0.0, 5.0]]), 1.0, md._capi.POT_PSEUDOHS)   # singular
    assert ei.value.code == md._capi.ERR_INVALID_ARG
    with pytest.raises(md.MdbError) as ei:
        md.Engine(3, 100, CELL3, 1.5, md._capi.POT_PSEUDOHS, rank=0, nranks=2)       # slabs need a diagonal cell
    assert ei.value.code == md._capi.ERR_UNSUPPORTED_CELL
    e = md.Engine(3, 100, np.array([[4.0, 3.9, 0.0], [0.0, 1.2, 0.0], [0.0, 0.0, 5.0]]), 1.0, md._capi.POT_PSEUDOHS)
    with pytest.raises(md.MdbError) as ei:                                            # perpendicular width 1.2 < 2 cutoff
        e.upload(np.zeros((100, 3)), np.ones(100))
    assert ei.value.code == md._capi.ERR_BOX_TOO_SMALL
    e.close()


def test_against_committed_golden_vectors(md):
    """tests/golden/setup_streams.npz (oracle-generated, committed): general-cell wrap bit-exact, nearest-image forces and
    pair counts, device velocity / position streams"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "setup_streams.npz"))
    cell, x = g["cell"], g["x"]
    e = md.Engine(3, 200, cell, 1.6, md._capi.POT_SOFT, (1.0, 1.6), seed=1234)
    e.upload(x, np.ones(200))
    xw, _, _, img = e.download()
    assert np.array_equal(xw, g["xw"]) and np.array_equal(img, g["img"])
    E, W, npairs = e.compute_forces()
    F = e.download()[2]
    assert npairs == int(g["n_int"]) and e.count_pairs(1.6) == int(g["n_cut"])
    assert relerr(E, float(g["E"])) <= 1e-12 and relerr(W, float(g["W"])) <= 1e-12 and force_error(F, g["F"]) <= 1e-12
    e.close()
    for dim, box, kt, vkey, pkey in ((3, (9.0, 11.0, 13.0), 1.4737, "velocities3", "positions3"), (2, (9.0, 11.0), 0.11, "velocities2", "positions2")):
        e = md.Engine(dim, 64, np.array(box), 1.0, md._capi.POT_SOFT, (1.0, 1.0), seed=1234)
        e.upload(np.zeros((64, dim)), np.ones(64))
        e.random_positions(stream=5)
        assert np.array_equal(e.download()[0], g[pkey])
        e.init_velocities(kt, stream=5)
        assert np.max(np.abs(e.download()[1] - g[vkey])) <= 1e-12 * np.max(np.abs(g[vkey]))
        e.close()
