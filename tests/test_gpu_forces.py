"""GPU parity of the pair-force path (C ABI: mdb_compute_forces / mdb_count_pairs) against the CPU oracle.
Bars (north_star / SURVEY 8c): pair counts bit-exact; per-particle forces, energy and virial within 1e-12 relative."""
import numpy as np
import pytest

from conftest import force_error, relerr

pytestmark = pytest.mark.gpu
TOL = 1e-12


def _engine(md, cfg, dim, cutoff, tag, params=(), mode="auto", **kw):
    modes = {"auto": md._capi.MODE_AUTO, "cells": md._capi.MODE_CELLS, "list": md._capi.MODE_LIST}
    n = cfg["x"].shape[0]
    e = md.Engine(dim, n, cfg["box"], cutoff, tag, params, seed=7, mode=modes[mode], **kw)
    e.upload(cfg["x"], cfg["diam"])
    return e


def _golden(name):
    import os
    return dict(np.load(os.path.join(os.path.dirname(__file__), "golden", name)))


def _melt(md, cfg, dim, tag, params, cutoff, kt, dt, nsteps, seed=17):
    """thermalise a lattice start on the GPU so that pairs actually interact; the force comparison that follows is
    against the oracle on whatever positions come out"""
    from mdjl_b200 import workloads
    n = cfg["x"].shape[0]
    e = md.Engine(dim, n, cfg["box"], cutoff, tag, params, seed=seed)
    e.upload(cfg["x"], cfg["diam"], velocities=workloads.velocities(n, dim, kt))
    e.run_nvt(nsteps, dt, kt, 100 * dt, thermo=False)
    x = e.download()[0]
    e.close()
    out = dict(cfg)
    out["x"] = x
    return out


def _check(md, orc, cfg, dim, cutoff, tag, params, mode, brute_oracle=False):
    e = _engine(md, cfg, dim, cutoff, tag, params, mode)
    E, W, npairs = e.compute_forces()
    _, _, F, _ = e.download()
    ref = orc.forces(cfg["x"], cfg["diam"], cfg["box"], cutoff, tag, params, brute=brute_oracle)
    assert npairs == ref["n_int"]
    assert relerr(E, ref["E"]) <= TOL, (E, ref["E"])
    assert relerr(W, ref["W"]) <= TOL, (W, ref["W"])
    assert force_error(F, ref["F"]) <= TOL
    # what map_pairwise! visits: pairs with d2 <= cutoff^2, bit-exact, total and per particle
    ncut, per = e.count_pairs(cutoff, per_particle=True)
    refc = orc.forces(cfg["x"], cfg["diam"], cfg["box"], cutoff, tag, params, brute=brute_oracle, counts=True)
    assert ncut == refc["n_cut"]
    assert np.array_equal(per, refc["nbr"])
    # the count re-sorted the state: the force path must still give the same answer afterwards
    E2, W2, np2 = e.compute_forces()
    assert np2 == npairs and relerr(E2, E) <= 1e-14
    st = e.stats()
    e.close()
    return st


@pytest.mark.parametrize("mode", ["cells", "list"])
def test_phs_c1_vs_brute_force(md, orc, mode):
    """C1: 3-D monodisperse pseudo-hard spheres, N=1024, phi=0.47 (README example) vs the O(N^2) oracle."""
    g = _golden("c1_phs_n1024.npz")   # lattice start melted by 2000 oracle NVT steps (tests/golden/make_golden.py)
    cfg = dict(x=g["x"], diam=g["diam"], box=g["box"])
    st = _check(md, orc, cfg, 3, 1.5, orc.POT_PHS, (), mode, brute_oracle=True)
    assert st["mode"] == {"cells": 1, "list": 2}[mode]
    # and against the committed golden vector itself
    e = _engine(md, cfg, 3, 1.5, orc.POT_PHS, (), mode)
    E, W, npairs = e.compute_forces()
    assert npairs == int(g["n_int"]) > 400 and relerr(E, float(g["E"])) <= TOL and relerr(W, float(g["W"])) <= TOL
    assert force_error(e.download()[2], g["F"]) <= TOL
    assert e.count_pairs(1.5) == int(g["n_cut"])
    e.close()


@pytest.mark.parametrize("mode", ["cells", "list"])
def test_poly2d_c2(md, orc, mode):
    """C2: 2-D non-additive polydisperse plugin, N=1200, cutoff 1.5 truncates the potential range."""
    from mdjl_b200 import workloads
    cfg = workloads.poly2d(1200)
    _check(md, orc, cfg, 2, 1.5, orc.POT_POLY, (1.25, 0.2), mode, brute_oracle=True)
    _check(md, orc, cfg, 2, 2.03, orc.POT_POLY, (1.25, 0.2), mode)


@pytest.mark.parametrize("mode", ["cells", "list"])
@pytest.mark.parametrize("dim", [2, 3])
def test_lennard_jones(md, orc, mode, dim):
    from mdjl_b200 import workloads
    n = 4096 if dim == 3 else 2500
    cfg = workloads.lj_fluid(n, rho=0.8, dim=dim)
    cfg = _melt(md, cfg, dim, orc.POT_LJ, (1.0, 2.5), 2.5, 1.2, 2e-3, 500)
    _check(md, orc, cfg, dim, 2.5, orc.POT_LJ, (1.0, 2.5), mode)
    _check(md, orc, cfg, dim, 3.0, orc.POT_LJ, (0.7, 2.2), mode)  # potential range below the neighbour cutoff


@pytest.mark.parametrize("mode", ["cells", "list"])
def test_lj_xplor_bug_for_bug(md, orc, mode):
    from mdjl_b200 import workloads
    cfg = workloads.lj_fluid(4096, rho=0.8, dim=3)
    cfg = _melt(md, cfg, 3, orc.POT_LJ, (1.0, 2.5), 2.5, 1.2, 2e-3, 500)
    _check(md, orc, cfg, 3, 2.5, orc.POT_XPLOR, (1.0, 2.0, 2.5), mode)


def test_polydisperse_3d_phs(md, orc):
    """sigma != 1 with the absolute PseudoHS cut"""
    from mdjl_b200 import workloads
    cfg = workloads.phs_fluid(4000, phi=0.40)
    rng = np.random.default_rng(3)
    cfg["diam"] = rng.uniform(0.9, 1.0, size=4000)
    cfg = _melt(md, cfg, 3, orc.POT_PHS, (), 1.5, 1.5, 1e-3, 1500)
    st = _check(md, orc, cfg, 3, 1.5, orc.POT_PHS, (), "auto")
    assert st["mode"] == 2


def test_tiny_box_all_pairs_kernel(md, orc):
    """boxes with fewer than three cells per side go through the all-pairs kernel (same arithmetic as the brute oracle)"""
    rng = np.random.default_rng(5)
    n, L = 64, 3.7
    g = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3)
    cfg = dict(x=(g + 0.5) * (L / 4) + rng.uniform(-0.03, 0.03, (n, 3)), diam=np.ones(n), box=np.full(3, L))
    _check(md, orc, cfg, 3, 1.5, orc.POT_LJ, (1.0, 1.5), "auto", brute_oracle=True)


def test_unwrapped_input_and_images(md, orc):
    """positions outside the cell are wrapped on upload with wrap_to_box arithmetic; x + L*img is preserved"""
    from mdjl_b200 import workloads
    g = _golden("c1_phs_n1024.npz")
    cfg = dict(x=g["x"], diam=g["diam"], box=g["box"])
    rng = np.random.default_rng(11)
    shift = rng.integers(-2, 3, size=cfg["x"].shape)
    xin = cfg["x"] + shift * cfg["box"]
    e = md.Engine(3, 1024, cfg["box"], 1.5, orc.POT_PHS, seed=1)
    e.upload(xin, cfg["diam"])
    E, W, npairs = e.compute_forces()
    x, _, F, img = e.download()
    ref = orc.forces(cfg["x"], cfg["diam"], cfg["box"], 1.5, orc.POT_PHS)
    assert npairs == ref["n_int"] and relerr(E, ref["E"]) < 1e-9
    assert np.all(x >= 0) and np.all(x <= cfg["box"])
    assert np.max(np.abs(x + img * cfg["box"] - xin)) < 1e-12
    for i in range(0, 1024, 97):  # same arithmetic as the oracle's wrap
        wx, wi = orc.wrap(xin[i], np.zeros(3, np.int32), cfg["box"])
        inside = (xin[i] >= 0) & (xin[i] < cfg["box"])
        assert np.array_equal(np.where(inside, xin[i], wx), x[i])
    e.close()


def test_bitwise_reproducible(md, orc):
    """canonical in-cell order + fixed-order reductions: two independent engines give identical bits"""
    from mdjl_b200 import workloads
    cfg = _melt(md, workloads.phs_fluid(8192), 3, orc.POT_PHS, (), 1.5, 1.4737, 1e-3, 1500)
    out = []
    for _ in range(2):
        e = _engine(md, cfg, 3, 1.5, orc.POT_PHS, (), "list")
        E, W, n = e.compute_forces()
        out.append((E, W, n, e.download()[2].copy()))
        e.close()
    assert out[0][:3] == out[1][:3]
    assert np.array_equal(out[0][3], out[1][3])


def test_error_paths(md):
    from mdjl_b200 import _capi
    with pytest.raises(md.MdbError) as ei:   # singular cell matrix (tilted cells themselves are supported: test_gpu_triclinic.py)
        md.Engine(3, 100, np.array([[5, 5, 0], [5, 5, 0], [0, 0, 5.0]]), 1.5, 0)
    assert ei.value.code == _capi.ERR_INVALID_ARG
    with pytest.raises(md.MdbError) as ei:
        md.Engine(3, 100, 5.0, 1.5, 42)
    assert ei.value.code == _capi.ERR_UNSUPPORTED_POTENTIAL
    e = md.Engine(3, 8, 2.5, 1.5, 1, (1.0, 2.5))
    with pytest.raises(md.MdbError) as ei:  # cutoff >= L/2
        e.upload(np.random.default_rng(0).uniform(0, 2.5, (8, 3)), np.ones(8))
    assert ei.value.code == _capi.ERR_BOX_TOO_SMALL
    e2 = md.Engine(3, 8, 10.0, 1.5, 0)
    with pytest.raises(md.MdbError) as ei:
        e2.run_nve(1, 1e-3)
    assert ei.value.code == _capi.ERR_STATE
    e2.upload(np.random.default_rng(0).uniform(0, 10, (8, 3)), np.ones(8))
    with pytest.raises(md.MdbError) as ei:  # velocities never set
        e2.run_nve(1, 1e-3)
    assert ei.value.code == _capi.ERR_STATE


LJ_BODY = """// MDB_DENSE_HITS -- same statements as lj_unshifted (src/potentials.jl:66-77)
double epsilon = p[0], r_cut = p[1];
double sigma = (sigma1 + sigma2) / 2.0;
if (r >= r_cut) { u = 0.0; f = 0.0; return false; }
double sr = sigma / r; double sr2 = sr * sr; double sr6 = (sr2 * sr2) * sr2; double sr12 = sr6 * sr6;
u = 4.0 * epsilon * (sr12 - sr6);
f = 24.0 * epsilon * (2.0 * sr12 - sr6) / r;
return true;"""

HARMONIC_BODY = """double s = 0.5 * (sigma1 + sigma2);
if (r >= s) { u = 0.0; f = 0.0; return false; }
double x = 1.0 - r / s;
u = 0.5 * p[0] * x * x;
f = p[0] * x / s;
return true;"""


def test_user_potential_nvrtc(md, orc):
    """the open plugin contract on the device: a user `evaluate` body compiled with NVRTC into the same kernel templates"""
    from mdjl_b200 import workloads
    cfg = workloads.lj_fluid(4096, rho=0.8, dim=3)
    cfg = _melt(md, cfg, 3, orc.POT_LJ, (1.0, 2.5), 2.5, 1.2, 2e-3, 400)
    # (1) user code restating Lennard-Jones must reproduce the built-in functor bit for bit, in both neighbour modes
    for mode in ("cells", "list"):
        ref = _engine(md, cfg, 3, 2.5, orc.POT_LJ, (1.0, 2.5), mode)
        E0, W0, n0 = ref.compute_forces()
        F0 = ref.download()[2]
        ref.close()
        modes = {"cells": md._capi.MODE_CELLS, "list": md._capi.MODE_LIST}
        e = md.Engine(3, 4096, cfg["box"], 2.5, 0, seed=7, mode=modes[mode])
        e.set_user_potential(LJ_BODY, (1.0, 2.5), 2.5)
        e.upload(cfg["x"], cfg["diam"])
        E, W, n = e.compute_forces()
        assert (E, W, n) == (E0, W0, n0) and np.array_equal(e.download()[2], F0)
        e.close()
    # (2) a potential the library does not ship: harmonic repulsion, against a direct O(N^2) numpy sum
    rng = np.random.default_rng(4)
    n, L, k = 1500, 12.0, 30.0
    x = rng.uniform(0, L, (n, 3))
    diam = rng.uniform(0.8, 1.2, n)
    e = md.Engine(3, n, L, 1.5, 0, seed=7)
    e.set_user_potential(HARMONIC_BODY, (k,), 1.2)
    e.upload(x, diam, velocities=np.zeros((n, 3)))
    E, W, npairs = e.compute_forces()
    F = e.download()[2]
    d = x[:, None, :] - x[None, :, :]
    d -= L * np.rint(d / L)
    r = np.sqrt((d * d).sum(-1))
    s = 0.5 * (diam[:, None] + diam[None, :])
    m = (r < s) & (r > 0)
    xx = np.where(m, 1.0 - r / np.where(m, s, 1.0), 0.0)
    Eref = 0.25 * k * (xx * xx).sum()
    fmag = np.where(m, k * xx / s, 0.0)
    Fref = ((fmag / np.where(m, r, 1.0))[:, :, None] * d).sum(axis=1)
    assert npairs == int(m.sum()) // 2 and relerr(E, Eref) < 1e-12 and force_error(F, Fref) < 1e-12
    t = e.run_nve(200, 2e-3)          # soft spheres relax: potential energy converts to kinetic, total is conserved
    Et = t[:, 0] + t[:, 2]
    assert abs(Et[-1] - Et[0]) < 2e-3 * abs(Et[0]) and t[-1, 2] > 0
    e.close()
    # (3) code that does not compile is reported with the compiler log
    e = md.Engine(3, 10, 10.0, 1.5, 0)
    with pytest.raises(md.MdbError) as ei:
        e.set_user_potential("u = nonsense; return true;", (), 1.0)
    assert ei.value.code == md._capi.ERR_NVRTC and "nonsense" in str(ei.value)
    e.close()
