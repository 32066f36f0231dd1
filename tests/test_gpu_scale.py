"""Full-size checks (BASELINE.json configs 3 and 5).  N = 2^20: GPU vs the cell-list oracle on a melted fluid.
N = 2^24 (too large for the oracle in seconds): size-independent properties -- two independent GPU code paths
(Verlet list vs per-step cells) agree bit-exactly on pair counts, Newton's third law, momentum and energy conservation,
a strided sample of particles re-checked by direct O(N) minimum-image sums on the host."""
import numpy as np
import pytest

from conftest import force_error, relerr

pytestmark = pytest.mark.gpu


def _melted(md, n, steps=1200, seed=3):
    from mdjl_b200 import workloads
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, 1.4737)
    e = md.Engine(3, n, cfg["box"], 1.5, 0, seed=seed)
    e.upload(cfg["x"], cfg["diam"], velocities=v0)
    e.run_nvt(steps, 1e-3, 1.4737, 0.1, thermo=False)
    return cfg, e


def test_n_2pow20_vs_cell_oracle(md, orc):
    n = 1 << 20
    cfg, e = _melted(md, n)
    E, W, npairs = e.compute_forces()
    x, v, F, img = e.download()
    ref = orc.forces(x, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS)
    assert npairs == ref["n_int"] > n // 4
    assert relerr(E, ref["E"]) <= 1e-12 and relerr(W, ref["W"]) <= 1e-12
    assert force_error(F, ref["F"]) <= 1e-12
    assert e.count_pairs(1.5) == ref["n_cut"]
    # one NVT step and one NVE step against the oracle loop from the same state
    e.upload(x, cfg["diam"], velocities=v, forces=F, images=img)
    e.rng_step = 5000
    t = np.vstack([e.run_nvt(1, 1e-3, 1.4737, 0.1), e.run_nve(1, 1e-3)])
    ox, ov, of, oi, t1 = orc.run(orc.NVT, x, v, F, img, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, (), 1e-3, 1, ktemp=1.4737,
                                 tau=0.1, seed=3, rng_step0=5000)
    ox, ov, of, oi, t2 = orc.run(orc.NVE, ox, ov, of, oi, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, (), 1e-3, 1, seed=3)
    x2, v2, f2, i2 = e.download()
    assert np.array_equal(i2, oi) and np.max(np.abs(x2 - ox)) < 1e-12 and np.max(np.abs(v2 - ov)) < 1e-11
    assert np.array_equal(t[:, 3], np.vstack([t1, t2])[:, 3]) and np.allclose(t[:, :3], np.vstack([t1, t2])[:, :3], rtol=1e-11)
    e.close()


def test_n_2pow20_list_mode_steps_vs_oracle(md, orc):
    """BASELINE configs 3 and 4 through the code path the bench runs: explicit MODE_LIST at N = 2^20 (Verlet list, two-level
    inner list, fused NVE and fused Brownian force kernels, graph replay) against the oracle loop from the same melted
    state: NVT, NVE and Brownian steps, per-step pair counts exact, positions/velocities/thermo to rounding."""
    n = 1 << 20
    cfg, e0 = _melted(md, n)
    x, v, F, img = e0.download()
    e0.close()
    e = md.Engine(3, n, cfg["box"], 1.5, 0, seed=3, mode=md._capi.MODE_LIST)
    e.upload(x, cfg["diam"], velocities=v, forces=F, images=img)
    e.rng_step = 7000
    ks = (4, 5, 4)
    t = np.vstack([e.run_nvt(ks[0], 1e-3, 1.4737, 0.1), e.run_nve(ks[1], 1e-3)])
    x1, v1, f1, i1 = e.download()
    ox, ov, of, oi, t1 = orc.run(orc.NVT, x, v, F, img, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, (), 1e-3, ks[0], ktemp=1.4737,
                                 tau=0.1, seed=3, rng_step0=7000)
    ox, ov, of, oi, t2 = orc.run(orc.NVE, ox, ov, of, oi, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, (), 1e-3, ks[1], seed=3)
    ot = np.vstack([t1, t2])
    assert np.array_equal(i1, oi) and np.max(np.abs(x1 - ox)) < 1e-11 and np.max(np.abs(v1 - ov)) < 1e-10
    assert np.array_equal(t[:, 3], ot[:, 3]) and np.allclose(t[:, :3], ot[:, :3], rtol=1e-10)
    assert force_error(f1, of) <= 1e-11
    # Brownian (C4): fused force + move kernel, noise keyed by (particle id, RNG step)
    tb = e.run_brownian(ks[2], 1e-5, 1.4737)
    x2, _, f2, i2 = e.download()
    bx, _, bf, bi, t3 = orc.run(orc.BROWNIAN, ox, None, of, oi, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, (), 1e-5, ks[2],
                                ktemp=1.4737, seed=3, rng_step0=7000 + ks[0] + ks[1])
    assert np.array_equal(i2, bi) and np.max(np.abs(x2 - bx)) < 1e-10
    assert np.array_equal(tb[:, 3], t3[:, 3]) and np.allclose(tb[:, :2], t3[:, :2], rtol=1e-9)
    assert e.stats()["mode"] == md._capi.MODE_LIST
    e.close()


@pytest.mark.parametrize("ensemble", ["nvt", "nve", "brownian"])
@pytest.mark.parametrize("use_graph", [True, False])
def test_list_mode_trajectory_from_melted_golden(md, orc, ensemble, use_graph):
    """the committed melted C1 snapshot (tests/golden/c1_phs_n1024.npz) stepped 40 times in EXPLICIT list mode (N = 1024
    would otherwise take the persistent small-system kernel) against the oracle loop: the multi-kernel list path of
    C3/C4/C5 on a state where pairs interact from the first step, across list rebuilds"""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_phs_n1024.npz"))
    x, v, diam, box = g["x"], np.ascontiguousarray(g["v"]), g["diam"], g["box"]
    n = x.shape[0]
    e = md.Engine(3, n, box, 1.5, 0, seed=1234, mode=md._capi.MODE_LIST, use_graph=use_graph)
    e.upload(x, diam, velocities=v)
    z, zi = np.zeros_like(x), np.zeros((n, 3), np.int32)
    nsteps = 40
    if ensemble == "nvt":
        t = e.run_nvt(nsteps, 1e-3, 1.4737, 0.1)
        ox, ov, of, oi, ot = orc.run(orc.NVT, x, v, z, zi, diam, box, 1.5, orc.POT_PHS, (), 1e-3, nsteps, ktemp=1.4737, tau=0.1, seed=1234)
    elif ensemble == "nve":
        t = e.run_nve(nsteps, 1e-3)
        ox, ov, of, oi, ot = orc.run(orc.NVE, x, v, z, zi, diam, box, 1.5, orc.POT_PHS, (), 1e-3, nsteps, seed=1234)
    else:
        t = e.run_brownian(nsteps, 1e-5, 1.4737)
        ox, ov, of, oi, ot = orc.run(orc.BROWNIAN, x, None, z, zi, diam, box, 1.5, orc.POT_PHS, (), 1e-5, nsteps, ktemp=1.4737, seed=1234)
    x1, v1, f1, i1 = e.download()
    assert e.stats()["mode"] == md._capi.MODE_LIST
    assert np.array_equal(i1, oi) and np.max(np.abs(x1 - ox)) < 1e-10
    if ensemble != "brownian":
        assert np.max(np.abs(v1 - ov)) < 1e-9 and np.allclose(t[:, 2], ot[:, 2], rtol=1e-10)
    assert np.array_equal(t[:, 3], ot[:, 3]) and np.allclose(t[:, :2], ot[:, :2], rtol=1e-10)
    if ensemble != "brownian":
        assert force_error(f1, of) <= 1e-10
    e.close()


def test_n_2pow24_properties(md):
    n = 1 << 24
    cfg, e = _melted(md, n, steps=400)
    box = cfg["box"]
    E, W, npairs = e.compute_forces()
    x, v, F, img = e.download()
    # Newton's third law over 16.7 M particles
    assert np.max(np.abs(F.sum(axis=0))) <= 1e-9 * np.abs(F).sum()
    assert np.all(x >= 0) and np.all(x <= box)
    # independent code path: per-step cell traversal instead of the Verlet list
    c = md.Engine(3, n, box, 1.5, 0, seed=3, mode=md._capi.MODE_CELLS)
    c.upload(x, cfg["diam"])
    Ec, Wc, npc = c.compute_forces()
    Fc = c.download()[2]
    assert npc == npairs and relerr(Ec, E) <= 1e-12 and relerr(Wc, W) <= 1e-12
    assert force_error(Fc, F) <= 1e-12
    c.close()
    # direct host recomputation for a strided sample of particles (O(N) each, numpy, same minimum-image formula)
    b, a = 1.0204081632653061, 134.5526623421209
    for i in range(0, n, n // 5):
        d = x[i] - x
        d -= box * np.rint(d / box)
        r2 = np.einsum("ij,ij->i", d, d)
        m = (r2 < b * b) & (r2 > 0)
        r = np.sqrt(r2[m])
        f = a * (50.0 * r ** -51.0 - 49.0 * r ** -50.0)
        Fi = ((f / r)[:, None] * d[m]).sum(axis=0)
        assert np.max(np.abs(Fi - F[i])) <= 1e-11 * max(np.max(np.abs(Fi)), 1.0)
    # NVE: momentum and energy over a short horizon
    p0 = v.sum(axis=0)
    t = e.run_nve(60, 1e-3)
    v1 = e.download(positions=False, forces=False, images=False)[1]
    assert np.max(np.abs(v1.sum(axis=0) - p0)) < 1e-6
    Et = t[:, 0] + t[:, 2]
    assert (Et.max() - Et.min()) / abs(Et[0]) < 1e-4
    assert e.stats()["rebuilds"] >= 2
    e.close()
