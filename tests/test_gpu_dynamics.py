"""GPU parity of the step loop (mdb_run_nve / mdb_run_nvt / mdb_run_brownian) against the oracle's restatement of
src/simulation.jl:88-108 / :231-250 on identical inputs, with the same counter-based RNG streams."""
import numpy as np
import pytest

from conftest import relerr

pytestmark = pytest.mark.gpu


def _setup(md, n=1024, mode="auto", use_graph=True, seed=99, kt=1.4737):
    from mdjl_b200 import workloads
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, kt)
    modes = {"auto": md._capi.MODE_AUTO, "cells": md._capi.MODE_CELLS, "list": md._capi.MODE_LIST}
    e = md.Engine(3, n, cfg["box"], 1.5, md._capi.POT_PSEUDOHS, seed=seed, mode=modes[mode], use_graph=use_graph)
    e.upload(cfg["x"], cfg["diam"], velocities=v0)
    return cfg, v0, e


def _oracle_run(orc, ens, cfg, v0, nsteps, dt, seed=99, rng_step0=0, **kw):
    n, dim = cfg["x"].shape
    return orc.run(ens, cfg["x"], v0, np.zeros_like(cfg["x"]), np.zeros((n, dim), np.int32), cfg["diam"], cfg["box"], 1.5,
                   orc.POT_PHS, (), dt, nsteps, seed=seed, rng_step0=rng_step0, **kw)


@pytest.mark.parametrize("mode,use_graph", [("cells", False), ("cells", True), ("list", False), ("list", True)])
def test_nve_matches_oracle(md, orc, mode, use_graph):
    cfg, v0, e = _setup(md, mode=mode, use_graph=use_graph)
    dt, nsteps = 1e-3, 50
    t = e.run_nve(nsteps, dt)
    x, v, f, img = e.download()
    ox, ov, of, oimg, ot = _oracle_run(orc, orc.NVE, cfg, v0, nsteps, dt)
    assert np.array_equal(img, oimg)
    assert np.max(np.abs(x - ox)) < 1e-10
    assert np.max(np.abs(v - ov)) < 1e-9
    assert np.array_equal(t[:, 3], ot[:, 3])                      # interacting pair counts, every step, exact
    assert np.allclose(t[:, :3], ot[:, :3], rtol=1e-10, atol=0)    # U, W, KE every step
    # step 0 uses zero forces: positions after step 0 are x0 + v0*dt wrapped -- covered by the oracle run
    e.close()


def test_graph_and_eager_are_bit_identical(md, orc):
    out = []
    for use_graph in (False, True):
        cfg, v0, e = _setup(md, n=4096, mode="list", use_graph=use_graph)
        t = e.run_nvt(120, 1e-3, 1.4737, 0.1)
        out.append((t, e.download(), e.stats()))
        e.close()
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)
    assert out[0][2]["rebuilds"] == out[1][2]["rebuilds"] > 1


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_nve_step_is_bit_identical_to_reference_order(md, orc, use_graph):
    """NVE list-mode runs fold the next step's kick-drift into the force kernel (no K5 sweep); positions, velocities,
    forces, images and every thermo row must equal the reference's kernel order (no_fuse) bit for bit, for any split of
    the run into calls (1-step calls exercise first-step == last-step), across list rebuilds."""
    from mdjl_b200 import workloads
    n = 4096
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, 1.4737)
    out = []
    for no_fuse in (True, False):
        e = md.Engine(3, n, cfg["box"], 1.5, md._capi.POT_PSEUDOHS, seed=99, mode=md._capi.MODE_LIST, use_graph=use_graph, no_fuse=no_fuse)
        e.upload(cfg["x"], cfg["diam"], velocities=v0)
        rows = [e.run_nve(k, 1e-3) for k in (1, 2, 1, 157, 40)]
        out.append((np.concatenate(rows), e.download(), e.stats()))
        e.close()
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)
    assert out[0][2]["rebuilds"] == out[1][2]["rebuilds"] > 1
    # one kick-drift launch per run call instead of one per step
    assert out[1][2]["kernel_launches"] < out[0][2]["kernel_launches"]


@pytest.mark.parametrize("use_graph", [False, True])
def test_fused_brownian_step_is_bit_identical_to_reference_order(md, orc, use_graph):
    """Brownian list-mode runs do the move inside the force kernel (no separate K8 sweep): same bits as forces-then-move
    in two kernels, across list rebuilds (exact displacement test included) and several run calls"""
    from mdjl_b200 import workloads
    n = 4096
    cfg = workloads.phs_fluid(n)
    out = []
    for no_fuse in (True, False):
        e = md.Engine(3, n, cfg["box"], 1.5, md._capi.POT_PSEUDOHS, seed=31, mode=md._capi.MODE_LIST, use_graph=use_graph, no_fuse=no_fuse)
        e.upload(cfg["x"], cfg["diam"])
        rows = [e.run_brownian(k, 2e-5, 1.4737) for k in (1, 2, 400, 37)]
        out.append((np.concatenate(rows), e.download(), e.stats(), e.rng_step))
        e.close()
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)
    assert out[0][2]["rebuilds"] == out[1][2]["rebuilds"] > 1 and out[0][3] == out[1][3]
    assert out[1][2]["kernel_launches"] < out[0][2]["kernel_launches"]


def test_list_mode_equals_cell_mode(md, orc):
    """same pair set every step: pair counts identical, energies to rounding, over several list rebuilds"""
    res = {}
    for mode in ("cells", "list"):
        cfg, v0, e = _setup(md, n=4096, mode=mode)
        res[mode] = (e.run_nve(300, 1e-3), e.stats())
        e.close()
    a, b = res["cells"][0], res["list"][0]
    assert np.array_equal(a[:, 3], b[:, 3])
    assert np.allclose(a[:, :3], b[:, :3], rtol=1e-9)
    assert res["list"][1]["rebuilds"] < res["cells"][1]["rebuilds"] == 300


def test_nvt_bussi_matches_oracle(md, orc):
    cfg, v0, e = _setup(md)
    dt, nsteps, tau = 1e-3, 40, 0.1
    kt = np.linspace(1.4737, 1.2, nsteps)  # ktemp(step+1) schedule, as a LinearRamp would give
    t = e.run_nvt(nsteps, dt, kt, tau)
    x, v, f, img = e.download()
    ox, ov, of, oimg, ot = _oracle_run(orc, orc.NVT, cfg, v0, nsteps, dt, ktemp=kt, tau=tau)
    assert np.array_equal(img, oimg)
    assert np.max(np.abs(x - ox)) < 1e-10 and np.max(np.abs(v - ov)) < 1e-9
    assert np.allclose(t[:, :3], ot[:, :3], rtol=1e-10)
    assert e.rng_step == nsteps
    # chained call continues the RNG stream and carries forces
    t2 = e.run_nvt(10, dt, 1.2, tau)
    ox2, ov2, of2, oimg2, ot2 = orc.run(orc.NVT, ox, ov, of, oimg, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, (), dt, 10,
                                        ktemp=1.2, tau=tau, seed=99, rng_step0=nsteps)
    assert np.allclose(t2[:, :3], ot2[:, :3], rtol=1e-9)
    e.close()


def test_bussi_hooks(md, orc):
    e = md.Engine(3, 1024, 10.0, 1.5, 0, seed=4242)
    for nf in (3069.0, 3068.0, 2.0, 3.0, 1.0, 3145725.0):
        for step in (0, 1, 12345678901):
            r1, r2 = e.bussi_noises(step, nf)
            o1, o2 = orc.bussi_noises(4242, step, nf)
            assert relerr(r1, o1) < 1e-12 and relerr(r2, o2) < 1e-12, (nf, step, r1, o1, r2, o2)
    rng = np.random.default_rng(0)
    for _ in range(20):
        ke, kt = rng.uniform(500, 5000), rng.uniform(0.5, 2.0)
        r1, r2 = rng.standard_normal(), rng.chisquare(3068)
        a = e.bussi_scale_from(ke, kt, 3069.0, 1e-3, 0.1, r1, r2)
        assert relerr(a, orc.bussi_scale(ke, kt, 3069.0, 1e-3, 0.1, r1, r2)) < 1e-14
    e.close()


def test_brownian_matches_oracle(md, orc):
    from mdjl_b200 import workloads
    cfg = workloads.phs_fluid(1024)
    e = md.Engine(3, 1024, cfg["box"], 1.5, 0, seed=31337)
    e.upload(cfg["x"], cfg["diam"])
    dt, nsteps, kt = 1e-5, 30, 1.4737
    t = e.run_brownian(nsteps, dt, kt)
    x, _, f, img = e.download()
    ox, _, of, oimg, ot = orc.run(orc.BROWNIAN, cfg["x"], None, np.zeros_like(cfg["x"]), np.zeros((1024, 3), np.int32),
                                  cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, (), dt, nsteps, ktemp=kt, seed=31337)
    assert np.array_equal(img, oimg)
    assert np.max(np.abs(x - ox)) < 1e-10
    assert np.array_equal(t[:, 3], ot[:, 3]) and np.allclose(t[:, :2], ot[:, :2], rtol=1e-10)
    e.close()


def test_nve_energy_drift_no_worse_than_oracle(md, orc):
    """E = U + KE over a short NVE horizon from an NVT-melted state."""
    cfg, v0, e = _setup(md, mode="list")
    dt = 1e-3
    e.run_nvt(2000, dt, 1.4737, 0.1, thermo=False)
    x, v, f, img = e.download()
    t = e.run_nve(2000, dt)
    ox, ov, of, oimg, ot = orc.run(orc.NVE, x, v, f, img, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS, (), dt, 2000)
    Eg, Eo = t[:, 0] + t[:, 2], ot[:, 0] + ot[:, 2]
    drift_g = np.max(np.abs(Eg - Eg[0])) / abs(Eg[0])
    drift_o = np.max(np.abs(Eo - Eo[0])) / abs(Eo[0])
    assert drift_g <= 1.05 * drift_o + 1e-12, (drift_g, drift_o)
    assert drift_g < 5e-3
    e.close()


def test_momentum_and_temperature(md, orc):
    cfg, v0, e = _setup(md, n=8192)
    t = e.run_nvt(3000, 1e-3, 1.4737, 0.1)
    _, v, _, _ = e.download()
    assert np.max(np.abs(v.sum(axis=0))) < 1e-8          # velocity Verlet + uniform rescale conserve total momentum
    nf = 3 * (8192 - 1.0)
    T = 2 * t[1000:, 2] / nf
    assert abs(T.mean() - 1.4737) < 0.02, T.mean()      # Bussi thermostat holds the target temperature
    assert relerr(orc.kinetic(v), t[-1, 2]) < 1e-12
    e.close()


def test_fire_minimizer_matches_oracle(md, orc):
    """mdb_fire_minimize vs the statement-by-statement restatement of fire_minimize! (src/minimize.jl:31-135)"""
    from mdjl_b200 import workloads
    p = workloads.poly2d(1200)
    kw = dict(max_steps=150, tol=1e-6, dt_initial=1e-4, dt_max=1e-2, alpha0=0.1, f_inc=1.2, f_dec=0.2, n_min=5)
    e = md.Engine(2, 1200, p["box"], 1.5, md._capi.POT_POLY, (1.25, 0.2), seed=1)
    v_user = workloads.velocities(1200, 2, 0.11)
    e.upload(p["x"], p["diam"], velocities=v_user)
    energy, frms, steps, conv = e.fire_minimize(**kw)
    x, v, f, img = e.download()
    ox, oimg, oE, osteps, oconv, trace = orc.fire(p["x"], np.zeros((1200, 2), np.int32), p["diam"], p["box"], 1.5, orc.POT_POLY,
                                                  (1.25, 0.2), **kw)
    assert conv == oconv is False and steps == osteps == 150
    assert relerr(energy, oE) < 1e-9 and energy < 0.2 * trace[0, 0]
    assert np.array_equal(img, oimg) and np.max(np.abs(x - ox)) < 1e-9
    assert np.array_equal(v, v_user)                       # the caller's velocities are untouched (the reference uses a local v)
    # a configuration already at a force-free state converges at the first test
    g = np.stack(np.meshgrid(np.arange(20), np.arange(20), indexing="ij"), -1).reshape(-1, 2) * 2.0 + 1.0
    e2 = md.Engine(2, 400, 40.0, 1.5, md._capi.POT_LJ, (1.0, 1.5), seed=1)
    e2.upload(g.astype(float), np.ones(400))
    energy, frms, steps, conv = e2.fire_minimize()
    assert conv and steps == 1 and energy == 0.0
    e.close(); e2.close()


@pytest.mark.parametrize("ensemble", ["nve", "nvt", "brownian"])
def test_small_system_persistent_kernel_matches_oracle(md, orc, ensemble):
    """K0-small (one persistent CTA, thousands of steps per launch) against the oracle loop and the list-mode engine"""
    g = dict(np.load(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "c1_phs_n1024.npz")))
    n = 1024
    out = {}
    for mode in ("small", "list"):
        modes = {"small": md._capi.MODE_SMALL, "list": md._capi.MODE_LIST}
        e = md.Engine(3, n, g["box"], 1.5, 0, seed=99, mode=modes[mode])
        e.upload(g["x"], g["diam"], velocities=g["v"], forces=g["f"], images=g["img"])
        if ensemble == "nve":
            t = e.run_nve(120, 1e-3)
        elif ensemble == "nvt":
            t = e.run_nvt(120, 1e-3, np.linspace(1.4737, 1.3, 120), 0.1)
        else:
            t = e.run_brownian(120, 1e-5, 1.4737)
        out[mode] = (t, e.download(), e.stats(), e.rng_step)
        e.close()
    assert out["small"][2]["mode"] == 3 and out["list"][2]["mode"] == 2
    ens = {"nve": orc.NVE, "nvt": orc.NVT, "brownian": orc.BROWNIAN}[ensemble]
    kw = dict(ktemp=np.linspace(1.4737, 1.3, 120), tau=0.1) if ensemble == "nvt" else (dict(ktemp=1.4737) if ensemble == "brownian" else {})
    dt = 1e-5 if ensemble == "brownian" else 1e-3
    ox, ov, of, oi, ot = orc.run(ens, g["x"], g["v"], g["f"], g["img"], g["diam"], g["box"], 1.5, orc.POT_PHS, (), dt, 120, seed=99, **kw)
    t, (x, v, f, img), st, rs = out["small"]
    assert rs == 120 and np.array_equal(img, oi)
    assert np.array_equal(t[:, 3], ot[:, 3]) and np.allclose(t[:, :3], ot[:, :3], rtol=1e-9, atol=1e-12)
    assert np.max(np.abs(x - ox)) < 1e-9
    if ensemble != "brownian":
        assert np.max(np.abs(v - ov)) < 1e-8
    assert np.array_equal(t[:, 3], out["list"][0][:, 3])
    # the large-system entry points keep working on the state the small kernel left behind
    e = md.Engine(3, n, g["box"], 1.5, 0, seed=99, mode=md._capi.MODE_SMALL)
    e.upload(g["x"], g["diam"], velocities=g["v"])
    e.run_nve(10, 1e-3)
    E, W, npairs = e.compute_forces()
    xs = e.download()[0]
    ref = orc.forces(xs, g["diam"], g["box"], 1.5, orc.POT_PHS)
    assert npairs == ref["n_int"] and relerr(E, ref["E"]) < 1e-12
    e.close()


@pytest.mark.parametrize("n,lpp", [(300, 1), (1024, 1), (1200, 1), (2000, 1), (300, 8), (1024, 4), (1200, 2), (1200, 4), (640, 8), (777, 4), (2048, 2)])
def test_small_system_cluster_kernel_equals_cooperative_kernel(md, monkeypatch, n, lpp):
    """K0-small as one thread-block cluster (cluster barriers, positions pushed through distributed shared memory, lpp lanes
    per particle) against the cooperative-grid version of the same loop.  With one lane per particle the per-particle
    arithmetic is the same statement by statement, so NVE and Brownian trajectories are bit-identical (only sums that
    never feed back are folded in another order: thermo rows to 1e-13); NVT feeds the kinetic-energy sum back through
    the thermostat: a few ulps per step.  With several lanes per particle the partial forces are added in another order:
    pair counts stay exact, everything else agrees to rounding error carried through the run."""
    from mdjl_b200 import workloads
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, 1.4737)
    out = {}
    for name, flag in (("cluster", "1"), ("grid", "0")):
        monkeypatch.setenv("MDB200_SMALL_CLUSTER", flag)
        monkeypatch.setenv("MDB200_SMALL_LPP", str(lpp))
        e = md.Engine(3, n, cfg["box"], 1.5, 0, seed=5, mode=md._capi.MODE_SMALL)
        e.upload(cfg["x"], cfg["diam"], velocities=v0)
        assert e.stats()["mode"] == md._capi.MODE_SMALL
        t1 = e.run_nve(400, 1e-3)
        s1 = e.download()
        t2 = e.run_brownian(150, 1e-5, 1.4737)
        s2 = e.download()
        t3 = e.run_nvt(60, 1e-3, 1.4737, 0.1)
        s3 = e.download()
        out[name] = (t1, s1, t2, s2, t3, s3, e.stats()["rebuilds"])
        e.close()
    a, b = out["cluster"], out["grid"]
    assert a[6] == b[6] and a[6] >= 2
    for k in (0, 2, 4):
        assert np.array_equal(a[k][:, 3], b[k][:, 3])                      # pair counts, every step
    if lpp == 1:
        for k in (1, 3):
            for u, w in zip(a[k], b[k]):
                assert np.array_equal(u, w)
        for k in (0, 2):
            assert np.allclose(a[k][:, :3], b[k][:, :3], rtol=1e-13, atol=1e-13)
    else:
        for k in (1, 3):
            assert np.array_equal(a[k][3], b[k][3])                        # image counters
            assert np.max(np.abs(a[k][0] - b[k][0])) < 1e-10 and np.max(np.abs(a[k][1] - b[k][1])) < 1e-9
            assert np.max(np.abs(a[k][2] - b[k][2])) < 1e-8 * max(1.0, np.max(np.abs(b[k][2])))
        for k in (0, 2):
            assert np.allclose(a[k][:, :3], b[k][:, :3], rtol=1e-10, atol=1e-12)
    assert np.allclose(a[4][:, :3], b[4][:, :3], rtol=1e-10)
    assert np.max(np.abs(a[5][0] - b[5][0])) < 1e-10 and np.max(np.abs(a[5][1] - b[5][1])) < 1e-9


def test_small_system_2d_polydisperse(md, orc):
    from mdjl_b200 import workloads
    p = workloads.poly2d(1200)
    e = md.Engine(2, 1200, p["box"], 1.5, md._capi.POT_POLY, (1.25, 0.2), seed=3)
    e.upload(p["x"], p["diam"], velocities=np.zeros((1200, 2)))
    e.fire_minimize(max_steps=400, tol=1e-3, dt_initial=1e-4, dt_max=5e-3)
    x0 = e.download()[0]
    v0 = workloads.velocities(1200, 2, 0.11)
    e.upload(x0, p["diam"], velocities=v0)
    assert e.stats()["mode"] == 3
    t = e.run_nve(60, 1e-3)
    x, v, f, img = e.download()
    ox, ov, of, oi, ot = orc.run(orc.NVE, x0, v0, np.zeros_like(x0), np.zeros((1200, 2), np.int32), p["diam"], p["box"], 1.5, orc.POT_POLY,
                                 (1.25, 0.2), 1e-3, 60)
    assert np.array_equal(t[:, 3], ot[:, 3]) and np.allclose(t[:, :3], ot[:, :3], rtol=1e-9)
    assert np.max(np.abs(x - ox)) < 1e-9 and np.array_equal(img, oi)
    e.close()


def test_mixed_ensembles_keep_the_list_exact(md, orc):
    """NVE, Brownian and stand-alone force calls interleaved on one engine: the displacement bookkeeping (running bound
    for velocity Verlet, exact displacement for Brownian moves) must never let a pair slip out of the Verlet list"""
    from mdjl_b200 import workloads
    n = 8192
    cfg = workloads.phs_fluid(n)
    e = md.Engine(3, n, cfg["box"], 1.5, 0, seed=12, mode=md._capi.MODE_LIST)
    e.upload(cfg["x"], cfg["diam"], velocities=workloads.velocities(n, 3, 1.4737))
    e.run_nvt(1200, 1e-3, 1.4737, 0.1, thermo=False)
    for rounds in range(6):
        e.run_brownian(37, 2e-5, 1.4737, thermo=False)
        e.run_nve(23, 1e-3, thermo=False)
        t = e.run_brownian(11, 2e-5, 1.4737)
        x = e.download()[0]
        E, W, npairs = e.compute_forces()
        ref = orc.forces(x, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS)
        assert npairs == ref["n_int"] and relerr(E, ref["E"]) <= 1e-12
    st = e.stats()
    assert st["rebuilds"] < 6 * 71      # far fewer rebuilds than steps
    e.close()
