"""The reference-facing API mirror end to end on the GPU: the README example flow (README.md:11-66) through
Parameters / initialize_state / initialize_velocities / run_simulation with the files the reference writes."""
import os
import re

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_readme_example_flow(md, orc, tmp_path):
    packing_fraction, ktemp, n, dt = 0.47, 1.4737, 2 ** 10, 0.001
    density = 6.0 * packing_fraction / np.pi
    params = md.Parameters(density, n, dt, md.PseudoHS())
    path = str(tmp_path)
    rng = np.random.default_rng(7)
    state = md.initialize_state(params, path, random_init=True, rng=rng, seed=11)
    assert os.path.isfile(os.path.join(path, "init.xyz")) and state.velocities.shape == (0, 3)
    with pytest.raises(md.MdbError):                       # velocities must be set explicitly (README.md:38-41)
        md.run_simulation(state, params, md.NVE(), 10, 5, path)
    state.velocities = md.initialize_velocities(md.initial_temperature_for_velocities(ktemp), rng, n, 3)
    t1 = md.run_simulation(state, params, md.NVT(ktemp, 100.0 * dt), 600, 100, path)
    rows = open(os.path.join(path, "thermo.txt")).read().splitlines()
    assert rows[0] == "# Step Energy Temperature Pressure" and len(rows) == 1 + 6
    assert all(re.fullmatch(r"\d+ -?\d+\.\d{6} -?\d+\.\d{6} -?\d+\.\d{6}", r) for r in rows[1:])
    assert [int(r.split()[0]) for r in rows[1:]] == [0, 100, 200, 300, 400, 500]
    # the row at step s holds the post-update values of step s: U/N, T = 2 KE / nf, P = W/(3V) + rho T
    vol = n / density
    for r in rows[1:]:
        s_, e_, T_, P_ = r.split()
        U, W, KE, _ = t1[int(s_)]
        assert abs(float(e_) - U / n) < 1e-6 and abs(float(T_) - 2 * KE / state.nf) < 1e-6
        assert abs(float(P_) - (W / (3 * vol) + density * 2 * KE / state.nf)) < 1e-6
    traj = open(os.path.join(path, "trajectory.xyz")).read()
    assert traj.count("ITEM: TIMESTEP") == 6 and "ITEM: ATOMS id type radius x y z xu yu zu" in traj
    assert os.path.isfile(os.path.join(path, "final.xyz"))
    # chained NVE production run: state (forces included) carries over, RNG step continues
    x_before = state.system.positions
    t2 = md.run_simulation(state, params, md.NVE(), 300, 100, path, traj_name="production.xyz", thermo_name="production_thermo.txt")
    assert open(os.path.join(path, "production_thermo.txt")).read().count("\n") == 1 + 3
    E = t2[:, 0] + t2[:, 2]
    assert (E.max() - E.min()) / abs(E[0]) < 1e-3 and abs(2 * t2[-1, 2] / state.nf - ktemp) < 0.2
    # ... and equals the oracle continuing from the same state
    eng = state.system.engine
    assert eng.rng_step == 900 and not np.array_equal(x_before, state.system.positions)
    assert state.images.dtype == np.int32 and state.system.energy_and_forces.forces.shape == (n, 3)
    cell2, pos2, diam2 = md.read_file(os.path.join(path, "final.xyz"))
    assert np.allclose(np.diag(cell2), np.diag(state.unitcell)) and np.allclose(pos2, state.system.positions, atol=1e-6)


def test_temperature_ramp_and_brownian_run(md, tmp_path):
    from mdjl_b200 import workloads
    n = 4096
    cfg = workloads.phs_fluid(n)
    params = md.Parameters(cfg["rho"], n, 1e-3, md.PseudoHS())
    state = md.initialize_state(params, str(tmp_path), positions=cfg["x"], diameters=cfg["diam"], unitcell=cfg["box"], seed=5,
                                rng=np.random.default_rng(1))
    state.velocities = workloads.velocities(n, 3, 2.0)
    ramp = md.LinearRamp(2.0, 1.0, 1500)
    t = md.run_simulation(state, params, md.NVT(ramp, 0.05), 2000, 500, str(tmp_path), write_trajectory=False)
    T = 2 * t[:, 2] / state.nf
    # the thermostat follows ktemp(step+1) (with lag tau; the lattice start first converts kinetic into potential energy)
    assert abs(T[0] - 2.0) < 0.05 and abs(T[700:800].mean() - ramp(750)) < 0.08 and abs(T[-300:].mean() - 1.0) < 0.05
    p2 = md.Parameters(cfg["rho"], n, 1e-5, md.PseudoHS())
    tb = md.run_simulation(state, p2, md.Brownian(1.0), 200, 50, str(tmp_path), thermo_name="bd.txt", write_trajectory=False)
    rows = open(os.path.join(str(tmp_path), "bd.txt")).read().splitlines()[1:]
    assert len(rows) == 4 and all(float(r.split()[2]) == 1.0 for r in rows) and np.all(tb[:, 2] == 0)
    # pressure column averages the virial sampled every 10 steps (src/simulation.jl:253-266)
    vol = n / cfg["rho"]
    w = [tb[s, 1] for s in range(0, 1) if s % 10 == 0]
    assert abs(float(rows[0].split()[3]) - (np.mean(w) / (3 * vol) + cfg["rho"] * 1.0)) < 1e-6


def test_bussi_canonical_distribution_on_device(md):
    """KE under the device thermostat samples Gamma(nf/2, kT) (mean nf kT/2, variance nf kT^2/2) for a small system"""
    from mdjl_b200 import workloads
    n, kt = 64, 1.3
    L = 30.0   # dilute: essentially an ideal gas, the thermostat alone shapes the distribution
    rng = np.random.default_rng(2)
    g = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3) * 7.0 + 1.0
    e = md.Engine(3, n, L, 1.5, 0, seed=21)
    e.upload(g.astype(float), np.ones(n), velocities=workloads.velocities(n, 3, kt))
    t = e.run_nvt(40000, 1e-3, kt, 0.01)
    nf = 3 * (n - 1.0)
    ke = t[2000:, 2]
    assert abs(ke.mean() - 0.5 * nf * kt) < 0.02 * 0.5 * nf * kt
    assert abs(ke.var() - 0.5 * nf * kt * kt) < 0.12 * 0.5 * nf * kt * kt
    e.close()
