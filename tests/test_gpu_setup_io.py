"""The steps either side of the hot path, through the C ABI: device-packed trajectory frames
with the library's background LAMMPS writer (src/io.jl:62-170), initialize_velocities on the device
(src/initialization.jl:32-47) against the oracle, and the exact binary checkpoint."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _engine(md, n=4096, mode="list", seed=11, **kw):
    from mdjl_b200 import workloads
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, 1.4737)
    modes = {"auto": md._capi.MODE_AUTO, "cells": md._capi.MODE_CELLS, "list": md._capi.MODE_LIST}
    e = md.Engine(3, n, cfg["box"], 1.5, md._capi.POT_PSEUDOHS, seed=seed, mode=modes[mode], **kw)
    e.upload(cfg["x"], cfg["diam"], velocities=v0)
    return cfg, v0, e


def test_frame_is_download_plus_unwrap(md):
    cfg, v0, e = _engine(md)
    e.run_nve(400, 2e-3, thermo=False)            # long enough for particles to cross the box faces
    e.frame_capture(0)
    e.run_nve(50, 2e-3, thermo=False)             # the step loop goes on while frame 0 travels
    e.frame_capture(1)
    fr1 = e.frame_wait(1).copy()
    x, v, f, img = e.download()
    L = np.asarray(cfg["box"], dtype=np.float64) * np.ones(3)
    assert np.array_equal(fr1[:, 0], cfg["diam"] / 2.0)
    assert np.array_equal(fr1[:, 1:4], x)
    assert np.array_equal(fr1[:, 4:7], x + L * img)
    assert np.any(img != 0)
    fr0 = e.frame_wait(0)
    assert not np.array_equal(fr0[:, 1:4], x)      # slot 0 still holds the earlier frame
    e.close()


@pytest.mark.parametrize("dim", [3, 2])
def test_lammps_writer_matches_reference_layout(md, tmp_path, dim):
    """the library's writer thread produces the bytes of write_to_file_lammps (src/io.jl:78-170) for the same state"""
    from mdjl_b200 import api, workloads
    if dim == 3:
        cfg, v0, e = _engine(md, n=3000, mode="auto")
        box = np.asarray(cfg["box"], dtype=np.float64) * np.ones(3)
        diam = cfg["diam"]
    else:
        p = workloads.poly2d(1200)
        e = md.Engine(2, 1200, p["box"], 1.5, md._capi.POT_POLY, (1.25, 0.2), seed=3)
        e.upload(p["x"], p["diam"], velocities=workloads.velocities(1200, 2, 0.11))
        e.fire_minimize(max_steps=200, tol=1e-3, dt_initial=1e-4, dt_max=5e-3)
        e.set_velocities(workloads.velocities(1200, 2, 0.11))
        box = np.asarray(p["box"], dtype=np.float64) * np.ones(2)
        diam = p["diam"]
    ours, ref = str(tmp_path / "ours.lammpstrj"), str(tmp_path / "ref.lammpstrj")
    for k, step in enumerate((0, 25, 50)):
        if step:
            e.run_nve(25, 1e-3, thermo=False)
        e.frame_capture(k % 2)
        e.frame_write_lammps(k % 2, ours, step, append=True)
        x, _, _, img = e.download(velocities=False, forces=False)
        api.write_to_file_lammps(ref, step, np.diag(box), x.shape[0], x, img, diam, dim, mode="a")
    e.frame_flush()
    with open(ours, "rb") as a, open(ref, "rb") as b:
        assert a.read() == b.read()
    # mode="w" (snapshot files, src/simulation.jl:152-166) overwrites
    e.frame_capture(0)
    e.frame_write_lammps(0, ours, 50, append=False)
    e.frame_flush()
    assert open(ours).read().count("ITEM: TIMESTEP") == 1
    with pytest.raises(md.MdbError) as ei:
        e.frame_capture(1)
        e.frame_write_lammps(1, str(tmp_path / "no_such_dir" / "f"), 1)
        e.frame_flush()
    assert ei.value.code == md._capi.ERR_IO
    e.close()


@pytest.mark.parametrize("dim,n", [(3, 50000), (2, 1200)])
def test_init_velocities_matches_oracle(md, orc, dim, n):
    from mdjl_b200 import workloads
    rng = np.random.default_rng(5)
    box = (n / 0.5) ** (1.0 / dim)
    x = rng.uniform(0, box, (n, dim))
    e = md.Engine(dim, n, box, 1.5, md._capi.POT_LJ, (1.0, 1.2), seed=20261018)
    e.upload(x, np.full(n, 0.1))
    e.init_velocities(1.4737, stream=3)
    v = e.download()[1]
    ref = orc.init_velocities(dim, n, 1.4737, 20261018, 3)
    assert np.max(np.abs(v - ref)) <= 1e-12 * np.max(np.abs(ref))
    assert np.max(np.abs(v.sum(axis=0))) < 1e-9                        # centre of mass at rest
    assert abs(np.sum(v * v) / ((n - 1) * dim) - 1.4737) < 1e-12      # exactly the requested temperature
    e.init_velocities(1.4737, stream=4)                                # another stream, other numbers
    assert not np.allclose(e.download()[1], v)
    e.run_nve(5, 1e-4)                                                 # velocities count as set
    e.close()


@pytest.mark.parametrize("n,ensemble", [(4096, "nve"), (4096, "nvt"), (1024, "nvt"), (4096, "brownian")])
def test_checkpoint_restart_is_bit_identical(md, tmp_path, n, ensemble):
    def run(e, k):
        if ensemble == "nve":
            return e.run_nve(k, 1e-3)
        if ensemble == "nvt":
            return e.run_nvt(k, 1e-3, 1.4737, 0.1)
        return e.run_brownian(k, 1e-5, 1.4737)
    path = str(tmp_path / "state.ckpt")
    cfg, v0, a = _engine(md, n=n, mode="auto")
    run(a, 130)
    a.checkpoint_save(path)
    ta = run(a, 90)
    sa = a.download()
    _, _, b = _engine(md, n=n, mode="auto")        # same configuration, unrelated state
    b.checkpoint_load(path)
    tb = run(b, 90)
    sb = b.download()
    assert np.array_equal(ta, tb)
    for p, q in zip(sa, sb):
        assert np.array_equal(p, q)
    assert a.rng_step == b.rng_step
    a.close()
    b.close()


def test_checkpoint_errors(md, tmp_path):
    cfg, v0, e = _engine(md, n=2048)
    path = str(tmp_path / "s.ckpt")
    e.checkpoint_save(path)
    with pytest.raises(md.MdbError) as ei:
        e.checkpoint_load(str(tmp_path / "missing.ckpt"))
    assert ei.value.code == md._capi.ERR_IO
    with open(str(tmp_path / "junk.ckpt"), "wb") as fh:
        fh.write(b"not a checkpoint" * 20)
    with pytest.raises(md.MdbError) as ei:
        e.checkpoint_load(str(tmp_path / "junk.ckpt"))
    assert ei.value.code == md._capi.ERR_IO
    other = md.Engine(3, 1000, 12.0, 1.5, md._capi.POT_PSEUDOHS)
    with pytest.raises(md.MdbError) as ei:
        other.checkpoint_load(path)
    assert ei.value.code == md._capi.ERR_INVALID_ARG
    other.close()
    e.close()


@pytest.mark.parametrize("dim", [3, 2])
def test_random_positions_match_oracle(md, orc, dim):
    n, box = 5000, (9.0, 11.0, 13.0)[:dim]
    e = md.Engine(dim, n, np.array(box), 1.0, md._capi.POT_SOFT, (1.0, 1.0), seed=77)
    e.upload(np.zeros((n, dim)), np.ones(n))
    e.random_positions(stream=9)
    x, _, _, img = e.download(velocities=False, forces=False)
    assert np.array_equal(x, orc.random_positions(dim, n, np.array(box + (1.0,) * (3 - dim)), 77, 9))   # same bits
    assert np.all(x >= 0) and np.all(x < np.array(box)) and not np.any(img)
    e.close()


def test_soft_penalty_forces_match_oracle(md, orc):
    from conftest import force_error, relerr
    n, box = 4000, 14.0
    x = orc.random_positions(3, n, box, 3, 0)
    diam = np.ones(n)
    e = md.Engine(3, n, box, 1.2, md._capi.POT_SOFT, (2.0, 1.2), seed=1)
    e.upload(x, diam)
    E, W, npairs = e.compute_forces()
    F = e.download()[2]
    ref = orc.forces(x, diam, np.full(3, box), 1.2, orc.POT_SOFT, (2.0, 1.2))
    assert npairs == ref["n_int"] > 1000
    assert relerr(E, ref["E"]) <= 1e-12 and relerr(W, ref["W"]) <= 1e-12 and force_error(F, ref["F"]) <= 1e-12
    e.close()


@pytest.mark.parametrize("dim,n,rho", [(3, 4096, 0.8976338790382897), (3, 1024, 0.8976338790382897), (2, 1200, 0.6)])
def test_initialize_random_removes_every_overlap(md, orc, dim, n, rho):
    """initialize_random (src/initialization.jl:20-30): random points, then no pair closer than tol -- what Packmol's
    pack_monoatomic! delivers in the reference -- checked by an independent O(N^2)-free recount in the oracle"""
    box = (n / rho) ** (1.0 / dim)
    x = md.initialize_random(box, n, np.random.default_rng(4), dim, tol=1.0)
    assert x.shape == (n, dim) and np.all(x >= 0) and np.all(x < box)
    ref = orc.forces(x, np.ones(n), np.full(3, box), 1.0, orc.POT_SOFT, (1.0, 1.0))
    assert ref["n_int"] == 0                                             # nobody within tol of anybody
    # and it is a disordered configuration, not a lattice: nearest-neighbour distances are spread out
    ref2 = orc.forces(x, np.ones(n), np.full(3, box), 1.3, orc.POT_SOFT, (1.0, 1.3))
    assert ref2["n_int"] > n // 2
