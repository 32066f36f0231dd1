"""Regression tests for defects found in review (round-1 ADVICE.md): each test names the failure it pins."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_small_path_brownian_long_run_equals_list_path(md):
    """K0-small runs Brownian steps with ONE grid barrier per step; the per-CTA partial sums are double-buffered by step
    parity so that a CTA that is a step ahead cannot overwrite what a slower CTA is still folding.  A stale fold would
    corrupt thermo rows and the displacement bound that drives list rebuilds (missed pairs).  Every thermo row of a long
    run must equal the multi-kernel list path bit for bit (same arithmetic, noise keyed by particle id and step)."""
    from mdjl_b200 import workloads
    n = 2048
    cfg = workloads.phs_fluid(n)
    rows = []
    for mode in (md._capi.MODE_SMALL, md._capi.MODE_LIST):
        e = md.Engine(3, n, cfg["box"], 1.5, md._capi.POT_PSEUDOHS, seed=17, mode=mode)
        e.upload(cfg["x"], cfg["diam"])
        t = np.concatenate([e.run_brownian(k, 2e-5, 1.4737) for k in (700, 1, 299)])
        x, _, f, img = e.download()
        rows.append((t, x, f, img, e.stats()))
        e.close()
    (ta, xa, fa, ia, sa), (tb, xb, fb, ib, sb) = rows
    assert sa["mode"] == md._capi.MODE_SMALL and sb["mode"] == md._capi.MODE_LIST
    assert np.array_equal(ta[:, 3], tb[:, 3])                 # interacting pairs, every step
    assert np.allclose(ta[:, :2], tb[:, :2], rtol=1e-9, atol=1e-9)
    assert np.array_equal(ia, ib) and np.max(np.abs(xa - xb)) < 1e-9


def test_reupload_with_wider_diameters_does_not_replay_a_stale_graph(md, orc):
    """mdb_upload on a handle whose buffers are unchanged used to keep the captured step graph, which bakes in the search
    radius of the OLD diameter range (Polydisperse: range = rcut*smax*(1+|eps|(smax-smin))).  The same handle re-uploaded
    with a wider range must step exactly like a fresh handle."""
    n, dim, L = 1600, 2, 40.0
    rng = np.random.default_rng(8)
    g = np.stack(np.meshgrid(np.arange(40), np.arange(40), indexing="ij"), -1).reshape(-1, 2).astype(float) + 0.5
    x = g + rng.uniform(-0.05, 0.05, g.shape)
    v = rng.normal(0, 0.3, (n, dim))
    d_narrow = rng.uniform(0.9, 1.0, n)
    d_wide = d_narrow.copy()
    d_wide[::7] = 1.28     # widens smax: r_search grows, the cell grid (floor) and the buffers stay
    pp = (1.25, 0.2)

    def run(e, diam):
        e.upload(x, diam, velocities=v)
        t = e.run_nve(60, 2e-3)
        return t, e.download()

    for mode in (md._capi.MODE_LIST, md._capi.MODE_AUTO):
        a = md.Engine(dim, n, L, 1.5, md._capi.POT_POLY, pp, seed=3, mode=mode)
        run(a, d_narrow)
        t1, s1 = run(a, d_wide)
        b = md.Engine(dim, n, L, 1.5, md._capi.POT_POLY, pp, seed=3, mode=mode)
        t2, s2 = run(b, d_wide)
        assert np.array_equal(t1, t2)
        for p, q in zip(s1, s2):
            assert np.array_equal(p, q)
        # and the pair set is the right one: forces of the final state against the oracle
        ref = orc.forces(s2[0], d_wide, np.array([L, L]), 1.5, orc.POT_POLY, pp)
        E, W, npairs = b.compute_forces()
        assert npairs == ref["n_int"]
        a.close()
        b.close()


def test_checkpoint_with_corrupted_ids_is_rejected(md, tmp_path):
    """the loader scatters through id[]: a file whose ids are not a permutation of 0..n-1 must fail with MDB_ERR_IO"""
    from mdjl_b200 import workloads
    n = 4096
    cfg = workloads.phs_fluid(n)
    e = md.Engine(3, n, cfg["box"], 1.5, md._capi.POT_PSEUDOHS, seed=4, mode=md._capi.MODE_LIST)
    e.upload(cfg["x"], cfg["diam"], velocities=workloads.velocities(n, 3, 1.0))
    e.run_nve(5, 1e-3)
    path = str(tmp_path / "state.ckpt")
    e.checkpoint_save(path)
    raw = bytearray(open(path, "rb").read())
    ids_off = len(raw) - 4 * n
    for bad, what in ((np.int32(n + 5), "out of range"), (np.frombuffer(bytes(raw[ids_off:ids_off + 4]), np.int32)[0], "duplicate")):
        broken = bytearray(raw)
        broken[ids_off + 4 * 10: ids_off + 4 * 11] = np.int32(bad).tobytes()
        p2 = str(tmp_path / ("broken_%s.ckpt" % what.replace(" ", "_")))
        open(p2, "wb").write(bytes(broken))
        with pytest.raises(md._capi.MdbError) as ei:
            e.checkpoint_load(p2)
        assert ei.value.code == md._capi.ERR_IO
    e.checkpoint_load(path)   # the intact file still loads
    e.close()


def test_brownian_thermo_row_without_a_virial_sample_is_nan_not_a_crash(md, tmp_path):
    """frequency = 5: output intervals without a step % 10 == 0 have nprom == 0; the reference writes NaN and goes on
    (src/simulation.jl:252-262)"""
    from mdjl_b200 import workloads
    n = 1024
    cfg = workloads.phs_fluid(n)
    params = md.Parameters(cfg["rho"], n, 1e-5, md.PseudoHS())
    state = md.initialize_state(params, str(tmp_path), positions=cfg["x"], diameters=cfg["diam"], unitcell=cfg["box"], seed=5,
                                rng=np.random.default_rng(1))
    md.run_simulation(state, params, md.Brownian(1.0), 30, 5, str(tmp_path), thermo_name="bd5.txt", write_trajectory=False)
    rows = open(os.path.join(str(tmp_path), "bd5.txt")).read().splitlines()[1:]
    p = [float(r.split()[3]) for r in rows]
    assert len(rows) == 6 and np.isnan(p[1]) and np.isfinite(p[0]) and np.isfinite(p[2])


def test_lammps_writer_survives_huge_coordinates(md, tmp_path):
    """a finite coordinate near 1e300 prints ~300 digits with %lf: the row formatter must not run past its buffer"""
    n = 64
    g = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3) * 2.0 + 0.5
    e = md.Engine(3, n, 8.0, 1.5, md._capi.POT_PSEUDOHS, seed=1)
    img = np.zeros((n, 3), np.int32)
    img[0] = (2**31 - 1, -(2**31), 5)
    e.upload(g.astype(float), np.ones(n), images=img)
    e.frame_capture(0)
    fr = e.frame_wait(0)
    fr[1, 0] = 1e300         # poke the pinned frame: what a blow-up would hand the writer
    fr[2, 4] = -1.7e308
    path = str(tmp_path / "huge.lammpstrj")
    e.frame_write_lammps(0, path, 0, append=False)
    e.frame_flush()
    lines = open(path).read().splitlines()
    atoms = lines[lines.index([l for l in lines if l.startswith("ITEM: ATOMS")][0]) + 1:]
    assert len(atoms) == n
    assert len(atoms[1].split()) == 9 and float(atoms[1].split()[2]) == 1e300 and float(atoms[2].split()[6]) == -1.7e308
    e.close()


@pytest.mark.parametrize("dim", [3, 2])
def test_f32_list_build_is_a_superset_with_identical_results(md, monkeypatch, dim):
    """the Verlet list is built from single-precision copies of the positions against a padded radius (k_build_list_f32):
    a superset of the FP64 list in the same order, so every force sum, thermo row and rebuild decision keeps its bits
    (MDB200_BUILD_F64=1 restores the FP64 build)"""
    from mdjl_b200 import workloads
    if dim == 3:
        n = 32768
        cfg = workloads.phs_fluid(n)
        tag, pp, dt, kt = md._capi.POT_PSEUDOHS, (), 1e-3, 1.4737
    else:
        n = 4900
        cfg = workloads.poly2d(n)
        tag, pp, dt, kt = md._capi.POT_POLY, (1.25, 0.2), 1e-3, 0.11
    v0 = workloads.velocities(n, dim, kt)
    out = []
    for f64, wc in ((True, "0"), (False, "0"), (False, "1")):   # FP64 build, FP32 build, FP32 warp-cooperative build
        if f64:
            monkeypatch.setenv("MDB200_BUILD_F64", "1")
        else:
            monkeypatch.delenv("MDB200_BUILD_F64", raising=False)
        monkeypatch.setenv("MDB200_BUILD_WC", wc)
        e = md.Engine(dim, n, cfg["box"], 1.5, tag, pp, seed=6, mode=md._capi.MODE_LIST)
        e.upload(cfg["x"], cfg["diam"], velocities=v0)
        if dim == 2:
            e.fire_minimize(max_steps=200, tol=1e-3, dt_initial=1e-4, dt_max=5e-3)
            e.set_velocities(v0)
        t = np.vstack([e.run_nvt(150, dt, kt, 100 * dt), e.run_nve(150, dt)])
        out.append((t, e.download(), e.stats()))
        e.close()
    for o in out[1:]:
        assert np.array_equal(out[0][0], o[0])
        for a, b in zip(out[0][1], o[1]):
            assert np.array_equal(a, b)
        assert out[0][2]["rebuilds"] == o[2]["rebuilds"] >= 3
        assert 0 <= o[2]["max_neighbors"] - out[0][2]["max_neighbors"] <= 2
    assert out[1][2]["max_neighbors"] == out[2][2]["max_neighbors"]      # the two FP32 builds produce the same list
