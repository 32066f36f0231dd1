"""x-slab decomposition on ONE GPU: P slab engines in an in-process ring (device copies instead of NCCL,
same kernels, same protocol) must reproduce the single-domain engine: identical pair counts every step, energies and
trajectories to rounding, ownership a partition of the particles, migration and ghost exchange exercised."""
import numpy as np
import pytest

from conftest import force_error, relerr

pytestmark = pytest.mark.gpu

# every test of this module runs over both in-process transports: 0 = device copies (stand-in for ncclSend/ncclRecv),
# 2 = the peer-memory kernels of the multi-GPU path (mailboxes + epoch flags, the step replayed as one CUDA graph)
TRANSPORTS = [0, 2]


@pytest.fixture(params=TRANSPORTS, ids=["copies", "peer"])
def tr(request):
    return request.param


def _cfg(md, n=8192, melt=1500):
    from mdjl_b200 import workloads
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, 1.4737)
    e = md.Engine(3, n, cfg["box"], 1.5, 0, seed=5)
    e.upload(cfg["x"], cfg["diam"], velocities=v0)
    e.run_nvt(melt, 1e-3, 1.4737, 0.1, thermo=False)
    x, v, f, img = e.download()
    e.close()
    return cfg, x, v, f, img


@pytest.mark.parametrize("nranks", [2, 3, 4])
def test_forces_match_single_domain_and_oracle(md, orc, nranks, tr):
    cfg, x, v, f, img = _cfg(md)
    n = x.shape[0]
    ring = md.SlabRing.local(nranks, 3, n, cfg["box"], 1.5, 0, seed=5, slab_transport=tr)
    ring.upload(x, cfg["diam"], velocities=v)
    E, W, npairs = ring.compute_forces()
    _, _, F, _ = ring.download()
    ref = orc.forces(x, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS)
    assert npairs == ref["n_int"] and relerr(E, ref["E"]) <= 1e-12 and relerr(W, ref["W"]) <= 1e-12
    assert force_error(F, ref["F"]) <= 1e-12
    st = ring.stats()
    assert sum(s["n_owned"] for s in st) == n and all(s["n_owned"] > 0 for s in st)
    ring.close()


@pytest.mark.parametrize("nranks,ensemble", [(2, "nve"), (4, "nve"), (3, "nvt"), (2, "brownian")])
def test_dynamics_match_single_domain(md, orc, nranks, ensemble, tr):
    cfg, x, v, f, img = _cfg(md)
    n = x.shape[0]
    single = md.Engine(3, n, cfg["box"], 1.5, 0, seed=77)
    single.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
    ring = md.SlabRing.local(nranks, 3, n, cfg["box"], 1.5, 0, seed=77, slab_transport=tr)
    ring.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
    nsteps = 150
    if ensemble == "nve":
        a, b = single.run_nve(nsteps, 1e-3), ring.run_nve(nsteps, 1e-3)
    elif ensemble == "nvt":
        a, b = single.run_nvt(nsteps, 1e-3, 1.4737, 0.1), ring.run_nvt(nsteps, 1e-3, 1.4737, 0.1)
    else:
        a, b = single.run_brownian(nsteps, 1e-5, 1.4737), ring.run_brownian(nsteps, 1e-5, 1.4737)
    assert np.array_equal(a[:, 3], b[:, 3])                     # interacting pairs, every step, exact
    assert np.allclose(a[:, :3], b[:, :3], rtol=1e-9, atol=1e-9)
    xs, vs, fs, ims = single.download()
    xr, vr, fr, imr = ring.download()
    assert np.array_equal(ims, imr)
    assert np.max(np.abs(xs - xr)) < 1e-9 and np.max(np.abs(vs - vr)) < 1e-8
    st = ring.stats()
    assert all(s["rebuilds"] >= 2 for s in st)
    single.close()
    ring.close()


def test_fused_slab_step_is_bit_identical(md, orc, tr):
    """slab NVE runs use the fused step too (the force kernel moves the owned particles for the next step before the
    ghost exchange): same bits as the reference kernel order, across rebuilds, migrations and several run calls"""
    cfg, x, v, f, img = _cfg(md)
    n = x.shape[0]
    out = []
    for no_fuse in (True, False):
        ring = md.SlabRing.local(3, 3, n, cfg["box"], 1.5, 0, seed=77, no_fuse=no_fuse, slab_transport=tr)
        ring.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
        rows = [ring.run_nve(k, 1e-3) for k in (1, 2, 120, 60)]
        out.append((np.concatenate(rows), ring.download(), ring.stats()))
        ring.close()
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)
    assert all(s["rebuilds"] >= 2 for s in out[1][2])


def test_upload_owned_round_trip(md, orc, tr):
    """a rank can hand back the rows it owns (download_owned -> upload_owned) instead of the global arrays: the run
    continues exactly like after a global re-upload of the same state (slab sorts are canonical in particle id)"""
    cfg, x, v, f, img = _cfg(md)
    n = x.shape[0]
    out = []
    for owned in (False, True):
        ring = md.SlabRing.local(3, 3, n, cfg["box"], 1.5, 0, seed=21, slab_transport=tr)
        ring.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
        ring.run_nve(80, 1e-3)
        if owned:
            parts = []
            for e in ring.engines:
                ids, xo, vo, fo, io = e.download_owned()
                parts.append((ids, xo, cfg["diam"][ids], vo, fo, io))
            ring.upload_owned(parts)
        else:
            X, V, F, I = ring.download()
            ring.upload(X, cfg["diam"], velocities=V, forces=F, images=I)
        t = ring.run_nve(120, 1e-3)
        out.append((t, ring.download()))
        assert sum(s["n_owned"] for s in ring.stats()) == n
        ring.close()
    assert np.array_equal(out[0][0], out[1][0])
    for a, b in zip(out[0][1], out[1][1]):
        assert np.array_equal(a, b)
    e = md.Engine(3, n, cfg["box"], 1.5, 0)
    e.upload(x, cfg["diam"])
    with pytest.raises(md.MdbError):          # single-domain handles use mdb_upload
        e.upload_owned(np.arange(4, dtype=np.int32), x[:4], cfg["diam"][:4])
    e.close()


def test_migration_over_long_run(md, orc, tr):
    """particles cross slab boundaries (and the periodic box face) during a longer run; ownership stays a partition,
    the pair count still matches an independent recount by the oracle at the end"""
    cfg, x, v, f, img = _cfg(md, n=4096)
    n = x.shape[0]
    ring = md.SlabRing.local(3, 3, n, cfg["box"], 1.5, 0, seed=9, slab_transport=tr)
    ring.upload(x, cfg["diam"], velocities=v * 1.5, forces=f, images=img)
    own0 = [s["n_owned"] for s in ring.stats()]
    ids0 = [e.download_owned()[0] for e in ring.engines]
    t = ring.run_nve(1500, 1e-3)
    ids1 = [e.download_owned()[0] for e in ring.engines]
    moved = sum(len(set(a.tolist()) ^ set(b.tolist())) for a, b in zip(ids0, ids1))
    assert moved > 0                                           # migration happened
    xr, vr, fr, imr = ring.download()                          # raises unless ownership is a partition
    ref = orc.forces(xr, cfg["diam"], cfg["box"], 1.5, orc.POT_PHS)
    # forces resident after the last step belong to the last positions
    assert int(t[-1, 3]) == ref["n_int"] and relerr(t[-1, 0], ref["E"]) <= 1e-11
    assert force_error(fr, ref["F"]) <= 1e-11
    assert np.max(np.abs(vr.sum(axis=0) - (v * 1.5).sum(axis=0))) < 1e-8
    ring.close()


def test_slab_errors(md):
    from mdjl_b200 import _capi
    e = md.Engine(3, 1000, 12.0, 1.5, 0, rank=0, nranks=2, use_graph=False)
    rng = np.random.default_rng(0)
    e.upload(rng.uniform(0, 12, (1000, 3)), np.ones(1000) * 0.3)
    with pytest.raises(md.MdbError) as ei:   # no communicator yet
        e.compute_forces()
    assert ei.value.code == _capi.ERR_STATE
    e.close()


def test_2d_polydisperse_slabs(md, orc, tr):
    """2-D non-additive mixture (BASELINE config 2 family) cut into slabs: forces vs oracle, dynamics vs single domain"""
    from mdjl_b200 import workloads
    p = workloads.poly2d(4900)
    e = md.Engine(2, 4900, p["box"], 1.5, md._capi.POT_POLY, (1.25, 0.2), seed=3, mode=md._capi.MODE_LIST)
    e.upload(p["x"], p["diam"], velocities=np.zeros((4900, 2)))
    e.fire_minimize(max_steps=300, tol=1e-3, dt_initial=1e-4, dt_max=5e-3)
    x0 = e.download()[0]
    v0 = workloads.velocities(4900, 2, 0.11)
    e.upload(x0, p["diam"], velocities=v0)
    ring = md.SlabRing.local(3, 2, 4900, p["box"], 1.5, md._capi.POT_POLY, (1.25, 0.2), seed=3, slab_transport=tr)
    ring.upload(x0, p["diam"], velocities=v0)
    E, W, npairs = ring.compute_forces()
    ref = orc.forces(x0, p["diam"], p["box"], 1.5, orc.POT_POLY, (1.25, 0.2))
    assert npairs == ref["n_int"] and relerr(E, ref["E"]) <= 1e-12
    assert force_error(ring.download()[2], ref["F"]) <= 1e-12
    # the stand-alone evaluation primed the resident forces of the ring; start both runs from the same (unprimed) state,
    # as the reference does not compute forces before step 0
    ring.upload(x0, p["diam"], velocities=v0)
    a, b = e.run_nve(200, 1e-3), ring.run_nve(200, 1e-3)
    assert np.array_equal(a[:, 3], b[:, 3]) and np.allclose(a[:, :3], b[:, :3], rtol=1e-9)
    assert np.max(np.abs(e.download()[0] - ring.download()[0])) < 1e-9
    e.close()
    ring.close()


@pytest.mark.parametrize("ensemble", ["nve", "nvt", "brownian"])
def test_peer_transport_is_bit_identical_to_copies(md, orc, ensemble):
    """the peer-memory transport (graph replay and eager launches) moves the same bytes as the copy transport: thermo rows,
    final state, ownership and rebuild counts agree bit for bit, across rebuilds, migrations and several run calls"""
    cfg, x, v, f, img = _cfg(md)
    n = x.shape[0]
    out = []
    for kw in (dict(slab_transport=0), dict(slab_transport=2), dict(slab_transport=2, use_graph=False)):
        ring = md.SlabRing.local(4, 3, n, cfg["box"], 1.5, 0, seed=31, **kw)
        ring.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
        rows = []
        for k in (1, 130, 2, 90):
            if ensemble == "nve":
                rows.append(ring.run_nve(k, 1e-3))
            elif ensemble == "nvt":
                rows.append(ring.run_nvt(k, 1e-3, 1.4737, 0.1))
            else:
                rows.append(ring.run_brownian(k, 1e-5, 1.4737))
        E, W, npairs = ring.compute_forces()
        st = ring.stats()
        out.append((np.concatenate(rows), ring.download(), [s["rebuilds"] for s in st], [s["n_owned"] for s in st], (E, W, npairs),
                    st[0]["slab_transport"], st[0]["slab_graph"]))
        ring.close()
    assert (out[0][5], out[1][5], out[2][5]) == (1, 3, 3) and (out[0][6], out[1][6], out[2][6]) == (0, 1, 0)
    for o in out[1:]:
        assert np.array_equal(out[0][0], o[0])
        for a, b in zip(out[0][1], o[1]):
            assert np.array_equal(a, b)
        assert out[0][2] == o[2] and out[0][3] == o[3] and out[0][4] == o[4]
    assert min(out[0][2]) >= 2


def test_peer_wait_times_out_instead_of_hanging(md, monkeypatch):
    """a neighbour that never sends must end the run with MDB_ERR_STATE after MDB200_PEER_TIMEOUT_S, not hang the GPU
    (fault injection: MDB200_PEER_TEST_MUTE_RANK makes one rank of the ring skip its head message)"""
    import time
    from mdjl_b200 import workloads, _capi
    monkeypatch.setenv("MDB200_PEER_TIMEOUT_S", "0.25")
    n = 4096
    cfg = workloads.phs_fluid(n)
    ring = md.SlabRing.local(2, 3, n, cfg["box"], 1.5, 0, seed=1, slab_transport=2)
    ring.upload(cfg["x"], cfg["diam"], velocities=workloads.velocities(n, 3, 1.0))
    assert np.all(np.isfinite(ring.run_nve(20, 1e-3)))       # healthy ring
    monkeypatch.setenv("MDB200_PEER_TEST_MUTE_RANK", "1")
    ring.upload(cfg["x"], cfg["diam"], velocities=workloads.velocities(n, 3, 1.0))
    t0 = time.perf_counter()
    with pytest.raises(md.MdbError) as ei:
        ring.run_nve(50, 1e-3)
    assert ei.value.code == _capi.ERR_STATE and "did not arrive" in str(ei.value)
    assert time.perf_counter() - t0 < 10.0                     # one timeout per waiting rank, then sticky
    ring.close()


def test_slab_init_velocities_match_single_domain(md, orc, tr):
    """initialize_velocities (src/initialization.jl:32-47) on a ring: every rank draws the normals of the particles it owns
    (keyed by particle id), the centre-of-mass and temperature sums are all-reduced; same velocities as one domain"""
    cfg, x, v, f, img = _cfg(md, n=4096, melt=200)
    n = x.shape[0]
    single = md.Engine(3, n, cfg["box"], 1.5, 0, seed=12)
    single.upload(x, cfg["diam"])
    single.init_velocities(1.3, stream=5)
    vs = single.download()[1]
    single.close()
    ring = md.SlabRing.local(3, 3, n, cfg["box"], 1.5, 0, seed=12, slab_transport=tr)
    ring.upload(x, cfg["diam"])
    ring.init_velocities(1.3, stream=5)
    vr = ring.download()[1]
    assert np.max(np.abs(vr - vs)) < 1e-13
    assert np.max(np.abs(vr.sum(axis=0))) < 1e-10 and abs((vr ** 2).sum() / (3 * (n - 1.0)) - 1.3) < 1e-12
    t = ring.run_nve(30, 1e-3)      # the ring steps with them (have_vel set on every member)
    assert np.all(np.isfinite(t))
    ring.close()


@pytest.mark.parametrize("ensemble", ["nve", "nvt"])
def test_slab_checkpoint_restart_is_bit_identical(md, orc, tr, tmp_path, ensemble):
    """one checkpoint file per slab (<path>.<rank>); a ring of FRESH handles restores from the files alone (no global
    arrays) and continues exactly like the run that was saved: thermo rows and final state bit for bit"""
    cfg, x, v, f, img = _cfg(md)
    n = x.shape[0]
    path = str(tmp_path / "ring.ckpt")

    def step(ring, k):
        return ring.run_nve(k, 1e-3) if ensemble == "nve" else ring.run_nvt(k, 1e-3, 1.4737, 0.1)

    a = md.SlabRing.local(3, 3, n, cfg["box"], 1.5, 0, seed=44, slab_transport=tr)
    a.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
    step(a, 70)
    a.checkpoint_save(path)
    ta = step(a, 110)
    fa = a.download()
    own_a = [s["n_owned"] for s in a.stats()]
    a.close()
    b = md.SlabRing.local(3, 3, n, cfg["box"], 1.5, 0, seed=44, slab_transport=tr)
    b.checkpoint_load(path)
    tb = step(b, 110)
    fb = b.download()
    assert np.array_equal(ta, tb)
    for p, q in zip(fa, fb):
        assert np.array_equal(p, q)
    assert own_a == [s["n_owned"] for s in b.stats()]
    with pytest.raises(md.MdbError):      # a file of another rank is refused
        import shutil
        shutil.copy(path + ".1", str(tmp_path / "swap.ckpt.0"))
        b.engines[0].checkpoint_load(str(tmp_path / "swap.ckpt"))
    b.close()


def test_slab_frames_one_lammps_file_per_rank(md, orc, tr, tmp_path):
    """trajectory frames of a ring: every slab writes <path>.<rank> with the atoms it owns (ids = original index + 1); the
    union of the files is the single-domain frame (radius, wrapped and unwrapped coordinates)"""
    cfg, x, v, f, img = _cfg(md, n=4096, melt=300)
    n = x.shape[0]
    ring = md.SlabRing.local(3, 3, n, cfg["box"], 1.5, 0, seed=2, slab_transport=tr)
    ring.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
    ring.run_nve(40, 1e-3)
    path = str(tmp_path / "ring.lammpstrj")
    ring.frame_capture(0)
    ring.frame_write_lammps(0, path, 40, append=False)
    ring.run_nve(10, 1e-3)              # the step loop goes on while the frame is written
    ring.frame_flush()
    xr, vr, fr, ir = ring.download()
    rows = {}
    for r in range(3):
        lines = open("%s.%d" % (path, r)).read().splitlines()
        assert lines[1] == "40" and lines[2] == "ITEM: NUMBER OF ATOMS"
        k = lines.index([l for l in lines if l.startswith("ITEM: ATOMS")][0])
        assert int(lines[3]) == len(lines) - k - 1
        for l in lines[k + 1:]:
            c = l.split()
            rows[int(c[0]) - 1] = [float(t) for t in c[2:]]
    assert sorted(rows) == list(range(n))
    # the frame was taken at step 40: compare with a single-domain engine stepped the same way
    single = md.Engine(3, n, cfg["box"], 1.5, 0, seed=2, mode=md._capi.MODE_LIST)
    single.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
    single.run_nve(40, 1e-3)
    single.frame_capture(0)
    fr0 = np.array(single.frame_wait(0))
    single.close()
    got = np.array([rows[i] for i in range(n)])
    assert np.max(np.abs(got - fr0)) < 2e-6      # "%lf" keeps six decimals
    ring.close()


@pytest.mark.parametrize("ensemble", ["nve", "nvt"])
def test_peer_graph_batching_is_bit_identical(md, orc, monkeypatch, ensemble):
    """MDB200_GRAPH_BATCH = B captures B consecutive slab steps (each with its own conditional rebuild node) into one
    graph; any split of a run into batched launches, single steps and the fused run's plain last step gives the same bits"""
    cfg, x, v, f, img = _cfg(md)
    n = x.shape[0]
    out = []
    for batch in ("1", "4", "7"):
        monkeypatch.setenv("MDB200_GRAPH_BATCH", batch)
        ring = md.SlabRing.local(3, 3, n, cfg["box"], 1.5, 0, seed=31, slab_transport=2)
        ring.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
        rows = []
        for k in (1, 3, 64, 29):
            rows.append(ring.run_nve(k, 1e-3) if ensemble == "nve" else ring.run_nvt(k, 1e-3, 1.4737, 0.1))
        out.append((np.concatenate(rows), ring.download(), [s["rebuilds"] for s in ring.stats()]))
        ring.close()
    for o in out[1:]:
        assert np.array_equal(out[0][0], o[0]) and out[0][2] == o[2]
        for a, b in zip(out[0][1], o[1]):
            assert np.array_equal(a, b)
    assert min(out[0][2]) >= 2


def test_lattice_start_fits_the_ghost_buffers(md, tr):
    """a lattice start puts whole lattice planes into single cell columns: a boundary column can hold ceil(w/a) planes where
    the mean is w/a (here 2 vs 1.33), which overflowed ghost buffers sized at 1.3x the mean column (round-2 finding: the
    N = 2^24 bench start failed on 2 GPUs the moment the default skin moved the grid)"""
    from mdjl_b200 import slabs
    m, a = 64, 1.0367
    n, L = m ** 3, m * a
    g = np.stack(np.meshgrid(*[np.arange(m)] * 3, indexing="ij"), -1).reshape(-1, 3) * a
    rng = np.random.default_rng(3)
    x = (g + 0.5 + rng.uniform(-0.015, 0.015, g.shape)) % L     # offset: two whole planes inside the boundary column at x ~ 16-17
    box = np.array([L, L, L])
    ring = md.SlabRing.local(4, 3, n, box, 1.5, 0, seed=1, slab_transport=tr)
    ring.upload(x, np.ones(n), velocities=rng.normal(0, 1.2, (n, 3)))
    st = ring.lead.stats()
    pl = slabs.plan(box, 3, st["r_search"], 0.3 * st["r_search"], 4)   # the engine's own grid (default skin)
    assert pl["nc"][0] == st["ncell"][0]
    col = slabs.column_of(x[:, 0], pl)
    pop = np.bincount(col, minlength=pl["nc"][0])
    boundary = sorted({c for lo, hi in pl["columns"] for c in (lo, hi - 1)})
    assert pop[boundary].max() > 1.3 * n / pl["nc"][0] + 1024      # the case that used to overflow is really present
    E, W, npairs = ring.compute_forces()
    t = ring.run_nve(30, 1e-3)
    assert np.all(np.isfinite(t)) and sum(s["n_owned"] for s in ring.stats()) == n
    ring.close()
