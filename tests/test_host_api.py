"""CPU tests of the host-side mirror of the reference interface (no device work)."""
import math
import os

import numpy as np
import pytest


def test_ramps(md):
    r = md.LinearRamp(2.0, 1.0, 11)          # src/temperature_ramps.jl:7-29 (1-indexed steps)
    assert r(1) == 2.0 and r(11) == 1.0 and r(12) == 1.0 and r(0) == 2.0 and abs(r(6) - 1.5) < 1e-15
    assert md.LinearRamp(2.0, 1.0, 1)(1) == 1.0
    e = md.ExponentialRamp(2.0, 0.5, 3)      # :36-60
    assert e(1) == 2.0 and abs(e(2) - 1.0) < 1e-15 and abs(e(3) - 0.5) < 1e-15 and e(99) == 0.5
    assert md.ExponentialRamp(1.0, 1.0, 5)(3) == 1.0
    assert md.initial_temperature_for_velocities(r) == 2.0 and md.initial_temperature_for_velocities(1.3) == 1.3


def test_ensembles_and_parameters(md):
    t = md.NVT(1.5, 0.1)
    assert t.ktemp(17) == 1.5 and t.tau == 0.1
    t2 = md.NVT(md.LinearRamp(2.0, 1.0, 11), 0.1)
    assert t2.ktemp(11) == 1.0
    assert md.Brownian(1.2).ktemp == 1.2
    p = md.Parameters(0.9, 1024, 1e-3, md.PseudoHS())
    assert (p.rho, p.n_particles, p.dt) == (0.9, 1024, 1e-3) and p.potential.tag == 0
    with pytest.raises(TypeError):
        md.Parameters(0.9, 10, 1e-3, "lj")

    class Mine(md.Potential):
        pass
    with pytest.raises(NotImplementedError):   # fallback `evaluate` errors, src/types.jl:4-6
        md.evaluate(Mine(), 1.0)
    lj = md.LennardJones(epsilon=1.0, sigma=1.0, r_cut=2.5)
    assert abs(lj.V_cut - 4 * ((1 / 2.5) ** 12 - (1 / 2.5) ** 6)) < 1e-16 and lj.params() == (1.0, 2.5)
    assert md.LennardJonesXPLOR(1.0, 1.0, 2.0, 2.5).params() == (1.0, 2.0, 2.5)
    assert md.Polydisperse().params() == (1.25, 0.2)


def test_long_range_corrections(md):
    lj = md.LennardJones(r_cut=2.5, tail_correction=True)
    n, vol = 1000, 1250.0
    rho = n / vol
    e = (8.0 * math.pi * rho / 3.0) * (((1 / 2.5) ** 9) / 3.0 - (1 / 2.5) ** 3) * n
    pr = (16.0 * math.pi * rho ** 2 / 3.0) * ((2.0 / 3.0) * (1 / 2.5) ** 9 - (1 / 2.5) ** 3)
    assert abs(md.energy_lrc(lj, n, vol) - e) < 1e-12 * abs(e) and abs(md.pressure_lrc(lj, n, vol) - pr) < 1e-12 * abs(pr)
    assert md.energy_lrc(md.LennardJones(), n, vol) == 0.0 and md.pressure_lrc(md.PseudoHS(), n, vol) == 0.0
    x = md.LennardJonesXPLOR(1.0, 1.0, 2.0, 2.5, True)
    assert abs(md.energy_lrc(x, n, vol) - e) < 1e-12 * abs(e)


def test_initialize_velocities(md):
    rng = np.random.default_rng(0)
    v = md.initialize_velocities(1.4737, rng, 1000, 3)
    assert v.shape == (1000, 3) and np.max(np.abs(v.sum(0))) < 1e-10
    assert abs(np.sum(v * v) / (3 * 999) - 1.4737) < 1e-12     # src/initialization.jl:39-42


def test_to_unitcell_and_files(md, tmp_path):
    assert np.array_equal(md.to_unitcell(2.0, 3), 2 * np.eye(3))
    assert np.array_equal(md.to_unitcell([1.0, 2.0], 2), np.diag([1.0, 2.0]))
    assert np.array_equal(md.to_unitcell(np.arange(16.0).reshape(4, 4), 3), np.arange(16.0).reshape(4, 4)[:3, :3])
    rng = np.random.default_rng(1)
    pos, diam, cell = rng.uniform(0, 5, (7, 3)), rng.uniform(0.8, 1.2, 7), np.diag([5.0, 6.0, 7.0])
    f = tmp_path / "c.xyz"
    md.write_to_file(str(f), 3, cell, 7, pos, diam, 3, mode="w")
    cell2, pos2, diam2 = md.read_file(str(f), dimension=3)
    assert np.allclose(cell2, cell) and np.allclose(pos2, pos, atol=1e-6) and np.allclose(diam2, diam, atol=2e-6)
    g = tmp_path / "t.lammpstrj"
    img = rng.integers(-2, 3, (7, 3)).astype(np.int32)
    md.write_to_file_lammps(str(g), 10, cell, 7, pos, img, diam, 3, mode="w")
    lines = open(g).read().splitlines()
    assert lines[0] == "ITEM: TIMESTEP" and lines[1] == "10" and lines[3] == "7"
    assert lines[8] == "ITEM: ATOMS id type radius x y z xu yu zu" and len(lines) == 9 + 7
    row = [float(t) for t in lines[9].split()]
    assert abs(row[6] - (pos[0, 0] + img[0, 0] * 5.0)) < 1e-5


def test_workloads_are_overlap_free(md, orc):
    from mdjl_b200 import workloads
    for n in (1024, 4000):
        c = workloads.phs_fluid(n)
        assert abs(c["rho"] - 6 * 0.47 / math.pi) < 1e-15 and c["x"].shape == (n, 3)
        r = orc.forces(c["x"], c["diam"], c["box"], 1.5, orc.POT_PHS)
        assert r["E"] < 50.0 * n and np.all(np.isfinite(r["F"]))
    p = workloads.poly2d()
    assert p["diam"].min() >= 0.73 and p["diam"].max() <= 1.62 and abs(p["box"][0] - math.sqrt(1200.0)) < 1e-12


def test_reference_arm_under_torchrun_prints_one_line_and_uses_the_host_cores():
    """bench.py --impl reference launched like the driver does for N > 1: rank 0 alone prints ONE JSON line, the other rank
    exits 0 without work, and the OpenMP team is not pinned to the single thread torchrun exports to its workers"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29571", os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2",
           "--warmup", "1", "--particles", "32768"]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=600, cwd=root)
    assert p.returncode == 0, p.stderr.decode()[-2000:]
    lines = [l for l in p.stdout.decode().splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    ncpu = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    assert d["cpu_baseline"]["cores"] == ncpu
