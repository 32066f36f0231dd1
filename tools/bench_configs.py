"""Throughput of every BASELINE.json config on one B200 (C1..C4; C5 is bench.py's default) next to the reference-shaped
OpenMP port on the host cores.  Prints one JSON object per config.  usage: python tools/bench_configs.py [--quick]"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
import numpy as np
import mdjl_b200 as md
from mdjl_b200 import workloads
import mdoracle as orc

quick = "--quick" in sys.argv
out = []


def cpu_rate(ens, x, v, diam, box, cutoff, tag, params, dt, steps, kt):
    n, dim = x.shape
    x, v = np.array(x), np.array(v)
    f, img = np.zeros_like(x), np.zeros((n, dim), np.int32)
    orc.run_timing(ens, x, v, f, img, diam, box, cutoff, tag, params, dt, 3, ktemp=kt, tau=100 * dt)
    t0 = time.perf_counter()
    orc.run_timing(ens, x, v, f, img, diam, box, cutoff, tag, params, dt, steps, ktemp=kt, tau=100 * dt)
    return n * steps / (time.perf_counter() - t0)


def gpu_case(name, dim, cfg, v0, cutoff, tag, params, dt, kt, ens, steps, melt, cpu_steps):
    n = cfg["x"].shape[0]
    e = md.Engine(dim, n, cfg["box"], cutoff, tag, params, seed=20261018)
    e.upload(cfg["x"], cfg["diam"], velocities=v0)
    if melt:
        e.run_nvt(melt, dt if ens != "brownian" else 1e-3, kt, 100 * (dt if ens != "brownian" else 1e-3), thermo=False)
    run = {"nve": lambda k: e.run_nve(k, dt, thermo=False), "nvt": lambda k: e.run_nvt(k, dt, kt, 100 * dt, thermo=False),
           "brownian": lambda k: e.run_brownian(k, dt, kt, thermo=False)}[ens]
    run(max(20, steps // 10))
    run(steps)
    st = e.stats()
    ms = st["last_run_ms"] / steps
    x, v, f, img = e.download()
    ens_id = {"nve": orc.NVE, "nvt": orc.NVT, "brownian": orc.BROWNIAN}[ens]
    cpu = cpu_rate(ens_id, x, v, cfg["diam"], cfg["box"], cutoff, tag, params, dt, cpu_steps, kt)
    rec = {"config": name, "n": n, "dim": dim, "ensemble": ens, "steps": steps, "us_per_step": 1e3 * ms, "steps_per_s": 1e3 / ms,
           "particle_steps_per_s": n / ms * 1e3, "mode": st["mode"], "rebuilds": st["rebuilds"],
           "cpu_port_particle_steps_per_s": cpu, "cpu_threads": orc.threads()}
    print(json.dumps(rec), flush=True)
    e.close()


kt = workloads.KT_README
c1 = workloads.phs_fluid(1024)
gpu_case("C1 3-D PseudoHS N=1024 NVT (README)", 3, c1, workloads.velocities(1024, 3, kt), 1.5, 0, (), 1e-3, kt, "nvt", 20000, 2000, 2000)
gpu_case("C1 3-D PseudoHS N=1024 NVE (README)", 3, c1, workloads.velocities(1024, 3, kt), 1.5, 0, (), 1e-3, kt, "nve", 20000, 2000, 2000)
c2 = workloads.poly2d(1200)
# the lattice start of the polydisperse mixture overlaps: relax it first with FIRE on the device, then thermalise
e = md.Engine(2, 1200, c2["box"], 1.5, md._capi.POT_POLY, (1.25, 0.2), seed=1)
e.upload(c2["x"], c2["diam"])
e.fire_minimize(max_steps=3000, tol=1e-3, dt_initial=1e-4, dt_max=5e-3)
c2["x"] = e.download()[0]
e.close()
gpu_case("C2 2-D polydisperse N=1200 NVE", 2, c2, workloads.velocities(1200, 2, 0.11), 1.5, md._capi.POT_POLY, (1.25, 0.2), 5e-3, 0.11, "nve",
         20000, 2000, 2000)
n20 = 1 << (18 if quick else 20)
c3 = workloads.phs_fluid(n20)
gpu_case("C3 3-D PseudoHS N=2^20 NVT", 3, c3, workloads.velocities(n20, 3, kt), 1.5, 0, (), 1e-3, kt, "nvt", 1000, 1500, 10)
gpu_case("C4 3-D PseudoHS N=2^20 Brownian", 3, c3, workloads.velocities(n20, 3, kt), 1.5, 0, (), 1e-5, kt, "brownian", 1000, 1500, 10)
