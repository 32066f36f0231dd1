"""Timing of the pair-force kernel and the whole step for one build variant (GPU box).
usage: MDB200_LIB=... python tools/tune_force.py [n] [skin]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mdjl_b200 as md
from mdjl_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
skin = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
skin_in = float(sys.argv[3]) if len(sys.argv) > 3 else 0.0
cfg = workloads.phs_fluid(n)
v0 = workloads.velocities(n, 3, workloads.KT_README)
e = md.Engine(3, n, cfg["box"], 1.5, 0, seed=1, skin=skin, skin_inner=skin_in, use_graph=True)
e.upload(cfg["x"], cfg["diam"], velocities=v0)
e.run_nvt(300, 1e-3, workloads.KT_README, 0.1, thermo=False)
e.run_nve(20, 1e-3, thermo=False)
e.run_nve(200, 1e-3, thermo=False)
s = e.stats()
step_ms = s["last_run_ms"] / 200
ft = []
for _ in range(5):
    e.compute_forces()
    ft.append(e.stats()["last_force_ms"])
x, v, f, img = e.download()
e2 = md.Engine(3, n, cfg["box"], 1.5, 0, seed=1, skin=skin, skin_inner=skin_in, use_graph=False)
e2.upload(x, cfg["diam"], velocities=v, forces=f, images=img)
e2.run_nve(5, 1e-3, thermo=False)
e2.run_nve(60, 1e-3, thermo=False)
p = e2.stats()
print("skin_in=%g " % skin_in, end="")
print("lib=%s n=%d skin=%g step_ms=%.4f rate=%.3e force_alone_ms=%.4f | eager: kick=%.4f force=%.4f rebuild/step=%.4f maxnbr=%d kmax=%d rebuilds=%d" % (
    os.path.basename(md._capi.lib_path()), n, s["r_search"] and skin, step_ms, n / step_ms * 1e3, min(ft), p["prof_kick_ms"] / 60, p["prof_force_ms"] / 60,
    p["prof_rebuild_ms"] / 60, s["max_neighbors"], s["list_capacity"], s["rebuilds"]))
