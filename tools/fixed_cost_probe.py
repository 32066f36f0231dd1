"""Per-step cost of the list-mode (multi-kernel, graph-replayed) step against N: where the fixed part sits.
usage: python tools/fixed_cost_probe.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mdjl_b200 as md
from mdjl_b200 import workloads

for n in (4096, 32768, 262144, 1 << 20, 1 << 21):
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, workloads.KT_README)
    res = {}
    for graph in (True, False):
        e = md.Engine(3, n, cfg["box"], 1.5, 0, seed=1, mode=md._capi.MODE_LIST, use_graph=graph)
        e.upload(cfg["x"], cfg["diam"], velocities=v0)
        e.run_nvt(300, 1e-3, workloads.KT_README, 0.1, thermo=False)
        e.run_nve(50, 1e-3, thermo=False)
        k = 400 if graph else 100
        e.run_nve(k, 1e-3, thermo=False)
        s = e.stats()
        if graph:
            res["graph_us"] = round(1e3 * s["last_run_ms"] / k, 2)
        else:
            res.update(eager_force_us=round(1e3 * s["prof_force_ms"] / k, 2), eager_rebuild_us=round(1e3 * s["prof_rebuild_ms"] / k, 2))
        e.close()
    print(n, res, flush=True)
