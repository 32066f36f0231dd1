import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mdjl_b200 as md
from mdjl_b200 import workloads
for n in (256, 1024, 4096):
    cfg = workloads.phs_fluid(n); v0 = workloads.velocities(n, 3, 1.4737)
    for mode, name in ((md._capi.MODE_SMALL, "small"), (md._capi.MODE_LIST, "list")):
        e = md.Engine(3, n, cfg["box"], 1.5, 0, seed=1, mode=mode)
        e.upload(cfg["x"], cfg["diam"], velocities=v0)
        e.run_nvt(2000, 1e-3, 1.4737, 0.1, thermo=False)
        res = {}
        for ens in ("nve", "nvt", "brownian"):
            run = {"nve": lambda k: e.run_nve(k, 1e-3, thermo=False), "nvt": lambda k: e.run_nvt(k, 1e-3, 1.4737, 0.1, thermo=False), "brownian": lambda k: e.run_brownian(k, 1e-5, 1.4737, thermo=False)}[ens]
            run(500); run(4000)
            res[ens] = e.stats()["last_run_ms"] / 4000 * 1e3
        print(n, name, {k: round(v, 2) for k, v in res.items()}, "us/step")
        e.close()
