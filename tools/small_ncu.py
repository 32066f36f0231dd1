import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, mdjl_b200 as md
g = dict(np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "c1_phs_n1024.npz")))
e = md.Engine(3, 1024, g["box"], 1.5, 0, seed=1, mode=md._capi.MODE_SMALL)
e.upload(g["x"], g["diam"], velocities=g["v"], forces=g["f"], images=g["img"])
e.run_nve(300, 1e-3, thermo=False)
print(e.stats()["last_run_ms"] / 300 * 1e3, "us/step")
