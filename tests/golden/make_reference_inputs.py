"""Snapshots of BASELINE configs 1 and 2 as raw little-endian files for julia/make_reference_golden.jl (Julia's stdlib reads
no .npz).  (n, D) C-order arrays are what Julia reads as column-major D x n matrices."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mdjl_b200 import workloads  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
out = os.path.join(HERE, "ref_inputs")
os.makedirs(out, exist_ok=True)
for name, src in (("c1", "c1_phs_n1024.npz"), ("c2", "c2_poly_n1200_cut1.5.npz")):
    g = np.load(os.path.join(HERE, src))
    n, d = g["x"].shape
    open(os.path.join(out, name + "_n.txt"), "w").write("%d\n" % n)
    # C2's snapshot carries no velocities: the seeded Maxwell-Boltzmann draw the tests use (kT = 0.11, README.md:147-176)
    v = g["v"] if "v" in g.files else workloads.velocities(n, d, 0.11)
    for key, arr in (("x", g["x"]), ("v", v), ("diam", g["diam"]), ("box", np.asarray(g["box"], dtype=np.float64).ravel()[:d])):
        np.ascontiguousarray(arr, dtype="<f8").tofile(os.path.join(out, "%s_%s.f64" % (name, key)))
print("wrote", out)
