"""tests/golden/ref_outputs/* (written by julia/make_reference_golden.jl from the STOCK reference package) -> ref_c1.npz,
ref_c2.npz, the files tests/test_reference_golden.py consumes."""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
src = os.path.join(HERE, "ref_outputs")
for name, d in (("c1", 3), ("c2", 2)):
    f = lambda key, dt: np.fromfile(os.path.join(src, "%s_%s" % (name, key)), dtype=dt)
    n = f("F.f64", "<f8").size // d
    np.savez(os.path.join(HERE, "ref_%s.npz" % name), F=f("F.f64", "<f8").reshape(n, d), E=f("EW.f64", "<f8")[0], W=f("EW.f64", "<f8")[1],
             n_cut=f("counts.i64", "<i8")[0], n_int=f("counts.i64", "<i8")[1], nve_x=f("nve_x.f64", "<f8").reshape(n, d),
             nve_v=f("nve_v.f64", "<f8").reshape(n, d), nve_img=f("nve_img.i32", "<i4").reshape(n, d),
             nve_thermo=f("nve_thermo.f64", "<f8").reshape(-1, 3), source="MolecularDynamics.jl (stock), julia/make_reference_golden.jl")
    print("wrote ref_%s.npz (n = %d)" % (name, n))
