"""A/B of the list-mode pair-force kernel variants on one GPU (run on the GPU box).

    [MDB200_LIB=variant.so] python tools/force_ab.py [n] [variants e.g. 0,1] [ensemble nve|brownian]

For every variant (MDB200_FORCE_VARIANT: 0 = k_force_list, direct gathers; 1 = k_force_list_staged, cp.async staging):
the same melted state is stepped (a) eagerly with CUDA events around the force kernel (kernel time per step), (b) as the
graph-replayed production step (ms/step).  The final state of every variant must equal variant 0's BIT FOR BIT
(positions, velocities, forces, images); thermo rows agree to rounding (the per-CTA grouping of the sums follows the grid).
One JSON line per variant."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import mdjl_b200 as md
from mdjl_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
variants = [int(v) for v in (sys.argv[2] if len(sys.argv) > 2 else "0,1").split(",")]
ensemble = sys.argv[3] if len(sys.argv) > 3 else "nve"
DT, KT = 1e-3, workloads.KT_README
cfg = workloads.phs_fluid(n)
v0 = workloads.velocities(n, 3, KT)


def engine(variant, graph):
    os.environ["MDB200_FORCE_VARIANT"] = str(variant)
    return md.Engine(3, n, cfg["box"], 1.5, 0, seed=1, use_graph=graph, mode=md._capi.MODE_LIST)


def run(e, k, thermo=False):
    if ensemble == "nve":
        return e.run_nve(k, DT, thermo=thermo)
    return e.run_brownian(k, 1e-5, KT, thermo=thermo)


e = engine(0, True)
e.upload(cfg["x"], cfg["diam"], velocities=v0)
e.run_nvt(int(os.environ.get("AB_MELT", "400")), DT, KT, 0.1, thermo=False)
state = e.download()
e.close()
ref = None
for v in variants:
    eg = engine(v, False)
    eg.upload(state[0], cfg["diam"], velocities=state[1], forces=state[2], images=state[3])
    run(eg, 5)
    run(eg, 60)
    p = eg.stats()
    eg.close()
    g = engine(v, True)
    g.upload(state[0], cfg["diam"], velocities=state[1], forces=state[2], images=state[3])
    run(g, 20)
    t = run(g, 200, thermo=True)
    s = g.stats()
    fin = g.download()
    g.close()
    same = None
    cross = None
    # across build variants (separate processes): AB_SAVE=path keeps variant 0's final state, AB_REF=path compares with it
    if os.environ.get("AB_SAVE") and ref is None:
        np.savez(os.environ["AB_SAVE"], x=fin[0], v=fin[1], f=fin[2], img=fin[3], t=t)
    if os.environ.get("AB_REF"):
        r = np.load(os.environ["AB_REF"])
        cross = {"x_max_abs_diff": float(np.max(np.abs(r["x"] - fin[0]))), "v_max_abs_diff": float(np.max(np.abs(r["v"] - fin[1]))),
                 "img_equal": bool(np.array_equal(r["img"], fin[3])), "pair_counts_equal": bool(np.array_equal(r["t"][:, 3], t[:, 3])),
                 "thermo_max_rel_diff": float(np.max(np.abs(r["t"][:, :3] - t[:, :3]) / np.abs(r["t"][:, :3])))}
    if ref is None:
        ref = (fin, t)
    else:
        same = all(np.array_equal(a, b) for a, b in zip(ref[0], fin)) and bool(np.array_equal(ref[1][:, 3], t[:, 3])) and \
            bool(np.allclose(ref[1][:, :3], t[:, :3], rtol=1e-12))
    print(json.dumps({"lib": os.path.basename(md._capi.lib_path()), "variant": v, "n": n, "ensemble": ensemble,
                      "force_kernel_ms": p["prof_force_ms"] / 60, "kick_ms": p["prof_kick_ms"] / 60, "rebuild_ms_per_step": p["prof_rebuild_ms"] / 60,
                      "step_ms_graph": s["last_run_ms"] / 200, "rate": n / (s["last_run_ms"] / 200) * 1e3,
                      "bit_identical_to_first": same, "vs_saved_reference": cross, "rebuilds": int(s["rebuilds"])}), flush=True)
