"""Kernel times of the slab step's head on ONE GPU (in-process ring of 2 slabs, peer-memory kernels with local mailboxes):
run under `ncu --metrics gpu__time_duration.sum -k regex:k_peer` to see what the pack / wait kernels cost by themselves.
usage: python tools/peer_head_probe.py [n]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mdjl_b200 as md
from mdjl_b200 import workloads

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
cfg = workloads.phs_fluid(n)
v0 = workloads.velocities(n, 3, workloads.KT_README)
ring = md.SlabRing.local(2, 3, n, cfg["box"], 1.5, md._capi.POT_PSEUDOHS, seed=1, slab_transport=2, use_graph=False)
ring.upload(cfg["x"], cfg["diam"], velocities=v0)
ring.run_nvt(100, 1e-3, workloads.KT_README, 0.1, thermo=False)
ring.run_nve(30, 1e-3, thermo=False)
st = ring.lead.stats()
print("ms/step (2 slabs on one GPU, eager)", st["last_run_ms"] / 30, "ncell", st["ncell"])
ring.close()
