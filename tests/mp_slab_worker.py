"""Worker for tests/test_gpu_multirank.py: run under torchrun with one process per GPU.  Every rank steps its slab over
NCCL; rank 0 also runs the single-domain engine on its GPU and compares."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import numpy as np
import torch
import torch.distributed as dist

import mdjl_b200 as md
from mdjl_b200 import slabs, workloads


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    uid = slabs.broadcast_unique_id(dist, rank, md.unique_id, device=dev)
    n = 32768
    cfg = workloads.phs_fluid(n)
    v0 = workloads.velocities(n, 3, 1.4737)
    ring = md.SlabRing.nccl(rank, world, uid, 3, n, cfg["box"], 1.5, 0, seed=77, device=local)
    ring.upload(cfg["x"], cfg["diam"], velocities=v0)
    t1 = ring.run_nvt(400, 1e-3, 1.4737, 0.1)
    t2 = ring.run_nve(200, 1e-3)
    st = ring.lead.stats()
    expect = os.environ.get("MDB200_EXPECT_TRANSPORT")
    if expect and int(expect) != st["slab_transport"]:
        print("MULTIRANK FAIL transport %d, expected %s" % (st["slab_transport"], expect), flush=True)
        sys.exit(1)
    ids, x, v, f, img = ring.download_local()
    parts = [None] * world
    dist.all_gather_object(parts, (ids, x, v, f, img))
    ok = True
    msg = ""
    if rank == 0:
        X, V = np.empty((n, 3)), np.empty((n, 3))
        seen = np.zeros(n, dtype=np.int64)
        for (i_, x_, v_, f_, m_) in parts:
            X[i_], V[i_] = x_, v_
            seen[i_] += 1
        single = md.Engine(3, n, cfg["box"], 1.5, 0, seed=77, device=local)
        single.upload(cfg["x"], cfg["diam"], velocities=v0)
        s1 = single.run_nvt(400, 1e-3, 1.4737, 0.1)
        s2 = single.run_nve(200, 1e-3)
        xs, vs, _, _ = single.download()
        checks = {
            "partition": bool(np.all(seen == 1)),
            "pairs_nvt": bool(np.array_equal(t1[:, 3], s1[:, 3])),
            "pairs_nve": bool(np.array_equal(t2[:, 3], s2[:, 3])),
            "thermo": bool(np.allclose(t1[:, :3], s1[:, :3], rtol=1e-8) and np.allclose(t2[:, :3], s2[:, :3], rtol=1e-8)),
            "positions": float(np.max(np.abs(X - xs))),
            "velocities": float(np.max(np.abs(V - vs))),
        }
        ok = checks["partition"] and checks["pairs_nvt"] and checks["pairs_nve"] and checks["thermo"] and \
            checks["positions"] < 1e-7 and checks["velocities"] < 1e-6
        msg = repr(checks)
        print("MULTIRANK", "OK" if ok else "FAIL", world, msg, flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, src=0)
    ring.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
