"""Host-side logic of the x-slab decomposition, mirrored from csrc/engine.cu (plan_neighbors, mdb_upload):
cell-column ownership, boundary/ghost columns, migration targets, and the NCCL unique-id hand-off over torch.distributed.
The device path never calls this module; it exists so that the partition arithmetic and the rank plumbing can be
exercised on CPU (world_size-2 gloo tests) and reused by bench.py."""
import math

import numpy as np


def plan(box, dim, r_search, skin, nranks):
    """global cell grid at r_grid = r_search + skin and the owned column range of every rank"""
    r_grid = r_search + skin
    nc = [1, 1, 1]
    for k in range(dim):
        nc[k] = int(math.floor(box[k] / (r_grid * (1.0 + 1e-6))))
        if nc[k] < 3:
            raise ValueError("slab decomposition needs at least 3 cells per direction")
    if nc[0] // nranks < 2:
        raise ValueError("each slab needs at least two cell columns")
    cols = [(r * nc[0] // nranks, (r + 1) * nc[0] // nranks) for r in range(nranks)]
    return dict(nc=nc, r_grid=r_grid, columns=cols, box=np.asarray(box, dtype=np.float64), dim=dim, nranks=nranks)


def wrap_x(x, L):
    """wrap_to_box arithmetic (src/boundary.jl:7-17) applied where k_import applies it: only outside [0, L)"""
    x = np.array(x, dtype=np.float64)
    out = (x < 0.0) | (x >= L)
    frac = (1.0 / L) * x[out]
    x[out] = L * (frac - np.floor(frac))
    return x


def cell_coord(x, L, nc):
    c = (x * (nc / L)).astype(np.int64)
    return np.clip(c, 0, nc - 1)


def column_of(x, pl):
    return cell_coord(wrap_x(x, pl["box"][0]), pl["box"][0], pl["nc"][0])


def owner_of_column(cx, pl):
    nx, P = pl["nc"][0], pl["nranks"]
    owner = np.empty_like(cx)
    for r, (c0, c1) in enumerate(pl["columns"]):
        owner[(cx >= c0) & (cx < c1)] = r
    return owner


def split(x, pl):
    """indices of the particles each rank owns (a partition of range(n))"""
    owner = owner_of_column(column_of(x[:, 0], pl), pl)
    return [np.nonzero(owner == r)[0] for r in range(pl["nranks"])]


def boundary_indices(x_owned, pl, rank):
    """(to_left, to_right): owned particles in the first / last owned column = what the neighbours need as ghosts"""
    c0, c1 = pl["columns"][rank]
    cx = column_of(x_owned[:, 0], pl)
    return np.nonzero(cx == c0)[0], np.nonzero(cx == c1 - 1)[0]


def migration_targets(x_owned, pl, rank):
    """-1 stay, 0 to the left neighbour, 1 to the right neighbour; raises on a jump of more than one column"""
    nx = pl["nc"][0]
    c0, c1 = pl["columns"][rank]
    cx = column_of(x_owned[:, 0], pl)
    tgt = np.full(cx.shape, -1, dtype=np.int64)
    tgt[cx == (c0 - 1) % nx] = 0
    tgt[cx == c1 % nx] = 1
    inside = (cx >= c0) & (cx < c1)
    tgt[inside] = -1
    if np.any(~inside & (tgt < 0)):
        raise RuntimeError("a particle crossed more than one cell column")
    return tgt


def ring_neighbours(rank, nranks):
    return (rank - 1) % nranks, (rank + 1) % nranks


def broadcast_unique_id(dist, rank, make_id, device=None):
    """rank 0 creates the 128-byte NCCL unique id, everybody receives it (works with the gloo and the nccl backend)"""
    import torch
    buf = torch.zeros(128, dtype=torch.uint8, device=device if device is not None else "cpu")
    if rank == 0:
        raw = make_id()
        buf.copy_(torch.tensor(list(raw), dtype=torch.uint8))
    dist.broadcast(buf, src=0)
    return bytes(buf.cpu().tolist())
