// slab.cuh -- kernels of the x-slab decomposition (new: the reference is single-process).
// Each rank owns the global cell columns [c0, c0 + nxo).  At every neighbour rebuild:
//   classify (leavers -> migration buffers) -> exchange -> unpack arrivals -> counting sort of the owned set
//   -> pack the two boundary columns -> exchange -> ghost cell ranges -> Verlet list.
// Between rebuilds only the boundary columns' positions travel, packed in the same order, so the receiver's ghost
// buffer is updated in place with no index translation.
#pragma once
#include "kernels.cuh"

namespace mdb {

constexpr int kErrMigrationOverflow = 1;   // more leavers than the migration buffer holds
constexpr int kErrGhostOverflow = 2;       // boundary column larger than the ghost buffer
constexpr int kErrOwnedOverflow = 4;       // owned + arrivals exceed the slab capacity
constexpr int kErrLongJump = 8;            // a particle moved more than one cell column between rebuilds

struct MigRec {  // 96 bytes; record 0 of every buffer is a header with p.x = record count
    double4 p;
    double v[3];
    double f[3];
    int32_t img[3];
    int32_t id;
};

// owned particles: stay (cell + arrival slot) or leave (packed for the left/right neighbour)
template <int DIM>
__global__ void k_slab_classify(DevCtl *ctl, Grid g, uint32_t *__restrict__ cell_of, uint32_t *__restrict__ slot_of,
                                uint32_t *__restrict__ counts, MigRec *__restrict__ out_l, MigRec *__restrict__ out_r, int mig_cap)
{
    const StatePtrs s = ctl->st[ctl->cur];
    const int n = ctl->n_own;
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) ctl->n_tmp = n;
    if (i >= n) return;
    double4 p = s.pos[i];
    int cx = cell_coord(p.x, g.cinv[0], g.nc[0]);
    int cy = cell_coord(p.y, g.cinv[1], g.nc[1]);
    int cz = (DIM == 3) ? cell_coord(p.z, g.cinv[2], g.nc[2]) : 0;
    int lx = cx - g.c0;
    if (lx >= 0 && lx < g.nxo) {
        uint32_t c = ((uint32_t)cz * g.nc[1] + cy) * g.nxo + lx;
        cell_of[i] = c;
        slot_of[i] = atomicAdd(&counts[c], 1u);
        return;
    }
    cell_of[i] = kInvalidCell;
    const int nxg = g.nc[0];
    int left_col = (g.c0 - 1 + nxg) % nxg, right_col = (g.c0 + g.nxo) % nxg;
    int dir;
    if (cx == left_col) dir = 0;
    else if (cx == right_col) dir = 1;
    else {
        atomicOr(&ctl->error, kErrLongJump);
        return;
    }
    int k = atomicAdd(&ctl->mig_count[dir], 1);
    if (k >= mig_cap) {
        atomicOr(&ctl->error, kErrMigrationOverflow);
        return;
    }
    MigRec r;
    r.p = p;
#pragma unroll
    for (int q = 0; q < 3; q++) {
        r.v[q] = q < DIM ? s.vel[q * s.cap + i] : 0.0;
        r.f[q] = q < DIM ? s.frc[q * s.cap + i] : 0.0;
        r.img[q] = q < DIM ? s.img[q * s.cap + i] : 0;
    }
    r.id = s.id[i];
    (dir == 0 ? out_l : out_r)[1 + k] = r;
}

// graph replay exchanges the migration buffers every step: without a rebuild they must say "nobody leaves"
__global__ void k_slab_zero_headers(MigRec *out_l, MigRec *out_r)
{
    out_l[0].p.x = 0.0;
    out_r[0].p.x = 0.0;
}

__global__ void k_slab_mig_headers(DevCtl *ctl, MigRec *out_l, MigRec *out_r, int mig_cap)
{
    out_l[0].p.x = (double)min(ctl->mig_count[0], mig_cap);
    out_r[0].p.x = (double)min(ctl->mig_count[1], mig_cap);
}

// arrivals are appended behind the old owned range and join the counting sort
template <int DIM>
__global__ void k_slab_unpack(DevCtl *ctl, Grid g, const MigRec *__restrict__ in_l, const MigRec *__restrict__ in_r,
                              uint32_t *__restrict__ cell_of, uint32_t *__restrict__ slot_of, uint32_t *__restrict__ counts,
                              int cap_own)
{
    const StatePtrs s = ctl->st[ctl->cur];
    const int n = ctl->n_own;
    const int cl = (int)in_l[0].p.x, cr = (int)in_r[0].p.x;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        if (n + cl + cr > cap_own) atomicOr(&ctl->error, kErrOwnedOverflow);
        ctl->n_tmp = min(n + cl + cr, cap_own);
    }
    if (t >= cl + cr) return;
    int dst = n + t;
    if (dst >= cap_own) return;
    const MigRec r = t < cl ? in_l[1 + t] : in_r[1 + (t - cl)];
    s.pos[dst] = r.p;
#pragma unroll
    for (int q = 0; q < DIM; q++) {
        s.vel[q * s.cap + dst] = r.v[q];
        s.frc[q * s.cap + dst] = r.f[q];
        s.img[q * s.cap + dst] = r.img[q];
    }
    s.id[dst] = r.id;
    int cx = cell_coord(r.p.x, g.cinv[0], g.nc[0]);
    int cy = cell_coord(r.p.y, g.cinv[1], g.nc[1]);
    int cz = (DIM == 3) ? cell_coord(r.p.z, g.cinv[2], g.nc[2]) : 0;
    int lx = cx - g.c0;
    if (lx < 0 || lx >= g.nxo) {
        atomicOr(&ctl->error, kErrLongJump);
        cell_of[dst] = kInvalidCell;
        return;
    }
    uint32_t c = ((uint32_t)cz * g.nc[1] + cy) * g.nxo + lx;
    cell_of[dst] = c;
    slot_of[dst] = atomicAdd(&counts[c], 1u);
}

// per (cz,cy) row: population of the first and the last owned column
__global__ void k_slab_rowcounts(int nrows, int nxo, const uint32_t *__restrict__ start, uint32_t *__restrict__ cnt_l,
                                 uint32_t *__restrict__ cnt_r)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    uint32_t b = (uint32_t)row * nxo;
    cnt_l[row] = start[b + 1] - start[b];
    cnt_r[row] = start[b + nxo] - start[b + nxo - 1];
}

// single CTA: out[r] = base + exclusive prefix of cnt[0..r), out[nrows] = base + total (two arrays per launch)
__global__ void k_slab_rowscan(int nrows, const uint32_t *__restrict__ cnt_a, uint32_t *__restrict__ out_a, uint32_t base_a,
                               const uint32_t *__restrict__ cnt_b, uint32_t *__restrict__ out_b, uint32_t base_b)
{
    __shared__ uint32_t sm[1024];
    __shared__ uint32_t carry;
    for (int which = 0; which < 2; which++) {
        const uint32_t *cnt = which ? cnt_b : cnt_a;
        uint32_t *out = which ? out_b : out_a;
        uint32_t base = which ? base_b : base_a;
        if (threadIdx.x == 0) carry = 0;
        __syncthreads();
        for (int b0 = 0; b0 < nrows; b0 += 1024) {
            int q = b0 + threadIdx.x;
            uint32_t v = q < nrows ? cnt[q] : 0;
            sm[threadIdx.x] = v;
            __syncthreads();
            for (int o = 1; o < 1024; o <<= 1) {
                uint32_t t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
                __syncthreads();
                sm[threadIdx.x] += t;
                __syncthreads();
            }
            uint32_t incl = sm[threadIdx.x];
            if (q < nrows) out[q] = base + carry + incl - v;
            __syncthreads();
            if (threadIdx.x == 1023) carry += incl;
            __syncthreads();
        }
        if (threadIdx.x == 0) out[nrows] = base + carry;
        __syncthreads();
    }
}

// boundary columns -> send buffers, row by row in slot order (used at rebuilds AND every step: same order, so the
// receiver's ghost records keep their indices).  rowoff_* are 0-based exclusive prefixes.
__global__ void k_slab_pack_ghost(const DevCtl *__restrict__ ctl, int nrows, int nxo, const uint32_t *__restrict__ start,
                                  const uint32_t *__restrict__ rowoff_l, const uint32_t *__restrict__ rowoff_r,
                                  double4 *__restrict__ out_l, double4 *__restrict__ out_r, int ghost_cap, DevCtl *ctl_w)
{
    const double4 *__restrict__ pos = ctl->st[ctl->cur].pos;
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row == 0) {
        uint32_t tl = rowoff_l[nrows], tr = rowoff_r[nrows];
        if (tl > (uint32_t)ghost_cap || tr > (uint32_t)ghost_cap) atomicOr(&ctl_w->error, kErrGhostOverflow);
        out_l[0] = make_double4((double)min(tl, (uint32_t)ghost_cap), 0, 0, 0);
        out_r[0] = make_double4((double)min(tr, (uint32_t)ghost_cap), 0, 0, 0);
    }
    if (row >= nrows) return;
    uint32_t b = (uint32_t)row * nxo;
    uint32_t o = rowoff_l[row];
    for (uint32_t j = start[b]; j < start[b + 1]; j++, o++)
        if (o < (uint32_t)ghost_cap) out_l[1 + o] = pos[j];
    o = rowoff_r[row];
    for (uint32_t j = start[b + nxo - 1]; j < start[b + nxo]; j++, o++)
        if (o < (uint32_t)ghost_cap) out_r[1 + o] = pos[j];
}

// receiver: population of every ghost cell (one cell per (cz,cy) row and side) from the received records
template <int DIM>
__global__ void k_slab_ghost_count(Grid g, const double4 *__restrict__ gl, const double4 *__restrict__ gr,
                                   uint32_t *__restrict__ cnt_l, uint32_t *__restrict__ cnt_r)
{
    const int nl = (int)gl[0].x, nr = (int)gr[0].x;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nl + nr) return;
    const double4 p = t < nl ? gl[1 + t] : gr[1 + (t - nl)];
    int cy = cell_coord(p.y, g.cinv[1], g.nc[1]);
    int cz = (DIM == 3) ? cell_coord(p.z, g.cinv[2], g.nc[2]) : 0;
    atomicAdd(&(t < nl ? cnt_l : cnt_r)[cz * g.nc[1] + cy], 1u);
}

}  // namespace mdb
