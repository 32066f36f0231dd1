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
constexpr int kErrPeerTimeout = 16;        // a ring neighbour's message did not arrive in time (peer-memory transport)

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
    __threadfence_system();  // the record may live in a neighbour's mailbox (peer-memory transport)
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


// ------------------------------------------------------------------------------------------------
// Peer-memory transport (one process per GPU, NVLink/NVSwitch): every rank owns one MAILBOX in its own HBM, mapped into
// its ring neighbours (and, for the tiny reduction slots, into every rank) with cudaIpc*.  Messages are written by the
// SENDER's kernels straight into the receiver's mailbox (st.global on the mapped address), followed by a system-scope
// fence and a 64-bit epoch flag; the receiver's stream waits in a one-warp kernel with a bounded spin.  No NCCL call, no
// host round trip and no copy engine in the step: the whole slab step (head, conditional rebuild with its two
// exchanges, forces, thermo) is plain kernels and replays as ONE CUDA graph.
//   * epoch E = number of force-evaluation heads executed so far (ctl->epoch, identical on every rank);
//   * per-step ghost data and the reduction slots are double-buffered by the parity of E: a sender can only be one
//     head ahead of a receiver (head E+1 waits for the neighbour's head-E+1 message, which the neighbour sends after its
//     forces of head E), so parity E is never overwritten while a rank still reads it;
//   * flags only grow (wait for >= E).
// The in-process ring of the single-GPU tests runs the same kernels with the other engines' mailboxes as "peers"; the
// kernels of all ranks are then launched phase by phase on one stream, so every wait finds its flag already set.
// ------------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 16;

struct PeerHdr {
    unsigned long long ghost_flag[2];   // [side] epoch of the last per-step ghost message from the left (0) / right (1) neighbour
    unsigned long long rghost_flag[2];  // same, rebuild-time ghost message (new ghost set)
    unsigned long long mig_flag[2];     // same, migration message
    unsigned long long pad[2];
    unsigned long long red_tag[2][kMaxRanks];   // [parity][source rank] epoch of red_val
    unsigned long long red_val[2][kMaxRanks];   // bit pattern of the rank's largest squared displacement
    unsigned long long sum_tag[2][kMaxRanks];   // thermostat sums (NVT: global kinetic energy inside the step)
    double sum_val[2][kMaxRanks][4];
};
struct PeerView {
    PeerHdr *hdr;
    double4 *ghost;  // [2 parity][2 side][1 + ghost_cap]; side 0 = column received from the left neighbour, 1 = from the right
    MigRec *mig;     // [2 side][1 + mig_cap]
};
struct PeerLinks {
    PeerView self, left, right;
    PeerHdr *all[kMaxRanks];  // every rank's header; all[me] == self.hdr
    int me, nranks, ghost_cap, mig_cap;
    long long timeout_ns;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// flag store behind an explicit __threadfence_system(): fence + relaxed store is a release, and ONE fence covers every flag
// that follows it.  A st.release.sys per flag carries a system-scope fence each: P + 2 of them in a row made the publishing
// thread of the per-step pack kernel the longest part of a slab step's head (ncu: 21 us per launch with 2 ranks)
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// bounded spin until *flag >= want; false on timeout.  A timeout is sticky (ctl->error): later waits return at once, so a
// dead neighbour costs one timeout, not one per step, and the run ends with MDB_ERR_STATE instead of hanging the GPU.
__device__ __forceinline__ bool peer_wait_flag(const unsigned long long *flag, unsigned long long want, long long timeout_ns, DevCtl *ctl)
{
    if (ld_acquire_sys_u64(flag) >= want) return true;
    if (*(volatile int *)&ctl->error & kErrPeerTimeout) return false;
    const unsigned long long t0 = global_timer_ns();
    for (;;) {
        if (ld_acquire_sys_u64(flag) >= want) return true;
        if ((long long)(global_timer_ns() - t0) > timeout_ns) {
            atomicOr(&ctl->error, kErrPeerTimeout);
            return false;
        }
        __nanosleep(64);
    }
}

__device__ __forceinline__ double4 *peer_ghost(const PeerView &v, int parity, int side, int ghost_cap)
{
    return v.ghost + (size_t)(parity * 2 + side) * (size_t)(1 + ghost_cap);
}

// Boundary columns -> the NEIGHBOURS' ghost buffers (same row-by-row order as k_slab_pack_ghost), then the flags.
// rebuild: source slot of every record of the two boundary columns, in the order the ghost buffers keep (row by row, slot
// order inside a cell): gsrc_l[rowoff_l[row] + k] = start[row * nxo] + k, likewise for the last column
__global__ void k_slab_ghost_src(int nrows, int nxo, const uint32_t *__restrict__ start, const uint32_t *__restrict__ rowoff_l,
                                 const uint32_t *__restrict__ rowoff_r, uint32_t *__restrict__ gsrc_l, uint32_t *__restrict__ gsrc_r, int ghost_cap)
{
    int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const uint32_t cap = (uint32_t)ghost_cap;
    uint32_t b = (uint32_t)row * nxo;
    uint32_t o = rowoff_l[row];
    for (uint32_t j = start[b]; j < start[b + 1]; j++, o++)
        if (o < cap) gsrc_l[o] = j;
    o = rowoff_r[row];
    for (uint32_t j = start[b + nxo - 1]; j < start[b + nxo]; j++, o++)
        if (o < cap) gsrc_r[o] = j;
}

// kind 0: head of a force evaluation (epoch E = ctl->epoch + 1): per-step ghost flag + this rank's displacement bound
//         into every rank's reduction slot;  kind 1: rebuild (E = ctl->epoch, already advanced): rebuild ghost flag.
// `done` is a zeroed counter used to find the last CTA (it re-zeroes it).
constexpr int kPackBlock = 1024;   // one system-scope fence per CTA: few, large CTAs
__global__ void __launch_bounds__(kPackBlock)
k_peer_pack_ghost(DevCtl *ctl, int nrows, const uint32_t *__restrict__ rowoff_l, const uint32_t *__restrict__ rowoff_r,
                  const uint32_t *__restrict__ gsrc_l, const uint32_t *__restrict__ gsrc_r, PeerLinks lk, int kind, unsigned int *done)
{
    // one thread per ghost RECORD (source slots tabulated at the rebuild by k_slab_ghost_src): index -> record -> remote store,
    // two independent chains per thread.  The first version walked one cell row per thread (row offsets -> cell ranges -> a serial
    // loop per column): 26 us per launch at 16 000 rows, the bulk of a slab step's head (ncu, tools/peer_head_probe.py)
    const double4 *__restrict__ pos = ctl->st[ctl->cur].pos;
    const unsigned long long E = ctl->epoch + (kind == 0 ? 1ull : 0ull);
    const int par = (int)(E & 1ull);
    double4 *out_l = peer_ghost(lk.left, par, 1, lk.ghost_cap);   // my first column is the left neighbour's RIGHT ghost column
    double4 *out_r = peer_ghost(lk.right, par, 0, lk.ghost_cap);
    const uint32_t cap = (uint32_t)lk.ghost_cap;
    const uint32_t tl = rowoff_l[nrows], tr = rowoff_r[nrows];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t == 0) {
        if (tl > cap || tr > cap) atomicOr(&ctl->error, kErrGhostOverflow);
        out_l[0] = make_double4((double)min(tl, cap), 0, 0, 0);
        out_r[0] = make_double4((double)min(tr, cap), 0, 0, 0);
    }
    const bool hl = t < min(tl, cap), hr = t < min(tr, cap);
    uint32_t jl = 0, jr = 0;
    if (hl) jl = gsrc_l[t];
    if (hr) jr = gsrc_r[t];
    double4 pl = make_double4(0, 0, 0, 0), pr = pl;
    if (hl) pl = ld_pos(pos + jl);
    if (hr) pr = ld_pos(pos + jr);
    if (hl) st_pos(out_l + 1 + t, pl);
    if (hr) st_pos(out_r + 1 + t, pr);
    // CTA barrier, then ONE system-scope fence by the reporting thread: the barrier orders every thread's stores before it and
    // the fence is cumulative (the pattern of a cooperative-groups multi-device barrier).  A fence per thread measured 20 us per
    // launch for 66 000 records -- MEMBAR.SYS, not the copies, was the cost of this kernel
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int k = atomicAdd(done, 1u);
        if (k == gridDim.x - 1) {  // last CTA: everybody's stores are out
            *done = 0u;
            if (kind == 0) {
                const unsigned long long bits = ctl->dmax2_bits;
                for (int r = 0; r < lk.nranks; r++) lk.all[r]->red_val[par][lk.me] = bits;
            }
            // one fence: every CTA's records (each CTA fenced before it reported to `done`; fence - atomic - atomic - fence
            // synchronises) and the bound above are visible system-wide before any of the flags below
            __threadfence_system();
            if (kind == 0) {
                for (int r = 0; r < lk.nranks; r++) st_relaxed_sys_u64(&lk.all[r]->red_tag[par][lk.me], E);
                st_relaxed_sys_u64(&lk.left.hdr->ghost_flag[1], E);
                st_relaxed_sys_u64(&lk.right.hdr->ghost_flag[0], E);
            } else {
                st_relaxed_sys_u64(&lk.left.hdr->rghost_flag[1], E);
                st_relaxed_sys_u64(&lk.right.hdr->rghost_flag[0], E);
            }
        }
    }
}

// the skin test of k_skin_check (kernels.cuh) as a function: the peer head runs it behind its own reduction
__device__ __forceinline__ int skin_decide(double m, double scale, double skin, double skin_in, int always, int exact, DevCtl *ctl)
{
    double step = sqrt(m) * scale;
    const unsigned long long dref = ctl->dref2_bits;
    double disp = (exact && dref != 0ull) ? sqrt(__longlong_as_double((long long)dref)) : ctl->disp + step;
    if (!exact) ctl->dref2_bits = 0ull;
    int need = always || !ctl->list_valid || !(2.0 * disp <= skin);
    ctl->disp = need ? 0.0 : disp;
    ctl->need_rebuild = need;
    double disp_in = ctl->disp_in + step;
    int need_in = need || !(2.0 * disp_in <= skin_in);
    ctl->disp_in = need_in ? 0.0 : disp_in;
    ctl->inner_refresh = need_in;
    return need;
}

#ifndef __CUDACC_RTC__
// Head of a force evaluation on the receiving side: wait for both neighbours' ghost columns and every rank's
// displacement bound of epoch E, take the global maximum (bit patterns of non-negative doubles order like the values;
// NaN sorts above everything and forces a rebuild, as in the single-domain path), decide the rebuild -- every rank
// reaches the same decision from the same P numbers -- and make parity E the live ghost buffer.
__global__ void k_peer_wait_head(PeerLinks lk, double skin, double skin_in, int always, DevCtl *ctl, const double4 *ghost_base_biased,
                                 CondHandles hs)
{
    const unsigned long long E = ctl->epoch + 1ull;
    const int par = (int)(E & 1ull);
    const int lane = threadIdx.x;
    unsigned long long v = 0ull;
    bool ok = true;
    if (lane < lk.nranks) {
        ok = peer_wait_flag(&lk.self.hdr->red_tag[par][lane], E, lk.timeout_ns, ctl);
        v = ok ? *(volatile unsigned long long *)&lk.self.hdr->red_val[par][lane] : 0x7ff8000000000000ull;
    }
    if (lane < 2) ok = peer_wait_flag(&lk.self.hdr->ghost_flag[lane], E, lk.timeout_ns, ctl) && ok;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    if (lane == 0) {
        ctl->dmax2_bits = 0ull;  // consumed
        const int need = skin_decide(__longlong_as_double((long long)v), 1.0, skin, skin_in, always, 0, ctl);
        ctl->epoch = E;
        ctl->gpos_m = ghost_base_biased + (size_t)par * 2 * (size_t)(1 + lk.ghost_cap);
        for (int q = 0; q < hs.n; q++) cudaGraphSetConditional(hs.h[q], need ? 1u : 0u);
    }
}
#endif

// rebuild, after k_slab_classify wrote the leavers into the neighbours' migration buffers: counts, fence, flags
__global__ void k_peer_mig_publish(DevCtl *ctl, PeerLinks lk)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const unsigned long long E = ctl->epoch;
    MigRec *out_l = lk.left.mig + (size_t)(1 + lk.mig_cap);  // arrives "from the right" at the left neighbour
    MigRec *out_r = lk.right.mig;
    out_l[0].p.x = (double)min(ctl->mig_count[0], lk.mig_cap);
    out_r[0].p.x = (double)min(ctl->mig_count[1], lk.mig_cap);
    __threadfence_system();
    st_relaxed_sys_u64(&lk.left.hdr->mig_flag[1], E);
    st_relaxed_sys_u64(&lk.right.hdr->mig_flag[0], E);
}
// which: 0 migration flags, 1 rebuild-ghost flags (both sides), epoch = ctl->epoch
__global__ void k_peer_wait(PeerLinks lk, int which, DevCtl *ctl)
{
    const unsigned long long E = ctl->epoch;
    if (threadIdx.x < 2) {
        const unsigned long long *f = which == 0 ? &lk.self.hdr->mig_flag[threadIdx.x] : &lk.self.hdr->rghost_flag[threadIdx.x];
        peer_wait_flag(f, E, lk.timeout_ns, ctl);
    }
}

// thermostat steps: this rank's {U, W, n_pairs, |v|^2} sums (ctl->red, left by k_finalize stage 1) to every rank ...
__global__ void k_peer_sum_publish(DevCtl *ctl, PeerLinks lk, int guard)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (guard && ctl->need_rebuild) return;
    const unsigned long long E = ctl->epoch;
    const int par = (int)(E & 1ull);
    for (int r = 0; r < lk.nranks; r++)
        for (int c = 0; c < 4; c++) lk.all[r]->sum_val[par][lk.me][c] = ctl->red[c];
    __threadfence_system();
    for (int r = 0; r < lk.nranks; r++) st_relaxed_sys_u64(&lk.all[r]->sum_tag[par][lk.me], E);
}
// ... and the global sums, added in rank order (the same bits on every rank), back into ctl->red for k_finalize stage 2
__global__ void k_peer_sum_wait(DevCtl *ctl, PeerLinks lk, int guard)
{
    if (guard && ctl->need_rebuild) return;
    const unsigned long long E = ctl->epoch;
    const int par = (int)(E & 1ull);
    const int lane = threadIdx.x;
    bool ok = true;
    if (lane < lk.nranks) ok = peer_wait_flag(&lk.self.hdr->sum_tag[par][lane], E, lk.timeout_ns, ctl);
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        double acc[4] = {0.0, 0.0, 0.0, 0.0};
        for (int r = 0; r < lk.nranks; r++)
            for (int c = 0; c < 4; c++) acc[c] += *(volatile double *)&lk.self.hdr->sum_val[par][r][c];
        for (int c = 0; c < 4; c++) ctl->red[c] = ok ? acc[c] : __longlong_as_double(0x7ff8000000000000ll);
    }
}

// ghost cell populations from the live parity of this rank's own mailbox
template <int DIM>
__global__ void k_peer_ghost_count(Grid g, const DevCtl *__restrict__ ctl, PeerLinks lk, uint32_t *__restrict__ cnt_l, uint32_t *__restrict__ cnt_r)
{
    const int par = (int)(ctl->epoch & 1ull);
    const double4 *gl = peer_ghost(lk.self, par, 0, lk.ghost_cap), *gr = peer_ghost(lk.self, par, 1, lk.ghost_cap);
    const int nl = (int)gl[0].x, nr = (int)gr[0].x;
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nl + nr) return;
    const double4 p = t < nl ? gl[1 + t] : gr[1 + (t - nl)];
    int cy = cell_coord(p.y, g.cinv[1], g.nc[1]);
    int cz = (DIM == 3) ? cell_coord(p.z, g.cinv[2], g.nc[2]) : 0;
    atomicAdd(&(t < nl ? cnt_l : cnt_r)[cz * g.nc[1] + cy], 1u);
}

}  // namespace mdb
