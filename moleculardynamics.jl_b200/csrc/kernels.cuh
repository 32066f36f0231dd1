// kernels.cuh -- sm_100a kernels of the per-step hot path (FP64 throughout, compiled with -fmad=false).
//
// Device layout (all resident in HBM, "slot order" = order of the last cell sort):
//   pos   double4[n]   {x, y, z, sigma}  one 32-byte sector per particle: a neighbour gather costs one sector
//   vel   double[3][cap], frc double[3][cap]   SoA, coalesced for the integrator sweeps
//   img   int32[3][cap]  periodic image counters, id int32[cap] original particle index
// Reference functions replaced are cited per kernel (paths relative to /root/reference).
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#include <cuda_runtime.h>
#endif

#include "rng.cuh"
#include "potentials.cuh"

namespace mdb {

constexpr int kForceBlock = 128;   // threads per CTA in the pair kernels
constexpr int kStreamBlock = 256;  // threads per CTA in the streaming (integrator / sort) kernels
constexpr int kQueue = 8;          // deferred-hit queue depth per thread (sparse-hit potentials)
constexpr int kMaxPartials = 4096;  // per-CTA partial slots (persistent grids are far smaller)

struct Grid {
    int nc[3];
    double cinv[3];  // nc / L
    double L[3];
    double invL[3];  // 1 / L  (IEEE division on the host == the reference's inv(U) for a diagonal cell)
    double hL[3];    // L / 2
    // x-slab decomposition (nranks > 1): this rank owns the global cell columns [c0, c0 + nxo); the two neighbouring
    // columns are ghosts held in a separate buffer.  Single domain: slab = 0, c0 = 0, nxo = nc[0] (periodic in x).
    int slab, c0, nxo;
    int kx_left, kx_right;           // image shift of the left/right ghost column (-1 / +1 when it wraps around the box)
    uint32_t g0;                     // neighbour indices >= g0 address ghosts: gpos_m[j]
    const uint32_t *gstart_l, *gstart_r;  // [ny*nz + 1] ghost index ranges per (cz,cy) row
    const double4 *gpos_m;           // ghost positions, biased so that gpos_m[g0 + k] is ghost record k
    // general (triclinic) unit cell, x = U frac with the lattice vectors in the columns of U (src/boundary.jl:7-17,
    // src/initialization.jl:7-18), row-major 3x3 (a 2-D cell is embedded with U[2][2] = 1).  tri == 0: diagonal cell,
    // the per-axis arithmetic above is used and U, Ui are ignored.  With tri != 0 the cell grid lives in fractional
    // coordinates (nc cells per lattice direction) and L holds the perpendicular widths of the cell.
    int tri;
    double U[9], Ui[9];
};

// o = M v, every row summed left to right (the oracle's order, oracle/md_oracle.c "general unit cells")
__device__ __forceinline__ void mat3_mul(const double *M, const double (&v)[3], double (&o)[3])
{
#pragma unroll
    for (int r = 0; r < 3; r++) o[r] = M[3 * r] * v[0] + M[3 * r + 1] * v[1] + M[3 * r + 2] * v[2];
}
// wrap_to_box (src/boundary.jl:7-17): frac = U^-1 x; n = floor(frac); x = U (frac - n).  ncr receives n (the image update).
// TRI template argument of the geometry helpers: -1 decide at run time from g.tri, 0 / 1 decided at compile time (the hot
// kernels are instantiated per cell type so the diagonal-cell code carries nothing of the general one)
template <int DIM, int TRI = -1>
__device__ __forceinline__ void wrap_point(const Grid &g, double (&x)[3], double (&ncr)[3])
{
    const bool tri = (TRI < 0) ? (g.tri != 0) : (TRI != 0);
    if (!tri) {
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            double frac = g.invL[k] * x[k];
            ncr[k] = floor(frac);
            x[k] = g.L[k] * (frac - ncr[k]);
        }
        if (DIM == 2) ncr[2] = 0.0;
    } else {
        double fr[3];
        if (DIM == 2) x[2] = 0.0;
        mat3_mul(g.Ui, x, fr);
#pragma unroll
        for (int k = 0; k < 3; k++) {
            ncr[k] = (k < DIM) ? floor(fr[k]) : 0.0;
            fr[k] = (k < DIM) ? fr[k] - ncr[k] : 0.0;
        }
        mat3_mul(g.U, fr, x);
    }
}
// unwrapped(p, img, boxmat) = p + boxmat * img (src/io.jl:62-70)
template <int DIM>
__device__ __forceinline__ void unwrap_point(const Grid &g, const double (&x)[3], const int32_t (&im)[3], double (&xu)[3])
{
    if (!g.tri) {
#pragma unroll
        for (int k = 0; k < 3; k++) xu[k] = (k < DIM) ? x[k] + g.L[k] * (double)im[k] : 0.0;
    } else {
        const double iv[3] = {(double)im[0], (double)im[1], (DIM == 3) ? (double)im[2] : 0.0};
        double sh[3];
        mat3_mul(g.U, iv, sh);
#pragma unroll
        for (int k = 0; k < 3; k++) xu[k] = (k < DIM) ? x[k] + sh[k] : 0.0;
    }
}
// minimum image of a separation vector: k = nearbyint(d / L) per axis, or k = nearbyint(U^-1 d), d -= U k
template <int DIM, int TRI = -1>
__device__ __forceinline__ void min_image(const Grid &g, double &dx, double &dy, double &dz)
{
    const bool tri = (TRI < 0) ? (g.tri != 0) : (TRI != 0);
    if (!tri) {
        dx = dx - nearbyint(dx * g.invL[0]) * g.L[0];
        dy = dy - nearbyint(dy * g.invL[1]) * g.L[1];
        if (DIM == 3) dz = dz - nearbyint(dz * g.invL[2]) * g.L[2];
    } else {
        const double d[3] = {dx, dy, (DIM == 3) ? dz : 0.0};
        double fr[3], sh[3];
        mat3_mul(g.Ui, d, fr);
        const double kk[3] = {nearbyint(fr[0]), nearbyint(fr[1]), (DIM == 3) ? nearbyint(fr[2]) : 0.0};
        mat3_mul(g.U, kk, sh);
        dx = dx - sh[0];
        dy = dy - sh[1];
        if (DIM == 3) dz = dz - sh[2];
    }
}
// Cartesian shift of the periodic image (kx, ky, kz) of the cell
template <int DIM, int TRI = -1>
__device__ __forceinline__ void image_shift(const Grid &g, int kx, int ky, int kz, double &sx, double &sy, double &sz)
{
    const bool tri = (TRI < 0) ? (g.tri != 0) : (TRI != 0);
    if (!tri) {
        sx = kx * g.L[0];
        sy = ky * g.L[1];
        sz = (DIM == 3) ? kz * g.L[2] : 0.0;
    } else {
        const double kk[3] = {(double)kx, (double)ky, (DIM == 3) ? (double)kz : 0.0};
        double sh[3];
        mat3_mul(g.U, kk, sh);
        sx = sh[0];
        sy = sh[1];
        sz = (DIM == 3) ? sh[2] : 0.0;
    }
}

struct StatePtrs {
    double4 *pos;
    double *vel;
    double *frc;
    int32_t *img;
    int32_t *id;
    int64_t cap;  // SoA component stride
};

// device-resident control block: lets one captured CUDA graph replay every step unchanged
struct DevCtl {
    unsigned long long step;      // steps done in the current chunk (row of thermo / ktemp arrays)
    unsigned long long rng_step;  // engine-wide step counter keying the RNG
    double alpha;                 // Bussi scale produced by the previous step, applied by the next kick (1 = none)
    double disp;                  // sum over steps of the largest per-step displacement since the last rebuild
    double last[4];               // U, W, KE, n_pairs of the most recent evaluation
    int need_rebuild;
    int list_valid;
    int max_nnbr;                 // largest neighbour count at the last list build
    int nonfinite;
    int cur;                      // which of the two state buffers is live (flipped on device by the re-sort)
    int n_overflow;               // particles whose neighbour count exceeded the list capacity at the last build
    int n_own;                    // particles owned by this handle (changes on the device when slabs migrate)
    int n_tmp;                    // owned + arrived migrants during a slab rebuild
    int mig_count[2];             // migrants packed for the left / right neighbour
    int error;                    // sticky device-side error bits (see kErr*)
    double red[4];                // slab mode: local sums of U, W, n_pairs, |v|^2 awaiting the all-reduce
    double disp_in;               // displacement bound since the inner (tight) list was last refreshed
    unsigned long long dref2_bits;  // Brownian: largest squared displacement from the positions of the last list build
    int inner_refresh;            // 1: the coming force evaluation re-derives the inner list from the outer one
    int max_nnbr_in;
    unsigned long long dmax2_bits;  // bit pattern of the largest squared displacement bound of the last move
    unsigned long long rebuilds;
    StatePtrs st[2];
    double fire[8];               // FIRE scalars: dt, alpha, steps_since_neg, converged, P, vnorm2, fnorm2, energy
    double scratch[4];            // initialize_velocities: mean per component, scale factor (setup_io.cuh)
    // Brownian parameters of the run in progress, read by the fused force + move kernel (KICK2 == 3)
    double bd_ktemp, bd_sigma;
    unsigned long long bd_seed;
    const double *bd_xref;        // unwrapped build-time positions for the exact displacement test, or null
    // x-slabs: force-evaluation heads executed so far (the epoch of the peer-memory transport, slab.cuh) and the live ghost
    // buffer, biased like Grid::gpos_m (the peer transport double-buffers the ghosts by epoch parity, so captured kernels
    // read the pointer here instead of from their by-value Grid)
    unsigned long long epoch;
    const double4 *gpos_m;
};


// ------------------------------------------------------------------------------------------------
// small helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_max(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// deterministic block reduction of NV values (fixed tree); result valid in thread 0
template <int NV, int BLOCK, bool MAX = false>
__device__ __forceinline__ void block_reduce(double (&v)[NV])
{
    __shared__ double sm[NV][BLOCK / 32];
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NV; q++) {
        v[q] = MAX ? warp_max(v[q]) : warp_sum(v[q]);
        if (lane == 0) sm[q][w] = v[q];
    }
    __syncthreads();
    if (w == 0) {
#pragma unroll
        for (int q = 0; q < NV; q++) {
            double x = (lane < BLOCK / 32) ? sm[q][lane] : (MAX ? 0.0 : 0.0);
            v[q] = MAX ? warp_max(x) : warp_sum(x);
        }
    }
}

// one 256-bit transaction per particle record (LDG.E.256 / STG.E.256 on sm_100a); pos is 32-byte aligned.
// 256-bit vector accesses need PTX ISA 8.8 (CUDA 12.9); an older NVRTC (user potentials compiled inside a process that
// already loaded a 12.8 runtime) falls back to two 128-bit accesses.
#if defined(__CUDACC_RTC__) && defined(__CUDACC_VER_MAJOR__) && (__CUDACC_VER_MAJOR__ * 100 + __CUDACC_VER_MINOR__ < 1209)
#define MDB_VEC256 0
#else
#define MDB_VEC256 1
#endif
__device__ __forceinline__ double4 ldg_pos(const double4 *p)
{
    double4 r;
#if MDB_VEC256
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
#else
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
    asm volatile("ld.global.nc.v2.f64 {%0,%1}, [%2+16];" : "=d"(r.z), "=d"(r.w) : "l"(p));
#endif
    return r;
}
__device__ __forceinline__ double4 ld_pos(const double4 *p)
{
    double4 r;
#if MDB_VEC256
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
#else
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p) : "memory");
    asm volatile("ld.global.v2.f64 {%0,%1}, [%2+16];" : "=d"(r.z), "=d"(r.w) : "l"(p) : "memory");
#endif
    return r;
}
__device__ __forceinline__ void st_pos(double4 *p, const double4 &v)
{
#if MDB_VEC256
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
#else
    asm volatile("st.global.v2.f64 [%0], {%1,%2};" ::"l"(p), "d"(v.x), "d"(v.y) : "memory");
    asm volatile("st.global.v2.f64 [%0+16], {%1,%2};" ::"l"(p), "d"(v.z), "d"(v.w) : "memory");
#endif
}

__device__ __forceinline__ int cell_coord(double x, double cinv, int nc)
{
    int c = (int)(x * cinv);
    return c < nc - 1 ? (c < 0 ? 0 : c) : nc - 1;
}
// cell of a (wrapped) position: per axis for a diagonal cell, from the fractional coordinates otherwise
template <int DIM, int TRI = -1>
__device__ __forceinline__ void cell_of_point(const Grid &g, const double4 &p, int &cx, int &cy, int &cz)
{
    const bool tri = (TRI < 0) ? (g.tri != 0) : (TRI != 0);
    if (!tri) {
        cx = cell_coord(p.x, g.cinv[0], g.nc[0]);
        cy = cell_coord(p.y, g.cinv[1], g.nc[1]);
        cz = (DIM == 3) ? cell_coord(p.z, g.cinv[2], g.nc[2]) : 0;
    } else {
        const double x[3] = {p.x, p.y, (DIM == 3) ? p.z : 0.0};
        double fr[3];
        mat3_mul(g.Ui, x, fr);
        cx = cell_coord(fr[0], (double)g.nc[0], g.nc[0]);
        cy = cell_coord(fr[1], (double)g.nc[1], g.nc[1]);
        cz = (DIM == 3) ? cell_coord(fr[2], (double)g.nc[2], g.nc[2]) : 0;
    }
}

// ------------------------------------------------------------------------------------------------
// K0  upload: AoS host image -> device layout, with wrap_to_box (src/boundary.jl:7-17)
// ------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void k_import(int64_t n, const double *__restrict__ x, const double *__restrict__ v, const double *__restrict__ f,
                         const double *__restrict__ diam, const int32_t *__restrict__ im, const int32_t *__restrict__ ids,
                         StatePtrs s, Grid g)
{
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double p[3] = {0.0, 0.0, 0.0};
    int32_t m[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        p[k] = x[i * DIM + k];
        m[k] = im ? im[i * DIM + k] : 0;
    }
    if (!g.tri) {
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            double xv = p[k];
            if (xv < 0.0 || xv >= g.L[k]) {
                double frac = g.invL[k] * xv;
                double ncr = floor(frac);
                m[k] += (int32_t)ncr;
                xv = g.L[k] * (frac - ncr);
            }
            p[k] = xv;
        }
    } else {  // a general cell is always passed through wrap_to_box (wrapping a point inside the cell leaves n = 0)
        double ncr[3];
        wrap_point<DIM>(g, p, ncr);
#pragma unroll
        for (int k = 0; k < DIM; k++) m[k] += (int32_t)ncr[k];
    }
    s.pos[i] = make_double4(p[0], p[1], p[2], diam[i]);
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        s.vel[k * s.cap + i] = v ? v[i * DIM + k] : 0.0;
        s.frc[k * s.cap + i] = f ? f[i * DIM + k] : 0.0;
        s.img[k * s.cap + i] = m[k];
    }
    s.id[i] = ids ? ids[i] : (int32_t)i;
}

template <int DIM>
__global__ void k_import_vel(int64_t n, const double *__restrict__ v, const DevCtl *__restrict__ ctl)
{
    const StatePtrs s = ctl->st[ctl->cur];
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int32_t o = s.id[i];
#pragma unroll
    for (int k = 0; k < DIM; k++) s.vel[k * s.cap + i] = v[(int64_t)o * DIM + k];
}

// device layout -> AoS host image in ORIGINAL particle order (scatter through id)
template <int DIM>
__global__ void k_export(int64_t n, const DevCtl *__restrict__ ctl, double *__restrict__ x, double *__restrict__ v, double *__restrict__ f,
                         int32_t *__restrict__ im, int to_original)
{
    const StatePtrs s = ctl->st[ctl->cur];
    if (n < 0) n = ctl->n_own;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t o = to_original ? s.id[i] : i;
    double4 p = s.pos[i];
    double pp[3] = {p.x, p.y, p.z};
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        if (x) x[o * DIM + k] = pp[k];
        if (v) v[o * DIM + k] = s.vel[k * s.cap + i];
        if (f) f[o * DIM + k] = s.frc[k * s.cap + i];
        if (im) im[o * DIM + k] = s.img[k * s.cap + i];
    }
}

// ------------------------------------------------------------------------------------------------
// K1-K3  cell list build: hash + count, scan, fill, canonical in-cell order, gather-reorder.
// Replaces CellListMap's UpdateCellList! inside map_pairwise! (src/simulation.jl:100).
// ------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void k_hash(int64_t n, const DevCtl *__restrict__ ctl, Grid g, uint32_t *__restrict__ cell_of,
                       uint32_t *__restrict__ slot_of, uint32_t *__restrict__ counts)
{
    if (n < 0) n = ctl->n_own;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double4 p = ctl->st[ctl->cur].pos[i];
    int cx, cy, cz;
    cell_of_point<DIM>(g, p, cx, cy, cz);
    uint32_t c = ((uint32_t)cz * g.nc[1] + cy) * g.nxo + (cx - g.c0);
    cell_of[i] = c;
    slot_of[i] = atomicAdd(&counts[c], 1u);
}

constexpr int kScanItems = 8;
constexpr int kScanTile = kStreamBlock * kScanItems;

__global__ void k_scan_tile_sums(int64_t n, const uint32_t *__restrict__ in, uint32_t *__restrict__ tile_sums)
{
    __shared__ uint32_t sm[kStreamBlock / 32];
    int64_t base = (int64_t)blockIdx.x * kScanTile;
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        int64_t q = base + k * kStreamBlock + threadIdx.x;
        if (q < n) s += in[q];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t t = 0;
        for (int w = 0; w < kStreamBlock / 32; w++) t += sm[w];
        tile_sums[blockIdx.x] = t;
    }
}

// single CTA: exclusive scan of the tile sums in place
__global__ void k_scan_tiles(int ntiles, uint32_t *__restrict__ tile_sums)
{
    __shared__ uint32_t sm[1024];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < ntiles; base += 1024) {
        int q = base + threadIdx.x;
        uint32_t v = q < ntiles ? tile_sums[q] : 0;
        sm[threadIdx.x] = v;
        __syncthreads();
        for (int o = 1; o < 1024; o <<= 1) {
            uint32_t t = threadIdx.x >= o ? sm[threadIdx.x - o] : 0;
            __syncthreads();
            sm[threadIdx.x] += t;
            __syncthreads();
        }
        uint32_t incl = sm[threadIdx.x];
        if (q < ntiles) tile_sums[q] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
}

// exclusive scan within each tile + tile offset; writes start[0..n] (start[n] = total)
__global__ void k_scan_apply(int64_t n, const uint32_t *__restrict__ in, const uint32_t *__restrict__ tile_off,
                             uint32_t *__restrict__ start)
{
    __shared__ uint32_t sm[kStreamBlock / 32];
    int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    // exclusive scan of s over the CTA
    uint32_t incl = s;
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) sm[w] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (int q = 0; q < w; q++) woff += sm[q];
    uint32_t run = tile_off[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) start[base + k] = run;
        run += v[k];
        if (base + k == n - 1) start[n] = run;
    }
}

constexpr uint32_t kInvalidCell = 0xffffffffu;  // particle that left this slab (migrated away)

__global__ void k_fill(int64_t n, const DevCtl *__restrict__ ctl, const uint32_t *__restrict__ cell_of,
                       const uint32_t *__restrict__ slot_of, const uint32_t *__restrict__ start, uint32_t *__restrict__ order)
{
    if (n < 0) n = ctl->n_tmp;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t c = cell_of[i];
    if (c != kInvalidCell) order[start[c] + slot_of[i]] = (uint32_t)i;
}

// canonical (ascending previous-slot) order inside each cell: removes the atomic-arrival nondeterminism,
// so the whole pipeline is a stable counting sort and reruns are bit-identical
// by_id: order by original particle id instead of previous slot (slab mode: migrants arrive in arbitrary order)
__global__ void k_cellsort(int64_t ncell, const uint32_t *__restrict__ start, uint32_t *__restrict__ order, int by_id,
                           const DevCtl *__restrict__ ctl)
{
    int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (c >= ncell) return;
    uint32_t b = start[c], e = start[c + 1];
    if (!by_id) {
        for (uint32_t a = b + 1; a < e; a++) {
            uint32_t key = order[a];
            uint32_t q = a;
            while (q > b && order[q - 1] > key) {
                order[q] = order[q - 1];
                q--;
            }
            order[q] = key;
        }
    } else {
        const int32_t *__restrict__ id = ctl->st[ctl->cur].id;
        for (uint32_t a = b + 1; a < e; a++) {
            uint32_t key = order[a];
            int32_t kid = id[key];
            uint32_t q = a;
            while (q > b && id[order[q - 1]] > kid) {
                order[q] = order[q - 1];
                q--;
            }
            order[q] = key;
        }
    }
}

// posf (optional): single-precision shadow of the re-sorted positions, read by k_build_list_f32
template <int DIM>
__global__ void k_gather(int64_t n, const uint32_t *__restrict__ order, const DevCtl *__restrict__ ctl,
                         const uint32_t *__restrict__ n_new, float4 *__restrict__ posf = nullptr)
{
    if (n < 0) n = *n_new;  // slab mode: the owned count after migration = start[number of owned cells]
    const StatePtrs src = ctl->st[ctl->cur], dst = ctl->st[ctl->cur ^ 1];
    int64_t p = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t s = order[p];
    const double4 rec = src.pos[s];
    dst.pos[p] = rec;
    if (posf) posf[p] = make_float4((float)rec.x, (float)rec.y, (float)rec.z, (float)rec.w);
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        dst.vel[k * dst.cap + p] = src.vel[k * src.cap + s];
        dst.frc[k * dst.cap + p] = src.frc[k * src.cap + s];
        dst.img[k * dst.cap + p] = src.img[k * src.cap + s];
    }
    dst.id[p] = src.id[s];
}

// ------------------------------------------------------------------------------------------------
// neighbour-cell traversal shared by the pair-force kernel, the list builder and the pair counter.
// Rows of three x-adjacent cells are contiguous slot ranges (cells are numbered x-fastest).
// Minimum image: dx = (xi - xj) - k*L with k in {-1,0,1} given by the periodic wrap of the cell row
// (same value as the oracle's k = nearbyint(dx/L) for every pair within the search radius).
// visit(j, dx, dy, dz, d2, sigma_j, code) is called for every candidate j != i with d2 <= r2.
// ------------------------------------------------------------------------------------------------
// ALWAYS: visit is called for EVERY candidate, with d2 = -1 for the ones that fail the test, so that a visitor can stay
// branch-free (the list build: a divergent accept path was half of its instructions at 5 of 32 lanes)
template <int DIM, int TRI = -1, bool ALWAYS = false, class Visit>
__device__ __forceinline__ void traverse_cells(const Grid &g, const uint32_t *__restrict__ start,
                                               const double4 *__restrict__ pos, int i, const double4 &pi, int cx, int cy,
                                               int cz, double r2, Visit &&visit)
{
    const int nx = g.nxo, ny = g.nc[1], nz = g.nc[2];
    const int lx = cx - g.c0;
    for (int dz = (DIM == 3 ? -1 : 0); dz <= (DIM == 3 ? 1 : 0); dz++) {
        int oz = cz + dz;
        int kz = 0;
        if (DIM == 3) {
            if (oz < 0) { oz += nz; kz = -1; }
            else if (oz >= nz) { oz -= nz; kz = 1; }
        }
        for (int dy = -1; dy <= 1; dy++) {
            int oy = cy + dy;
            int ky = 0;
            if (oy < 0) { oy += ny; ky = -1; }
            else if (oy >= ny) { oy -= ny; ky = 1; }
            const uint32_t rowg = (uint32_t)oz * ny + oy;
            const uint32_t row = rowg * nx;
            // up to two contiguous segments along x
            uint32_t seg_b[2], seg_e[2];
            int seg_k[2], nseg = 1;
            const double4 *seg_p[2] = {pos, pos};
            if (!g.slab) {
                if (lx == 0) {
                    seg_b[0] = start[row]; seg_e[0] = start[row + 2]; seg_k[0] = 0;
                    seg_b[1] = start[row + nx - 1]; seg_e[1] = start[row + nx]; seg_k[1] = -1;
                    nseg = 2;
                } else if (lx == nx - 1) {
                    seg_b[0] = start[row + nx - 2]; seg_e[0] = start[row + nx]; seg_k[0] = 0;
                    seg_b[1] = start[row]; seg_e[1] = start[row + 1]; seg_k[1] = 1;
                    nseg = 2;
                } else {
                    seg_b[0] = start[row + lx - 1]; seg_e[0] = start[row + lx + 2]; seg_k[0] = 0;
                }
            } else {
                const int lo = lx > 0 ? lx - 1 : 0, hi = lx < nx - 1 ? lx + 1 : nx - 1;
                seg_b[0] = start[row + lo]; seg_e[0] = start[row + hi + 1]; seg_k[0] = 0;
                if (lx == 0) {
                    seg_b[1] = g.gstart_l[rowg]; seg_e[1] = g.gstart_l[rowg + 1]; seg_k[1] = g.kx_left; seg_p[1] = g.gpos_m;
                    nseg = 2;
                } else if (lx == nx - 1) {
                    seg_b[1] = g.gstart_r[rowg]; seg_e[1] = g.gstart_r[rowg + 1]; seg_k[1] = g.kx_right; seg_p[1] = g.gpos_m;
                    nseg = 2;
                }
            }
            for (int sgi = 0; sgi < nseg; sgi++) {
                const uint32_t jb = seg_b[sgi], je = seg_e[sgi];
                const int kx = seg_k[sgi];
                const double4 *__restrict__ src = seg_p[sgi];
                const int code = (kx + 1) + 3 * (ky + 1) + 9 * (kz + 1);
                if (code == 13) {
                    for (uint32_t j = jb; j < je; j++) {
                        double4 pj = ldg_pos(&src[j]);
                        double dx = pi.x - pj.x, dy_ = pi.y - pj.y;
                        double d2 = fma(dy_, dy_, dx * dx);
                        double dz_ = 0.0;
                        if (DIM == 3) {
                            dz_ = pi.z - pj.z;
                            d2 = fma(dz_, dz_, d2);
                        }
                        if (ALWAYS) visit((int)j, dx, dy_, dz_, (d2 <= r2 && (int)j != i) ? d2 : -1.0, pj.w, code);
                        else if (d2 <= r2 && (int)j != i) visit((int)j, dx, dy_, dz_, d2, pj.w, code);
                    }
                } else {
                    double sx, sy, sz;
                    image_shift<DIM, TRI>(g, kx, ky, kz, sx, sy, sz);
                    for (uint32_t j = jb; j < je; j++) {
                        double4 pj = ldg_pos(&src[j]);
                        double dx = (pi.x - pj.x) - sx, dy_ = (pi.y - pj.y) - sy;
                        double d2 = fma(dy_, dy_, dx * dx);
                        double dz_ = 0.0;
                        if (DIM == 3) {
                            dz_ = (pi.z - pj.z) - sz;
                            d2 = fma(dz_, dz_, d2);
                        }
                        if (ALWAYS) visit((int)j, dx, dy_, dz_, (d2 <= r2 && (int)j != i) ? d2 : -1.0, pj.w, code);
                        else if (d2 <= r2 && (int)j != i) visit((int)j, dx, dy_, dz_, d2, pj.w, code);
                    }
                }
            }
        }
    }
}

// owned neighbours live in the (double-buffered) state, ghosts of an x-slab in their own receive buffer
__device__ __forceinline__ const double4 *nbr_ptr(const Grid &g, const double4 *pos, uint32_t j)
{
    return j >= g.g0 ? g.gpos_m + j : pos + j;
}

// per-pair update: src/pairwise.jl:26-39 (one side of it: the gather evaluates each pair from both ends,
// the antisymmetric halves are bit-exact negatives, energy/virial/pair-count totals are halved at the end)
// pair_force: the vector s = f*r/d of one pair and its energy / virial / count terms
template <int DIM, class Pot>
__device__ __forceinline__ void pair_force(const Pot &pot, const PotParams &pp, double dx, double dy, double dz, double d2, double si,
                                           double sj, double &sx, double &sy, double &sz, double &e, double &w, double &np)
{
    double d = sqrt(d2);
    double u, f;
    bool in;
    sz = 0.0;
    if constexpr (has_rcp_eval<Pot>::value) {
        const double r = rcp_refined(d);  // (f*r_k)/d and sigma/d share one reciprocal: same bits as four IEEE divisions
        in = pot.eval_rcp(pp, d, r, si, sj, u, f);
        sx = div_by(f * dx, d, r);
        sy = div_by(f * dy, d, r);
        if (DIM == 3) sz = div_by(f * dz, d, r);
    } else {
        in = pot.eval(pp, d, si, sj, u, f);
        sx = (f * dx) / d;
        sy = (f * dy) / d;
        if (DIM == 3) sz = (f * dz) / d;
    }
    double dot = sx * dx + sy * dy;
    if (DIM == 3) dot += sz * dz;
    w += dot;
    e += u;
    np += in ? 1.0 : 0.0;
}
template <int DIM, class Pot>
__device__ __forceinline__ void pair_accumulate(const Pot &pot, const PotParams &pp, double dx, double dy, double dz, double d2,
                                                double si, double sj, double (&F)[3], double &e, double &w, double &np)
{
    double sx, sy, sz;
    pair_force<DIM>(pot, pp, dx, dy, dz, d2, si, sj, sx, sy, sz, e, w, np);
    F[0] += sx;
    F[1] += sy;
    if (DIM == 3) F[2] += sz;
}

struct ForceOut {
    double *part;  // [4][kMaxPartials]: per-CTA sums of e_i, w_i, n_i, |v_i|^2
};

// per-thread running sums across the tiles a persistent CTA walks
struct ThreadSums {
    double e = 0.0, w = 0.0, np = 0.0, v2 = 0.0;
};

// per-particle epilogue shared by the force kernels: store f, optional second half kick (src/integrate.jl:28-38;
// x*0.5 == x/2.0 bit-for-bit) with the kinetic-energy term (src/thermostat.jl:50-60).
template <int DIM, bool KICK2>
__device__ __forceinline__ void particle_epilogue(int i, const double (&F)[3], const StatePtrs &s, double dt, ThreadSums &acc)
{
#pragma unroll
    for (int k = 0; k < DIM; k++) s.frc[k * s.cap + i] = F[k];
    if (KICK2) {
        double v2 = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            double v = s.vel[k * s.cap + i];
            v += (F[k] * dt) * 0.5;
            s.vel[k * s.cap + i] = v;
            v2 = (k == 0) ? v * v : v2 + v * v;
        }
        acc.v2 += v2;
    }
}

// Fused NVE step (KICK2 == 2): second half kick of this step, then the NEXT step's first half kick, drift and wrap
// (k_kick_drift's arithmetic, operation for operation: src/integrate.jl:8-38, src/boundary.jl:7-17) while F, v and x are
// still in registers.  The moved position goes to the other position buffer (neighbours still read this step's
// positions from the live one); k_finalize swaps the two pointers afterwards.  Saves the whole K5 sweep of the next step.
template <int DIM, int TRI = -1>
__device__ __forceinline__ void leap_epilogue(int i, const double (&F)[3], const double (&vel)[3], const double4 &pi, const StatePtrs &s,
                                              double4 *__restrict__ pos_next, const Grid &g, double dt, double &ke2, double &vmax2)
{
    double x[3] = {pi.x, pi.y, pi.z};
    double v2 = 0.0, w2 = 0.0;
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        const double h = (F[k] * dt) * 0.5;
        double v = vel[k] + h;  // v(t + dt): the kinetic energy of this step (src/thermostat.jl:50-60)
        v2 = (k == 0) ? v * v : v2 + v * v;
        v += h;                 // next step's first half kick (alpha == 1 in NVE)
        s.vel[k * s.cap + i] = v;
        w2 = (k == 0) ? v * v : w2 + v * v;
        x[k] = x[k] + v * dt;
    }
    double ncr[3];
    wrap_point<DIM, TRI>(g, x, ncr);
#pragma unroll
    for (int k = 0; k < DIM; k++)
        if (ncr[k] != 0.0) s.img[k * s.cap + i] += (int32_t)ncr[k];
    st_pos(&pos_next[i], make_double4(x[0], x[1], x[2], pi.w));
    ke2 += v2;
    vmax2 = fmax(vmax2, w2);
}
// displacement bound of the fused move, folded like k_kick_drift does
template <int BLOCK>
__device__ __forceinline__ void leap_report(double vmax2, double dt, DevCtl *ctl)
{
    double r[1] = {vmax2 * (dt * dt)};
    block_reduce<1, BLOCK, true>(r);
    if (threadIdx.x == 0) atomicMax(&ctl->dmax2_bits, (unsigned long long)__double_as_longlong(r[0]));
}

// Fused Brownian step (KICK2 == 3): the move x += f*dt/kT + noise*sigma, wrap (k_brownian's arithmetic, operation for
// operation: src/integrate.jl:66-82) while F and x are in registers; the moved position goes to the other position
// buffer like in the fused NVE step.
template <int DIM, int TRI = -1>
__device__ __forceinline__ void brown_epilogue(int i, const double (&F)[3], const double4 &pi, const StatePtrs &s,
                                               double4 *__restrict__ pos_next, const Grid &g, double dt, const DevCtl *__restrict__ ctl,
                                               unsigned long long rng_step, double &dmax2, double &dref2)
{
    const double ktemp = ctl->bd_ktemp, sigma = ctl->bd_sigma;
    const double *__restrict__ xref = ctl->bd_xref;
    double noise[3];
    brownian_noise<DIM>(ctl->bd_seed, rng_step, (uint32_t)s.id[i], noise);
    double x[3] = {pi.x, pi.y, pi.z};
    double d2 = 0.0, r2 = 0.0;
    double ncrv[3];
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        double xv = x[k] + (F[k] * dt / ktemp) + (noise[k] * sigma);
        double del = xv - x[k];
        d2 = (k == 0) ? del * del : d2 + del * del;
        x[k] = xv;
    }
    wrap_point<DIM, TRI>(g, x, ncrv);
#pragma unroll
    for (int k = 0; k < DIM; k++) {
        const double ncr = ncrv[k];
        if (xref) {
            int32_t im = s.img[k * s.cap + i];
            if (ncr != 0.0) {
                im += (int32_t)ncr;
                s.img[k * s.cap + i] = im;
            }
            double dr = (x[k] + g.L[k] * (double)im) - xref[k * s.cap + i];
            r2 = (k == 0) ? dr * dr : r2 + dr * dr;
        } else if (ncr != 0.0) {
            s.img[k * s.cap + i] += (int32_t)ncr;
        }
    }
    st_pos(&pos_next[i], make_double4(x[0], x[1], x[2], pi.w));
    dmax2 = fmax(dmax2, d2);
    dref2 = fmax(dref2, r2);
}
template <int BLOCK>
__device__ __forceinline__ void brown_report(double dmax2, double dref2, DevCtl *ctl)
{
    double r[2] = {dmax2, dref2};
    block_reduce<2, BLOCK, true>(r);
    if (threadIdx.x == 0) {
        atomicMax(&ctl->dmax2_bits, (unsigned long long)__double_as_longlong(r[0]));
        if (ctl->bd_xref) atomicMax(&ctl->dref2_bits, (unsigned long long)__double_as_longlong(r[1]));
    }
}

// one deterministic CTA reduction at the end of the persistent loop
__device__ __forceinline__ void cta_epilogue(const ThreadSums &acc, ForceOut out, int slot)
{
    double r[4] = {acc.e, acc.w, acc.np, acc.v2};
    block_reduce<4, kForceBlock>(r);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 4; q++) out.part[q * kMaxPartials + slot] = r[q];
    }
}

// ------------------------------------------------------------------------------------------------
// K4a  pair forces straight from the cell list (MDB_MODE_CELLS: the reference's rebuild-every-step shape)
// Replaces map_pairwise! + energy_and_forces! + evaluate + reducer (src/pairwise.jl:17-39, src/potentials.jl).
// Persistent CTAs walk 128-particle tiles; candidates that pass the cutoff and the potential's conservative range
// test are parked in a per-thread shared-memory queue and evaluated together at the end of the tile, so the rare
// sqrt/divide path (about 1 interacting neighbour out of ~26 candidates for PseudoHS) is not replayed per candidate.
// ------------------------------------------------------------------------------------------------
template <int DIM, class Pot, bool KICK2>
__global__ void __launch_bounds__(kForceBlock)
k_force_cells(int n, const DevCtl *__restrict__ ctl, Grid g, const uint32_t *__restrict__ start, double cutoff2, Pot pot, PotParams pp, double dt,
              ForceOut out, int guard)
{
    // guard: launched speculatively behind the rebuild decision (slab steps); a pending rebuild turns the launch into a no-op
    if (guard && ctl->need_rebuild) return;
    if (g.slab) g.gpos_m = ctl->gpos_m;
    __shared__ uint32_t queue[kQueue][kForceBlock];
    const StatePtrs s = ctl->st[ctl->cur];
    const double4 *__restrict__ pos = s.pos;
    ThreadSums acc;
    if (n < 0) n = ctl->n_own;
    const int ntiles = (n + kForceBlock - 1) / kForceBlock;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * kForceBlock + threadIdx.x;
        const bool active = i < n;
        double F[3] = {0.0, 0.0, 0.0};
        double4 pi = make_double4(0, 0, 0, 1);
        int nq = 0;
        auto drain_one = [&]() {
            if (nq > 0) {
                uint32_t ent = queue[--nq][threadIdx.x];
                int j = (int)(ent & 0x7ffffffu);
                int code = (int)(ent >> 27);
                int kx = code % 3 - 1, ky = (code / 3) % 3 - 1, kz = code / 9 - 1;
                double4 pj = ldg_pos(nbr_ptr(g, pos, (uint32_t)j));
                double sx, sy, sz;
                image_shift<DIM>(g, kx, ky, kz, sx, sy, sz);
                double dx = (pi.x - pj.x) - sx, dy = (pi.y - pj.y) - sy;
                double d2 = fma(dy, dy, dx * dx), dz = 0.0;
                if (DIM == 3) {
                    dz = (pi.z - pj.z) - sz;
                    d2 = fma(dz, dz, d2);
                }
                pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, acc.e, acc.w, acc.np);
            }
        };
        if (active) {
            pi = pos[i];
            int cx, cy, cz;
            cell_of_point<DIM>(g, pi, cx, cy, cz);
            traverse_cells<DIM>(g, start, pos, i, pi, cx, cy, cz, cutoff2,
                                [&](int j, double dx, double dy, double dz, double d2, double sj, int code) {
                                    if (pot.may_interact(pp, d2, pi.w, sj)) {
                                        if (Pot::kSparseHits) {
                                            queue[nq++][threadIdx.x] = (uint32_t)j | ((uint32_t)code << 27);
                                            if (nq == kQueue) {
                                                while (nq > 0) drain_one();
                                            }
                                        } else {
                                            pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, sj, F, acc.e, acc.w, acc.np);
                                        }
                                    }
                                });
        }
        // all lanes drain together: iterations = deepest queue in the warp
        if (Pot::kSparseHits) {
            while (__any_sync(0xffffffffu, nq > 0)) drain_one();
        }
        if (active) particle_epilogue<DIM, KICK2>(i, F, s, dt, acc);
    }
    cta_epilogue(acc, out, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
// K4b  Verlet list build (r_list = r_search + skin) and list-driven pair forces (MDB_MODE_LIST)
// nl is column-major: nl[k * stride + i], coalesced across the warp.  Particles with more than kmax neighbours are
// appended to the overflow list and handled exactly by k_force_overflow.
// ------------------------------------------------------------------------------------------------
template <int DIM, bool TRI = false>
__global__ void __launch_bounds__(kForceBlock)
k_build_list(int n, Grid g, const uint32_t *__restrict__ start, double rlist2,
             uint32_t *__restrict__ nl, int64_t stride, int kmax, int32_t *__restrict__ nnbr, uint32_t *__restrict__ ovf, DevCtl *ctl,
             double *__restrict__ xref)
{
    const StatePtrs st = ctl->st[ctl->cur];
    const double4 *__restrict__ pos = st.pos;
    if (g.slab) g.gpos_m = ctl->gpos_m;
    if (n < 0) n = ctl->n_own;
    int i = blockIdx.x * kForceBlock + threadIdx.x;
    int cnt = 0;
    if (i < n) {
        double4 pi = pos[i];
        if (xref) {  // unwrapped position at build time: reference for the exact displacement test of Brownian runs
            const double pk[3] = {pi.x, pi.y, pi.z};
#pragma unroll
            for (int k = 0; k < DIM; k++) xref[k * st.cap + i] = pk[k] + g.L[k] * (double)st.img[k * st.cap + i];  // (diagonal cells only)
        }
        int cx, cy, cz;
        cell_of_point<DIM, TRI ? 1 : 0>(g, pi, cx, cy, cz);
        uint32_t *slot = nl + i;  // next free row of this particle's column
        traverse_cells<DIM, TRI ? 1 : 0, true>(g, start, pos, i, pi, cx, cy, cz, rlist2,
                                               [&](int j, double, double, double, double d2, double, int) {
                                                   const bool hit = d2 >= 0.0;
                                                   if (hit && cnt < kmax) *slot = (uint32_t)j;
                                                   slot += hit ? stride : 0;
                                                   cnt += hit ? 1 : 0;
                                               });
        nnbr[i] = cnt;
        if (cnt > kmax) ovf[atomicAdd(&ctl->n_overflow, 1)] = (uint32_t)i;
    }
    int m = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&ctl->max_nnbr, m);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl->list_valid = 1;
        ctl->disp = 0.0;
        ctl->dref2_bits = 0ull;
    }
}

// K4b (single-precision membership): the Verlet list only has to be a SUPERSET of the pairs within r_list -- every decision
// that matters (inner-list membership, cutoff, potential range) is re-taken in FP64 by the force kernel -- so the build may
// test membership on single-precision copies of the positions against a radius padded by the worst-case rounding error
// (rl2f = (r_list^2 + 2*sqrt(3)*r_list*eps + 3*eps^2)(1 + 1e-5), eps = 4 L 2^-24; the engine falls back to k_build_list when
// the pad is not small against the skin).  Why: ncu shows k_build_list bound by L1 wavefronts -- at every iteration the 32
// lanes of a warp load 32 different 32-byte records spread over ~10 cache lines; 16-byte float4 records halve the lines
// (and the FP64 pipe drops out of the loop).  Same traversal order as k_build_list, so list order -- and with it the order
// of every force sum -- is unchanged; the list may hold a few extra candidates within ~1e-4 of r_list.
template <int DIM>
__global__ void __launch_bounds__(kForceBlock)
k_build_list_f32(int n, Grid g, const uint32_t *__restrict__ start, float rl2f, const float4 *__restrict__ posf,
                 uint32_t *__restrict__ nl, int64_t stride, int kmax, int32_t *__restrict__ nnbr, uint32_t *__restrict__ ovf, DevCtl *ctl,
                 double *__restrict__ xref)
{
    const StatePtrs st = ctl->st[ctl->cur];
    if (g.slab) g.gpos_m = ctl->gpos_m;
    if (n < 0) n = ctl->n_own;
    const int i = blockIdx.x * kForceBlock + threadIdx.x;
    int cnt = 0;
    if (i < n) {
        const double4 pd = st.pos[i];
        if (xref) {
            const double pk[3] = {pd.x, pd.y, pd.z};
#pragma unroll
            for (int k = 0; k < DIM; k++) xref[k * st.cap + i] = pk[k] + g.L[k] * (double)st.img[k * st.cap + i];
        }
        int cx, cy, cz;
        cell_of_point<DIM, 0>(g, pd, cx, cy, cz);
        const float xi = (float)pd.x, yi = (float)pd.y, zi = (float)pd.z;
        const float Lx = (float)g.L[0], Ly = (float)g.L[1], Lz = (float)g.L[2];
        uint32_t *slot = nl + i;
        const int nx = g.nxo, ny = g.nc[1], nz = g.nc[2];
        const int lx = cx - g.c0;
        for (int dz = (DIM == 3 ? -1 : 0); dz <= (DIM == 3 ? 1 : 0); dz++) {
            int oz = cz + dz, kz = 0;
            if (DIM == 3) {
                if (oz < 0) { oz += nz; kz = -1; }
                else if (oz >= nz) { oz -= nz; kz = 1; }
            }
            for (int dy = -1; dy <= 1; dy++) {
                int oy = cy + dy, ky = 0;
                if (oy < 0) { oy += ny; ky = -1; }
                else if (oy >= ny) { oy -= ny; ky = 1; }
                const uint32_t rowg = (uint32_t)oz * ny + oy;
                const uint32_t row = rowg * nx;
                uint32_t seg_b[2], seg_e[2];
                int seg_k[2], nseg = 1;
                bool seg_ghost[2] = {false, false};
                if (!g.slab) {
                    if (lx == 0) {
                        seg_b[0] = start[row]; seg_e[0] = start[row + 2]; seg_k[0] = 0;
                        seg_b[1] = start[row + nx - 1]; seg_e[1] = start[row + nx]; seg_k[1] = -1;
                        nseg = 2;
                    } else if (lx == nx - 1) {
                        seg_b[0] = start[row + nx - 2]; seg_e[0] = start[row + nx]; seg_k[0] = 0;
                        seg_b[1] = start[row]; seg_e[1] = start[row + 1]; seg_k[1] = 1;
                        nseg = 2;
                    } else {
                        seg_b[0] = start[row + lx - 1]; seg_e[0] = start[row + lx + 2]; seg_k[0] = 0;
                    }
                } else {
                    const int lo = lx > 0 ? lx - 1 : 0, hi = lx < nx - 1 ? lx + 1 : nx - 1;
                    seg_b[0] = start[row + lo]; seg_e[0] = start[row + hi + 1]; seg_k[0] = 0;
                    if (lx == 0) {
                        seg_b[1] = g.gstart_l[rowg]; seg_e[1] = g.gstart_l[rowg + 1]; seg_k[1] = g.kx_left; seg_ghost[1] = true;
                        nseg = 2;
                    } else if (lx == nx - 1) {
                        seg_b[1] = g.gstart_r[rowg]; seg_e[1] = g.gstart_r[rowg + 1]; seg_k[1] = g.kx_right; seg_ghost[1] = true;
                        nseg = 2;
                    }
                }
                for (int sgi = 0; sgi < nseg; sgi++) {
                    const uint32_t jb = seg_b[sgi], je = seg_e[sgi];
                    // the shifted own position: (xi - sx) - xj; the same two roundings whichever of the pair does the test
                    const float sx = (float)seg_k[sgi] * Lx, sy = (float)ky * Ly, sz = (float)kz * Lz;
                    const bool ghost = seg_ghost[sgi];
                    for (uint32_t j = jb; j < je; j++) {
                        float xj, yj, zj;
                        if (!ghost) {
                            const float4 pj = __ldg(&posf[j]);
                            xj = pj.x; yj = pj.y; zj = pj.z;
                        } else {  // ghost records live in the mailbox as FP64
                            const double4 pj = ldg_pos(&g.gpos_m[j]);
                            xj = (float)pj.x; yj = (float)pj.y; zj = (float)pj.z;
                        }
                        const float dx = (xi - xj) - sx, dy_ = (yi - yj) - sy;
                        float d2 = dx * dx + dy_ * dy_;
                        if (DIM == 3) {
                            const float dz_ = (zi - zj) - sz;
                            d2 += dz_ * dz_;
                        }
                        const bool hit = d2 <= rl2f && (int)j != i;
                        if (hit && cnt < kmax) *slot = j;
                        slot += hit ? stride : 0;
                        cnt += hit ? 1 : 0;
                    }
                }
            }
        }
        nnbr[i] = cnt;
        if (cnt > kmax) ovf[atomicAdd(&ctl->n_overflow, 1)] = (uint32_t)i;
    }
    int m = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0) atomicMax(&ctl->max_nnbr, m);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl->list_valid = 1;
        ctl->disp = 0.0;
        ctl->dref2_bits = 0ull;
    }
}

// K4b (warp-cooperative): k_build_list_f32 with the candidates of a warp staged ONCE per neighbour row.  The 32 particles of a
// warp are consecutive slots, i.e. ~20 consecutive cells of one x-row; for each of the 9 neighbour rows their candidate
// ranges overlap almost entirely (union ~34 records), yet k_build_list_f32 loads them lane by lane, 42 scattered 16-byte
// loads per lane.  Here the warp loads the union window with one or two coalesced loads into its shared-memory buffer and
// every lane walks its own [jb, je) out of shared memory.  Lanes of a warp that straddles two x-rows, periodic wrap
// segments, ghost segments and windows larger than the buffer take the per-lane path of k_build_list_f32.  Same
// traversal order, same tests: the list is identical to k_build_list_f32's.
constexpr int kWcWindow = 96;
template <int DIM>
__global__ void __launch_bounds__(kForceBlock)
k_build_list_wc(int n, Grid g, const uint32_t *__restrict__ start, float rl2f, const float4 *__restrict__ posf,
                uint32_t *__restrict__ nl, int64_t stride, int kmax, int32_t *__restrict__ nnbr, uint32_t *__restrict__ ovf, DevCtl *ctl,
                double *__restrict__ xref)
{
    __shared__ float4 win[kForceBlock / 32][kWcWindow];
    const StatePtrs st = ctl->st[ctl->cur];
    if (g.slab) g.gpos_m = ctl->gpos_m;
    if (n < 0) n = ctl->n_own;
    const int i = blockIdx.x * kForceBlock + threadIdx.x;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const bool active = i < n;
    int cnt = 0;
    double4 pd = make_double4(0, 0, 0, 1);
    int cx = 0, cy = 0, cz = 0;
    if (active) {
        pd = st.pos[i];
        if (xref) {
            const double pk[3] = {pd.x, pd.y, pd.z};
#pragma unroll
            for (int k = 0; k < DIM; k++) xref[k * st.cap + i] = pk[k] + g.L[k] * (double)st.img[k * st.cap + i];
        }
        cell_of_point<DIM, 0>(g, pd, cx, cy, cz);
    }
    const float xi = (float)pd.x, yi = (float)pd.y, zi = (float)pd.z;
    const float Lx = (float)g.L[0], Ly = (float)g.L[1], Lz = (float)g.L[2];
    uint32_t *slot = nl + i;
    const int nx = g.nxo, ny = g.nc[1], nz = g.nc[2];
    const int lx = cx - g.c0;
    // the cooperative path needs every active lane in the same x-row (same cy, cz)
    const unsigned act = __ballot_sync(0xffffffffu, active);
    const int leader = act ? (__ffs(act) - 1) : 0;
    const int cy0 = __shfl_sync(0xffffffffu, cy, leader), cz0 = __shfl_sync(0xffffffffu, cz, leader);
    const bool same_row = __all_sync(0xffffffffu, !active || (cy == cy0 && cz == cz0));
    auto test = [&](float xj, float yj, float zj, float sx, float sy, float sz, uint32_t j) {
        const float dx = (xi - xj) - sx, dy_ = (yi - yj) - sy;
        float d2 = dx * dx + dy_ * dy_;
        if (DIM == 3) {
            const float dz_ = (zi - zj) - sz;
            d2 += dz_ * dz_;
        }
        const bool hit = d2 <= rl2f && (int)j != i;
        if (hit && cnt < kmax) *slot = j;
        slot += hit ? stride : 0;
        cnt += hit ? 1 : 0;
    };
    for (int dz = (DIM == 3 ? -1 : 0); dz <= (DIM == 3 ? 1 : 0); dz++) {
        int oz = cz + dz, kz = 0;
        if (DIM == 3) {
            if (oz < 0) { oz += nz; kz = -1; }
            else if (oz >= nz) { oz -= nz; kz = 1; }
        }
        for (int dy = -1; dy <= 1; dy++) {
            int oy = cy + dy, ky = 0;
            if (oy < 0) { oy += ny; ky = -1; }
            else if (oy >= ny) { oy -= ny; ky = 1; }
            const uint32_t rowg = (uint32_t)oz * ny + oy;
            const uint32_t row = rowg * nx;
            uint32_t seg_b[2] = {0, 0}, seg_e[2] = {0, 0};
            int seg_k[2] = {0, 0}, nseg = 0;
            bool seg_ghost[2] = {false, false};
            if (active) {
                nseg = 1;
                if (!g.slab) {
                    if (lx == 0) {
                        seg_b[0] = start[row]; seg_e[0] = start[row + 2]; seg_k[0] = 0;
                        seg_b[1] = start[row + nx - 1]; seg_e[1] = start[row + nx]; seg_k[1] = -1;
                        nseg = 2;
                    } else if (lx == nx - 1) {
                        seg_b[0] = start[row + nx - 2]; seg_e[0] = start[row + nx]; seg_k[0] = 0;
                        seg_b[1] = start[row]; seg_e[1] = start[row + 1]; seg_k[1] = 1;
                        nseg = 2;
                    } else {
                        seg_b[0] = start[row + lx - 1]; seg_e[0] = start[row + lx + 2]; seg_k[0] = 0;
                    }
                } else {
                    const int lo = lx > 0 ? lx - 1 : 0, hi = lx < nx - 1 ? lx + 1 : nx - 1;
                    seg_b[0] = start[row + lo]; seg_e[0] = start[row + hi + 1]; seg_k[0] = 0;
                    if (lx == 0) {
                        seg_b[1] = g.gstart_l[rowg]; seg_e[1] = g.gstart_l[rowg + 1]; seg_k[1] = g.kx_left; seg_ghost[1] = true;
                        nseg = 2;
                    } else if (lx == nx - 1) {
                        seg_b[1] = g.gstart_r[rowg]; seg_e[1] = g.gstart_r[rowg + 1]; seg_k[1] = g.kx_right; seg_ghost[1] = true;
                        nseg = 2;
                    }
                }
            }
            const float sy = (float)ky * Ly, sz = (float)kz * Lz;
            // main segment (always unshifted in x): the warp's union window, staged once
            uint32_t wb = active ? seg_b[0] : 0xffffffffu, we = active ? seg_e[0] : 0u;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                wb = min(wb, __shfl_xor_sync(0xffffffffu, wb, o));
                we = max(we, __shfl_xor_sync(0xffffffffu, we, o));
            }
            const bool coop = same_row && we > wb && (we - wb) <= (uint32_t)kWcWindow;  // warp-uniform
            if (coop) {
                for (uint32_t q = wb + lane; q < we; q += 32) win[wid][q - wb] = __ldg(&posf[q]);
                __syncwarp();
                for (uint32_t j = seg_b[0]; j < seg_e[0]; j++) {
                    const float4 pj = win[wid][j - wb];
                    test(pj.x, pj.y, pj.z, 0.0f, sy, sz, j);
                }
                __syncwarp();
            } else {
                for (uint32_t j = seg_b[0]; j < seg_e[0]; j++) {
                    const float4 pj = __ldg(&posf[j]);
                    test(pj.x, pj.y, pj.z, 0.0f, sy, sz, j);
                }
            }
            // second segment: periodic wrap piece or ghost column, per lane
            if (nseg == 2) {
                const float sx = (float)seg_k[1] * Lx;
                for (uint32_t j = seg_b[1]; j < seg_e[1]; j++) {
                    float xj, yj, zj;
                    if (!seg_ghost[1]) {
                        const float4 pj = __ldg(&posf[j]);
                        xj = pj.x; yj = pj.y; zj = pj.z;
                    } else {
                        const double4 pj = ldg_pos(&g.gpos_m[j]);
                        xj = (float)pj.x; yj = (float)pj.y; zj = (float)pj.z;
                    }
                    test(xj, yj, zj, sx, sy, sz, j);
                }
            }
        }
    }
    if (active) {
        nnbr[i] = cnt;
        if (cnt > kmax) ovf[atomicAdd(&ctl->n_overflow, 1)] = (uint32_t)i;
    }
    int m = cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0 && m > 0) atomicMax(&ctl->max_nnbr, m);
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl->list_valid = 1;
        ctl->disp = 0.0;
        ctl->dref2_bits = 0ull;
    }
}

// minimum image of a listed pair: positions are re-wrapped every step, so the image is decided per evaluation,
// dx = (xi - xj) - k*L with k = +-1 when |xi - xj| > L/2 (the oracle's nearbyint gives the same k for listed pairs)
template <int DIM, int TRI = -1>
__device__ __forceinline__ double separation_wrap(const Grid &g, const double4 &pi, const double4 &pj, double &dx, double &dy, double &dz)
{
    const bool tri = (TRI < 0) ? (g.tri != 0) : (TRI != 0);
    if (tri) {  // general cell: nearest image through the fractional coordinates
        dx = pi.x - pj.x;
        dy = pi.y - pj.y;
        dz = (DIM == 3) ? pi.z - pj.z : 0.0;
        min_image<DIM, 1>(g, dx, dy, dz);
        double q2 = fma(dy, dy, dx * dx);
        if (DIM == 3) q2 = fma(dz, dz, q2);
        return q2;
    }
    dx = pi.x - pj.x;
    if (dx > g.hL[0]) dx -= g.L[0];
    else if (dx < -g.hL[0]) dx += g.L[0];
    dy = pi.y - pj.y;
    if (dy > g.hL[1]) dy -= g.L[1];
    else if (dy < -g.hL[1]) dy += g.L[1];
    double d2 = fma(dy, dy, dx * dx);
    dz = 0.0;
    if (DIM == 3) {
        dz = pi.z - pj.z;
        if (dz > g.hL[2]) dz -= g.L[2];
        else if (dz < -g.hL[2]) dz += g.L[2];
        d2 = fma(dz, dz, d2);
    }
    return d2;
}
template <int DIM>
__device__ __forceinline__ double separation_plain(const double4 &pi, const double4 &pj, double &dx, double &dy, double &dz)
{
    dx = pi.x - pj.x;
    dy = pi.y - pj.y;
    double d2 = fma(dy, dy, dx * dx);
    dz = 0.0;
    if (DIM == 3) {
        dz = pi.z - pj.z;
        d2 = fma(dz, dz, d2);
    }
    return d2;
}

#ifndef MDB_UNROLL
#define MDB_UNROLL 2  // tools/tune_force.py on B200: the inner list holds ~2 candidates; 2 -> no spills at 72 registers
                      // (4 spilled 176 B per thread and the spills missed L1: ncu saw more local than global L2 sectors)
#endif
#ifndef MDB_FORCE_MIN_CTAS
#define MDB_FORCE_MIN_CTAS 7  // 72 registers, 28 warps/SM
#endif
constexpr int kUnroll = MDB_UNROLL;  // independent neighbour gathers in flight per thread
#ifndef MDB_PARK
#define MDB_PARK 1    // 0: evaluate every hit where the list walk finds it (no queue, no second gather)
#endif


// Two-level Verlet list + software pipelining.
//  * outer list (radius r_search + skin): built from the cells, ~8 candidates per particle for PseudoHS at phi = 0.47;
//  * inner list (radius r_search + skin_in, skin_in << skin): the candidates of the outer list that can come within
//    r_search before the inner displacement budget is used up, ~2 per particle.  When ctl->inner_refresh is set this
//    kernel walks the outer list, rewrites the inner list and evaluates forces in the same pass; otherwise it walks
//    only the inner list.  Both are supersets of the pairs within r_search, so the pair set is identical.
//  * the chain of dependent memory round trips per tile (own record -> indices -> neighbour records -> velocity) bounds
//    the kernel (ncu: long-scoreboard stalls, nothing saturated): the next tile's operands and this tile's velocities
//    are requested at the top of the tile, and index chunk c+1 is requested with the gathers of chunk c.
struct ListView {
    const uint32_t *nl;     // outer list, column-major [kmax][stride]
    const int32_t *nnbr;
    uint32_t *nl_in;        // inner list, column-major [kmax_in][stride]
    int32_t *nnbr_in;
    int64_t stride;
    int kmax, kmax_in;
    double rin2;            // (r_search + skin_in)^2
};

// SLAB: neighbour indices >= g.g0 address the ghost buffer of an x-slab (single domain: no such indices, no select)
// KICK2: 0 forces only, 1 + second half kick, 2 + second half kick and the next step's kick-drift-wrap (leap_epilogue)
template <int DIM, class Pot, int KICK2, bool SLAB, bool TRI = false>
__global__ void __launch_bounds__(kForceBlock, MDB_FORCE_MIN_CTAS)
k_force_list(int n, DevCtl *__restrict__ ctl, Grid g, ListView lv, double cutoff2, double rwrap, Pot pot, PotParams pp, double dt,
             ForceOut out, int guard)
{
    // guard: launched speculatively behind the rebuild decision (slab steps); a pending rebuild turns the launch into a no-op
    if (guard && ctl->need_rebuild) return;
    if (SLAB) g.gpos_m = ctl->gpos_m;
    constexpr bool kPark = Pot::kSparseHits && (MDB_PARK != 0);  // park hits in a queue, or evaluate them where they are found
    __shared__ uint32_t queue[kPark ? kQueue : 1][kForceBlock];
    const StatePtrs s = ctl->st[ctl->cur];
    const double4 *__restrict__ pos = s.pos;
    double4 *__restrict__ pos_next = ctl->st[ctl->cur ^ 1].pos;  // KICK2 == 2 only
    double vmax2 = 0.0, dref2 = 0.0;
    const unsigned long long rng_step = ctl->rng_step;  // KICK2 == 3 only
    const bool refresh = ctl->inner_refresh != 0;  // uniform over the grid
    const uint32_t *__restrict__ nl = refresh ? lv.nl : lv.nl_in;
    const int32_t *__restrict__ nnbr = refresh ? lv.nnbr : lv.nnbr_in;
    const int kcap = refresh ? lv.kmax : lv.kmax_in;
    const int64_t stride = lv.stride;
    ThreadSums acc;
    if (n < 0) n = ctl->n_own;
    const int ntiles = (n + kForceBlock - 1) / kForceBlock;
    int max_in = 0;
    // this thread's column of the parked-hit queue as a shared-window address, computed once: ncu's source view showed seven
    // address instructions (S2R tid, S2UR cta id, ULEA, LEA, IMAD ...) in front of every queue access, ~7 % of all instructions
    const uint32_t qaddr = smem_u32(&queue[0][threadIdx.x]);
    constexpr uint32_t kQRow = kForceBlock * (uint32_t)sizeof(uint32_t);
    // prologue: first tile's operands
    // grid-strided tiles: CTAs that run at the same time work on adjacent tiles, so one SM's gathers are another's L2
    // hits (a contiguous range of tiles per CTA measured 24% slower)
    int tile = blockIdx.x;
    const int tile_end = ntiles;
    const int tile_step = gridDim.x;
    int i = tile * kForceBlock + threadIdx.x;
    double4 pi_n = make_double4(0, 0, 0, 1);
    int cnt_n = 0;
    uint32_t jn[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; u++) jn[u] = 0;
    if (tile < tile_end && i < n) {
        pi_n = pos[i];
        cnt_n = nnbr[i];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) jn[u] = nl[(int64_t)u * stride + i];  // rows < kUnroll always exist
    }
    for (; tile < tile_end; tile += tile_step) {
        i = tile * kForceBlock + threadIdx.x;
        bool active = i < n;
        const double4 pi = pi_n;
        int cnt = cnt_n;
        uint32_t jj[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) jj[u] = jn[u];
        // this tile's velocities (needed only by the epilogue) and the next tile's operands: in flight during the gathers
        double vel[3] = {0.0, 0.0, 0.0};
        if ((KICK2 == 1 || KICK2 == 2) && active) {
#pragma unroll
            for (int k = 0; k < DIM; k++) vel[k] = s.vel[k * s.cap + i];
        }
        {
            const int tn = tile + tile_step;
            const int in = tn * kForceBlock + threadIdx.x;
            cnt_n = 0;
            if (tn < tile_end && in < n) {
                pi_n = pos[in];
                cnt_n = nnbr[in];
#pragma unroll
                for (int u = 0; u < kUnroll; u++) jn[u] = nl[(int64_t)u * stride + in];
            }
        }
        // which list this particle walks: normally the grid-uniform choice; a particle whose inner list overflowed
        // keeps walking its outer list (exact either way)
        const uint32_t *__restrict__ mynl = nl;
        bool write_inner = refresh;
        if (active && !refresh && cnt > kcap) {
            mynl = lv.nl;
            cnt = lv.nnbr[i];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) jj[u] = mynl[(int64_t)u * stride + i];
        }
        if (active && cnt > lv.kmax) {  // outer list overflow: handled in full by k_force_overflow
            active = false;
        }
        if (!active) cnt = 0;
        double F[3] = {0.0, 0.0, 0.0};
        int nq = 0, nin = 0;
        // a listed neighbour can sit across a periodic face only if this particle is within rwrap of one
        bool wrap = pi.x < rwrap || pi.x > g.L[0] - rwrap || pi.y < rwrap || pi.y > g.L[1] - rwrap;
        if (DIM == 3) wrap = wrap || pi.z < rwrap || pi.z > g.L[2] - rwrap;
        if (TRI) wrap = true;  // general cells: every listed pair goes through the fractional nearest image
        const bool wrap_any = __any_sync(0xffffffffu, wrap);  // warp-uniform: the wrapped separation is exact for every pair
        auto drain_one = [&]() {
            if (nq > 0) {
                uint32_t j;
                --nq;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(j) : "r"(qaddr + (uint32_t)nq * kQRow) : "memory");
                double4 pj = ldg_pos(SLAB ? nbr_ptr(g, pos, j) : pos + j);
                double dx, dy, dz, d2;
                if (wrap_any) d2 = separation_wrap<DIM, TRI ? 1 : 0>(g, pi, pj, dx, dy, dz);
                else d2 = separation_plain<DIM>(pi, pj, dx, dy, dz);
                pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, acc.e, acc.w, acc.np);
            }
        };
        // row k0 + kUnroll of this particle's list column: a running pointer instead of a 64-bit multiply per index load
        const uint32_t *__restrict__ prow = mynl + (int64_t)kUnroll * stride + i;
        for (int k0 = 0; k0 < cnt; k0 += kUnroll, prow += (int64_t)kUnroll * stride) {
            double4 pj[kUnroll];
            uint32_t jc[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                // slots past the end of this particle's list issue no load at all: the kernel is bound by L1 wavefronts
                // (ncu: 74% of the LSU data pipe), and a dummy gather costs a sector per lane like a real one
                jc[u] = jj[u];
                pj[u] = make_double4(0.0, 0.0, 0.0, 1.0);
                if (k0 + u < cnt) pj[u] = ldg_pos(SLAB ? nbr_ptr(g, pos, jc[u]) : pos + jc[u]);
            }
            // next index chunk travels while the gathers above are in flight
            if (k0 + kUnroll < cnt) {
#pragma unroll
                for (int u = 0; u < kUnroll; u++)
                    jj[u] = (k0 + kUnroll + u < cnt) ? prow[(int64_t)u * stride] : (uint32_t)i;
            }
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                double dx, dy, dz;
                double d2 = wrap ? separation_wrap<DIM, TRI ? 1 : 0>(g, pi, pj[u], dx, dy, dz) : separation_plain<DIM>(pi, pj[u], dx, dy, dz);
                const bool valid = k0 + u < cnt;
                if (write_inner && valid && d2 <= lv.rin2) {
                    if (nin < lv.kmax_in) lv.nl_in[(int64_t)nin * stride + i] = jc[u];
                    nin++;
                }
                if (valid && d2 <= cutoff2 && pot.may_interact(pp, d2, pi.w, pj[u].w)) {
                    if (kPark) {
                        asm volatile("st.shared.u32 [%0], %1;" ::"r"(qaddr + (uint32_t)nq * kQRow), "r"(jc[u]) : "memory");
                        nq++;
                    } else
                        pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj[u].w, F, acc.e, acc.w, acc.np);
                }
            }
            if (kPark && nq > kQueue - kUnroll) {
                while (nq > 0) drain_one();
            }
        }
        if (kPark) {
            while (__any_sync(0xffffffffu, nq > 0)) drain_one();
        }
        if (write_inner && i < n) {
            lv.nnbr_in[i] = active ? nin : 0x7fffffff;  // outer-overflow particles never use the inner list
            max_in = max(max_in, active ? nin : 0);
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < DIM; k++) s.frc[k * s.cap + i] = F[k];
            if (KICK2 == 1) {
                double v2 = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    double v = vel[k];
                    v += (F[k] * dt) * 0.5;
                    s.vel[k * s.cap + i] = v;
                    v2 = (k == 0) ? v * v : v2 + v * v;
                }
                acc.v2 += v2;
            }
            if (KICK2 == 2) leap_epilogue<DIM, TRI ? 1 : 0>(i, F, vel, pi, s, pos_next, g, dt, acc.v2, vmax2);
            if (KICK2 == 3) brown_epilogue<DIM, TRI ? 1 : 0>(i, F, pi, s, pos_next, g, dt, ctl, rng_step, vmax2, dref2);
        }
    }
    if (KICK2 == 2) leap_report<kForceBlock>(vmax2, dt, ctl);
    if (KICK2 == 3) brown_report<kForceBlock>(vmax2, dref2, ctl);
    if (refresh) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) max_in = max(max_in, __shfl_xor_sync(0xffffffffu, max_in, o));
        if ((threadIdx.x & 31) == 0 && max_in > 0) atomicMax(&ctl->max_nnbr_in, max_in);
    }
    cta_epilogue(acc, out, blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
// K4 (TMA operands): k_force_list with the tile's OWN operands -- its 128 particle records, list counts and the first
// index rows, all contiguous in slot order -- brought into shared memory by TMA bulk copies (cp.async.bulk + mbarrier,
// SASS UBLKCP) two tiles ahead, through a three-stage ring filled by one elected thread.
//   ncu on k_force_list (profiles/r02_force_kernel_ncu.md): besides the gathers themselves (29 % of all warp-stall
//   samples), 9 % sit on STL instructions -- the compiler spills the next tile's prefetched operands, and a spill of a
//   value that is still in flight waits for the load -- and 5 % on the first uses of those operands at the top of a
//   tile.  Here the next tiles' operands never occupy registers: the copy engine writes them to shared memory, a thread
//   reads its own element when the tile starts (one mbarrier wait per tile, normally long satisfied) and hands the
//   stage back at once.  The neighbour gathers stay direct LDG.256s (staging scattered records cost more LSU wavefronts
//   than it saved, k_force_list_staged below; staging whole neighbour rows costs as much as the whole kernel,
//   tools/stage_probe.cu).  Arithmetic, candidate order and reduction order are those of k_force_list: bit-identical.
// ------------------------------------------------------------------------------------------------
constexpr int kTmaStages = 3;
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tMDBW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra MDBD_%=;\n\tbra MDBW_%=;\n\tMDBD_%=:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gsrc, uint32_t bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"((unsigned long long)__cvta_generic_to_global(gsrc)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ double4 lds_rec(const double4 *p)
{
    double4 r;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p);
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%4];\n\tld.shared.v2.f64 {%2,%3}, [%4+16];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "r"(a) : "memory");
    return r;
}

template <int DIM, class Pot, int KICK2, bool SLAB>
__global__ void __launch_bounds__(kForceBlock, MDB_FORCE_MIN_CTAS)
k_force_list_tma(int n, DevCtl *__restrict__ ctl, Grid g, ListView lv, double cutoff2, double rwrap, Pot pot, PotParams pp, double dt,
             ForceOut out, int guard)
{
    // guard: launched speculatively behind the rebuild decision (slab steps); a pending rebuild turns the launch into a no-op
    if (guard && ctl->need_rebuild) return;
    if (SLAB) g.gpos_m = ctl->gpos_m;
    constexpr bool kPark = Pot::kSparseHits && (MDB_PARK != 0);  // park hits in a queue, or evaluate them where they are found
    constexpr bool TRI = false;  // (general cells keep k_force_list)
    constexpr int NS = kTmaStages;
    __shared__ __align__(128) double4 s_pos[NS][kForceBlock];
    __shared__ __align__(128) int32_t s_cnt[NS][kForceBlock];
    __shared__ __align__(128) uint32_t s_idx[NS][kUnroll][kForceBlock];
    __shared__ __align__(8) unsigned long long bar_full[NS], bar_empty[NS];
    __shared__ uint32_t queue[kPark ? kQueue : 1][kForceBlock];
    const StatePtrs s = ctl->st[ctl->cur];
    const double4 *__restrict__ pos = s.pos;
    double4 *__restrict__ pos_next = ctl->st[ctl->cur ^ 1].pos;  // KICK2 == 2 only
    double vmax2 = 0.0, dref2 = 0.0;
    const unsigned long long rng_step = ctl->rng_step;  // KICK2 == 3 only
    const bool refresh = ctl->inner_refresh != 0;  // uniform over the grid
    const uint32_t *__restrict__ nl = refresh ? lv.nl : lv.nl_in;
    const int32_t *__restrict__ nnbr = refresh ? lv.nnbr : lv.nnbr_in;
    const int kcap = refresh ? lv.kmax : lv.kmax_in;
    const int64_t stride = lv.stride;
    ThreadSums acc;
    if (n < 0) n = ctl->n_own;
    const int ntiles = (n + kForceBlock - 1) / kForceBlock;
    int max_in = 0;
    // operand ring: the elected thread posts the bulk copies of tile k+2 while tile k is evaluated
    const int tid = threadIdx.x;
    if (tid == 0) {
#pragma unroll
        for (int q = 0; q < NS; q++) {
            mbar_init(&bar_full[q], 1);
            mbar_init(&bar_empty[q], kForceBlock / 32);
        }
        mbar_fence_init();
    }
    __syncthreads();
    const int tile_end = ntiles;
    const int tile_step = gridDim.x;
    const int64_t cap = s.cap;
    auto post = [&](int t, int st) {  // elected thread: own records, list counts and the first kUnroll index rows of tile t
        const int base = t * kForceBlock;
        const uint32_t lim = (uint32_t)min((int64_t)kForceBlock, cap - base);  // a multiple of 32 (cap is)
        mbar_expect_tx(&bar_full[st], lim * (uint32_t)(sizeof(double4) + 4 + 4 * kUnroll));
        bulk_g2s(&s_pos[st][0], pos + base, lim * (uint32_t)sizeof(double4), &bar_full[st]);
        bulk_g2s(&s_cnt[st][0], nnbr + base, lim * 4u, &bar_full[st]);
#pragma unroll
        for (int u = 0; u < kUnroll; u++) bulk_g2s(&s_idx[st][u][0], nl + (int64_t)u * stride + base, lim * 4u, &bar_full[st]);
    };
    int tile = blockIdx.x;
    int i = tile * kForceBlock + tid;
    if (tid == 0) {
        if (tile < tile_end) post(tile, 0);
        if (tile + tile_step < tile_end) post(tile + tile_step, 1);
    }
    int kt = 0;  // ordinal of this CTA's tile
    for (; tile < tile_end; tile += tile_step, kt++) {
        i = tile * kForceBlock + tid;
        bool active = i < n;
        const int st = kt % NS;
        if (tid == 0 && tile + 2 * tile_step < tile_end) {
            const int st2 = (kt + 2) % NS;
            if (kt + 2 >= NS) mbar_wait(&bar_empty[st2], (uint32_t)(((kt + 2) / NS - 1) & 1));  // every warp has let go of that stage
            post(tile + 2 * tile_step, st2);
        }
        mbar_wait(&bar_full[st], (uint32_t)((kt / NS) & 1));
        double4 pi = make_double4(0, 0, 0, 1);
        int cnt = 0;
        uint32_t jj[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; u++) jj[u] = 0;
        if (active) {
            pi = lds_rec(&s_pos[st][tid]);
            cnt = s_cnt[st][tid];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) jj[u] = s_idx[st][u][tid];
        }
        // the operands are in registers: hand the stage back (one arrival per warp)
        __syncwarp();
        if ((tid & 31) == 0) mbar_arrive(&bar_empty[st]);
        // this tile's velocities (needed only by the epilogue) and the next tile's operands: in flight during the gathers
        double vel[3] = {0.0, 0.0, 0.0};
        if ((KICK2 == 1 || KICK2 == 2) && active) {
#pragma unroll
            for (int k = 0; k < DIM; k++) vel[k] = s.vel[k * s.cap + i];
        }
        // which list this particle walks: normally the grid-uniform choice; a particle whose inner list overflowed
        // keeps walking its outer list (exact either way)
        const uint32_t *__restrict__ mynl = nl;
        bool write_inner = refresh;
        if (active && !refresh && cnt > kcap) {
            mynl = lv.nl;
            cnt = lv.nnbr[i];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) jj[u] = mynl[(int64_t)u * stride + i];
        }
        if (active && cnt > lv.kmax) {  // outer list overflow: handled in full by k_force_overflow
            active = false;
        }
        if (!active) cnt = 0;
        double F[3] = {0.0, 0.0, 0.0};
        int nq = 0, nin = 0;
        // a listed neighbour can sit across a periodic face only if this particle is within rwrap of one
        bool wrap = pi.x < rwrap || pi.x > g.L[0] - rwrap || pi.y < rwrap || pi.y > g.L[1] - rwrap;
        if (DIM == 3) wrap = wrap || pi.z < rwrap || pi.z > g.L[2] - rwrap;
        if (TRI) wrap = true;  // general cells: every listed pair goes through the fractional nearest image
        const bool wrap_any = __any_sync(0xffffffffu, wrap);  // warp-uniform: the wrapped separation is exact for every pair
        auto drain_one = [&]() {
            if (nq > 0) {
                int j = (int)queue[--nq][tid];
                double4 pj = ldg_pos(SLAB ? nbr_ptr(g, pos, (uint32_t)j) : pos + j);
                double dx, dy, dz, d2;
                if (wrap_any) d2 = separation_wrap<DIM, TRI ? 1 : 0>(g, pi, pj, dx, dy, dz);
                else d2 = separation_plain<DIM>(pi, pj, dx, dy, dz);
                pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, acc.e, acc.w, acc.np);
            }
        };
        for (int k0 = 0; k0 < cnt; k0 += kUnroll) {
            double4 pj[kUnroll];
            uint32_t jc[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                // slots past the end of this particle's list issue no load at all: the kernel is bound by L1 wavefronts
                // (ncu: 74% of the LSU data pipe), and a dummy gather costs a sector per lane like a real one
                jc[u] = jj[u];
                pj[u] = make_double4(0.0, 0.0, 0.0, 1.0);
                if (k0 + u < cnt) pj[u] = ldg_pos(SLAB ? nbr_ptr(g, pos, jc[u]) : pos + jc[u]);
            }
            // next index chunk travels while the gathers above are in flight
            if (k0 + kUnroll < cnt) {
#pragma unroll
                for (int u = 0; u < kUnroll; u++)
                    jj[u] = (k0 + kUnroll + u < cnt) ? mynl[(int64_t)(k0 + kUnroll + u) * stride + i] : (uint32_t)i;
            }
#pragma unroll
            for (int u = 0; u < kUnroll; u++) {
                double dx, dy, dz;
                double d2 = wrap ? separation_wrap<DIM, TRI ? 1 : 0>(g, pi, pj[u], dx, dy, dz) : separation_plain<DIM>(pi, pj[u], dx, dy, dz);
                const bool valid = k0 + u < cnt;
                if (write_inner && valid && d2 <= lv.rin2) {
                    if (nin < lv.kmax_in) lv.nl_in[(int64_t)nin * stride + i] = jc[u];
                    nin++;
                }
                if (valid && d2 <= cutoff2 && pot.may_interact(pp, d2, pi.w, pj[u].w)) {
                    if (kPark) queue[nq++][tid] = jc[u];
                    else pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj[u].w, F, acc.e, acc.w, acc.np);
                }
            }
            if (kPark && nq > kQueue - kUnroll) {
                while (nq > 0) drain_one();
            }
        }
        if (kPark) {
            while (__any_sync(0xffffffffu, nq > 0)) drain_one();
        }
        if (write_inner && i < n) {
            lv.nnbr_in[i] = active ? nin : 0x7fffffff;  // outer-overflow particles never use the inner list
            max_in = max(max_in, active ? nin : 0);
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < DIM; k++) s.frc[k * s.cap + i] = F[k];
            if (KICK2 == 1) {
                double v2 = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    double v = vel[k];
                    v += (F[k] * dt) * 0.5;
                    s.vel[k * s.cap + i] = v;
                    v2 = (k == 0) ? v * v : v2 + v * v;
                }
                acc.v2 += v2;
            }
            if (KICK2 == 2) leap_epilogue<DIM, TRI ? 1 : 0>(i, F, vel, pi, s, pos_next, g, dt, acc.v2, vmax2);
            if (KICK2 == 3) brown_epilogue<DIM, TRI ? 1 : 0>(i, F, pi, s, pos_next, g, dt, ctl, rng_step, vmax2, dref2);
        }
    }
    if (KICK2 == 2) leap_report<kForceBlock>(vmax2, dt, ctl);
    if (KICK2 == 3) brown_report<kForceBlock>(vmax2, dref2, ctl);
    if (refresh) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) max_in = max(max_in, __shfl_xor_sync(0xffffffffu, max_in, o));
        if ((tid & 31) == 0 && max_in > 0) atomicMax(&ctl->max_nnbr_in, max_in);
    }
    cta_epilogue(acc, out, blockIdx.x);
}


// ------------------------------------------------------------------------------------------------
// K4 (staged): the list-driven pair-force kernel with the NEIGHBOUR RECORDS STAGED IN SHARED MEMORY by asynchronous
// copies (cp.async / LDGSTS) one tile ahead of their use.
//   ncu on k_force_list (profiles/r01_s9_fused_ncu.md): 37 % of all warp-stall samples sit on the first use of a gathered
//   neighbour record (long scoreboard), with 7 warps per scheduler nothing else is saturated.  Here every thread posts the
//   gathers of tile t+1 (list indices prefetched one tile earlier still, into registers) as 16-byte asynchronous copies
//   into its own shared-memory slots, then evaluates tile t out of the slots filled during tile t-1: the round trip to
//   L2/HBM overlaps a whole tile of arithmetic instead of stalling the warp, and the parked-hit queue re-reads its records
//   from shared memory instead of gathering them a second time.
//   * slots: stage[2][kStageSlots][128] double4, private to the thread that filled them: no CTA barrier in the tile loop,
//     cp.async.wait_group is all the synchronisation there is;
//   * a particle with more than kStageSlots listed candidates (about 5 % on the inner list; every particle while the outer
//     list is walked to refresh the inner one) takes the rest through direct gathers as before;
//   * arithmetic, candidate order and reduction order are those of k_force_list: results are bit-identical.
// ------------------------------------------------------------------------------------------------
#ifndef MDB_STAGE_SLOTS
#define MDB_STAGE_SLOTS 4
#endif
#ifndef MDB_STAGED_MIN_CTAS
#define MDB_STAGED_MIN_CTAS 6
#endif
constexpr int kStageSlots = MDB_STAGE_SLOTS;

__device__ __forceinline__ void cp_async_rec(double4 *smem_dst, const double4 *gsrc)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const unsigned long long s = (unsigned long long)__cvta_generic_to_global(gsrc);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n\tcp.async.cg.shared.global [%0+16], [%1+16], 16;" ::"r"(d), "l"(s) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int DIM, class Pot, int KICK2, bool SLAB>
__global__ void __launch_bounds__(kForceBlock, MDB_STAGED_MIN_CTAS)
k_force_list_staged(int n, DevCtl *__restrict__ ctl, Grid g, ListView lv, double cutoff2, double rwrap, Pot pot, PotParams pp, double dt,
                    ForceOut out, int guard)
{
    if (guard && ctl->need_rebuild) return;
    if (SLAB) g.gpos_m = ctl->gpos_m;
    constexpr int KS = kStageSlots;
    constexpr bool kPark = Pot::kSparseHits && (MDB_PARK != 0);
    __shared__ __align__(32) double4 stage[2][KS][kForceBlock];
    __shared__ uint32_t queue[kPark ? kQueue : 1][kForceBlock];
    const StatePtrs s = ctl->st[ctl->cur];
    const double4 *__restrict__ pos = s.pos;
    double4 *__restrict__ pos_next = ctl->st[ctl->cur ^ 1].pos;  // KICK2 >= 2 only
    double vmax2 = 0.0, dref2 = 0.0;
    const unsigned long long rng_step = ctl->rng_step;  // KICK2 == 3 only
    const bool refresh = ctl->inner_refresh != 0;       // uniform over the grid
    const uint32_t *__restrict__ nl = refresh ? lv.nl : lv.nl_in;
    const int32_t *__restrict__ nnbr = refresh ? lv.nnbr : lv.nnbr_in;
    const int kcap = refresh ? lv.kmax : lv.kmax_in;
    const int64_t stride = lv.stride;
    ThreadSums acc;
    if (n < 0) n = ctl->n_own;
    const int ntiles = (n + kForceBlock - 1) / kForceBlock;
    const int tile_step = gridDim.x;
    const int tid = threadIdx.x;
    int max_in = 0;

    // list head of one particle: which list it walks (a particle whose inner list overflowed keeps walking its outer list,
    // a particle whose outer list overflowed is left to k_force_overflow), how many candidates, the first KS indices
    auto fetch = [&](int tile, int &cnt, bool &outer, uint32_t (&idx)[KS]) {
        const int i = tile * kForceBlock + tid;
        cnt = 0;
        outer = refresh;
        if (tile < ntiles && i < n) {
            int c = nnbr[i];
            const uint32_t *__restrict__ src = nl;
            if (!refresh && c > kcap) {
                src = lv.nl;
                c = lv.nnbr[i];
                outer = true;
            }
            if (c > lv.kmax) c = -1;  // outer overflow: not ours
#pragma unroll
            for (int u = 0; u < KS; u++) idx[u] = (u < c) ? src[(int64_t)u * stride + i] : 0u;
            cnt = c;
        }
    };
    auto post = [&](int buf, int cnt, const uint32_t (&idx)[KS]) {
#pragma unroll
        for (int u = 0; u < KS; u++)
            if (u < cnt) cp_async_rec(&stage[buf][u][tid], SLAB ? nbr_ptr(g, pos, idx[u]) : pos + idx[u]);
        cp_async_commit();
    };

    int tile = blockIdx.x;
    int cnt, cnt_s;
    bool outer, outer_s;
    uint32_t idx_s[KS];
    double4 pi_n = make_double4(0, 0, 0, 1);
    // prologue: stage the first tile, fetch the list heads of the second
    fetch(tile, cnt, outer, idx_s);
    post(0, cnt, idx_s);
    {
        const int i0 = tile * kForceBlock + tid;
        if (tile < ntiles && i0 < n) pi_n = pos[i0];
    }
    fetch(tile + tile_step, cnt_s, outer_s, idx_s);
    int buf = 0;
    for (; tile < ntiles; tile += tile_step, buf ^= 1) {
        const int i = tile * kForceBlock + tid;
        const bool in_range = i < n;
        const double4 pi = pi_n;
        // the next tile's gathers go out now; its own record and the list heads of the tile after it follow
        post(buf ^ 1, cnt_s, idx_s);
        const int cnt_next = cnt_s;
        const bool outer_next = outer_s;
        {
            const int in = (tile + tile_step) * kForceBlock + tid;
            if (tile + tile_step < ntiles && in < n) pi_n = pos[in];
        }
        fetch(tile + 2 * tile_step, cnt_s, outer_s, idx_s);
        double vel[3] = {0.0, 0.0, 0.0};
        if ((KICK2 == 1 || KICK2 == 2) && in_range) {
#pragma unroll
            for (int k = 0; k < DIM; k++) vel[k] = s.vel[k * s.cap + i];
        }
        const bool active = in_range && cnt >= 0;
        const int ncand = active ? cnt : 0;
        const uint32_t *__restrict__ mynl = outer ? lv.nl : nl;
        const bool write_inner = refresh;
        double F[3] = {0.0, 0.0, 0.0};
        int nq = 0, nin = 0;
        bool wrap = pi.x < rwrap || pi.x > g.L[0] - rwrap || pi.y < rwrap || pi.y > g.L[1] - rwrap;
        if (DIM == 3) wrap = wrap || pi.z < rwrap || pi.z > g.L[2] - rwrap;
        const bool wrap_any = __any_sync(0xffffffffu, wrap);
        cp_async_wait<1>();  // this tile's slots are filled (the group posted above may still be in flight)
        auto candidate_record = [&](int k) -> double4 {
            if (k < KS) return lds_rec(&stage[buf][k][tid]);
            const uint32_t j = mynl[(int64_t)k * stride + i];
            return ldg_pos(SLAB ? nbr_ptr(g, pos, j) : pos + j);
        };
        auto drain_one = [&]() {
            if (nq > 0) {
                const int k = (int)queue[--nq][tid];
                const double4 pj = candidate_record(k);
                double dx, dy, dz, d2;
                if (wrap_any) d2 = separation_wrap<DIM, 0>(g, pi, pj, dx, dy, dz);
                else d2 = separation_plain<DIM>(pi, pj, dx, dy, dz);
                pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, acc.e, acc.w, acc.np);
            }
        };
        auto consider = [&](int k, const double4 &pj, uint32_t j) {
            double dx, dy, dz;
            const double d2 = wrap ? separation_wrap<DIM, 0>(g, pi, pj, dx, dy, dz) : separation_plain<DIM>(pi, pj, dx, dy, dz);
            if (write_inner && d2 <= lv.rin2) {
                if (nin < lv.kmax_in) lv.nl_in[(int64_t)nin * stride + i] = j;
                nin++;
            }
            if (d2 <= cutoff2 && pot.may_interact(pp, d2, pi.w, pj.w)) {
                if (kPark) queue[nq++][tid] = (uint32_t)k;
                else pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, acc.e, acc.w, acc.np);
            }
        };
        // staged candidates: shared memory
#pragma unroll
        for (int u = 0; u < KS; u++) {
            if (u < ncand) {
                const double4 pj = lds_rec(&stage[buf][u][tid]);
                // the index is only needed when the inner list is being rewritten: re-read it then (coalesced, L1/L2-resident)
                const uint32_t j = write_inner ? mynl[(int64_t)u * stride + i] : 0u;
                consider(u, pj, j);
            }
        }
        if (kPark && nq > kQueue - kUnroll) {
            while (nq > 0) drain_one();
        }
        // the rest of a long list: direct gathers, index chunk c+1 requested with the gathers of chunk c
        if (ncand > KS) {
            uint32_t jj[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; u++) jj[u] = (KS + u < ncand) ? mynl[(int64_t)(KS + u) * stride + i] : (uint32_t)i;
            for (int k0 = KS; k0 < ncand; k0 += kUnroll) {
                double4 pj[kUnroll];
                uint32_t jc[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; u++) {
                    jc[u] = jj[u];
                    pj[u] = make_double4(0.0, 0.0, 0.0, 1.0);
                    if (k0 + u < ncand) pj[u] = ldg_pos(SLAB ? nbr_ptr(g, pos, jc[u]) : pos + jc[u]);
                }
                if (k0 + kUnroll < ncand) {
#pragma unroll
                    for (int u = 0; u < kUnroll; u++)
                        jj[u] = (k0 + kUnroll + u < ncand) ? mynl[(int64_t)(k0 + kUnroll + u) * stride + i] : (uint32_t)i;
                }
#pragma unroll
                for (int u = 0; u < kUnroll; u++)
                    if (k0 + u < ncand) consider(k0 + u, pj[u], jc[u]);
                if (kPark && nq > kQueue - kUnroll) {
                    while (nq > 0) drain_one();
                }
            }
        }
        if (kPark) {
            while (__any_sync(0xffffffffu, nq > 0)) drain_one();
        }
        if (write_inner && in_range) {
            lv.nnbr_in[i] = active ? nin : 0x7fffffff;  // outer-overflow particles never use the inner list
            max_in = max(max_in, active ? nin : 0);
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < DIM; k++) s.frc[k * s.cap + i] = F[k];
            if (KICK2 == 1) {
                double v2 = 0.0;
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    double v = vel[k];
                    v += (F[k] * dt) * 0.5;
                    s.vel[k * s.cap + i] = v;
                    v2 = (k == 0) ? v * v : v2 + v * v;
                }
                acc.v2 += v2;
            }
            if (KICK2 == 2) leap_epilogue<DIM, 0>(i, F, vel, pi, s, pos_next, g, dt, acc.v2, vmax2);
            if (KICK2 == 3) brown_epilogue<DIM, 0>(i, F, pi, s, pos_next, g, dt, ctl, rng_step, vmax2, dref2);
        }
        cnt = cnt_next;
        outer = outer_next;
    }
    cp_async_wait<0>();
    if (KICK2 == 2) leap_report<kForceBlock>(vmax2, dt, ctl);
    if (KICK2 == 3) brown_report<kForceBlock>(vmax2, dref2, ctl);
    if (refresh) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) max_in = max(max_in, __shfl_xor_sync(0xffffffffu, max_in, o));
        if ((tid & 31) == 0 && max_in > 0) atomicMax(&ctl->max_nnbr_in, max_in);
    }
    cta_epilogue(acc, out, blockIdx.x);
}

// list overflow (more than kmax neighbours within r_list): exact fallback through the stale-but-conservative
// build-time cells.  Slot order is the build-time cell order, so the home cell is found by bisection on `start`.
// One warp-lane per overflowing particle; normally the overflow list is empty and this kernel exits at once.
template <int DIM, class Pot, int KICK2>
__global__ void __launch_bounds__(kForceBlock)
k_force_overflow(DevCtl *__restrict__ ctl, Grid g, const uint32_t *__restrict__ start, const uint32_t *__restrict__ ovf,
                 double cutoff2, Pot pot, PotParams pp, double dt, ForceOut out, int slot0, int guard)
{
    if (guard && ctl->need_rebuild) return;
    if (g.slab) g.gpos_m = ctl->gpos_m;
    const StatePtrs s = ctl->st[ctl->cur];
    const double4 *__restrict__ pos = s.pos;
    const int novf = ctl->n_overflow;
    double4 *__restrict__ pos_next = ctl->st[ctl->cur ^ 1].pos;  // KICK2 >= 2 only
    double vmax2 = 0.0, dref2 = 0.0;
    const unsigned long long rng_step = ctl->rng_step;
    ThreadSums acc;
    for (int q = blockIdx.x * kForceBlock + threadIdx.x; q < novf; q += gridDim.x * kForceBlock) {
        const int i = (int)ovf[q];
        double F[3] = {0.0, 0.0, 0.0};
        double4 pi = pos[i];
        int64_t ncell = (int64_t)g.nxo * g.nc[1] * g.nc[2];
        int64_t lo = 0, hi = ncell;  // start[lo] <= i < start[hi]
        while (hi - lo > 1) {
            int64_t mid = (lo + hi) >> 1;
            if (start[mid] <= (uint32_t)i) lo = mid;
            else hi = mid;
        }
        int cx = (int)(lo % g.nxo) + g.c0, cy = (int)((lo / g.nxo) % g.nc[1]), cz = (int)(lo / ((int64_t)g.nxo * g.nc[1]));
        traverse_cells<DIM>(g, start, pos, i, pi, cx, cy, cz, 1e300, [&](int j, double, double, double, double, double, int) {
            double4 pj = ldg_pos(nbr_ptr(g, pos, (uint32_t)j));
            double dx, dy, dz;
            double d2 = separation_wrap<DIM>(g, pi, pj, dx, dy, dz);
            if (d2 <= cutoff2 && pot.may_interact(pp, d2, pi.w, pj.w))
                pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, acc.e, acc.w, acc.np);
        });
        if (KICK2 == 2) {
            double vel[3] = {0.0, 0.0, 0.0};
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                s.frc[k * s.cap + i] = F[k];
                vel[k] = s.vel[k * s.cap + i];
            }
            leap_epilogue<DIM>(i, F, vel, pi, s, pos_next, g, dt, acc.v2, vmax2);
        } else if (KICK2 == 3) {
#pragma unroll
            for (int k = 0; k < DIM; k++) s.frc[k * s.cap + i] = F[k];
            brown_epilogue<DIM>(i, F, pi, s, pos_next, g, dt, ctl, rng_step, vmax2, dref2);
        } else
            particle_epilogue<DIM, KICK2 != 0>(i, F, s, dt, acc);
    }
    if (KICK2 == 2 && novf > 0) leap_report<kForceBlock>(vmax2, dt, ctl);  // novf is grid-uniform
    if (KICK2 == 3 && novf > 0) brown_report<kForceBlock>(vmax2, dref2, ctl);
    cta_epilogue(acc, out, slot0 + blockIdx.x);
}

// ------------------------------------------------------------------------------------------------
// K4c  all-pairs kernel for boxes with fewer than three cells per direction (tiny systems) -- same arithmetic,
// k = nearbyint(dx/L) exactly as the oracle's brute force.
// ------------------------------------------------------------------------------------------------
template <int DIM, class Pot, bool KICK2>
__global__ void __launch_bounds__(kForceBlock)
k_force_brute(int n, const DevCtl *__restrict__ ctl, Grid g, double cutoff2, Pot pot, PotParams pp, double dt, ForceOut out)
{
    const StatePtrs s = ctl->st[ctl->cur];
    ThreadSums acc;
    if (n < 0) n = ctl->n_own;
    const int ntiles = (n + kForceBlock - 1) / kForceBlock;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int i = tile * kForceBlock + threadIdx.x;
        if (i < n) {
            double F[3] = {0.0, 0.0, 0.0};
            double4 pi = s.pos[i];
            double inv[3] = {g.invL[0], g.invL[1], g.invL[2]};
            for (int j = 0; j < n; j++) {
                if (j == i) continue;
                double4 pj = ldg_pos(&s.pos[j]);
                double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = (DIM == 3) ? pi.z - pj.z : 0.0;
                if (!g.tri) {
                    dx = dx - nearbyint(dx * inv[0]) * g.L[0];
                    dy = dy - nearbyint(dy * inv[1]) * g.L[1];
                    if (DIM == 3) dz = dz - nearbyint(dz * inv[2]) * g.L[2];
                } else
                    min_image<DIM>(g, dx, dy, dz);
                double d2 = fma(dy, dy, dx * dx);
                if (DIM == 3) d2 = fma(dz, dz, d2);
                if (d2 <= cutoff2 && pot.may_interact(pp, d2, pi.w, pj.w))
                    pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, acc.e, acc.w, acc.np);
            }
            particle_epilogue<DIM, KICK2>(i, F, s, dt, acc);
        }
    }
    cta_epilogue(acc, out, blockIdx.x);
}


// debug pair counter: pairs with d2 <= cutoff^2 (what map_pairwise! visits); per-particle counts in original order
template <int DIM>
__global__ void __launch_bounds__(kForceBlock)
k_count_pairs(int n, const DevCtl *__restrict__ ctl, Grid g,
              const uint32_t *__restrict__ start, double cutoff2, int use_cells, int32_t *__restrict__ per_particle,
              unsigned long long *__restrict__ total)
{
    const double4 *__restrict__ pos = ctl->st[ctl->cur].pos;
    const int32_t *__restrict__ id = ctl->st[ctl->cur].id;
    int i = blockIdx.x * kForceBlock + threadIdx.x;
    int cnt = 0;
    if (i < n) {
        double4 pi = pos[i];
        if (use_cells) {
            int cx, cy, cz;
            cell_of_point<DIM>(g, pi, cx, cy, cz);
            traverse_cells<DIM>(g, start, pos, i, pi, cx, cy, cz, cutoff2,
                                [&](int, double, double, double, double, double, int) { cnt++; });
        } else {
            double inv[3] = {g.invL[0], g.invL[1], g.invL[2]};
            for (int j = 0; j < n; j++) {
                if (j == i) continue;
                double4 pj = ldg_pos(&pos[j]);
                double dx = pi.x - pj.x, dy = pi.y - pj.y, dz = (DIM == 3) ? pi.z - pj.z : 0.0;
                if (!g.tri) {
                    dx = dx - nearbyint(dx * inv[0]) * g.L[0];
                    dy = dy - nearbyint(dy * inv[1]) * g.L[1];
                    if (DIM == 3) dz = dz - nearbyint(dz * inv[2]) * g.L[2];
                } else
                    min_image<DIM>(g, dx, dy, dz);
                double d2 = fma(dy, dy, dx * dx);
                if (DIM == 3) d2 = fma(dz, dz, d2);
                if (d2 <= cutoff2) cnt++;
            }
        }
        if (per_particle) per_particle[id[i]] = cnt;
    }
    unsigned long long c = (unsigned long long)cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((threadIdx.x & 31) == 0 && c) atomicAdd(total, c);
}

// ------------------------------------------------------------------------------------------------
// K5  first half of velocity Verlet, fused with the pending Bussi rescale and the periodic wrap:
//   v <- v*alpha (bussi!, src/thermostat.jl:45-47, deferred from the previous step)
//   v += (f*dt)/2 ; x += v*dt ; x = wrap_to_box(x)   (src/integrate.jl:8-21, src/boundary.jl:7-17)
// Grid-stride CTAs; each folds its largest |v|^2 into ctl->dmax2_bits with one atomicMax (order-independent, so
// deterministic): that bounds this step's displacement for the Verlet-skin bookkeeping.
// ------------------------------------------------------------------------------------------------
template <int DIM>
__global__ void __launch_bounds__(kStreamBlock)
k_kick_drift(int n, Grid g, double dt, DevCtl *__restrict__ ctl)
{
    const StatePtrs s = ctl->st[ctl->cur];
    const double alpha = ctl->alpha;
    if (n < 0) n = ctl->n_own;
    double vmax2 = 0.0;
    for (int i = blockIdx.x * kStreamBlock + threadIdx.x; i < n; i += gridDim.x * kStreamBlock) {
        double4 p = ld_pos(&s.pos[i]);
        double x[3] = {p.x, p.y, p.z};
        double v2 = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            double v = s.vel[k * s.cap + i];
            double f = s.frc[k * s.cap + i];
            v = v * alpha;
            v += (f * dt) * 0.5;  // == f*dt/2.0 bit-for-bit
            s.vel[k * s.cap + i] = v;
            v2 = (k == 0) ? v * v : v2 + v * v;
            x[k] = x[k] + v * dt;
        }
        double ncr[3];
        wrap_point<DIM>(g, x, ncr);
#pragma unroll
        for (int k = 0; k < DIM; k++)
            if (ncr[k] != 0.0) s.img[k * s.cap + i] += (int32_t)ncr[k];
        st_pos(&s.pos[i], make_double4(x[0], x[1], x[2], p.w));
        vmax2 = fmax(vmax2, v2);
    }
    double r[1] = {vmax2 * (dt * dt)};  // squared displacement of this step: every mover reports the same quantity
    block_reduce<1, kStreamBlock, true>(r);
    // non-negative doubles order like their bit patterns; NaN (0x7ff8...) compares above everything and forces a rebuild
    if (threadIdx.x == 0) atomicMax(&ctl->dmax2_bits, (unsigned long long)__double_as_longlong(r[0]));
}

// apply a pending Bussi scale to the resident velocities (end of an NVT run) : src/thermostat.jl:45-47
template <int DIM>
__global__ void k_scale(int n, DevCtl *ctl)
{
    const StatePtrs s = ctl->st[ctl->cur];
    const double alpha = ctl->alpha;
    if (n < 0) n = ctl->n_own;
    for (int i = blockIdx.x * kStreamBlock + threadIdx.x; i < n; i += gridDim.x * kStreamBlock) {
#pragma unroll
        for (int k = 0; k < DIM; k++) s.vel[k * s.cap + i] = s.vel[k * s.cap + i] * alpha;
    }
}
__global__ void k_reset_alpha(DevCtl *ctl) { ctl->alpha = 1.0; }
__global__ void k_set_brownian(DevCtl *ctl, double ktemp, double sigma, unsigned long long seed, const double *xref)
{
    ctl->bd_ktemp = ktemp;
    ctl->bd_sigma = sigma;
    ctl->bd_seed = seed;
    ctl->bd_xref = xref;
}
// after the gather-reorder: the other state buffer becomes live
__global__ void k_flip(DevCtl *ctl, const uint32_t *__restrict__ n_new)
{
    if (n_new) ctl->n_own = (int)*n_new;
    ctl->mig_count[0] = ctl->mig_count[1] = 0;
    ctl->cur ^= 1;
    ctl->max_nnbr = 0;
    ctl->max_nnbr_in = 0;
    ctl->n_overflow = 0;
    ctl->rebuilds += 1;
}

// K8  Brownian step: x = x + (f*dt/kT) + (noise*sigma), wrap (src/integrate.jl:66-82 intended semantics, SURVEY Q5;
// sigma = sqrt(2 dt), src/simulation.jl:212).  Noise keyed by (original particle id, RNG step).
template <int DIM>
__global__ void __launch_bounds__(kStreamBlock)
k_brownian(int n, Grid g, double dt, double ktemp, double sigma, uint64_t seed, DevCtl *__restrict__ ctl,
           const double *__restrict__ xref, int guard)
{
    if (guard && ctl->need_rebuild) return;
    const StatePtrs s = ctl->st[ctl->cur];
    const unsigned long long rng_step = ctl->rng_step;
    if (n < 0) n = ctl->n_own;
    double dmax2 = 0.0, dref2 = 0.0;
    for (int i = blockIdx.x * kStreamBlock + threadIdx.x; i < n; i += gridDim.x * kStreamBlock) {
        double noise[3];
        brownian_noise<DIM>(seed, rng_step, (uint32_t)s.id[i], noise);
        double4 p = ld_pos(&s.pos[i]);
        double x[3] = {p.x, p.y, p.z};
        double d2 = 0.0, r2 = 0.0;
        double ncrv[3];
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            double f = s.frc[k * s.cap + i];
            double xv = x[k] + (f * dt / ktemp) + (noise[k] * sigma);
            double del = xv - x[k];
            d2 = (k == 0) ? del * del : d2 + del * del;
            x[k] = xv;
        }
        wrap_point<DIM>(g, x, ncrv);
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            const double ncr = ncrv[k];
            if (xref) {  // (diagonal cells only: the engine passes no xref for a general cell)
                // a random walk moves far less than the sum of its per-step maxima: measure the true displacement from
                // the positions the Verlet list was built for (unwrapped through the image counters)
                int32_t im = s.img[k * s.cap + i];
                if (ncr != 0.0) {
                    im += (int32_t)ncr;
                    s.img[k * s.cap + i] = im;
                }
                double dr = (x[k] + g.L[k] * (double)im) - xref[k * s.cap + i];
                r2 = (k == 0) ? dr * dr : r2 + dr * dr;
            } else if (ncr != 0.0) {
                s.img[k * s.cap + i] += (int32_t)ncr;
            }
        }
        st_pos(&s.pos[i], make_double4(x[0], x[1], x[2], p.w));
        dmax2 = fmax(dmax2, d2);
        dref2 = fmax(dref2, r2);
    }
    double r[2] = {dmax2, dref2};
    block_reduce<2, kStreamBlock, true>(r);
    if (threadIdx.x == 0) {
        atomicMax(&ctl->dmax2_bits, (unsigned long long)__double_as_longlong(r[0]));
        if (xref) atomicMax(&ctl->dref2_bits, (unsigned long long)__double_as_longlong(r[1]));
    }
}

#ifndef __CUDACC_RTC__
// ------------------------------------------------------------------------------------------------
// K-skin: consume the displacement bound of the move that just happened, decide whether the Verlet list must be
// rebuilt before the coming force evaluation, and drive the conditional graph node.
// dmax2 holds the largest squared displacement of the move that just happened (scale = 1).
// ------------------------------------------------------------------------------------------------
struct CondHandles {  // a rebuild split into several conditional graph nodes (slab protocol) shares one decision
    cudaGraphConditionalHandle h[3];
    int n;
};
__global__ void k_skin_check(double scale, double skin, double skin_in, int always, int exact, DevCtl *ctl,
                             cudaGraphConditionalHandle handle, int use_handle, CondHandles extra = CondHandles{{0, 0, 0}, 0})
{
    double m = __longlong_as_double((long long)ctl->dmax2_bits);
    ctl->dmax2_bits = 0ull;  // consumed
    double step = sqrt(m) * scale;
    // exact: the mover measured the true largest displacement since the list build (Brownian); it replaces the running sum
    // of per-step maxima and remains a valid starting point if later steps go back to accumulating bounds
    const unsigned long long dref = ctl->dref2_bits;
    double disp = (exact && dref != 0ull) ? sqrt(__longlong_as_double((long long)dref)) : ctl->disp + step;
    if (!exact) ctl->dref2_bits = 0ull;  // a mover that does not measure it leaves the exact displacement unknown
    // NaN-safe: a non-finite bound forces a rebuild
    int need = always || !ctl->list_valid || !(2.0 * disp <= skin);
    ctl->disp = need ? 0.0 : disp;
    ctl->need_rebuild = need;
    // two-level Verlet list: the tight inner list (r_search + skin_in) is re-derived from the outer one by the force
    // kernel itself whenever its own, smaller, displacement budget is used up
    double disp_in = ctl->disp_in + step;
    int need_in = need || !(2.0 * disp_in <= skin_in);
    ctl->disp_in = need_in ? 0.0 : disp_in;
    ctl->inner_refresh = need_in;
    if (use_handle) cudaGraphSetConditional(handle, need ? 1u : 0u);
    for (int q = 0; q < extra.n; q++) cudaGraphSetConditional(extra.h[q], need ? 1u : 0u);
}

#endif  // !__CUDACC_RTC__

// ------------------------------------------------------------------------------------------------
// K9  thermo: fixed-order second stage of the per-CTA partials (deterministic), kinetic energy / temperature
// (src/thermostat.jl:50-67), and the Bussi-Donadio-Parrinello scale for NVT (src/thermostat.jl:20-48) drawn from the
// counter-based RNG.  Feeds the scalars read at src/simulation.jl:118-131.
// ------------------------------------------------------------------------------------------------
__global__ void k_finalize(int nslots, const double *__restrict__ part, int ensemble, double nf, double dt, double tau,
                           const double *__restrict__ ktemp, uint64_t seed, double *__restrict__ thermo, int advance, DevCtl *ctl,
                           int stage, int guard, int swap_pos = 0)
{
    if (guard && ctl->need_rebuild) return;
    // fused NVE step: the force kernels wrote the moved positions into the other buffer; make it the live one
    if (swap_pos && threadIdx.x == 0) {
        double4 *t = ctl->st[0].pos;
        ctl->st[0].pos = ctl->st[1].pos;
        ctl->st[1].pos = t;
    }
    // stage 0: single domain.  Slabs: stage 1 leaves this rank's sums in ctl->red for the all-reduce,
    // stage 2 continues from the globally summed ctl->red (identical on every rank).
    double r[4] = {0.0, 0.0, 0.0, 0.0};
    if (stage != 2) {
        for (int q = threadIdx.x; q < nslots; q += blockDim.x) {
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] += part[c * kMaxPartials + q];
        }
        block_reduce<4, kStreamBlock>(r);
    }
    if (threadIdx.x == 0 && stage == 1) {
#pragma unroll
        for (int c = 0; c < 4; c++) ctl->red[c] = r[c];
        return;
    }
    if (threadIdx.x == 0) {
        if (stage == 2) {
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] = ctl->red[c];
        }
        double U = 0.5 * r[0], W = 0.5 * r[1], NP = 0.5 * r[2];
        double KE = r[3] / 2.0;
        unsigned long long st = ctl->step;
        if (ensemble == 1) {
            ThermoRng rng;
            rng.init(seed, ctl->rng_step);
            double r1 = rng.normal();
            double r2 = rng.sum_noises(nf - 1.0);
            double scale = bussi_scale(KE, ktemp[st], nf, dt, tau, r1, r2);
            ctl->alpha = scale;
            KE = (scale * scale) * KE;
        }
        if (ensemble == 2) KE = 0.0;
        ctl->last[0] = U;
        ctl->last[1] = W;
        ctl->last[2] = KE;
        ctl->last[3] = NP;
        if (!(isfinite(U) && isfinite(KE))) ctl->nonfinite = 1;
        if (thermo) {
            thermo[4 * st + 0] = U;
            thermo[4 * st + 1] = W;
            thermo[4 * st + 2] = KE;
            thermo[4 * st + 3] = NP;
        }
        if (advance) {
            ctl->step = st + 1;
            ctl->rng_step += 1;
        }
    }
}


// ------------------------------------------------------------------------------------------------
// FIRE minimiser (src/minimize.jl:31-135), second caller of the force path.  ctl->fire = {dt, alpha, steps_since_neg,
// converged, mix_a, mix_b, zero_v, steps_done}.  The engine's velocity array holds the FIRE velocities for the
// duration of the call (it travels with the particle re-sorts); the caller's velocities are saved and restored.
// ------------------------------------------------------------------------------------------------
// v += dt*f ; partial sums of P = v.f, |v|^2, |f|^2   (src/minimize.jl:89-96)
template <int DIM>
__global__ void __launch_bounds__(kStreamBlock)
k_fire_kick(int n, const DevCtl *__restrict__ ctl, double *__restrict__ part)
{
    const StatePtrs s = ctl->st[ctl->cur];
    const double dt = ctl->fire[0];
    double P = 0.0, vv = 0.0, ff = 0.0;
    for (int i = blockIdx.x * kStreamBlock + threadIdx.x; i < n; i += gridDim.x * kStreamBlock) {
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            double f = s.frc[k * s.cap + i];
            double v = s.vel[k * s.cap + i];
            v += dt * f;
            s.vel[k * s.cap + i] = v;
            P += v * f;
            vv += v * v;
            ff += f * f;
        }
    }
    double r[3] = {P, vv, ff};
    block_reduce<3, kStreamBlock>(r);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int q = 0; q < 3; q++) part[q * kMaxPartials + blockIdx.x] = r[q];
    }
}

// scalar logic of one FIRE iteration: convergence test, velocity mixing factors, time-step adaptation (:76-115)
__global__ void k_fire_decide(int nslots, const double *__restrict__ part, double ndof, double tol, double dt_initial, double dt_max,
                              double alpha0, double f_inc, double f_dec, int n_min, DevCtl *ctl)
{
    double r[3] = {0.0, 0.0, 0.0};
    for (int q = threadIdx.x; q < nslots; q += blockDim.x) {
#pragma unroll
        for (int c = 0; c < 3; c++) r[c] += part[c * kMaxPartials + q];
    }
    block_reduce<3, kStreamBlock>(r);
    if (threadIdx.x == 0) {
        double *F = ctl->fire;
        if (F[3] != 0.0) return;  // already converged: the remaining iterations of the chunk are no-ops
        double P = r[0], v_norm = sqrt(r[1]), f_norm = sqrt(r[2]);
        F[7] += 1.0;
        ctl->last[2] = f_norm / sqrt(ndof);
        if (f_norm / sqrt(ndof) < tol) {
            F[3] = 1.0;
            return;
        }
        double alpha = F[1], dt = F[0];
        if (v_norm > 0 && f_norm > 0) {
            F[4] = 1.0 - alpha;
            F[5] = alpha * (v_norm / f_norm);
        } else {
            F[4] = 1.0;
            F[5] = 0.0;
        }
        if (P > 0) {
            F[2] += 1.0;
            if (F[2] > (double)n_min) {
                dt = fmin(dt * f_inc, dt_max);
                alpha *= 0.99;
            }
            F[6] = 0.0;
        } else {
            dt = fmax(dt * f_dec, dt_initial);
            alpha = alpha0;
            F[2] = 0.0;
            F[6] = 1.0;
        }
        F[0] = dt;
        F[1] = alpha;
    }
}

// v = (1-alpha) v + scale f (or 0) ; x += dt v ; wrap   (:98-123)
template <int DIM>
__global__ void __launch_bounds__(kStreamBlock)
k_fire_move(int n, Grid g, DevCtl *__restrict__ ctl)
{
    const StatePtrs s = ctl->st[ctl->cur];
    const double *F = ctl->fire;
    if (F[3] != 0.0) return;
    const double dt = F[0], ma = F[4], mb = F[5];
    const bool zero_v = F[6] != 0.0, mix = !(ma == 1.0 && mb == 0.0);
    double dmax2 = 0.0;
    for (int i = blockIdx.x * kStreamBlock + threadIdx.x; i < n; i += gridDim.x * kStreamBlock) {
        double4 p = ld_pos(&s.pos[i]);
        double x[3] = {p.x, p.y, p.z};
        double d2 = 0.0;
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            double v = s.vel[k * s.cap + i];
            if (mix) v = ma * v + mb * s.frc[k * s.cap + i];
            if (zero_v) v = 0.0;
            s.vel[k * s.cap + i] = v;
            double step = dt * v;
            d2 = (k == 0) ? step * step : d2 + step * step;
            x[k] = x[k] + step;
        }
        double ncr[3];
        wrap_point<DIM>(g, x, ncr);
#pragma unroll
        for (int k = 0; k < DIM; k++)
            if (ncr[k] != 0.0) s.img[k * s.cap + i] += (int32_t)ncr[k];
        st_pos(&s.pos[i], make_double4(x[0], x[1], x[2], p.w));
        dmax2 = fmax(dmax2, d2);
    }
    double r[1] = {dmax2};
    block_reduce<1, kStreamBlock, true>(r);
    if (threadIdx.x == 0) atomicMax(&ctl->dmax2_bits, (unsigned long long)__double_as_longlong(r[0]));
}

template <int DIM>
__global__ void k_zero_vel(int n, const DevCtl *__restrict__ ctl)
{
    const StatePtrs s = ctl->st[ctl->cur];
    for (int i = blockIdx.x * kStreamBlock + threadIdx.x; i < n; i += gridDim.x * kStreamBlock) {
#pragma unroll
        for (int k = 0; k < DIM; k++) s.vel[k * s.cap + i] = 0.0;
    }
}

__global__ void k_bussi_hooks(int what, double a0, double a1, double a2, double a3, double a4, double a5, double a6, uint64_t seed,
                              uint64_t step, double *out)
{
    if (what == 0) {
        out[0] = bussi_scale(a0, a1, a2, a3, a4, a5, a6);
    } else {
        ThermoRng rng;
        rng.init(seed, step);
        out[0] = rng.normal();
        out[1] = rng.sum_noises(a0 - 1.0);
    }
}

}  // namespace mdb
