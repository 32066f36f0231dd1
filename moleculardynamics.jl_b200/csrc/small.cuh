// small.cuh -- K0-small: the whole step loop of src/simulation.jl:88-108 / :231-250 as ONE persistent kernel for systems
// of a few thousand particles.  Two versions: k_small_run (cooperative grid, below; N up to 4096, fallback) and k_small_cluster
// (one thread-block cluster, further down; default for N <= 2048).  k_small_run: a cooperative
// kernel for systems of a few thousand particles (BASELINE configs 1 and 2: N = 1024 / 1200), where a step is a few
// microseconds of work and per-kernel launch latency would dominate ("K0-small").
//   * one thread per particle, 64-thread CTAs spread over as many SMs as there are CTAs (N = 1024 -> 16 SMs); the
//     particle's position, velocity, force and image counters live in registers for the whole call;
//   * two grid-wide barriers per step (cooperative launch): after the drift (new positions visible) and after the
//     forces (thermo scalars / Bussi scale); Brownian dynamics needs one;
//   * every CTA stages all positions in shared memory once per step (coalesced L2 reads), so the neighbour gathers
//     and the all-pairs list rebuild never leave the SM;
//   * Verlet list in global memory (column-major), membership decided in FP32 with a safety margin (a superset is all
//     that is needed); every force-loop decision in FP64 with the same arithmetic as the large-system kernels;
//   * reductions: per-CTA partials in global memory, folded in fixed CTA order by every CTA after the barrier
//     (deterministic, no atomics); thread 0 of CTA 0 writes the thermo row; each CTA's thread 0 draws the same
//     Philox numbers for the Bussi scale.
#pragma once
#include <cooperative_groups.h>

#include "kernels.cuh"

namespace mdb {

constexpr int kSmallBlock = 64;
constexpr int kSmallMaxN = 4096;
constexpr int kSmallMaxGrid = kSmallMaxN / kSmallBlock;

struct SmallArgs {
    int n, ensemble;         // 0 NVE, 1 NVT, 2 Brownian
    long long nsteps;
    double dt, tau, ktemp, nf;
    const double *ktemp_per_step;
    double *thermo;          // [nsteps][4] or null
    uint32_t *nl;            // [kmax][n]
    int kmax;
    double skin, cutoff2;
    float rlist2_f;          // (r_list^2 + margin) for the FP32 membership test
    unsigned long long seed;
    double *gpart;           // [3][kSmallMaxGrid][5] per-CTA partials: phase 0 = {vmax2}, phase 1 = {e, w, np, ke2, dmax2},
                             // phase 1 double-buffered by step parity (see the step loop)
    double *rng_pre;         // cluster version, NVT: [nsteps][2] the thermostat's Gaussian and chi-square draws of every step
};

template <int DIM, class Pot>
__global__ void __launch_bounds__(kSmallBlock)
k_small_run(DevCtl *__restrict__ ctl, Grid g, SmallArgs a, Pot pot, PotParams pp)
{
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ double4 spos[];
    __shared__ double red[5][kSmallBlock / 32];
    __shared__ double s_alpha;
    const StatePtrs s = ctl->st[ctl->cur];
    const int n = a.n, tid = threadIdx.x, G = gridDim.x;
    const int i = blockIdx.x * kSmallBlock + tid;
    const bool active = i < n;
    // Brownian steps have a single grid barrier (B): a CTA that is one step ahead must not overwrite the phase-1 partials a
    // slower CTA is still folding, so they alternate between two buffers by step parity (two steps ahead is impossible:
    // barrier B of the step in between needs every CTA)
    double *part0 = a.gpart, *part1_base = a.gpart + kSmallMaxGrid * 5;
    // Brownian dynamics has a single barrier per step, so the move must not overwrite positions other CTAs may still be
    // staging: it ping-pongs between the two state buffers (velocity Verlet writes before barrier A and needs no such care)
    double4 *pbuf[2] = {s.pos, ctl->st[ctl->cur ^ 1].pos};
    int par = 0;

    // the particle's state stays in registers for the whole call
    double x[3] = {0, 0, 0}, v[3] = {0, 0, 0}, f[3] = {0, 0, 0}, sig = 1.0;
    int32_t im[3] = {0, 0, 0};
    uint32_t pid = 0;
    if (active) {
        double4 p = s.pos[i];
        x[0] = p.x; x[1] = p.y; x[2] = p.z; sig = p.w;
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            v[k] = s.vel[k * s.cap + i];
            f[k] = s.frc[k * s.cap + i];
            im[k] = s.img[k * s.cap + i];
        }
        pid = (uint32_t)s.id[i];
    }
    int cnt = 0;
    double alpha = 1.0, disp = 0.0;
    bool rebuild = true;          // the list of a previous call is never trusted
    double dmax2_prev = 0.0;      // Brownian: displacement bound of the previous step's move (grid-wide)
    unsigned long long rng_step = ctl->rng_step;
    const double sigma_bd = sqrt(2.0 * a.dt);

    auto block_fold = [&](double *vals, int nv, bool is_max, double *dst) {
        // per-CTA reduction (2 warps), result written by thread 0 to dst[0..nv)
        const int lane = tid & 31, w = tid >> 5;
        for (int q = 0; q < nv; q++) {
            double r = is_max ? warp_max(vals[q]) : warp_sum(vals[q]);
            if (lane == 0) red[q][w] = r;
        }
        __syncthreads();
        if (tid == 0) {
            for (int q = 0; q < nv; q++) {
                double r = red[q][0];
                for (int k = 1; k < kSmallBlock / 32; k++) r = is_max ? fmax(r, red[q][k]) : r + red[q][k];
                dst[q] = r;
            }
        }
        __syncthreads();
    };

    for (long long step = 0; step < a.nsteps; step++) {
        double *part1 = part1_base + (step & 1) * (kSmallMaxGrid * 5);
        double dmax2;
        if (a.ensemble != 2) {
            // ---- first half kick + drift + wrap (src/integrate.jl:8-21, src/boundary.jl:7-17); Bussi scale of the previous step
            double v2 = 0.0;
            if (active) {
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    double vk = v[k] * alpha;
                    vk += (f[k] * a.dt) * 0.5;
                    v[k] = vk;
                    v2 = (k == 0) ? vk * vk : v2 + vk * vk;
                    double xv = x[k] + vk * a.dt;
                    double frac = g.invL[k] * xv;
                    double ncr = floor(frac);
                    if (ncr != 0.0) im[k] += (int32_t)ncr;
                    x[k] = g.L[k] * (frac - ncr);
                }
                s.pos[i] = make_double4(x[0], x[1], x[2], sig);
            }
            double vals[1] = {v2};
            block_fold(vals, 1, true, part0 + blockIdx.x * 5);
            grid.sync();  // barrier A: new positions and the per-CTA maxima are visible everywhere
            dmax2 = 0.0;
            for (int b = 0; b < G; b++) dmax2 = fmax(dmax2, __ldcg(part0 + b * 5));
            dmax2 = dmax2 * (a.dt * a.dt);
        } else {
            dmax2 = dmax2_prev;
        }
        {
            double d = disp + sqrt(dmax2);
            rebuild = rebuild || !(2.0 * d <= a.skin);
            disp = rebuild ? 0.0 : d;
        }
        // ---- all positions into shared memory (coalesced, L2)
        __syncthreads();
        for (int j = tid; j < n; j += kSmallBlock) {
            const double2 *src = reinterpret_cast<const double2 *>(pbuf[par] + j);
            double2 lo = __ldcg(src), hi = __ldcg(src + 1);
            spos[j] = make_double4(lo.x, lo.y, hi.x, hi.y);
        }
        __syncthreads();
        const double4 pi = make_double4(x[0], x[1], x[2], sig);
        // ---- Verlet list rebuild: all pairs from shared memory, FP32 membership with margin (superset)
        if (rebuild) {
            if (active) {
                const float xi = (float)pi.x, yi = (float)pi.y, zi = (float)pi.z;
                const float Lx = (float)g.L[0], Ly = (float)g.L[1], Lz = (float)g.L[2];
                const float hx = 0.5f * Lx, hy = 0.5f * Ly, hz = 0.5f * Lz;
                cnt = 0;
                for (int j = 0; j < n; j++) {
                    const double4 pj = spos[j];
                    float dx = xi - (float)pj.x, dy = yi - (float)pj.y;
                    dx = dx > hx ? dx - Lx : (dx < -hx ? dx + Lx : dx);
                    dy = dy > hy ? dy - Ly : (dy < -hy ? dy + Ly : dy);
                    float d2 = dx * dx + dy * dy;
                    if (DIM == 3) {
                        float dz = zi - (float)pj.z;
                        dz = dz > hz ? dz - Lz : (dz < -hz ? dz + Lz : dz);
                        d2 += dz * dz;
                    }
                    if (d2 <= a.rlist2_f && j != i) {
                        if (cnt < a.kmax) a.nl[(size_t)cnt * n + i] = (uint32_t)j;
                        cnt++;
                    }
                }
            }
            if (i == 0) ctl->rebuilds += 1;
            rebuild = false;
        }
        // ---- pair forces (src/pairwise.jl:26-39)
        double e = 0.0, w = 0.0, np = 0.0, ke2 = 0.0, bd2 = 0.0;
        if (active) {
            double F[3] = {0.0, 0.0, 0.0};
            auto candidate = [&](int j) {
                const double4 pj = spos[j];
                double dx, dy, dz;
                double d2 = separation_wrap<DIM>(g, pi, pj, dx, dy, dz);
                if (d2 <= a.cutoff2 && pot.may_interact(pp, d2, pi.w, pj.w)) pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, e, w, np);
            };
            if (cnt <= a.kmax) {
                for (int k = 0; k < cnt; k++) candidate((int)a.nl[(size_t)k * n + i]);
            } else {
                for (int j = 0; j < n; j++)
                    if (j != i) candidate(j);  // list overflow: exact all-pairs fallback for this particle
            }
#pragma unroll
            for (int k = 0; k < DIM; k++) f[k] = F[k];
            if (a.ensemble != 2) {
                // ---- second half kick (src/integrate.jl:28-38)
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    v[k] += (F[k] * a.dt) * 0.5;
                    ke2 = (k == 0) ? v[k] * v[k] : ke2 + v[k] * v[k];
                }
            } else {
                // ---- Brownian move (src/integrate.jl:66-82, intended semantics)
                double noise[3];
                brownian_noise<DIM>(a.seed, rng_step, pid, noise);
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    double xv = x[k] + (F[k] * a.dt / a.ktemp) + (noise[k] * sigma_bd);
                    double del = xv - x[k];
                    bd2 = (k == 0) ? del * del : bd2 + del * del;
                    double frac = g.invL[k] * xv;
                    double ncr = floor(frac);
                    if (ncr != 0.0) im[k] += (int32_t)ncr;
                    x[k] = g.L[k] * (frac - ncr);
                }
            }
        }
        if (a.ensemble == 2) {
            par ^= 1;
            if (active) pbuf[par][i] = make_double4(x[0], x[1], x[2], sig);
        }
        {
            double vals[4] = {e, w, np, ke2};
            block_fold(vals, 4, false, part1 + blockIdx.x * 5);
            double m[1] = {bd2};
            block_fold(m, 1, true, part1 + blockIdx.x * 5 + 4);
        }
        grid.sync();  // barrier B: per-CTA sums (and Brownian positions) are visible everywhere
        // ---- thermo scalars and thermostat (src/thermostat.jl:20-67, src/simulation.jl:118-131): fixed CTA order
        // the first warp folds the G per-CTA partials: lane b takes CTA b (G <= 64: at most two per lane), then a fixed
        // butterfly -- one L2 round trip instead of G dependent ones in a single thread (2.5 us of a 17 us step at N = 1024);
        // every CTA computes the same sums from the same numbers in the same order
        double r[4] = {0.0, 0.0, 0.0, 0.0}, dm = 0.0;
        if (tid < 32) {
            for (int b = tid; b < G; b += 32) {
#pragma unroll
                for (int c = 0; c < 4; c++) r[c] += __ldcg(part1 + b * 5 + c);
                dm = fmax(dm, __ldcg(part1 + b * 5 + 4));
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
                for (int c = 0; c < 4; c++) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
                dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
            }
        }
        if (tid == 0) {
            double U = 0.5 * r[0], W = 0.5 * r[1], NP = 0.5 * r[2], KE = r[3] / 2.0;
            double scale = 1.0;
            if (a.ensemble == 1) {
                ThermoRng rng;
                rng.init(a.seed, rng_step);
                double r1 = rng.normal();
                double r2 = rng.sum_noises(a.nf - 1.0);
                scale = bussi_scale(KE, a.ktemp_per_step[step], a.nf, a.dt, a.tau, r1, r2);
                KE = (scale * scale) * KE;
            }
            if (a.ensemble == 2) KE = 0.0;
            s_alpha = scale;
            red[0][0] = dm;
            if (blockIdx.x == 0) {
                ctl->last[0] = U; ctl->last[1] = W; ctl->last[2] = KE; ctl->last[3] = NP;
                if (!(isfinite(U) && isfinite(KE))) ctl->nonfinite = 1;
                if (a.thermo) {
                    a.thermo[4 * step + 0] = U; a.thermo[4 * step + 1] = W; a.thermo[4 * step + 2] = KE; a.thermo[4 * step + 3] = NP;
                }
            }
        }
        __syncthreads();
        alpha = s_alpha;
        dmax2_prev = red[0][0];
        rng_step++;
        __syncthreads();
    }
    // ---- leave the resident state as the step loop would: pending Bussi scale applied, everything back in global memory
    if (active) {
        s.pos[i] = make_double4(x[0], x[1], x[2], sig);
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            s.vel[k * s.cap + i] = (a.ensemble == 1) ? v[k] * alpha : v[k];
            s.frc[k * s.cap + i] = f[k];
            s.img[k * s.cap + i] = im[k];
        }
    }
    if (i == 0) {
        ctl->rng_step = rng_step;
        ctl->list_valid = 0;  // the large-system neighbour structures no longer match the positions
        ctl->dmax2_bits = 0ull;
        ctl->disp = 0.0;
    }
}


// ------------------------------------------------------------------------------------------------
// K0-small, thread-block-cluster version (default for N <= 2048): the same step loop with the CTAs of ONE cluster (<= 16
// CTAs on the SMs of one GPC) instead of a cooperative grid, LPP lanes of a warp per particle, and no barrier in the step.
//   * nothing is staged from L2: the lanes of a particle PUSH its new position into the position table of every CTA of the
//     cluster through distributed shared memory (st.async ... mbarrier::complete_tx::bytes, SASS STAS), so the neighbour
//     gathers and the all-pairs list rebuild read local shared memory only;
//   * each CTA waits on its OWN byte-counting mbarrier for an exchange to be complete: "A" (velocity Verlet) = the drifted
//     positions + the per-warp velocity maxima, "B" = the per-warp partial sums (+ the moved positions, Brownian dynamics).
//     Thread 0 re-arms a barrier (arrive.expect_tx) right after its wait; bytes that land before that are counted against
//     the new phase, which cannot complete without thread 0's arrival;
//   * position tables, maxima and partial sums alternate between two buffers by step parity, and the two exchanges of a step
//     order each other, so nothing is overwritten while it is still read and no phase is overrun:
//       - a peer pushes positions(s+1) only after its B(s) wait, i.e. after it received THIS CTA's sums of step s, which this
//         CTA sends after its pair loop of step s -- the last reader of the table positions(s+1) replaces (the one of s-1);
//       - a peer pushes sums(s) only after its A(s) wait, i.e. after it received this CTA's positions(s), which this CTA
//         sends after folding sums(s-1) -- and sums(s-2), the previous content of that buffer, before that;
//       - Brownian dynamics has exchange B only: a peer pushes sums(s) and positions(s+1) after its B(s-1) wait, i.e. after
//         it received this CTA's sums(s-1), sent after this CTA's pair loop of step s-1 and its fold of sums(s-2);
//       - every warp of the cluster contributes bytes to every phase, so no warp can fall a whole phase behind
//         (mbarrier.try_wait.parity tells the current phase from the previous one only);
//     the barrier.cluster version of this loop spent 21 % of its samples in ERRBAR / barrier wait (ncu);
//   * a step of a small system is one dependent chain per thread (ncu: half of the step was the pair loop at ~14 iterations
//     per warp, most of them entering the interaction branch for a few lanes).  The LPP lanes of a particle hold the same
//     state (they repeat the cheap kick / drift / move arithmetic, bit for bit), split the candidates j = q, q + LPP, ...
//     between them -- list rebuild and pair loop -- and add their partial forces with a fixed butterfly, so every lane of
//     the group ends with the same force bits.  LPP = 1 is the cooperative kernel's arithmetic statement by statement;
//   * reductions: every warp folds the per-warp partials itself in one fixed order (lane l takes warps l, l+32, ... in
//     sequence, then a fixed butterfly) -- deterministic for a given (block, LPP) shape, no block-level barrier;
//   * NVT: the thermostat's Gaussian and chi-square draws depend on (seed, step) only; all threads draw them for the whole
//     chunk up front, and every thread evaluates the Bussi scale from the same numbers.
// ------------------------------------------------------------------------------------------------
constexpr int kSmallClusterMaxN = 2048;           // two position tables of 32 B per particle in shared memory
constexpr int kSmallClusterMaxBlock = 320;        // 204 registers per thread available (the 3-D kernels use ~180)
constexpr int kSmallMaxWarps = 16 * kSmallClusterMaxBlock / 32;

__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// address of `local` (a shared-memory address of this CTA) in the shared memory of CTA `rank` of the cluster
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local), "r"(rank));
    return r;
}
// 16 bytes into a peer CTA's shared memory; the peer's mbarrier counts the bytes as they land (no fence, no cluster barrier)
__device__ __forceinline__ void st_async_f64x2(uint32_t remote, double lo, double hi, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];" ::"r"(remote), "d"(lo), "d"(hi),
                 "r"(remote_bar)
                 : "memory");
}
__device__ __forceinline__ void st_async_f64(uint32_t remote, double val, uint32_t remote_bar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];" ::"r"(remote), "d"(val), "r"(remote_bar) : "memory");
}

template <int DIM, class Pot, int LPP>
__global__ void __launch_bounds__(kSmallClusterMaxBlock, 1)
k_small_cluster(DevCtl *__restrict__ ctl, Grid g, SmallArgs a, Pot pot, PotParams pp)
{
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    extern __shared__ __align__(32) double4 spos_all[];             // [2][n] positions of ALL particles, pushed by their owners
    __shared__ __align__(16) double part0[2][kSmallMaxWarps];       // per-warp max |v|^2 (pushed by every warp of the cluster)
    __shared__ __align__(16) double part1[2][kSmallMaxWarps][6];    // per-warp {e, w, np, ke2, dmax2, 0}
    __shared__ __align__(8) unsigned long long barA, barB;          // byte-counting barriers of the two exchanges of a step
    const StatePtrs s = ctl->st[ctl->cur];
    const int n = a.n, tid = threadIdx.x, lane = tid & 31;
    const int C = (int)gridDim.x;                     // the grid is one cluster
    const int t = blockIdx.x * blockDim.x + tid;      // thread number in the cluster
    const int i = t / LPP, q = t % LPP;               // particle and the lane's place among the particle's lanes
    const int gw = t >> 5, NW = (C * (int)blockDim.x) >> 5;
    const bool active = i < n;
    uint32_t *const nl = a.nl + ((size_t)(active ? i : 0) * LPP + q);   // [k][i][q]: a warp's entries of one k are contiguous
    const size_t nl_stride = (size_t)n * LPP;
    const uint32_t barA_u = smem_u32(&barA), barB_u = smem_u32(&barB);
    // bytes every CTA receives per exchange.  A (velocity Verlet): the drifted positions and the per-warp velocity maxima;
    // B: the per-warp sums (and, Brownian dynamics, the moved positions)
    const uint32_t bytesA = (uint32_t)n * 32u + (uint32_t)NW * 8u;
    const uint32_t bytesB = (uint32_t)NW * 48u + (a.ensemble == 2 ? (uint32_t)n * 32u : 0u);

    double x[3] = {0, 0, 0}, v[3] = {0, 0, 0}, f[3] = {0, 0, 0}, sig = 1.0;
    int32_t im[3] = {0, 0, 0};
    uint32_t pid = 0;
    if (active) {
        double4 p = s.pos[i];
        x[0] = p.x; x[1] = p.y; x[2] = p.z; sig = p.w;
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            v[k] = s.vel[k * s.cap + i];
            f[k] = s.frc[k * s.cap + i];
            im[k] = s.img[k * s.cap + i];
        }
        pid = (uint32_t)s.id[i];
    }
    int cnt = 0;
    double alpha = 1.0, disp = 0.0;
    bool rebuild = true;          // the list of a previous call is never trusted
    double dmax2_prev = 0.0;      // Brownian: displacement bound of the previous step's move (cluster-wide)
    unsigned long long rng_step = ctl->rng_step;
    const double sigma_bd = sqrt(2.0 * a.dt);

    // the particle's lanes share the fan-out; table `buf`, counted by the peers' barrier `bar`
    auto push_position = [&](int buf, uint32_t bar) {
        if (active) {
            const uint32_t slot = smem_u32(spos_all + (size_t)buf * n + i);
            for (int r = q; r < C; r += LPP) {
                const uint32_t dst = dsmem_addr(slot, (uint32_t)r), rb = dsmem_addr(bar, (uint32_t)r);
                st_async_f64x2(dst, x[0], x[1], rb);
                st_async_f64x2(dst + 16, x[2], sig, rb);
            }
        }
    };
    if (tid == 0) {
        mbar_init(&barA, 1);
        mbar_init(&barB, 1);
        mbar_fence_init();
        if (a.ensemble != 2) mbar_expect_tx(&barA, bytesA);
        mbar_expect_tx(&barB, bytesB);
    }
    if (active)           // Brownian dynamics evaluates forces before it moves anything: table 0, plain remote stores
        for (int r = q; r < C; r += LPP) cluster.map_shared_rank(spos_all, r)[i] = make_double4(x[0], x[1], x[2], sig);
    if (a.ensemble == 1) {
        // the thermostat's random numbers depend on (seed, step) only: a serial chain of ~2.4 us per step when one thread
        // draws them between the exchanges -- here every thread of the cluster draws the pair of a few steps up front
        const int T = C * (int)blockDim.x;
        for (long long st = t; st < a.nsteps; st += T) {
            ThermoRng rng;
            rng.init(a.seed, rng_step + (unsigned long long)st);
            double r1 = rng.normal();
            double r2 = rng.sum_noises(a.nf - 1.0);
            a.rng_pre[2 * st] = r1;
            a.rng_pre[2 * st + 1] = r2;
        }
    }
    // the only cluster barrier before the end: barriers initialised and armed, table 0 filled, the draws visible (release /
    // acquire at cluster scope; the table is read through L2 below)
    cluster_sync_all();

    for (long long step = 0; step < a.nsteps; step++) {
        const int par = (int)(step & 1);
        const double4 *spos = spos_all + (size_t)par * n;    // this step's positions
        double dmax2;
        double2 rr = make_double2(0.0, 0.0);
        double kt_step = 0.0;
        if (a.ensemble == 1) {   // needed after exchange B: on their way from L2 during the whole step
            rr = __ldcg(reinterpret_cast<const double2 *>(a.rng_pre) + step);
            kt_step = a.ktemp_per_step[step];
        }
        if (a.ensemble != 2) {
            // ---- first half kick + drift + wrap (src/integrate.jl:8-21, src/boundary.jl:7-17); Bussi scale of the previous step
            double v2 = 0.0;
            if (active) {
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    double vk = v[k] * alpha;
                    vk += (f[k] * a.dt) * 0.5;
                    v[k] = vk;
                    v2 = (k == 0) ? vk * vk : v2 + vk * vk;
                    double xv = x[k] + vk * a.dt;
                    double frac = g.invL[k] * xv;
                    double ncr = floor(frac);
                    if (ncr != 0.0) im[k] += (int32_t)ncr;
                    x[k] = g.L[k] * (frac - ncr);
                }
            }
            // ---- exchange A: positions into table `par` and the per-warp maxima, counted by every peer's barrier A
            push_position(par, barA_u);
            v2 = warp_max(v2);
            {
                const uint32_t slot = smem_u32(&part0[par][gw]);
                for (int r = lane; r < C; r += 32) st_async_f64(dsmem_addr(slot, (uint32_t)r), v2, dsmem_addr(barA_u, (uint32_t)r));
            }
            mbar_wait(&barA, (uint32_t)par);
            if (tid == 0) mbar_expect_tx(&barA, bytesA);     // arm the next phase (bytes that land early are counted against it)
            dmax2 = 0.0;
            for (int wq = lane; wq < NW; wq += 32) dmax2 = fmax(dmax2, part0[par][wq]);
            dmax2 = warp_max(dmax2) * (a.dt * a.dt);
        } else {
            dmax2 = dmax2_prev;
        }
        {
            double d = disp + sqrt(dmax2);
            rebuild = rebuild || !(2.0 * d <= a.skin);
            disp = rebuild ? 0.0 : d;
        }
        const double4 pi = make_double4(x[0], x[1], x[2], sig);
        // ---- Verlet list rebuild: all pairs from shared memory, FP32 membership with margin (superset); lane q keeps the
        // neighbours j = q (mod LPP) in ascending order
        if (rebuild) {
            if (active) {
                const float xi = (float)pi.x, yi = (float)pi.y, zi = (float)pi.z;
                const float Lx = (float)g.L[0], Ly = (float)g.L[1], Lz = (float)g.L[2];
                const float hx = 0.5f * Lx, hy = 0.5f * Ly, hz = 0.5f * Lz;
                cnt = 0;
                for (int j = q; j < n; j += LPP) {
                    const double4 pj = spos[j];
                    float dx = xi - (float)pj.x, dy = yi - (float)pj.y;
                    dx = dx > hx ? dx - Lx : (dx < -hx ? dx + Lx : dx);
                    dy = dy > hy ? dy - Ly : (dy < -hy ? dy + Ly : dy);
                    float d2 = dx * dx + dy * dy;
                    if (DIM == 3) {
                        float dz = zi - (float)pj.z;
                        dz = dz > hz ? dz - Lz : (dz < -hz ? dz + Lz : dz);
                        d2 += dz * dz;
                    }
                    if (d2 <= a.rlist2_f && j != i) {
                        if (cnt < a.kmax) nl[(size_t)cnt * nl_stride] = (uint32_t)j;
                        cnt++;
                    }
                }
            }
            if (t == 0) ctl->rebuilds += 1;
            rebuild = false;
        }
        // ---- pair forces (src/pairwise.jl:26-39)
        double e = 0.0, w = 0.0, np = 0.0, ke2 = 0.0, bd2 = 0.0;
        double F[3] = {0.0, 0.0, 0.0};
        if (active) {
            auto candidate = [&](int j) {
                const double4 pj = spos[j];
                double dx, dy, dz;
                double d2 = separation_wrap<DIM>(g, pi, pj, dx, dy, dz);
                if (d2 <= a.cutoff2 && pot.may_interact(pp, d2, pi.w, pj.w)) pair_accumulate<DIM>(pot, pp, dx, dy, dz, d2, pi.w, pj.w, F, e, w, np);
            };
            if (cnt <= a.kmax) {
                uint32_t jn = cnt > 0 ? nl[0] : 0u;
                for (int k = 0; k < cnt; k++) {
                    const uint32_t j = jn;
                    if (k + 1 < cnt) jn = nl[(size_t)(k + 1) * nl_stride];   // the next index is on its way while this pair is evaluated
                    candidate((int)j);
                }
            } else {
                for (int j = q; j < n; j += LPP)
                    if (j != i) candidate(j);  // list overflow: exact scan of this lane's share
            }
        }
        if (LPP > 1) {
            // the particle's lanes add their shares with a fixed butterfly: every lane ends with the same bits
#pragma unroll
            for (int o = LPP / 2; o > 0; o >>= 1) {
#pragma unroll
                for (int k = 0; k < DIM; k++) F[k] += __shfl_xor_sync(0xffffffffu, F[k], o);
            }
        }
        if (active) {
#pragma unroll
            for (int k = 0; k < DIM; k++) f[k] = F[k];
            if (a.ensemble != 2) {
                // ---- second half kick (src/integrate.jl:28-38)
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    v[k] += (F[k] * a.dt) * 0.5;
                    ke2 = (k == 0) ? v[k] * v[k] : ke2 + v[k] * v[k];
                }
            } else {
                // ---- Brownian move (src/integrate.jl:66-82, intended semantics)
                double noise[3];
                brownian_noise<DIM>(a.seed, rng_step, pid, noise);
#pragma unroll
                for (int k = 0; k < DIM; k++) {
                    double xv = x[k] + (F[k] * a.dt / a.ktemp) + (noise[k] * sigma_bd);
                    double del = xv - x[k];
                    bd2 = (k == 0) ? del * del : bd2 + del * del;
                    double frac = g.invL[k] * xv;
                    double ncr = floor(frac);
                    if (ncr != 0.0) im[k] += (int32_t)ncr;
                    x[k] = g.L[k] * (frac - ncr);
                }
            }
        }
        // ---- exchange B: the per-warp sums (Brownian dynamics: and the moved positions, into the OTHER table -- peers may
        // still be reading this step's), counted by every peer's barrier B
        if (a.ensemble == 2) push_position(par ^ 1, barB_u);
        {
            // e, w, np are per-lane shares (their warp sum is the sum over the warp's particles); the kinetic term is the
            // same in all lanes of a particle and is counted once
            const double s0 = warp_sum(e), s1 = warp_sum(w), s2 = warp_sum(np), s3 = warp_sum(q == 0 ? ke2 : 0.0), s4 = warp_max(bd2);
            const uint32_t slot = smem_u32(&part1[par][gw][0]);
            for (int r = lane; r < C; r += 32) {
                const uint32_t dst = dsmem_addr(slot, (uint32_t)r), rb = dsmem_addr(barB_u, (uint32_t)r);
                st_async_f64x2(dst, s0, s1, rb);
                st_async_f64x2(dst + 16, s2, s3, rb);
                st_async_f64x2(dst + 32, s4, 0.0, rb);
            }
        }
        mbar_wait(&barB, (uint32_t)par);
        if (tid == 0) mbar_expect_tx(&barB, bytesB);
        // ---- thermo scalars and thermostat (src/thermostat.jl:20-67, src/simulation.jl:118-131): every warp folds the same
        // numbers in the same order
        double r[4] = {0.0, 0.0, 0.0, 0.0}, dm = 0.0;
        for (int wq = lane; wq < NW; wq += 32) {
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] += part1[par][wq][c];
            dm = fmax(dm, part1[par][wq][4]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int c = 0; c < 4; c++) r[c] += __shfl_xor_sync(0xffffffffu, r[c], o);
            dm = fmax(dm, __shfl_xor_sync(0xffffffffu, dm, o));
        }
        dmax2_prev = dm;
        if (a.ensemble == 1) {
            // every thread evaluates the scale from the same numbers (no block barrier, no broadcast)
            alpha = bussi_scale(r[3] / 2.0, kt_step, a.nf, a.dt, a.tau, rr.x, rr.y);
        }
        if (t == 0) {
            double U = 0.5 * r[0], W = 0.5 * r[1], NP = 0.5 * r[2], KE = r[3] / 2.0;
            if (a.ensemble == 1) KE = (alpha * alpha) * KE;
            if (a.ensemble == 2) KE = 0.0;
            ctl->last[0] = U; ctl->last[1] = W; ctl->last[2] = KE; ctl->last[3] = NP;
            if (!(isfinite(U) && isfinite(KE))) ctl->nonfinite = 1;
            if (a.thermo) {
                a.thermo[4 * step + 0] = U; a.thermo[4 * step + 1] = W; a.thermo[4 * step + 2] = KE; a.thermo[4 * step + 3] = NP;
            }
        }
        rng_step++;
    }
    // ---- leave the resident state as the step loop would: pending Bussi scale applied, everything back in global memory
    if (active && q == 0) {
        s.pos[i] = make_double4(x[0], x[1], x[2], sig);
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            s.vel[k * s.cap + i] = (a.ensemble == 1) ? v[k] * alpha : v[k];
            s.frc[k * s.cap + i] = f[k];
            s.img[k * s.cap + i] = im[k];
        }
    }
    if (t == 0) {
        ctl->rng_step = rng_step;
        ctl->list_valid = 0;  // the large-system neighbour structures no longer match the positions
        ctl->dmax2_bits = 0ull;
        ctl->disp = 0.0;
    }
    cluster_sync_all();           // no CTA may exit while peers can still address its shared memory
}

}  // namespace mdb
