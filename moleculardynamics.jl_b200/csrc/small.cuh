// small.cuh -- K0-small: the whole step loop of src/simulation.jl:88-108 / :231-250 as ONE persistent cooperative
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

}  // namespace mdb
