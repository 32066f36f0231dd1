// setup_io.cuh -- kernels of the steps either side of the hot path:
//   * trajectory frames: unwrapped coordinates p + U*img (/root/reference/src/io.jl:62-70) packed on the device in the
//     caller's particle order, ready for write_to_file_lammps' columns (src/io.jl:97-167);
//   * initialize_velocities (src/initialization.jl:32-47) on the device with the counter-based RNG;
//   * slot-order images of the state for the exact binary checkpoint.
// Included by engine.cu only (never by NVRTC).
#pragma once
#include "kernels.cuh"

namespace mdb {

constexpr uint32_t kTagVel = 0x1E10Cu;  // velocity-initialisation stream (spec: oracle/md_oracle.c "Counter-based RNG")

// frame record of particle `id` (original order): {radius, x[DIM], xu[DIM]}; radius = diameter / 2 as written at
// src/io.jl:141,153; xu = x + U*img.
template <int DIM>
__global__ void __launch_bounds__(kStreamBlock)
k_pack_frame(int64_t n, const DevCtl *__restrict__ ctl, Grid g, double *__restrict__ frame)
{
    const StatePtrs s = ctl->st[ctl->cur];
    constexpr int W = 2 * DIM + 1;
    for (int64_t i = blockIdx.x * (int64_t)kStreamBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStreamBlock) {
        const double4 p = ld_pos(&s.pos[i]);
        const double x[3] = {p.x, p.y, p.z};
        double *out = frame + (int64_t)s.id[i] * W;
        out[0] = p.w / 2.0;
        int32_t im[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < DIM; k++) im[k] = s.img[k * s.cap + i];
        double xu[3];
        unwrap_point<DIM>(g, x, im, xu);
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            out[1 + k] = x[k];
            out[1 + DIM + k] = xu[k];
        }
    }
}

// x-slab handles: the rank's owned particles in SLOT order plus their original ids (the writer prints id + 1 per row; a frame
// is one file per rank, LAMMPS' "file.%" multi-file dump convention)
template <int DIM>
__global__ void __launch_bounds__(kStreamBlock)
k_pack_frame_slab(const DevCtl *__restrict__ ctl, Grid g, double *__restrict__ frame, int32_t *__restrict__ ids)
{
    const StatePtrs s = ctl->st[ctl->cur];
    const int64_t n = ctl->n_own;
    constexpr int W = 2 * DIM + 1;
    for (int64_t i = blockIdx.x * (int64_t)kStreamBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStreamBlock) {
        const double4 p = ld_pos(&s.pos[i]);
        const double x[3] = {p.x, p.y, p.z};
        double *out = frame + i * W;
        out[0] = p.w / 2.0;
        int32_t im[3] = {0, 0, 0};
#pragma unroll
        for (int k = 0; k < DIM; k++) im[k] = s.img[k * s.cap + i];
        double xu[3];
        unwrap_point<DIM>(g, x, im, xu);
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            out[1 + k] = x[k];
            out[1 + DIM + k] = xu[k];
        }
        ids[i] = s.id[i];
    }
}

// standard normals keyed by (particle id, stream): ctr = (id, stream_lo, stream_hi, kTagVel<<8 | block), Box-Muller on
// (u53_open(w0,w1), u53(w2,w3)); block 0 gives (v_x, v_y), block 1 gives v_z
template <int DIM>
__device__ __forceinline__ void velocity_normals(uint64_t seed, uint64_t stream, uint32_t id, double *v)
{
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    Philox4 o = philox4x32_10(id, (uint32_t)stream, (uint32_t)(stream >> 32), (kTagVel << 8) | 0u, k0, k1);
    double r = sqrt(-2.0 * log(u53_open(o.w[0], o.w[1])));
    double th = 6.283185307179586 * u53(o.w[2], o.w[3]);
    v[0] = r * cos(th);
    v[1] = r * sin(th);
    if (DIM == 3) {
        Philox4 q = philox4x32_10(id, (uint32_t)stream, (uint32_t)(stream >> 32), (kTagVel << 8) | 1u, k0, k1);
        double r2 = sqrt(-2.0 * log(u53_open(q.w[0], q.w[1])));
        v[2] = r2 * cos(6.283185307179586 * u53(q.w[2], q.w[3]));
    }
}

// initialize_velocities in three sweeps with deterministic two-stage reductions (per-CTA partials in `part`, combined
// in fixed order by k_vel_reduce):  stage 0  V = randn            -> partial sums of V        (src/initialization.jl:34)
//                                   stage 1  V -= mean(V)         -> partial sums of V.^2     (:36-38)
//                                   stage 2  V *= fs                                          (:40-42)
template <int DIM>
__global__ void __launch_bounds__(kStreamBlock)
k_vel_init(int stage, int64_t n, uint64_t seed, uint64_t stream, DevCtl *__restrict__ ctl, double *__restrict__ part)
{
    const StatePtrs s = ctl->st[ctl->cur];
    double acc[3] = {0.0, 0.0, 0.0};
    const double m[3] = {ctl->scratch[0], ctl->scratch[1], ctl->scratch[2]};
    const double fs = ctl->scratch[3];
    if (n < 0) n = ctl->n_own;  // x-slab handles
    for (int64_t i = blockIdx.x * (int64_t)kStreamBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStreamBlock) {
        double v[3] = {0.0, 0.0, 0.0};
        if (stage == 0) {
            velocity_normals<DIM>(seed, stream, (uint32_t)s.id[i], v);
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                s.vel[k * s.cap + i] = v[k];
                acc[k] += v[k];
            }
        } else if (stage == 1) {
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                double w = s.vel[k * s.cap + i] - m[k];
                s.vel[k * s.cap + i] = w;
                acc[k] += w * w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < DIM; k++) s.vel[k * s.cap + i] = s.vel[k * s.cap + i] * fs;
        }
    }
    if (stage < 2) {
        block_reduce<3, kStreamBlock>(acc);
        if (threadIdx.x == 0) {
#pragma unroll
            for (int k = 0; k < 3; k++) part[k * kMaxPartials + blockIdx.x] = acc[k];
        }
    }
}
// stage 0: scratch[0..2] = mean per component; stage 1: scratch[3] = fs = sqrt(ktemp / (sum_v2 / ((N-1) dim)))
// slab: 1 = leave this rank's sums in ctl->red[0..2] for the all-reduce, 2 = continue from the globally summed ctl->red
__global__ void k_vel_reduce(int stage, int nslots, const double *__restrict__ part, double n_particles, int dim, double ktemp, DevCtl *ctl,
                             int slab = 0)
{
    double r[3] = {0.0, 0.0, 0.0};
    if (slab != 2) {
        for (int q = threadIdx.x; q < nslots; q += blockDim.x) {
#pragma unroll
            for (int k = 0; k < 3; k++) r[k] += part[k * kMaxPartials + q];
        }
        block_reduce<3, kStreamBlock>(r);
    }
    if (threadIdx.x == 0 && slab == 1) {
        for (int k = 0; k < 3; k++) ctl->red[k] = r[k];
        ctl->red[3] = 0.0;
        return;
    }
    if (threadIdx.x == 0) {
        if (slab == 2)
            for (int k = 0; k < 3; k++) r[k] = ctl->red[k];
        if (stage == 0) {
            for (int k = 0; k < 3; k++) ctl->scratch[k] = r[k] / n_particles;
        } else {
            double sum_v2 = r[0] + r[1] + r[2];
            ctl->scratch[3] = sqrt(ktemp / (sum_v2 / ((n_particles - 1.0) * dim)));
        }
    }
}

// slot-order copies for the checkpoint: compact SoA rows [DIM][n] out of / into the capacity-strided state arrays
template <int DIM>
__global__ void k_ckpt_pack(int64_t n, const DevCtl *__restrict__ ctl, double4 *__restrict__ pos, double *__restrict__ vel,
                            double *__restrict__ frc, int32_t *__restrict__ img, int32_t *__restrict__ id)
{
    const StatePtrs s = ctl->st[ctl->cur];
    for (int64_t i = blockIdx.x * (int64_t)kStreamBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStreamBlock) {
        pos[i] = s.pos[i];
        id[i] = s.id[i];
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            vel[k * n + i] = s.vel[k * s.cap + i];
            frc[k * n + i] = s.frc[k * s.cap + i];
            img[k * n + i] = s.img[k * s.cap + i];
        }
    }
}
template <int DIM>
__global__ void k_ckpt_unpack(int64_t n, StatePtrs s, const double4 *__restrict__ pos, const double *__restrict__ vel,
                              const double *__restrict__ frc, const int32_t *__restrict__ img, const int32_t *__restrict__ id)
{
    for (int64_t i = blockIdx.x * (int64_t)kStreamBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStreamBlock) {
        s.pos[i] = pos[i];
        s.id[i] = id[i];
#pragma unroll
        for (int k = 0; k < DIM; k++) {
            s.vel[k * s.cap + i] = vel[k * n + i];
            s.frc[k * s.cap + i] = frc[k * n + i];
            s.img[k * s.cap + i] = img[k * n + i];
        }
    }
}

// uniform positions in the cell keyed by (particle id, stream): ctr = (id, stream_lo, stream_hi, kTagPos<<8 | block);
// block 0 gives u_x = (w0,w1), u_y = (w2,w3), block 1 gives u_z = (w0,w1); x_k = L_k * u53 (0 if that rounds up to L_k).
// Images are reset and the neighbour structures invalidated.  (rand(rng, dim) .* (maxs .- mins) .+ mins with mins = 0,
// src/initialization.jl:22-27)
constexpr uint32_t kTagPos = 0x9051u;
template <int DIM>
__global__ void __launch_bounds__(kStreamBlock)
k_random_positions(int64_t n, Grid g, uint64_t seed, uint64_t stream, DevCtl *__restrict__ ctl)
{
    const StatePtrs s = ctl->st[ctl->cur];
    const uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int64_t i = blockIdx.x * (int64_t)kStreamBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kStreamBlock) {
        const uint32_t id = (uint32_t)s.id[i];
        Philox4 o = philox4x32_10(id, (uint32_t)stream, (uint32_t)(stream >> 32), (kTagPos << 8) | 0u, k0, k1);
        double u[3] = {u53(o.w[0], o.w[1]), u53(o.w[2], o.w[3]), 0.0};
        if (DIM == 3) {
            Philox4 q = philox4x32_10(id, (uint32_t)stream, (uint32_t)(stream >> 32), (kTagPos << 8) | 1u, k0, k1);
            u[2] = u53(q.w[0], q.w[1]);
        }
        double x[3] = {0.0, 0.0, 0.0};
        if (!g.tri) {
#pragma unroll
            for (int k = 0; k < DIM; k++) {
                x[k] = g.L[k] * u[k];
                if (!(x[k] < g.L[k])) x[k] = 0.0;
            }
        } else {  // uniform in the cell: uniform fractional coordinates (wrapped like any other position)
            double ncr[3];
            mat3_mul(g.U, u, x);
            wrap_point<DIM>(g, x, ncr);
        }
#pragma unroll
        for (int k = 0; k < DIM; k++) s.img[k * s.cap + i] = 0;
        const double w = s.pos[i].w;
        s.pos[i] = make_double4(x[0], x[1], x[2], w);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctl->list_valid = 0;
        ctl->disp = 0.0;
        ctl->disp_in = 0.0;
    }
}

// DFMA throughput probe for the FP64-pipe figures in bench.py / BASELINE.md (not part of the path): 8 independent
// dependent-FMA chains per thread so the pipe, not the latency, is what is measured
__global__ void __launch_bounds__(256)
k_fp64_probe(int iters, double seed, double *__restrict__ sink)
{
    double a[8];
#pragma unroll
    for (int q = 0; q < 8; q++) a[q] = seed + (double)(threadIdx.x + q);
    const double m = 1.0000001, c = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int q = 0; q < 8; q++) a[q] = fma(a[q], m, c);
    }
    double r = 0.0;
#pragma unroll
    for (int q = 0; q < 8; q++) r += a[q];
    if (r == 123.456) sink[0] = r;  // keeps the chains alive
}

}  // namespace mdb
