// rng.cuh -- Philox4x32-10 counter-based RNG and the thermostat / Brownian streams.
// Replaces Random.Xoshiro + Distributions.Gamma on the hot path (/root/reference/src/thermostat.jl:1-18,35-36;
// src/integrate.jl:55-59) as north_star prescribes.  Stream layout is specified in oracle/md_oracle.c
// ("Counter-based RNG") and is identical on both sides.
#pragma once
#ifndef __CUDACC_RTC__
#include <cstdint>
#include <cuda_runtime.h>
#else
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
#endif

namespace mdb {

constexpr uint32_t kTagBussi = 0xB0551u;
constexpr uint32_t kTagBrown = 0xB12Du;

struct Philox4 {
    uint32_t w[4];
};

__host__ __device__ __forceinline__ Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                         uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        if (r > 0) {
            k0 += 0x9E3779B9u;
            k1 += 0xBB67AE85u;
        }
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0;
        c1 = n1;
        c2 = n2;
        c3 = n3;
    }
    Philox4 o;
    o.w[0] = c0;
    o.w[1] = c1;
    o.w[2] = c2;
    o.w[3] = c3;
    return o;
}

__host__ __device__ __forceinline__ double u53(uint32_t hi, uint32_t lo)
{
    return (double)((((uint64_t)hi << 32) | lo) >> 11) * 0x1.0p-53;
}
__host__ __device__ __forceinline__ double u53_open(uint32_t hi, uint32_t lo)
{
    return (double)(((((uint64_t)hi << 32) | lo) >> 11) + 1) * 0x1.0p-53;
}

// sequential sampler for the thermostat (one thread per step)
struct ThermoRng {
    uint32_t k0, k1;
    uint64_t step;
    uint32_t nblock, ublock;
    int nhave, uhave;
    double nbuf[2], ubuf[2];
    __device__ void init(uint64_t seed, uint64_t s)
    {
        k0 = (uint32_t)seed;
        k1 = (uint32_t)(seed >> 32);
        step = s;
        nblock = ublock = 0;
        nhave = uhave = 0;
    }
    __device__ double normal()
    {
        if (nhave == 0) {
            Philox4 o = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), nblock++, (kTagBussi << 8) | 0u, k0, k1);
            double u1 = u53_open(o.w[0], o.w[1]), u2 = u53(o.w[2], o.w[3]);
            double r = sqrt(-2.0 * log(u1));
            double th = 6.283185307179586 * u2;
            nbuf[0] = r * cos(th);
            nbuf[1] = r * sin(th);
            nhave = 2;
        }
        double v = nbuf[2 - nhave];
        nhave--;
        return v;
    }
    __device__ double uniform()
    {
        if (uhave == 0) {
            Philox4 o = philox4x32_10((uint32_t)step, (uint32_t)(step >> 32), ublock++, (kTagBussi << 8) | 1u, k0, k1);
            ubuf[0] = u53_open(o.w[0], o.w[1]);
            ubuf[1] = u53_open(o.w[2], o.w[3]);
            uhave = 2;
        }
        double v = ubuf[2 - uhave];
        uhave--;
        return v;
    }
    // Marsaglia-Tsang Gamma(k, 1), k >= 1
    __device__ double gamma(double k)
    {
        double d = k - 1.0 / 3.0;
        double c = 1.0 / sqrt(9.0 * d);
        for (;;) {
            double x = normal();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            double u = uniform();
            double x2 = x * x;
            if (u < 1.0 - 0.0331 * x2 * x2) return d * v;
            if (log(u) < 0.5 * x2 + d * (1.0 - v + log(v))) return d * v;
        }
    }
    // sum_noises(nf, rng): src/thermostat.jl:1-18
    __device__ double sum_noises(double nf)
    {
        if (nf == 0.0) return 0.0;
        if (nf == 1.0) {
            double z = normal();
            return z * z;
        }
        if (fmod(nf, 2.0) == 0.0) return 2.0 * gamma(floor(nf / 2.0));
        double r = 2.0 * gamma(floor((nf - 1.0) / 2.0));
        double z = normal();
        return r + z * z;
    }
};

// scale factor of bussi!: src/thermostat.jl:21-43
__device__ __forceinline__ double bussi_scale(double kinetic_energy, double ktemp, double nf, double dt, double tau, double r1,
                                              double r2)
{
    double dt_ratio = dt / tau;
    double current_temperature = 2.0 * kinetic_energy / nf;
    double term_1 = exp(-dt_ratio);
    double c2 = (1.0 - term_1) * ktemp / (current_temperature * nf);
    double term_2 = c2 * (r2 + r1 * r1);
    double term_3 = 2.0 * r1 * sqrt(term_1 * c2);
    return sqrt(term_1 + term_2 + term_3);
}

// sample_uniform!: src/integrate.jl:55-59, keyed by (particle id, step)
template <int DIM>
__device__ __forceinline__ void brownian_noise(uint64_t seed, uint64_t step, uint32_t id, double *noise)
{
    const double sqthree = 1.7320508075688772;  // sqrt(3.0)
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    Philox4 o = philox4x32_10(id, (uint32_t)step, (uint32_t)(step >> 32), (kTagBrown << 8) | 0u, k0, k1);
    noise[0] = (2.0 * u53(o.w[0], o.w[1]) - 1.0) * sqthree;
    noise[1] = (2.0 * u53(o.w[2], o.w[3]) - 1.0) * sqthree;
    if (DIM == 3) {
        Philox4 q = philox4x32_10(id, (uint32_t)step, (uint32_t)(step >> 32), (kTagBrown << 8) | 1u, k0, k1);
        noise[2] = (2.0 * u53(q.w[0], q.w[1]) - 1.0) * sqthree;
    }
}

}  // namespace mdb
