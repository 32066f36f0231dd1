// stage_probe.cu -- what does it cost to stage a tile's NEIGHBOUR-CELL ROWS in shared memory with TMA bulk copies?
// Evidence for the force-kernel design (profiles/r02_force_ab.md): VERDICT r01 asked for a kernel whose
// CTA tile = 128 consecutive slots bulk-copies (cp.async.bulk + mbarrier, double-buffered) the 9 contiguous neighbour
// x-rows (about 37 KB) into shared memory.  This probe does exactly that data movement on a cell-sorted fluid of the
// bench's density -- every tile's 9 (or 18, when the tile straddles two x-rows; plus the periodic wrap pieces) row
// segments, double-buffered, a token read of the staged records -- and nothing else: no list walk, no distance test, no
// force.  Its time is a LOWER bound for any kernel built on that staging; compare with the whole pair-force kernel.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stage_probe stage_probe.cu && ./stage_probe [log2 N]
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include <cuda_runtime.h>

#define CK(call)                                                                              \
    do {                                                                                      \
        cudaError_t _e = (call);                                                              \
        if (_e != cudaSuccess) {                                                              \
            printf("%s failed: %s (line %d)\n", #call, cudaGetErrorString(_e), __LINE__);     \
            return 2;                                                                         \
        }                                                                                     \
    } while (0)

constexpr int kTile = 128;
constexpr int kMaxSeg = 48;   // <= 64: the issuing warp takes two segments per lane
constexpr int kStageRecords = 1792;  // 56 KB per stage; 2 stages = 112 KB per CTA -> 2 CTAs per SM

struct Seg {
    uint32_t begin, count;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"((unsigned long long)__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}" ::"r"(
            smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// mode 0: stage the neighbour rows (TMA bulk copies) + token read;  mode 1: token work only (the tile's own records)
__global__ void __launch_bounds__(kTile) k_stage(int ntiles, const double4 *__restrict__ pos, const Seg *__restrict__ segs,
                                                const int *__restrict__ nseg, double *__restrict__ sink, int mode)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double4 *stage0 = (double4 *)smem_raw;
    double4 *stage1 = stage0 + kStageRecords;
    __shared__ uint64_t bar[2];
    __shared__ uint32_t total[2];
    const int tid = threadIdx.x;
    if (tid == 0) {
        mbar_init(&bar[0], 1);
        mbar_init(&bar[1], 1);
    }
    __syncthreads();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    double acc = 0.0;
    // warp 0 issues a tile: lane q owns segment q (kMaxSeg <= 64: two passes), offsets by a warp prefix sum, ONE expect_tx
    // for the whole tile, then every lane posts its own bulk copy -- no serial loop over the segments
    auto issue = [&](int tile, int buf) {
        if (tid < 32 && tile < ntiles) {
            const int ns = nseg[tile];
            const Seg *sg = segs + (size_t)tile * kMaxSeg;
            Seg mine[2];
            uint32_t cnt[2], off[2], run = 0;
            for (int h = 0; h < 2; h++) {
                const int q = h * 32 + tid;
                mine[h] = q < ns ? sg[q] : Seg{0u, 0u};
                cnt[h] = mine[h].count;
                uint32_t incl = cnt[h];
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += t;
                }
                off[h] = run + incl - cnt[h];
                run += __shfl_sync(0xffffffffu, incl, 31);
            }
            const uint32_t recs = min(run, (uint32_t)kStageRecords);
            if (tid == 0) {
                total[buf] = recs;
                mbar_expect_tx(&bar[buf], recs * 32u);
            }
            __syncwarp();
            double4 *dst = buf ? stage1 : stage0;
            for (int h = 0; h < 2; h++) {
                if (off[h] < recs) {
                    const uint32_t c = min(cnt[h], recs - off[h]);
                    if (c) bulk_g2s(dst + off[h], pos + mine[h].begin, c * 32u, &bar[buf]);
                }
            }
        }
    };
    int tile = blockIdx.x;
    uint32_t phase[2] = {0, 0};
    if (mode == 0) issue(tile, 0);
    int buf = 0;
    for (; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        if (mode == 0) {
            issue(tile + gridDim.x, buf ^ 1);
            mbar_wait(&bar[buf], phase[buf]);
            phase[buf] ^= 1;
            const double4 *st = buf ? stage1 : stage0;
            const uint32_t recs = total[buf];
            // token read: 8 staged records per thread (a real kernel reads ~2 list candidates + its own record)
            for (int q = 0; q < 8; q++) {
                const uint32_t k = (uint32_t)(tid * 13 + q * 211) % max(recs, 1u);
                acc += st[k].x + st[k].w;
            }
            __syncthreads();  // everybody is done with this stage before it is refilled two tiles later
        } else {
            const double4 p = pos[(size_t)tile * kTile + tid];
            acc += p.x + p.w;
        }
    }
    sink[blockIdx.x * kTile + tid] = acc;
}

int main(int argc, char **argv)
{
    const int lg = argc > 1 ? atoi(argv[1]) : 22;
    const int64_t N = 1ll << lg;
    const double rho = 0.8976338790382897, rgrid = 1.0204081632653061 * 1.25;
    const double L = std::cbrt((double)N / rho);
    const int nc = (int)std::floor(L / (rgrid * (1.0 + 1e-6)));
    const int64_t ncell = (int64_t)nc * nc * nc;
    printf("N = 2^%d, L = %.3f, %d^3 cells of %.4f (%.2f particles per cell)\n", lg, L, nc, L / nc, (double)N / ncell);
    std::mt19937_64 rng(20261018);
    std::uniform_real_distribution<double> U(0.0, L);
    std::vector<double4> p(N);
    std::vector<uint32_t> cell(N), order(N);
    for (int64_t i = 0; i < N; i++) {
        p[i] = make_double4(U(rng), U(rng), U(rng), 1.0);
        int cx = std::min(nc - 1, (int)(p[i].x * nc / L)), cy = std::min(nc - 1, (int)(p[i].y * nc / L)), cz = std::min(nc - 1, (int)(p[i].z * nc / L));
        cell[i] = ((uint32_t)cz * nc + cy) * nc + cx;
        order[i] = (uint32_t)i;
    }
    std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return cell[a] < cell[b]; });
    std::vector<double4> sorted(N);
    std::vector<uint32_t> start(ncell + 1, 0);
    for (int64_t s = 0; s < N; s++) {
        sorted[s] = p[order[s]];
        start[cell[order[s]] + 1]++;
    }
    for (int64_t c = 0; c < ncell; c++) start[c + 1] += start[c];
    std::vector<uint32_t> cell_of_slot(N);
    for (int64_t s = 0; s < N; s++) cell_of_slot[s] = cell[order[s]];
    const int ntiles = (int)(N / kTile);
    std::vector<Seg> segs((size_t)ntiles * kMaxSeg);
    std::vector<int> nseg(ntiles, 0);
    double seg_total = 0, rec_total = 0;
    int seg_max = 0, clipped = 0;
    for (int t = 0; t < ntiles; t++) {
        const uint32_t c_lo = cell_of_slot[(size_t)t * kTile], c_hi = cell_of_slot[(size_t)t * kTile + kTile - 1];
        int ns = 0;
        uint32_t recs = 0;
        for (uint32_t row = c_lo / nc; row <= c_hi / nc; row++) {  // x-rows the tile touches
            const int cx0 = row == c_lo / nc ? (int)(c_lo % nc) : 0, cx1 = row == c_hi / nc ? (int)(c_hi % nc) : nc - 1;
            const int cy = (int)(row % nc), cz = (int)(row / nc);
            for (int dz = -1; dz <= 1; dz++)
                for (int dy = -1; dy <= 1; dy++) {
                    const int oy = (cy + dy + nc) % nc, oz = (cz + dz + nc) % nc;
                    const uint32_t base = ((uint32_t)oz * nc + oy) * nc;
                    auto add = [&](int a, int b) {  // cells [a, b] of that row
                        if (ns < kMaxSeg && b >= a) {
                            Seg s{start[base + a], start[base + b + 1] - start[base + a]};
                            segs[(size_t)t * kMaxSeg + ns++] = s;
                            recs += s.count;
                        }
                    };
                    add(std::max(cx0 - 1, 0), std::min(cx1 + 1, nc - 1));
                    if (cx0 == 0) add(nc - 1, nc - 1);
                    if (cx1 == nc - 1) add(0, 0);
                }
        }
        nseg[t] = ns;
        seg_total += ns;
        rec_total += std::min<uint32_t>(recs, kStageRecords);
        seg_max = std::max(seg_max, ns);
        if (recs > kStageRecords) clipped++;
    }
    printf("tiles %d: %.1f segments per tile (max %d), %.0f staged records per tile = %.1f KB, %.0f B per particle; %d tiles clipped\n", ntiles,
           seg_total / ntiles, seg_max, rec_total / ntiles, rec_total / ntiles * 32 / 1024, rec_total / ntiles * 32 / kTile, clipped);
    double4 *d_pos;
    Seg *d_segs;
    int *d_nseg;
    double *d_sink;
    CK(cudaMalloc(&d_pos, sizeof(double4) * N));
    CK(cudaMalloc(&d_segs, sizeof(Seg) * segs.size()));
    CK(cudaMalloc(&d_nseg, sizeof(int) * ntiles));
    CK(cudaMemcpy(d_pos, sorted.data(), sizeof(double4) * N, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_segs, segs.data(), sizeof(Seg) * segs.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_nseg, nseg.data(), sizeof(int) * ntiles, cudaMemcpyHostToDevice));
    const size_t smem = sizeof(double4) * 2 * kStageRecords;
    CK(cudaFuncSetAttribute(k_stage, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0, nsm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stage, kTile, smem));
    CK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0));
    const int grid = std::min(ntiles, per_sm * nsm);
    CK(cudaMalloc(&d_sink, sizeof(double) * grid * kTile));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        for (int rep = 0; rep < 6; rep++) {
            CK(cudaEventRecord(e0));
            k_stage<<<grid, kTile, smem>>>(ntiles, d_pos, d_segs, d_nseg, d_sink, mode);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (rep > 0) best = std::min(best, ms);
        }
        CK(cudaGetLastError());
        const double bytes = mode == 0 ? rec_total * 32.0 : (double)N * 32.0;
        printf("STAGE PROBE mode %d (%s): grid %d x %d threads (%d CTAs/SM), %.4f ms at N = 2^%d -> %.3f ms scaled to 2^24; %.2f TB/s into %s\n", mode,
               mode == 0 ? "TMA bulk staging of the 9 neighbour rows, double-buffered" : "own records only (plain coalesced loads)", grid, kTile,
               per_sm, best, lg, best * (double)(1ll << 24) / (double)N, bytes / (best * 1e-3) / 1e12, mode == 0 ? "shared memory" : "registers");
    }
    return 0;
}
