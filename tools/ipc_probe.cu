// ipc_probe.cu -- can two processes (one per GPU) map each other's device memory with cudaIpc* in this sandbox, and what
// does a flag round trip over NVLink peer memory cost?  Evidence for the peer-memory slab transport.
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a -o ipc_probe ipc_probe.cu
//   ./ipc_probe 0 /tmp/ipcp & ./ipc_probe 1 /tmp/ipcp ; wait
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unistd.h>

#include <cuda_runtime.h>

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t _e = (call);                                                                   \
        if (_e != cudaSuccess) {                                                                   \
            printf("rank %d: %s failed: %s\n", g_rank, #call, cudaGetErrorString(_e));            \
            return 2;                                                                              \
        }                                                                                          \
    } while (0)
static int g_rank = 0;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// fill the peer's payload, then publish a flag there
__global__ void k_send(double *peer_payload, int n, unsigned long long *peer_flag, unsigned long long tag, unsigned int *done)
{
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) peer_payload[i] = (double)tag + i;
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int k = atomicAdd(done, 1u);
        if (k == gridDim.x - 1) {
            *done = 0;
            __threadfence_system();
            st_release_sys(peer_flag, tag);
        }
    }
}
// bounded wait for a flag in OUR memory written by the peer; out[0] = 1 ok / 0 timed out, out[1] = payload check
__global__ void k_wait(const unsigned long long *flag, unsigned long long tag, const double *payload, int n, long long max_ns, int *out)
{
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    int ok = 0;
    for (;;) {
        if (ld_acquire_sys(flag) >= tag) { ok = 1; break; }
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
        if ((long long)(t1 - t0) > max_ns) break;
        __nanosleep(100);
    }
    out[0] = ok;
    int good = 1;
    if (ok)
        for (int i = 0; i < n; i += 97)
            if (payload[i] != (double)tag + i) good = 0;
    out[1] = good;
}
// flag ping-pong inside one kernel per GPU: rank 0 sends i, rank 1 echoes i
__global__ void k_pingpong(int rank, unsigned long long *mine, unsigned long long *peer, int iters, long long max_ns, long long *ns_out)
{
    unsigned long long t0, t1, ts;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(ts));
    for (int i = 1; i <= iters; i++) {
        if (rank == 0) st_release_sys(peer, (unsigned long long)i);
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
        while (ld_acquire_sys(mine) < (unsigned long long)i) {
            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
            if ((long long)(t1 - t0) > max_ns) { ns_out[0] = -1; return; }
        }
        if (rank == 1) st_release_sys(peer, (unsigned long long)i);
    }
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t1));
    ns_out[0] = (long long)(t1 - ts);
}

static bool read_file(const std::string &path, void *buf, size_t n, int timeout_s)
{
    for (int t = 0; t < timeout_s * 100; t++) {
        FILE *fp = fopen(path.c_str(), "rb");
        if (fp) {
            size_t got = fread(buf, 1, n, fp);
            fclose(fp);
            if (got == n) return true;
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(10));
    }
    return false;
}
static void write_file(const std::string &path, const void *buf, size_t n)
{
    std::string tmp = path + ".tmp";
    FILE *fp = fopen(tmp.c_str(), "wb");
    fwrite(buf, 1, n, fp);
    fclose(fp);
    rename(tmp.c_str(), path.c_str());
}

int main(int argc, char **argv)
{
    if (argc < 3) return 1;
    g_rank = atoi(argv[1]);
    const std::string prefix = argv[2];
    const int peer = 1 - g_rank, n = 1 << 17;  // 1 MiB payload
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (ndev < 2) { printf("rank %d: only %d device(s)\n", g_rank, ndev); return 3; }
    CK(cudaSetDevice(g_rank));
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, g_rank, peer));
    char *base = nullptr;
    const size_t bytes = sizeof(double) * n + 4096;
    CK(cudaMalloc(&base, bytes));
    CK(cudaMemset(base, 0, bytes));
    cudaIpcMemHandle_t mine, theirs;
    CK(cudaIpcGetMemHandle(&mine, base));
    write_file(prefix + "." + std::to_string(g_rank), &mine, sizeof(mine));
    if (!read_file(prefix + "." + std::to_string(peer), &theirs, sizeof(theirs), 60)) { printf("rank %d: no handle from the peer\n", g_rank); return 4; }
    char *pbase = nullptr;
    CK(cudaIpcOpenMemHandle((void **)&pbase, theirs, cudaIpcMemLazyEnablePeerAccess));
    double *payload = (double *)(base + 4096), *peer_payload = (double *)(pbase + 4096);
    unsigned long long *flag = (unsigned long long *)base, *peer_flag = (unsigned long long *)pbase;
    unsigned long long *pp_mine = flag + 16, *pp_peer = peer_flag + 16;
    unsigned int *done;
    int *out, h_out[2] = {0, 0};
    long long *ns, h_ns = 0;
    CK(cudaMalloc(&done, 4));
    CK(cudaMemset(done, 0, 4));
    CK(cudaMalloc(&out, 8));
    CK(cudaMalloc(&ns, 8));
    CK(cudaDeviceSynchronize());
    // barrier through files so both sides have mapped before anyone writes
    char one = 1, got = 0;
    write_file(prefix + ".ready." + std::to_string(g_rank), &one, 1);
    if (!read_file(prefix + ".ready." + std::to_string(peer), &got, 1, 60)) return 5;
    for (unsigned long long tag = 1; tag <= 3; tag++) {
        k_send<<<64, 256>>>(peer_payload, n, peer_flag, tag * 1000, done);
        k_wait<<<1, 1>>>(flag, tag * 1000, payload, n, 5000000000ll, out);
        CK(cudaMemcpy(h_out, out, 8, cudaMemcpyDeviceToHost));
        if (!h_out[0] || !h_out[1]) { printf("rank %d: exchange %llu FAILED (flag %d payload %d)\n", g_rank, tag, h_out[0], h_out[1]); return 6; }
    }
    const int iters = 2000;
    k_pingpong<<<1, 1>>>(g_rank, pp_mine, pp_peer, iters, 5000000000ll, ns);
    CK(cudaMemcpy(&h_ns, ns, 8, cudaMemcpyDeviceToHost));
    printf("IPC PROBE rank %d: canAccessPeer=%d, payload+flag exchange OK x3, flag round trip %.2f us (%d iterations)\n", g_rank, can,
           h_ns > 0 ? 1e-3 * (double)h_ns / iters : -1.0, iters);
    write_file(prefix + ".done." + std::to_string(g_rank), &one, 1);
    read_file(prefix + ".done." + std::to_string(peer), &got, 1, 30);
    cudaIpcCloseMemHandle(pbase);
    cudaFree(base);
    return 0;
}
