// Micro-benchmark: throughput of shared-memory atomics across a thread-block cluster (DSMEM) on B200.
// Each cluster owns one "plane" of CS*SLICE 32-bit words spread over its CTAs' shared memory; every
// thread scatters random adds to the whole plane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o dsmem_atomics dsmem_atomics.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16; return h;
}

constexpr int SLICE = 38400;   // words per CTA = 60 rows x 640 px

template <int MODE>   // 0: red (no return)  1: atom with return  2: plain remote store (st.shared::cluster)
__global__ void dsmem_kernel(int iters, unsigned* sink) {
    extern __shared__ unsigned s[];
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned cs = cluster.num_blocks();
    for (int i = threadIdx.x; i < SLICE; i += blockDim.x) s[i] = 0;
    cluster.sync();
    uint32_t h = mix(blockIdx.x * 1315423911u + threadIdx.x);
    unsigned acc = 0;
    const unsigned total = cs * SLICE;
    for (int it = 0; it < iters; ++it) {
        h = h * 1664525u + 1013904223u;
        const unsigned p = mix(h) % total;
        const unsigned r = p / SLICE, off = p - r * SLICE;
        unsigned* dst = cluster.map_shared_rank(s, r) + off;
        if (MODE == 0) atomicAdd(dst, h | 1u);
        else if (MODE == 1) { unsigned o = atomicAdd(dst, h | 1u); acc += (o + (h | 1u) < o); }
        else if (MODE == 2) *dst = h;
        else {
            const unsigned local = static_cast<unsigned>(__cvta_generic_to_shared(s + off));
            unsigned remote;
            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
            if (MODE == 3) asm volatile("red.relaxed.cluster.shared::cluster.add.u32 [%0], %1;" ::"r"(remote), "r"(h | 1u) : "memory");
            else { unsigned o; asm volatile("atom.relaxed.cluster.shared::cluster.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(remote), "r"(h | 1u) : "memory"); acc += (o + (h | 1u) < o); }
        }
    }
    cluster.sync();
    unsigned v = acc;
    for (int i = threadIdx.x; i < SLICE; i += blockDim.x) v ^= s[i];
    if (v == 0x12345678u) sink[0] = v;
}

template <typename K>
int run(K kern, const char* name, int cs, int threads, int iters, unsigned* sink) {
    const size_t shm = SLICE * 4;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm));
    if (cs > 8) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    int nclusters = 148 / cs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nclusters * cs); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = shm;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxc = 0;
    cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    CK(cudaLaunchKernelEx(&cfg, kern, iters, sink)); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(a); CK(cudaLaunchKernelEx(&cfg, kern, iters, sink)); cudaEventRecord(b); CK(cudaEventSynchronize(b));
        float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms;
    }
    double ops = (double)nclusters * cs * threads * iters;
    printf("%-12s cluster=%2d ctas=%3d (max active clusters %d) thr=%4d : %8.3f ms  %8.1f Gop/s  %.2f op/clk/SM\n", name, cs,
           nclusters * cs, maxc, threads, best, ops / best / 1e6, ops / best / 1e6 / (nclusters * cs) / 1.9);
    return 0;
}

int main() {
    unsigned* sink; CK(cudaMalloc(&sink, 64));
    for (int cs : {1, 2, 4, 8, 16})
        for (int threads : {512, 1024}) {
            if (run(dsmem_kernel<0>, "red", cs, threads, 2048, sink)) return 1;
            if (run(dsmem_kernel<1>, "atom.ret", cs, threads, 2048, sink)) return 1;
            if (run(dsmem_kernel<2>, "st", cs, threads, 2048, sink)) return 1;
            if (run(dsmem_kernel<3>, "red.ptx", cs, threads, 2048, sink)) return 1;
            if (run(dsmem_kernel<4>, "atom.ret.ptx", cs, threads, 2048, sink)) return 1;
        }
    return 0;
}
