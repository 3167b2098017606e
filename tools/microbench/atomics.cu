// Micro-benchmarks that size the voxel scatter design on B200 (not part of the product):
// throughput of the candidate accumulation primitives at full-chip occupancy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomics atomics.cu && ./atomics
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s @%d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352dU; h ^= h >> 15; h *= 0x846ca68bU; h ^= h >> 16; return h;
}

// mode 0: ATOMS.ADD u32 no return; 1: with return; 2: LDS+STS (non atomic); 3: ATOMS pair adjacent (2 per idx)
template <int MODE>
__global__ void smem_kernel(int iters, int words, unsigned* sink) {
    extern __shared__ unsigned s[];
    for (int i = threadIdx.x; i < words; i += blockDim.x) s[i] = 0;
    __syncthreads();
    uint32_t h = mix(blockIdx.x * 1315423911u + threadIdx.x);
    unsigned acc = 0;
    for (int it = 0; it < iters; ++it) {
        h = h * 1664525u + 1013904223u;
        uint32_t a = (mix(h) % (uint32_t)words);
        if (MODE == 0) atomicAdd(&s[a], h | 1);
        else if (MODE == 1) { unsigned o = atomicAdd(&s[a], h | 1); acc += (o + (h | 1) < o); }
        else if (MODE == 2) { s[a] += h; }
        else { atomicAdd(&s[a], h | 1); atomicAdd(&s[(a + 1) % words], h | 3); }
    }
    __syncthreads();
    unsigned v = acc;
    for (int i = threadIdx.x; i < words; i += blockDim.x) v ^= s[i];
    if (v == 0x12345678u) sink[0] = v;
}

// global RED: mode 0 u64, 1 f32, 2 u32; region words
template <int MODE>
__global__ void red_kernel(int iters, size_t elems, void* grid) {
    uint32_t h = mix(blockIdx.x * 1315423911u + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
        h = h * 1664525u + 1013904223u;
        size_t a = (size_t)(mix(h)) % elems;
        if (MODE == 0) atomicAdd((unsigned long long*)grid + a, (unsigned long long)h);
        else if (MODE == 1) atomicAdd((float*)grid + a, 1.0f);
        else atomicAdd((unsigned*)grid + a, h);
    }
}

// random float2 gather from a map (L2 resident)
__global__ void gather_kernel(int iters, size_t elems, const float2* map, float* sink) {
    uint32_t h = mix(blockIdx.x * 1315423911u + threadIdx.x);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        h = h * 1664525u + 1013904223u;
        size_t a = (size_t)(mix(h)) % elems;
        float2 m = __ldg(map + a);
        acc += m.x + m.y;
    }
    if (acc == 1.2345f) sink[0] = acc;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    printf("device %s SMs %d\n", pr.name, pr.multiProcessorCount);
    const int sms = pr.multiProcessorCount;
    unsigned* sink; CK(cudaMalloc(&sink, 64));
    // ---- shared memory
    const int iters = 4096;
    for (int words : {1024, 8192, 32768}) {
        for (int threads : {256, 512, 1024}) {
            int ctas_per_sm = (words * 4 <= 48 * 1024) ? (2048 / threads) : 1;
            if (ctas_per_sm * words * 4 > 200 * 1024) ctas_per_sm = 200 * 1024 / (words * 4);
            if (ctas_per_sm < 1) ctas_per_sm = 1;
            int grid = sms * ctas_per_sm;
            size_t shm = (size_t)words * 4;
            auto run = [&](auto kern, const char* name, int per_iter) {
                cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
                float ms = time_ms([&] { kern<<<grid, threads, shm>>>(iters, words, sink); });
                double ops = (double)grid * threads * iters * per_iter;
                printf("smem %-14s words=%6d thr=%4d cta/sm=%d : %8.3f ms  %8.1f Gop/s  (%.2f op/clk/SM @1.9GHz)\n", name, words,
                       threads, ctas_per_sm, ms, ops / ms / 1e6, ops / ms / 1e6 / sms / 1.9);
            };
            run(smem_kernel<0>, "atoms.add", 1);
            run(smem_kernel<1>, "atoms.add.ret", 1);
            run(smem_kernel<2>, "lds+sts", 1);
            run(smem_kernel<3>, "atoms.add x2", 2);
        }
    }
    // ---- global RED
    for (size_t mb : {2, 12, 40, 200}) {
        size_t bytes = mb << 20;
        void* g; CK(cudaMalloc(&g, bytes)); CK(cudaMemset(g, 0, bytes));
        int grid = sms * 8, threads = 256, it = 1024;
        auto run = [&](auto kern, const char* name, size_t esz) {
            float ms = time_ms([&] { kern<<<grid, threads>>>(it, bytes / esz, g); });
            double ops = (double)grid * threads * it;
            printf("global %-10s region=%4zu MB : %8.3f ms  %8.1f Gop/s\n", name, mb, ms, ops / ms / 1e6);
        };
        run(red_kernel<0>, "red.u64", 8);
        run(red_kernel<1>, "red.f32", 4);
        run(red_kernel<2>, "red.u32", 4);
        cudaFree(g);
    }
    // ---- gather
    {
        size_t elems = 480 * 640;
        float2* map; CK(cudaMalloc(&map, elems * 8)); CK(cudaMemset(map, 0, elems * 8));
        int grid = sms * 8, threads = 256, it = 1024;
        float ms = time_ms([&] { gather_kernel<<<grid, threads>>>(it, elems, map, (float*)sink); });
        double ops = (double)grid * threads * it;
        printf("gather float2 from 2.4MB map: %8.3f ms  %8.1f Gop/s\n", ms, ops / ms / 1e6);
    }
    return 0;
}
