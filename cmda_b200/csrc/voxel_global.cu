// K2 (mode GLOBAL) -- trilinear voxel scatter with one 64-bit integer RED per corner into
// an L2-resident fixed-point grid, plus the integer side outputs used by the parity tests.
// Follows /root/reference/mmseg/datasets/dsec.py:26-58 and 341-357.
//
// Determinism: each float32 corner weight (bit-identical to the reference's) is quantised
// to 2^-30 and accumulated as a 64-bit integer.  Integer addition is associative, so the
// grid is bit-for-bit reproducible whatever the atomic order; the final value is the
// correctly rounded float32 of the exact sum of the quantised weights.
//
// This is the simple path: 8 REDG.64 per event bound it by L2 atomic throughput.  It is
// used for small windows and as the in-library cross-check of the TILED mode.
#include "event_math.cuh"

namespace cmda {

constexpr int kScatterThreads = 256;

template <bool RAW>
__global__ void __launch_bounds__(kScatterThreads)
voxel_scatter_global_kernel(const uint32_t* __restrict__ t, const uint16_t* __restrict__ x,
                            const uint16_t* __restrict__ y, const uint8_t* __restrict__ p,
                            const float* __restrict__ ft, const float* __restrict__ fx,
                            const float* __restrict__ fy, const float* __restrict__ fp, WindowTable tab,
                            const float2* __restrict__ maps, int H, int W, int B,
                            unsigned long long* __restrict__ acc, unsigned long long* __restrict__ bin_counts) {
    extern __shared__ unsigned int s_bins[];   // B counters (only when bin_counts != nullptr)
    const int s = blockIdx.y;
    const WindowDesc wd = tab.w[s];
    const long long n = wd.end - wd.start;
    if (n <= 0) return;
    if (bin_counts != nullptr) {
        for (int b = threadIdx.x; b < B; b += blockDim.x) s_bins[b] = 0u;
        __syncthreads();
    }
    const size_t V = static_cast<size_t>(B) * H * W;
    unsigned long long* g = acc + static_cast<size_t>(s) * V;
    const float2* map = maps ? maps + static_cast<size_t>(wd.map_id) * H * W : nullptr;

    RawWindowTime rw{};
    F32WindowTime fw{};
    if (RAW) rw = raw_window_time(t, wd.start, wd.end, B);
    else fw = f32_window_time(ft + wd.start, n, B);

    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long gi = wd.start + i;
        Event e;
        bool ok = true;
        if (RAW) e = make_raw_event(__ldg(t + gi), __ldg(x + gi), __ldg(y + gi), __ldg(p + gi), map, H, W, rw, ok);
        else e = make_f32_event(__ldg(ft + gi), __ldg(fx + gi), __ldg(fy + gi), __ldg(fp + gi), fw);
        const Origin o = origin_of(e, H, W, B);
        if (bin_counts != nullptr && ok && o.t0 >= 0 && o.t0 < B) atomicAdd(&s_bins[o.t0], 1u);
        if (!ok || !o.any) continue;
        for_each_corner(e, o, H, W, B, [&](int xl, int yl, int tl, float w) {
            const long long q = quantise(w);
            if (q != 0)
                atomicAdd(g + (static_cast<size_t>(tl) * H + yl) * W + xl, static_cast<unsigned long long>(q));
        });
    }
    if (bin_counts != nullptr) {
        __syncthreads();
        for (int b = threadIdx.x; b < B; b += blockDim.x)
            if (s_bins[b]) atomicAdd(bin_counts + static_cast<size_t>(s) * B + b, static_cast<unsigned long long>(s_bins[b]));
    }
}

// Integer side outputs of one window (cmda_remap_events).
__global__ void remap_events_kernel(const uint32_t* __restrict__ t, const uint16_t* __restrict__ x,
                                    const uint16_t* __restrict__ y, const uint8_t* __restrict__ p, long long start,
                                    long long end, const float2* __restrict__ map, int H, int W, int B,
                                    float* __restrict__ xr, float* __restrict__ yr, float* __restrict__ tn,
                                    int* __restrict__ x0, int* __restrict__ y0, int* __restrict__ t0) {
    const long long n = end - start;
    const RawWindowTime rw = raw_window_time(t, start, end, B);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long gi = start + i;
        bool ok;
        const Event e = make_raw_event(__ldg(t + gi), __ldg(x + gi), __ldg(y + gi), __ldg(p + gi), map, H, W, rw, ok);
        const Origin o = origin_of(e, H, W, B);
        if (xr) xr[i] = e.x;
        if (yr) yr[i] = e.y;
        if (tn) tn[i] = e.tn;
        if (x0) x0[i] = o.x0;
        if (y0) y0[i] = o.y0;
        if (t0) t0[i] = o.t0;
    }
}

static int scatter_grid_x(long long max_events) {
    long long b = (max_events + kScatterThreads * 4 - 1) / (kScatterThreads * 4);
    if (b < 1) b = 1;
    if (b > 148 * 8) b = 148 * 8;
    return static_cast<int>(b);
}

int launch_scatter_global_raw(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                              const WindowTable& tab, int S, long long max_events, const float* maps, int H, int W,
                              int B, long long* acc, int64_t* bin_counts, cudaStream_t s) {
    if (max_events <= 0) return CMDA_OK;
    dim3 grid(scatter_grid_x(max_events), S);
    voxel_scatter_global_kernel<true><<<grid, kScatterThreads, bin_counts ? sizeof(unsigned) * B : 0, s>>>(
        t, x, y, p, nullptr, nullptr, nullptr, nullptr, tab, reinterpret_cast<const float2*>(maps), H, W, B,
        reinterpret_cast<unsigned long long*>(acc), reinterpret_cast<unsigned long long*>(bin_counts));
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

int launch_scatter_global_f32(const float* ft, const float* fx, const float* fy, const float* fp, long long n, int H,
                              int W, int B, long long* acc, int64_t* bin_counts, cudaStream_t s) {
    if (n <= 0) return CMDA_OK;
    WindowTable tab{};
    tab.w[0].start = 0;
    tab.w[0].end = n;
    dim3 grid(scatter_grid_x(n), 1);
    voxel_scatter_global_kernel<false><<<grid, kScatterThreads, bin_counts ? sizeof(unsigned) * B : 0, s>>>(
        nullptr, nullptr, nullptr, nullptr, ft, fx, fy, fp, tab, nullptr, H, W, B,
        reinterpret_cast<unsigned long long*>(acc), reinterpret_cast<unsigned long long*>(bin_counts));
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

int launch_remap(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, long long start,
                 long long end, const float* map, int H, int W, int B, float* xr, float* yr, float* tn, int* x0,
                 int* y0, int* t0, cudaStream_t s) {
    if (end <= start) return CMDA_OK;
    remap_events_kernel<<<scatter_grid_x(end - start), kScatterThreads, 0, s>>>(
        t, x, y, p, start, end, reinterpret_cast<const float2*>(map), H, W, B, xr, yr, tn, x0, y0, t0);
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

}  // namespace cmda
