// K2 (mode EXACT) -- the reference's own float32 additions in the reference's own order.
// Follows /root/reference/mmseg/datasets/dsec.py:47-58: eight corner passes (x outer, y, t inner), each
// one `voxel_grid.put_(index[mask], weights[mask], accumulate=True)`, which the deterministic
// single-thread reference executes as sequential float32 additions in event order.  So the value of a
// voxel is the left-to-right float32 sum of its contributions ordered by (corner pass, event index).
//
// A verification mode, not a fast one: every (event, corner) pair is written out as
// (voxel * 8 + pass, weight), a STABLE radix sort groups them by key while keeping the event order inside
// a key, and one thread per voxel adds its run sequentially.  The result is BIT-IDENTICAL to the
// reference's raw grid, including the rounding residue it leaves where ON and OFF events cancel.
// The sort is cub::DeviceRadixSort (a library sort is fine here: this mode exists to pin the other modes
// to the reference's bits, it is not on the hot path).  Windows are processed one at a time; workspace is
// 64 bytes per event-corner pair of the largest window.
#include <cub/device/device_radix_sort.cuh>

#include "event_math.cuh"

namespace cmda {

constexpr unsigned kExactSentinel = 0xffffffffu;
constexpr int kExactThreads = 256;

template <bool RAW>
__global__ void __launch_bounds__(kExactThreads)
exact_emit_kernel(const uint32_t* __restrict__ t, const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                  const uint8_t* __restrict__ p, const float* __restrict__ ft, const float* __restrict__ fx,
                  const float* __restrict__ fy, const float* __restrict__ fp, long long start, long long end,
                  const float2* __restrict__ map, int H, int W, int B, unsigned* __restrict__ keys, float* __restrict__ vals,
                  unsigned long long* __restrict__ bin_counts) {
    extern __shared__ unsigned int s_bins[];
    const long long n = end - start;
    if (bin_counts != nullptr) {
        for (int b = threadIdx.x; b < B; b += blockDim.x) s_bins[b] = 0u;
        __syncthreads();
    }
    RawWindowTime rw{};
    F32WindowTime fw{};
    if (RAW) rw = raw_window_time(t, start, end, B);
    else fw = f32_window_time(ft + start, n, B);
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long gi = start + i;
        Event e;
        bool ok = true;
        if (RAW) e = make_raw_event(__ldg(t + gi), __ldg(x + gi), __ldg(y + gi), __ldg(p + gi), map, H, W, rw, ok);
        else e = make_f32_event(__ldg(ft + gi), __ldg(fx + gi), __ldg(fy + gi), __ldg(fp + gi), fw);
        const Origin o = origin_of(e, H, W, B);
        if (bin_counts != nullptr && ok && o.t0 >= 0 && o.t0 < B) atomicAdd(&s_bins[o.t0], 1u);
        unsigned k8[8];
        float w8[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) { k8[c] = kExactSentinel; w8[c] = 0.0f; }
        if (ok && o.any) {
            for_each_corner(e, o, H, W, B, [&](int xl, int yl, int tl, float w) {
                const int pass = ((xl - o.x0) << 2) | ((yl - o.y0) << 1) | (tl - o.t0);       // dsec.py:47-49 nest order
                const unsigned voxel = (static_cast<unsigned>(tl) * H + yl) * W + xl;         // dsec.py:54-56
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    if (c == pass) { k8[c] = voxel * 8u + static_cast<unsigned>(pass); w8[c] = w; }
            });
        }
        unsigned* ko = keys + i * 8;
        float* vo = vals + i * 8;
        reinterpret_cast<uint4*>(ko)[0] = make_uint4(k8[0], k8[1], k8[2], k8[3]);
        reinterpret_cast<uint4*>(ko)[1] = make_uint4(k8[4], k8[5], k8[6], k8[7]);
        reinterpret_cast<float4*>(vo)[0] = make_float4(w8[0], w8[1], w8[2], w8[3]);
        reinterpret_cast<float4*>(vo)[1] = make_float4(w8[4], w8[5], w8[6], w8[7]);
    }
    if (bin_counts != nullptr) {
        __syncthreads();
        for (int b = threadIdx.x; b < B; b += blockDim.x)
            if (s_bins[b]) atomicAdd(bin_counts + b, static_cast<unsigned long long>(s_bins[b]));
    }
}

// one thread per voxel: its contributions are the sorted run with keys in [voxel * 8, voxel * 8 + 8)
__global__ void __launch_bounds__(kExactThreads)
exact_reduce_kernel(const unsigned* __restrict__ keys, const float* __restrict__ vals, long long n_items, unsigned n_voxels,
                    float* __restrict__ grid) {
    const unsigned v = blockIdx.x * blockDim.x + threadIdx.x;
    if (v >= n_voxels) return;
    const unsigned lo_key = v * 8u;
    long long lo = 0, hi = n_items;
    while (lo < hi) {                                  // first item with key >= lo_key
        const long long mid = lo + ((hi - lo) >> 1);
        if (__ldg(keys + mid) < lo_key) lo = mid + 1;
        else hi = mid;
    }
    float acc = 0.0f;                                  // torch.zeros (dsec.py:31)
    for (long long i = lo; i < n_items && (__ldg(keys + i) >> 3) == v; ++i) acc = __fadd_rn(acc, __ldg(vals + i));
    grid[v] = acc;
}

static size_t exact_sort_temp_bound(long long items) {
    // closed-form bound on cub's temporary storage: histograms + per-tile look-back state, and -- SortPairs leaves
    // its inputs intact, so it ping-pongs through a key and a value array of its own -- 8 bytes per item
    return align_up(static_cast<size_t>(64) << 20, 256) + align_up(static_cast<size_t>(items) * 8 + 4096, 256);
}

int exact_supported(int H, int W, int B) {
    return static_cast<long long>(B) * H * W * 8 < 0xffffffffLL;
}

size_t exact_workspace_bytes(long long max_window_events) {
    const long long items = max_window_events * 8;
    return 4 * align_up(sizeof(unsigned) * static_cast<size_t>(items), 256) + exact_sort_temp_bound(items);
}

// One window: [start, end) of the raw store (RAW) or of the float arrays (!RAW) -> grid [B, H, W].
template <bool RAW>
static int exact_window(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, const float* ft,
                        const float* fx, const float* fy, const float* fp, long long start, long long end, const float2* map,
                        int H, int W, int B, float* grid, int64_t* bin_counts, void* ws, size_t ws_bytes, cudaStream_t st) {
    const long long n = end - start;
    const unsigned n_voxels = static_cast<unsigned>(B) * H * W;
    if (n <= 0) {
        CMDA_CUDA_TRY(cudaMemsetAsync(grid, 0, sizeof(float) * n_voxels, st));
        return CMDA_OK;
    }
    const long long items = n * 8;
    if (items >= (1LL << 31)) return CMDA_ERR_UNSUPPORTED;
    const size_t arr = align_up(sizeof(unsigned) * static_cast<size_t>(items), 256);
    if (ws_bytes < 4 * arr) return CMDA_ERR_WORKSPACE;
    char* base = static_cast<char*>(ws);
    unsigned* k_in = reinterpret_cast<unsigned*>(base);
    unsigned* k_out = reinterpret_cast<unsigned*>(base + arr);
    float* v_in = reinterpret_cast<float*>(base + 2 * arr);
    float* v_out = reinterpret_cast<float*>(base + 3 * arr);
    void* temp = base + 4 * arr;
    size_t temp_have = ws_bytes - 4 * arr;
    long long blocks = (n + kExactThreads - 1) / kExactThreads;
    if (blocks > 148 * 16) blocks = 148 * 16;
    exact_emit_kernel<RAW><<<static_cast<unsigned>(blocks), kExactThreads, sizeof(unsigned) * B, st>>>(
        t, x, y, p, ft, fx, fy, fp, start, end, map, H, W, B, k_in, v_in, reinterpret_cast<unsigned long long*>(bin_counts));
    CMDA_LAUNCH_CHECK();
    int end_bit = 1;
    while (end_bit < 32 && (1ULL << end_bit) <= static_cast<unsigned long long>(n_voxels) * 8ULL) ++end_bit;
    end_bit = 32;    // the sentinel (all ones) must sort last: use every bit
    size_t temp_need = 0;
    CMDA_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, temp_need, k_in, k_out, v_in, v_out, static_cast<int>(items), 0,
                                                  end_bit, st));
    if (temp_need > temp_have) return CMDA_ERR_WORKSPACE;
    CMDA_CUDA_TRY(cub::DeviceRadixSort::SortPairs(temp, temp_need, k_in, k_out, v_in, v_out, static_cast<int>(items), 0,
                                                  end_bit, st));
    exact_reduce_kernel<<<(n_voxels + kExactThreads - 1) / kExactThreads, kExactThreads, 0, st>>>(k_out, v_out, items, n_voxels,
                                                                                                grid);
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

int launch_exact_raw(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, const WindowTable& tab, int S,
                     const float* maps, int H, int W, int B, float* raw, int64_t* bin_counts, void* ws, size_t ws_bytes,
                     cudaStream_t st) {
    if (!exact_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    const size_t V = static_cast<size_t>(B) * H * W;
    for (int s = 0; s < S; ++s) {
        const float2* map = maps ? reinterpret_cast<const float2*>(maps) + static_cast<size_t>(tab.w[s].map_id) * H * W : nullptr;
        const int rc = exact_window<true>(t, x, y, p, nullptr, nullptr, nullptr, nullptr, tab.w[s].start, tab.w[s].end, map, H, W,
                                          B, raw + s * V, bin_counts ? bin_counts + static_cast<size_t>(s) * B : nullptr, ws,
                                          ws_bytes, st);
        if (rc != CMDA_OK) return rc;
    }
    return CMDA_OK;
}

int launch_exact_f32(const float* ft, const float* fx, const float* fy, const float* fp, long long n, int H, int W, int B,
                     float* grid, int64_t* bin_counts, void* ws, size_t ws_bytes, cudaStream_t st) {
    if (!exact_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    return exact_window<false>(nullptr, nullptr, nullptr, nullptr, ft, fx, fy, fp, 0, n, nullptr, H, W, B, grid, bin_counts, ws,
                               ws_bytes, st);
}

}  // namespace cmda
