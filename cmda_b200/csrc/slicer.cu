// K1 -- event-window slicer: batched "last event with t <= ts" lookups over the
// device-resident, time-sorted t array.
// Follows /root/reference/create_dsec_dataset_txt.py:19-42 (np.searchsorted 'right'
// inside the ms_to_idx bracket).  Integer only: results are bit-exact.
#include "common.cuh"

namespace cmda {

// number of elements of t[lo, hi) that are <= q  (+ lo): classic upper bound
__device__ __forceinline__ long long upper_bound_u32(const uint32_t* __restrict__ t, long long lo, long long hi,
                                                     long long q) {
    while (lo < hi) {
        const long long mid = lo + ((hi - lo) >> 1);
        if (static_cast<long long>(__ldg(t + mid)) <= q) lo = mid + 1;
        else hi = mid;
    }
    return lo;
}

__global__ void searchsorted_right_kernel(const uint32_t* __restrict__ t, long long n,
                                          const long long* __restrict__ q, int nq, long long* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    out[i] = upper_bound_u32(t, 0, n, q[i]);
}

__global__ void images_to_events_index_kernel(const uint32_t* __restrict__ t, long long n,
                                              const long long* __restrict__ ms_to_idx, long long n_ms,
                                              long long t_offset, const long long* __restrict__ ts, int n_ts,
                                              long long* __restrict__ index, int* __restrict__ status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ts) return;
    const long long ts_us = ts[i] - t_offset;                                     // :20
    int st = 0;
    long long res = -1;
    if (!(ts_us <= 0 || ts_us > static_cast<long long>(__ldg(t + n - 1)))) {      // :21-22
        // math.floor(ts_us / 1000) on a positive value; Python's true division of an
        // int64 < 2^53 followed by floor equals integer division here
        long long ms = ts_us / 1000 - 1;                                          // :24
        if (ms < 0) ms = 0;                                                       // :25
        if (ms + 2 >= n_ms) {
            st = 2;  // the reference would raise IndexError on ms_to_idx[ms + 2]
        } else {
            const long long left = ms_to_idx[ms];                                 // :26
            long long right = ms_to_idx[ms + 2];                                  // :33
            if (right > n - 1) right = n - 1;                                     // :34-35
            if (left < 0 || left >= n || right < 0) {
                st = 2;  // t[left] / t[right] would raise IndexError in the reference (:37)
            } else if (!(static_cast<long long>(__ldg(t + left)) <= ts_us &&
                         ts_us <= static_cast<long long>(__ldg(t + right)))) {    // :37-39
                st = 1;
            } else {
                res = upper_bound_u32(t, left, right + 1, ts_us) - 1;             // :40-42
            }
        }
    }
    index[i] = res;
    status[i] = st;
}

int launch_searchsorted(const uint32_t* t, int64_t n, const int64_t* q, int nq, int64_t* out, cudaStream_t s) {
    if (nq == 0) return CMDA_OK;
    const int threads = 128;
    searchsorted_right_kernel<<<(nq + threads - 1) / threads, threads, 0, s>>>(
        t, static_cast<long long>(n), reinterpret_cast<const long long*>(q), nq, reinterpret_cast<long long*>(out));
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

int launch_images_to_events_index(const uint32_t* t, int64_t n, const int64_t* ms_to_idx, int64_t n_ms,
                                  int64_t t_offset, const int64_t* ts, int n_ts, int64_t* index, int32_t* status,
                                  cudaStream_t s) {
    if (n_ts == 0) return CMDA_OK;
    const int threads = 128;
    images_to_events_index_kernel<<<(n_ts + threads - 1) / threads, threads, 0, s>>>(
        t, static_cast<long long>(n), reinterpret_cast<const long long*>(ms_to_idx), static_cast<long long>(n_ms),
        static_cast<long long>(t_offset), reinterpret_cast<const long long*>(ts), n_ts,
        reinterpret_cast<long long*>(index), status);
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

// ------------------------------------------------------------------ packed (P4) event stream
// cmda_pack_events_p4: SoA events (DSEC dtypes, t ascending) -> one 32-bit record per event
//   x | y << 11 | p << 21 | (t - t_base - 1000 * ms) << 22,   ms = (t - t_base) / 1000
// plus the table ms_to_idx[k] = index of the first event with t - t_base >= 1000 * k, k = 0 .. n_ms (the definition
// of DSEC's own ms_to_idx, create_dsec_dataset_txt.py:26-35).  d_status counts the events the format cannot hold
// (x > 2047, y > 1023, polarity beyond {0, 1}, t < t_base, t beyond the table, t descending): 0 = the packed
// stream reproduces the SoA stream exactly.
__global__ void __launch_bounds__(256)
p4_ms_table_kernel(const uint32_t* __restrict__ t, long long n, uint32_t t_base, long long n_ms, long long* __restrict__ ms_to_idx) {
    const long long k = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (k > n_ms) return;
    if (k == n_ms) { ms_to_idx[k] = n; return; }
    // first event with t >= t_base + 1000 k  (64-bit: the threshold may pass 2^32)
    const unsigned long long q = static_cast<unsigned long long>(t_base) + 1000ull * static_cast<unsigned long long>(k);
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = lo + ((hi - lo) >> 1);
        if (static_cast<unsigned long long>(__ldg(t + mid)) < q) lo = mid + 1; else hi = mid;
    }
    ms_to_idx[k] = lo;
}
__global__ void __launch_bounds__(256)
p4_pack_kernel(const uint32_t* __restrict__ t, const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
               const uint8_t* __restrict__ p, long long n, uint32_t t_base, long long n_ms, uint32_t* __restrict__ rec,
               int* __restrict__ status) {
    int bad = 0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const uint32_t ti = __ldg(t + i);
        const unsigned xi = __ldg(x + i), yi = __ldg(y + i), pi = __ldg(p + i);
        const uint32_t rel = ti - t_base;
        const uint32_t ms = rel / 1000u;
        bad += (xi > 2047u) | (yi > 1023u) | (pi > 1u) | (ti < t_base) | (static_cast<long long>(ms) >= n_ms) |
               (i > 0 && __ldg(t + i - 1) > ti);
        rec[i] = (xi & 2047u) | ((yi & 1023u) << 11) | ((pi & 1u) << 21) | ((rel - ms * 1000u) << 22);
    }
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(status, bad);
}

int launch_pack_p4(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, int64_t n, uint32_t t_base,
                   int64_t n_ms, uint32_t* rec, int64_t* ms_to_idx, int32_t* status, cudaStream_t s) {
    CMDA_CUDA_TRY(cudaMemsetAsync(status, 0, sizeof(int32_t), s));
    p4_ms_table_kernel<<<static_cast<unsigned>((n_ms + 1 + 255) / 256), 256, 0, s>>>(t, n, t_base, n_ms,
                                                                                    reinterpret_cast<long long*>(ms_to_idx));
    if (n > 0) {
        long long blocks = (n + 256 * 8 - 1) / (256 * 8);
        if (blocks > 148 * 8) blocks = 148 * 8;
        p4_pack_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(t, x, y, p, n, t_base, n_ms, rec, status);
    }
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

}  // namespace cmda
