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

}  // namespace cmda
