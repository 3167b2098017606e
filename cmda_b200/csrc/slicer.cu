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

// ---- the 3-byte WIRE form of the packed stream ("P3") -> P4 records --------------------------------------
// record = x | y << 10 | p << 19 | d << 20 (x < 1024, y < 512), d = (t_us - t_base) mod 16; the 16-microsecond bucket
// j of an event follows from its index through sub_to_idx (sub_to_idx[j] = index of the first event with
// t_us - t_base >= 16 j), exactly as the millisecond bucket of a P4 record follows from ms_to_idx.  One warp per
// bucket: its events are contiguous, no search.  t_base is a multiple of 1000, so (16 j + d) mod 1000 is the P4
// record's sub-millisecond field.
__global__ void __launch_bounds__(256)
p3_unpack_kernel(const uint8_t* __restrict__ rec3, const long long* __restrict__ sub_to_idx, long long j_lo, long long j_hi,
                 long long first, long long last, uint32_t* __restrict__ rec4) {
    const long long warps = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (long long j = j_lo + ((static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5); j <= j_hi; j += warps) {
        long long lo = __ldg(sub_to_idx + j), hi = __ldg(sub_to_idx + j + 1);
        lo = lo > first ? lo : first;
        hi = hi < last ? hi : last;
        const uint32_t ms_off = static_cast<uint32_t>((static_cast<unsigned long long>(j) * 16ull) % 1000ull);
        for (long long i = lo + lane; i < hi; i += 32) {
            const uint8_t* r = rec3 + 3 * (i - first);
            const uint32_t v = static_cast<uint32_t>(__ldg(r)) | (static_cast<uint32_t>(__ldg(r + 1)) << 8) |
                               (static_cast<uint32_t>(__ldg(r + 2)) << 16);
            uint32_t sub = ms_off + (v >> 20);                    // < 999 + 16
            sub = sub >= 1000u ? sub - 1000u : sub;
            rec4[i - first] = (v & 1023u) | (((v >> 10) & 511u) << 11) | (((v >> 19) & 1u) << 21) | (sub << 22);
        }
    }
}

int launch_unpack_p3(const uint8_t* rec3, const int64_t* sub_to_idx, int64_t j_lo, int64_t j_hi, int64_t first, int64_t last,
                     uint32_t* rec4, cudaStream_t s) {
    if (last <= first || j_hi < j_lo) return CMDA_OK;
    long long blocks = (j_hi - j_lo + 1 + 7) / 8;                  // 8 warps per block, one bucket per warp and trip
    if (blocks > 148 * 16) blocks = 148 * 16;
    p3_unpack_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(rec3, reinterpret_cast<const long long*>(sub_to_idx), j_lo, j_hi,
                                                                   first, last, rec4);
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

}  // namespace cmda
