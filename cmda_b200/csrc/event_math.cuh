// Per-event arithmetic shared by every voxel kernel: window-relative time, rectify-map
// gather, truncation, tent weights, fixed-point quantisation.  Every float operation is
// an explicit round-to-nearest intrinsic so that nvcc can neither contract into FMA nor
// reassociate: the per-contribution float32 weights are bit-identical to the reference's
// (mmseg/datasets/dsec.py:38-52, 347-355).
#pragma once
#include "common.cuh"

namespace cmda {

// What one event contributes, before the corner expansion.
struct Event {
    float x, y;     // rectified coordinates (dsec.py:351-355)
    float tn;       // t_norm in [0, B-1] or NaN (dsec.py:38-39)
    float value;    // 2*pol - 1 (dsec.py:45)
};

// Window constants of the raw DSEC path (dsec.py:347-348 followed by dsec.py:38-39).
struct RawWindowTime {
    uint32_t t_first;
    float fdT;        // float32(t[-1] - t[0])
    float t01_first;  // (t - t[0])[0] / fdT  = 0 or NaN
    float den;        // t01[-1] - t01[0]     = 1 or NaN
    float cm1;        // C - 1
};

__device__ __forceinline__ RawWindowTime raw_window_time(const uint32_t* __restrict__ t, long long start,
                                                         long long end, int B) {
    RawWindowTime w;
    w.t_first = __ldg(t + start);
    w.fdT = __uint2float_rn(__ldg(t + end - 1) - w.t_first);       // uint32 subtraction, then astype(float32)
    w.t01_first = __fdiv_rn(0.0f, w.fdT);                          // events_t[0] after the division
    const float t01_last = __fdiv_rn(w.fdT, w.fdT);                // events_t[-1] after the division
    w.den = __fsub_rn(t01_last, w.t01_first);
    w.cm1 = static_cast<float>(B - 1);
    return w;
}

// t_norm of an event dt = t - t[0] microseconds into its window (uint32 subtraction first).
__device__ __forceinline__ float raw_t_norm_dt(uint32_t dt, const RawWindowTime& w) {
    const float t01 = __fdiv_rn(__uint2float_rn(dt), w.fdT);              // dsec.py:347-348
    const float a = __fsub_rn(t01, w.t01_first);                          // dsec.py:39
    const float b = __fmul_rn(w.cm1, a);
    return __fdiv_rn(b, w.den);
}
__device__ __forceinline__ float raw_t_norm(uint32_t t, const RawWindowTime& w) {
    return raw_t_norm_dt(t - w.t_first, w);
}

__device__ __forceinline__ Event make_raw_event(uint32_t t, unsigned x, unsigned y, unsigned p,
                                                const float2* __restrict__ map, int H, int W,
                                                const RawWindowTime& w, bool& in_map) {
    Event e;
    in_map = (x < static_cast<unsigned>(W)) && (y < static_cast<unsigned>(H));
    if (map != nullptr) {
        // numpy would raise IndexError for an out-of-range (y, x); such events are dropped
        const float2 m = in_map ? __ldg(map + static_cast<size_t>(y) * W + x) : make_float2(-2.0f, -2.0f);
        e.x = m.x; e.y = m.y;                                             // dsec.py:351-353
    } else {
        e.x = static_cast<float>(x); e.y = static_cast<float>(y);
        in_map = true;
    }
    e.tn = raw_t_norm(t, w);
    e.value = __fsub_rn(__fmul_rn(2.0f, static_cast<float>(p)), 1.0f);    // dsec.py:349, 45
    return e;
}

// ---- packed (P4) event source -----------------------------------------------------------------------
// One 32-bit record per event: x | y << 11 | p << 21 | sub << 22 with sub = t_us - t_base - 1000 * ms, ms the
// millisecond bucket of the event, which is not stored: it follows from the event's INDEX through the store's
// ms_to_idx table (DSEC's own events.h5 carries that table; create_dsec_dataset_txt.py:16-35 uses it the same
// way).  A CTA works on a run of consecutive events, so it looks its first event's bucket up once and keeps the
// next bucket boundaries in shared memory; a thread then finds the bucket of an event by comparing its index
// with at most a few boundaries (one, for any stream denser than a few events per millisecond).
constexpr int kMsBounds = 32;
struct MsWindow {
    long long bound[kMsBounds];   // global index of the first event of bucket ms0 + 1 + j (LLONG_MAX past the table)
    int ms0;                      // bucket of the run's first event
};
constexpr unsigned kP4XMask = 0x7ffu, kP4YMask = 0x3ffu;
constexpr int kP4YShift = 11, kP4PShift = 21, kP4SubShift = 22;

// largest k in [lo, hi] with ms_to_idx[k] <= idx (lo qualifies by construction)
__device__ __forceinline__ int ms_search(const long long* __restrict__ ms_to_idx, int lo, int hi, long long idx) {
    while (lo < hi) {
        const int mid = lo + (hi - lo + 1) / 2;
        if (__ldg(ms_to_idx + mid) <= idx) lo = mid; else hi = mid - 1;
    }
    return lo;
}
// Block-wide: every thread calls it; `first` is the global index of the run's first event (clamped into the window).
__device__ __forceinline__ void ms_window_init(MsWindow& mw, const PackedSrc& pk, const WindowDesc& wd, long long first) {
    if (threadIdx.x == 0) mw.ms0 = ms_search(pk.ms_to_idx, wd.ms_lo, wd.ms_hi, first);
    __syncthreads();
    for (int j = threadIdx.x; j < kMsBounds; j += blockDim.x) {
        const long long k = static_cast<long long>(mw.ms0) + 1 + j;
        mw.bound[j] = k <= pk.n_ms ? __ldg(pk.ms_to_idx + k) : 0x7fffffffffffffffLL;
    }
    __syncthreads();
}
// bucket of global event index idx >= the run's first event; k is the thread's cursor into mw.bound (monotone)
__device__ __forceinline__ int ms_advance(const MsWindow& mw, const PackedSrc& pk, const WindowDesc& wd, long long idx, int& k) {
    while (k < kMsBounds && mw.bound[k] <= idx) ++k;
    if (k < kMsBounds) return mw.ms0 + k;
    // a run that spans more than kMsBounds buckets (a stream of a few events per millisecond): search the table
    return ms_search(pk.ms_to_idx, mw.ms0 + kMsBounds, wd.ms_hi, idx);
}
// window-relative microseconds of a record (t_base cancels in every difference the path takes)
__device__ __forceinline__ uint32_t p4_time(uint32_t rec, int ms) {
    return static_cast<uint32_t>(ms) * 1000u + (rec >> kP4SubShift);
}
__device__ __forceinline__ RawWindowTime raw_window_time_p4(const PackedSrc& pk, const WindowDesc& wd, int B) {
    RawWindowTime w;
    w.t_first = p4_time(__ldg(pk.rec + wd.start), wd.ms_lo);
    w.fdT = __uint2float_rn(p4_time(__ldg(pk.rec + wd.end - 1), wd.ms_hi) - w.t_first);
    w.t01_first = __fdiv_rn(0.0f, w.fdT);
    const float t01_last = __fdiv_rn(w.fdT, w.fdT);
    w.den = __fsub_rn(t01_last, w.t01_first);
    w.cm1 = static_cast<float>(B - 1);
    return w;
}

// Window constants of the float path (events_to_voxel_grid called directly).
struct F32WindowTime {
    float t_first, den, cm1;
};
__device__ __forceinline__ F32WindowTime f32_window_time(const float* __restrict__ time, long long n, int B) {
    F32WindowTime w;
    w.t_first = __ldg(time);
    w.den = __fsub_rn(__ldg(time + n - 1), w.t_first);
    w.cm1 = static_cast<float>(B - 1);
    return w;
}
__device__ __forceinline__ Event make_f32_event(float t, float x, float y, float pol, const F32WindowTime& w) {
    Event e;
    e.x = x; e.y = y;
    e.tn = __fdiv_rn(__fmul_rn(w.cm1, __fsub_rn(t, w.t_first)), w.den);   // dsec.py:38-39
    e.value = __fsub_rn(__fmul_rn(2.0f, pol), 1.0f);
    return e;
}

// Corner origin of an event; `any` is false when no corner can be inside the grid.
struct Origin {
    int x0, y0, t0;
    bool any;
};
__device__ __forceinline__ Origin origin_of(const Event& e, int H, int W, int B) {
    Origin o;
    o.x0 = trunc_like_x86(e.x);                                           // dsec.py:41
    o.y0 = trunc_like_x86(e.y);                                           // dsec.py:42
    o.t0 = trunc_like_x86(e.tn);                                          // dsec.py:43
    o.any = (o.x0 >= -1) && (o.x0 < W) && (o.y0 >= -1) && (o.y0 < H) && (o.t0 >= -1) && (o.t0 < B);
    return o;
}

// Quantise one float32 contribution to a 2^-30 fixed-point integer.  Scaling by a power
// of two is exact, so the only rounding is the float->int conversion, which is exact for
// |w| >= 2^-7 and otherwise rounds at 2^-30 (9.3e-10 absolute).
__device__ __forceinline__ long long quantise(float w) {
    return __float2ll_rn(__fmul_rn(w, kFixScale));
}

// Calls f(xl, yl, tl, w) for each in-bounds corner, in the reference's nest order
// (x outer, y, t inner: dsec.py:47-52) with the reference's left-to-right products.
template <typename F>
__device__ __forceinline__ void for_each_corner(const Event& e, const Origin& o, int H, int W, int B, F&& f) {
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
        const int xl = o.x0 + dx;
        if (xl < 0 || xl >= W) continue;
        const float vx = __fmul_rn(e.value, tent(xl, e.x));
#pragma unroll
        for (int dy = 0; dy < 2; ++dy) {
            const int yl = o.y0 + dy;
            if (yl < 0 || yl >= H) continue;
            const float vxy = __fmul_rn(vx, tent(yl, e.y));
#pragma unroll
            for (int dt = 0; dt < 2; ++dt) {
                const int tl = o.t0 + dt;
                if (tl < 0 || tl >= B) continue;
                f(xl, yl, tl, __fmul_rn(vxy, tent(tl, e.tn)));
            }
        }
    }
}

}  // namespace cmda
