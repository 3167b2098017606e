// K3 -- events_norm: global z-score over the non-zero voxels of one window, sign split,
// clamp, min-max to [0, r] / [-r, 0], sum.
// Follows /root/reference/mmseg/datasets/dsec.py:80-121 (numeric clip_range) and
// tensor_normalize_to_range dsec.py:73-77.
//
// Two phases per window: (1) statistics (count / sum / sum of squares / min and max of the
// non-zero voxels) reduced in a FIXED order -- kStatBlocks partials per window, each a fixed
// thread->element mapping and a fixed butterfly, merged sequentially -- so the result is
// bit-reproducible; (2) an element-wise apply.  The min/max of the clamped positive and
// negative parts that the reference obtains with four more full-grid reductions follow
// from the non-zero min/max, because every step between the raw value and the clamped
// part is a monotone non-decreasing float32 map.  Sums are accumulated in float64 (the
// reference's float32 torch.sum order is third-party; effect on the output <= 2e-7).
#include "common.cuh"

namespace cmda {

constexpr int kNormThreads = 256;

struct BlockStats {
    double sum, sumsq;
    long long nnz;
    float mn, mx;
};

__device__ __forceinline__ void stats_add(BlockStats& st, float v) {
    if (v != 0.0f) {                                   // dsec.py:88
        st.nnz += 1;
        st.mn = fminf(st.mn, v);
        st.mx = fmaxf(st.mx, v);
    }
    st.sum += static_cast<double>(v);                  // dsec.py:91 events.sum()
    st.sumsq += static_cast<double>(__fmul_rn(v, v));  // dsec.py:92 (events ** 2).sum()
}

__device__ void block_stats_store(BlockStats st, PartialStats* __restrict__ dst) {
    __shared__ double s_sum[kNormThreads / 32], s_sq[kNormThreads / 32];
    __shared__ long long s_n[kNormThreads / 32];
    __shared__ float s_mn[kNormThreads / 32], s_mx[kNormThreads / 32];
    st.sum = warp_sum(st.sum);
    st.sumsq = warp_sum(st.sumsq);
    st.nnz = warp_sum(st.nnz);
    st.mn = warp_min(st.mn);
    st.mx = warp_max(st.mx);
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_sum[wid] = st.sum; s_sq[wid] = st.sumsq; s_n[wid] = st.nnz; s_mn[wid] = st.mn; s_mx[wid] = st.mx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        PartialStats o;
        o.sum = 0.0; o.sumsq = 0.0; o.nnz = 0; o.min_nz = INFINITY; o.max_nz = -INFINITY;
        for (int w = 0; w < kNormThreads / 32; ++w) {   // fixed order
            o.sum += s_sum[w]; o.sumsq += s_sq[w]; o.nnz += s_n[w];
            o.min_nz = fminf(o.min_nz, s_mn[w]); o.max_nz = fmaxf(o.max_nz, s_mx[w]);
        }
        *dst = o;
    }
}

// FROM_I64: src is the 2^-30 fixed-point accumulator grid; the float32 raw grid is written
// to raw (value = correctly rounded float32 of the exact integer sum, times 2^-30).
template <bool FROM_I64>
__global__ void __launch_bounds__(kNormThreads)
convert_stats_kernel(const long long* __restrict__ acc, float* __restrict__ raw, long long V,
                     PartialStats* __restrict__ partials) {
    const int s = blockIdx.y;
    const long long seg = (V + kStatBlocks - 1) / kStatBlocks;
    const long long lo = static_cast<long long>(blockIdx.x) * seg;
    const long long hi = (lo + seg < V) ? lo + seg : V;
    BlockStats st{0.0, 0.0, 0, INFINITY, -INFINITY};
    float* r = raw + static_cast<size_t>(s) * V;
    const long long* a = FROM_I64 ? acc + static_cast<size_t>(s) * V : nullptr;
    for (long long i = lo + threadIdx.x; i < hi; i += kNormThreads) {
        float v;
        if (FROM_I64) {
            v = __fmul_rn(__ll2float_rn(a[i]), kFixInv);
            r[i] = v;
        } else {
            v = r[i];
        }
        stats_add(st, v);
    }
    block_stats_store(st, partials + static_cast<size_t>(s) * kStatBlocks + blockIdx.x);
}

struct NormParams {
    int has_nz;      // num_nonzeros > 0 (dsec.py:90)
    int all_nan;     // stddev is NaN -> every output is NaN, as in the reference
    float mean, den; // dsec.py:91-93
    float clip, final_range;
    float pmin, pden, nmin, nden;
    float rden, rpden, rnden;     // correctly rounded reciprocals of den / pden / nden (div_by_reused)
    float p_of_zero, n_of_zero;   // the normalised positive / negative part of a voxel whose part is 0
};

__device__ __forceinline__ float zscore(float e, const NormParams& q) {
    if (!q.has_nz) return e;
    const float m = (e != 0.0f) ? 1.0f : 0.0f;                          // dsec.py:93 mask
    return div_by_reused(__fmul_rn(m, __fsub_rn(e, q.mean)), q.den, q.rden);   // dsec.py:94
}
__device__ __forceinline__ float pos_part(float z, float clip) {
    const float p = z < 0.0f ? 0.0f : z;                                // dsec.py:108
    return fminf(fmaxf(p, 0.0f), clip);                                 // dsec.py:110
}
__device__ __forceinline__ float neg_part(float z, float clip) {
    const float n = z > 0.0f ? 0.0f : z;                                // dsec.py:112
    return fminf(fmaxf(n, -clip), 0.0f);                                // dsec.py:114
}

__device__ NormParams make_norm_params(const PartialStats* __restrict__ partials, long long V, float clip,
                                       float final_range) {
    double sum = 0.0, sumsq = 0.0;
    long long nnz = 0;
    float mn = INFINITY, mx = -INFINITY;
    for (int b = 0; b < kStatBlocks; ++b) {             // fixed order
        const PartialStats p = partials[b];
        sum += p.sum; sumsq += p.sumsq; nnz += p.nnz;
        mn = fminf(mn, p.min_nz); mx = fmaxf(mx, p.max_nz);
    }
    NormParams q{};
    q.clip = clip;
    q.final_range = final_range;
    q.has_nz = nnz > 0;
    q.mean = 0.0f;
    q.den = 1.0f;
    q.rden = 1.0f;
    if (q.has_nz) {
        const float fn = __ll2float_rn(nnz);
        q.mean = __fdiv_rn(__double2float_rn(sum), fn);                                  // dsec.py:91
        const float var = __fsub_rn(__fdiv_rn(__double2float_rn(sumsq), fn), __fmul_rn(q.mean, q.mean));
        const float sd = __fsqrt_rn(var);                                                // dsec.py:92
        q.den = __fadd_rn(sd, 1e-8f);
        q.rden = __frcp_rn(q.den);
        q.all_nan = isnan(sd);
    }
    // min / max of the clamped parts over the whole grid: values come from the non-zero
    // extremes and, when any voxel is zero, from z(0)
    const bool has_zero = nnz < V;
    float pmax = -INFINITY, pmin = INFINITY, nmax = -INFINITY, nmin = INFINITY;
    if (q.has_nz) {
        const float zhi = zscore(mx, q), zlo = zscore(mn, q);
        pmax = pos_part(zhi, clip); pmin = pos_part(zlo, clip);
        nmax = neg_part(zhi, clip); nmin = neg_part(zlo, clip);
    }
    if (has_zero) {
        const float z0 = zscore(0.0f, q);
        pmax = fmaxf(pmax, pos_part(z0, clip)); pmin = fminf(pmin, pos_part(z0, clip));
        nmax = fmaxf(nmax, neg_part(z0, clip)); nmin = fminf(nmin, neg_part(z0, clip));
    }
    q.pmin = pmin;
    q.pden = __fadd_rn(__fsub_rn(pmax, pmin), 1e-8f);                                    // dsec.py:76
    q.nmin = nmin;
    q.nden = __fadd_rn(__fsub_rn(nmax, nmin), 1e-8f);
    q.rpden = __frcp_rn(q.pden);
    q.rnden = __frcp_rn(q.nden);
    // a voxel with z >= 0 has negative part 0 and vice versa: that half of dsec.py:111 / 115 is
    // the same value for every such voxel
    q.p_of_zero = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(0.0f, q.pmin), q.pden), final_range), 0.0f);
    q.n_of_zero = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(0.0f, q.nmin), q.nden), final_range), -final_range);
    return q;
}

__device__ __forceinline__ float norm_one(float e, const NormParams& q, bool enforce) {
    if (q.all_nan) return __int_as_float(0x7fc00000);
    const float z = zscore(e, q);
    if (enforce) {                                                                       // dsec.py:106-117
        // only the part on z's side of zero differs from voxel to voxel; the other one is the
        // per-grid constant computed in make_norm_params (same operations, same bits)
        const bool neg = z < 0.0f;
        // clamp(z, 0, clip) for z >= 0 and clamp(z, -clip, 0) for z < 0 are both clamp(z, -clip, clip) when clip >= 0
        const float part = q.clip >= 0.0f ? fminf(fmaxf(z, -q.clip), q.clip)
                                          : (neg ? neg_part(z, q.clip) : pos_part(z, q.clip));
        float v = div_by_reused(__fsub_rn(part, neg ? q.nmin : q.pmin), neg ? q.nden : q.pden, neg ? q.rnden : q.rpden);
        v = __fadd_rn(__fmul_rn(v, q.final_range), neg ? -q.final_range : 0.0f);         // * (r - 0) + 0 | * (0 - (-r)) + (-r)
        return neg ? __fadd_rn(q.p_of_zero, v) : __fadd_rn(v, q.n_of_zero);
    }
    float c = fminf(fmaxf(z, -q.clip), q.clip);                                          // dsec.py:119
    c = __fmul_rn(c, q.final_range);
    return __fmul_rn(__fdiv_rn(c, q.clip), q.final_range);                               // dsec.py:120
}

template <bool VEC>
__global__ void __launch_bounds__(kNormThreads, 4)
norm_apply_kernel(const float* raw, float* out, long long V,   // raw may alias out (in-place call)
                  const PartialStats* __restrict__ partials, WindowTable tab, float final_range, int enforce) {
    __shared__ NormParams s_q;
    dependency_wait();
    const int s = blockIdx.y;
    if (threadIdx.x == 0)
        s_q = make_norm_params(partials + static_cast<size_t>(s) * kStatBlocks, V, tab.w[s].clip, final_range);
    __syncthreads();
    const NormParams q = s_q;
    const float* r = raw + static_cast<size_t>(s) * V;
    float* o = out + static_cast<size_t>(s) * V;
    if (VEC) {
        // four independent 16-byte loads in flight per thread before the first is consumed
        const long long n4 = V / 4;
        const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
        long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n4; i += 4 * stride) {
            float4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] = reinterpret_cast<const float4*>(r)[i + u * stride];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                stg_stream_f4(o + (i + u * stride) * 4, make_float4(norm_one(v[u].x, q, enforce), norm_one(v[u].y, q, enforce),
                                                                    norm_one(v[u].z, q, enforce), norm_one(v[u].w, q, enforce)));
        }
        for (; i < n4; i += stride) {
            const float4 v = reinterpret_cast<const float4*>(r)[i];
            stg_stream_f4(o + i * 4, make_float4(norm_one(v.x, q, enforce), norm_one(v.y, q, enforce),
                                                 norm_one(v.z, q, enforce), norm_one(v.w, q, enforce)));
        }
    } else {
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < V;
             i += static_cast<long long>(gridDim.x) * blockDim.x)
            o[i] = norm_one(r[i], q, enforce);
    }
}

// events_norm apply fused with the dataset's post-voxel augmentation (dsec.py:304-319): every output
// pixel of the resized crop normalises its four raw taps on the fly, so the normalised full grid is never
// materialised.  Bilinear weights follow ATen's upsample_bilinear2d with align_corners=False
// (src = scale * (dst + 0.5) - 0.5 clamped at 0, scale = in / out in float32).
struct AugTable {
    int crop_x[kMaxWindows], crop_y[kMaxWindows];
    unsigned char flip[kMaxWindows];
};
struct AugSpec {
    int crop_w, crop_h, out_w, out_h, avg_bins, repeat;
};

__global__ void __launch_bounds__(256)
norm_augment_kernel(const float* __restrict__ raw, float* __restrict__ out, int B, int H, int W,
                    const PartialStats* __restrict__ partials, WindowTable tab, AugTable aug, AugSpec sp, float final_range,
                    int enforce) {
    __shared__ NormParams s_q;
    const int s = blockIdx.y;
    const long long V = static_cast<long long>(B) * H * W;
    if (threadIdx.x == 0) s_q = make_norm_params(partials + static_cast<size_t>(s) * kStatBlocks, V, tab.w[s].clip, final_range);
    __syncthreads();
    const NormParams q = s_q;
    const float* r = raw + static_cast<size_t>(s) * V;
    const int Bo = sp.avg_bins ? 1 : B;
    const size_t oplane = static_cast<size_t>(sp.out_h) * sp.out_w;
    float* o = out + static_cast<size_t>(s) * sp.repeat * Bo * oplane;
    const float sy = static_cast<float>(sp.crop_h) / static_cast<float>(sp.out_h);
    const float sx = static_cast<float>(sp.crop_w) / static_cast<float>(sp.out_w);
    const int cx = aug.crop_x[s], cy = aug.crop_y[s];
    const bool flip = aug.flip[s] != 0;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < static_cast<long long>(oplane);
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int oy = static_cast<int>(i / sp.out_w), ox = static_cast<int>(i - static_cast<long long>(oy) * sp.out_w);
        float fy = sy * (static_cast<float>(oy) + 0.5f) - 0.5f;
        float fx = sx * (static_cast<float>(ox) + 0.5f) - 0.5f;
        fy = fy < 0.0f ? 0.0f : fy;
        fx = fx < 0.0f ? 0.0f : fx;
        const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
        const int y1 = y0 + (y0 < sp.crop_h - 1 ? 1 : 0), x1 = x0 + (x0 < sp.crop_w - 1 ? 1 : 0);
        const float ly1 = fy - static_cast<float>(y0), ly0 = 1.0f - ly1;
        const float lx1 = fx - static_cast<float>(x0), lx0 = 1.0f - lx1;
        // crop then flip: flipped[x] = crop[crop_w - 1 - x]
        const int gx0 = cx + (flip ? sp.crop_w - 1 - x0 : x0), gx1 = cx + (flip ? sp.crop_w - 1 - x1 : x1);
        const int gy0 = cy + y0, gy1 = cy + y1;
        float m00 = 0.0f, m01 = 0.0f, m10 = 0.0f, m11 = 0.0f;
        for (int b = 0; b < B; ++b) {
            const float* pb = r + static_cast<size_t>(b) * H * W;
            const float v00 = norm_one(__ldg(pb + static_cast<size_t>(gy0) * W + gx0), q, enforce);
            const float v01 = norm_one(__ldg(pb + static_cast<size_t>(gy0) * W + gx1), q, enforce);
            const float v10 = norm_one(__ldg(pb + static_cast<size_t>(gy1) * W + gx0), q, enforce);
            const float v11 = norm_one(__ldg(pb + static_cast<size_t>(gy1) * W + gx1), q, enforce);
            if (sp.avg_bins) {
                m00 += v00; m01 += v01; m10 += v10; m11 += v11;            // torch.mean(dim=1): sum, then / B
                if (b + 1 < B) continue;
                const float fb = static_cast<float>(B);
                m00 /= fb; m01 /= fb; m10 /= fb; m11 /= fb;
            } else {
                m00 = v00; m01 = v01; m10 = v10; m11 = v11;
            }
            const float val = ly0 * (lx0 * m00 + lx1 * m01) + ly1 * (lx0 * m10 + lx1 * m11);
            const int ob = sp.avg_bins ? 0 : b;
            for (int rep = 0; rep < sp.repeat; ++rep) o[static_cast<size_t>(rep * Bo + ob) * oplane + i] = val;
        }
    }
}

int launch_norm_augment(const float* raw, float* out, int S, int B, int H, int W, const PartialStats* partials,
                        const WindowTable& tab, const int* crop_x, const int* crop_y, const int* flip, int crop_w, int crop_h,
                        int out_w, int out_h, int avg_bins, int repeat, float final_range, int enforce, cudaStream_t s) {
    AugTable aug{};
    for (int k = 0; k < S; ++k) { aug.crop_x[k] = crop_x[k]; aug.crop_y[k] = crop_y[k]; aug.flip[k] = flip[k] ? 1 : 0; }
    AugSpec sp{crop_w, crop_h, out_w, out_h, avg_bins, repeat};
    const long long oplane = static_cast<long long>(out_h) * out_w;
    long long gx = (oplane + kNormThreads - 1) / kNormThreads;
    long long cap = (148LL * 8 + S - 1) / S;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    norm_augment_kernel<<<dim3(static_cast<unsigned>(gx), S), kNormThreads, 0, s>>>(raw, out, B, H, W, partials, tab, aug, sp,
                                                                                  final_range, enforce);
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

int launch_convert_stats(const long long* acc, float* raw, int S, long long V, PartialStats* partials,
                         cudaStream_t s) {
    dim3 grid(kStatBlocks, S);
    if (acc) convert_stats_kernel<true><<<grid, kNormThreads, 0, s>>>(acc, raw, V, partials);
    else convert_stats_kernel<false><<<grid, kNormThreads, 0, s>>>(nullptr, raw, V, partials);
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

int launch_norm_apply(const float* raw, float* out, int S, long long V, const PartialStats* partials,
                      const WindowTable& tab, float final_range, int enforce, cudaStream_t s) {
    const bool vec = (V % 4 == 0) && ((reinterpret_cast<uintptr_t>(raw) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    long long per = vec ? V / 4 : V;
    long long gx = (per + kNormThreads - 1) / kNormThreads;
    long long cap = (148LL * 8 + S - 1) / S;
    if (cap < 1) cap = 1;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(static_cast<unsigned>(gx), S);
    // programmatic dependent launch: the CTAs may queue up while the kernel before (statistics regroup / conversion) drains
    if (vec) CMDA_CUDA_TRY(launch_dependent(norm_apply_kernel<true>, grid, dim3(kNormThreads), 0, s, raw, out, V, partials, tab, final_range, enforce));
    else CMDA_CUDA_TRY(launch_dependent(norm_apply_kernel<false>, grid, dim3(kNormThreads), 0, s, raw, out, V, partials, tab, final_range, enforce));
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

}  // namespace cmda
