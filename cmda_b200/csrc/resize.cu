// f-4 -- PIL-exact bilinear resize of 8-bit images and the uint8 -> [-1, 1] float hand-off, so that the
// Cityscapes source branch of the loader can stay on the device:
//   img_time_res = Image.open(IC png).convert('L').resize(size, Image.BILINEAR) -> crop -> flip
//                  -> (x / 255 - 0.5) / 0.5 -> repeat(3, 1, 1)          /root/reference/mmseg/datasets/cityscapes_ic.py:175-183, 207-209
//   crop_image   = raw_image.resize(size, Image.BILINEAR).crop(...)      cityscapes_ic.py:152-156 (feeds get_image_change_from_pil, :237-240)
// Image.resize(BILINEAR) is third-party arithmetic (Pillow's Resample.c, absent from /root/reference; Pillow
// is installed here and the parity tests compare against it bit for bit).  Its published algorithm, followed
// here: a separable convolution, horizontal pass then vertical pass with an 8-bit intermediate image;
// per output index the taps are the input samples whose centres lie within `support` = max(scale, 1) of the
// output centre (in0 + (xx + 0.5) * scale), weighted by the triangle filter, normalised in double precision,
// converted to 22-bit fixed point ((int)(0.5 + w * 2^22)), accumulated in int32 starting from 2^21 and
// clipped to 0..255 after the shift.  The coefficients are computed on the device in IEEE double (no FMA
// contraction: -fmad=false), one thread per output index, which keeps the call free of host buffers.
#include "common.cuh"

namespace cmda {

constexpr int kPrecisionBits = 32 - 8 - 2;       // Pillow: PRECISION_BITS
constexpr int kMaxTaps = 64;                      // ksize = ceil(support) * 2 + 1 <= 64: down-scaling up to 31x

struct ResizeAxis {
    int in_size, out_size, ksize;
};

__host__ __device__ inline int resize_ksize(int in_size, int out_size) {
    double filterscale = static_cast<double>(in_size) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 1.0 * filterscale;     // bilinear: support 1.0
    int c = static_cast<int>(support);
    if (static_cast<double>(c) < support) ++c;    // ceil
    return c * 2 + 1;
}

// Pillow precompute_coeffs + normalize_coeffs_8bpc for one axis: bounds[xx] = (xmin, count), kk[xx][ksize].
__global__ void resize_coeffs_kernel(ResizeAxis ax, int2* __restrict__ bounds, int* __restrict__ kk) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= ax.out_size) return;
    const double scale = static_cast<double>(ax.in_size) / ax.out_size;
    double filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 1.0 * filterscale;
    const double center = 0.0 + (xx + 0.5) * scale;
    const double ss = 1.0 / filterscale;
    int xmin = static_cast<int>(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = static_cast<int>(center + support + 0.5);
    if (xmax > ax.in_size) xmax = ax.in_size;
    xmax -= xmin;
    double w[kMaxTaps];
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        double a = (x + xmin - center + 0.5) * ss;
        if (a < 0.0) a = -a;
        const double v = a < 1.0 ? 1.0 - a : 0.0;  // bilinear_filter
        w[x] = v;
        ww += v;
    }
    int* k = kk + static_cast<size_t>(xx) * ax.ksize;
    for (int x = 0; x < ax.ksize; ++x) {
        double v = 0.0;
        if (x < xmax) v = ww != 0.0 ? w[x] / ww : w[x];
        k[x] = v < 0.0 ? static_cast<int>(-0.5 + v * (1 << kPrecisionBits)) : static_cast<int>(0.5 + v * (1 << kPrecisionBits));
    }
    bounds[xx] = make_int2(xmin, xmax);
}

__device__ __forceinline__ unsigned clip8(int v) {
    v >>= kPrecisionBits;
    return static_cast<unsigned>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src [rows, in_w, C] -> dst [rows, out_w, C]
template <int C>
__global__ void __launch_bounds__(256)
resize_horizontal_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long rows, int in_w, int out_w, int ksize,
                         const int2* __restrict__ bounds, const int* __restrict__ kk) {
    const long long n = rows * out_w;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = i / out_w;
        const int xx = static_cast<int>(i - r * out_w);
        const int2 b = __ldg(bounds + xx);
        const int* k = kk + static_cast<size_t>(xx) * ksize;
        const uint8_t* in = src + (static_cast<size_t>(r) * in_w + b.x) * C;
        int acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
        for (int x = 0; x < b.y; ++x) {
            const int kv = __ldg(k + x);
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] += static_cast<int>(__ldg(in + x * C + c)) * kv;
        }
#pragma unroll
        for (int c = 0; c < C; ++c) dst[static_cast<size_t>(i) * C + c] = static_cast<uint8_t>(clip8(acc[c]));
    }
}

// vertical pass: src [S, in_h, w, C] -> dst [S, out_h, w, C]; a thread produces one byte column element
__global__ void __launch_bounds__(256)
resize_vertical_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, int S, int in_h, int out_h, long long row_bytes,
                       int ksize, const int2* __restrict__ bounds, const int* __restrict__ kk) {
    const long long n = static_cast<long long>(S) * out_h * row_bytes;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long col = i % row_bytes;
        const long long sy = i / row_bytes;
        const int yy = static_cast<int>(sy % out_h);
        const long long s = sy / out_h;
        const int2 b = __ldg(bounds + yy);
        const int* k = kk + static_cast<size_t>(yy) * ksize;
        const uint8_t* in = src + (static_cast<size_t>(s) * in_h + b.x) * row_bytes + col;
        int acc = 1 << (kPrecisionBits - 1);
        for (int y = 0; y < b.y; ++y) acc += static_cast<int>(__ldg(in + static_cast<size_t>(y) * row_bytes)) * __ldg(k + y);
        dst[i] = static_cast<uint8_t>(clip8(acc));
    }
}

// crop -> horizontal flip -> float32 -> (x / 255.0 - 0.5) / 0.5 -> repeat (cityscapes_ic.py:177-183, 207-209)
struct CropTable {
    int crop_x[kMaxWindows], crop_y[kMaxWindows];
    unsigned char flip[kMaxWindows];
};
__global__ void __launch_bounds__(256)
u8_crop_center_kernel(const uint8_t* __restrict__ src, int H, int W, CropTable tab, int crop_w, int crop_h, int repeat,
                      float* __restrict__ out) {
    const int s = blockIdx.y;
    const long long n = static_cast<long long>(crop_w) * crop_h;
    const uint8_t* in = src + static_cast<size_t>(s) * H * W;
    float* o = out + static_cast<size_t>(s) * repeat * n;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int y = static_cast<int>(i / crop_w), x = static_cast<int>(i - static_cast<long long>(y) * crop_w);
        const int sx = tab.crop_x[s] + (tab.flip[s] ? crop_w - 1 - x : x), sy = tab.crop_y[s] + y;
        const float v = static_cast<float>(__ldg(in + static_cast<size_t>(sy) * W + sx));
        const float r = __fdiv_rn(__fsub_rn(__fdiv_rn(v, 255.0f), 0.5f), 0.5f);
        for (int k = 0; k < repeat; ++k) o[static_cast<size_t>(k) * n + i] = r;
    }
}

size_t resize_workspace_bytes(int S, int H, int W, int C, int out_h, int out_w) {
    const int kh = resize_ksize(W, out_w), kv = resize_ksize(H, out_h);
    return align_up(sizeof(int2) * out_w, 256) + align_up(sizeof(int) * static_cast<size_t>(out_w) * kh, 256) +
           align_up(sizeof(int2) * out_h, 256) + align_up(sizeof(int) * static_cast<size_t>(out_h) * kv, 256) +
           align_up(static_cast<size_t>(S) * H * out_w * C, 256) + 256;
}

int launch_resize_bilinear(const uint8_t* src, int C, int S, int H, int W, int out_h, int out_w, uint8_t* dst, void* ws,
                           size_t ws_bytes, cudaStream_t st) {
    const int kh = resize_ksize(W, out_w), kv = resize_ksize(H, out_h);
    if (kh > kMaxTaps || kv > kMaxTaps) return CMDA_ERR_UNSUPPORTED;
    if (ws_bytes < resize_workspace_bytes(S, H, W, C, out_h, out_w)) return CMDA_ERR_WORKSPACE;
    char* b = static_cast<char*>(ws);
    int2* bounds_h = reinterpret_cast<int2*>(b); b += align_up(sizeof(int2) * out_w, 256);
    int* kk_h = reinterpret_cast<int*>(b); b += align_up(sizeof(int) * static_cast<size_t>(out_w) * kh, 256);
    int2* bounds_v = reinterpret_cast<int2*>(b); b += align_up(sizeof(int2) * out_h, 256);
    int* kk_v = reinterpret_cast<int*>(b); b += align_up(sizeof(int) * static_cast<size_t>(out_h) * kv, 256);
    uint8_t* tmp = reinterpret_cast<uint8_t*>(b);
    resize_coeffs_kernel<<<(out_w + 127) / 128, 128, 0, st>>>(ResizeAxis{W, out_w, kh}, bounds_h, kk_h);
    resize_coeffs_kernel<<<(out_h + 127) / 128, 128, 0, st>>>(ResizeAxis{H, out_h, kv}, bounds_v, kk_v);
    // Pillow skips a pass whose size does not change; an identity pass has a single tap of weight 2^22 and copies
    const long long rows = static_cast<long long>(S) * H;
    const uint8_t* mid = src;
    if (out_w != W) {
        long long n = rows * out_w;
        const int blocks = static_cast<int>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
        if (C == 1) resize_horizontal_kernel<1><<<blocks, 256, 0, st>>>(src, out_h != H ? tmp : dst, rows, W, out_w, kh, bounds_h, kk_h);
        else resize_horizontal_kernel<3><<<blocks, 256, 0, st>>>(src, out_h != H ? tmp : dst, rows, W, out_w, kh, bounds_h, kk_h);
        mid = tmp;
    }
    if (out_h != H) {
        const long long row_bytes = static_cast<long long>(out_w) * C;
        long long n = static_cast<long long>(S) * out_h * row_bytes;
        const int blocks = static_cast<int>(n / 256 + 1 < 148 * 16 ? n / 256 + 1 : 148 * 16);
        resize_vertical_kernel<<<blocks, 256, 0, st>>>(mid, dst, S, H, out_h, row_bytes, kv, bounds_v, kk_v);
    } else if (out_w == W) {
        CMDA_CUDA_TRY(cudaMemcpyAsync(dst, src, static_cast<size_t>(S) * H * W * C, cudaMemcpyDeviceToDevice, st));
    }
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

int launch_u8_crop_center(const uint8_t* src, int S, int H, int W, const int* crop_x, const int* crop_y, const int* flip, int crop_w,
                          int crop_h, int repeat, float* out, cudaStream_t st) {
    for (int s0 = 0; s0 < S; s0 += kMaxWindows) {
        const int sn = (S - s0) < kMaxWindows ? (S - s0) : kMaxWindows;
        CropTable tab{};
        for (int k = 0; k < sn; ++k) { tab.crop_x[k] = crop_x[s0 + k]; tab.crop_y[k] = crop_y[s0 + k]; tab.flip[k] = flip[s0 + k] ? 1 : 0; }
        const long long n = static_cast<long long>(crop_w) * crop_h;
        long long gx = (n + 255) / 256;
        const long long cap = (148LL * 8 + sn - 1) / sn;
        if (gx > cap) gx = cap;
        u8_crop_center_kernel<<<dim3(static_cast<unsigned>(gx), sn), 256, 0, st>>>(
            src + static_cast<size_t>(s0) * H * W, H, W, tab, crop_w, crop_h, repeat, out + static_cast<size_t>(s0) * repeat * n);
        CMDA_LAUNCH_CHECK();
    }
    return CMDA_OK;
}

}  // namespace cmda
