// K4 / K5 -- log-intensity-change pseudo-events.
//   K4 frame pair : /root/reference/create_cityscapes_image_change.py:16-35 (get_image_change)
//   K5 shift pair : /root/reference/mmseg/datasets/utils.py:87-152 (get_ic, get_image_change_from_pil)
//
// Both are u8 -> f32 element-wise maps around two global min/max pairs per term, so each
// is two passes over the (L2-resident) u8 input: pass 1 reduces the min/max of the log
// difference (the min/max of the clamped parts follow from it, every step in between being
// monotone), pass 2 re-evaluates the difference and writes the result with 128-bit stores.  There are only 256 distinct log values, so the logarithm is a
// 256-entry table the CALLER computes with numpy exactly as the reference does; every
// other operation is an IEEE float32 op with one rounding, which makes the whole path
// bit-exact against the reference.  HBM-bound: algorithmic bytes (1+1+4)·H·W per pair,
// (1+4)·H·W per gray shift-pair image.
#include "common.cuh"

namespace cmda {

constexpr int kPxPerThread = 4;
constexpr int kImgThreads = 256;

// The 256-entry log table lives in shared memory once per bank (32 KB): lane l reads entry v at
// word v * 32 + l, i.e. always from its own bank -- a data-dependent table lookup with no bank
// conflicts.  The fill reads the kernel-parameter copy with a warp-uniform index (one broadcast).
constexpr int kLutWords = 256 * 32;
__device__ __forceinline__ void load_banked_lut(float* __restrict__ s_lut, const LogLut& in) {
    for (int i = threadIdx.x; i < kLutWords; i += blockDim.x) s_lut[i] = in.v[i >> 5];
}
#define CMDA_LUT(v) s_lut[((v) << 5) | (threadIdx.x & 31)]

// Direction codes of one term: 0 left, 1 right, 2 up, 3 down (utils.py:129-132).
struct TermList {
    int n;
    int dir[4];
};

__host__ __device__ inline TermList terms_of(int direction) {
    TermList t{};
    if (direction == CMDA_DIR_ALL) {  // up, left, down, right -- utils.py:133-137
        t.n = 4; t.dir[0] = 2; t.dir[1] = 0; t.dir[2] = 3; t.dir[3] = 1;
    } else {                           // column-shift term first, row-shift term second -- utils.py:139-151
        t.n = 2;
        t.dir[0] = (direction == CMDA_DIR_LEFTDOWN || direction == CMDA_DIR_LEFTUP) ? 0 : 1;
        t.dir[1] = (direction == CMDA_DIR_RIGHTUP || direction == CMDA_DIR_LEFTUP) ? 2 : 3;
    }
    return t;
}

// gray value of the shifted copy at (r, c): the first (right/down) or last (left/up)
// `s` columns/rows are left unshifted (utils.py:129-132, 140-148)
__device__ __forceinline__ int shifted_at(const uint8_t* __restrict__ g, int H, int W, int r, int c, int s, int dir) {
    int rr = r, cc = c;
    if (dir == 0) { if (c < W - s) cc = c + s; }
    else if (dir == 1) { if (c >= s) cc = c - s; }
    else if (dir == 2) { if (r < H - s) rr = r + s; }
    else { if (r >= s) rr = r - s; }
    return __ldg(g + static_cast<size_t>(rr) * W + cc);
}

// dead zone + sign split + clamp (utils.py:95-101 / create_cityscapes_image_change.py:22-28)
__device__ __forceinline__ void split_clamp(float d, float thr, float clip, float& pos, float& neg) {
    if (fabsf(d) <= thr) d = 0.0f;
    pos = d < 0.0f ? 0.0f : d;
    pos = fminf(fmaxf(pos, 0.0f), clip);
    neg = d > 0.0f ? 0.0f : d;
    neg = fminf(fmaxf(neg, -clip), 0.0f);
}

// Every step between the log difference d and the clamped positive / negative part is a monotone
// non-decreasing float32 map (dead zone, sign split, clamp: utils.py:95-100), so the global min / max of
// the two parts that tensor_normalize_to_range needs (utils.py:10-14) are the parts of the global min /
// max of d.  Pass 1 therefore only tracks d's range, as order-preserving unsigned keys so that a
// zero-initialised workspace is the neutral element and the merge is an exact, order-independent
// atomicMax: slot [0] = max key(d), slot [1] = max ~key(d) (i.e. the minimum).
__device__ __forceinline__ unsigned ordered_key(float f) {
    const unsigned b = __float_as_uint(f);
    return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}
__device__ __forceinline__ float ordered_unkey(unsigned k) {
    return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xffffffffu));
}

struct MinMaxAcc {
    float dmin, dmax;
    __device__ __forceinline__ void init() { dmin = INFINITY; dmax = -INFINITY; }
    __device__ __forceinline__ void add(float d) { dmin = fminf(dmin, d); dmax = fmaxf(dmax, d); }
};

struct TermRange {
    float pmin, pden, nmin, nden;
    float rpden, rnden;           // correctly rounded reciprocals of the two denominators
    float p_of_zero, n_of_zero;   // normalised part of a pixel whose part is 0 (the side of zero it is not on)
};

// tensor_normalize_to_range of one part (utils.py:10-14): positive part -> [0, 1], negative part -> [-1, 0]
// tensor_normalize_to_range of one part (utils.py:10-14): positive part -> [0, 1], negative part -> [-1, 0].
// "* (1 - 0) + 0" is the identity on the non-negative quotient and is not spelled out; the divisions by the
// per-image denominators are div_by_reused (correctly rounded, common.cuh).
__device__ __forceinline__ float normalize_pos(float pos, const TermRange& r) {
    return div_by_reused(__fsub_rn(pos, r.pmin), r.pden, r.rpden);
}
__device__ __forceinline__ float normalize_neg(float neg, const TermRange& r) {
    return __fadd_rn(div_by_reused(__fsub_rn(neg, r.nmin), r.nden, r.rnden), -1.0f);      // * (0 - (-1)) + (-1)
}

__device__ __forceinline__ TermRange decode_range(const unsigned* __restrict__ slots, float thr, float clip) {
    const float dmax = ordered_unkey(slots[0]);
    const float dmin = ordered_unkey(~slots[1]);
    float pmax, pmin, nmax, nmin;
    split_clamp(dmax, thr, clip, pmax, nmax);
    split_clamp(dmin, thr, clip, pmin, nmin);
    TermRange r;
    r.pmin = pmin;
    r.pden = __fadd_rn(__fsub_rn(pmax, pmin), 1e-8f);   // tensor_max - tensor_min + 1e-8
    r.nmin = nmin;
    r.nden = __fadd_rn(__fsub_rn(nmax, nmin), 1e-8f);
    r.rpden = __frcp_rn(r.pden);
    r.rnden = __frcp_rn(r.nden);
    r.p_of_zero = normalize_pos(0.0f, r);
    r.n_of_zero = normalize_neg(0.0f, r);
    return r;
}

// dead zone, sign split, clamp, both normalisations and their sum (utils.py:95-104) for one pixel.  Only
// the part on d's side of zero varies; the other part is 0 and its normalised value is the per-image
// constant above (same operations, same bits).
__device__ __forceinline__ float normalize_term(float d, float thr, float clip, const TermRange& r) {
    if (fabsf(d) <= thr) d = 0.0f;
    const bool neg = d < 0.0f;
    float part;
    if (clip >= 0.0f) {
        // clamp(d, 0, clip) for d >= 0 and clamp(d, -clip, 0) for d < 0 are both clamp(d, -clip, clip)
        part = fminf(fmaxf(d, -clip), clip);
    } else {        // a negative clip range (min > max in the reference's clamps): spelled out
        part = neg ? fminf(fmaxf(d, -clip), 0.0f) : fminf(fmaxf(d, 0.0f), clip);
    }
    // the side's own min / denominator; the other side contributes its per-image constant
    float v = div_by_reused(__fsub_rn(part, neg ? r.nmin : r.pmin), neg ? r.nden : r.pden, neg ? r.rnden : r.rpden);
    if (neg) v = __fadd_rn(v, -1.0f);                                  // * (0 - (-1)) + (-1)
    return __fadd_rn(v, neg ? r.p_of_zero : r.n_of_zero);               // events + neg (commutative)
}

// Compile-time view of terms_of(): number of terms of a direction code and the shift of term k.
__host__ __device__ constexpr int term_count(int direction) { return direction == CMDA_DIR_ALL ? 4 : 2; }
__host__ __device__ constexpr int term_dir(int direction, int k) {
    if (direction == CMDA_DIR_ALL) return k == 0 ? 2 : (k == 1 ? 0 : (k == 2 ? 3 : 1));
    if (k == 0) return (direction == CMDA_DIR_LEFTDOWN || direction == CMDA_DIR_LEFTUP) ? 0 : 1;
    return (direction == CMDA_DIR_RIGHTUP || direction == CMDA_DIR_LEFTUP) ? 2 : 3;
}

// Table value of byte J of a packed word: the byte scaled to its 128-byte row of the bank-replicated table and the
// lane's bank offset merged in by one logic op (shift, and-or, load).
template <int J>
__device__ __forceinline__ float lut_of_byte(const float* s_lut, unsigned w, unsigned lane_bytes) {
    const unsigned row = J == 0 ? (w << 7) : (w >> (8 * J - 7));
    return *reinterpret_cast<const float*>(reinterpret_cast<const char*>(s_lut) + ((row & 0x7f80u) | lane_bytes));
}

template <int NT>
__device__ __forceinline__ void flush_minmax(MinMaxAcc (&acc)[NT], unsigned* __restrict__ slots) {
#pragma unroll
    for (int k = 0; k < NT; ++k) {
        const unsigned hi = __reduce_max_sync(0xffffffffu, ordered_key(acc[k].dmax));
        const unsigned lo = __reduce_max_sync(0xffffffffu, ~ordered_key(acc[k].dmin));
        if ((threadIdx.x & 31) == 0) {
            atomicMax(slots + k * 4 + 0, hi);
            atomicMax(slots + k * 4 + 1, lo);
        }
    }
}

// ------------------------------------------------------------------ K5 shift pair
template <int NT>
__global__ void __launch_bounds__(kImgThreads)
isr_minmax_kernel(const uint8_t* __restrict__ gray, int H, int W, int shift, TermList terms, LogLut lut_in,
                  float thr, float clip, unsigned* __restrict__ ws) {
    __shared__ float s_lut[kLutWords];
    load_banked_lut(s_lut, lut_in);
    __syncthreads();
    const int img = blockIdx.y;
    const uint8_t* g = gray + static_cast<size_t>(img) * H * W;
    const long long npx = static_cast<long long>(H) * W;
    MinMaxAcc acc[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) acc[k].init();
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npx;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / W), c = static_cast<int>(i - static_cast<long long>(r) * W);
        const float base = CMDA_LUT(__ldg(g + i));
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            acc[k].add(__fsub_rn(CMDA_LUT(shifted_at(g, H, W, r, c, shift, terms.dir[k])), base));  // utils.py:92
        }
    }
    flush_minmax<NT>(acc, ws + static_cast<size_t>(img) * 16);
}

template <int NT, bool VEC>
__global__ void __launch_bounds__(kImgThreads)
isr_apply_kernel(const uint8_t* __restrict__ gray, int H, int W, int shift, TermList terms, LogLut lut_in, float thr,
                 float clip, const unsigned* __restrict__ ws, float* __restrict__ out) {
    __shared__ float s_lut[kLutWords];
    load_banked_lut(s_lut, lut_in);
    __syncthreads();
    const int img = blockIdx.y;
    const uint8_t* g = gray + static_cast<size_t>(img) * H * W;
    float* o = out + static_cast<size_t>(img) * H * W;
    TermRange rng[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) rng[k] = decode_range(ws + static_cast<size_t>(img) * 16 + k * 4, thr, clip);
    const float inv = NT == 4 ? 0.25f : 0.5f;   // x / 4 and x / 2 are exact scalings
    const long long npx = static_cast<long long>(H) * W;
    const long long ngroups = (npx + kPxPerThread - 1) / kPxPerThread;
    for (long long gi = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; gi < ngroups;
         gi += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long i0 = gi * kPxPerThread;
        float res[kPxPerThread];
#pragma unroll
        for (int j = 0; j < kPxPerThread; ++j) {
            const long long i = i0 + j;
            res[j] = 0.0f;
            if (i < npx) {
                const int r = static_cast<int>(i / W), c = static_cast<int>(i - static_cast<long long>(r) * W);
                const float base = CMDA_LUT(__ldg(g + i));
                float sum = 0.0f;
#pragma unroll
                for (int k = 0; k < NT; ++k) {
                    const float d = __fsub_rn(CMDA_LUT(shifted_at(g, H, W, r, c, shift, terms.dir[k])), base);
                    const float t = __fmul_rn(normalize_term(d, thr, clip, rng[k]), inv);
                    sum = (k == 0) ? t : __fadd_rn(sum, t);       // utils.py:137 / 151, left to right
                }
                res[j] = sum;
            }
        }
        if (VEC) {
            stg_stream_f4(o + i0, make_float4(res[0], res[1], res[2], res[3]));
        } else {
#pragma unroll
            for (int j = 0; j < kPxPerThread; ++j)
                if (i0 + j < npx) o[i0 + j] = res[j];
        }
    }
}

// ------------------------------------------------------------------ K5 shift pair, vector path
// The fast path (W % 4 == 0, 4-byte aligned rows): a thread produces 4 consecutive pixels of one row from
// 32-bit loads -- its own word, the word `shift` rows above / below, and the two aligned words that hold the
// columns `shift` to the left / right (funnel-shifted into place); all of them L1 / L2 hits after the first
// touch.  The shifted copy of utils.py:129-132 never reads outside the image: border pixels stay unshifted,
// which is a per-byte select here.  Same arithmetic as the generic kernels above; rows are the work items of
// persistent CTAs so that the bank-replicated log table is filled once per CTA.
// Column-shift geometry of one thread (4 consecutive pixels starting at column c of every row it visits):
// which two aligned words hold the shifted columns, by how many bits to funnel them, and which of the four
// bytes are shifted at all (border pixels stay unshifted, utils.py:129-130).  Row invariant: computed once.
struct ColShift {
    int off0, off1;        // byte offsets inside a row of the two aligned words (clamped into the row)
    unsigned funnel;       // bit count of the funnel shift
    unsigned mask;         // 0xff per byte that takes the shifted value
};
__device__ __forceinline__ ColShift col_shift_of(int c, int W, int delta /* +s: left (c + s), -s: right (c - s) */) {
    ColShift g;
    const int start = c + delta;
    g.mask = 0u;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int cj = c + j;
        const bool shifted = delta > 0 ? (cj < W - delta) : (cj >= -delta);
        if (shifted) g.mask |= 255u << (8 * j);
    }
    // byte j of funnel(w0, w1) must be row[start + j] for every shifted byte; unshifted bytes are masked away,
    // so the offsets only need to be valid addresses
    int a = start & ~3;                       // floor to a multiple of 4 (also for negative start)
    g.funnel = static_cast<unsigned>(start - a) * 8u;
    g.off0 = min(max(a, 0), W - 4);
    g.off1 = min(max(a + 4, 0), W - 4);
    // a shifted byte always comes from a column inside [0, W): when a < 0 (or a + 4 > W - 4) the clamped word is
    // only read for bytes that are masked away, except when the clamp moved a word that still matters; that
    // happens only for the word that starts inside the row, which the clamp leaves where it is
    return g;
}

// The packed words one thread needs for its four pixels of one row: its own word and, per term, the word `shift`
// rows above / below or the two aligned words that hold the columns `shift` to the left / right.  Loaded one row
// ahead of their use so that the L2 latency of row r + 1 hides behind the arithmetic of row r.
template <int DIRECTION>
struct IsrRowWords {
    unsigned base;
    unsigned a[term_count(DIRECTION)], b[term_count(DIRECTION)];     // b only for the column-shift terms
};
template <int DIRECTION>
__device__ __forceinline__ IsrRowWords<DIRECTION> isr_load_row(const unsigned* __restrict__ row_w, int cw, int r, int H, int shift,
                                                               long long shift_words, const ColShift& left,
                                                               const ColShift& right) {
    IsrRowWords<DIRECTION> w;
    w.base = __ldg(row_w + cw);
#pragma unroll
    for (int k = 0; k < term_count(DIRECTION); ++k) {
        const int dir = term_dir(DIRECTION, k);
        if (dir == 2) {                              // up: rows r < H - s take row r + s (utils.py:131)
            w.a[k] = __ldg(row_w + cw + (r < H - shift ? shift_words : 0));
            w.b[k] = 0u;
        } else if (dir == 3) {                       // down: rows r >= s take row r - s (utils.py:132)
            w.a[k] = __ldg(row_w + cw - (r >= shift ? shift_words : 0));
            w.b[k] = 0u;
        } else {                                     // left (utils.py:129) / right (utils.py:130)
            const ColShift& cs = dir == 0 ? left : right;
            w.a[k] = __ldg(row_w + (cs.off0 >> 2));
            w.b[k] = __ldg(row_w + (cs.off1 >> 2));
        }
    }
    return w;
}

template <int DIRECTION, bool APPLY>
__global__ void __launch_bounds__(256)
isr_vec_kernel(const uint8_t* __restrict__ gray, int S, int H, int W, int shift, LogLut lut_in, float thr, float clip,
               unsigned* __restrict__ ws, float* __restrict__ out) {
    constexpr int NT = term_count(DIRECTION);
    __shared__ float s_lut[kLutWords];
    load_banked_lut(s_lut, lut_in);
    __syncthreads();
    const int words = W >> 2;
    const int cw = blockIdx.x * 256 + threadIdx.x;
    const int c = cw * 4;
    const float inv = NT == 4 ? 0.25f : 0.5f;   // x / 4 and x / 2 are exact scalings
    const long long n_rows = static_cast<long long>(S) * H;
    // consecutive rows per CTA: an image's rows stay together (one min/max flush per image)
    const long long per_cta = (n_rows + gridDim.y - 1) / gridDim.y;
    const long long row_begin = per_cta * blockIdx.y, row_end = min(row_begin + per_cta, n_rows);
    if (row_begin >= row_end || cw >= words) return;
    const unsigned lane_bytes = (threadIdx.x & 31u) * 4u;
    const ColShift left = col_shift_of(c, W, shift), right = col_shift_of(c, W, -shift);
    int img = static_cast<int>(row_begin / H), r = static_cast<int>(row_begin - static_cast<long long>(img) * H);
    int cur_img = -1;
    // the images are contiguous: row `row` of the batch starts at word row * words
    const unsigned* row_w = reinterpret_cast<const unsigned*>(gray) + row_begin * words;
    float4* out4 = APPLY ? reinterpret_cast<float4*>(out) + row_begin * words + cw : nullptr;
    const long long shift_words = static_cast<long long>(shift) * words;
    TermRange rng[NT];
    MinMaxAcc acc[NT];
#pragma unroll
    for (int k = 0; k < NT; ++k) acc[k].init();
    IsrRowWords<DIRECTION> cur = isr_load_row<DIRECTION>(row_w, cw, r, H, shift, shift_words, left, right);
    for (long long row = row_begin; row < row_end; ++row, row_w += words, out4 += APPLY ? words : 0) {
        const int r_next = r + 1 == H ? 0 : r + 1;
        IsrRowWords<DIRECTION> nxt = cur;
        if (row + 1 < row_end) nxt = isr_load_row<DIRECTION>(row_w + words, cw, r_next, H, shift, shift_words, left, right);
        if (img != cur_img) {
            if (!APPLY && cur_img >= 0) {
                flush_minmax<NT>(acc, ws + static_cast<size_t>(cur_img) * 16);
#pragma unroll
                for (int k = 0; k < NT; ++k) acc[k].init();
            }
            if (APPLY) {
#pragma unroll
                for (int k = 0; k < NT; ++k) rng[k] = decode_range(ws + static_cast<size_t>(img) * 16 + k * 4, thr, clip);
            }
            cur_img = img;
        }
        const unsigned base_w = cur.base;
        const float base[4] = {lut_of_byte<0>(s_lut, base_w, lane_bytes), lut_of_byte<1>(s_lut, base_w, lane_bytes),
                               lut_of_byte<2>(s_lut, base_w, lane_bytes), lut_of_byte<3>(s_lut, base_w, lane_bytes)};
        float res[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int k = 0; k < NT; ++k) {
            const int dir = term_dir(DIRECTION, k);       // a constant once the loop is unrolled
            unsigned shw = cur.a[k];
            if (dir < 2) {                                // column shift: funnel the two words, border bytes stay unshifted
                const ColShift& cs = dir == 0 ? left : right;
                shw = (__funnelshift_r(cur.a[k], cur.b[k], cs.funnel) & cs.mask) | (base_w & ~cs.mask);
            }
            const float sh[4] = {lut_of_byte<0>(s_lut, shw, lane_bytes), lut_of_byte<1>(s_lut, shw, lane_bytes),
                                 lut_of_byte<2>(s_lut, shw, lane_bytes), lut_of_byte<3>(s_lut, shw, lane_bytes)};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float d = __fsub_rn(sh[j], base[j]);                                    // utils.py:92
                if (APPLY) {
                    const float tv = __fmul_rn(normalize_term(d, thr, clip, rng[k]), inv);
                    res[j] = (k == 0) ? tv : __fadd_rn(res[j], tv);                           // utils.py:137 / 151, left to right
                } else {
                    acc[k].add(d);
                }
            }
        }
        if (APPLY) stg_stream_f4(reinterpret_cast<float*>(out4), make_float4(res[0], res[1], res[2], res[3]));
        cur = nxt;
        r = r_next;
        if (r == 0) ++img;
    }
    if (!APPLY && cur_img >= 0) flush_minmax<NT>(acc, ws + static_cast<size_t>(cur_img) * 16);
}

// ------------------------------------------------------------------ K4 frame pair
__global__ void __launch_bounds__(kImgThreads)
pair_minmax_kernel(const uint8_t* __restrict__ now, const uint8_t* __restrict__ front, long long npx, bool vec,
                   LogLut lut_in, float thr, float clip, unsigned* __restrict__ ws) {
    __shared__ float s_lut[kLutWords];
    load_banked_lut(s_lut, lut_in);
    __syncthreads();
    const int img = blockIdx.y;
    const uint8_t* a = now + static_cast<size_t>(img) * npx;
    const uint8_t* b = front + static_cast<size_t>(img) * npx;
    MinMaxAcc acc[1];
    acc[0].init();
    const unsigned lane_bytes = (threadIdx.x & 31u) * 4u;
    if (vec) {
        const long long n16 = npx / 16;
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const uint4 va = __ldg(reinterpret_cast<const uint4*>(a) + i);
            const uint4 vb = __ldg(reinterpret_cast<const uint4*>(b) + i);
            const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {                                                        // :21
                acc[0].add(__fsub_rn(lut_of_byte<0>(s_lut, wa[q], lane_bytes), lut_of_byte<0>(s_lut, wb[q], lane_bytes)));
                acc[0].add(__fsub_rn(lut_of_byte<1>(s_lut, wa[q], lane_bytes), lut_of_byte<1>(s_lut, wb[q], lane_bytes)));
                acc[0].add(__fsub_rn(lut_of_byte<2>(s_lut, wa[q], lane_bytes), lut_of_byte<2>(s_lut, wb[q], lane_bytes)));
                acc[0].add(__fsub_rn(lut_of_byte<3>(s_lut, wa[q], lane_bytes), lut_of_byte<3>(s_lut, wb[q], lane_bytes)));
            }
        }
    } else {
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npx;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            acc[0].add(__fsub_rn(CMDA_LUT(__ldg(a + i)), CMDA_LUT(__ldg(b + i))));
        }
    }
    flush_minmax<1>(acc, ws + static_cast<size_t>(img) * 16);
}

__device__ __forceinline__ unsigned quantise_u8(float v) {
    // np.uint8(np.around((v + 1) / 2 * 255)) -- create_cityscapes_image_change.py:33
    const float q = __fmul_rn(__fmul_rn(__fadd_rn(v, 1.0f), 0.5f), 255.0f);
    return static_cast<unsigned>(__float2int_rn(q)) & 255u;   // rint = round half to even
}

template <bool VEC>
__global__ void __launch_bounds__(kImgThreads)
pair_apply_kernel(const uint8_t* __restrict__ now, const uint8_t* __restrict__ front, long long npx, LogLut lut_in,
                  float thr, float clip, const unsigned* __restrict__ ws, float* __restrict__ out_f32,
                  uint8_t* __restrict__ out_u8) {
    __shared__ float s_lut[kLutWords];
    load_banked_lut(s_lut, lut_in);
    __syncthreads();
    const int img = blockIdx.y;
    const uint8_t* a = now + static_cast<size_t>(img) * npx;
    const uint8_t* b = front + static_cast<size_t>(img) * npx;
    const TermRange rng = decode_range(ws + static_cast<size_t>(img) * 16, thr, clip);
    const unsigned lane_bytes = (threadIdx.x & 31u) * 4u;
    float* of = out_f32 ? out_f32 + static_cast<size_t>(img) * npx : nullptr;
    uint8_t* ou = out_u8 ? out_u8 + static_cast<size_t>(img) * npx : nullptr;
    if (VEC) {
        const long long n16 = npx / 16;
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n16;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const uint4 va = __ldg(reinterpret_cast<const uint4*>(a) + i);
            const uint4 vb = __ldg(reinterpret_cast<const uint4*>(b) + i);
            const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
            unsigned packed[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float d[4] = {
                    __fsub_rn(lut_of_byte<0>(s_lut, wa[q], lane_bytes), lut_of_byte<0>(s_lut, wb[q], lane_bytes)),
                    __fsub_rn(lut_of_byte<1>(s_lut, wa[q], lane_bytes), lut_of_byte<1>(s_lut, wb[q], lane_bytes)),
                    __fsub_rn(lut_of_byte<2>(s_lut, wa[q], lane_bytes), lut_of_byte<2>(s_lut, wb[q], lane_bytes)),
                    __fsub_rn(lut_of_byte<3>(s_lut, wa[q], lane_bytes), lut_of_byte<3>(s_lut, wb[q], lane_bytes))};
                float r[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) r[j] = normalize_term(d[j], thr, clip, rng);
                if (of) stg_stream_f4(of + i * 16 + q * 4, make_float4(r[0], r[1], r[2], r[3]));
                packed[q] = quantise_u8(r[0]) | (quantise_u8(r[1]) << 8) | (quantise_u8(r[2]) << 16) |
                            (quantise_u8(r[3]) << 24);
            }
            if (ou) reinterpret_cast<uint4*>(ou)[i] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
    } else {
        for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npx;
             i += static_cast<long long>(gridDim.x) * blockDim.x) {
            const float d = __fsub_rn(CMDA_LUT(__ldg(a + i)), CMDA_LUT(__ldg(b + i)));
            const float r = normalize_term(d, thr, clip, rng);
            if (of) of[i] = r;
            if (ou) ou[i] = static_cast<uint8_t>(quantise_u8(r));
        }
    }
}

// ------------------------------------------------------------------ K4 frame pair, table-driven u8 apply (large images)
// The output of a pixel is a function of its (now, front) byte pair and of the image's four extrema only, and the
// arithmetic kernels above are bound by instruction issue, not by memory (34 instructions per pixel:
// profiles/r02_ncu_pseudo_summary.txt).  For the uint8 output of get_image_change (what the reference writes to its
// PNGs, create_cityscapes_image_change.py:32-34) and images large enough to pay for it, the per-pixel arithmetic is
// done ONCE per byte pair instead: pair_table_kernel evaluates normalize_term + the quantisation (the same device
// functions, hence the same bits) for all 65 536 pairs of an image into a 64 KB table in the workspace, and the
// apply pass is a gather from a copy of that table in shared memory (persistent CTAs, two per SM, walk an even share
// of the batch's 16-pixel groups and refill the table when they cross into the next image): 0.100 vs 0.129 ms on C3.
// The same idea for the float32 outputs (frame pair and shift pair: banded float tables in shared memory, global
// table for the pairs outside the band) was built and measured in round 2 and dropped: it halves the instructions
// per pixel but moves the bound to the load / store unit (gathers with 2-3-way bank conflicts, l1tex 60-75 % busy) and
// ends up slower than the arithmetic kernels (frame pair 0.19 vs 0.154 ms, shift pair 0.28 vs 0.237 ms;
// profiles/r02_pseudo_tables.txt).
constexpr int kTabThreads = 1024;
#ifndef CMDA_TABLE_MIN_PIXELS
#define CMDA_TABLE_MIN_PIXELS (1LL << 17)
#endif
constexpr long long kTableMinPixels = CMDA_TABLE_MIN_PIXELS;      // below this the arithmetic kernels are cheaper than the table

constexpr int kTabRowsPerCta = 8;
__global__ void __launch_bounds__(256)
pair_table_kernel(LogLut lut_in, float thr, float clip, const unsigned* __restrict__ ws, uint8_t* __restrict__ tab8) {
    __shared__ float s_l[256];
    __shared__ TermRange s_rng;
    s_l[threadIdx.x] = lut_in.v[threadIdx.x];
    const int img = blockIdx.y, b = threadIdx.x;
    if (threadIdx.x == 0) s_rng = decode_range(ws + static_cast<size_t>(img) * 16, thr, clip);
    __syncthreads();
    const TermRange rng = s_rng;
    const float lb = s_l[b];
#pragma unroll
    for (int q = 0; q < kTabRowsPerCta; ++q) {
        const int a = blockIdx.x * kTabRowsPerCta + q;
        const float r = normalize_term(__fsub_rn(s_l[a], lb), thr, clip, rng);            // d = log(now) - log(front), :21
        tab8[(static_cast<size_t>(img) << 16) + (a << 8) + b] = static_cast<uint8_t>(quantise_u8(r));
    }
}

// u8 output only: the whole table of quantised values in shared memory
__global__ void __launch_bounds__(kTabThreads, 2)
pair_apply_table8_kernel(const uint8_t* __restrict__ now, const uint8_t* __restrict__ front, long long npx, int S,
                         const uint8_t* __restrict__ tab8, uint8_t* __restrict__ out_u8) {
    extern __shared__ __align__(16) float s_tab[];
    uint8_t* s8 = reinterpret_cast<uint8_t*>(s_tab);         // [65536]: entry (a, b) at a << 8 | b
    const long long gpi = npx / 16, total = gpi * S;
    const long long g_begin = total * blockIdx.x / gridDim.x, g_end = total * (blockIdx.x + 1) / gridDim.x;
    for (long long img = g_begin / gpi; img < S && img * gpi < g_end; ++img) {
        const long long lo = max(g_begin, img * gpi) - img * gpi, hi = min(g_end, (img + 1) * gpi) - img * gpi;
        __syncthreads();
        const uint4* src = reinterpret_cast<const uint4*>(tab8 + (static_cast<size_t>(img) << 16));
        for (int i = threadIdx.x; i < 65536 / 16; i += kTabThreads) reinterpret_cast<uint4*>(s8)[i] = __ldg(src + i);
        __syncthreads();
        const uint4* pa = reinterpret_cast<const uint4*>(now + static_cast<size_t>(img) * npx);
        const uint4* pb = reinterpret_cast<const uint4*>(front + static_cast<size_t>(img) * npx);
        uint4* ou = reinterpret_cast<uint4*>(out_u8 + static_cast<size_t>(img) * npx);
        for (long long gi = lo + threadIdx.x; gi < hi; gi += kTabThreads) {
            const uint4 va = __ldg(pa + gi), vb = __ldg(pb + gi);
            const unsigned wa[4] = {va.x, va.y, va.z, va.w}, wb[4] = {vb.x, vb.y, vb.z, vb.w};
            unsigned packed[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                // (now byte j) << 8 | (front byte j), one byte permute each
                const unsigned i0 = __byte_perm(wb[q], wa[q], 0x7740) & 0xffffu, i1 = __byte_perm(wb[q], wa[q], 0x7751) & 0xffffu;
                const unsigned i2 = __byte_perm(wb[q], wa[q], 0x7762) & 0xffffu, i3 = __byte_perm(wb[q], wa[q], 0x7773) & 0xffffu;
                packed[q] = static_cast<unsigned>(s8[i0]) | (static_cast<unsigned>(s8[i1]) << 8) |
                            (static_cast<unsigned>(s8[i2]) << 16) | (static_cast<unsigned>(s8[i3]) << 24);
            }
            ou[gi] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
    }
}

// ------------------------------------------------------------------ PIL 'L'
__global__ void rgb_to_gray_kernel(const uint8_t* __restrict__ rgb, long long n, uint8_t* __restrict__ gray) {
    // Pillow rgb2l: (19595 R + 38470 G + 7471 B + 0x8000) >> 16  (utils.py:126 convert('L'))
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(rgb) & 3) == 0) &&
                     ((reinterpret_cast<uintptr_t>(gray) & 3) == 0);
    if (vec) {
        for (long long k = i; k < n / 4; k += stride) {   // 4 pixels = 12 bytes in, 4 bytes out
            const unsigned* src = reinterpret_cast<const unsigned*>(rgb) + k * 3;
            const unsigned w0 = __ldg(src), w1 = __ldg(src + 1), w2 = __ldg(src + 2);
            const unsigned char bts[12] = {
                (unsigned char)(w0), (unsigned char)(w0 >> 8), (unsigned char)(w0 >> 16), (unsigned char)(w0 >> 24),
                (unsigned char)(w1), (unsigned char)(w1 >> 8), (unsigned char)(w1 >> 16), (unsigned char)(w1 >> 24),
                (unsigned char)(w2), (unsigned char)(w2 >> 8), (unsigned char)(w2 >> 16), (unsigned char)(w2 >> 24)};
            unsigned o = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const unsigned l = (19595u * bts[3 * j] + 38470u * bts[3 * j + 1] + 7471u * bts[3 * j + 2] + 32768u) >> 16;
                o |= l << (8 * j);
            }
            reinterpret_cast<unsigned*>(gray)[k] = o;
        }
    } else {
        for (long long k = i; k < n; k += stride) {
            const unsigned r = rgb[3 * k], g = rgb[3 * k + 1], b = rgb[3 * k + 2];
            gray[k] = static_cast<uint8_t>((19595u * r + 38470u * g + 7471u * b + 32768u) >> 16);
        }
    }
}

// ------------------------------------------------------------------ a9: normalised float image -> 'L' bytes
// dacs.py:730-733: clamp(denorm(img, mean, std), 0, 1) * 255 -> np.uint8 (truncation) -> PIL 'L'.  denorm is
// img.mul(std).add(mean) / 255.0 (mmseg/models/utils/dacs_transforms.py:52-53); every step one float32
// rounding, as torch evaluates it ON CUDA TENSORS (dacs.py:729): a tensor divided by a Python scalar is computed
// as a * fl(1 / 255) there (ATen's div kernel, CPU-scalar fast path), not as a true division.  d_img is [S, 3, H, W]; the gray plane (and optionally the HWC bytes PIL
// would have been handed) come out without the image ever leaving the device.
// fl(1 / 255) in float32, as torch computes the reciprocal of the scalar divisor
__device__ constexpr float kInv255 = 1.0f / 255.0f;
struct Denorm3 {
    float mean[3], std[3];
};
__global__ void __launch_bounds__(256)
denorm_to_gray_kernel(const float* __restrict__ img, long long npx, Denorm3 q, const float* __restrict__ d_mean,
                      const float* __restrict__ d_std, uint8_t* __restrict__ gray, uint8_t* __restrict__ rgb) {
    const int s = blockIdx.y;
    if (d_mean != nullptr) {          // constants that live in device memory (the reference's CUDA tensors)
#pragma unroll
        for (int c = 0; c < 3; ++c) { q.mean[c] = __ldg(d_mean + c); q.std[c] = __ldg(d_std + c); }
    }
    const float* base = img + static_cast<size_t>(s) * 3 * npx;
    for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < npx;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        unsigned c8[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float v = __fmul_rn(__fadd_rn(__fmul_rn(__ldg(base + c * npx + i), q.std[c]), q.mean[c]), kInv255);   // denorm
            v = fminf(fmaxf(v, 0.0f), 1.0f);                                                                    // torch.clamp(0, 1)
            c8[c] = static_cast<unsigned>(__float2int_rz(__fmul_rn(v, 255.0f))) & 255u;                           // * 255 -> np.uint8
        }
        gray[static_cast<size_t>(s) * npx + i] =
            static_cast<uint8_t>((19595u * c8[0] + 38470u * c8[1] + 7471u * c8[2] + 32768u) >> 16);              // convert('L')
        if (rgb != nullptr) {
            uint8_t* o = rgb + (static_cast<size_t>(s) * npx + i) * 3;
            o[0] = static_cast<uint8_t>(c8[0]); o[1] = static_cast<uint8_t>(c8[1]); o[2] = static_cast<uint8_t>(c8[2]);
        }
    }
}

static bool is_device_pointer(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { (void)cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

int launch_denorm_to_gray(const float* img, int S, int H, int W, const float* mean, const float* stdv, uint8_t* gray,
                          uint8_t* rgb, cudaStream_t st) {
    Denorm3 q = {};
    const bool dev_m = is_device_pointer(mean), dev_s = is_device_pointer(stdv);
    if (dev_m != dev_s) return CMDA_ERR_BAD_ARG;
    const float* d_mean = dev_m ? mean : nullptr;
    const float* d_std = dev_m ? stdv : nullptr;
    if (!dev_m)
        for (int c = 0; c < 3; ++c) { q.mean[c] = mean[c]; q.std[c] = stdv[c]; }
    const long long npx = static_cast<long long>(H) * W;
    long long gx = (npx + 255) / 256;
    const long long cap = (148LL * 8 + S - 1) / S;
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    for (int s0 = 0; s0 < S; s0 += 32768) {
        const int sn = (S - s0) < 32768 ? (S - s0) : 32768;
        denorm_to_gray_kernel<<<dim3(static_cast<unsigned>(gx), sn), 256, 0, st>>>(
            img + static_cast<size_t>(s0) * 3 * npx, npx, q, d_mean, d_std, gray + static_cast<size_t>(s0) * npx,
            rgb ? rgb + static_cast<size_t>(s0) * npx * 3 : nullptr);
        CMDA_LAUNCH_CHECK();
    }
    return CMDA_OK;
}

// ------------------------------------------------------------------ launchers
static int image_grid_x(long long work_items) {
    // 148 SMs x 8 resident 256-thread CTAs; never more blocks than work
    long long b = (work_items + kImgThreads - 1) / kImgThreads;
    const long long cap = 148LL * 8;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return static_cast<int>(b);
}

int launch_rgb_to_gray(const uint8_t* rgb, int64_t n, uint8_t* gray, cudaStream_t s) {
    if (n <= 0) return CMDA_OK;
    rgb_to_gray_kernel<<<image_grid_x(n / 4 + 1), kImgThreads, 0, s>>>(rgb, n, gray);
    CMDA_LAUNCH_CHECK();
    return CMDA_OK;
}

size_t image_slot_bytes(int S);
size_t image_table_bytes(int S, int H, int W);

int launch_isr(const uint8_t* gray, int S, int H, int W, int shift, int direction, const float* h_lut, float thr,
               float clip, float* out, unsigned* ws, size_t /*ws_bytes*/, cudaStream_t s) {
    LogLut lut;
    for (int i = 0; i < 256; ++i) lut.v[i] = h_lut[i];
    const TermList terms = terms_of(direction);
    const long long npx = static_cast<long long>(H) * W;
    CMDA_CUDA_TRY(cudaMemsetAsync(ws, 0, sizeof(unsigned) * 16 * S, s));
    if ((W % 4) == 0 && W >= 8 && (reinterpret_cast<uintptr_t>(gray) & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
        // vector fast path
        const int gx = (W / 4 + 255) / 256;
        const long long n_rows = static_cast<long long>(S) * H;
        long long gy = (148LL * 6 + gx - 1) / gx;          // 6 CTAs of 33 KB shared memory per SM
        if (gy > n_rows) gy = n_rows;
        dim3 grid(gx, static_cast<unsigned>(gy));
#define CMDA_ISR_VEC(D)                                                                                          \
    case D:                                                                                                      \
        isr_vec_kernel<D, false><<<grid, 256, 0, s>>>(gray, S, H, W, shift, lut, thr, clip, ws, out);             \
        isr_vec_kernel<D, true><<<grid, 256, 0, s>>>(gray, S, H, W, shift, lut, thr, clip, ws, out);              \
        break
        switch (direction) {
            CMDA_ISR_VEC(CMDA_DIR_RIGHTDOWN);
            CMDA_ISR_VEC(CMDA_DIR_RIGHTUP);
            CMDA_ISR_VEC(CMDA_DIR_LEFTDOWN);
            CMDA_ISR_VEC(CMDA_DIR_LEFTUP);
            CMDA_ISR_VEC(CMDA_DIR_ALL);
            default: return CMDA_ERR_BAD_ARG;
        }
#undef CMDA_ISR_VEC
        CMDA_LAUNCH_CHECK();
        return CMDA_OK;
    }
    // images per launch are bounded by gridDim.y
    for (int s0 = 0; s0 < S; s0 += 32768) {
        const int sn = (S - s0) < 32768 ? (S - s0) : 32768;
        const uint8_t* g = gray + static_cast<size_t>(s0) * npx;
        float* o = out + static_cast<size_t>(s0) * npx;
        unsigned* w = ws + static_cast<size_t>(s0) * 16;
        // split the 148x8 resident CTAs across the images of the batch
        int gx = image_grid_x(npx);
        int per_img = (148 * 8 + sn - 1) / sn;
        if (per_img < 1) per_img = 1;
        if (gx > per_img) gx = per_img;
        dim3 grid(gx, sn);
        int gxa = image_grid_x((npx + kPxPerThread - 1) / kPxPerThread);
        if (gxa > per_img) gxa = per_img;
        dim3 grida(gxa, sn);
        const bool vec = (npx % 4 == 0) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
        if (terms.n == 4) {
            isr_minmax_kernel<4><<<grid, kImgThreads, 0, s>>>(g, H, W, shift, terms, lut, thr, clip, w);
            if (vec) isr_apply_kernel<4, true><<<grida, kImgThreads, 0, s>>>(g, H, W, shift, terms, lut, thr, clip, w, o);
            else isr_apply_kernel<4, false><<<grida, kImgThreads, 0, s>>>(g, H, W, shift, terms, lut, thr, clip, w, o);
        } else {
            isr_minmax_kernel<2><<<grid, kImgThreads, 0, s>>>(g, H, W, shift, terms, lut, thr, clip, w);
            if (vec) isr_apply_kernel<2, true><<<grida, kImgThreads, 0, s>>>(g, H, W, shift, terms, lut, thr, clip, w, o);
            else isr_apply_kernel<2, false><<<grida, kImgThreads, 0, s>>>(g, H, W, shift, terms, lut, thr, clip, w, o);
        }
        CMDA_LAUNCH_CHECK();
    }
    return CMDA_OK;
}

// Workspace of the image entry points: 16 words of min / max slots per image, then (large images only) the 64 KB
// table of quantised values per image of the frame pair's table-driven uint8 pass.
size_t image_slot_bytes(int S) { return align_up(sizeof(unsigned) * 16 * static_cast<size_t>(S), 256); }
size_t image_table_bytes(int S, int H, int W) {
    if (static_cast<long long>(H) * W < kTableMinPixels) return 0;
    return static_cast<size_t>(S) * 65536;
}

int launch_pair(const uint8_t* now, const uint8_t* front, int S, int H, int W, const float* h_lut, float thr,
                float clip, float* out_f32, uint8_t* out_u8, unsigned* ws, size_t ws_bytes, cudaStream_t s) {
    LogLut lut;
    for (int i = 0; i < 256; ++i) lut.v[i] = h_lut[i];
    const long long npx = static_cast<long long>(H) * W;
    CMDA_CUDA_TRY(cudaMemsetAsync(ws, 0, sizeof(unsigned) * 16 * S, s));
    const bool all_vec = (npx % 16 == 0) && ((reinterpret_cast<uintptr_t>(now) & 15) == 0) && ((reinterpret_cast<uintptr_t>(front) & 15) == 0) &&
                         (!out_f32 || (reinterpret_cast<uintptr_t>(out_f32) & 15) == 0) && (!out_u8 || (reinterpret_cast<uintptr_t>(out_u8) & 15) == 0);
    const bool tables = all_vec && out_f32 == nullptr && S <= 32768 && image_table_bytes(S, H, W) != 0 &&
                        ws_bytes >= image_slot_bytes(S) + image_table_bytes(S, H, W);
    for (int s0 = 0; s0 < S; s0 += 32768) {
        const int sn = (S - s0) < 32768 ? (S - s0) : 32768;
        const size_t off = static_cast<size_t>(s0) * npx;
        int per_img = (148 * 8 + sn - 1) / sn;
        const bool vec = (npx % 16 == 0) && ((reinterpret_cast<uintptr_t>(now + off) & 15) == 0) &&
                         ((reinterpret_cast<uintptr_t>(front + off) & 15) == 0) &&
                         (!out_f32 || (reinterpret_cast<uintptr_t>(out_f32 + off) & 15) == 0) &&
                         (!out_u8 || (reinterpret_cast<uintptr_t>(out_u8 + off) & 15) == 0);
        int gx = image_grid_x(vec ? npx / 16 : npx);
        if (gx > per_img) gx = per_img;
        dim3 grid(gx, sn);
        unsigned* w = ws + static_cast<size_t>(s0) * 16;
        pair_minmax_kernel<<<grid, kImgThreads, 0, s>>>(now + off, front + off, npx, vec, lut, thr, clip, w);
        if (tables) {
            // uint8 output only: one evaluation per byte pair and image, then a gather per pixel
            uint8_t* tab8 = reinterpret_cast<uint8_t*>(ws) + image_slot_bytes(S);
            pair_table_kernel<<<dim3(256 / kTabRowsPerCta, sn), 256, 0, s>>>(lut, thr, clip, w, tab8);
            CMDA_CUDA_TRY(cudaFuncSetAttribute(pair_apply_table8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
            pair_apply_table8_kernel<<<148 * 2, kTabThreads, 65536, s>>>(now, front, npx, sn, tab8, out_u8);
        } else if (vec)
            pair_apply_kernel<true><<<grid, kImgThreads, 0, s>>>(now + off, front + off, npx, lut, thr, clip, w,
                                                                 out_f32 ? out_f32 + off : nullptr,
                                                                 out_u8 ? out_u8 + off : nullptr);
        else
            pair_apply_kernel<false><<<grid, kImgThreads, 0, s>>>(now + off, front + off, npx, lut, thr, clip, w,
                                                                  out_f32 ? out_f32 + off : nullptr,
                                                                  out_u8 ? out_u8 + off : nullptr);
        CMDA_LAUNCH_CHECK();
    }
    return CMDA_OK;
}

}  // namespace cmda
