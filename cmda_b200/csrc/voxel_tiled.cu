// K2 (mode TILED) -- tile-partitioned voxel scatter with shared-memory accumulation.
// Follows /root/reference/mmseg/datasets/dsec.py:26-58 and 341-357, same arithmetic as
// voxel_global.cu (event_math.cuh), different data movement:
//
//   bbox      per (window, raw tile): the rectified footprint of the tile's pixels, from the
//             window's rectify map (min/max of the map over the tile, truncated like .int()).
//   count     histogram of the window's events over raw-sensor tiles (reads x, y only).
//   scan      bucket offsets per (window, tile) + the work-item list (buckets cut into
//             sub-ranges of at most kItemRecords records, for load balance).
//   partition multisplit: every CTA sorts one chunk of events by tile in shared memory and
//             appends each tile's run to the tile's bucket as compact records (local pixel,
//             polarity and, for B > 1, the window-relative timestamp) -- coalesced runs.
//   accumulate persistent CTAs pull work items; each stages the tile's patch of the rectify
//             map in shared memory (the gather becomes an LDS), accumulates the corner weights
//             of its records into a shared-memory copy of the tile's footprint with native
//             32-bit integer ATOMS (2^-30 fixed point; the carries into the upper word are
//             tracked through the value the atomic returns), and flushes the non-zero voxels
//             to the global 64-bit grid with one RED each.
//
// The sums are exact integer sums of the same quantised weights as mode GLOBAL, so the two
// modes are bit-identical and the result does not depend on the order in which records land
// in a bucket (bucket order is the only non-deterministic thing here).
#include <type_traits>

#include "event_math.cuh"

namespace cmda {

constexpr int kCountThreads = 256;
constexpr int kCountGroupsPerThread = 16;                                  // 8 events per group
constexpr int kAccThreads = 512;
constexpr int kItemRecords = 32768;    // records per work item (load-balance granularity of the accumulate pass)
constexpr int kMaxTiles = 4096;

struct TileGeom {
    int tw_log2, th_log2;   // raw-sensor tile size (powers of two)
    int ntx, nty, T;        // tiles per row / column / total
    int cap_voxels;         // shared-memory accumulator capacity of one work item
};

struct TiledTable {
    long long rec_base[kMaxWindows];   // first record of each window in the record buffer
};

struct Rec8 {
    uint32_t dt;      // t - t[window start], microseconds
    uint32_t lxyp;    // local x | local y << 8 | polarity << 16
};
typedef uint16_t Rec2;   // local x | local y << 6 | polarity << 12  (B == 1: no time needed)

static TileGeom pick_geom(int H, int W, int B) {
    TileGeom g{};
    if (B == 1) { g.tw_log2 = 6; g.th_log2 = 5; g.cap_voxels = 6144; }
    else if (B <= 5) { g.tw_log2 = 5; g.th_log2 = 5; g.cap_voxels = 12288; }
    else if (B <= 10) { g.tw_log2 = 5; g.th_log2 = 4; g.cap_voxels = 12288; }
    else { g.tw_log2 = 4; g.th_log2 = 4; g.cap_voxels = 16384; }
    g.ntx = (W + (1 << g.tw_log2) - 1) >> g.tw_log2;
    g.nty = (H + (1 << g.th_log2) - 1) >> g.th_log2;
    g.T = g.ntx * g.nty;
    return g;
}

int tiled_supported(int H, int W, int B) {
    if (B < 1 || B > 24 || H < 1 || W < 1 || H > 32768 || W > 32768) return 0;
    const TileGeom g = pick_geom(H, W, B);
    return g.T <= kMaxTiles;
}

// ---- workspace carving (everything the TILED passes need besides the int64 grid) -------------
struct TiledScratch {
    unsigned* counts;       // [S][T]   events per (window, tile)            (zeroed per call)
    unsigned* cursor;       // [S][T]   append cursors of the partition pass (zeroed per call)
    unsigned* queue;        // [64]     [0] = number of work items, [1] = next item (zeroed per call)
    unsigned* bucket_off;   // [S][T]   first record of the bucket, relative to the window's records
    int4* bbox;             // [S][T]   x0, y0, w, h of the tile's rectified footprint
    int4* items;            // [max_items]  window, tile, first record, one past the last record
    void* records;
    size_t control_bytes;   // counts + cursor + queue (contiguous, one memset)
    size_t total_bytes;
    int max_items;
};

static TiledScratch carve(char* base, long long total_events, int S, const TileGeom& g, int B) {
    TiledScratch z{};
    const size_t st = static_cast<size_t>(S) * g.T;
    size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += align_up(bytes, 256); return p; };
    z.counts = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * st));
    z.cursor = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * st));
    z.queue = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * 64));
    z.control_bytes = off;
    z.bucket_off = reinterpret_cast<unsigned*>(take(sizeof(unsigned) * st));
    z.bbox = reinterpret_cast<int4*>(take(sizeof(int4) * st));
    z.max_items = static_cast<int>(st + static_cast<size_t>(total_events / kItemRecords) + 1);
    z.items = reinterpret_cast<int4*>(take(sizeof(int4) * z.max_items));
    const size_t rec = (B == 1) ? sizeof(Rec2) : sizeof(Rec8);
    z.records = take(rec * static_cast<size_t>(total_events > 0 ? total_events : 1) + 64);
    z.total_bytes = off;
    return z;
}

size_t tiled_workspace_bytes(int64_t total_events, int S, int H, int W, int B) {
    if (!tiled_supported(H, W, B)) return 0;
    const TileGeom g = pick_geom(H, W, B);
    const int group = S < kMaxWindows ? S : kMaxWindows;
    return carve(nullptr, total_events, group, g, B).total_bytes;
}

// ---- bbox ---------------------------------------------------------------------------------
// Corner range of one axis from the extreme map values of the tile: corners are
// trunc(v) and trunc(v) + 1 (dsec.py:41-49); trunc is monotone, so the range of corners is
// [trunc(min), trunc(max) + 1] clipped to the grid.
__device__ __forceinline__ void corner_range(float vmin, float vmax, int size, int& lo, int& len) {
    if (!(vmin <= vmax)) { lo = 0; len = 0; return; }           // no finite map value in the tile
    const int a = (vmin >= 0.0f) ? ((vmin < static_cast<float>(size)) ? __float2int_rz(vmin) : size) : 0;
    const int b = (vmax < -1.0f) ? -1 : ((vmax >= static_cast<float>(size - 1)) ? size - 1 : __float2int_rz(vmax) + 1);
    lo = a;
    len = b >= a ? b - a + 1 : 0;
}

__global__ void __launch_bounds__(256)
tile_bbox_kernel(WindowTable tab, const float2* __restrict__ maps, int H, int W, TileGeom g, int4* __restrict__ bbox) {
    const int k = blockIdx.x, s = blockIdx.y;
    const int TW = 1 << g.tw_log2, TH = 1 << g.th_log2;
    const int ox = (k % g.ntx) << g.tw_log2, oy = (k / g.ntx) << g.th_log2;
    float xmin = INFINITY, xmax = -INFINITY, ymin = INFINITY, ymax = -INFINITY;
    if (maps != nullptr) {
        const float2* map = maps + static_cast<size_t>(tab.w[s].map_id) * H * W;
        for (int i = threadIdx.x; i < TW * TH; i += blockDim.x) {
            const int gx = ox + (i & (TW - 1)), gy = oy + (i >> g.tw_log2);
            if (gx < W && gy < H) {
                const float2 m = __ldg(map + static_cast<size_t>(gy) * W + gx);
                xmin = fminf(xmin, m.x); xmax = fmaxf(xmax, m.x);      // fminf / fmaxf skip NaN
                ymin = fminf(ymin, m.y); ymax = fmaxf(ymax, m.y);
            }
        }
    } else if (threadIdx.x == 0) {   // no remap: x = float(x), y = float(y)
        xmin = static_cast<float>(ox); xmax = static_cast<float>(min(ox + TW, W) - 1);
        ymin = static_cast<float>(oy); ymax = static_cast<float>(min(oy + TH, H) - 1);
    }
    xmin = warp_min(xmin); xmax = warp_max(xmax); ymin = warp_min(ymin); ymax = warp_max(ymax);
    __shared__ float s_v[4][8];
    const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { s_v[0][wid] = xmin; s_v[1][wid] = xmax; s_v[2][wid] = ymin; s_v[3][wid] = ymax; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            xmin = fminf(xmin, s_v[0][w]); xmax = fmaxf(xmax, s_v[1][w]);
            ymin = fminf(ymin, s_v[2][w]); ymax = fmaxf(ymax, s_v[3][w]);
        }
        int4 bb;
        corner_range(xmin, xmax, W, bb.x, bb.z);
        corner_range(ymin, ymax, H, bb.y, bb.w);
        bbox[static_cast<size_t>(s) * g.T + k] = bb;
    }
}

// ---- vector access to 8 consecutive events (global index multiple of 8) ---------------------
struct XY8 { uint4 x, y; };
__device__ __forceinline__ unsigned u16_of(const uint4& v, int j) {
    const unsigned w = (j < 2) ? v.x : (j < 4) ? v.y : (j < 6) ? v.z : v.w;
    return (j & 1) ? (w >> 16) : (w & 0xffffu);
}
template <bool VEC>
__device__ __forceinline__ XY8 load_xy8(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, long long i0,
                                        long long lo, long long hi) {
    XY8 r;
    if (VEC && i0 >= lo && i0 + 8 <= hi) {
        r.x = ldg_stream_u4(x + i0);
        r.y = ldg_stream_u4(y + i0);
    } else {
        unsigned vx[8], vy[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool in = (i0 + j >= lo) && (i0 + j < hi);
            vx[j] = in ? __ldg(x + i0 + j) : 0xffffu;       // 0xffff is outside any sensor: dropped
            vy[j] = in ? __ldg(y + i0 + j) : 0xffffu;
        }
        r.x = make_uint4(vx[0] | (vx[1] << 16), vx[2] | (vx[3] << 16), vx[4] | (vx[5] << 16), vx[6] | (vx[7] << 16));
        r.y = make_uint4(vy[0] | (vy[1] << 16), vy[2] | (vy[3] << 16), vy[4] | (vy[5] << 16), vy[6] | (vy[7] << 16));
    }
    return r;
}

// ---- count ----------------------------------------------------------------------------------
template <bool VEC>
__global__ void __launch_bounds__(kCountThreads)
tile_count_kernel(const uint16_t* __restrict__ x, const uint16_t* __restrict__ y, WindowTable tab, TileGeom g, int H,
                  int W, unsigned* __restrict__ counts) {
    extern __shared__ unsigned s_hist[];
    const int s = blockIdx.y;
    const WindowDesc wd = tab.w[s];
    const long long g0 = wd.start >> 3, g1 = (wd.end + 7) >> 3;          // groups of 8 events
    const long long first = g0 + static_cast<long long>(blockIdx.x) * (kCountThreads * kCountGroupsPerThread);
    if (wd.end <= wd.start || first >= g1) return;
    for (int k = threadIdx.x; k < g.T; k += kCountThreads) s_hist[k] = 0u;
    __syncthreads();
#pragma unroll 4
    for (int j = 0; j < kCountGroupsPerThread; ++j) {
        const long long grp = first + static_cast<long long>(j) * kCountThreads + threadIdx.x;
        if (grp >= g1) break;
        const XY8 v = load_xy8<VEC>(x, y, grp << 3, wd.start, wd.end);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const unsigned ex = u16_of(v.x, e), ey = u16_of(v.y, e);
            if (ex < static_cast<unsigned>(W) && ey < static_cast<unsigned>(H))
                atomicAdd(&s_hist[(ey >> g.th_log2) * g.ntx + (ex >> g.tw_log2)], 1u);
        }
    }
    __syncthreads();
    unsigned* dst = counts + static_cast<size_t>(s) * g.T;
    for (int k = threadIdx.x; k < g.T; k += kCountThreads) {
        const unsigned c = s_hist[k];
        if (c) atomicAdd(dst + k, c);
    }
}

// ---- scan: bucket offsets + work items --------------------------------------------------------
__global__ void __launch_bounds__(1024)
tile_scan_kernel(const unsigned* __restrict__ counts, TileGeom g, unsigned* __restrict__ bucket_off,
                 int4* __restrict__ items, unsigned* __restrict__ queue, int max_items) {
    constexpr int PER = kMaxTiles / 1024;
    __shared__ unsigned s_warp[2][32];
    __shared__ unsigned s_item_base;
    const int s = blockIdx.x;
    const unsigned* c = counts + static_cast<size_t>(s) * g.T;
    unsigned cnt[PER], sub[PER];
    unsigned tc = 0, ts = 0;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int k = threadIdx.x * PER + j;
        cnt[j] = (k < g.T) ? c[k] : 0u;
        sub[j] = (cnt[j] + kItemRecords - 1) / kItemRecords;
        tc += cnt[j]; ts += sub[j];
    }
    // block-wide exclusive scan of (tc, ts)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned ic = tc, is = ts;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned a = __shfl_up_sync(0xffffffffu, ic, o), b = __shfl_up_sync(0xffffffffu, is, o);
        if (lane >= o) { ic += a; is += b; }
    }
    if (lane == 31) { s_warp[0][wid] = ic; s_warp[1][wid] = is; }
    __syncthreads();
    if (wid == 0) {
        unsigned a = s_warp[0][lane], b = s_warp[1][lane];
        unsigned ia = a, ib = b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, ia, o), v = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) { ia += u; ib += v; }
        }
        s_warp[0][lane] = ia - a; s_warp[1][lane] = ib - b;
        if (lane == 31) s_item_base = atomicAdd(&queue[0], ib);     // reserve this window's items
    }
    __syncthreads();
    unsigned oc = s_warp[0][wid] + ic - tc;          // exclusive prefix of records
    unsigned os = s_item_base + s_warp[1][wid] + is - ts;
#pragma unroll
    for (int j = 0; j < PER; ++j) {
        const int k = threadIdx.x * PER + j;
        if (k < g.T) {
            bucket_off[static_cast<size_t>(s) * g.T + k] = oc;
            for (unsigned q = 0; q < sub[j]; ++q) {
                const unsigned b = oc + q * kItemRecords;
                const unsigned e = min(oc + cnt[j], b + kItemRecords);
                if (os + q < static_cast<unsigned>(max_items)) items[os + q] = make_int4(s, k, static_cast<int>(b), static_cast<int>(e));
            }
            oc += cnt[j]; os += sub[j];
        }
    }
}

// ---- partition --------------------------------------------------------------------------------
#ifndef CMDA_PART_THREADS
#define CMDA_PART_THREADS 512
#endif
#ifndef CMDA_PART_GROUPS
#define CMDA_PART_GROUPS 1
#endif
#ifndef CMDA_PART_MINBLOCKS
#define CMDA_PART_MINBLOCKS 2
#endif
constexpr int kPartThreads = CMDA_PART_THREADS;
constexpr int kPartGroupsPerThread = CMDA_PART_GROUPS;                     // 8 events per group
constexpr int kPartChunk = kPartThreads * kPartGroupsPerThread * 8;        // events per CTA

// 8 consecutive events of one window in registers (vector loads when aligned and interior).
struct Ev8 {
    uint4 x, y, t0, t1;
    uint2 p;
};
template <bool HAS_T, bool VEC>
__device__ __forceinline__ Ev8 load_ev8(const uint32_t* __restrict__ t, const uint16_t* __restrict__ x,
                                        const uint16_t* __restrict__ y, const uint8_t* __restrict__ p, long long i0,
                                        long long lo, long long hi) {
    Ev8 r;
    if (VEC && i0 >= lo && i0 + 8 <= hi) {
        r.x = ldg_stream_u4(x + i0);
        r.y = ldg_stream_u4(y + i0);
        if (HAS_T) { r.t0 = ldg_stream_u4(t + i0); r.t1 = ldg_stream_u4(t + i0 + 4); }
        else { r.t0 = make_uint4(0, 0, 0, 0); r.t1 = r.t0; }
        r.p = ldg_stream_u2(p + i0);
    } else {
        unsigned vx[8], vy[8], tv[8];
        unsigned long long pv = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const bool in = (i0 + j >= lo) && (i0 + j < hi);
            vx[j] = in ? __ldg(x + i0 + j) : 0xffffu;       // 0xffff is outside any sensor: dropped
            vy[j] = in ? __ldg(y + i0 + j) : 0xffffu;
            tv[j] = (HAS_T && in) ? __ldg(t + i0 + j) : 0u;
            if (in) pv |= static_cast<unsigned long long>(__ldg(p + i0 + j)) << (8 * j);
        }
        r.x = make_uint4(vx[0] | (vx[1] << 16), vx[2] | (vx[3] << 16), vx[4] | (vx[5] << 16), vx[6] | (vx[7] << 16));
        r.y = make_uint4(vy[0] | (vy[1] << 16), vy[2] | (vy[3] << 16), vy[4] | (vy[5] << 16), vy[6] | (vy[7] << 16));
        r.t0 = make_uint4(tv[0], tv[1], tv[2], tv[3]);
        r.t1 = make_uint4(tv[4], tv[5], tv[6], tv[7]);
        r.p = make_uint2(static_cast<unsigned>(pv), static_cast<unsigned>(pv >> 32));
    }
    return r;
}

template <bool HAS_T, bool VEC>
__global__ void __launch_bounds__(kPartThreads, CMDA_PART_MINBLOCKS)
tile_partition_kernel(const uint32_t* __restrict__ t, const uint16_t* __restrict__ x, const uint16_t* __restrict__ y,
                      const uint8_t* __restrict__ p, WindowTable tab, TiledTable tt, TileGeom g, int H, int W,
                      const unsigned* __restrict__ bucket_off, unsigned* __restrict__ cursor, void* __restrict__ records) {
    using Rec = typename std::conditional<HAS_T, Rec8, Rec2>::type;
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned* s_hist = reinterpret_cast<unsigned*>(s_raw);     // [T] counts, later the run's global base
    unsigned* s_loff = s_hist + g.T;                            // [T + 1] local exclusive offsets
    unsigned* s_dst = s_loff + g.T + 2;                         // [kPartChunk] destination record index
    Rec* s_stage = reinterpret_cast<Rec*>(s_dst + kPartChunk);  // [kPartChunk] records sorted by tile
    __shared__ unsigned s_warp[kPartThreads / 32];

    const int s = blockIdx.y;
    const WindowDesc wd = tab.w[s];
    const long long g0 = wd.start >> 3, g1 = (wd.end + 7) >> 3;
    const long long first = g0 + static_cast<long long>(blockIdx.x) * (kPartThreads * kPartGroupsPerThread);
    if (wd.end <= wd.start || first >= g1) return;

    // every load of the chunk is issued before anything waits on one
    Ev8 ev[kPartGroupsPerThread];
#pragma unroll
    for (int j = 0; j < kPartGroupsPerThread; ++j) {
        const long long grp = first + static_cast<long long>(j) * kPartThreads + threadIdx.x;
        if (grp < g1) {
            ev[j] = load_ev8<HAS_T, VEC>(t, x, y, p, grp << 3, wd.start, wd.end);
        } else {
            ev[j].x = make_uint4(~0u, ~0u, ~0u, ~0u);
            ev[j].y = ev[j].x; ev[j].t0 = ev[j].x; ev[j].t1 = ev[j].x; ev[j].p = make_uint2(0u, 0u);
        }
    }
    const uint32_t t_first = HAS_T ? __ldg(t + wd.start) : 0u;
    for (int k = threadIdx.x; k < g.T; k += kPartThreads) s_hist[k] = 0u;
    __syncthreads();

    const unsigned tile_mask_x = (1u << g.tw_log2) - 1u, tile_mask_y = (1u << g.th_log2) - 1u;
    unsigned slot[kPartGroupsPerThread][8];      // tile << 16 | rank   (0xffffffff: dropped)
    unsigned lxyp[kPartGroupsPerThread][8];      // the record's pixel / polarity word
#pragma unroll
    for (int j = 0; j < kPartGroupsPerThread; ++j) {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const unsigned ex = u16_of(ev[j].x, e), ey = u16_of(ev[j].y, e);
            // reference dsec.py:349 casts p to float32 and 2*p-1 follows; DSEC stores 0 / 1
            const unsigned pol = ((e < 4 ? ev[j].p.x : ev[j].p.y) >> (8 * (e & 3))) & 0xffu;
            const unsigned lx = ex & tile_mask_x, ly = ey & tile_mask_y;
            lxyp[j][e] = HAS_T ? (lx | (ly << 8) | (pol << 16)) : (lx | (ly << 6) | ((pol & 15u) << 12));
            slot[j][e] = 0xffffffffu;
            if (ex < static_cast<unsigned>(W) && ey < static_cast<unsigned>(H)) {
                const unsigned tile = (ey >> g.th_log2) * g.ntx + (ex >> g.tw_log2);
                slot[j][e] = (tile << 16) | atomicAdd(&s_hist[tile], 1u);
            }
        }
    }
    __syncthreads();
    // exclusive scan of the tile histogram (each thread owns a contiguous run of tiles)
    const int per = (g.T + kPartThreads - 1) / kPartThreads;
    unsigned mine = 0;
    for (int j = 0; j < per; ++j) {
        const int k = threadIdx.x * per + j;
        if (k < g.T) mine += s_hist[k];
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned a = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += a;
    }
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const unsigned a = (lane < kPartThreads / 32) ? s_warp[lane] : 0u;
        unsigned ia = a;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned u = __shfl_up_sync(0xffffffffu, ia, o);
            if (lane >= o) ia += u;
        }
        if (lane < kPartThreads / 32) s_warp[lane] = ia - a;
    }
    __syncthreads();
    unsigned run = s_warp[wid] + inc - mine;
    const unsigned* boff = bucket_off + static_cast<size_t>(s) * g.T;
    unsigned* cur = cursor + static_cast<size_t>(s) * g.T;
    for (int j = 0; j < per; ++j) {
        const int k = threadIdx.x * per + j;
        if (k < g.T) {
            const unsigned c = s_hist[k];
            s_loff[k] = run;
            run += c;
            // append position of this chunk's run inside the tile's bucket
            s_hist[k] = c ? (boff[k] + atomicAdd(cur + k, c)) : 0u;
            if (k == g.T - 1) s_loff[g.T] = run;
        }
    }
    __syncthreads();
    // stage the records sorted by tile
#pragma unroll
    for (int j = 0; j < kPartGroupsPerThread; ++j) {
        const unsigned tv[8] = {ev[j].t0.x, ev[j].t0.y, ev[j].t0.z, ev[j].t0.w,
                                ev[j].t1.x, ev[j].t1.y, ev[j].t1.z, ev[j].t1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const unsigned sl = slot[j][e];
            if (sl == 0xffffffffu) continue;
            const unsigned tile = sl >> 16, rank = sl & 0xffffu;
            Rec r;
            if constexpr (HAS_T) {
                r.dt = tv[e] - t_first;
                r.lxyp = lxyp[j][e];
            } else {
                r = static_cast<Rec2>(lxyp[j][e]);
            }
            const unsigned pos = s_loff[tile] + rank;
            s_stage[pos] = r;
            s_dst[pos] = s_hist[tile] + rank;      // bucket base of this chunk's run + rank inside the run
        }
    }
    __syncthreads();
    // copy out: consecutive staged records of one tile go to consecutive bucket slots, so a
    // warp's stores coalesce into the tile runs
    Rec* out = reinterpret_cast<Rec*>(records) + tt.rec_base[s];
    const unsigned total = s_loff[g.T];
    for (unsigned i = threadIdx.x; i < total; i += kPartThreads) out[s_dst[i]] = s_stage[i];
}

// ---- accumulate -------------------------------------------------------------------------------
// One contribution into the shared-memory footprint: a native 32-bit ATOMS.ADD on the low
// word; the value it returns tells whether this very addition carried (or borrowed) across
// bit 32, and only then (about 1 contribution in 16) a predicated RED updates the upper word.
// (lo, hi) is therefore the exact 64-bit sum, like the RED.64 of mode GLOBAL.  Straight-line
// PTX: the compiler would otherwise wrap each of the 8 corners in divergence regions.
// `addr` is the shared-window byte address of the low word, XOFF selects the x + 1 neighbour,
// `hi_off` the byte distance from the low to the high array.
template <int XOFF>
__device__ __forceinline__ void smem_accumulate_at(unsigned addr, unsigned hi_off, int q) {
    unsigned old;
    if constexpr (XOFF == 0) asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(q) : "memory");
    else asm volatile("atom.shared.add.u32 %0, [%1+4], %2;" : "=r"(old) : "r"(addr), "r"(q) : "memory");
    const unsigned nw = old + static_cast<unsigned>(q);
    const int d = static_cast<int>(nw < old) + (q >> 31);      // carry out of the unsigned add, minus 1 for a negative addend
    const unsigned haddr = addr + hi_off;
    if constexpr (XOFF == 0)
        asm volatile("{ .reg .pred p; setp.ne.s32 p, %1, 0; @p red.shared.add.s32 [%0], %1; }" ::"r"(haddr), "r"(d) : "memory");
    else
        asm volatile("{ .reg .pred p; setp.ne.s32 p, %1, 0; @p red.shared.add.s32 [%0+4], %1; }" ::"r"(haddr), "r"(d) : "memory");
}
__device__ __forceinline__ void smem_accumulate(unsigned* __restrict__ lo, int* __restrict__ hi, unsigned v, int q) {
    const unsigned addr = static_cast<unsigned>(__cvta_generic_to_shared(lo + v));
    const unsigned hi_off = static_cast<unsigned>(__cvta_generic_to_shared(hi)) - static_cast<unsigned>(__cvta_generic_to_shared(lo));
    smem_accumulate_at<0>(addr, hi_off, q);
}

struct AccRecord {
    unsigned a, b;
};
template <bool HAS_T>
__device__ __forceinline__ AccRecord load_record(const void* __restrict__ rec, unsigned i) {
    AccRecord r;
    if constexpr (HAS_T) {
        const uint2 v = ldg_stream_u2(reinterpret_cast<const Rec8*>(rec) + i);
        r.a = v.x; r.b = v.y;
    } else {
        r.a = 0u;
        r.b = __ldg(reinterpret_cast<const unsigned short*>(rec) + i);
    }
    return r;
}

template <bool HAS_T>
__global__ void __launch_bounds__(kAccThreads)
tile_accumulate_kernel(const void* __restrict__ records, const uint32_t* __restrict__ t, WindowTable tab, TiledTable tt,
                       TileGeom g, const float2* __restrict__ maps, int H, int W, int B,
                       const int4* __restrict__ bbox, const int4* __restrict__ items, unsigned* __restrict__ queue,
                       unsigned long long* __restrict__ acc, unsigned long long* __restrict__ bin_counts) {
    using Rec = typename std::conditional<HAS_T, Rec8, Rec2>::type;
    extern __shared__ __align__(16) unsigned char s_raw[];
    const int TW = 1 << g.tw_log2, TH = 1 << g.th_log2;
    const unsigned cap = static_cast<unsigned>(g.cap_voxels);
    unsigned* s_lo = reinterpret_cast<unsigned*>(s_raw);                        // [cap]
    int* s_hi = reinterpret_cast<int*>(s_lo + cap);                             // [cap]
    float2* s_map = reinterpret_cast<float2*>(s_hi + cap);                      // [TW * TH]
    unsigned* s_bins = reinterpret_cast<unsigned*>(s_map + TW * TH);            // [32]
    __shared__ unsigned s_item;

    for (unsigned i = threadIdx.x; i < cap; i += kAccThreads) { s_lo[i] = 0u; s_hi[i] = 0; }
    if (threadIdx.x < 32) s_bins[threadIdx.x] = 0u;
    const unsigned n_items = queue[0];
    const size_t V = static_cast<size_t>(B) * H * W;
    const unsigned lo_base = static_cast<unsigned>(__cvta_generic_to_shared(s_lo));
    const unsigned hi_off = cap * 4u;

    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(&queue[1], 1u);
        __syncthreads();
        const unsigned item = s_item;
        if (item >= n_items) break;
        const int4 it = __ldg(items + item);
        const int s = it.x, k = it.y;
        const WindowDesc wd = tab.w[s];
        const int4 bb = __ldg(bbox + static_cast<size_t>(s) * g.T + k);
        const int ox = (k % g.ntx) << g.tw_log2, oy = (k / g.ntx) << g.th_log2;
        const float2* map = maps ? maps + static_cast<size_t>(wd.map_id) * H * W : nullptr;
        // stage the tile's patch of the rectify map (dsec.py:351: rectify_map[y, x])
        for (int i = threadIdx.x; i < TW * TH; i += kAccThreads) {
            const int gx = ox + (i & (TW - 1)), gy = oy + (i >> g.tw_log2);
            float2 m = make_float2(static_cast<float>(gx), static_cast<float>(gy));
            if (map != nullptr && gx < W && gy < H) m = __ldg(map + static_cast<size_t>(gy) * W + gx);
            s_map[i] = m;
        }
        const RawWindowTime rw = raw_window_time(t, wd.start, wd.end, B);
        // den is 1 (then t01[0] is 0 and t_norm = (C-1) * (dt / dT), dsec.py:38-39 after 347-348)
        // or NaN (single-timestamp window: every t_norm is NaN, nothing is accumulated or counted)
        const bool window_nan = !(rw.den == 1.0f);
        const Rec* rec = reinterpret_cast<const Rec*>(records) + tt.rec_base[s];
        unsigned long long* gacc = acc + static_cast<size_t>(s) * V;
        const unsigned bw = static_cast<unsigned>(bb.z), bh = static_cast<unsigned>(bb.w);
        const unsigned plane = bw * bh;
        const bool fits = plane * static_cast<unsigned>(B) <= cap;
        const unsigned end = static_cast<unsigned>(it.w);
        unsigned i = static_cast<unsigned>(it.z) + threadIdx.x;
        AccRecord cur{0u, 0u};
        if (!window_nan && i < end) cur = load_record<HAS_T>(rec, i);
        __syncthreads();
        while (!window_nan && i < end) {
            const unsigned inext = i + kAccThreads;
            AccRecord nxt{0u, 0u};
            if (inext < end) nxt = load_record<HAS_T>(rec, inext);      // in flight while this record is processed
            unsigned lx, ly, pol;
            if constexpr (HAS_T) { lx = cur.b & 0xffu; ly = (cur.b >> 8) & 0xffu; pol = cur.b >> 16; }
            else { lx = cur.b & 63u; ly = (cur.b >> 6) & 63u; pol = cur.b >> 12; }
            const float2 m = s_map[(ly << g.tw_log2) + lx];
            const float value = __fsub_rn(__fmul_rn(2.0f, static_cast<float>(pol)), 1.0f);     // dsec.py:45
            // dsec.py:347-348 then 38-39 with t01[0] = 0 and den = 1 (both exact identities)
            const float tn = HAS_T ? __fmul_rn(rw.cm1, __fdiv_rn(__uint2float_rn(cur.a), rw.fdT)) : 0.0f;
            const int x0 = trunc_like_x86(m.x), y0 = trunc_like_x86(m.y);                       // dsec.py:41-42
            const int t0 = HAS_T ? trunc_like_x86(tn) : 0;                                      // dsec.py:43
            if (bin_counts != nullptr && t0 >= 0 && t0 < B) atomicAdd(&s_bins[t0], 1u);
            const bool interior = fits && x0 >= 0 && x0 < W - 1 && y0 >= 0 && y0 < H - 1 &&
                                  (!HAS_T || (t0 >= 0 && t0 < B - 1));
            if (interior) {
                // all corners are inside the grid and inside the footprint: no per-corner tests.
                // Weights are carried pre-scaled by 2^30 (exact: a power of two commutes with
                // every rounding of the left-to-right product of dsec.py:51-52).
                const float v30 = __fmul_rn(value, kFixScale);
                const float vx0 = __fmul_rn(v30, tent(x0, m.x)), vx1 = __fmul_rn(v30, tent(x0 + 1, m.x));
                const float wy0 = tent(y0, m.y), wy1 = tent(y0 + 1, m.y);
                const float w00 = __fmul_rn(vx0, wy0), w01 = __fmul_rn(vx0, wy1);
                const float w10 = __fmul_rn(vx1, wy0), w11 = __fmul_rn(vx1, wy1);
                const unsigned v = (static_cast<unsigned>(t0) * bh + static_cast<unsigned>(y0 - bb.y)) * bw +
                                   static_cast<unsigned>(x0 - bb.x);
                const unsigned a00 = lo_base + v * 4u, a01 = a00 + bw * 4u;     // (y0, t0), (y0 + 1, t0)
                if constexpr (HAS_T) {
                    const float wt0 = tent(t0, tn), wt1 = tent(t0 + 1, tn);
                    const unsigned a10 = a00 + plane * 4u, a11 = a01 + plane * 4u;   // bin t0 + 1
                    smem_accumulate_at<0>(a00, hi_off, __float2int_rn(__fmul_rn(w00, wt0)));
                    smem_accumulate_at<0>(a10, hi_off, __float2int_rn(__fmul_rn(w00, wt1)));
                    smem_accumulate_at<0>(a01, hi_off, __float2int_rn(__fmul_rn(w01, wt0)));
                    smem_accumulate_at<0>(a11, hi_off, __float2int_rn(__fmul_rn(w01, wt1)));
                    smem_accumulate_at<4>(a00, hi_off, __float2int_rn(__fmul_rn(w10, wt0)));
                    smem_accumulate_at<4>(a10, hi_off, __float2int_rn(__fmul_rn(w10, wt1)));
                    smem_accumulate_at<4>(a01, hi_off, __float2int_rn(__fmul_rn(w11, wt0)));
                    smem_accumulate_at<4>(a11, hi_off, __float2int_rn(__fmul_rn(w11, wt1)));
                } else {
                    // B == 1: t_norm = 0, the only temporal corner is bin 0 with weight 1
                    smem_accumulate_at<0>(a00, hi_off, __float2int_rn(w00));
                    smem_accumulate_at<0>(a01, hi_off, __float2int_rn(w01));
                    smem_accumulate_at<4>(a00, hi_off, __float2int_rn(w10));
                    smem_accumulate_at<4>(a01, hi_off, __float2int_rn(w11));
                }
            } else {
                // border events (some corner outside the grid) and over-capacity footprints
                Event e;
                e.x = m.x; e.y = m.y; e.tn = tn; e.value = value;
                const Origin o = origin_of(e, H, W, B);
                if (o.any) {
                    for_each_corner(e, o, H, W, B, [&](int xl, int yl, int tl, float w) {
                        const int q = __float2int_rn(__fmul_rn(w, kFixScale));     // |w| <= 1: fits 32 bits
                        if (q == 0) return;
                        const unsigned cx = static_cast<unsigned>(xl - bb.x), cy = static_cast<unsigned>(yl - bb.y);
                        const unsigned vv = (static_cast<unsigned>(tl) * bh + cy) * bw + cx;
                        if (cx < bw && cy < bh && vv < cap) smem_accumulate(s_lo, s_hi, vv, q);
                        else atomicAdd(gacc + (static_cast<size_t>(tl) * H + yl) * W + xl,
                                       static_cast<unsigned long long>(static_cast<long long>(q)));
                    });
                }
            }
            cur = nxt;
            i = inext;
        }
        __syncthreads();
        // flush the non-zero voxels of the footprint and restore the all-zero state
        const unsigned nvox = min(plane * static_cast<unsigned>(B), cap);
        for (unsigned v = threadIdx.x; v < nvox; v += kAccThreads) {
            const unsigned l = s_lo[v];
            const int h = s_hi[v];
            if (l == 0u && h == 0) continue;
            s_lo[v] = 0u;
            s_hi[v] = 0;
            const long long val = (static_cast<long long>(h) << 32) + static_cast<long long>(l);
            const unsigned cx = v % bw, r = v / bw;
            const unsigned cy = r % bh, tl = r / bh;
            atomicAdd(gacc + (static_cast<size_t>(tl) * H + (bb.y + cy)) * W + (bb.x + cx),
                      static_cast<unsigned long long>(val));
        }
        if (bin_counts != nullptr && threadIdx.x < B) {
            const unsigned c = s_bins[threadIdx.x];
            if (c) {
                atomicAdd(bin_counts + static_cast<size_t>(s) * B + threadIdx.x, static_cast<unsigned long long>(c));
                s_bins[threadIdx.x] = 0u;
            }
        }
    }
}

// ---- launch sequence --------------------------------------------------------------------------
template <typename K>
static int set_smem(K kernel, size_t bytes) {
    CMDA_CUDA_TRY(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(bytes)));
    return CMDA_OK;
}

int launch_tiled_scatter(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p,
                         const WindowTable& tab, int S, const float* maps, int H, int W, int B, long long* acc,
                         int64_t* bin_counts, void* scratch, size_t scratch_bytes, cudaStream_t st) {
    if (!tiled_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    const TileGeom g = pick_geom(H, W, B);
    long long total = 0, max_events = 0;
    TiledTable tt{};
    for (int s = 0; s < S; ++s) {
        const long long n = tab.w[s].end - tab.w[s].start;
        tt.rec_base[s] = total;
        if (n > 0xfffffff0LL) return CMDA_ERR_UNSUPPORTED;     // 32-bit record offsets inside a window
        total += n > 0 ? n : 0;
        if (n > max_events) max_events = n;
    }
    if (total == 0) return CMDA_OK;
    const TiledScratch z = carve(static_cast<char*>(scratch), total, S, g, B);
    if (z.total_bytes > scratch_bytes) return CMDA_ERR_WORKSPACE;
    const bool vec = ((reinterpret_cast<uintptr_t>(t) & 15) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p) & 7) == 0);
    const float2* maps2 = reinterpret_cast<const float2*>(maps);

    CMDA_CUDA_TRY(cudaMemsetAsync(z.counts, 0, z.control_bytes, st));
    tile_bbox_kernel<<<dim3(g.T, S), 256, 0, st>>>(tab, maps2, H, W, g, z.bbox);
    CMDA_LAUNCH_CHECK();
    {
        const long long groups = (max_events + 7) / 8 + 1;
        const long long per = static_cast<long long>(kCountThreads) * kCountGroupsPerThread;
        dim3 grid(static_cast<unsigned>((groups + per - 1) / per), S);
        const size_t shm = sizeof(unsigned) * g.T;
        if (vec) tile_count_kernel<true><<<grid, kCountThreads, shm, st>>>(x, y, tab, g, H, W, z.counts);
        else tile_count_kernel<false><<<grid, kCountThreads, shm, st>>>(x, y, tab, g, H, W, z.counts);
        CMDA_LAUNCH_CHECK();
    }
    tile_scan_kernel<<<S, 1024, 0, st>>>(z.counts, g, z.bucket_off, z.items, z.queue, z.max_items);
    CMDA_LAUNCH_CHECK();
    phase_mark(st);
    {
        const long long groups = (max_events + 7) / 8 + 1;
        const long long per = static_cast<long long>(kPartThreads) * kPartGroupsPerThread;
        dim3 grid(static_cast<unsigned>((groups + per - 1) / per), S);
        const size_t rec = (B == 1) ? sizeof(Rec2) : sizeof(Rec8);
        const size_t shm = sizeof(unsigned) * (2 * g.T + 2) + (rec + sizeof(unsigned)) * kPartChunk;
        int rc;
        if (B == 1) {
            if (vec) {
                if ((rc = set_smem(tile_partition_kernel<false, true>, shm)) != CMDA_OK) return rc;
                tile_partition_kernel<false, true><<<grid, kPartThreads, shm, st>>>(t, x, y, p, tab, tt, g, H, W, z.bucket_off, z.cursor, z.records);
            } else {
                if ((rc = set_smem(tile_partition_kernel<false, false>, shm)) != CMDA_OK) return rc;
                tile_partition_kernel<false, false><<<grid, kPartThreads, shm, st>>>(t, x, y, p, tab, tt, g, H, W, z.bucket_off, z.cursor, z.records);
            }
        } else {
            if (vec) {
                if ((rc = set_smem(tile_partition_kernel<true, true>, shm)) != CMDA_OK) return rc;
                tile_partition_kernel<true, true><<<grid, kPartThreads, shm, st>>>(t, x, y, p, tab, tt, g, H, W, z.bucket_off, z.cursor, z.records);
            } else {
                if ((rc = set_smem(tile_partition_kernel<true, false>, shm)) != CMDA_OK) return rc;
                tile_partition_kernel<true, false><<<grid, kPartThreads, shm, st>>>(t, x, y, p, tab, tt, g, H, W, z.bucket_off, z.cursor, z.records);
            }
        }
        CMDA_LAUNCH_CHECK();
    }
    phase_mark(st);
    {
        const size_t shm = sizeof(unsigned) * 2 * g.cap_voxels +
                           sizeof(float2) * (static_cast<size_t>(1) << (g.tw_log2 + g.th_log2)) + sizeof(unsigned) * 32;
        int per_sm = static_cast<int>((220 * 1024) / (shm + 1024));
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 2048 / kAccThreads) per_sm = 2048 / kAccThreads;
        int grid = 148 * per_sm;
        if (grid > z.max_items) grid = z.max_items;
        int rc;
        unsigned long long* uacc = reinterpret_cast<unsigned long long*>(acc);
        unsigned long long* ubins = reinterpret_cast<unsigned long long*>(bin_counts);
        if (B == 1) {
            if ((rc = set_smem(tile_accumulate_kernel<false>, shm)) != CMDA_OK) return rc;
            tile_accumulate_kernel<false><<<grid, kAccThreads, shm, st>>>(z.records, t, tab, tt, g, maps2, H, W, B, z.bbox, z.items, z.queue, uacc, ubins);
        } else {
            if ((rc = set_smem(tile_accumulate_kernel<true>, shm)) != CMDA_OK) return rc;
            tile_accumulate_kernel<true><<<grid, kAccThreads, shm, st>>>(z.records, t, tab, tt, g, maps2, H, W, B, z.bbox, z.items, z.queue, uacc, ubins);
        }
        CMDA_LAUNCH_CHECK();
    }
    return CMDA_OK;
}

}  // namespace cmda
