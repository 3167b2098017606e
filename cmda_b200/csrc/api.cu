// extern "C" entry points of libcmda_b200.so (see include/cmda_b200.h): argument
// validation, workspace carving, launch sequencing.  No allocation, no global state, no
// synchronisation, no CPU fallback.
#include <algorithm>

#include "common.cuh"

namespace cmda {

thread_local int g_last_cuda_error = 0;
thread_local PhaseTimer g_phase_timer = {nullptr, 0, 0};

// launchers defined in the kernel translation units
int launch_scatter_global_raw(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, const WindowTable&, int,
                              long long, const float*, int, int, int, long long*, int64_t*, cudaStream_t);
int launch_scatter_global_f32(const float*, const float*, const float*, const float*, long long, int, int, int,
                              long long*, int64_t*, cudaStream_t);
int launch_remap(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, long long, long long, const float*,
                 int, int, int, float*, float*, float*, int*, int*, int*, cudaStream_t);
int launch_convert_stats(const long long*, float*, int, long long, PartialStats*, cudaStream_t);
int launch_norm_apply(const float*, float*, int, long long, const PartialStats*, const WindowTable&, float, int,
                      cudaStream_t);
int launch_norm_augment(const float*, float*, int, int, int, int, const PartialStats*, const WindowTable&, const int*,
                        const int*, const int*, int, int, int, int, int, int, float, int, cudaStream_t);
size_t resize_workspace_bytes(int, int, int, int, int, int);
int launch_resize_bilinear(const uint8_t*, int, int, int, int, int, int, uint8_t*, void*, size_t, cudaStream_t);
int launch_u8_crop_center(const uint8_t*, int, int, int, const int*, const int*, const int*, int, int, int, float*, cudaStream_t);
int launch_rgb_to_gray(const uint8_t*, int64_t, uint8_t*, cudaStream_t);
int launch_denorm_to_gray(const float*, int, int, int, const float*, const float*, uint8_t*, uint8_t*, cudaStream_t);
int launch_isr(const uint8_t*, int, int, int, int, int, const float*, float, float, float*, unsigned*, size_t, cudaStream_t);
int launch_pair(const uint8_t*, const uint8_t*, int, int, int, const float*, float, float, float*, uint8_t*, unsigned*, size_t,
                cudaStream_t);
size_t image_slot_bytes(int);
size_t image_table_bytes(int, int, int);
// TILED mode (voxel_tiled.cu)
size_t tiled_workspace_bytes(int64_t total_events, int S, int H, int W, int B);
int tiled_supported(int H, int W, int B);
int factored_supported(int H, int W, int B);
size_t factored_scratch_bytes(int group, int H, int W, int B);
int factored_max_maps(void);
size_t factored_plan_bytes(int H, int W);
int launch_plan_build(const float*, int, int, int, void*, cudaStream_t);
int launch_factored(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, const PackedSrc*, const WindowTable&, int,
                    long long, const float*, int, int, int, void*, int64_t*, float*, PartialStats*, void*, size_t, const void*, int,
                    cudaStream_t);
int launch_unpack_p3(const uint8_t*, const int64_t*, int64_t, int64_t, int64_t, int64_t, uint32_t*, cudaStream_t);
int launch_pack_p4(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, int64_t, uint32_t, int64_t, uint32_t*, int64_t*,
                   int32_t*, cudaStream_t);
int banded_supported(int H, int W, int B);
size_t banded_scratch_bytes(long long total_events, int group, int H, int W, int B);
int exact_supported(int H, int W, int B);
size_t exact_workspace_bytes(long long max_window_events);
int launch_exact_raw(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, const WindowTable&, int, const float*, int,
                     int, int, float*, int64_t*, void*, size_t, cudaStream_t);
int launch_exact_f32(const float*, const float*, const float*, const float*, long long, int, int, int, float*, int64_t*, void*,
                     size_t, cudaStream_t);
int launch_tiled_scatter(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, const WindowTable&, int,
                         const float*, int, int, int, long long*, int64_t*, void*, size_t, cudaStream_t);

static size_t stats_bytes(int S) { return align_up(sizeof(PartialStats) * kStatBlocks * static_cast<size_t>(S), 256); }
static size_t acc_bytes(int S, int H, int W, int B) {
    return align_up(sizeof(long long) * static_cast<size_t>(S) * B * H * W, 256);
}

// AUTO: FACTORED wherever the sensor-space formulation applies.  Its stage A is the L2 RED kernel, except for
// B == 1 (the shipped events_bins) on large batches, where the BANDED stage A with the second cut of the partition
// pass (band partition + shared-memory counting, bit-identical R) is faster: measured on C2 0.443 vs 0.536 ms per
// 80 M events (first cut 0.463; profiles/r02_banded_phases.txt); at B = 5 the three forms are within 5 % of each
// other (0.77 / 0.80 / 0.81 ms) and the RED kernel stays.  Below the threshold the banded passes' fixed cost per
// (window, band) item loses to the 2-6 us the RED kernel needs.
constexpr long long kAutoBandedMinEvents = 8LL << 20;
static int resolve_mode(int mode, long long total_events, int S, int H, int W, int B) {
    (void)S;
    if (mode != CMDA_VOXEL_AUTO) return mode;
    if (!factored_supported(H, W, B)) return CMDA_VOXEL_GLOBAL;
    if (B == 1 && total_events >= kAutoBandedMinEvents && banded_supported(H, W, B)) return CMDA_VOXEL_BANDED2;
    return CMDA_VOXEL_FACTORED;
}

}  // namespace cmda

using namespace cmda;

extern "C" {

const char* cmda_strerror(int code) {
    switch (code) {
        case CMDA_OK: return "ok";
        case CMDA_ERR_BAD_ARG: return "bad argument";
        case CMDA_ERR_CUDA: return "CUDA runtime error";
        case CMDA_ERR_WORKSPACE: return "workspace too small or misaligned";
        case CMDA_ERR_UNSUPPORTED: return "unsupported shape";
        case CMDA_ERR_NO_DEVICE: return "no sm_100 device";
        default: return "unknown error";
    }
}

int cmda_version(void) { return CMDA_B200_VERSION; }

int cmda_profiler_attach(void* const* h_events, int n) {
    if (!h_events || n <= 0) return CMDA_ERR_BAD_ARG;
    g_phase_timer = {h_events, n, 0};
    return CMDA_OK;
}
int cmda_profiler_detach(void) {
    const int used = g_phase_timer.next;
    g_phase_timer = {nullptr, 0, 0};
    return used;
}
void* cmda_event_create(void) {
    cudaEvent_t e = nullptr;
    if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
    return e;
}
int cmda_event_destroy(void* e) {
    if (e) CMDA_CUDA_TRY(cudaEventDestroy(static_cast<cudaEvent_t>(e)));
    return CMDA_OK;
}
int cmda_event_elapsed_ms(void* a, void* b, float* h_ms) {
    if (!a || !b || !h_ms) return CMDA_ERR_BAD_ARG;
    CMDA_CUDA_TRY(cudaEventElapsedTime(h_ms, static_cast<cudaEvent_t>(a), static_cast<cudaEvent_t>(b)));
    return CMDA_OK;
}
int cmda_last_cuda_error(void) { return g_last_cuda_error; }

int cmda_searchsorted_right_u32(const uint32_t* d_t, int64_t n, const int64_t* d_q, int nq, int64_t* d_out,
                                void* stream) {
    if (n < 0 || nq < 0) return CMDA_ERR_BAD_ARG;
    if (nq == 0) return CMDA_OK;
    if ((n > 0 && !d_t) || !d_q || !d_out) return CMDA_ERR_BAD_ARG;
    return launch_searchsorted(d_t, n, d_q, nq, d_out, static_cast<cudaStream_t>(stream));
}

int cmda_images_to_events_index(const uint32_t* d_t, int64_t n, const int64_t* d_ms_to_idx, int64_t n_ms,
                                int64_t t_offset, const int64_t* d_timestamps, int n_ts, int64_t* d_index,
                                int32_t* d_status, void* stream) {
    if (n <= 0 || n_ms <= 0 || n_ts < 0) return CMDA_ERR_BAD_ARG;
    if (n_ts == 0) return CMDA_OK;
    if (!d_t || !d_ms_to_idx || !d_timestamps || !d_index || !d_status) return CMDA_ERR_BAD_ARG;
    return launch_images_to_events_index(d_t, n, d_ms_to_idx, n_ms, t_offset, d_timestamps, n_ts, d_index, d_status,
                                         static_cast<cudaStream_t>(stream));
}

int cmda_events_vg_resolved_mode(int64_t total_events, int S, int H, int W, int B, int mode) {
    if (S <= 0 || H <= 0 || W <= 0 || B <= 0 || total_events < 0) return CMDA_ERR_BAD_ARG;
    if (mode < CMDA_VOXEL_GLOBAL || mode > CMDA_VOXEL_BANDED2) return CMDA_ERR_BAD_ARG;
    return resolve_mode(mode, total_events, S, H, W, B);
}

size_t cmda_events_vg_workspace_bytes(int64_t total_events, int S, int H, int W, int B, int mode) {
    if (S <= 0 || H <= 0 || W <= 0 || B <= 0 || total_events < 0) return 0;
    const int group = S < kMaxWindows ? S : kMaxWindows;
    size_t need = stats_bytes(S) + acc_bytes(group, H, W, B);       // both paths sum into the int64 grid
    if (mode == CMDA_VOXEL_TILED && tiled_supported(H, W, B)) need += tiled_workspace_bytes(total_events, S, H, W, B);
    // AUTO may resolve to FACTORED or (B == 1) BANDED depending on the batch: sized for either, so that a workspace
    // sized for a larger batch always serves a smaller one
    const bool banded = (mode == CMDA_VOXEL_BANDED || mode == CMDA_VOXEL_BANDED2 || (mode == CMDA_VOXEL_AUTO && B == 1)) &&
                        banded_supported(H, W, B);
    if ((mode == CMDA_VOXEL_FACTORED || mode == CMDA_VOXEL_AUTO || banded) && factored_supported(H, W, B))
        need += factored_scratch_bytes(group, H, W, B);
    if (banded) need += 256 + banded_scratch_bytes(total_events, group, H, W, B);
    if (mode == CMDA_VOXEL_EXACT) need += exact_workspace_bytes(total_events);    // total_events bounds the largest window
    return need + 256;
}

}  // extern "C"

namespace {
struct AugmentArgs {      // post-voxel augmentation fused into the normaliser (dsec.py:304-319), or none
    const cmda_vg_augment* h_aug;
    int crop_w, crop_h, out_w, out_h, avg_bins, repeat;
};
struct PackedArgs {       // the packed (P4) source of cmda_events_vg_batch_p4, or none
    PackedSrc src;
    const int64_t* h_ms_to_idx;     // host copy of the table: the window-level bucket search happens here
    const int64_t* h_win_src;       // [S] store index of each window's first event, or NULL (= h_win_start)
};

int events_vg_impl(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p,
                   const int64_t* h_win_start, const int64_t* h_win_end, int S, const float* d_rectify_map,
                   const int32_t* h_map_id, int H, int W, int B, const float* h_clip, float final_range,
                   int enforce_no_events_zero, int normalize, float* d_out, float* d_raw_out, int64_t* d_bin_counts,
                   void* d_workspace, size_t workspace_bytes, int mode, void* stream, const AugmentArgs* aug,
                   const void* d_plans, const PackedArgs* packed = nullptr) {
    if (S < 0 || H <= 0 || W <= 0 || B <= 0) return CMDA_ERR_BAD_ARG;
    if (S == 0) return CMDA_OK;
    if (!packed && (!d_t || !d_x || !d_y || !d_p)) return CMDA_ERR_BAD_ARG;
    if (packed && (!packed->src.rec || !packed->src.ms_to_idx || !packed->h_ms_to_idx || packed->src.n_ms < 1 ||
                   packed->src.n_ms > 0x7ffffff0LL || packed->h_ms_to_idx[0] != 0))
        return CMDA_ERR_BAD_ARG;
    if (!h_win_start || !h_win_end || !d_out || !d_workspace) return CMDA_ERR_BAD_ARG;
    if (normalize && !h_clip) return CMDA_ERR_BAD_ARG;
    if (mode < CMDA_VOXEL_GLOBAL || mode > CMDA_VOXEL_BANDED2) return CMDA_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(d_workspace) & 255) return CMDA_ERR_WORKSPACE;
    long long total = 0;
    for (int s = 0; s < S; ++s) {
        if (h_win_start[s] < 0) return CMDA_ERR_BAD_ARG;
        if (h_map_id && h_map_id[s] < 0) return CMDA_ERR_BAD_ARG;
        const long long n = h_win_end[s] - h_win_start[s];
        if (n > 0) total += n;
    }
    size_t base_need = cmda_events_vg_workspace_bytes(total, S, H, W, B, mode);
    float* raw_ws = nullptr;      // augmented output: the raw grid lives in the workspace unless the caller wants it
    if (aug) {
        if (!aug->h_aug || aug->crop_w <= 0 || aug->crop_h <= 0 || aug->out_w <= 0 || aug->out_h <= 0 || aug->repeat <= 0)
            return CMDA_ERR_BAD_ARG;
        for (int s = 0; s < S; ++s)
            if (aug->h_aug[s].crop_x < 0 || aug->h_aug[s].crop_y < 0 || aug->h_aug[s].crop_x + aug->crop_w > W ||
                aug->h_aug[s].crop_y + aug->crop_h > H)
                return CMDA_ERR_BAD_ARG;
        if (!d_raw_out) {
            raw_ws = reinterpret_cast<float*>(static_cast<char*>(d_workspace) + base_need);
            base_need += sizeof(float) * static_cast<size_t>(S) * B * H * W;
        }
    }
    if (workspace_bytes < base_need) return CMDA_ERR_WORKSPACE;
    const int use_mode = resolve_mode(mode, total, S, H, W, B);
    // the packed source feeds the sensor-space formulation only (FACTORED's RED kernel and the BANDED cuts)
    if (packed && use_mode != CMDA_VOXEL_FACTORED && use_mode != CMDA_VOXEL_BANDED && use_mode != CMDA_VOXEL_BANDED2) return CMDA_ERR_UNSUPPORTED;
    if (use_mode == CMDA_VOXEL_TILED && !tiled_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    if (use_mode == CMDA_VOXEL_FACTORED && !factored_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    if (use_mode == CMDA_VOXEL_EXACT && !exact_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    const int banded = use_mode == CMDA_VOXEL_BANDED ? 1 : use_mode == CMDA_VOXEL_BANDED2 ? 2 : 0;     // which cut of the BANDED stage A
    if (banded && !banded_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    const bool factored_like = use_mode == CMDA_VOXEL_FACTORED || banded;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long V = static_cast<long long>(B) * H * W;

    char* ws = static_cast<char*>(d_workspace);
    PartialStats* partials = reinterpret_cast<PartialStats*>(ws);
    char* scratch = ws + stats_bytes(S);
    const size_t scratch_bytes = workspace_bytes - stats_bytes(S);
    if (d_bin_counts) CMDA_CUDA_TRY(cudaMemsetAsync(d_bin_counts, 0, sizeof(int64_t) * S * B, st));

    for (int s0 = 0, sn = 0; s0 < S; s0 += sn) {
        // a launch group: at most kMaxWindows windows and, for FACTORED, at most factored_max_maps()
        // distinct rectify maps (one inverse index each)
        sn = (S - s0) < kMaxWindows ? (S - s0) : kMaxWindows;
        if (factored_like && h_map_id && d_rectify_map && !d_plans) {
            int ids[kMaxWindows], n_ids = 0, k = 0;
            for (; k < sn; ++k) {
                const int id = h_map_id[s0 + k];
                int j = 0;
                while (j < n_ids && ids[j] != id) ++j;
                if (j == n_ids) {
                    if (n_ids == factored_max_maps()) break;
                    ids[n_ids++] = id;
                }
            }
            sn = k;
        }
        WindowTable tab{};
        long long max_events = 0;
        for (int k = 0; k < sn; ++k) {
            const int s = s0 + k;
            tab.w[k].start = h_win_start[s];
            tab.w[k].end = h_win_end[s] > h_win_start[s] ? h_win_end[s] : h_win_start[s];
            tab.w[k].map_id = h_map_id ? h_map_id[s] : 0;
            tab.w[k].clip = h_clip ? h_clip[s] : 1.0f;
            const long long n = tab.w[k].end - tab.w[k].start;
            if (n > max_events) max_events = n;
            if (packed && n > 0) {
                // millisecond buckets of the window's first and last event: largest k with ms_to_idx[k] <= index
                const long long src = packed->h_win_src ? packed->h_win_src[s] : tab.w[k].start;
                const int64_t* tb = packed->h_ms_to_idx;
                const long long n_ms = packed->src.n_ms;
                if (src < 0 || src + n > tb[n_ms]) return CMDA_ERR_BAD_ARG;
                tab.w[k].src_shift = src - tab.w[k].start;
                tab.w[k].ms_lo = static_cast<int>(std::upper_bound(tb, tb + n_ms, static_cast<int64_t>(src)) - tb - 1);
                tab.w[k].ms_hi = static_cast<int>(std::upper_bound(tb, tb + n_ms, static_cast<int64_t>(src + n - 1)) - tb - 1);
            }
        }
        float* out_g = d_out + static_cast<size_t>(s0) * V;
        float* raw_g = (normalize && d_raw_out) ? d_raw_out + static_cast<size_t>(s0) * V : out_g;
        PartialStats* part_g = partials + static_cast<size_t>(s0) * kStatBlocks;
        int crop_x[kMaxWindows], crop_y[kMaxWindows], flip[kMaxWindows];
        if (aug) {
            const int Bo = aug->avg_bins ? 1 : B;
            out_g = d_out + static_cast<size_t>(s0) * aug->repeat * Bo * aug->out_h * aug->out_w;
            raw_g = (d_raw_out ? d_raw_out : raw_ws) + static_cast<size_t>(s0) * V;
            for (int k = 0; k < sn; ++k) {
                crop_x[k] = aug->h_aug[s0 + k].crop_x; crop_y[k] = aug->h_aug[s0 + k].crop_y; flip[k] = aug->h_aug[s0 + k].flip;
            }
        }
        // the last stage: events_norm apply, alone or fused with the dataset's crop / flip / resize / repeat
        auto finish = [&]() -> int {
            if (aug)
                return launch_norm_augment(raw_g, out_g, sn, B, H, W, part_g, tab, crop_x, crop_y, flip, aug->crop_w, aug->crop_h,
                                           aug->out_w, aug->out_h, aug->avg_bins, aug->repeat, final_range,
                                           enforce_no_events_zero, st);
            return launch_norm_apply(raw_g, out_g, sn, V, part_g, tab, final_range, enforce_no_events_zero, st);
        };
        int64_t* bins_g = d_bin_counts ? d_bin_counts + static_cast<size_t>(s0) * B : nullptr;
        int rc;
        phase_mark(st);
        long long* acc = reinterpret_cast<long long*>(scratch);
        if (factored_like) {
            const size_t ab = acc_bytes(sn, H, W, B);
            rc = launch_factored(d_t, d_x, d_y, d_p, packed ? &packed->src : nullptr, tab, sn, max_events, d_rectify_map, H, W, B, acc, bins_g, raw_g,
                                 part_g, scratch + ab, scratch_bytes - ab, d_plans, banded, st);   // marks: memset | plans | accumulate
            if (rc != CMDA_OK) return rc;
            phase_mark(st);
            if (normalize) {
                rc = finish();
                if (rc != CMDA_OK) return rc;
                phase_mark(st);
            }
            continue;
        }
        if (use_mode == CMDA_VOXEL_EXACT) {
            // reference-order float32 sums straight into the raw grid; statistics from the float grid
            const size_t ab = acc_bytes(sn, H, W, B);
            rc = launch_exact_raw(d_t, d_x, d_y, d_p, tab, sn, d_rectify_map, H, W, B, raw_g, bins_g, scratch + ab,
                                  scratch_bytes - ab, st);
            if (rc != CMDA_OK) return rc;
            phase_mark(st);
            rc = launch_convert_stats(nullptr, raw_g, sn, V, part_g, st);
            if (rc != CMDA_OK) return rc;
            phase_mark(st);
            if (normalize) {
                rc = finish();
                if (rc != CMDA_OK) return rc;
                phase_mark(st);
            }
            continue;
        }
        CMDA_CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(long long) * sn * V, st));
        phase_mark(st);
        if (use_mode == CMDA_VOXEL_TILED) {
            const size_t ab = acc_bytes(sn, H, W, B);
            rc = launch_tiled_scatter(d_t, d_x, d_y, d_p, tab, sn, d_rectify_map, H, W, B, acc, bins_g, scratch + ab,
                                      scratch_bytes - ab, st);       // marks: count+scan | partition
        } else {
            rc = launch_scatter_global_raw(d_t, d_x, d_y, d_p, tab, sn, max_events, d_rectify_map, H, W, B, acc, bins_g, st);
        }
        if (rc != CMDA_OK) return rc;
        phase_mark(st);
        rc = launch_convert_stats(acc, raw_g, sn, V, part_g, st);
        if (rc != CMDA_OK) return rc;
        phase_mark(st);
        if (normalize) {
            rc = finish();
            if (rc != CMDA_OK) return rc;
            phase_mark(st);
        }
    }
    return CMDA_OK;
}
}  // namespace

extern "C" {

int cmda_events_vg_batch(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p,
                         const int64_t* h_win_start, const int64_t* h_win_end, int S, const float* d_rectify_map,
                         const int32_t* h_map_id, int H, int W, int B, const float* h_clip, float final_range,
                         int enforce_no_events_zero, int normalize, float* d_out, float* d_raw_out,
                         int64_t* d_bin_counts, void* d_workspace, size_t workspace_bytes, int mode, void* stream) {
    return events_vg_impl(d_t, d_x, d_y, d_p, h_win_start, h_win_end, S, d_rectify_map, h_map_id, H, W, B, h_clip, final_range,
                          enforce_no_events_zero, normalize, d_out, d_raw_out, d_bin_counts, d_workspace, workspace_bytes, mode,
                          stream, nullptr, nullptr);
}

int cmda_events_vg_batch_planned(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p,
                                 const int64_t* h_win_start, const int64_t* h_win_end, int S, const float* d_rectify_map,
                                 const int32_t* h_map_id, int H, int W, int B, const float* h_clip, float final_range,
                                 int enforce_no_events_zero, int normalize, float* d_out, float* d_raw_out,
                                 int64_t* d_bin_counts, void* d_workspace, size_t workspace_bytes, int mode,
                                 const void* d_plans, void* stream) {
    return events_vg_impl(d_t, d_x, d_y, d_p, h_win_start, h_win_end, S, d_rectify_map, h_map_id, H, W, B, h_clip, final_range,
                          enforce_no_events_zero, normalize, d_out, d_raw_out, d_bin_counts, d_workspace, workspace_bytes, mode,
                          stream, nullptr, d_plans);
}

int cmda_events_vg_batch_p4(const uint32_t* d_rec, const int64_t* d_ms_to_idx, const int64_t* h_ms_to_idx, int64_t n_ms,
                            const int64_t* h_win_start, const int64_t* h_win_end, const int64_t* h_win_src, int S,
                            const float* d_rectify_map, const int32_t* h_map_id, int H, int W, int B, const float* h_clip,
                            float final_range, int enforce_no_events_zero, int normalize, float* d_out, float* d_raw_out,
                            int64_t* d_bin_counts, void* d_workspace, size_t workspace_bytes, int mode, const void* d_plans,
                            void* stream) {
    if (W > 2048 || H > 1024) return CMDA_ERR_UNSUPPORTED;      // 11-bit x, 10-bit y
    const PackedArgs pa{PackedSrc{d_rec, reinterpret_cast<const long long*>(d_ms_to_idx), n_ms}, h_ms_to_idx, h_win_src};
    return events_vg_impl(nullptr, nullptr, nullptr, nullptr, h_win_start, h_win_end, S, d_rectify_map, h_map_id, H, W, B, h_clip,
                          final_range, enforce_no_events_zero, normalize, d_out, d_raw_out, d_bin_counts, d_workspace,
                          workspace_bytes, mode, stream, nullptr, d_plans, &pa);
}

int cmda_pack_events_p4(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p, int64_t n,
                        uint32_t t_base_us, int64_t n_ms, uint32_t* d_rec, int64_t* d_ms_to_idx, int32_t* d_status,
                        void* stream) {
    if (n < 0 || n_ms < 1) return CMDA_ERR_BAD_ARG;
    if (!d_ms_to_idx || !d_status || (n > 0 && (!d_t || !d_x || !d_y || !d_p || !d_rec))) return CMDA_ERR_BAD_ARG;
    return launch_pack_p4(d_t, d_x, d_y, d_p, n, t_base_us, n_ms, d_rec, d_ms_to_idx, d_status, static_cast<cudaStream_t>(stream));
}

int cmda_unpack_p3_to_p4(const uint8_t* d_rec3, const int64_t* d_sub_to_idx, int64_t sub_lo, int64_t sub_hi, int64_t first,
                         int64_t last, uint32_t* d_rec4, void* stream) {
    if (first < 0 || last < first || sub_lo < 0) return CMDA_ERR_BAD_ARG;
    if (last == first) return CMDA_OK;
    if (sub_hi < sub_lo || !d_rec3 || !d_sub_to_idx || !d_rec4) return CMDA_ERR_BAD_ARG;
    return launch_unpack_p3(d_rec3, d_sub_to_idx, sub_lo, sub_hi, first, last, d_rec4, static_cast<cudaStream_t>(stream));
}

size_t cmda_rectify_plan_bytes(int H, int W) {
    if (H <= 0 || W <= 0) return 0;
    return factored_plan_bytes(H, W);
}

int cmda_rectify_plan_build(const float* d_rectify_map, int n_maps, int H, int W, void* d_plans, void* stream) {
    if (n_maps < 0 || H <= 0 || W <= 0) return CMDA_ERR_BAD_ARG;
    if (n_maps == 0) return CMDA_OK;
    if (!d_rectify_map || !d_plans) return CMDA_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(d_plans) & 255) return CMDA_ERR_WORKSPACE;
    return launch_plan_build(d_rectify_map, n_maps, H, W, d_plans, static_cast<cudaStream_t>(stream));
}

size_t cmda_events_vg_augmented_workspace_bytes(int64_t total_events, int S, int H, int W, int B, int mode) {
    const size_t base = cmda_events_vg_workspace_bytes(total_events, S, H, W, B, mode);
    return base ? base + sizeof(float) * static_cast<size_t>(S) * B * H * W : 0;
}

int cmda_events_vg_augmented_batch(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p,
                                   const int64_t* h_win_start, const int64_t* h_win_end, int S,
                                   const float* d_rectify_map, const int32_t* h_map_id, int H, int W, int B,
                                   const float* h_clip, float final_range, int enforce_no_events_zero,
                                   const cmda_vg_augment* h_aug, int crop_w, int crop_h, int out_w, int out_h,
                                   int avg_bins, int repeat, float* d_out, float* d_raw_out, int64_t* d_bin_counts,
                                   void* d_workspace, size_t workspace_bytes, int mode, const void* d_plans, void* stream) {
    const AugmentArgs aug{h_aug, crop_w, crop_h, out_w, out_h, avg_bins, repeat};
    return events_vg_impl(d_t, d_x, d_y, d_p, h_win_start, h_win_end, S, d_rectify_map, h_map_id, H, W, B, h_clip, final_range,
                          enforce_no_events_zero, 1, d_out, d_raw_out, d_bin_counts, d_workspace, workspace_bytes, mode, stream,
                          &aug, d_plans);
}

int cmda_voxel_grid_f32(const float* d_time, const float* d_x, const float* d_y, const float* d_pol, int64_t n, int W,
                        int H, int B, float* d_grid, int64_t* d_bin_counts, void* d_workspace, size_t workspace_bytes,
                        int mode, void* stream) {
    if (n < 0 || H <= 0 || W <= 0 || B <= 0 || !d_grid || !d_workspace) return CMDA_ERR_BAD_ARG;
    if (n > 0 && (!d_time || !d_x || !d_y || !d_pol)) return CMDA_ERR_BAD_ARG;
    if (mode < CMDA_VOXEL_GLOBAL || mode > CMDA_VOXEL_BANDED2) return CMDA_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(d_workspace) & 255) return CMDA_ERR_WORKSPACE;
    if (workspace_bytes < cmda_events_vg_workspace_bytes(n, 1, H, W, B, mode)) return CMDA_ERR_WORKSPACE;
    // float32 events are already rectified: there is no map gather to tile, so this entry point
    // runs the GLOBAL scatter (AUTO resolves to it); TILED / EXACT are refused explicitly
    if (mode == CMDA_VOXEL_TILED || mode >= CMDA_VOXEL_FACTORED) return CMDA_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long V = static_cast<long long>(B) * H * W;
    char* ws = static_cast<char*>(d_workspace);
    PartialStats* partials = reinterpret_cast<PartialStats*>(ws);
    char* scratch = ws + stats_bytes(1);
    if (d_bin_counts) CMDA_CUDA_TRY(cudaMemsetAsync(d_bin_counts, 0, sizeof(int64_t) * B, st));
    if (mode == CMDA_VOXEL_EXACT) {
        const size_t ab = acc_bytes(1, H, W, B);
        return launch_exact_f32(d_time, d_x, d_y, d_pol, n, H, W, B, d_grid, d_bin_counts, scratch + ab,
                                workspace_bytes - stats_bytes(1) - ab, st);
    }
    long long* acc = reinterpret_cast<long long*>(scratch);
    CMDA_CUDA_TRY(cudaMemsetAsync(acc, 0, sizeof(long long) * V, st));
    int rc = launch_scatter_global_f32(d_time, d_x, d_y, d_pol, n, H, W, B, acc, d_bin_counts, st);
    if (rc != CMDA_OK) return rc;
    return launch_convert_stats(acc, d_grid, 1, V, partials, st);
}

size_t cmda_events_norm_workspace_bytes(int S) { return S > 0 ? stats_bytes(S) + 256 : 0; }

int cmda_events_norm_batch(float* d_grid, int S, int64_t voxels, const float* h_clip, float final_range,
                           int enforce_no_events_zero, void* d_workspace, size_t workspace_bytes, void* stream) {
    if (S < 0 || voxels <= 0) return CMDA_ERR_BAD_ARG;
    if (S == 0) return CMDA_OK;
    if (!d_grid || !h_clip || !d_workspace) return CMDA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) || workspace_bytes < cmda_events_norm_workspace_bytes(S))
        return CMDA_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    PartialStats* partials = static_cast<PartialStats*>(d_workspace);
    for (int s0 = 0; s0 < S; s0 += kMaxWindows) {
        const int sn = (S - s0) < kMaxWindows ? (S - s0) : kMaxWindows;
        WindowTable tab{};
        for (int k = 0; k < sn; ++k) tab.w[k].clip = h_clip[s0 + k];
        float* g = d_grid + static_cast<size_t>(s0) * voxels;
        PartialStats* pg = partials + static_cast<size_t>(s0) * kStatBlocks;
        int rc = launch_convert_stats(nullptr, g, sn, voxels, pg, st);
        if (rc != CMDA_OK) return rc;
        rc = launch_norm_apply(g, g, sn, voxels, pg, tab, final_range, enforce_no_events_zero, st);
        if (rc != CMDA_OK) return rc;
    }
    return CMDA_OK;
}

int cmda_remap_events(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p, int64_t start,
                      int64_t end, const float* d_rectify_map, int H, int W, int B, float* d_xr, float* d_yr,
                      float* d_tn, int32_t* d_x0, int32_t* d_y0, int32_t* d_t0, void* stream) {
    if (start < 0 || H <= 0 || W <= 0 || B <= 0) return CMDA_ERR_BAD_ARG;
    if (end <= start) return CMDA_OK;
    if (!d_t || !d_x || !d_y || !d_p) return CMDA_ERR_BAD_ARG;
    return launch_remap(d_t, d_x, d_y, d_p, start, end, d_rectify_map, H, W, B, d_xr, d_yr, d_tn, d_x0, d_y0, d_t0,
                        static_cast<cudaStream_t>(stream));
}

size_t cmda_image_workspace_bytes(int S, int H, int W, int channels) {
    if (S <= 0 || H <= 0 || W <= 0) return 0;
    // min / max slots | pair tables of the table-driven apply passes (large images) | gray plane of an RGB input
    size_t need = image_slot_bytes(S) + align_up(image_table_bytes(S, H, W), 256);
    if (channels == 3) need += align_up(static_cast<size_t>(S) * H * W, 256);
    return need + 256;
}

int cmda_logdiff_pair_u8(const uint8_t* d_now, const uint8_t* d_front, int S, int H, int W, const float* h_lut,
                         float thr, float clip, float* d_out_f32, uint8_t* d_out_u8, void* d_workspace,
                         size_t workspace_bytes, void* stream) {
    if (S < 0 || H <= 0 || W <= 0) return CMDA_ERR_BAD_ARG;
    if (S == 0) return CMDA_OK;
    if (!d_now || !d_front || !h_lut || (!d_out_f32 && !d_out_u8) || !d_workspace) return CMDA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) || workspace_bytes < cmda_image_workspace_bytes(S, H, W, 1))
        return CMDA_ERR_WORKSPACE;
    return launch_pair(d_now, d_front, S, H, W, h_lut, thr, clip, d_out_f32, d_out_u8,
                       static_cast<unsigned*>(d_workspace), image_slot_bytes(S) + image_table_bytes(S, H, W),
                       static_cast<cudaStream_t>(stream));
}

int cmda_isr_shift_u8(const uint8_t* d_img, int channels, int S, int H, int W, int shift_pixel, int direction,
                      const float* h_lut, float thr, float clip, float* d_out, void* d_workspace,
                      size_t workspace_bytes, void* stream) {
    if (S < 0 || H <= 0 || W <= 0 || (channels != 1 && channels != 3)) return CMDA_ERR_BAD_ARG;
    if (direction < CMDA_DIR_RIGHTDOWN || direction > CMDA_DIR_ALL) return CMDA_ERR_BAD_ARG;
    // numpy slicing of the reference (utils.py:129-132) needs 0 <= shift <= min(H, W)
    if (shift_pixel < 0 || shift_pixel > W || shift_pixel > H) return CMDA_ERR_BAD_ARG;
    if (S == 0) return CMDA_OK;
    if (!d_img || !h_lut || !d_out || !d_workspace) return CMDA_ERR_BAD_ARG;
    if ((reinterpret_cast<uintptr_t>(d_workspace) & 255) ||
        workspace_bytes < cmda_image_workspace_bytes(S, H, W, channels))
        return CMDA_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned* slots = static_cast<unsigned*>(d_workspace);
    const uint8_t* gray = d_img;
    if (channels == 3) {
        uint8_t* g = reinterpret_cast<uint8_t*>(d_workspace) + image_slot_bytes(S) + align_up(image_table_bytes(S, H, W), 256);
        int rc = launch_rgb_to_gray(d_img, static_cast<int64_t>(S) * H * W, g, st);
        if (rc != CMDA_OK) return rc;
        gray = g;
    }
    return launch_isr(gray, S, H, W, shift_pixel, direction, h_lut, thr, clip, d_out, slots,
                      image_slot_bytes(S) + image_table_bytes(S, H, W), st);
}

int cmda_denorm_rgb_to_gray_u8(const float* d_img, int S, int H, int W, const float* mean, const float* stdv,
                               uint8_t* d_gray, uint8_t* d_rgb, void* stream) {
    if (S < 0 || H <= 0 || W <= 0) return CMDA_ERR_BAD_ARG;
    if (S == 0) return CMDA_OK;
    if (!d_img || !mean || !stdv || !d_gray) return CMDA_ERR_BAD_ARG;
    return launch_denorm_to_gray(d_img, S, H, W, mean, stdv, d_gray, d_rgb, static_cast<cudaStream_t>(stream));
}

size_t cmda_resize_bilinear_workspace_bytes(int S, int H, int W, int channels, int out_h, int out_w) {
    if (S <= 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || (channels != 1 && channels != 3)) return 0;
    return resize_workspace_bytes(S, H, W, channels, out_h, out_w);
}

int cmda_resize_bilinear_u8(const uint8_t* d_src, int channels, int S, int H, int W, int out_h, int out_w, uint8_t* d_dst,
                            void* d_workspace, size_t workspace_bytes, void* stream) {
    if (S < 0 || H <= 0 || W <= 0 || out_h <= 0 || out_w <= 0 || (channels != 1 && channels != 3)) return CMDA_ERR_BAD_ARG;
    if (S == 0) return CMDA_OK;
    if (!d_src || !d_dst || !d_workspace) return CMDA_ERR_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(d_workspace) & 255) return CMDA_ERR_WORKSPACE;
    return launch_resize_bilinear(d_src, channels, S, H, W, out_h, out_w, d_dst, d_workspace, workspace_bytes,
                                  static_cast<cudaStream_t>(stream));
}

int cmda_u8_crop_to_centered_f32(const uint8_t* d_src, int S, int H, int W, const cmda_vg_augment* h_aug, int crop_w, int crop_h,
                                 int repeat, float* d_out, void* stream) {
    if (S < 0 || H <= 0 || W <= 0 || crop_w <= 0 || crop_h <= 0 || repeat <= 0) return CMDA_ERR_BAD_ARG;
    if (S == 0) return CMDA_OK;
    if (!d_src || !h_aug || !d_out) return CMDA_ERR_BAD_ARG;
    int cx[kMaxWindows], cy[kMaxWindows], fl[kMaxWindows];
    for (int s0 = 0; s0 < S; s0 += kMaxWindows) {
        const int sn = (S - s0) < kMaxWindows ? (S - s0) : kMaxWindows;
        for (int k = 0; k < sn; ++k) {
            const cmda_vg_augment& a = h_aug[s0 + k];
            if (a.crop_x < 0 || a.crop_y < 0 || a.crop_x + crop_w > W || a.crop_y + crop_h > H) return CMDA_ERR_BAD_ARG;
            cx[k] = a.crop_x; cy[k] = a.crop_y; fl[k] = a.flip;
        }
        const int rc = launch_u8_crop_center(d_src + static_cast<size_t>(s0) * H * W, sn, H, W, cx, cy, fl, crop_w, crop_h, repeat,
                                             d_out + static_cast<size_t>(s0) * repeat * crop_w * crop_h,
                                             static_cast<cudaStream_t>(stream));
        if (rc != CMDA_OK) return rc;
    }
    return CMDA_OK;
}

int cmda_rgb_to_gray_u8(const uint8_t* d_rgb, int64_t n_pixels, uint8_t* d_gray, void* stream) {
    if (n_pixels < 0) return CMDA_ERR_BAD_ARG;
    if (n_pixels == 0) return CMDA_OK;
    if (!d_rgb || !d_gray) return CMDA_ERR_BAD_ARG;
    return launch_rgb_to_gray(d_rgb, n_pixels, d_gray, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
