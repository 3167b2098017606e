/* cmda_b200 -- C ABI of the B200-native event-representation path of XiaRho/CMDA.
 *
 * The reference is pure Python and has NO FFI: its boundary for this path is a set of
 * Python call signatures.  This header is the operator ABI those
 * signatures bind to in this build; each entry point cites the reference function it
 * replaces.  INTEGRATION.md shows the ctypes stub a maintainer of the reference adds.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch / CUDA types in the signatures
 *    (`stream` is a cudaStream_t passed as void*; NULL = legacy default stream);
 *  - pointers named `d_*` are DEVICE pointers, `h_*` are HOST pointers read during the
 *    call (small per-window tables, 256-entry LUTs); they may be freed on return;
 *  - the caller owns every buffer including the workspace; the library allocates no
 *    device memory and keeps no global state, except one internal side stream and two
 *    events per (host thread, device), made on first use by the voxel entry points that
 *    build map plans per call (those kernels run beside the event accumulation);
 *    calls are asynchronous on `stream`, never synchronise -- event record / wait and
 *    programmatic dependent launches only, so a call can be captured into a CUDA graph --
 *    and are re-entrant with one workspace per concurrent stream;
 *  - every function returns CMDA_OK (0) or a negative CMDA_ERR_* code; nothing throws
 *    or exits.  There is no CPU fallback anywhere behind this ABI.
 */
#ifndef CMDA_B200_H
#define CMDA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMDA_B200_VERSION 100

enum {
    CMDA_OK = 0,
    CMDA_ERR_BAD_ARG = -1,     /* NULL pointer, non-positive size, unknown mode ...        */
    CMDA_ERR_CUDA = -2,        /* a CUDA runtime call or launch failed (see cmda_last_cuda_error) */
    CMDA_ERR_WORKSPACE = -3,   /* workspace too small or misaligned (needs 256-byte alignment) */
    CMDA_ERR_UNSUPPORTED = -4, /* shape outside what the kernels are built for              */
    CMDA_ERR_NO_DEVICE = -5    /* no sm_100 device / driver                                  */
};

/* Accumulation modes of the voxel scatter.  All are deterministic (bit-reproducible).
 * GLOBAL and TILED evaluate the reference's float32 corner weight of every event
 * (dsec.py:51-52, bit-identical per contribution), quantise it to 2^-30 and sum 64-bit
 * integers, which is order independent: they are bit-identical to each other and closer to
 * the exact sum than the reference's float32 accumulation.
 * FACTORED uses that rectify_map is a function of the raw pixel only: it sums the temporal
 * weights of each raw pixel exactly (one 64-bit RED per EVENT, no map gather) and multiplies
 * by the pixel's four spatial weights once per pixel (gather through an inverse index of the
 * map, no atomics).  Each contribution differs from the reference's by at most one float32
 * rounding (2^-24 relative); for B == 1 it is the exact sum of the reference's weights.
 * Capacity: fewer than 2^19 events of net polarity per (raw pixel, temporal bin, window).
 * EXACT performs the reference's own float32 additions in the reference's own order
 * (corner pass major, event index minor: dsec.py:47-58): its raw grid is BIT-IDENTICAL to the
 * single-thread reference.  A verification mode (stable sort of every event-corner pair, 64 bytes of
 * workspace per pair of the largest window), not a fast one. */
enum {
    CMDA_VOXEL_GLOBAL = 0,   /* one 64-bit integer RED per corner into an L2-resident grid     */
    CMDA_VOXEL_TILED = 1,    /* raw-tile multisplit + shared-memory fixed-point accumulation   */
    CMDA_VOXEL_AUTO = 2,     /* FACTORED where it applies (raw DSEC events), else GLOBAL       */
    CMDA_VOXEL_EXACT = 3,    /* stable sort by (voxel, corner pass) + ordered float32 accumulation */
    CMDA_VOXEL_FACTORED = 4, /* sensor-space temporal accumulation + per-pixel rectify gather  */
    CMDA_VOXEL_BANDED = 5,   /* FACTORED with its per-event L2 atomics replaced by a band partition +
                                shared-memory accumulation; bit-identical to FACTORED; polarity in {0, 1} */
    CMDA_VOXEL_BANDED2 = 6   /* second cut of BANDED: fewer instructions per event, any polarity byte; same
                                output; verified on the CPU emulation of tests/emu, first hardware run pending */
};

/* Directions of the shift-pair generator (reference mmseg/datasets/utils.py:128-151). */
enum {
    CMDA_DIR_RIGHTDOWN = 0,
    CMDA_DIR_RIGHTUP = 1,
    CMDA_DIR_LEFTDOWN = 2,
    CMDA_DIR_LEFTUP = 3,
    CMDA_DIR_ALL = 4
};

const char* cmda_strerror(int code);
int cmda_version(void);
/* cudaError_t of the last failing CUDA call made by this thread inside the library. */
int cmda_last_cuda_error(void);

/* Optional phase timer used by bench.py for the per-kernel roofline: attach a host array
 * of n cudaEvent_t (created with cmda_event_create) to the CALLING THREAD; every phase
 * boundary of the following cmda_events_vg_batch calls records the next event of the list
 * on the call's stream (GLOBAL: start | memset | scatter | convert+stats | apply;
 * TILED: start | memset | bbox+count+scan | partition | accumulate | convert+stats | apply;
 * FACTORED: start | memset | inverse index | sensor accumulate | rectify gather+stats | apply).
 * detach returns how many events were recorded.  Off by default. */
int cmda_profiler_attach(void* const* h_events, int n);
int cmda_profiler_detach(void);
void* cmda_event_create(void);
int cmda_event_destroy(void* event);
int cmda_event_elapsed_ms(void* start_event, void* end_event, float* h_ms);

/* ------------------------------------------------------------------------------------
 * K1  window slicer.
 * Replaces: create_images_to_events_index, /root/reference/create_dsec_dataset_txt.py:19-42
 * (np.searchsorted(..., 'right') inside the ms_to_idx bracket).
 * out[i] = number of elements of the ascending array d_t[0..n) that are <= d_q[i]. */
int cmda_searchsorted_right_u32(const uint32_t* d_t, int64_t n, const int64_t* d_q, int nq,
                                int64_t* d_out, void* stream);

/* The whole per-image computation of create_dsec_dataset_txt.py:19-42 for n_ts image
 * timestamps: d_index[i] = index of the last event with t <= ts - t_offset, or -1 when
 * ts - t_offset <= 0 or > t[n-1]; d_status[i] = 1 when the reference would raise
 * ValueError('range error!') (line 37-39), else 0. */
int cmda_images_to_events_index(const uint32_t* d_t, int64_t n, const int64_t* d_ms_to_idx, int64_t n_ms,
                                int64_t t_offset, const int64_t* d_timestamps, int n_ts,
                                int64_t* d_index, int32_t* d_status, void* stream);

/* ------------------------------------------------------------------------------------
 * K2+K3  event windows -> voxel grids.
 * Replaces: DSECDataset.get_events_vg, /root/reference/mmseg/datasets/dsec.py:341-366
 * (window-relative time, rectify_map[y, x] gather, events_to_voxel_grid dsec.py:26-58,
 * events_norm dsec.py:80-121), batched over S independent windows of one SoA store.
 *
 *  d_t/d_x/d_y/d_p   SoA event store in DSEC dtypes (16-byte aligned for the fast path)
 *  h_win_start/end   [S] window bounds as event indices, end EXCLUSIVE (= finish + 1)
 *  d_rectify_map     [n_maps, H, W, 2] float32 (channel 0 = x, 1 = y) or NULL (no remap)
 *  h_map_id          [S] map index per window, or NULL (all windows use map 0)
 *  h_clip            [S] clip_range per window (dsec.py:359-362); ignored if !normalize
 *  normalize         0: d_out receives the raw grid (events_to_voxel_grid output)
 *                    1: d_out receives events_norm(raw, clip, final_range, enforce)
 *  d_out             [S, B, H, W] float32
 *  d_raw_out         optional [S, B, H, W] raw grid when normalize=1 (may be NULL)
 *  d_bin_counts      optional [S, B] int64: events per temporal bin t0 (may be NULL)
 *  empty windows (end <= start) and single-timestamp windows (NaN time)
 *  produce an all-zero raw grid, exactly like the reference. */
size_t cmda_events_vg_workspace_bytes(int64_t total_events, int S, int H, int W, int B, int mode);
/* The concrete mode CMDA_VOXEL_AUTO resolves to for this batch (or `mode` itself). */
int cmda_events_vg_resolved_mode(int64_t total_events, int S, int H, int W, int B, int mode);

int cmda_events_vg_batch(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p,
                         const int64_t* h_win_start, const int64_t* h_win_end, int S,
                         const float* d_rectify_map, const int32_t* h_map_id, int H, int W, int B,
                         const float* h_clip, float final_range, int enforce_no_events_zero, int normalize,
                         float* d_out, float* d_raw_out, int64_t* d_bin_counts,
                         void* d_workspace, size_t workspace_bytes, int mode, void* stream);

/* Rectify-map plans.  Everything the FACTORED mode derives from a rectify map alone -- the inverse index
 * of the map, the sparse gather stencil, the source box of every output tile -- is a "plan" of
 * cmda_rectify_plan_bytes(H, W) bytes per map.  cmda_events_vg_batch builds the plans of a call's maps
 * into its workspace on every call (the reference, too, re-reads the map for every sample:
 * dsec.py:289-291); a caller that keeps a sequence resident builds them once and passes them to the
 * *_planned entry point: d_plans is [n_maps] plans, 256-byte aligned, plan m belongs to map id m.
 * Plans are only read by the voxel calls; NULL means "build per call". */
size_t cmda_rectify_plan_bytes(int H, int W);
int cmda_rectify_plan_build(const float* d_rectify_map, int n_maps, int H, int W, void* d_plans, void* stream);
int cmda_events_vg_batch_planned(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p,
                                 const int64_t* h_win_start, const int64_t* h_win_end, int S,
                                 const float* d_rectify_map, const int32_t* h_map_id, int H, int W, int B,
                                 const float* h_clip, float final_range, int enforce_no_events_zero, int normalize,
                                 float* d_out, float* d_raw_out, int64_t* d_bin_counts, void* d_workspace,
                                 size_t workspace_bytes, int mode, const void* d_plans, void* stream);

/* ------------------------------------------------------------------------------------
 * K2+K3 from the PACKED event stream ("P4"): 4 bytes per event instead of the 9 of the SoA arrays.
 * Replaces: the same DSECDataset.get_events_vg (dsec.py:341-366); the packed stream stands in for the
 * decoded events/{t,x,y,p} datasets of events.h5 (dsec.py:342-345) on the wire between host and device and
 * in HBM -- the host->device copy of the events is what bounds the end-to-end rate of this path.
 *
 *   record  = x | y << 11 | p << 21 | sub << 22        (x < 2048, y < 1024, p in {0, 1}, sub < 1000)
 *   t_us    = t_base + 1000 * ms + sub, ms = the event's millisecond bucket = the largest k with
 *             ms_to_idx[k] <= event index  (ms_to_idx[k] = index of the first event with t_us - t_base >= 1000 k;
 *             DSEC's events.h5 carries exactly this table: create_dsec_dataset_txt.py:16, 26-35)
 *
 * t_base never enters: the path only takes time differences inside a window.  Results are BIT-IDENTICAL to the SoA
 * entry points on the stream the records were packed from (tests/test_gpu_parity.py).
 *
 *  cmda_pack_events_p4   packs a device-resident SoA stream (t ascending) and builds ms_to_idx [n_ms + 1]
 *                        (entry n_ms = n); *d_status = number of events the format cannot hold (0 = lossless)
 *  d_rec                 the packed records the windows index: the whole store, or a staging buffer holding only
 *                        the windows' events (then h_win_src[s] = store index of window s's first event)
 *  d_ms_to_idx / h_ms_to_idx   the table on the device and a host copy (the per-window bucket bracket is found on
 *                        the host, the per-event bucket on the device); h_ms_to_idx[0] must be 0
 *  mode                  AUTO, FACTORED, BANDED or BANDED2 (the sensor-space formulation); others: CMDA_ERR_UNSUPPORTED
 *  everything else as cmda_events_vg_batch_planned; the workspace size is the same function of (events, S, ...). */
int cmda_pack_events_p4(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p, int64_t n,
                        uint32_t t_base_us, int64_t n_ms, uint32_t* d_rec, int64_t* d_ms_to_idx, int32_t* d_status,
                        void* stream);
int cmda_events_vg_batch_p4(const uint32_t* d_rec, const int64_t* d_ms_to_idx, const int64_t* h_ms_to_idx, int64_t n_ms,
                            const int64_t* h_win_start, const int64_t* h_win_end, const int64_t* h_win_src, int S,
                            const float* d_rectify_map, const int32_t* h_map_id, int H, int W, int B,
                            const float* h_clip, float final_range, int enforce_no_events_zero, int normalize,
                            float* d_out, float* d_raw_out, int64_t* d_bin_counts, void* d_workspace,
                            size_t workspace_bytes, int mode, const void* d_plans, void* stream);

/* ------------------------------------------------------------------------------------
 * The 3-byte WIRE form of the packed stream ("P3") -> P4 records on the device.
 * Replaces: nothing in the reference -- it is the host->device transport of the decoded events/{t,x,y,p} slices of
 * DSECDataset.get_events_vg (dsec.py:342-345) at 3 bytes per event; what the path computes on is the P4 stream above.
 *
 *   record3 = x | y << 10 | p << 19 | d << 20   (24 bits, little endian; x < 1024, y < 512, p in {0, 1})
 *   t_us    = t_base + 16 * j + d, j = the event's 16-microsecond bucket = the largest k with
 *             sub_to_idx[k] <= event index (sub_to_idx[k] = index of the first event with t_us - t_base >= 16 k;
 *             entry n_sub = n); t_base is the P4 stream's (a multiple of 1000 us)
 *
 *  d_rec3 / d_rec4   the record of store event `first` and where its P4 record goes: events [first, last) are
 *                    unpacked (a staging buffer that holds one copied range, or the whole store with first = 0)
 *  sub_lo, sub_hi    the buckets of event `first` and of event `last - 1` (found on the host's copy of the table)
 * The P4 records written are bit-identical to cmda_pack_events_p4's for the same stream. */
int cmda_unpack_p3_to_p4(const uint8_t* d_rec3, const int64_t* d_sub_to_idx, int64_t sub_lo, int64_t sub_hi, int64_t first,
                         int64_t last, uint32_t* d_rec4, void* stream);

/* ------------------------------------------------------------------------------------
 * K2+K3 with the post-voxel augmentation of the dataset fused into the normaliser's apply phase
 * (SURVEY.md 8 f-1).
 * Replaces: DSECDataset.__getitem__, events branch after get_events_vg,
 * /root/reference/mmseg/datasets/dsec.py:304-319 --
 *   mean over the bins (events_bins_5_avg_1, :304-305) -> crop [crop_h, crop_w] at (crop_x, crop_y) (:310)
 *   -> HorizontalFlip (:311-312) -> F.interpolate(size=(out_h, out_w), mode='bilinear',
 *   align_corners=False) (:313-315) -> repeat(3, 1, 1) (:318-319).
 * Test mode (:316-317) is crop (0, 0, W, 440), flip 0, out = crop size: the resize is then an exact copy.
 * The normalised [S, B, H, W] grid is never written: every output pixel normalises its four raw taps.
 *  h_aug      [S] per-window crop origin and flip flag (the augmentation draws of dsec.py:206-210)
 *  avg_bins   1: average the B bins first (output has 1 bin), 0: keep B bins
 *  repeat     copies along the channel axis (3 for enforce_3_channels, else 1)
 *  d_out      [S, repeat * Bo, out_h, out_w] float32, Bo = avg_bins ? 1 : B; channel r * Bo + b = bin b
 * Workspace: cmda_events_vg_workspace_bytes(...) + 4 * S * B * H * W bytes (the raw grid), unless
 * d_raw_out is given (then the raw grid goes there). */
typedef struct cmda_vg_augment {
    int32_t crop_x, crop_y;
    int32_t flip;
} cmda_vg_augment;

size_t cmda_events_vg_augmented_workspace_bytes(int64_t total_events, int S, int H, int W, int B, int mode);
int cmda_events_vg_augmented_batch(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p,
                                   const int64_t* h_win_start, const int64_t* h_win_end, int S,
                                   const float* d_rectify_map, const int32_t* h_map_id, int H, int W, int B,
                                   const float* h_clip, float final_range, int enforce_no_events_zero,
                                   const cmda_vg_augment* h_aug, int crop_w, int crop_h, int out_w, int out_h,
                                   int avg_bins, int repeat, float* d_out, float* d_raw_out, int64_t* d_bin_counts,
                                   void* d_workspace, size_t workspace_bytes, int mode, const void* d_plans /* or NULL */,
                                   void* stream);

/* Replaces: events_to_voxel_grid(time, x, y, pol, width, height, num_bins),
 * /root/reference/mmseg/datasets/dsec.py:26-58 (normalize_flag=False), on float32 SoA
 * inputs that are already rectified.  d_grid is [B, H, W] float32. */
int cmda_voxel_grid_f32(const float* d_time, const float* d_x, const float* d_y, const float* d_pol, int64_t n,
                        int W, int H, int B, float* d_grid, int64_t* d_bin_counts,
                        void* d_workspace, size_t workspace_bytes, int mode, void* stream);

/* Replaces: events_norm(events, clip_range, final_range, enforce_no_events_zero),
 * /root/reference/mmseg/datasets/dsec.py:80-121 (numeric clip_range), for S independent
 * grids of `voxels` elements each; in place on d_grid. */
size_t cmda_events_norm_workspace_bytes(int S);
int cmda_events_norm_batch(float* d_grid, int S, int64_t voxels, const float* h_clip, float final_range,
                           int enforce_no_events_zero, void* d_workspace, size_t workspace_bytes, void* stream);

/* Integer side outputs for parity checks (bit-exact against dsec.py:41-43, 347-355):
 * per event of ONE window [start, end): rectified float coordinates, t_norm, and the
 * truncated integer corner origin (INT32_MIN where the reference's .int() is
 * 'integer indefinite').  Any output pointer may be NULL. */
int cmda_remap_events(const uint32_t* d_t, const uint16_t* d_x, const uint16_t* d_y, const uint8_t* d_p,
                      int64_t start, int64_t end, const float* d_rectify_map, int H, int W, int B,
                      float* d_xr, float* d_yr, float* d_tn, int32_t* d_x0, int32_t* d_y0, int32_t* d_t0,
                      void* stream);

/* ------------------------------------------------------------------------------------
 * K4  frame-pair pseudo-events.
 * Replaces: get_image_change(image_now, image_front),
 * /root/reference/create_cityscapes_image_change.py:16-35.
 *  d_now/d_front  [S, H, W] uint8 gray
 *  h_lut          256 floats = np.log(arange(256, float32) + log_add), computed by the
 *                 caller with numpy exactly as the reference evaluates it
 *  thr, clip      float32(threshold), float32(clip_range)
 *  d_out_f32      optional [S, H, W] float32 in [-1, 1] (line 31)
 *  d_out_u8       optional [S, H, W] uint8 = uint8(around((d+1)/2*255)) (line 33) */
size_t cmda_image_workspace_bytes(int S, int H, int W, int channels);
int cmda_logdiff_pair_u8(const uint8_t* d_now, const uint8_t* d_front, int S, int H, int W, const float* h_lut,
                         float thr, float clip, float* d_out_f32, uint8_t* d_out_u8,
                         void* d_workspace, size_t workspace_bytes, void* stream);

/* K5  shift-pair pseudo-events (ISR).
 * Replaces: get_image_change_from_pil + get_ic, /root/reference/mmseg/datasets/utils.py:87-152.
 *  d_img      [S, H, W] gray (channels=1) or [S, H, W, 3] RGB (channels=3; converted with
 *             PIL's 'L' formula (19595 R + 38470 G + 7471 B + 32768) >> 16, utils.py:126)
 *  h_lut      256 floats = np.log(arange(256, float32)/255*(v1-v0)+v0) (utils.py:88-91)
 *  thr, clip  float32((ln v1 - ln v0)*threshold), float32((ln v1 - ln v0)*clip_range)
 *  d_out      [S, H, W] float32 */
int cmda_isr_shift_u8(const uint8_t* d_img, int channels, int S, int H, int W, int shift_pixel, int direction,
                      const float* h_lut, float thr, float clip, float* d_out,
                      void* d_workspace, size_t workspace_bytes, void* stream);

/* Mixed-image ISR inside the train step without the GPU -> CPU -> PIL -> GPU bounce.
 * Replaces: /root/reference/mmseg/models/uda/dacs.py:730-733 --
 *   clamp(denorm(img, mean, std), 0, 1) * 255 -> np.uint8 -> Image.fromarray -> convert('L')
 * (denorm = img.mul(std).add(mean) / 255.0, mmseg/models/utils/dacs_transforms.py:52-53).
 *  d_img   [S, 3, H, W] float32, the normalised image batch as the model sees it
 *  mean, stdv   3 floats each (img_norm_cfg mean / std), in HOST or DEVICE memory (both in the same kind;
 *               told apart with cudaPointerGetAttributes).  The reference keeps them in CUDA tensors
 *               (dacs_transforms.py:38-49): passing those pointers avoids a synchronising read-back.
 *  d_gray  [S, H, W] uint8 'L' plane: feed it to cmda_isr_shift_u8(channels = 1)
 *  d_rgb   optional [S, H, W, 3] uint8, the bytes PIL would have been handed (may be NULL) */
int cmda_denorm_rgb_to_gray_u8(const float* d_img, int S, int H, int W, const float* mean, const float* stdv,
                               uint8_t* d_gray, uint8_t* d_rgb, void* stream);

/* ------------------------------------------------------------------------------------
 * Cityscapes source branch of the loader on the device.
 * cmda_resize_bilinear_u8 replaces PIL's Image.resize(size, Image.BILINEAR) on 8-bit 'L' (channels = 1) or
 * 'RGB' (channels = 3, interleaved) images, bit for bit
 * (/root/reference/mmseg/datasets/cityscapes_ic.py:152-153, 175-176: raw_image / image-change PNG -> 1024 x 512).
 *  d_src [S, H, W, channels] uint8 -> d_dst [S, out_h, out_w, channels] uint8.
 *  Down-scaling by more than 31x along an axis (more than 64 filter taps) returns CMDA_ERR_UNSUPPORTED.
 * cmda_u8_crop_to_centered_f32 replaces cityscapes_ic.py:177-183, 207-209 on a gray image:
 *   crop(box=(x, y, x + crop_w, y + crop_h)) -> HorizontalFlip -> float32 -> (v / 255.0 - 0.5) / 0.5 -> repeat(3, 1, 1)
 *  d_src [S, H, W] uint8, h_aug [S] crop origin + flip, d_out [S, repeat, crop_h, crop_w] float32. */
size_t cmda_resize_bilinear_workspace_bytes(int S, int H, int W, int channels, int out_h, int out_w);
int cmda_resize_bilinear_u8(const uint8_t* d_src, int channels, int S, int H, int W, int out_h, int out_w, uint8_t* d_dst,
                            void* d_workspace, size_t workspace_bytes, void* stream);
int cmda_u8_crop_to_centered_f32(const uint8_t* d_src, int S, int H, int W, const cmda_vg_augment* h_aug, int crop_w,
                                 int crop_h, int repeat, float* d_out, void* stream);

/* PIL 'L' conversion alone (parity of the integer stage). d_rgb [n,3] -> d_gray [n]. */
int cmda_rgb_to_gray_u8(const uint8_t* d_rgb, int64_t n_pixels, uint8_t* d_gray, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CMDA_B200_H */
