// TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h): the voxel path of libcmda_b200 -- launch_factored (plans,
// stage A in its three forms, fused gather) and launch_norm_apply, i.e. the REAL host launch code of
// cmda_b200/csrc/voxel_factored.cu and norm.cu with every kernel under it -- executed on the host through the fiber
// emulation, for one launch group of windows.  The carving of the workspace restates events_vg_impl (api.cu).
#define EMU_DEFINE_SWITCH 1
#include "common.cuh"

namespace cmda {
thread_local int g_last_cuda_error = 0;
thread_local PhaseTimer g_phase_timer = {nullptr, 0, 0};
// the kernels' dynamic shared memory
alignas(16) unsigned char s_band_raw[256 * 1024];
alignas(16) unsigned s_band_acc[64 * 1024];
alignas(16) double s_planes[32 * 1024];

int factored_supported(int H, int W, int B);
size_t factored_scratch_bytes(int group, int H, int W, int B);
int banded_supported(int H, int W, int B);
size_t banded_scratch_bytes(long long total_events, int group, int H, int W, int B);
int launch_factored(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, const PackedSrc*, const WindowTable&, int, long long,
                    const float*, int, int, int, void*, int64_t*, float*, PartialStats*, void*, size_t, const void*, int, cudaStream_t);
int launch_norm_apply(const float*, float*, int, long long, const PartialStats*, const WindowTable&, float, int, cudaStream_t);
}  // namespace cmda

using namespace cmda;

// out / raw: [S][B][H][W] float32 (raw may be NULL); bins: [S][B] (may be NULL); maps: [n_maps][H][W][2] or NULL;
// map_ids: [S] or NULL.  banded: 0 = FACTORED (L2 RED stage A), 1 = BANDED, 2 = its second cut.
extern "C" int emu_events_vg(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, const int64_t* starts,
                             const int64_t* ends, int S, const float* maps, const int32_t* map_ids, int H, int W, int B,
                             const float* clips, int banded, int normalize, float* out, float* raw, int64_t* bins) {
    if (S < 1 || S > kMaxWindows || !factored_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    if (banded && !banded_supported(H, W, B)) return CMDA_ERR_UNSUPPORTED;
    WindowTable tab{};
    long long max_events = 0, total = 0;
    for (int s = 0; s < S; ++s) {
        tab.w[s].start = starts[s];
        tab.w[s].end = ends[s] > starts[s] ? ends[s] : starts[s];
        tab.w[s].map_id = map_ids ? map_ids[s] : 0;
        tab.w[s].clip = clips ? clips[s] : 1.0f;
        const long long n = tab.w[s].end - tab.w[s].start;
        total += n;
        if (n > max_events) max_events = n;
    }
    const size_t V = static_cast<size_t>(B) * H * W;
    const size_t stats = align_up(sizeof(PartialStats) * kStatBlocks * static_cast<size_t>(S), 256);
    const size_t ab = align_up(sizeof(long long) * static_cast<size_t>(S) * V, 256);
    const size_t scratch = factored_scratch_bytes(S, H, W, B) + 256 + (banded ? banded_scratch_bytes(total, S, H, W, B) : 0);
    const size_t bytes = stats + ab + scratch + 256;
    char* ws = static_cast<char*>(std::aligned_alloc(256, align_up(bytes, 256)));
    if (!ws) return CMDA_ERR_WORKSPACE;
    std::memset(ws, 0xa5, bytes);                                   // nothing may rely on a zeroed workspace
    PartialStats* partials = reinterpret_cast<PartialStats*>(ws);
    if (bins) std::memset(bins, 0, sizeof(int64_t) * S * B);
    float* raw_g = (normalize && raw) ? raw : out;
    int rc = launch_factored(t, x, y, p, nullptr, tab, S, max_events, maps, H, W, B, ws + stats, bins, raw_g, partials, ws + stats + ab,
                             scratch, nullptr, banded, nullptr);
    if (rc == CMDA_OK && normalize) rc = launch_norm_apply(raw_g, out, S, static_cast<long long>(V), partials, tab, 1.0f, 1, nullptr);
    std::free(ws);
    return rc;
}

char* emu_shared_window = nullptr;      // no kernel of this library uses shared-window addresses
