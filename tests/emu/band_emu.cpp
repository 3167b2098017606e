// TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h): runs stage A of the FACTORED voxel path -- the L2-RED kernel,
// the BANDED kernels and their second cut -- as host fibers on small inputs and returns the sensor-space grid R, so
// that the variants can be compared bit for bit where no GPU exists.  The kernel code is the transformed copy of
// cmda_b200/csrc/voxel_factored.cu (everything above its host launch section) made by build_emu.py; the launch
// bookkeeping below restates launch_factored's.
#define EMU_DEFINE_SWITCH 1
#include "gen/voxel_factored_kernels.inc"

namespace cmda {
thread_local int g_last_cuda_error = 0;
thread_local PhaseTimer g_phase_timer = {nullptr, 0, 0};
// the kernels' dynamic shared memory
alignas(16) unsigned char s_band_raw[256 * 1024];
alignas(16) unsigned s_band_acc[64 * 1024];
double s_planes[32 * 1024];
}  // namespace cmda

using namespace cmda;

// capacity guard buffers of the stage-A kernels (sketch of the RED path, flags of the BANDED cuts): scratch here,
// the guard itself is exercised through the C ABI (tests/test_emu_abi.py)
static unsigned g_flags[kMaxWindows];

template <bool HAS_T, bool VEC, int CUT>
static void run_banded(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, const WindowTable& tab,
                       const BandTable& bt, const BandGeom& g, int S, int H, int W, int B, long long max_chunks,
                       unsigned* table, unsigned* rec32, unsigned char* rec8, unsigned short* rec16, void* R,
                       unsigned long long* bins) {
    if (max_chunks > 0)
        emu_launch(dim3(static_cast<unsigned>(max_chunks), S), dim3(CUT == 1 ? kBandPartThreads : kPart3Threads), [&] {
            if (CUT == 1) band_partition_kernel<HAS_T, VEC, false>(t, x, y, p, PackedSrc{nullptr, nullptr, 0}, tab, bt, g, H, W, B, table, rec32, rec8, rec16, bins, g_flags);
            else band_partition3_kernel<HAS_T, VEC, false>(t, x, y, p, PackedSrc{nullptr, nullptr, 0}, tab, bt, g, H, W, B, table, rec32, rec8, rec16, bins, g_flags);
        });
    emu_launch(dim3(static_cast<unsigned>(S) * g.nbuckets), dim3(kBandAccThreads), [&] {
        band_accumulate_kernel<HAS_T>(table, rec32, rec8, rec16, bt, g, H, W, B, R, g_flags);      // both cuts share the accumulate pass
    });
}

template <bool HAS_T, bool VEC>
static void run_red(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, const WindowTable& tab, int S,
                    long long max_events, int H, int W, int B, void* R, unsigned long long* bins) {
    if (max_events <= 0) return;
    const long long groups = (max_events + 7) / 8 + 1;
    const long long per = static_cast<long long>(kSensThreads) * kSensGroupsPerThread;
    emu_launch(dim3(static_cast<unsigned>((groups + per - 1) / per), S), dim3(kSensThreads),
               [&] { sensor_accumulate_kernel<HAS_T, VEC, true, false>(t, x, y, p, PackedSrc{nullptr, nullptr, 0}, tab, H, W, B, R, bins, Guard{g_flags, B == 1 ? kCellLimit32 : kCellLimit64}); });
}

// variant 0: sensor_accumulate_kernel (R zeroed here, like the memset of launch_factored); 1: BANDED; 2: BANDED second cut.
// R: int64 [S][B][H][W] for B > 1, int32 [S][H][W] for B == 1 -- prefilled with garbage by the caller for variants 1 / 2,
// which must store every cell.  bins: [S][B], zeroed by the caller.  Returns 0, or -1 for an unsupported shape.
extern "C" int emu_stage_a(const uint32_t* t, const uint16_t* x, const uint16_t* y, const uint8_t* p, const int64_t* starts,
                           const int64_t* ends, int S, int H, int W, int B, int variant, void* R, unsigned long long* bins) {
    if (S < 1 || S > kMaxWindows) return -1;
    WindowTable tab{};
    long long max_events = 0;
    for (int s = 0; s < S; ++s) {
        tab.w[s].start = starts[s];
        tab.w[s].end = ends[s] > starts[s] ? ends[s] : starts[s];
        const long long n = tab.w[s].end - tab.w[s].start;
        if (n > max_events) max_events = n;
    }
    const bool vec = ((reinterpret_cast<uintptr_t>(t) & 15) == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(y) & 15) == 0) && ((reinterpret_cast<uintptr_t>(p) & 7) == 0);
    const size_t npx = static_cast<size_t>(H) * W;
    if (variant == 0) {
        std::memset(R, 0, B == 1 ? sizeof(int) * S * npx : sizeof(long long) * S * B * npx);
        if (B == 1) { if (vec) run_red<false, true>(t, x, y, p, tab, S, max_events, H, W, B, R, bins); else run_red<false, false>(t, x, y, p, tab, S, max_events, H, W, B, R, bins); }
        else { if (vec) run_red<true, true>(t, x, y, p, tab, S, max_events, H, W, B, R, bins); else run_red<true, false>(t, x, y, p, tab, S, max_events, H, W, B, R, bins); }
        return 0;
    }
    BandGeom g{};
    if (!pick_band_geom(H, W, B, g)) return -1;
    if (static_cast<size_t>(g.rows) * W * (B > 1 ? 8 : 4) > sizeof(s_band_acc)) return -1;
    BandTable bt{};
    long long chunks = 0, max_chunks = 0;
    for (int s = 0; s < S; ++s) {
        long long n = 0;
        if (tab.w[s].end > tab.w[s].start) {
            const long long groups = ((tab.w[s].end + 7) >> 3) - (tab.w[s].start >> 3);
            const long long per = static_cast<long long>(kBandPartThreads) * kBandPartGroups;
            n = (groups + per - 1) / per;
        }
        bt.rec_base[s] = chunks * kBandChunk;
        bt.chunk_base[s] = static_cast<int>(chunks);
        bt.nchunks[s] = static_cast<int>(n);
        chunks += n;
        if (n > max_chunks) max_chunks = n;
    }
    const size_t slots = static_cast<size_t>(chunks > 0 ? chunks : 1) * kBandChunk;
    // garbage-filled scratch: nothing may rely on zeroed workspace
    std::vector<unsigned> table(static_cast<size_t>(chunks > 0 ? chunks : 1) * (g.nbuckets + 2), 0xdeadbeefu);
    std::vector<unsigned> rec32(B > 1 ? slots : 4, 0xa5a5a5a5u);
    std::vector<unsigned char> rec8(B > 1 ? slots : 4, 0x5a);
    std::vector<unsigned short> rec16(B == 1 ? slots : 4, 0xa5a5);
#define RUN(HAS_T, VEC)                                                                                                        \
    do {                                                                                                                       \
        if (variant == 1) run_banded<HAS_T, VEC, 1>(t, x, y, p, tab, bt, g, S, H, W, B, max_chunks, table.data(), rec32.data(), rec8.data(), rec16.data(), R, bins); \
        else run_banded<HAS_T, VEC, 2>(t, x, y, p, tab, bt, g, S, H, W, B, max_chunks, table.data(), rec32.data(), rec8.data(), rec16.data(), R, bins); \
    } while (0)
    if (B == 1) { if (vec) RUN(false, true); else RUN(false, false); }
    else { if (vec) RUN(true, true); else RUN(true, false); }
#undef RUN
    return 0;
}

char* emu_shared_window = nullptr;      // no kernel of this library uses shared-window addresses
