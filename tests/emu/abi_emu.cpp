// TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h): what the emulated build of the C ABI needs besides the
// translation units themselves -- the fiber switch, the kernels' dynamic shared memory, and "unsupported" answers for
// the two voxel modes that cannot be emulated (TILED: inline-PTX shared-memory atomics; EXACT: cub's radix sort).
#define EMU_DEFINE_SWITCH 1
#include "common.cuh"

namespace cmda {
alignas(16) unsigned char s_band_raw[256 * 1024];
alignas(16) unsigned s_band_acc[64 * 1024];
alignas(16) double s_planes[32 * 1024];
alignas(16) unsigned int s_bins[1024];

size_t tiled_workspace_bytes(int64_t, int, int, int, int) { return 0; }
int tiled_supported(int, int, int) { return 0; }
int launch_tiled_scatter(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, const WindowTable&, int, const float*, int,
                         int, int, long long*, int64_t*, void*, size_t, cudaStream_t) { return CMDA_ERR_UNSUPPORTED; }
int exact_supported(int, int, int) { return 0; }
size_t exact_workspace_bytes(long long) { return 0; }
int launch_exact_raw(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, const WindowTable&, int, const float*, int, int,
                     int, float*, int64_t*, void*, size_t, cudaStream_t) { return CMDA_ERR_UNSUPPORTED; }
int launch_exact_f32(const float*, const float*, const float*, const float*, long long, int, int, int, float*, int64_t*, void*, size_t,
                     cudaStream_t) { return CMDA_ERR_UNSUPPORTED; }
}  // namespace cmda
