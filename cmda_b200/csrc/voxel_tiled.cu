// K2 (mode TILED) -- placeholder until the band-partitioned shared-memory path lands.
#include "common.cuh"

namespace cmda {

size_t tiled_workspace_bytes(int64_t, int, int, int, int) { return 0; }
int tiled_supported(int, int, int) { return 0; }
int launch_tiled_raw(const uint32_t*, const uint16_t*, const uint16_t*, const uint8_t*, const WindowTable&, int,
                     const float*, int, int, int, float*, PartialStats*, int64_t*, void*, size_t, cudaStream_t) {
    return CMDA_ERR_UNSUPPORTED;
}
int launch_tiled_f32(const float*, const float*, const float*, const float*, long long, int, int, int, float*,
                     PartialStats*, int64_t*, void*, size_t, cudaStream_t) {
    return CMDA_ERR_UNSUPPORTED;
}

}  // namespace cmda
