// TEST INFRASTRUCTURE ONLY (see include/cuda_runtime.h): what the emulated build of the C ABI needs besides the
// translation units themselves -- the fiber switch and the kernels' dynamic shared memory.
#define EMU_DEFINE_SWITCH 1
#include "common.cuh"

namespace cmda {
alignas(16) unsigned char s_band_raw[256 * 1024];
alignas(16) unsigned s_band_acc[64 * 1024];
alignas(16) double s_planes[32 * 1024];
alignas(16) unsigned int s_bins[1024];

alignas(16) unsigned s_hist[8192];
alignas(16) unsigned char s_raw[256 * 1024];
alignas(16) float s_tab[64 * 1024];                 // pair tables of the table-driven pseudo-event kernels
}  // namespace cmda

char* emu_shared_window = reinterpret_cast<char*>(cmda::s_raw);      // TILED addresses its accumulators through the shared window
