// api.cu -- library identification, error strings, launch counter.
#include <atomic>

#include "common.cuh"

namespace cnsn {
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
}  // namespace cnsn

extern "C" int cnsn_version(void) { return CNSN_ABI_VERSION; }

extern "C" unsigned long long cnsn_launch_count(void) {
    return cnsn::g_launches.load(std::memory_order_relaxed);
}

extern "C" const char* cnsn_error_string(int code) {
    switch (code) {
        case CNSN_OK: return "ok";
        case CNSN_E_BADARG: return "cnsn: bad argument (null pointer, non-positive dim, bad dtype or window)";
        case CNSN_E_WORKSPACE: return "cnsn: workspace too small";
        case CNSN_E_BATCH1: return "Expected more than 1 value per channel when training";
        case CNSN_E_ALIGN: return "cnsn: tensor pointer not aligned to its element size";
        case CNSN_E_UNSUPPORTED: return "cnsn: shape not supported by this operator (fused site: planes must be multiples of 16 bytes and a channel must fit on chip)";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "cnsn: unknown error";
}
