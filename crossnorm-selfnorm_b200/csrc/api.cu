// api.cu -- library identification, error strings, launch counter, tuning knobs, asynchronous error word.
#include <string.h>

#include <atomic>
#include <mutex>

#include "flow_common.cuh"

namespace cnsn {
static std::atomic<unsigned long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

Knobs& knobs() {
    static Knobs k;
    return k;
}

// One pinned, device-mapped word per process (portable: every device of the process sees it under UVA).  Allocated
// on first use -- a resident-kernel launch -- which is before any stream capture a caller may start later.
static std::once_flag g_err_once;
static unsigned* g_err_word = nullptr;
unsigned* async_error_word() {
    std::call_once(g_err_once, [] {
        void* p = nullptr;
        if (cudaHostAlloc(&p, 64, cudaHostAllocPortable | cudaHostAllocMapped) == cudaSuccess) {
            memset(p, 0, 64);
            g_err_word = static_cast<unsigned*>(p);
        } else {
            (void)cudaGetLastError();
        }
    });
    return g_err_word;
}
int async_error_peek() {
    return g_err_word ? (int)*reinterpret_cast<volatile unsigned*>(g_err_word) : 0;
}
}  // namespace cnsn

extern "C" int cnsn_version(void) { return CNSN_ABI_VERSION; }

extern "C" unsigned long long cnsn_launch_count(void) {
    return cnsn::g_launches.load(std::memory_order_relaxed);
}

extern "C" int cnsn_async_error(int clear) {
    const int v = cnsn::async_error_peek();
    if (v && clear) *reinterpret_cast<volatile unsigned*>(cnsn::g_err_word) = 0u;
    return v ? CNSN_E_TIMEOUT : CNSN_OK;
}

extern "C" int cnsn_tune(const char* name, int value) {
    if (!name) return CNSN_E_BADARG;
    cnsn::Knobs& k = cnsn::knobs();
    if (!strcmp(name, "reset")) { k = cnsn::Knobs(); return CNSN_OK; }
#define CNSN_KNOB(field) if (!strcmp(name, #field)) { k.field = value; return CNSN_OK; }
    CNSN_KNOB(selfnorm_impl) CNSN_KNOB(crossnorm_impl) CNSN_KNOB(flow_mode) CNSN_KNOB(flow_bwd) CNSN_KNOB(flow_d)
    CNSN_KNOB(flow_tpi) CNSN_KNOB(flow_batches) CNSN_KNOB(lookahead_mb) CNSN_KNOB(item_kb) CNSN_KNOB(grp_kb)
    CNSN_KNOB(keep) CNSN_KNOB(pf) CNSN_KNOB(rpf) CNSN_KNOB(poll_ns) CNSN_KNOB(i3) CNSN_KNOB(cooperative)
    CNSN_KNOB(grid_cap) CNSN_KNOB(debug) CNSN_KNOB(trace) CNSN_KNOB(tm) CNSN_KNOB(tm_items)
#undef CNSN_KNOB
    return CNSN_E_BADARG;
}

extern "C" const char* cnsn_error_string(int code) {
    switch (code) {
        case CNSN_OK: return "ok";
        case CNSN_E_BADARG: return "cnsn: bad argument (null pointer, non-positive dim, bad dtype or window)";
        case CNSN_E_WORKSPACE: return "cnsn: workspace too small";
        case CNSN_E_BATCH1: return "Expected more than 1 value per channel when training";
        case CNSN_E_ALIGN: return "cnsn: tensor pointer not aligned to its element size";
        case CNSN_E_UNSUPPORTED: return "cnsn: shape not supported by this operator (fused site: planes must be multiples of 16 bytes and a channel must fit on chip)";
        case CNSN_E_TIMEOUT: return "cnsn: an earlier kernel of this process gave up waiting for a peer CTA or a bulk copy (its results are undefined); cnsn_async_error(1) clears the condition";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "cnsn: unknown error";
}
