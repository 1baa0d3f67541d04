// Microbenchmark: HBM -> shared memory throughput of 1-D cp.async.bulk (TMA bulk copy, SASS UBLKCP),
// one persistent CTA per SM, a ring of `stages` slots of `bytes` each, no compute on the data.
// Compared with a plain LDG streaming kernel over the same bytes.
#include <cstdio>
#include <cstdlib>
#include "../../crossnorm-selfnorm_b200/csrc/fused_common.cuh"
using namespace cnsn; using namespace cnsn::fused;
namespace cnsn { void note_launch() {} }

__global__ void __launch_bounds__(64, 1) k_bulk(const unsigned char* src, size_t total_bytes, unsigned bytes, int stages, int issuers) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + 64;
    unsigned char* data = smem + 1024;
    if (threadIdx.x == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t items = total_bytes / bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) {
            long long it = 0;
            for (size_t i = blockIdx.x; i < items; i += gridDim.x, ++it) {
                const int st = (int)(it % stages), ph = (int)((it / stages) & 1);
                mbar_wait(&empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&full[st], bytes);
                tma_load_1d_plain(data + (size_t)st * bytes, src + i * bytes, bytes, &full[st]);
            }
        }
    } else {
        if (lane == 0) {
            long long it = 0;
            for (size_t i = blockIdx.x; i < items; i += gridDim.x, ++it) {
                const int st = (int)(it % stages), ph = (int)((it / stages) & 1);
                mbar_wait(&full[st], ph);
                mbar_arrive(&empty[st]);
            }
        }
    }
}

__global__ void k_ldg(const uint4* src, size_t nvec, float* sink) {
    float acc = 0.f;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x * 4) {
        uint4 r[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { size_t j = i + (size_t)u * gridDim.x * blockDim.x; if (j < nvec) r[u] = ldg_stream(src + j); else r[u] = make_uint4(0,0,0,0); }
#pragma unroll
        for (int u = 0; u < 4; ++u) acc += __uint_as_float(r[u].x) + __uint_as_float(r[u].w);
    }
    if (acc == 123.456f) *sink = acc;
}

int main() {
    const size_t total = (size_t)822083584;
    unsigned char* src; float* sink;
    cudaMalloc(&src, total); cudaMemset(src, 1, total); cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const unsigned sizes[] = {2048, 4096, 12544, 25088, 50176};
    for (unsigned bytes : sizes) {
        for (int depth_kb : {64, 200}) {
            int stages = depth_kb * 1024 / bytes; if (stages > 64) stages = 64; if (stages < 2) continue;
            float best = 1e9f;
            for (int rep = 0; rep < 3; ++rep) {
                cudaEventRecord(e0);
                k_bulk<<<148, 64, 1024 + (size_t)stages * bytes>>>(src, total, bytes, stages, 1);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("bulk copy %6u B x %2d stages (%3d KB in flight/SM): %.3f ms  %.0f GB/s  (%s)\n", bytes, stages, stages * bytes / 1024, best, total / best / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        k_ldg<<<148 * 8, 256>>>((const uint4*)src, total / 16, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    printf("LDG.128 streaming read (8 CTAs/SM x 256 thr, 4 in flight): %.3f ms  %.0f GB/s\n", best, total / best / 1e6);
    return 0;
}
