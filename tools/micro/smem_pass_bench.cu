// Microbenchmark: cycles for one warp to run the fused kernel's smem statistics (smem_mean_m2) over one
// 56x56 fp32 plane, alone and with other warps doing the same on their own planes.
#include <cstdio>
#include "../../crossnorm-selfnorm_b200/csrc/fused_common.cuh"
using namespace cnsn; using namespace cnsn::fused;
namespace cnsn { void note_launch() {} }

__global__ void k(int nwarps, int M, long long* out, float* sink) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* data = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < 16 * M; i += blockDim.x) data[i] = (float)(i % 97) * 0.01f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp < nwarps) {
        long long t0 = clock64();
        float2 r = smem_mean_m2<float>(data + warp * M, M, lane, 32, true, true);
        long long t1 = clock64();
        float sdv = sqrtf(r.y / (float)(M - 1) + 1e-12f);
        long long t2 = clock64();
        if (lane == 0) { out[blockIdx.x * 48 + warp * 3] = t1 - t0; out[blockIdx.x * 48 + warp * 3 + 1] = t2 - t1; sink[blockIdx.x * 16 + warp] = r.x + sdv; }
    }
}
int main() {
    const int M = 3136;
    long long* out; float* sink;
    cudaMalloc(&out, 148 * 48 * 8); cudaMalloc(&sink, 148 * 16 * 4);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * M * 4);
    for (int nw : {1, 2, 4, 8, 16}) {
        for (int rep = 0; rep < 2; ++rep) {
            k<<<148, 512, 16 * M * 4>>>(nw, M, out, sink);
            cudaDeviceSynchronize();
        }
        long long h[48];
        cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
        printf("warps=%2d smem_mean_m2 cycles(warp0)=%lld (last)=%lld  sqrt/div tail=%lld  err=%s\n", nw, h[0], h[(nw - 1) * 3], h[1], cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
