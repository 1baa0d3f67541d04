// Microbenchmark: does the TRAVERSAL ORDER of an NCHW tensor bound HBM throughput?
// y = 2*x over (N,C,M) fp32 planes with one CTA (256 threads) per group of planes, three orders:
//   seq   : CTA b handles planes b*P .. b*P+P-1 in memory order (what the three-kernel path does)
//   chan  : channel-major -- CTA b handles channel c = b / (N/P), samples n = (b % (N/P))*P .. : the order every
//           fused SelfNorm kernel is forced into (all N planes of a channel before the next channel)
//   chan kk: channel groups of kk adjacent channels (contiguous runs of kk planes per sample)
// Also: read-only (sum) variants, to separate read and read+write behaviour.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint4 ldg_stream(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" :: "l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w));
}

// mode 0 seq, 1 channel-major with groups of kk channels.  P planes per CTA, TPI = 256/P threads per plane.
template <int P, bool WRITE>
__global__ void __launch_bounds__(256) k_order(const float* x, float* y, float* sink, int N, int C, int M, int mode, int kk) {
    constexpr int TPI = 256 / P;
    const int nv = M / 4;
    const int sub = threadIdx.x / TPI, r = threadIdx.x % TPI;
    long long plane;
    if (mode == 0) {
        plane = (long long)blockIdx.x * P + sub;
    } else {
        // virtual index v inside group g: v = n*kk + cl
        const long long per_group = (long long)N * kk;
        const long long item = (long long)blockIdx.x * P + sub;
        const long long g = item / per_group, v = item - g * per_group;
        const long long n = v / kk, cl = v - n * kk;
        plane = n * C + g * kk + cl;
    }
    if (plane >= (long long)N * C) return;
    const uint4* px = reinterpret_cast<const uint4*>(x + plane * M);
    uint4* py = reinterpret_cast<uint4*>(y + plane * M);
    float acc = 0.f;
    for (int i0 = r; i0 < nv; i0 += TPI * 4) {
        uint4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i0 + u * TPI < nv) v[u] = ldg_stream(px + i0 + u * TPI);
#pragma unroll
        for (int u = 0; u < 4; ++u) if (i0 + u * TPI < nv) {
            if (WRITE) {
                uint4 o;
                o.x = __float_as_uint(2.f * __uint_as_float(v[u].x)); o.y = __float_as_uint(2.f * __uint_as_float(v[u].y));
                o.z = __float_as_uint(2.f * __uint_as_float(v[u].z)); o.w = __float_as_uint(2.f * __uint_as_float(v[u].w));
                stg_stream(py + i0 + u * TPI, o);
            } else {
                acc += __uint_as_float(v[u].x) + __uint_as_float(v[u].y) + __uint_as_float(v[u].z) + __uint_as_float(v[u].w);
            }
        }
    }
    if (!WRITE && acc == 123.456f) *sink = acc;
}

template <int P, bool WRITE>
static void run(const char* name, const float* x, float* y, float* sink, int N, int C, int M, int mode, int kk) {
    const long long planes = (long long)N * C;
    const unsigned grid = (unsigned)((planes + P - 1) / P);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) k_order<P, WRITE><<<grid, 256>>>(x, y, sink, N, C, M, mode, kk);
    cudaEventRecord(a);
    const int reps = 10;
    for (int i = 0; i < reps; ++i) k_order<P, WRITE><<<grid, 256>>>(x, y, sink, N, C, M, mode, kk);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    ms /= reps;
    const double bytes = (double)planes * M * 4 * (WRITE ? 2 : 1);
    printf("%-28s P=%d %s  %.3f ms  %.0f GB/s  (%s)\n", name, P, WRITE ? "copy" : "read", ms, bytes / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 256, C = argc > 2 ? atoi(argv[2]) : 256, M = argc > 3 ? atoi(argv[3]) : 3136;
    float *x, *y, *sink;
    const size_t bytes = (size_t)N * C * M * 4;
    cudaMalloc(&x, bytes); cudaMalloc(&y, bytes); cudaMalloc(&sink, 4);
    cudaMemset(x, 0, bytes); cudaMemset(y, 0, bytes);
    printf("N=%d C=%d M=%d  (%.0f MB per tensor)\n", N, C, M, bytes / 1e6);
    run<8, true>("seq", x, y, sink, N, C, M, 0, 1);
    run<8, true>("chan-major kk=1", x, y, sink, N, C, M, 1, 1);
    run<8, true>("chan-major kk=2", x, y, sink, N, C, M, 1, 2);
    run<8, true>("chan-major kk=4", x, y, sink, N, C, M, 1, 4);
    run<8, true>("chan-major kk=8", x, y, sink, N, C, M, 1, 8);
    run<8, true>("chan-major kk=16", x, y, sink, N, C, M, 1, 16);
    run<4, true>("seq", x, y, sink, N, C, M, 0, 1);
    run<4, true>("chan-major kk=1", x, y, sink, N, C, M, 1, 1);
    run<4, true>("chan-major kk=4", x, y, sink, N, C, M, 1, 4);
    run<1, true>("seq", x, y, sink, N, C, M, 0, 1);
    run<1, true>("chan-major kk=1", x, y, sink, N, C, M, 1, 1);
    run<8, false>("seq", x, y, sink, N, C, M, 0, 1);
    run<8, false>("chan-major kk=1", x, y, sink, N, C, M, 1, 1);
    run<8, false>("chan-major kk=4", x, y, sink, N, C, M, 1, 4);
    run<4, false>("seq", x, y, sink, N, C, M, 0, 1);
    run<4, false>("chan-major kk=1", x, y, sink, N, C, M, 1, 1);
    return 0;
}
