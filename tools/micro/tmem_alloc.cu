// Does tcgen05.alloc let 4 CTAs per SM hold 128 TMEM columns each at the same time?  Each CTA allocates COLS columns,
// holds them for ~50 us, frees them; the time it waited for its allocation is recorded.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/micro/tmem_alloc tools/micro/tmem_alloc.cu && tools/micro/tmem_alloc
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ unsigned long long gtime() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int COLS>
__global__ void __launch_bounds__(128) k(unsigned long long* out, int hold_us) {
    extern __shared__ unsigned char dsm[];
    __shared__ unsigned s_tmem;
    const int warp = threadIdx.x >> 5;
    unsigned long long t0 = gtime();
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared(&s_tmem)), "r"(COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    unsigned long long t1 = gtime();
    const unsigned base = s_tmem;
    // a few st / ld round trips, timed
    const unsigned trow = base + ((unsigned)(warp & 3) << 21);
    unsigned r0 = threadIdx.x, r1 = 1, r2 = 2, r3 = 3;
    unsigned long long t2 = gtime();
    for (int it = 0; it < 64; ++it) {
        for (int c = 0; c < COLS; c += 4)
            asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(trow + c), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        for (int c = 0; c < COLS; c += 4) {
            unsigned a, b, d, e;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(d), "=r"(e) : "r"(trow + c) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            r0 += a; r1 += b; r2 += d; r3 += e;
        }
    }
    unsigned long long t3 = gtime();
    while (gtime() - t1 < (unsigned long long)hold_us * 1000ull) {}
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(COLS) : "memory");
    if (threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        out[blockIdx.x * 4 + 0] = smid;
        out[blockIdx.x * 4 + 1] = t1 - t0;
        out[blockIdx.x * 4 + 2] = (t3 - t2) + (r0 + r1 + r2 + r3 == 12345u);
        out[blockIdx.x * 4 + 3] = t0;
    }
}

template <int COLS>
void run(int per_sm, size_t dsmem) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int grid = sms * per_sm;
    unsigned long long* d;
    cudaMalloc(&d, grid * 4 * sizeof(unsigned long long));
    cudaFuncSetAttribute(k<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsmem);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<COLS>, 128, dsmem);
    k<COLS><<<grid, 128, dsmem>>>(d, 50);
    cudaError_t e = cudaDeviceSynchronize();
    unsigned long long* h = (unsigned long long*)malloc(grid * 4 * sizeof(unsigned long long));
    cudaMemcpy(h, d, grid * 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    unsigned long long worst = 0, sum = 0, rt = 0;
    int late = 0;
    for (int i = 0; i < grid; ++i) {
        worst = h[i * 4 + 1] > worst ? h[i * 4 + 1] : worst;
        sum += h[i * 4 + 1];
        rt += h[i * 4 + 2];
        late += h[i * 4 + 1] > 10000;
    }
    printf("COLS %3d, %d CTAs/SM asked (occupancy %d), smem %zu: %s | alloc wait mean %.1f us, worst %.1f us, CTAs that waited > 10 us: %d of %d | 64 x (st + ld of %d cols): %.2f us per CTA\n",
           COLS, per_sm, occ, dsmem, cudaGetErrorString(e), sum / 1e3 / grid, worst / 1e3, late, grid, COLS, rt / 1e3 / grid);
    cudaFree(d);
    free(h);
}

int main() {
    run<128>(4, 50 * 1024);
    run<128>(3, 50 * 1024);
    run<64>(4, 50 * 1024);
    run<64>(8, 24 * 1024);
    run<32>(8, 24 * 1024);
    run<256>(2, 100 * 1024);
    return 0;
}
