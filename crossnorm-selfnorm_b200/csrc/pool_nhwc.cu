// pool_nhwc.cu -- nn.MaxPool2d on CHANNELS-LAST tensors (the stem of the ResNet hosts: relu(bn1(conv1(x))) -> MaxPool2d(3, 2, 1),
// models/imagenet/resnet_cnsn.py:183,238), forward and backward.
//
// Why it is here: once the ResNet-50 steps run channels_last with this package's SelfNorm / BatchNorm2d kernels, torch's
// max_pool2d is the slowest single launch of the step -- (768,64,112,112) bf16: forward 2.3 ms, backward 5.3 ms
// (profiles/r02_jsd_profile_channels_last.txt) for 1.5 GB / 1.8 GB of traffic, i.e. 0.7 / 0.3 TB/s; it keeps an int64
// index per output element.  Here: one 16-byte channel vector per thread, a uint8 window code per output element,
// the backward is a gather over the (at most ceil(k/stride)^2) windows that contain an input pixel -- no atomics, every
// dx element written once.
//
// Semantics = torch's: the window is scanned rows first, a later element replaces the running maximum only if it is
// strictly greater or NaN (so the FIRST of equal maxima gets the gradient -- after a ReLU whole windows are 0), padding
// never wins, dilation 1, ceil_mode off.
#include <stdio.h>

#include "flow_common.cuh"

namespace cnsn {
namespace pool {

constexpr int kT = 256;

struct Geom {
    int N, C, H, W, OH, OW;
    int k, stride, pad;
    int CG;                    // 16-byte vectors per pixel
    long long total;           // threads needed
};

template <int V> struct Code;                                  // V window codes, one byte each
template <> struct Code<4> { typedef unsigned int type; };
template <> struct Code<8> { typedef uint2 type; };

template <typename T>
__global__ void __launch_bounds__(kT) k_maxpool_nhwc_fwd(const T* __restrict__ x, T* __restrict__ y, unsigned char* __restrict__ code,
                                                         const Geom g) {
    constexpr int V = VecOf<T>::n;
    const long long t = (long long)blockIdx.x * kT + threadIdx.x;
    if (t >= g.total) return;
    const int cg = (int)(t % g.CG);
    long long p = t / g.CG;                                  // output pixel index: (n * OH + oh) * OW + ow
    const int ow = (int)(p % g.OW); p /= g.OW;
    const int oh = (int)(p % g.OH);
    const int n = (int)(p / g.OH);
    const int h0 = oh * g.stride - g.pad, w0 = ow * g.stride - g.pad;
    float best[V];
    unsigned char arg[V];
#pragma unroll
    for (int e = 0; e < V; ++e) { best[e] = -INFINITY; arg[e] = 0; }
    const uint4* vx = reinterpret_cast<const uint4*>(x) + (size_t)n * g.H * g.W * g.CG + cg;
    for (int kh = 0; kh < g.k; ++kh) {
        const int ih = h0 + kh;
        if (ih < 0 || ih >= g.H) continue;
        for (int kw = 0; kw < g.k; ++kw) {
            const int iw = w0 + kw;
            if (iw < 0 || iw >= g.W) continue;
            float a[V];
            unpack<T>(__ldg(vx + ((size_t)ih * g.W + iw) * g.CG), a);
            const unsigned char c = (unsigned char)(kh * g.k + kw);
#pragma unroll
            for (int e = 0; e < V; ++e)
                if (a[e] > best[e] || a[e] != a[e]) { best[e] = a[e]; arg[e] = c; }
        }
    }
    reinterpret_cast<uint4*>(y)[t] = pack<T>(best);
    typename Code<V>::type packed;
    unsigned char* pc = reinterpret_cast<unsigned char*>(&packed);
#pragma unroll
    for (int e = 0; e < V; ++e) pc[e] = arg[e];
    reinterpret_cast<typename Code<V>::type*>(code)[t] = packed;
}

template <typename T>
__global__ void __launch_bounds__(kT) k_maxpool_nhwc_bwd(const T* __restrict__ dy, const unsigned char* __restrict__ code, T* __restrict__ dx,
                                                         const Geom g) {
    constexpr int V = VecOf<T>::n;
    const long long t = (long long)blockIdx.x * kT + threadIdx.x;
    if (t >= g.total) return;
    const int cg = (int)(t % g.CG);
    long long p = t / g.CG;                                  // input pixel index: (n * H + ih) * W + iw
    const int iw = (int)(p % g.W); p /= g.W;
    const int ih = (int)(p % g.H);
    const int n = (int)(p / g.H);
    // output windows that contain (ih, iw): oh * stride - pad <= ih < oh * stride - pad + k
    const int oh_lo = max(0, (ih + g.pad - g.k + g.stride) / g.stride), oh_hi = min(g.OH - 1, (ih + g.pad) / g.stride);
    const int ow_lo = max(0, (iw + g.pad - g.k + g.stride) / g.stride), ow_hi = min(g.OW - 1, (iw + g.pad) / g.stride);
    float acc[V];
#pragma unroll
    for (int e = 0; e < V; ++e) acc[e] = 0.f;
    for (int oh = oh_lo; oh <= oh_hi; ++oh) {
        const int kh = ih + g.pad - oh * g.stride;
        if (kh < 0 || kh >= g.k) continue;
        for (int ow = ow_lo; ow <= ow_hi; ++ow) {
            const int kw = iw + g.pad - ow * g.stride;
            if (kw < 0 || kw >= g.k) continue;
            const size_t o = (((size_t)n * g.OH + oh) * g.OW + ow) * g.CG + cg;
            const typename Code<V>::type packed = __ldg(reinterpret_cast<const typename Code<V>::type*>(code) + o);
            const unsigned char* pc = reinterpret_cast<const unsigned char*>(&packed);
            float d[V];
            unpack<T>(__ldg(reinterpret_cast<const uint4*>(dy) + o), d);
            const unsigned char c = (unsigned char)(kh * g.k + kw);
#pragma unroll
            for (int e = 0; e < V; ++e) if (pc[e] == c) acc[e] += d[e];
        }
    }
    reinterpret_cast<uint4*>(dx)[t] = pack<T>(acc);
}

static int make_geom(Geom& g, int dtype, int N, int C, int H, int W, int k, int stride, int pad, bool backward) {
    const int esz = (int)esize(dtype);
    if (k < 1 || k > 15 || stride < 1 || pad < 0 || 2 * pad > k) return CNSN_E_BADARG;        // torch: pad <= k / 2
    if (((size_t)C * esz) % 16) return CNSN_E_UNSUPPORTED;
    g.N = N; g.C = C; g.H = H; g.W = W; g.k = k; g.stride = stride; g.pad = pad;
    g.OH = (H + 2 * pad - k) / stride + 1;
    g.OW = (W + 2 * pad - k) / stride + 1;
    if (g.OH < 1 || g.OW < 1) return CNSN_E_BADARG;
    g.CG = C * esz / 16;
    g.total = (long long)N * (backward ? (long long)H * W : (long long)g.OH * g.OW) * g.CG;
    if ((g.total + kT - 1) / kT > 0x7fffffffll) return CNSN_E_UNSUPPORTED;
    return 0;
}

}  // namespace pool
}  // namespace cnsn

using namespace cnsn;

extern "C" int cnsn_maxpool_nhwc_out(int H, int W, int k, int stride, int pad, int* OH, int* OW) {
    if (!OH || !OW || k < 1 || stride < 1 || pad < 0) return CNSN_E_BADARG;
    *OH = (H + 2 * pad - k) / stride + 1;
    *OW = (W + 2 * pad - k) / stride + 1;
    return (*OH < 1 || *OW < 1) ? CNSN_E_BADARG : CNSN_OK;
}

extern "C" int cnsn_maxpool_nhwc_fwd(const void* x, void* y, unsigned char* code, int dtype, int N, int C, int H, int W,
                                     int k, int stride, int pad, void* stream) {
    if (!x || !y || !code || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (!aligned16(x) || !aligned16(y) || (reinterpret_cast<uintptr_t>(code) & 7u)) return CNSN_E_ALIGN;
    pool::Geom g{};
    const int rc = pool::make_geom(g, dtype, N, C, H, W, k, stride, pad, false);
    if (rc) return rc;
    const unsigned grid = (unsigned)((g.total + pool::kT - 1) / pool::kT);
    CNSN_DISPATCH_DTYPE(dtype, T, (pool::k_maxpool_nhwc_fwd<T><<<grid, pool::kT, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, code, g)));
    return launch_status();
}

extern "C" int cnsn_maxpool_nhwc_bwd(const void* dy, const unsigned char* code, void* dx, int dtype, int N, int C, int H, int W,
                                     int k, int stride, int pad, void* stream) {
    if (!dy || !dx || !code || check_dims(N, C, H, W)) return CNSN_E_BADARG;
    if (dtype < CNSN_F32 || dtype > CNSN_F16) return CNSN_E_BADARG;
    if (!aligned16(dy) || !aligned16(dx) || (reinterpret_cast<uintptr_t>(code) & 7u)) return CNSN_E_ALIGN;
    pool::Geom g{};
    const int rc = pool::make_geom(g, dtype, N, C, H, W, k, stride, pad, true);
    if (rc) return rc;
    const unsigned grid = (unsigned)((g.total + pool::kT - 1) / pool::kT);
    CNSN_DISPATCH_DTYPE(dtype, T, (pool::k_maxpool_nhwc_bwd<T><<<grid, pool::kT, 0, (cudaStream_t)stream>>>((const T*)dy, code, (T*)dx, g)));
    return launch_status();
}
