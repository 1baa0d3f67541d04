// ibn_general.cuh -- argument block and entry points of the general (three-kernel) IBN path, ibn_general.cu.
#pragma once

#include "common.cuh"

namespace cnsn {
namespace ibn_general {

struct GArgs {
    const void* x; const void* dy; void* out;
    int N, C, M, half, training;
    float momentum, eps_in, eps_bn;
    const float* in_w; const float* in_b; const float* bn_w; const float* bn_b;
    float* run_mean; float* run_var; long long* nbt;
    float* in_mean; float* in_rstd;     // [N][half]
    float* bn_mean; float* bn_rstd;     // [C - half]
    float* d_in_w; float* d_in_b; float* d_bn_w; float* d_bn_b;
    float* w0; float* w1;               // [N][C] each: (mean, M2) -> (scale, shift) / (A, B) -> (cb, cc)
};

// ws: 2*N*C floats.  Both return 0 or a cuda error.
int ibn_general_fwd(GArgs& a, int dtype, float* ws, cudaStream_t s);
int ibn_general_bwd(GArgs& a, int dtype, float* ws, cudaStream_t s);

}  // namespace ibn_general
}  // namespace cnsn
