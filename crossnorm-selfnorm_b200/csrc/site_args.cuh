// site_args.cuh -- argument block of the fused CrossNorm -> SelfNorm site kernels (site_flow.cu, site_tmem.cu).
#pragma once

#include "selfnorm_fold.cuh"

namespace cnsn {
namespace flow {

struct SiteArgs {
    FArgs sn;               // SelfNorm half: x / dy / out, N, C, M, nI, parameters, save block, sn words (pub), chan, ticket
    int H, W;
    Window cw, sw;          // content / style window
    float lam, cn_eps;
    const int* perm;        // [N] style source of every sample
    float* mu_c; float* sd_c; float* mu_s; float* sd_s;     // CrossNorm save block
    float2* pub_cn;         // CrossNorm words, pre-filled with the sentinel: forward [C][N], backward [C][N][3]
};

// site_tmem.cu: the shared + tensor memory pipeline for whole-plane windows; -100 when it does not apply
int site_tmem_fwd(SiteArgs& s, int dtype, float* scratch, cudaStream_t stream);
int site_tmem_bwd(SiteArgs& s, int dtype, float* scratch, cudaStream_t stream);

}  // namespace flow
}  // namespace cnsn
