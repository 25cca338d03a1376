// stream_f32_fog.cu -- instantiations of the fp32 (FFMA) stream kernel: FoG / FBG encoders and the trunk stages of the fusion baselines
#include "stream_kernel.cuh"
#include "stream_dispatch.h"

namespace gaitk {
template <class Cfg> static StreamKernelFn kfn() { return &stream_kernel<Cfg>; }
#define GK_CASE(e_, ci_, kt_, h_, c_, s_, nfl_) \
    if (k.enc == e_ && k.CIN == ci_ && k.KT1 == kt_ && k.H == h_ && k.C == c_ && k.S == s_ && k.NFL == nfl_ && k.PROJ == 0) \
        return kfn<StreamCfg<e_, ci_, kt_, h_, c_, s_, nfl_>>();
#define GK_CASE_P(e_, ci_, kt_, h_, c_, s_, nfl_, p_) \
    if (k.enc == e_ && k.CIN == ci_ && k.KT1 == kt_ && k.H == h_ && k.C == c_ && k.S == s_ && k.NFL == nfl_ && k.PROJ == p_) \
        return kfn<StreamCfg<e_, ci_, kt_, h_, c_, s_, nfl_, p_>>();

StreamKernelFn find_kernel_fog(const KernelKey& k) {
    // FoG (configs.py:17-31) and FBG (:2-16)
    GK_CASE(ENC_LINEAR_LN_RELU, 21, 1, 0, 6, 16, 4)
    GK_CASE(ENC_CONV_POOL, 6, 3, 0, 6, 16, 4)
    GK_CASE(ENC_LINEAR_LN_RELU, 51, 1, 0, 3, 16, 4)
    GK_CASE(ENC_CONV_POOL, 3, 3, 0, 3, 16, 4)
    // the same sensor encoders in pooled-taps form (3 x raw channels, one tap; stream_common.cuh)
    GK_CASE(ENC_POOL_LINEAR, 18, 1, 0, 6, 16, 4)
    GK_CASE(ENC_POOL_LINEAR, 9, 1, 0, 3, 16, 4)
    // trunk stages of the fusion baselines (no encoder; C = CIN): EarlyFusion3 3 x 12, CheapXAttn3 12; 2-stream twins
    // (feature_encoder.py:346-596) early 6 + 6, late / cross-attention 6, shared latent 16
    GK_CASE(ENC_NONE, 36, 1, 0, 36, 16, 4)
    GK_CASE(ENC_NONE, 12, 1, 0, 12, 16, 4)
    GK_CASE(ENC_NONE, 6, 1, 0, 6, 16, 4)
    GK_CASE(ENC_NONE, 16, 1, 0, 16, 16, 4)
    return nullptr;
}

}  // namespace gaitk
