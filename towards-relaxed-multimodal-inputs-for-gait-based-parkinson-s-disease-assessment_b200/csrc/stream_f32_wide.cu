// stream_f32_wide.cu -- instantiations of the fp32 (FFMA) stream kernel: the scaled sweep's widths
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

StreamKernelFn find_kernel_wide(const KernelKey& k) {
    // scaled sweep (BASELINE configs[4]: --enc_out_ch 24 --shared_out_ch 32, H = 48, NF = 256; T = 256 runs as 2-CTA clusters)
    GK_CASE(ENC_CONV_GELU_LN, 2, 3, 0, 24, 32, 8)
    GK_CASE(ENC_INSOLE, 13, 5, 48, 24, 32, 8)
    GK_CASE(ENC_CONV_GELU_LN, 24, 3, 0, 24, 32, 8)
    return nullptr;
}

}  // namespace gaitk
