// stream_f32.cu -- instantiations of the fp32 (FFMA) stream kernel: WearGait default widths (one translation unit per family: they build in parallel)
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

StreamKernelFn find_kernel_fog(const KernelKey& k);       // stream_f32_fog.cu
StreamKernelFn find_kernel_wide(const KernelKey& k);      // stream_f32_wide.cu

StreamKernelFn find_kernel(const KernelKey& k) {
    // SharedLatent3 (weargait_encoders.py:284-322): per-stream Linear(12 -> proj_ch=16) before the backbone
    GK_CASE_P(ENC_CONV_GELU_LN, 2, 3, 0, 12, 16, 4, 16)
    GK_CASE_P(ENC_INSOLE, 13, 5, 24, 12, 16, 4, 16)
    GK_CASE_P(ENC_CONV_GELU_LN, 24, 3, 0, 12, 16, 4, 16)
    // WearGait defaults (weargait_train.py:655-673): C=12, H=24, S=16, bdim=8
    GK_CASE(ENC_CONV_GELU_LN, 2, 3, 0, 12, 16, 4)
    GK_CASE(ENC_INSOLE, 13, 5, 24, 12, 16, 4)
    GK_CASE(ENC_CONV_GELU_LN, 24, 3, 0, 12, 16, 4)
    if (StreamKernelFn f = find_kernel_fog(k)) return f;
    return find_kernel_wide(k);
}

}  // namespace gaitk
