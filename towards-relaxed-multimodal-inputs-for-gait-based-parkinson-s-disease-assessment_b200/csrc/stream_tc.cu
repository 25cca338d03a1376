// stream_tc.cu -- instantiations of the tf32 tensor-core stream kernel (tcgen05 + mma.sync)
#include "stream_kernel_tc.cuh"
#ifdef GAITK_WITH_TC2
#include "stream_kernel_tc2.cuh"
#endif
#include "stream_dispatch.h"

namespace gaitk {
template <class Cfg, bool FX> static StreamKernelTcFn kfn_tc() { return &stream_kernel_tc<Cfg, FX>; }
StreamKernelTcFn find_kernel_tc(const KernelKey& k, bool fixed_geometry) {
#define GK_CASE(e_, ci_, kt_, h_, c_, s_, nfl_) \
    if (k.enc == e_ && k.CIN == ci_ && k.KT1 == kt_ && k.H == h_ && k.C == c_ && k.S == s_ && k.NFL == nfl_ && k.PROJ == 0) \
        return fixed_geometry ? kfn_tc<StreamCfg<e_, ci_, kt_, h_, c_, s_, nfl_>, true>() : kfn_tc<StreamCfg<e_, ci_, kt_, h_, c_, s_, nfl_>, false>();
    GK_CASE(ENC_CONV_GELU_LN, 2, 3, 0, 12, 16, 4)
    GK_CASE(ENC_INSOLE, 13, 5, 24, 12, 16, 4)
    GK_CASE(ENC_CONV_GELU_LN, 24, 3, 0, 12, 16, 4)
#undef GK_CASE
    return nullptr;
}

#ifdef GAITK_WITH_TC2
template <class Cfg> static StreamKernelTcFn kfn_tc2() { return &stream_kernel_tc2<Cfg>; }
// two-threads-per-row variant: compile-time geometry T = 64, W = 2, bdim = 8
StreamKernelTcFn find_kernel_tc2(const KernelKey& k) {
#define GK_CASE(e_, ci_, kt_, h_, c_, s_, nfl_) \
    if (k.enc == e_ && k.CIN == ci_ && k.KT1 == kt_ && k.H == h_ && k.C == c_ && k.S == s_ && k.NFL == nfl_ && k.PROJ == 0) \
        return kfn_tc2<StreamCfg<e_, ci_, kt_, h_, c_, s_, nfl_>>();
    GK_CASE(ENC_CONV_GELU_LN, 2, 3, 0, 12, 16, 4)
    GK_CASE(ENC_INSOLE, 13, 5, 24, 12, 16, 4)
    GK_CASE(ENC_CONV_GELU_LN, 24, 3, 0, 12, 16, 4)
#undef GK_CASE
    return nullptr;
}
#endif

}  // namespace gaitk
