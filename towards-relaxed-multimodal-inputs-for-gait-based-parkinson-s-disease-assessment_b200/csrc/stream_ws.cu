// stream_ws.cu -- instantiations + launch geometry of the warp-specialised split-bf16 stream kernel (stream_kernel_ws.cuh)
#include <stdlib.h>

#include "stream_kernel_ws.cuh"
#include "stream_dispatch.h"

namespace gaitk {

template <class Cfg, int G, int K, int SL> static void fill(WsKernel* o) {
    o->fn = &stream_kernel_ws<Cfg, G, K, SL>; o->threads = WsLayout<Cfg, G, SL>::NTH; o->smem = WsLayout<Cfg, G, SL>::TOTAL; o->groups = G;
}
template <class Cfg, int G, int SL> static bool pick(int K, WsKernel* o) {
    if (K == 2) { fill<Cfg, G, 2, SL>(o); return true; }
    if (K == 3) { fill<Cfg, G, 3, SL>(o); return true; }
    if (K == 4) { fill<Cfg, G, 4, SL>(o); return true; }
    return false;
}
// G tiles in flight per CTA, as RG row warpgroups x SL slots (a row thread ping-pongs between the SL tiles of its warpgroup).
// Measured (profiles/r2_ws_history.md): the ping-pong variant (GAITK_WS_SLOTS=2) is correct but 4 % SLOWER than one tile per
// warpgroup (1.104 vs 1.060 ms per step): the step is not bound by the row warps' waits on the MMAs; default = 1.
// WearGait defaults (weargait_train.py:655-673): C = 12, H = 24, S = 16, bdim = 8, T = 64; plain linear head
bool find_kernel_ws(const KernelKey& k, int K, WsKernel* o) {
    if (k.PROJ != 0 || k.C != 12 || k.S != 16 || k.NFL != 4) return false;
    const char* e = getenv("GAITK_WS_SLOTS");
    const int slots = e ? atoi(e) : 1;
    if (k.enc == ENC_CONV_GELU_LN && k.CIN == 2 && k.KT1 == 3)
        return slots == 1 ? pick<StreamCfg<ENC_CONV_GELU_LN, 2, 3, 0, 12, 16, 4>, 4, 1>(K, o) : pick<StreamCfg<ENC_CONV_GELU_LN, 2, 3, 0, 12, 16, 4>, 4, 2>(K, o);
    if (k.enc == ENC_INSOLE && k.CIN == 13 && k.KT1 == 5 && k.H == 24)
        return slots == 1 ? pick<StreamCfg<ENC_INSOLE, 13, 5, 24, 12, 16, 4>, 3, 1>(K, o) : pick<StreamCfg<ENC_INSOLE, 13, 5, 24, 12, 16, 4>, 3, 3>(K, o);
    if (k.enc == ENC_CONV_GELU_LN && k.CIN == 24 && k.KT1 == 3)
        return slots == 1 ? pick<StreamCfg<ENC_CONV_GELU_LN, 24, 3, 0, 12, 16, 4>, 4, 1>(K, o) : pick<StreamCfg<ENC_CONV_GELU_LN, 24, 3, 0, 12, 16, 4>, 4, 2>(K, o);
    return false;
}

}  // namespace gaitk
