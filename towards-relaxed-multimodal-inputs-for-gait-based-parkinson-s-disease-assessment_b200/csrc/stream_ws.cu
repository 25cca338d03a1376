// stream_ws.cu -- instantiations + launch geometry of the warp-specialised split-bf16 stream kernel (stream_kernel_ws.cuh)
#include "stream_kernel_ws.cuh"
#include "stream_dispatch.h"

namespace gaitk {

template <class Cfg, int G, int K> static void fill(WsKernel* o) {
    o->fn = &stream_kernel_ws<Cfg, G, K>; o->threads = WsLayout<Cfg, G>::NTH; o->smem = WsLayout<Cfg, G>::TOTAL; o->groups = G;
}
template <class Cfg, int G> static bool pick(int K, WsKernel* o) {
    if (K == 2) { fill<Cfg, G, 2>(o); return true; }
    if (K == 3) { fill<Cfg, G, 3>(o); return true; }
    if (K == 4) { fill<Cfg, G, 4>(o); return true; }
    return false;
}
// WearGait defaults (weargait_train.py:655-673): C = 12, H = 24, S = 16, bdim = 8, T = 64; plain linear head
bool find_kernel_ws(const KernelKey& k, int K, WsKernel* o) {
    if (k.PROJ != 0 || k.C != 12 || k.S != 16 || k.NFL != 4) return false;
    if (k.enc == ENC_CONV_GELU_LN && k.CIN == 2 && k.KT1 == 3) return pick<StreamCfg<ENC_CONV_GELU_LN, 2, 3, 0, 12, 16, 4>, 4>(K, o);
    if (k.enc == ENC_INSOLE && k.CIN == 13 && k.KT1 == 5 && k.H == 24) return pick<StreamCfg<ENC_INSOLE, 13, 5, 24, 12, 16, 4>, 3>(K, o);
    if (k.enc == ENC_CONV_GELU_LN && k.CIN == 24 && k.KT1 == 3) return pick<StreamCfg<ENC_CONV_GELU_LN, 24, 3, 0, 12, 16, 4>, 4>(K, o);
    return false;
}

}  // namespace gaitk
