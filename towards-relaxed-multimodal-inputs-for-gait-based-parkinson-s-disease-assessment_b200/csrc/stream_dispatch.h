// stream_dispatch.h -- kernel lookup across translation units (one TU per kernel family so that they build in parallel)
#pragma once
#include "stream_common.cuh"

namespace gaitk {
struct TcPlan;
typedef void (*StreamKernelFn)(const StreamArgs, const SmemPlan);
typedef void (*StreamKernelTcFn)(const StreamArgs, const TcPlan);
struct KernelKey { int enc, CIN, KT1, H, C, S, NFL, PROJ; };
StreamKernelFn find_kernel(const KernelKey& k);                               // stream_f32.cu   (fp32 FFMA)
StreamKernelTcFn find_kernel_tc(const KernelKey& k, bool fixed_geometry);    // stream_tc.cu    (tf32 tcgen05 + mma.sync)
StreamKernelTcFn find_kernel_tc2(const KernelKey& k);                        // stream_tc.cu, -DGAITK_WITH_TC2 only
typedef void (*StreamKernelWsFn)(const StreamArgs);
struct WsKernel { StreamKernelWsFn fn; int threads, smem, groups; };
bool find_kernel_ws(const KernelKey& k, int num_classes, WsKernel* out);    // stream_ws.cu    (split-bf16, warp-specialised, all tcgen05)
}  // namespace gaitk
