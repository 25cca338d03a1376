// stream_tc_plan.h -- shared-memory plan of the tf32 tensor-core stream kernel (offsets in floats; filled on the host)
#pragma once
namespace gaitk {
struct TcPlan {
    int X, HA, D1, XH, D, F, RSTD, Z;
    int W1B, B1, W2B, B2, W2D, LNG, LNB, WBB, BB, WBD, HW, HB, HNG, HNB, INW;
    int DP, BINS, STAGE, STG, P;
    int total;
};
}  // namespace gaitk
