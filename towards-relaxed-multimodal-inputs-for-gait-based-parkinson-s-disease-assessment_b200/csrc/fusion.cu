// fusion.cu -- C-ABI entries of the fusion-baseline building blocks that need no plan: zero-parameter cross attention,
// the small dense layer (heads, per-stream projections) and the Adam update (fusion_kernels.cuh).
#include <math.h>
#include <string.h>

#include "../../include/gaitk.h"
#include "fusion_kernels.cuh"

extern "C" int gaitk_set_error(int code, const char* fmt, ...);           // gaitk_api.cu (thread-local message)

using namespace gaitk;
#define FUSION_LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return gaitk_set_error((int)e_, "kernel launch: %s", cudaGetErrorString(e_)); } while (0)

template <int D> static void xa_fwd(const float* A, const float* B, float* out, int n, int T, float sc, cudaStream_t st) {
    xattn_fwd_kernel<D><<<n, (T + 31) / 32 * 32, 0, st>>>(A, B, out, T, sc);
}
template <int D> static void xa_bwd(const float* A, const float* B, const float* dO, float* dA, float* dB, int n, int T, float sc, cudaStream_t st) {
    xattn_bwd_kernel<D><<<n, (T + 31) / 32 * 32, 0, st>>>(A, B, dO, dA, dB, T, sc);
}

extern "C" int gaitk_xattn_forward(const float* A, const float* B, float* out, int n_windows, int T, int d, void* stream) {
    if (!A || !B || !out || n_windows < 0 || T < 1 || T > XA_TMAX) return gaitk_set_error(GAITK_E_BADARG, "xattn: bad argument (T <= %d)", XA_TMAX);
    if (n_windows == 0) return 0;
    const float sc = 1.0f / sqrtf((float)d);                   // dim ** -0.5 (weargait_encoders.py:330)
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
        case 3: xa_fwd<3>(A, B, out, n_windows, T, sc, st); break;
        case 6: xa_fwd<6>(A, B, out, n_windows, T, sc, st); break;
        case 8: xa_fwd<8>(A, B, out, n_windows, T, sc, st); break;
        case 12: xa_fwd<12>(A, B, out, n_windows, T, sc, st); break;
        case 16: xa_fwd<16>(A, B, out, n_windows, T, sc, st); break;
        default: return gaitk_set_error(GAITK_E_SHAPE, "xattn: channel width %d not instantiated (3, 6, 8, 12, 16)", d);
    }
    FUSION_LAUNCH_CHECK();
    return 0;
}
extern "C" int gaitk_xattn_backward(const float* A, const float* B, const float* dout, float* dA, float* dB, int n_windows, int T, int d, void* stream) {
    if (!A || !B || !dout || !dA || !dB || n_windows < 0 || T < 1 || T > XA_TMAX) return gaitk_set_error(GAITK_E_BADARG, "xattn: bad argument");
    if (n_windows == 0) return 0;
    const float sc = 1.0f / sqrtf((float)d);
    cudaStream_t st = (cudaStream_t)stream;
    switch (d) {
        case 3: xa_bwd<3>(A, B, dout, dA, dB, n_windows, T, sc, st); break;
        case 6: xa_bwd<6>(A, B, dout, dA, dB, n_windows, T, sc, st); break;
        case 8: xa_bwd<8>(A, B, dout, dA, dB, n_windows, T, sc, st); break;
        case 12: xa_bwd<12>(A, B, dout, dA, dB, n_windows, T, sc, st); break;
        case 16: xa_bwd<16>(A, B, dout, dA, dB, n_windows, T, sc, st); break;
        default: return gaitk_set_error(GAITK_E_SHAPE, "xattn: channel width %d not instantiated (3, 6, 8, 12, 16)", d);
    }
    FUSION_LAUNCH_CHECK();
    return 0;
}

static const int LIN_ROWS_PER_CTA = 256;
extern "C" size_t gaitk_linear_workspace_bytes(int R, int I, int O) {
    const int nparts = (R + LIN_ROWS_PER_CTA - 1) / LIN_ROWS_PER_CTA;
    return (size_t)(nparts > 0 ? nparts : 1) * (size_t)(O * I + O) * sizeof(float);
}
extern "C" int gaitk_linear_forward(const float* x, const float* W, const float* bias, float* y, int R, int I, int O, void* stream) {
    if (!x || !W || !y || R < 0 || I < 1 || I > LIN_IMAX || O < 1 || O > LIN_OMAX)
        return gaitk_set_error(GAITK_E_BADARG, "linear: bad argument (in <= %d, out <= %d)", LIN_IMAX, LIN_OMAX);
    if (R == 0) return 0;
    const int grid = (R + 7) / 8 < 1184 ? (R + 7) / 8 : 1184;
    linear_fwd_kernel<<<grid, 256, (size_t)O * I * sizeof(float), (cudaStream_t)stream>>>(x, W, bias, y, R, I, O);
    FUSION_LAUNCH_CHECK();
    return 0;
}
extern "C" int gaitk_linear_backward(const float* x, const float* W, const float* dy, float* dx, float* dW, float* db, int R, int I, int O,
                                     void* workspace, size_t workspace_bytes, void* stream) {
    if (!x || !W || !dy || !dW || !workspace || R < 1 || I < 1 || I > LIN_IMAX || O < 1 || O > LIN_OMAX)
        return gaitk_set_error(GAITK_E_BADARG, "linear backward: bad argument");
    if (workspace_bytes < gaitk_linear_workspace_bytes(R, I, O)) return gaitk_set_error(GAITK_E_BADARG, "linear backward: workspace too small");
    const int nparts = (R + LIN_ROWS_PER_CTA - 1) / LIN_ROWS_PER_CTA, n = O * I + O;
    cudaStream_t st = (cudaStream_t)stream;
    linear_bwd_kernel<<<nparts, 256, (size_t)(O * I + 32 * O) * sizeof(float), st>>>(x, W, dy, dx, (float*)workspace, R, I, O, LIN_ROWS_PER_CTA);
    FUSION_LAUNCH_CHECK();
    linear_reduce_kernel<<<(n + 255) / 256, 256, 0, st>>>((const float*)workspace, nparts, n, dW, db, O * I);
    FUSION_LAUNCH_CHECK();
    return 0;
}

extern "C" int gaitk_adam(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq, const int64_t* numel,
                          int n_tensors, float lr, float beta1, float beta2, float eps, float weight_decay, int step, void* stream) {
    if (!params || !grads || !exp_avg || !exp_avg_sq || !numel || n_tensors < 1 || step < 1) return gaitk_set_error(GAITK_E_BADARG, "adam: bad argument");
    const float bc1 = 1.0f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, (float)step));
    for (int base = 0; base < n_tensors; base += ADAM_MAX_SEG) {
        AdamArgs A; memset(&A, 0, sizeof(A));
        A.nseg = n_tensors - base < ADAM_MAX_SEG ? n_tensors - base : ADAM_MAX_SEG;
        int64_t mx = 1;
        for (int i = 0; i < A.nseg; ++i) {
            A.seg[i].p = params[base + i]; A.seg[i].g = grads[base + i]; A.seg[i].m = exp_avg[base + i]; A.seg[i].v = exp_avg_sq[base + i];
            A.seg[i].n = (int)numel[base + i]; if (numel[base + i] > mx) mx = numel[base + i];
        }
        A.lr = lr; A.b1 = beta1; A.b2 = beta2; A.eps = eps; A.wd = weight_decay; A.bc1 = bc1; A.bc2_sqrt = bc2_sqrt;
        const int gx = (int)((mx + 255) / 256 < 64 ? (mx + 255) / 256 : 64);
        adam_kernel<<<dim3(gx, A.nseg), 256, 0, (cudaStream_t)stream>>>(A);
        FUSION_LAUNCH_CHECK();
    }
    return 0;
}
