// fusion_kernels.cuh -- the pieces of the fusion baselines that sit BETWEEN the stream-kernel stages
// (encoder stage -> fusion op -> trunk stage -> linear head), and the Adam update that trains them.
//
// Reference semantics (paths relative to the reference root):
//   CheapCrossAttention   data/WearGait/weargait_encoders.py:324-336, train/feature_encoder.py:497-528:
//                         out = softmax(A B^T / sqrt(d)) B per window, zero parameters
//   nn.Linear heads / per-stream projections   weargait_encoders.py:30-37,299-303; feature_encoder.py:380-386,470-474
//   torch.optim.Adam      train/baselines/fusion_train.py:202 (defaults: betas (0.9, 0.999), eps 1e-8, no weight decay)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gaitk {

// ------------------------------------------------------------------------------------------ zero-parameter cross attention
// One CTA per window, thread i = query row i (T <= 128 rows, d <= 16 channels).  A, Bm, out: (B, T, d) fp32.
// Forward keeps nothing: the backward recomputes the score rows (T x T x d MACs per window: tiny next to the HBM traffic).
constexpr int XA_TMAX = 128, XA_DMAX = 16;

template <int D>
__global__ void __launch_bounds__(XA_TMAX) xattn_fwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, float* __restrict__ out,
                                                            int T, float scale) {
    __shared__ float Bs[XA_TMAX][D];
    const int b = blockIdx.x, i = threadIdx.x;
    for (int e = threadIdx.x; e < T * D; e += blockDim.x) Bs[e / D][e % D] = Bm[(size_t)b * T * D + e];
    __syncthreads();
    if (i >= T) return;
    float a[D];
#pragma unroll
    for (int c = 0; c < D; ++c) a[c] = A[((size_t)b * T + i) * D + c] * scale;
    float mx = -INFINITY;
    for (int j = 0; j < T; ++j) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) s = fmaf(a[c], Bs[j][c], s);
        mx = fmaxf(mx, s);
    }
    float l = 0.f, o[D];
#pragma unroll
    for (int c = 0; c < D; ++c) o[c] = 0.f;
    for (int j = 0; j < T; ++j) {
        float s = 0.f;
#pragma unroll
        for (int c = 0; c < D; ++c) s = fmaf(a[c], Bs[j][c], s);
        const float p = expf(s - mx);
        l += p;
#pragma unroll
        for (int c = 0; c < D; ++c) o[c] = fmaf(p, Bs[j][c], o[c]);
    }
    const float inv = 1.0f / l;
#pragma unroll
    for (int c = 0; c < D; ++c) out[((size_t)b * T + i) * D + c] = o[c] * inv;
}

// backward: P = softmax(S), S = scale A B^T;  dP = dO B^T;  dS = P o (dP - rowsum(dP o P));
//   dA = scale dS B;   dB = P^T dO + scale dS^T A.
// Pass 1 (thread = query i): row max m_i, row sum l_i, D_i = rowsum(dP o P), and dA_i.
// Pass 2 (thread = key j): dB_j = sum_i p_ij dO_i + scale ds_ij A_i with p_ij recomputed from (m_i, l_i): no atomics, fixed order.
template <int D>
__global__ void __launch_bounds__(XA_TMAX) xattn_bwd_kernel(const float* __restrict__ A, const float* __restrict__ Bm, const float* __restrict__ dO,
                                                            float* __restrict__ dA, float* __restrict__ dB, int T, float scale) {
    __shared__ float As[XA_TMAX][D], Bs[XA_TMAX][D], Gs[XA_TMAX][D];
    __shared__ float ms[XA_TMAX], ls[XA_TMAX], Ds[XA_TMAX];
    const int b = blockIdx.x, i = threadIdx.x;
    for (int e = threadIdx.x; e < T * D; e += blockDim.x) {
        const size_t g = (size_t)b * T * D + e;
        As[e / D][e % D] = A[g]; Bs[e / D][e % D] = Bm[g]; Gs[e / D][e % D] = dO[g];
    }
    __syncthreads();
    if (i < T) {
        float a[D], g[D];
#pragma unroll
        for (int c = 0; c < D; ++c) { a[c] = As[i][c] * scale; g[c] = Gs[i][c]; }
        float mx = -INFINITY;
        for (int j = 0; j < T; ++j) {
            float s = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) s = fmaf(a[c], Bs[j][c], s);
            mx = fmaxf(mx, s);
        }
        float l = 0.f, dsum = 0.f;
        for (int j = 0; j < T; ++j) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) { s = fmaf(a[c], Bs[j][c], s); dp = fmaf(g[c], Bs[j][c], dp); }
            const float p = expf(s - mx);
            l += p; dsum = fmaf(p, dp, dsum);
        }
        const float inv = 1.0f / l;
        const float Di = dsum * inv;
        ms[i] = mx; ls[i] = inv; Ds[i] = Di;
        float da[D];
#pragma unroll
        for (int c = 0; c < D; ++c) da[c] = 0.f;
        for (int j = 0; j < T; ++j) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) { s = fmaf(a[c], Bs[j][c], s); dp = fmaf(g[c], Bs[j][c], dp); }
            const float ds = expf(s - mx) * inv * (dp - Di);
#pragma unroll
            for (int c = 0; c < D; ++c) da[c] = fmaf(ds, Bs[j][c], da[c]);
        }
#pragma unroll
        for (int c = 0; c < D; ++c) dA[((size_t)b * T + i) * D + c] = da[c] * scale;
    }
    __syncthreads();
    if (i < T) {
        const int j = i;
        float bj[D], db[D];
#pragma unroll
        for (int c = 0; c < D; ++c) { bj[c] = Bs[j][c]; db[c] = 0.f; }
        for (int q = 0; q < T; ++q) {
            float s = 0.f, dp = 0.f;
#pragma unroll
            for (int c = 0; c < D; ++c) { s = fmaf(As[q][c] * scale, bj[c], s); dp = fmaf(Gs[q][c], bj[c], dp); }
            const float p = expf(s - ms[q]) * ls[q];
            const float ds = p * (dp - Ds[q]) * scale;
#pragma unroll
            for (int c = 0; c < D; ++c) db[c] = fmaf(p, Gs[q][c], fmaf(ds, As[q][c], db[c]));
        }
#pragma unroll
        for (int c = 0; c < D; ++c) dB[((size_t)b * T + j) * D + c] = db[c];
    }
}

// ------------------------------------------------------------------------------------------ small dense layer y = x W^T + b
// x (R, I), W (O, I), b (O) or NULL, y (R, O); I <= 256, O <= 32 (heads: I = 128 / 256, O = K; projections: I = 6, O = 16).
// One warp per row: lanes stride the input features, the O dot products are reduced by shuffles.
constexpr int LIN_IMAX = 256, LIN_OMAX = 32;
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ bias,
                                                         float* __restrict__ y, int R, int I, int O) {
    extern __shared__ float lin_sh[];                          // (O, I)
    float* Ws = lin_sh;
    for (int e = threadIdx.x; e < O * I; e += blockDim.x) Ws[e] = W[e];
    __syncthreads();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    for (int r = blockIdx.x * wpb + (threadIdx.x >> 5); r < R; r += gridDim.x * wpb) {
        float xv[LIN_IMAX / 32];
#pragma unroll
        for (int q = 0; q < LIN_IMAX / 32; ++q) { const int j = lane + 32 * q; xv[q] = j < I ? x[(size_t)r * I + j] : 0.f; }
        for (int o = 0; o < O; ++o) {
            float s = 0.f;
#pragma unroll
            for (int q = 0; q < LIN_IMAX / 32; ++q) { const int j = lane + 32 * q; if (j < I) s = fmaf(xv[q], Ws[o * I + j], s); }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
            if (lane == 0) y[(size_t)r * O + o] = s + (bias ? bias[o] : 0.f);
        }
    }
}
// dx = dy W (R, I); per-CTA partials of dW (O, I) and db (O) over the CTA's rows -> part[cta][O * I + O]; linear_reduce_kernel sums
// the CTAs in order.  Thread t owns input feature t (I <= 256 = blockDim): dW[:, t] accumulates in registers, rows are visited in order.
__global__ void __launch_bounds__(256) linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ W, const float* __restrict__ dy,
                                                         float* __restrict__ dx, float* __restrict__ part, int R, int I, int O, int rows_per_cta) {
    extern __shared__ float lin_sh[];                          // W (O, I) | dy rows of the current chunk (32, O)
    float* Ws = lin_sh; float* dys = lin_sh + O * I;
    for (int e = threadIdx.x; e < O * I; e += blockDim.x) Ws[e] = W[e];
    const int t = threadIdx.x;
    float gw[LIN_OMAX];
#pragma unroll
    for (int o = 0; o < LIN_OMAX; ++o) gw[o] = 0.f;
    float gb = 0.f;                                            // thread o < O also owns db[o]
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(R, r0 + rows_per_cta);
    for (int rc = r0; rc < r1; rc += 32) {
        const int n = min(32, r1 - rc);
        __syncthreads();
        for (int e = threadIdx.x; e < n * O; e += blockDim.x) dys[e] = dy[(size_t)rc * O + e];
        __syncthreads();
        for (int q = 0; q < n; ++q) {
            const float* d = dys + q * O;
            if (t < I) {
                const float xv = x[(size_t)(rc + q) * I + t];
                float acc = 0.f;
#pragma unroll
                for (int o = 0; o < LIN_OMAX; ++o) if (o < O) { acc = fmaf(d[o], Ws[o * I + t], acc); gw[o] = fmaf(d[o], xv, gw[o]); }
                if (dx) dx[(size_t)(rc + q) * I + t] = acc;
            }
            if (t < O) gb += d[t];
        }
    }
    float* p = part + (size_t)blockIdx.x * (O * I + O);
    if (t < I) {
#pragma unroll
        for (int o = 0; o < LIN_OMAX; ++o) if (o < O) p[o * I + t] = gw[o];
    }
    if (t < O) p[O * I + t] = gb;
}
__global__ void linear_reduce_kernel(const float* __restrict__ part, int nparts, int n, float* __restrict__ dW, float* __restrict__ db, int OI) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float s = 0.f;
    for (int c = 0; c < nparts; ++c) s += part[(size_t)c * n + e];
    if (e < OI) dW[e] = s; else if (db) db[e - OI] = s;
}

// ------------------------------------------------------------------------------------------ Adam over a table of tensors
// torch.optim.Adam.step (single-tensor formulas): m = b1 m + (1 - b1) g; v = b2 v + (1 - b2) g^2;
// p -= (lr / (1 - b1^t)) * m / (sqrt(v) / sqrt(1 - b2^t) + eps); optional L2 weight decay g += wd p.
struct AdamSeg { float* p; const float* g; float* m; float* v; int n; };
constexpr int ADAM_MAX_SEG = 48;
struct AdamArgs { AdamSeg seg[ADAM_MAX_SEG]; int nseg; float lr, b1, b2, eps, wd, bc1, bc2_sqrt; };
__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs A) {
    const AdamSeg s = A.seg[blockIdx.y];
    const float step = A.lr / A.bc1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < s.n; i += gridDim.x * blockDim.x) {
        float g = s.g[i];
        const float p = s.p[i];
        if (A.wd != 0.f) g = fmaf(A.wd, p, g);
        const float m0 = s.m[i];
        const float m = m0 + (g - m0) * (1.f - A.b1);           // torch: exp_avg.lerp_(grad, 1 - beta1)
        const float v = A.b2 * s.v[i] + (1.f - A.b2) * g * g;   // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, value = 1 - beta2)
        s.m[i] = m; s.v[i] = v;
        const float denom = sqrtf(v) / A.bc2_sqrt + A.eps;
        s.p[i] = p - step * (m / denom);
    }
}

}  // namespace gaitk
