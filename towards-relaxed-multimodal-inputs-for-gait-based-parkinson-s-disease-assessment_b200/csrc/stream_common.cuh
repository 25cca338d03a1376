// stream_common.cuh -- pieces shared by the fp32 (FFMA) and the tensor-core stream kernels: argument
// structs, row helpers, GELU / LayerNorm, and the warp-per-window pooling + head + loss phase.
//
// Reference semantics (paths relative to the reference root):
//   encoders   data/WearGait/weargait_encoders.py:40-101, train/feature_encoder.py:27-77
//   backbone   weargait_encoders.py:103-113 / feature_encoder.py:80-109  (+ .flatten(1))
//   head       weargait_encoders.py:19-37 / feature_encoder.py:7-24,112-146
//   losses     train/learning/optimizers/classification_losses.py:54-109
//   backward   what autograd does for losses[i].backward() multitask_weighting.py:680-688
//
// Work decomposition: a tile = W windows interleaved row-wise (row r <-> time t = r / W, window
// w = r % W), so a time shift of one step is a shift of W rows and the zero "same" padding of the
// convolutions is simply the zero halo at both ends of the tile -- valid for every window at once.
// Thread i owns rows i, i+128, ...  Every activation lives in shared memory as [chunk][row][4]
// (chunk = channel / 4): a row's 4-channel group is one 16-byte word, consecutive rows are
// consecutive words (conflict-free LDS.128 / STS.128, and exactly the no-swizzle K-major /
// MN-major core-matrix order tcgen05 descriptors address with SBO = 128 B).
// Inputs are read from HBM exactly once per step; nothing but logits and per-CTA partial gradient
// sums is written back.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gaitk {

constexpr int NT = 128;        // threads per CTA == rows per slot
constexpr int KMAX = 4;        // GAITK_MAX_CLASSES
constexpr int WMAX = 4;        // windows per tile (one warp each in the head phase)
// rows per chunk plane (RB) modulo 8.  Row-wise float4 accesses are conflict-free for any RB (a quarter warp reads
// 128 contiguous bytes); the mma.sync fragment loads of the weight gradients read lanes (g, t) -> channel g, row t:
// chunk g >> 2, word 4 t + (g & 3), so the two chunks must sit 16 banks apart: 4 RB = 16 (mod 32)  <=>  RB = 4 (mod 8).
#ifndef GAITK_FOG_MINB
#define GAITK_FOG_MINB 3
#endif
#ifndef GAITK_RB_MOD
#define GAITK_RB_MOD 4
#endif

enum EncKind { ENC_CONV_GELU_LN = 0, ENC_INSOLE = 1, ENC_LINEAR_LN_RELU = 2, ENC_CONV_POOL = 3,
               ENC_NONE = 4,         // no encoder: x IS the backbone input (the trunk stage of the fusion baselines)
               // SensorEncoder when it pools (feature_encoder.py:27-58: Conv1d k3 -> AdaptiveAvgPool1d, NOTHING non-linear in
               // between): mean_bin(conv(x)) == Linear over the three tap-shifted bin means of x.  The kernel pools the raw clip
               // while it loads it (T_in -> T rows of 3 CIN_raw channels, channel ci * 3 + tap) and runs a 1-tap "conv" whose
               // weight matrix IS conv1d.weight (C, CIN_raw, 3) read as (C, 3 CIN_raw): T_in / T times fewer MACs in the
               // forward and in the weight gradient, no (T_in, C) intermediate.  Cfg::CIN = 3 CIN_raw, Cfg::KT1 = 1.
               ENC_POOL_LINEAR = 5 };
enum Mode { MODE_FWD = 0, MODE_FUSED = 1, MODE_BWD_EXT = 2 };

// stream-local gradient layout (offsets in floats into a partial-gradient row; -1 = absent)
struct GradOff {
    int w1, b1, w2, b2, wsk, bsk, lng, lnb, wbb, bbb, hng, hnb, hw, hb, wp, bp;
    int total;      // NG
};

struct StreamArgs {
    // input
    const float* x;                // (B, T_in, CIN) dense, or frame store (N, CIN) with win_start
    const long long* win_start;    // optional [B] first frame of each window
    int B, T_in, T, W, bdim, K, NF;
    int rows_in, rows, halo, RBi, RB;
    int mode, zero_input, pool_sensor;
    int cl;                        // > 1: a window is split by TIME over a thread-block cluster of `cl` CTAs (T = cl * rows, W = 1):
                                   // conv halos travel through distributed shared memory, pooled features are all-gathered
    // parameters (global memory, PyTorch layouts)
    const float *w1, *b1, *w2, *b2, *wsk, *bsk, *lng, *lnb, *wbb, *bbb, *hng, *hnb, *hw, *hb;
    const float *wp, *bp;          // optional per-stream projection Linear(C -> PROJ) between encoder and backbone (SharedLatent3)
    int head_norm, head_cos, skip_identity;
    // loss
    const long long* y;
    float scale; float margin[KMAX]; float cls_w[KMAX]; int nan_degenerate;
    const float* logit_off;        // optional (B,K)
    const float* denom;            // device scalar: sum_b w[y_b] over the GLOBAL batch
    const float* dlogits_ext;      // MODE_BWD_EXT (B,K)
    // outputs
    float* logits;                 // optional (B,K)
    float* partial;                // [gridDim.x][NGP]
    float* dx;                     // ENC_NONE, MODE_BWD_EXT: gradient of the backbone input (B, T, CIN), dense
    // stage cuts (fusion baselines, weargait_encoders.py:209-387 / feature_encoder.py:346-596): a model is run as
    // encoder stages -> fusion op -> trunk stage, with the intermediate tensors in global memory
    float* feat_out;               // encoder stage, forward: encoder output (B, T, C); backbone / head are skipped
    const float* dfeat_in;         // encoder stage, backward: gradient of the encoder output (B, T, C) (replaces the backbone dgrad)
    float* repr_out;               // trunk stage, forward: pooled + flattened backbone features (B, NF); the head is skipped
    const float* drepr_in;         // trunk stage, backward: their gradient (B, NF) (replaces head + loss)
    GradOff go; int NGP;
};

// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float f4get(const float4& v, int e) { return e == 0 ? v.x : e == 1 ? v.y : e == 2 ? v.z : v.w; }

// exact-erf GELU (nn.GELU() default) and its derivative
__device__ __forceinline__ void gelu_fwd(float a, float& g, float& dg) {
    const float cdf = 0.5f * (1.0f + erff(a * 0.70710678118654752440f));
    g = a * cdf;
    dg = cdf + a * 0.39894228040143267794f * __expf(-0.5f * a * a);
}

template <int N>
__device__ __forceinline__ void store_row(float* buf, int RBx, int halo, int r, const float (&v)[N]) {
    static_assert(N % 4 == 0, "");
    float4* p = reinterpret_cast<float4*>(buf) + (halo + r);
#pragma unroll
    for (int c4 = 0; c4 < N / 4; ++c4) p[c4 * RBx] = make_float4(v[c4 * 4], v[c4 * 4 + 1], v[c4 * 4 + 2], v[c4 * 4 + 3]);
}
template <int N>
__device__ __forceinline__ void load_row(const float* buf, int RBx, int halo, int r, float (&v)[N]) {
    static_assert(N % 4 == 0, "");
    const float4* p = reinterpret_cast<const float4*>(buf) + (halo + r);
#pragma unroll
    for (int c4 = 0; c4 < N / 4; ++c4) {
        const float4 t = p[c4 * RBx];
        v[c4 * 4] = t.x; v[c4 * 4 + 1] = t.y; v[c4 * 4 + 2] = t.z; v[c4 * 4 + 3] = t.w;
    }
}

// LayerNorm over the CR real channels of v (biased variance, eps 1e-5): xh, rstd
template <int CP, int CR>
__device__ __forceinline__ void ln_fwd(const float (&v)[CP], float (&xh)[CP], float& rstd) {
    float mu = 0.f;
#pragma unroll
    for (int c = 0; c < CR; ++c) mu += v[c];
    mu *= (1.0f / CR);
    float var = 0.f;
#pragma unroll
    for (int c = 0; c < CR; ++c) { const float d = v[c] - mu; var = fmaf(d, d, var); }
    var *= (1.0f / CR);
    rstd = rsqrtf(var + 1e-5f);
#pragma unroll
    for (int c = 0; c < CP; ++c) xh[c] = c < CR ? (v[c] - mu) * rstd : 0.f;
}
// dv from dxh (= dy * gamma)
template <int CP, int CR>
__device__ __forceinline__ void ln_bwd(const float (&dxh)[CP], const float (&xh)[CP], float rstd, float (&dv)[CP]) {
    float m1 = 0.f, m2 = 0.f;
#pragma unroll
    for (int c = 0; c < CR; ++c) { m1 += dxh[c]; m2 = fmaf(dxh[c], xh[c], m2); }
    m1 *= (1.0f / CR); m2 *= (1.0f / CR);
#pragma unroll
    for (int c = 0; c < CP; ++c) dv[c] = c < CR ? rstd * (dxh[c] - m1 - xh[c] * m2) : 0.f;
}

// GELU and derivative for the tensor-core path: erf by Abramowitz-Stegun 7.1.26 (|err| <= 1.5e-7), sharing
// exp(-a^2/2) with the Gaussian density of the derivative
// raw MUFU forms: __expf / __fdividef carry denormal-range fix-ups (several extra instructions each) that the
// erf approximation does not need (the arguments are O(1); a flushed exp(-a^2/2) below 1e-38 is zero either way)
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ void gelu_fwd_fast(float a, float& g, float& dg) {
    const float e = ex2_approx(a * a * -0.72134752044448170368f);        // exp(-a^2 / 2)
    const float z = fabsf(a) * 0.70710678118654752440f;
    const float t = rcp_approx(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f); p = fmaf(p, t, -0.284496736f); p = fmaf(p, t, 0.254829592f);
    const float er = 1.0f - p * t * e;                       // erf(|a|/sqrt2)
    const float cdf = 0.5f + copysignf(0.5f * er, a);
    g = a * cdf;
    dg = fmaf(a * 0.39894228040143267794f, e, cdf);
}
// ------------------------------------------------------------------------------------------
// compile-time description of one stream
template <int ENC_, int CIN_, int KT1_, int H_, int C_, int S_, int NFL_, int PROJ_ = 0>
struct StreamCfg {
    static constexpr int ENC = ENC_;
    static constexpr int CIN = CIN_;                 // real input channels
    static constexpr int CI4 = (CIN_ + 3) / 4;
    static constexpr int KT1 = KT1_;                 // taps of the first conv (1 for Linear)
    static constexpr int H = H_;                     // insole hidden channels (0 otherwise)
    static constexpr int H4 = (H_ + 3) / 4;
    static constexpr int C = C_;                     // encoder output channels (real)
    static constexpr int C4 = (C_ + 3) / 4;
    static constexpr int CP = C4 * 4;
    static constexpr int S = S_;                     // backbone channels (multiple of 4)
    static constexpr int S4 = S_ / 4;
    static constexpr int NFL = NFL_;                 // head features per lane (NF = 32 * NFL)
    static constexpr int PROJ = PROJ_;               // per-stream projection width (0 = none)
    static constexpr int CB = PROJ_ ? PROJ_ : C_;    // backbone input channels
    static constexpr int CB4 = (CB + 3) / 4;
    static constexpr int CBP = CB4 * 4;
    static constexpr int O1 = (ENC_ == ENC_INSOLE) ? H4 * 4 : CP;   // padded outputs of the first conv
    // resident CTAs per SM the fp32 kernel is compiled for.  The narrow FoG / FBG encoders fit 168 registers with a handful of
    // spilled words (GAITK_FOG_MINB = 3: 12 warps per SM); left alone ptxas takes ~250 registers = 2 CTAs per SM, and these
    // latency-bound kernels then run at IPC 0.9 instead of 1.4 (FoG step at B = 32768: 5.1 vs 3.7 ms, profiles/r3_fog.md -- the
    // round-2 note that occupancy made no difference compared two builds that both ran 3 CTAs per SM).  4 CTAs per SM (128
    // registers) spills 0.5 - 0.9 KB per thread.  The WearGait encoders would spill at 3 (160 - 1140 B): 1 there.
    static constexpr int MINB = (ENC_ == ENC_LINEAR_LN_RELU || ENC_ == ENC_CONV_POOL || ENC_ == ENC_POOL_LINEAR) ? GAITK_FOG_MINB : 1;
};

// shared-memory plan (offsets in floats); filled on the host, passed by value
struct SmemPlan {
    int X, HA, D1, XH, D, F, RSTD, Z, A;            // activation buffers
    int W1F, B1, W2F, B2, W2D, LNG, LNB, WBF, BB, WBD, HW, HB, HNG, HNB, INW;   // weights
    int P, DP, LOGIT, BINS, STAGE;                   // head / pooling scratch
    int L, DL, WPF, BP, WPD;                         // projection stage (SharedLatent3)
    int STG, STGN, MBAR;                             // bulk-prefetch staging: raw bytes of the NEXT tile (STGN floats per window, 0 = off)
    int total;
};


// ------------------------------------------------------------------------------------------
// adaptive pooling + task head + loss + their backward for ONE window, executed by one warp
// (weargait_encoders.py:19-37,110-113; classification_losses.py:54-109).  Lane l owns features l, l+32, ...
struct HeadCtx {
    const float* Zs; int RB, halo, W, S; const int* bin_s; const int* bin_e;
    const float *hws, *hbs, *hngs, *hnbs, *inws; float* DPs;
    const float* Ps;      // optional pre-pooled features [W][NF] (already divided by the bin size)
    const int* ys;        // optional labels of the tile's windows, prefetched to shared memory
    float inv_bin;        // > 0: every pooling bin has 1 / inv_bin frames (a power of two: the product is exact)
};

template <int NFL, int SC>
struct HeadState {
    float g_hw[KMAX][NFL], g_hng[NFL], g_hnb[NFL], g_hb[KMAX];
    float acc_loss, acc_correct;

    __device__ __forceinline__ void zero() {
        acc_loss = 0.f; acc_correct = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            g_hb[k] = 0.f;
#pragma unroll
            for (int i = 0; i < NFL; ++i) g_hw[k][i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < NFL; ++i) { g_hng[i] = 0.f; g_hnb[i] = 0.f; }
    }

    // FAST: MUFU exp / log in the loss (tensor-core path; the fp32 path keeps libdevice's expf / logf)
    // KT: compile-time number of classes (0 = read A.K): with K known the KMAX-wide loops lose their dead iterations
    // and, NF = 32 NFL being a constant, every weight address is an immediate offset
    template <bool FAST = false, int KT = 0>
    __device__ __forceinline__ void run(const StreamArgs& A, const HeadCtx& c, int w_, int lane, int win0, bool train, float inv_denom) {
            const int wi = win0 + w_; const int K = KT ? KT : A.K; constexpr int NF = NFL * 32;
            float f[NFL], xn[NFL], xh[NFL];
            float rstd_h = 1.f;
#pragma unroll
            for (int i = 0; i < NFL; ++i) {
                const int j = lane + 32 * i, b = j / SC, s = j - b * SC;
                if (c.Ps) { f[i] = c.Ps[w_ * NF + j]; continue; }
                const int t0 = c.bin_s[b], t1 = c.bin_e[b];
                float acc = 0.f;
                for (int t = t0; t < t1; ++t) acc += c.Zs[((s >> 2) * c.RB + c.halo + t * c.W + w_) * 4 + (s & 3)];
                f[i] = acc / (float)(t1 - t0);
            }
            if (A.repr_out || A.drepr_in) {           // trunk stage: features out / feature gradients in, no head
                if (A.repr_out && wi < A.B) {
#pragma unroll
                    for (int i = 0; i < NFL; ++i) A.repr_out[(size_t)wi * NF + lane + 32 * i] = f[i];
                }
                if (A.drepr_in) {
#pragma unroll
                    for (int i = 0; i < NFL; ++i) {
                        const int j = lane + 32 * i, b = j / SC;
                        const float g = wi < A.B ? A.drepr_in[(size_t)wi * NF + j] : 0.f;
                        c.DPs[w_ * NF + j] = c.inv_bin > 0.f ? g * c.inv_bin : g / (float)(c.bin_e[b] - c.bin_s[b]);
                    }
                }
                return;
            }
            if (A.head_norm) {
                float m = 0.f;
#pragma unroll
                for (int i = 0; i < NFL; ++i) m += f[i];
                m = warp_sum(m) / (float)NF;
                float v = 0.f;
#pragma unroll
                for (int i = 0; i < NFL; ++i) { const float d = f[i] - m; v = fmaf(d, d, v); }
                v = warp_sum(v) / (float)NF;
                rstd_h = rsqrtf(v + 1e-5f);
#pragma unroll
                for (int i = 0; i < NFL; ++i) { xh[i] = (f[i] - m) * rstd_h; xn[i] = fmaf(xh[i], c.hngs[lane + 32 * i], c.hnbs[lane + 32 * i]); }
            } else {
#pragma unroll
                for (int i = 0; i < NFL; ++i) { xh[i] = 0.f; xn[i] = f[i]; }
            }
            float nx = 0.f, inx = 1.f;
            if (A.head_cos) {
#pragma unroll
                for (int i = 0; i < NFL; ++i) nx = fmaf(xn[i], xn[i], nx);
                nx = sqrtf(warp_sum(nx));
                inx = 1.0f / fmaxf(nx, 1e-8f);
            }
            float logit[KMAX], dot[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                logit[k] = 0.f; dot[k] = 0.f;
                if (k < K) {
                    float d = 0.f;
#pragma unroll
                    for (int i = 0; i < NFL; ++i) d = fmaf(xn[i], c.hws[k * NF + lane + 32 * i], d);
                    d = warp_sum(d);
                    dot[k] = d;
                    if (A.head_cos) logit[k] = fminf(fmaxf(d * inx * c.inws[k], -1.0f + 1e-8f), 1.0f - 1e-8f);
                    else logit[k] = d + c.hbs[k];
                }
            }
            if (wi < A.B && A.logits && lane < K)
                A.logits[(size_t)wi * K + lane] = lane == 0 ? logit[0] : lane == 1 ? logit[1] : lane == 2 ? logit[2] : logit[3];
            if (train) {
                float dl[KMAX];
#pragma unroll
                for (int k = 0; k < KMAX; ++k) dl[k] = 0.f;
                if (wi < A.B) {
                    if (A.mode == MODE_FUSED) {
                        const int y = c.ys ? c.ys[w_] : (int)A.y[wi];
                        float zz[KMAX]; float mx = -INFINITY; int am = 0; float best = -INFINITY;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < K) {
                            float z = logit[k];
                            if (A.logit_off) z -= A.logit_off[(size_t)wi * K + k];
                            if (k == y) z -= A.margin[k];
                            z *= A.scale;
                            if (A.nan_degenerate) z = __int_as_float(0x7fc00000);
                            zz[k] = z; mx = fmaxf(mx, z);
                            if (logit[k] > best) { best = logit[k]; am = k; }
                        }
                        float se = 0.f;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < K) se += FAST ? __expf(zz[k] - mx) : expf(zz[k] - mx);
                        const float lse = mx + (FAST ? __logf(se) : logf(se));
                        float zy = 0.f, wy = 0.f;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < K && k == y) { zy = zz[k]; wy = A.cls_w[k]; }
                        acc_loss += wy * (lse - zy) * inv_denom;
                        acc_correct += (am == y) ? 1.f : 0.f;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < K)
                            dl[k] = A.scale * wy * inv_denom * ((FAST ? __expf(zz[k] - lse) : expf(zz[k] - lse)) - (k == y ? 1.f : 0.f));
                    } else {
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < K) dl[k] = A.dlogits_ext[(size_t)wi * K + k];
                    }
                }
                // head backward
                float dxn[NFL];
#pragma unroll
                for (int i = 0; i < NFL; ++i) dxn[i] = 0.f;
                if (A.head_cos) {
                    float dinx = 0.f;
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) if (k < K) {
                        const float cv = dot[k] * inx * c.inws[k];
                        const float g = (cv >= -1.0f + 1e-8f && cv <= 1.0f - 1e-8f) ? dl[k] : 0.f;
                        const float ddot = g * inx * c.inws[k];
                        dinx = fmaf(g, dot[k] * c.inws[k], dinx);
                        const float dinw = g * dot[k] * inx;                 // d/d(inw_k)
                        const float iw = c.inws[k];
                        const bool wfree = iw < 1e8f;                        // ||w|| > eps
#pragma unroll
                        for (int i = 0; i < NFL; ++i) {
                            const float wkj = c.hws[k * NF + lane + 32 * i];
                            dxn[i] = fmaf(ddot, wkj, dxn[i]);
                            float gw = ddot * xn[i];
                            if (wfree) gw = fmaf(dinw, -wkj * iw * iw * iw, gw);
                            g_hw[k][i] += gw;
                        }
                    }
                    if (nx > 1e-8f) {
                        const float cfac = -dinx * inx * inx * inx;
#pragma unroll
                        for (int i = 0; i < NFL; ++i) dxn[i] = fmaf(cfac, xn[i], dxn[i]);
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < KMAX; ++k) if (k < K) {
                        g_hb[k] += dl[k];
#pragma unroll
                        for (int i = 0; i < NFL; ++i) {
                            dxn[i] = fmaf(dl[k], c.hws[k * NF + lane + 32 * i], dxn[i]);
                            g_hw[k][i] = fmaf(dl[k], xn[i], g_hw[k][i]);
                        }
                    }
                }
                float df[NFL];
                if (A.head_norm) {
                    float m1 = 0.f, m2 = 0.f; float dxh[NFL];
#pragma unroll
                    for (int i = 0; i < NFL; ++i) {
                        g_hng[i] = fmaf(dxn[i], xh[i], g_hng[i]); g_hnb[i] += dxn[i];
                        dxh[i] = dxn[i] * c.hngs[lane + 32 * i];
                        m1 += dxh[i]; m2 = fmaf(dxh[i], xh[i], m2);
                    }
                    m1 = warp_sum(m1) / (float)NF; m2 = warp_sum(m2) / (float)NF;
#pragma unroll
                    for (int i = 0; i < NFL; ++i) df[i] = rstd_h * (dxh[i] - m1 - xh[i] * m2);
                } else {
#pragma unroll
                    for (int i = 0; i < NFL; ++i) df[i] = dxn[i];
                }
#pragma unroll
                for (int i = 0; i < NFL; ++i) {
                    const int j = lane + 32 * i, b = j / SC;
                    c.DPs[w_ * NF + j] = c.inv_bin > 0.f ? df[i] * c.inv_bin : df[i] / (float)(c.bin_e[b] - c.bin_s[b]);
                }
            }
            }

    // per warp (window slot), per lane (feature) accumulators: sum the warps in fixed order
    __device__ __forceinline__ void flush(const StreamArgs& A, float* stage, float* out, int tid) {
        const int lane = tid & 31, wrp = tid >> 5, K = A.K, NF = A.NF;
        const GradOff& go = A.go;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K)
#pragma unroll
            for (int i = 0; i < NFL; ++i) stage[(wrp * KMAX + k) * NF + lane + 32 * i] = g_hw[k][i];
        __syncthreads();
        for (int e = tid; e < K * NF; e += NT) {
            const int k = e / NF, j = e - k * NF;
            float s = 0.f;
            for (int w = 0; w < NT / 32; ++w) s += stage[(w * KMAX + k) * NF + j];
            out[go.hw + e] = s;
        }
        __syncthreads();
        if (A.head_norm) {
#pragma unroll
            for (int i = 0; i < NFL; ++i) { stage[wrp * 2 * NF + lane + 32 * i] = g_hng[i]; stage[wrp * 2 * NF + NF + lane + 32 * i] = g_hnb[i]; }
            __syncthreads();
            for (int e = tid; e < 2 * NF; e += NT) {
                float s = 0.f;
                for (int w = 0; w < NT / 32; ++w) s += stage[w * 2 * NF + e];
                if (e < NF) out[go.hng + e] = s; else out[go.hnb + e - NF] = s;
            }
            __syncthreads();
        }
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < KMAX; ++k) stage[wrp * 8 + k] = g_hb[k];
            stage[wrp * 8 + 4] = acc_loss; stage[wrp * 8 + 5] = acc_correct;
        }
        __syncthreads();
        if (tid < 6) {
            float s = 0.f;
            for (int w = 0; w < NT / 32; ++w) s += stage[w * 8 + tid];
            if (tid < 4) { if (tid < K && go.hb >= 0) out[go.hb + tid] = s; }
            else out[go.total + (tid - 4)] = s;      // [NG] = loss, [NG+1] = correct
        }
    }
};

// per-row-thread accumulators (bias / LayerNorm affine grads): deterministic block sum of N values
template <int N>
__device__ __forceinline__ void flush_rowacc(const float (&v)[N], int nreal, float* stage, float* dst, float* dst2, int tid) {
    __syncthreads();
    const int lane = tid & 31, wrp = tid >> 5;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const float s = warp_sum(v[i]);
        if (lane == 0) stage[i * 4 + wrp] = s;
    }
    __syncthreads();
    if (tid < nreal) {
        const float s = (stage[tid * 4] + stage[tid * 4 + 1]) + (stage[tid * 4 + 2] + stage[tid * 4 + 3]);
        dst[tid] = s;
        if (dst2) dst2[tid] = s;
    }
    __syncthreads();
}


}  // namespace gaitk
