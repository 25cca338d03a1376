// stream_kernel.cuh -- one gait stream (encoder -> shared backbone -> head -> loss) forward AND
// backward in a single persistent kernel, fp32 FFMA arithmetic ("GAITK_DTYPE_F32" path).
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

enum EncKind { ENC_CONV_GELU_LN = 0, ENC_INSOLE = 1, ENC_LINEAR_LN_RELU = 2, ENC_CONV_POOL = 3 };
enum Mode { MODE_FWD = 0, MODE_FUSED = 1, MODE_BWD_EXT = 2 };

// stream-local gradient layout (offsets in floats into a partial-gradient row; -1 = absent)
struct GradOff {
    int w1, b1, w2, b2, wsk, bsk, lng, lnb, wbb, bbb, hng, hnb, hw, hb;
    int total;      // NG
};

struct StreamArgs {
    // input
    const float* x;                // (B, T_in, CIN) dense, or frame store (N, CIN) with win_start
    const long long* win_start;    // optional [B] first frame of each window
    int B, T_in, T, W, bdim, K, NF;
    int rows_in, rows, halo, RBi, RB;
    int mode, zero_input, pool_sensor;
    // parameters (global memory, PyTorch layouts)
    const float *w1, *b1, *w2, *b2, *wsk, *bsk, *lng, *lnb, *wbb, *bbb, *hng, *hnb, *hw, *hb;
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
    float* dx;                     // optional input gradient (dense layout) -- MODE_BWD_EXT only
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

// out[o] = bias[o] + sum_{tap, ci} in[row + (tap - KT/2) * W][ci] * wf[tap][ci][o]
// `in` is a chunked buffer with RBx rows per chunk; wf/bias live in shared memory.
template <int KT, int CI4, int CO>
__device__ __forceinline__ void conv_row(const float* __restrict__ in, int RBx, int halo, int W, int r,
                                         const float* __restrict__ wf, const float* __restrict__ bias,
                                         float (&acc)[CO]) {
    static_assert(CO % 4 == 0, "CO must be a multiple of 4");
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = bias ? bias[o] : 0.f;
#pragma unroll
    for (int tap = 0; tap < KT; ++tap) {
        const float4* xp = reinterpret_cast<const float4*>(in) + (halo + r + (tap - KT / 2) * W);
#pragma unroll
        for (int c4 = 0; c4 < CI4; ++c4) {
            const float4 x = xp[c4 * RBx];
            const float4* wp = reinterpret_cast<const float4*>(wf + (size_t)(tap * CI4 * 4 + c4 * 4) * CO);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float xe = f4get(x, e);
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4) {
                    const float4 w = wp[e * (CO / 4) + o4];
                    acc[o4 * 4 + 0] = fmaf(xe, w.x, acc[o4 * 4 + 0]);
                    acc[o4 * 4 + 1] = fmaf(xe, w.y, acc[o4 * 4 + 1]);
                    acc[o4 * 4 + 2] = fmaf(xe, w.z, acc[o4 * 4 + 2]);
                    acc[o4 * 4 + 3] = fmaf(xe, w.w, acc[o4 * 4 + 3]);
                }
            }
        }
    }
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

// ------------------------------------------------------------------------------------------
// weight-gradient blocks: dW[tap][ci][co] += sum_r in[r + (tap-KT/2)W][ci] * dout[r][co].
// A thread owns 4(ci) x 4(co) register blocks that persist across all tiles of the kernel.
template <int KT, int CI4, int CO4>
struct Wgrad {
    static constexpr int NBLK = KT * CI4 * CO4;
    static constexpr int NB = (NBLK + NT - 1) / NT;              // blocks per thread
    static constexpr int SLICES = NBLK >= NT ? 1 : NT / NBLK;    // row slices when blocks are few
    float a[NB][16];

    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int b = 0; b < NB; ++b)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[b][i] = 0.f;
    }
    __device__ __forceinline__ void accumulate(const float* __restrict__ in, int RBin, const float* __restrict__ dout,
                                               int RBout, int halo, int W, int rows, int tid) {
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int blk = (SLICES == 1) ? tid + b * NT : tid % NBLK;
            const int sl = (SLICES == 1) ? 0 : tid / NBLK;
            if (blk >= NBLK || sl >= SLICES) continue;
            const int per = (rows + SLICES - 1) / SLICES;
            const int r0 = sl * per, r1 = min(rows, r0 + per);
            const int tap = blk / (CI4 * CO4), ci4 = (blk / CO4) % CI4, co4 = blk % CO4;
            const float4* xp = reinterpret_cast<const float4*>(in) + (size_t)ci4 * RBin + halo + (tap - KT / 2) * W;
            const float4* dp = reinterpret_cast<const float4*>(dout) + (size_t)co4 * RBout + halo;
            float* acc = a[b];
#pragma unroll 4
            for (int r = r0; r < r1; ++r) {
                const float4 x = xp[r];
                const float4 d = dp[r];
                acc[0] = fmaf(x.x, d.x, acc[0]);  acc[1] = fmaf(x.x, d.y, acc[1]);  acc[2] = fmaf(x.x, d.z, acc[2]);  acc[3] = fmaf(x.x, d.w, acc[3]);
                acc[4] = fmaf(x.y, d.x, acc[4]);  acc[5] = fmaf(x.y, d.y, acc[5]);  acc[6] = fmaf(x.y, d.z, acc[6]);  acc[7] = fmaf(x.y, d.w, acc[7]);
                acc[8] = fmaf(x.z, d.x, acc[8]);  acc[9] = fmaf(x.z, d.y, acc[9]);  acc[10] = fmaf(x.z, d.z, acc[10]); acc[11] = fmaf(x.z, d.w, acc[11]);
                acc[12] = fmaf(x.w, d.x, acc[12]); acc[13] = fmaf(x.w, d.y, acc[13]); acc[14] = fmaf(x.w, d.z, acc[14]); acc[15] = fmaf(x.w, d.w, acc[15]);
            }
        }
    }
    // Deterministic flush: slices are summed in fixed order through `stage` (>= SLICES*NBLK*16
    // floats of shared memory), then scattered to the PyTorch weight layout (CO, CI, KT).
    // dst2 (optional) receives the centre tap only, layout (CO, CI, 1): the folded 1x1 skip.
    __device__ __forceinline__ void flush(float* stage, float* dst, float* dst2, int CIN, int COUT, int tid) {
        __syncthreads();
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int blk = (SLICES == 1) ? tid + b * NT : tid % NBLK;
            const int sl = (SLICES == 1) ? 0 : tid / NBLK;
            if (blk >= NBLK || sl >= SLICES) continue;
#pragma unroll
            for (int i = 0; i < 16; ++i) stage[(sl * NBLK + blk) * 16 + i] = a[b][i];
        }
        __syncthreads();
        for (int e = tid; e < NBLK * 16; e += NT) {
            float s = 0.f;
            for (int sl = 0; sl < SLICES; ++sl) s += stage[sl * NBLK * 16 + e];
            const int blk = e / 16, i = (e % 16) / 4, j = e % 4;
            const int tap = blk / (CI4 * CO4), ci = ((blk / CO4) % CI4) * 4 + i, co = (blk % CO4) * 4 + j;
            if (ci < CIN && co < COUT) {
                dst[(co * CIN + ci) * KT + tap] = s;
                if (dst2 && tap == KT / 2) dst2[co * CIN + ci] = s;
            }
        }
        __syncthreads();
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

// ------------------------------------------------------------------------------------------
// compile-time description of one stream
template <int ENC_, int CIN_, int KT1_, int H_, int C_, int S_, int NFL_>
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
    static constexpr int O1 = (ENC_ == ENC_INSOLE) ? H4 * 4 : CP;   // padded outputs of the first conv
};

// shared-memory plan (offsets in floats); filled on the host, passed by value
struct SmemPlan {
    int X, HA, D1, XH, D, F, RSTD, Z, A;            // activation buffers
    int W1F, B1, W2F, B2, W2D, LNG, LNB, WBF, BB, WBD, HW, HB, HNG, HNB, INW;   // weights
    int P, DP, LOGIT, BINS, STAGE;                   // head / pooling scratch
    int total;
};

template <class Cfg>
__global__ void __launch_bounds__(NT) stream_kernel(const StreamArgs A, const SmemPlan SP) {
    extern __shared__ __align__(16) float sm[];
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    constexpr int ENC = Cfg::ENC, CIN = Cfg::CIN, CI4 = Cfg::CI4, KT1 = Cfg::KT1, H = Cfg::H, H4 = Cfg::H4;
    constexpr int C = Cfg::C, C4 = Cfg::C4, CP = Cfg::CP, S = Cfg::S, S4 = Cfg::S4, NFL = Cfg::NFL, O1 = Cfg::O1;
    const int W = A.W, halo = A.halo, RB = A.RB, RBi = A.RBi, rows = A.rows, rows_in = A.rows_in, T = A.T, T_in = A.T_in;
    const int K = A.K, NF = A.NF, bdim = A.bdim;
    const bool train = A.mode != MODE_FWD;

    float* Xs = sm + SP.X;   float* HAs = sm + SP.HA; float* D1s = sm + SP.D1; float* XHs = sm + SP.XH;
    float* Ds = sm + SP.D;   float* Fs = sm + SP.F;   float* RSTDs = sm + SP.RSTD; float* Zs = sm + SP.Z;
    float* As = sm + SP.A;
    float* w1f = sm + SP.W1F; float* b1s = sm + SP.B1; float* w2f = sm + SP.W2F; float* b2s = sm + SP.B2;
    float* w2d = sm + SP.W2D; float* lngs = sm + SP.LNG; float* lnbs = sm + SP.LNB;
    float* wbf = sm + SP.WBF; float* bbs = sm + SP.BB; float* wbd = sm + SP.WBD;
    float* hws = sm + SP.HW; float* hbs = sm + SP.HB; float* hngs = sm + SP.HNG; float* hnbs = sm + SP.HNB;
    float* inws = sm + SP.INW;
    float* Ps = sm + SP.P; float* DPs = sm + SP.DP; float* LGs = sm + SP.LOGIT;
    int* bins = reinterpret_cast<int*>(sm + SP.BINS);   // [0,bdim) start, [bdim,2bdim) end, then per-t lo / hi, then
                                                        // encoder-pool tables (CONV_POOL)
    float* stage = sm + SP.STAGE;

    // ---- one-time setup: zero everything (halos, padded channels), stage weights, bin tables
    for (int i = tid; i < SP.total; i += NT) sm[i] = 0.f;
    __syncthreads();
    // first conv / linear: PyTorch (O, CIN, KT1) -> wf[tap][ci][o]
    {
        const int OUT1 = (ENC == ENC_INSOLE) ? H : C;
        for (int i = tid; i < OUT1 * CIN * KT1; i += NT) {
            const int o = i / (CIN * KT1), ci = (i / KT1) % CIN, tap = i % KT1;
            w1f[(tap * CI4 * 4 + ci) * O1 + o] = A.w1[i];
        }
        for (int i = tid; i < OUT1; i += NT) b1s[i] = A.b1[i];
    }
    if constexpr (ENC == ENC_INSOLE) {
        // conv2 (C, H, 3) with the 1x1 skip (C, H, 1) folded into the centre tap (identity skip if H == C)
        for (int i = tid; i < C * H * 3; i += NT) {
            const int o = i / (H * 3), ci = (i / 3) % H, tap = i % 3;
            float w = A.w2[i];
            if (tap == 1) w += A.skip_identity ? (o == ci ? 1.f : 0.f) : A.wsk[o * H + ci];
            w2f[(tap * H4 * 4 + ci) * CP + o] = w;
            w2d[((2 - tap) * CP + o) * (H4 * 4) + ci] = w;      // dgrad: flipped taps, [tap][co][ci]
        }
        for (int i = tid; i < C; i += NT) b2s[i] = A.b2[i] + (A.skip_identity ? 0.f : A.bsk[i]);
    }
    if constexpr (ENC != ENC_CONV_POOL) {
        for (int i = tid; i < C; i += NT) { lngs[i] = A.lng[i]; lnbs[i] = A.lnb[i]; }
    }
    for (int i = tid; i < S * C * 3; i += NT) {
        const int o = i / (C * 3), ci = (i / 3) % C, tap = i % 3;
        const float w = A.wbb[i];
        wbf[(tap * CP + ci) * S + o] = w;
        wbd[((2 - tap) * S + o) * CP + ci] = w;
    }
    for (int i = tid; i < S; i += NT) bbs[i] = A.bbb[i];
    for (int i = tid; i < K * NF; i += NT) hws[i] = A.hw[i];
    if (A.hb) for (int i = tid; i < K; i += NT) hbs[i] = A.hb[i];
    if (A.head_norm) for (int i = tid; i < NF; i += NT) { hngs[i] = A.hng[i]; hnbs[i] = A.hnb[i]; }
    // adaptive pooling tables: bin b covers [floor(b*T/bdim), ceil((b+1)*T/bdim))
    int* bin_s = bins; int* bin_e = bins + bdim; int* t_lo = bins + 2 * bdim; int* t_hi = t_lo + T;
    for (int b = tid; b < bdim; b += NT) { bin_s[b] = (b * T) / bdim; bin_e[b] = ((b + 1) * T + bdim - 1) / bdim; }
    int* ep_s = t_hi + T; int* ep_e = ep_s + T; int* ep_lo = ep_e + T; int* ep_hi = ep_lo + T_in;   // encoder pool (T_in -> T)
    if constexpr (ENC == ENC_CONV_POOL) {
        if (A.pool_sensor)
            for (int i = tid; i < T; i += NT) { ep_s[i] = (i * T_in) / T; ep_e[i] = ((i + 1) * T_in + T - 1) / T; }
    }
    __syncthreads();
    for (int t = tid; t < T; t += NT) {
        int lo = bdim, hi = -1;
        for (int b = 0; b < bdim; ++b) if (t >= bin_s[b] && t < bin_e[b]) { lo = min(lo, b); hi = max(hi, b); }
        t_lo[t] = lo; t_hi[t] = hi;
    }
    if constexpr (ENC == ENC_CONV_POOL) {
        if (A.pool_sensor)
            for (int t = tid; t < T_in; t += NT) {
                int lo = T, hi = -1;
                // bins are monotone: search around the proportional position
                const int guess = (int)(((long long)t * T) / T_in);
                for (int i = max(0, guess - 2); i <= min(T - 1, guess + 2); ++i)
                    if (t >= ep_s[i] && t < ep_e[i]) { lo = min(lo, i); hi = max(hi, i); }
                ep_lo[t] = lo; ep_hi[t] = hi;
            }
    }
    if (A.head_cos) {   // 1 / max(||w_k||, 1e-8)
        if (wrp < K) {
            float s = 0.f;
            for (int j = lane; j < NF; j += 32) s = fmaf(hws[wrp * NF + j], hws[wrp * NF + j], s);
            s = warp_sum(s);
            if (lane == 0) inws[wrp] = 1.0f / fmaxf(sqrtf(s), 1e-8f);
        }
    }
    __syncthreads();

    // ---- persistent accumulators
    Wgrad<KT1, CI4, O1 / 4> g_w1;
    Wgrad<3, (ENC == ENC_INSOLE ? H4 : 1), (ENC == ENC_INSOLE ? C4 : 1)> g_w2;
    Wgrad<3, C4, S4> g_wb;
    float g_b1[O1], g_b2[ENC == ENC_INSOLE ? CP : 4], g_lng[CP], g_lnb[CP], g_bb[S];
    float g_hw[KMAX][NFL], g_hng[NFL], g_hnb[NFL], g_hb[KMAX];
    float acc_loss = 0.f, acc_correct = 0.f;
    if (train) {
        g_w1.zero(); g_w2.zero(); g_wb.zero();
#pragma unroll
        for (int i = 0; i < O1; ++i) g_b1[i] = 0.f;
#pragma unroll
        for (int i = 0; i < (ENC == ENC_INSOLE ? CP : 4); ++i) g_b2[i] = 0.f;
#pragma unroll
        for (int i = 0; i < CP; ++i) { g_lng[i] = 0.f; g_lnb[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < S; ++i) g_bb[i] = 0.f;
#pragma unroll
        for (int k = 0; k < KMAX; ++k) {
            g_hb[k] = 0.f;
#pragma unroll
            for (int i = 0; i < NFL; ++i) g_hw[k][i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < NFL; ++i) { g_hng[i] = 0.f; g_hnb[i] = 0.f; }
    }
    const float inv_denom = (A.mode == MODE_FUSED) ? 1.0f / A.denom[0] : 0.f;

    const int ntiles = (A.B + W - 1) / W;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int win0 = tile * W;
        // ================= load: global (window-major) -> Xs [chunk][row][4]
        {
            const int per_win = T_in * CIN;
            for (int e = tid; e < W * per_win; e += NT) {
                const int w = e / per_win, rem = e - w * per_win;
                const int t = rem / CIN, c = rem - t * CIN;
                const int wi = win0 + w;
                float v = 0.f;
                if (wi < A.B && !A.zero_input) {
                    const size_t base = A.win_start ? (size_t)A.win_start[wi] * CIN : (size_t)wi * per_win;
                    v = __ldg(A.x + base + rem);
                }
                Xs[((c >> 2) * RBi + halo + t * W + w) * 4 + (c & 3)] = v;
            }
        }
        __syncthreads();
        // ================= encoder forward
        if constexpr (ENC == ENC_INSOLE) {
            for (int r = tid; r < rows; r += NT) {
                float a1[O1], ha[O1], d1[O1];
                conv_row<KT1, CI4, O1>(Xs, RBi, halo, W, r, w1f, b1s, a1);
#pragma unroll
                for (int c = 0; c < O1; ++c) { if (c < H) gelu_fwd(a1[c], ha[c], d1[c]); else { ha[c] = 0.f; d1[c] = 0.f; } }
                store_row<O1>(HAs, RB, halo, r, ha);
                if (train) store_row<O1>(D1s, RB, halo, r, d1);
            }
            __syncthreads();
        }
        if constexpr (ENC == ENC_CONV_POOL) {
            // conv over the T_in input rows, no activation; optional adaptive pooling T_in -> T
            for (int r = tid; r < rows_in; r += NT) {
                float a[CP];
                conv_row<KT1, CI4, CP>(Xs, RBi, halo, W, r, w1f, b1s, a);
                store_row<CP>(A.pool_sensor ? As : Fs, A.pool_sensor ? RBi : RB, halo, r, a);
            }
            __syncthreads();
            if (A.pool_sensor) {
                for (int r = tid; r < rows; r += NT) {      // W == 1 for pooled streams
                    float f[CP];
#pragma unroll
                    for (int c = 0; c < CP; ++c) f[c] = 0.f;
                    const int s0 = ep_s[r], s1 = ep_e[r];
                    for (int t = s0; t < s1; ++t) {
                        float a[CP]; load_row<CP>(As, RBi, halo, t, a);
#pragma unroll
                        for (int c = 0; c < CP; ++c) f[c] += a[c];
                    }
                    const float inv = 1.0f / (float)(s1 - s0);
#pragma unroll
                    for (int c = 0; c < CP; ++c) f[c] *= inv;
                    store_row<CP>(Fs, RB, halo, r, f);
                }
                __syncthreads();
            }
        } else {
            for (int r = tid; r < rows; r += NT) {
                float a[CP], g[CP], d[CP], xh[CP], f[CP]; float rstd;
                if constexpr (ENC == ENC_INSOLE) conv_row<3, H4, CP>(HAs, RB, halo, W, r, w2f, b2s, a);
                else conv_row<KT1, CI4, CP>(Xs, RBi, halo, W, r, w1f, b1s, a);
                if constexpr (ENC == ENC_LINEAR_LN_RELU) {
                    ln_fwd<CP, C>(a, xh, rstd);
#pragma unroll
                    for (int c = 0; c < CP; ++c) f[c] = c < C ? fmaxf(fmaf(xh[c], lngs[c], lnbs[c]), 0.f) : 0.f;
                } else {
#pragma unroll
                    for (int c = 0; c < CP; ++c) { if (c < C) gelu_fwd(a[c], g[c], d[c]); else { g[c] = 0.f; d[c] = 0.f; } }
                    ln_fwd<CP, C>(g, xh, rstd);
#pragma unroll
                    for (int c = 0; c < CP; ++c) f[c] = c < C ? fmaf(xh[c], lngs[c], lnbs[c]) : 0.f;
                    if (train) store_row<CP>(Ds, RB, halo, r, d);
                }
                store_row<CP>(Fs, RB, halo, r, f);
                if (train) { store_row<CP>(XHs, RB, halo, r, xh); RSTDs[r] = rstd; }
            }
            __syncthreads();
        }
        // ================= shared backbone forward: conv3 -> ReLU
        for (int r = tid; r < rows; r += NT) {
            float z[S];
            conv_row<3, C4, S>(Fs, RB, halo, W, r, wbf, bbs, z);
#pragma unroll
            for (int s = 0; s < S; ++s) z[s] = fmaxf(z[s], 0.f);
            store_row<S>(Zs, RB, halo, r, z);
        }
        __syncthreads();
        // ================= adaptive pool + head + loss: warp w <-> window w of the tile
        if (wrp < W) {
            const int wi = win0 + wrp;
            float f[NFL], xn[NFL], xh[NFL];
            float rstd_h = 1.f;
#pragma unroll
            for (int i = 0; i < NFL; ++i) {
                const int j = lane + 32 * i, b = j / S, s = j - b * S;
                const int t0 = bin_s[b], t1 = bin_e[b];
                float acc = 0.f;
                for (int t = t0; t < t1; ++t) acc += Zs[((s >> 2) * RB + halo + t * W + wrp) * 4 + (s & 3)];
                f[i] = acc / (float)(t1 - t0);
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
                for (int i = 0; i < NFL; ++i) { xh[i] = (f[i] - m) * rstd_h; xn[i] = fmaf(xh[i], hngs[lane + 32 * i], hnbs[lane + 32 * i]); }
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
                    for (int i = 0; i < NFL; ++i) d = fmaf(xn[i], hws[k * NF + lane + 32 * i], d);
                    d = warp_sum(d);
                    dot[k] = d;
                    if (A.head_cos) logit[k] = fminf(fmaxf(d * inx * inws[k], -1.0f + 1e-8f), 1.0f - 1e-8f);
                    else logit[k] = d + hbs[k];
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
                        const int y = (int)A.y[wi];
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
                        for (int k = 0; k < KMAX; ++k) if (k < K) se += expf(zz[k] - mx);
                        const float lse = mx + logf(se);
                        float zy = 0.f, wy = 0.f;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < K && k == y) { zy = zz[k]; wy = A.cls_w[k]; }
                        acc_loss += wy * (lse - zy) * inv_denom;
                        acc_correct += (am == y) ? 1.f : 0.f;
#pragma unroll
                        for (int k = 0; k < KMAX; ++k) if (k < K)
                            dl[k] = A.scale * wy * inv_denom * (expf(zz[k] - lse) - (k == y ? 1.f : 0.f));
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
                        const float cv = dot[k] * inx * inws[k];
                        const float g = (cv >= -1.0f + 1e-8f && cv <= 1.0f - 1e-8f) ? dl[k] : 0.f;
                        const float ddot = g * inx * inws[k];
                        dinx = fmaf(g, dot[k] * inws[k], dinx);
                        const float dinw = g * dot[k] * inx;                 // d/d(inw_k)
                        const float iw = inws[k];
                        const bool wfree = iw < 1e8f;                        // ||w|| > eps
#pragma unroll
                        for (int i = 0; i < NFL; ++i) {
                            const float wkj = hws[k * NF + lane + 32 * i];
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
                            dxn[i] = fmaf(dl[k], hws[k * NF + lane + 32 * i], dxn[i]);
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
                        dxh[i] = dxn[i] * hngs[lane + 32 * i];
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
                    const int j = lane + 32 * i, b = j / S;
                    DPs[wrp * NF + j] = df[i] / (float)(bin_e[b] - bin_s[b]);
                }
            }
        }
        if (!train) { __syncthreads(); continue; }
        __syncthreads();
        // ================= backbone backward: dz (through pool + ReLU), in place over Z
        for (int r = tid; r < rows; r += NT) {
            const int t = r / W, w = r - t * W;
            float z[S], dz[S];
            load_row<S>(Zs, RB, halo, r, z);
            const int lo = t_lo[t], hi = t_hi[t];
#pragma unroll
            for (int s = 0; s < S; ++s) {
                float d = 0.f;
                for (int b = lo; b <= hi; ++b) d += DPs[w * NF + b * S + s];
                dz[s] = z[s] > 0.f ? d : 0.f;
                g_bb[s] += dz[s];
            }
            store_row<S>(Zs, RB, halo, r, dz);
        }
        __syncthreads();
        g_wb.accumulate(Fs, RB, Zs, RB, halo, W, rows, tid);
        // dgrad into the encoder output + encoder-specific backward up to the first-conv pre-activation
        if constexpr (ENC == ENC_CONV_POOL) {
            for (int r = tid; r < rows; r += NT) {
                float df[CP];
                conv_row<3, S4, CP>(Zs, RB, halo, W, r, wbd, nullptr, df);
                if (A.pool_sensor) {
                    const float inv = 1.0f / (float)(ep_e[r] - ep_s[r]);
#pragma unroll
                    for (int c = 0; c < CP; ++c) df[c] *= inv;
                    store_row<CP>(XHs, RB, halo, r, df);           // dF / binsize (feature rows)
                } else {
#pragma unroll
                    for (int c = 0; c < CP; ++c) g_b1[c] += df[c];
                    store_row<CP>(As, RBi, halo, r, df);            // dA == dF
                }
            }
            __syncthreads();
            if (A.pool_sensor) {
                for (int t = tid; t < rows_in; t += NT) {
                    float da[CP];
#pragma unroll
                    for (int c = 0; c < CP; ++c) da[c] = 0.f;
                    for (int i = ep_lo[t]; i <= ep_hi[t]; ++i) {
                        float d[CP]; load_row<CP>(XHs, RB, halo, i, d);
#pragma unroll
                        for (int c = 0; c < CP; ++c) da[c] += d[c];
                    }
#pragma unroll
                    for (int c = 0; c < CP; ++c) g_b1[c] += da[c];
                    store_row<CP>(As, RBi, halo, t, da);
                }
                __syncthreads();
            }
            g_w1.accumulate(Xs, RBi, As, RBi, halo, W, rows_in, tid);
        } else {
            for (int r = tid; r < rows; r += NT) {
                float df[CP], xh[CP], dxh[CP], dg[CP], da[CP];
                conv_row<3, S4, CP>(Zs, RB, halo, W, r, wbd, nullptr, df);
                load_row<CP>(XHs, RB, halo, r, xh);
                const float rstd = RSTDs[r];
                if constexpr (ENC == ENC_LINEAR_LN_RELU) {
                    float f[CP]; load_row<CP>(Fs, RB, halo, r, f);
#pragma unroll
                    for (int c = 0; c < CP; ++c) if (!(f[c] > 0.f)) df[c] = 0.f;
                }
#pragma unroll
                for (int c = 0; c < CP; ++c) {
                    g_lng[c] = fmaf(df[c], xh[c], g_lng[c]); g_lnb[c] += df[c];
                    dxh[c] = c < C ? df[c] * lngs[c] : 0.f;
                }
                ln_bwd<CP, C>(dxh, xh, rstd, dg);
                if constexpr (ENC == ENC_LINEAR_LN_RELU) {
#pragma unroll
                    for (int c = 0; c < CP; ++c) da[c] = dg[c];
                } else {
                    float d[CP]; load_row<CP>(Ds, RB, halo, r, d);
#pragma unroll
                    for (int c = 0; c < CP; ++c) da[c] = dg[c] * d[c];
                }
                if constexpr (ENC == ENC_INSOLE) {
#pragma unroll
                    for (int c = 0; c < CP; ++c) g_b2[c] += da[c];
                } else {
#pragma unroll
                    for (int c = 0; c < CP; ++c) g_b1[c] += da[c];
                }
                store_row<CP>(XHs, RB, halo, r, da);      // dA over XH (row-private)
            }
            __syncthreads();
            if constexpr (ENC == ENC_INSOLE) {
                g_w2.accumulate(HAs, RB, XHs, RB, halo, W, rows, tid);
                for (int r = tid; r < rows; r += NT) {
                    float dh[O1], d1[O1];
                    conv_row<3, C4, O1>(XHs, RB, halo, W, r, w2d, nullptr, dh);
                    load_row<O1>(D1s, RB, halo, r, d1);
#pragma unroll
                    for (int c = 0; c < O1; ++c) { dh[c] *= d1[c]; g_b1[c] += dh[c]; }
                    store_row<O1>(D1s, RB, halo, r, dh);   // dA1 over D1 (row-private)
                }
                __syncthreads();
                g_w1.accumulate(Xs, RBi, D1s, RB, halo, W, rows, tid);
            } else {
                g_w1.accumulate(Xs, RBi, XHs, RB, halo, W, rows, tid);
            }
        }
        __syncthreads();
    }

    // ================= flush per-CTA partial sums (deterministic order)
    float* out = A.partial + (size_t)blockIdx.x * A.NGP;
    if (A.mode == MODE_FWD) return;
    __syncthreads();
    const GradOff& go = A.go;
    {
        const int OUT1 = (ENC == ENC_INSOLE) ? H : C;
        g_w1.flush(stage, out + go.w1, nullptr, CIN, OUT1, tid);
        flush_rowacc<O1>(g_b1, OUT1, stage, out + go.b1, nullptr, tid);
    }
    if constexpr (ENC == ENC_INSOLE) {
        g_w2.flush(stage, out + go.w2, (A.skip_identity ? nullptr : out + go.wsk), H, C, tid);
        flush_rowacc<CP>(g_b2, C, stage, out + go.b2, (A.skip_identity ? nullptr : out + go.bsk), tid);
    }
    if constexpr (ENC != ENC_CONV_POOL) {
        flush_rowacc<CP>(g_lng, C, stage, out + go.lng, nullptr, tid);
        flush_rowacc<CP>(g_lnb, C, stage, out + go.lnb, nullptr, tid);
    }
    g_wb.flush(stage, out + go.wbb, nullptr, C, S, tid);
    flush_rowacc<S>(g_bb, S, stage, out + go.bbb, nullptr, tid);
    // head accumulators: per warp (window slot), per lane (feature); sum the warps in order
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

}  // namespace gaitk
