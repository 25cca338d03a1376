// stream_kernel.cuh -- one gait stream (encoder -> shared backbone -> head -> loss) forward AND
// backward in a single persistent kernel, fp32 FFMA arithmetic ("GAITK_DTYPE_F32" path, parity 1e-5).
// Layout, tiling and phase structure: see stream_common.cuh and DESIGN.md section 3.1.
#pragma once
#include <cooperative_groups.h>

#include "stream_common.cuh"
#include "umma.cuh"

namespace gaitk {

// (d0, d1) += a * (b0, b1) as ONE instruction: Blackwell's packed fp32 FMA (PTX fma.rn.f32x2, SASS FFMA2 with a scalar-broadcast
// operand).  Each lane is an IEEE fma.rn, so results are bit-identical to two fmaf; ptxas never forms FFMA2 on its own.  The
// FMA pipe does the same work per clock either way -- what halves is the number of issue slots, and these kernels are issue-bound.
__device__ __forceinline__ void fma2(float& d0, float& d1, float a, float b0, float b1) {
    unsigned long long d, aa, bb;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a), "f"(a));
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b0), "f"(b1));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(d) : "l"(aa), "l"(bb));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(d));
}

// out[o] = bias[o] + sum_{tap, ci} in[row + (tap - KT/2) * W][ci] * wf[tap][ci][o]
// `in` is a chunked buffer with RBx rows per chunk; wf/bias live in shared memory.
// CIR / COR: REAL input / output channels (<= the padded CI4 * 4 / CO): the unrolled loops drop the multiply-adds of the
// padding (the weights there are zero); a padded output keeps its bias (zero).  FoG widths (6 -> 8, 21 -> 24) spent a
// quarter of their multiply-adds on padding.
template <int KT, int CI4, int CO, int CIR = CI4 * 4, int COR = CO>
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
            if (c4 * 4 >= CIR) continue;
            const float4 x = xp[c4 * RBx];
            const float4* wp = reinterpret_cast<const float4*>(wf + (size_t)(tap * CI4 * 4 + c4 * 4) * CO);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                if (c4 * 4 + e >= CIR) continue;
                const float xe = f4get(x, e);
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4) {
                    if (o4 * 4 >= COR) continue;
                    const float4 w = wp[e * (CO / 4) + o4];
                    // output pairs; a pair whose second output is padding (odd COR) is computed anyway (zero weights)
                    fma2(acc[o4 * 4 + 0], acc[o4 * 4 + 1], xe, w.x, w.y);
                    if (o4 * 4 + 2 < COR) fma2(acc[o4 * 4 + 2], acc[o4 * 4 + 3], xe, w.z, w.w);
                }
            }
        }
    }
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
                fma2(acc[0], acc[1], x.x, d.x, d.y);   fma2(acc[2], acc[3], x.x, d.z, d.w);
                fma2(acc[4], acc[5], x.y, d.x, d.y);   fma2(acc[6], acc[7], x.y, d.z, d.w);
                fma2(acc[8], acc[9], x.z, d.x, d.y);   fma2(acc[10], acc[11], x.z, d.z, d.w);
                fma2(acc[12], acc[13], x.w, d.x, d.y); fma2(acc[14], acc[15], x.w, d.z, d.w);
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

template <class Cfg>
__global__ void __launch_bounds__(NT, Cfg::MINB) stream_kernel(const StreamArgs A, const SmemPlan SP) {
    extern __shared__ __align__(1024) float sm[];      // one alignment for every kernel of the translation unit
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    constexpr int ENC = Cfg::ENC, CIN = Cfg::CIN, CI4 = Cfg::CI4, KT1 = Cfg::KT1, H = Cfg::H, H4 = Cfg::H4;
    constexpr int C = Cfg::C, C4 = Cfg::C4, CP = Cfg::CP, S = Cfg::S, S4 = Cfg::S4, NFL = Cfg::NFL, O1 = Cfg::O1;
    constexpr int PROJ = Cfg::PROJ, CB = Cfg::CB, CB4 = Cfg::CB4, CBP = Cfg::CBP;
    static_assert(PROJ == 0 || (ENC != ENC_CONV_POOL && ENC != ENC_LINEAR_LN_RELU && ENC != ENC_POOL_LINEAR), "projection stage: WearGait encoders");
    constexpr bool POOLLIN = ENC == ENC_POOL_LINEAR;                 // pooled-taps sensor encoder (see EncKind)
    constexpr bool CONVP = ENC == ENC_CONV_POOL || POOLLIN;          // conv (or its pooled-taps form), no activation, no LayerNorm
    constexpr int RC = POOLLIN ? CIN / 3 : CIN;                      // channels of a raw input frame
    static_assert(!POOLLIN || (CIN % 3 == 0 && KT1 == 1), "pooled taps: CIN = 3 x raw channels, one tap");
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
    float* Ls = sm + SP.L; float* DLs = sm + SP.DL; float* wpf = sm + SP.WPF; float* bps = sm + SP.BP; float* wpd = sm + SP.WPD;
    float* BBin = PROJ ? Ls : (ENC == ENC_NONE ? Xs : Fs);   // what the backbone convolves
    const bool enc_only = A.feat_out != nullptr || A.dfeat_in != nullptr;      // encoder stage (no backbone, no head)
    // ---- long windows (T = cl * 128, W = 1): a thread-block cluster of `cl` CTAs shares one window, CTA k owns frames
    // [k rows, (k + 1) rows).  After every layer whose output feeds a k-tap convolution the CTAs copy their neighbours' boundary
    // rows into their own halo rows through distributed shared memory (the outer halos of the first / last CTA stay zero: the
    // reference pads every layer with zeros at the window edges); pooling bins never straddle CTAs (host check), each CTA pools
    // its own bins and writes them into every CTA's feature vector, all CTAs run the head, CTA 0 alone keeps its sums.
    namespace cg = cooperative_groups;
    const int CLn = A.cl > 1 ? A.cl : 1;
    const int crank = CLn > 1 ? (int)cg::this_cluster().block_rank() : 0;
    const int t_off = crank * rows;                       // first frame of this CTA (W == 1 when CLn > 1)
    auto xchg = [&](float* buf, int nch, int RBx) {      // buf: [chunk][RBx rows][4]; own rows written and block-synchronised
        if (CLn == 1) return;
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        for (int e = tid; e < 2 * nch * halo; e += NT) {
            const int side = e / (nch * halo), rem = e - side * nch * halo, ch = rem / halo, h = rem - ch * halo;
            const int peer = side == 0 ? crank - 1 : crank + 1;
            if (peer < 0 || peer >= CLn) continue;
            const float4* src = reinterpret_cast<const float4*>(cluster.map_shared_rank(buf, peer)) + (size_t)ch * RBx + halo + (side == 0 ? rows - halo + h : h);
            float4* dst = reinterpret_cast<float4*>(buf) + (size_t)ch * RBx + (side == 0 ? h : halo + rows + h);
            *dst = *src;
        }
        __syncthreads();
    };

    // ---- one-time setup: zero everything (halos, padded channels), stage weights, bin tables
    for (int i = tid; i < SP.total; i += NT) sm[i] = 0.f;
    __syncthreads();
    // first conv / linear: PyTorch (O, CIN, KT1) -> wf[tap][ci][o]
    if constexpr (ENC != ENC_NONE) {
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
    if constexpr (!CONVP && ENC != ENC_NONE) {
        for (int i = tid; i < C; i += NT) { lngs[i] = A.lng[i]; lnbs[i] = A.lnb[i]; }
    }
    if (!enc_only) for (int i = tid; i < S * CB * 3; i += NT) {
        const int o = i / (CB * 3), ci = (i / 3) % CB, tap = i % 3;
        const float w = A.wbb[i];
        wbf[(tap * CBP + ci) * S + o] = w;
        wbd[((2 - tap) * S + o) * CBP + ci] = w;
    }
    if constexpr (PROJ > 0) {
        // Linear (PROJ, C): forward layout [ci][o] (a 1-tap conv), backward keeps PyTorch's [o][ci]
        for (int i = tid; i < PROJ * C; i += NT) {
            const int o = i / C, ci = i - o * C;
            wpf[ci * CBP + o] = A.wp[i];
            wpd[o * CP + ci] = A.wp[i];
        }
        for (int i = tid; i < PROJ; i += NT) bps[i] = A.bp[i];
    }
    if (!enc_only) for (int i = tid; i < S; i += NT) bbs[i] = A.bbb[i];
    if (A.hw) for (int i = tid; i < K * NF; i += NT) hws[i] = A.hw[i];
    if (A.hb) for (int i = tid; i < K; i += NT) hbs[i] = A.hb[i];
    if (A.head_norm) for (int i = tid; i < NF; i += NT) { hngs[i] = A.hng[i]; hnbs[i] = A.hnb[i]; }
    // adaptive pooling tables: bin b covers [floor(b*T/bdim), ceil((b+1)*T/bdim))
    int* bin_s = bins; int* bin_e = bins + bdim; int* t_lo = bins + 2 * bdim; int* t_hi = t_lo + T;
    for (int b = tid; b < bdim; b += NT) { bin_s[b] = (b * T) / bdim; bin_e[b] = ((b + 1) * T + bdim - 1) / bdim; }
    int* ep_s = t_hi + T; int* ep_e = ep_s + T; int* ep_lo = ep_e + T; int* ep_hi = ep_lo + T_in;   // encoder pool (T_in -> T)
    if constexpr (CONVP) {
        if (A.pool_sensor || POOLLIN)
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
    Wgrad<3, CB4, S4> g_wb;
    Wgrad<1, (PROJ ? C4 : 1), (PROJ ? CB4 : 1)> g_wp;
    float g_bp[PROJ ? CBP : 4];
    float g_b1[O1], g_b2[ENC == ENC_INSOLE ? CP : 4], g_lng[CP], g_lnb[CP], g_bb[S];
    HeadState<NFL, S> head; head.zero();
    HeadCtx hc; hc.Zs = Zs; hc.RB = RB; hc.halo = halo; hc.W = W; hc.S = S; hc.bin_s = bins; hc.bin_e = bins + bdim;
    hc.hws = hws; hc.hbs = hbs; hc.hngs = hngs; hc.hnbs = hnbs; hc.inws = inws; hc.DPs = DPs; hc.Ps = nullptr; hc.ys = nullptr; hc.inv_bin = 0.f;
    if (train) {
        g_w1.zero(); g_w2.zero(); g_wb.zero(); g_wp.zero();
#pragma unroll
        for (int i = 0; i < (PROJ ? CBP : 4); ++i) g_bp[i] = 0.f;
#pragma unroll
        for (int i = 0; i < O1; ++i) g_b1[i] = 0.f;
#pragma unroll
        for (int i = 0; i < (ENC == ENC_INSOLE ? CP : 4); ++i) g_b2[i] = 0.f;
#pragma unroll
        for (int i = 0; i < CP; ++i) { g_lng[i] = 0.f; g_lnb[i] = 0.f; }
#pragma unroll
        for (int i = 0; i < S; ++i) g_bb[i] = 0.f;
    }
    const float inv_denom = (A.mode == MODE_FUSED) ? 1.0f / A.denom[0] : 0.f;

    const int ntiles = (A.B + W - 1) / W;
    // ---- bulk prefetch (host: SP.STGN > 0 -- the FoG / FBG streams): while a tile is computed the raw bytes of the CTA's NEXT
    // tile travel global -> staging by cp.async.bulk (the 16-byte aligned body of every window; a clip of 101 x 21 floats
    // starts on a 4-byte boundary only) plus up to 3 + 3 four-byte cp.async copies for the unaligned head / tail; nothing is
    // read outside the window.  The staged window keeps the source's misalignment so that both ends of the bulk copy are
    // 16-byte aligned.  The synchronous global -> shared scatter this replaces was a third of all stall samples.
    float* STGs = sm + SP.STG;
    uint64_t* ldbar = reinterpret_cast<uint64_t*>(sm + SP.MBAR);
    const int per_raw = T_in * RC;                                   // floats of one raw window
    const bool PF = SP.STGN > 0 && !A.zero_input && CLn == 1;
    uint32_t ldphase = 0;
    auto win_src = [&](int wi) { return A.x + (A.win_start ? (size_t)A.win_start[wi] * RC : (size_t)wi * per_raw); };
    auto prefetch = [&](int tile_) {
        if (tid == 0) {
            uint32_t bytes = 0;
            for (int w = 0; w < W; ++w) {
                const int wi = tile_ * W + w;
                if (wi >= A.B) continue;
                const int mis = (int)((reinterpret_cast<uintptr_t>(win_src(wi)) >> 2) & 3), head = (4 - mis) & 3;
                bytes += (uint32_t)((per_raw - head) & ~3) * 4u;
            }
            umma::mbar_expect_tx(ldbar, bytes);
        }
        for (int w = 0; w < W; ++w) {
            const int wi = tile_ * W + w;
            if (wi >= A.B) continue;
            const float* src = win_src(wi);
            const int mis = (int)((reinterpret_cast<uintptr_t>(src) >> 2) & 3), head = (4 - mis) & 3;
            const int body = (per_raw - head) & ~3, tail = per_raw - head - body;
            float* dst = STGs + w * SP.STGN + mis;
            if (tid == 0) { if (body > 0) umma::bulk_g2s(dst + head, src + head, (uint32_t)body * 4u, ldbar); }
            else if (tid <= head + tail) {
                const int e = tid - 1 < head ? tid - 1 : head + body + (tid - 1 - head);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(umma::smem_u32(dst + e)), "l"(src + e) : "memory");
            }
        }
    };
    if (PF) {
        if (tid == 0) { umma::mbar_init(ldbar, 1); umma::fence_mbar_init(); }
        __syncthreads();
        if ((int)blockIdx.x < ntiles) prefetch(blockIdx.x);
    }
    for (int tile = blockIdx.x / CLn; tile < ntiles; tile += gridDim.x / CLn) {
        const int win0 = tile * W;
        // ================= load: global (window-major) -> Xs [chunk][row][4]
        if (PF) {
            asm volatile("cp.async.wait_all;" ::: "memory");
            umma::mbar_wait(ldbar, ldphase); ldphase ^= 1u;
            __syncthreads();                                          // the 4-byte copies of the other threads
            if constexpr (POOLLIN) {
                for (int w = 0; w < W; ++w) {
                    const int wi = win0 + w;
                    const bool live = wi < A.B;
                    const float* sw = STGs + w * SP.STGN + (live ? (int)((reinterpret_cast<uintptr_t>(win_src(wi)) >> 2) & 3) : 0);
                    // row i, channel ci * 3 + tap = mean over bin i of x[t + tap - 1][ci] (zero outside the clip)
                    for (int it = tid; it < T * RC; it += NT) {
                        const int i = it / RC, ci = it - i * RC;
                        const int s0 = ep_s[i], s1 = ep_e[i];
                        float a0 = 0.f, a1 = 0.f, a2 = 0.f;
                        if (live) {
                            // bin [s0, s1): tap 0 sums frames [s0 - 1, s1 - 1), tap 1 [s0, s1), tap 2 [s0 + 1, s1 + 1) -- the three
                            // windows share their middle frames (s0, s1 - 1), summed once
                            const float* col = sw + ci;
                            const float xm = s0 > 0 ? col[(s0 - 1) * RC] : 0.f, xe = s1 < T_in ? col[s1 * RC] : 0.f;
                            const float x0 = col[s0 * RC];
                            if (s1 - s0 == 1) { a0 = xm; a1 = x0; a2 = xe; }
                            else {
                                const float xl = col[(s1 - 1) * RC];
                                float mid = 0.f;
                                for (int t = s0 + 1; t < s1 - 1; ++t) mid += col[t * RC];
                                const float inv = 1.0f / (float)(s1 - s0);
                                a0 = (xm + x0 + mid) * inv; a1 = (x0 + mid + xl) * inv; a2 = (mid + xl + xe) * inv;
                            }
                        }
                        const int c0 = ci * 3, row = halo + i * W + w;
                        Xs[(((c0) >> 2) * RBi + row) * 4 + ((c0) & 3)] = a0;
                        Xs[(((c0 + 1) >> 2) * RBi + row) * 4 + ((c0 + 1) & 3)] = a1;
                        Xs[(((c0 + 2) >> 2) * RBi + row) * 4 + ((c0 + 2) & 3)] = a2;
                    }
                }
            } else {
                // thread = row (frame t of window w): its CIN staged floats (stride CIN across the lanes: conflict-free for odd CIN)
                // -> one 16-byte word per channel chunk.  (A loop over the window's elements with a division per element was 18 %
                // of the skeleton kernel's instructions.)
                for (int r = tid; r < rows_in; r += NT) {
                    const int t = r / W, w = r - t * W, wi = win0 + w;
                    const bool live = wi < A.B;
                    const float* src = STGs + w * SP.STGN + (live ? (int)((reinterpret_cast<uintptr_t>(win_src(wi)) >> 2) & 3) : 0) + t * CIN;
                    float v[CI4 * 4];
#pragma unroll
                    for (int c = 0; c < CI4 * 4; ++c) v[c] = (c < CIN && live) ? src[c] : 0.f;
                    store_row<CI4 * 4>(Xs, RBi, halo, r, v);
                }
            }
            umma::fence_smem_to_async();                              // staging is read; the next bulk copy may overwrite it
            __syncthreads();
            const int nxt = tile + (int)gridDim.x;
            if (nxt < ntiles) prefetch(nxt);
        } else if constexpr (POOLLIN) {
            // masked / zero input: the pooled taps of zeros (a live input always takes the staged path: the host plans
            // the staging buffer with this encoder)
            for (int e = tid; e < CI4 * rows; e += NT) {
                const int c4 = e / rows, r = e - c4 * rows;
                reinterpret_cast<float4*>(Xs)[c4 * RBi + halo + r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
        } else if (CLn > 1) {
            // frames [t_off - halo, t_off + rows + halo) of the one window; zero outside [0, T)
            const int wi = win0;
            const size_t base = (wi < A.B) ? (A.win_start ? (size_t)A.win_start[wi] * CIN : (size_t)wi * T_in * CIN) : 0;
            for (int e = tid; e < (rows + 2 * halo) * CIN; e += NT) {
                const int rr = e / CIN, c = e - rr * CIN, t = t_off + rr - halo;
                float v = 0.f;
                if (t >= 0 && t < T && wi < A.B && !A.zero_input) v = __ldg(A.x + base + (size_t)t * CIN + c);
                Xs[((c >> 2) * RBi + rr) * 4 + (c & 3)] = v;
            }
        } else {
            const int per_win = T_in * CIN;
            for (int e = tid; e < W * per_win; e += NT) {
                const int w = (e >= per_win) + (e >= 2 * per_win) + (e >= 3 * per_win), rem = e - w * per_win;   // W <= 4
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
                conv_row<KT1, CI4, O1, CIN, H>(Xs, RBi, halo, W, r, w1f, b1s, a1);
#pragma unroll
                for (int c = 0; c < O1; ++c) { if (c < H) gelu_fwd(a1[c], ha[c], d1[c]); else { ha[c] = 0.f; d1[c] = 0.f; } }
                store_row<O1>(HAs, RB, halo, r, ha);
                if (train) store_row<O1>(D1s, RB, halo, r, d1);
            }
            __syncthreads();
            xchg(HAs, H4, RB);
        }
        if constexpr (ENC == ENC_NONE) {
            // nothing: Xs is the backbone input
        } else if constexpr (CONVP) {
            // conv over the T_in input rows, no activation; optional adaptive pooling T_in -> T
            for (int r = tid; r < rows_in; r += NT) {
                float a[CP];
                conv_row<KT1, CI4, CP, CIN, C>(Xs, RBi, halo, W, r, w1f, b1s, a);
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
                if constexpr (ENC == ENC_INSOLE) conv_row<3, H4, CP, H, C>(HAs, RB, halo, W, r, w2f, b2s, a);
                else conv_row<KT1, CI4, CP, CIN, C>(Xs, RBi, halo, W, r, w1f, b1s, a);
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
                if constexpr (PROJ > 0) {
                    float l[CBP];
#pragma unroll
                    for (int j = 0; j < CBP; ++j) l[j] = j < PROJ ? bps[j] : 0.f;
#pragma unroll
                    for (int cc = 0; cc < C; ++cc)
#pragma unroll
                        for (int j = 0; j < CBP; ++j) l[j] = fmaf(f[cc], wpf[cc * CBP + j], l[j]);
                    store_row<CBP>(Ls, RB, halo, r, l);
                }
            }
            __syncthreads();
            xchg(BBin, CB4, RB);
        }
        if (A.feat_out) {
            // ---- encoder stage, forward: rows of F -> (B, T, C) and on to the next tile
            for (int r = tid; r < rows; r += NT) {
                const int t = r / W, w = r - t * W, wi = win0 + w;
                if (wi >= A.B) continue;
                float f[CP]; load_row<CP>(Fs, RB, halo, r, f);
                float* dst = A.feat_out + ((size_t)wi * T + t) * C;
#pragma unroll
                for (int c = 0; c < CP; ++c) if (c < C) dst[c] = f[c];
            }
            __syncthreads();
            continue;
        }
        // ================= shared backbone forward: conv3 -> ReLU
        if (!A.dfeat_in) {
        for (int r = tid; r < rows; r += NT) {
            float z[S];
            conv_row<3, CB4, S, CB, S>(BBin, RB, halo, W, r, wbf, bbs, z);
#pragma unroll
            for (int s = 0; s < S; ++s) z[s] = fmaxf(z[s], 0.f);
            store_row<S>(Zs, RB, halo, r, z);
        }
        __syncthreads();
        // ================= adaptive pool + head + loss: warp w <-> window w of the tile
        if (CLn > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (wrp == 0) {
#pragma unroll
                for (int i = 0; i < NFL; ++i) {
                    const int j = lane + 32 * i, b = j / S, sch = j - b * S;
                    const int t0 = bin_s[b], t1 = bin_e[b];
                    if (t0 < t_off || t1 > t_off + rows) continue;          // another CTA's bin
                    float acc = 0.f;
                    for (int t = t0; t < t1; ++t) acc += Zs[((sch >> 2) * RB + halo + (t - t_off)) * 4 + (sch & 3)];
                    const float f = acc / (float)(t1 - t0);
                    for (int peer = 0; peer < CLn; ++peer) cluster.map_shared_rank(Ps, peer)[j] = f;
                }
            }
            cluster.sync();
            hc.Ps = Ps;
            if (wrp == 0) {
                head.run(A, hc, 0, lane, win0, train, inv_denom);
                if (crank != 0) head.zero();                               // CTA 0 alone accounts for the head, loss and accuracy
            }
        } else if (W == 1) {
            // one window per tile (FoG / FBG clips): all four warps pool (same sums, same order as the head's own pooling),
            // warp 0 runs head + loss -- instead of three warps waiting while one pools and runs the head
            for (int j = tid; j < NF; j += NT) {
                const int b = j / S, sch = j - b * S;
                const int t0 = bin_s[b], t1 = bin_e[b];
                float acc = 0.f;
                for (int t = t0; t < t1; ++t) acc += Zs[((sch >> 2) * RB + halo + t) * 4 + (sch & 3)];
                Ps[j] = acc / (float)(t1 - t0);
            }
            __syncthreads();
            hc.Ps = Ps;
            if (wrp == 0) head.run(A, hc, 0, lane, win0, train, inv_denom);
        } else
        if (wrp < W) head.run(A, hc, wrp, lane, win0, train, inv_denom);
        if (!train) { __syncthreads(); continue; }
        __syncthreads();
        // ================= backbone backward: dz (through pool + ReLU), in place over Z
        for (int r = tid; r < rows; r += NT) {
            const int t = r / W, w = r - t * W;
            float z[S], dz[S];
            load_row<S>(Zs, RB, halo, r, z);
            const int lo = t_lo[t + t_off], hi = t_hi[t + t_off];
            // a frame lies in one pooling bin, or in two where ATen's adaptive bins overlap: vector rows of DP instead of a
            // scalar loop per channel (that loop was 11 - 16 % of the FoG kernels' instructions)
            float dsum[S];
#pragma unroll
            for (int s = 0; s < S; ++s) dsum[s] = 0.f;
            for (int b = lo; b <= hi; ++b) {
                const float4* dp4 = reinterpret_cast<const float4*>(DPs + w * NF + b * S);
#pragma unroll
                for (int s4 = 0; s4 < S4; ++s4) {
                    const float4 v = dp4[s4];
                    dsum[s4 * 4] += v.x; dsum[s4 * 4 + 1] += v.y; dsum[s4 * 4 + 2] += v.z; dsum[s4 * 4 + 3] += v.w;
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) {
                dz[s] = z[s] > 0.f ? dsum[s] : 0.f;
                g_bb[s] += dz[s];
            }
            store_row<S>(Zs, RB, halo, r, dz);
        }
        __syncthreads();
        xchg(Zs, S4, RB);
        g_wb.accumulate(BBin, RB, Zs, RB, halo, W, rows, tid);
        }   // !A.dfeat_in
        // gradient of the encoder output: the backbone's data gradient, or (encoder stage) what the fusion op sent back
        auto load_dfeat = [&](int r, float (&df)[CP]) {
            const int t = r / W, w = r - t * W, wi = win0 + w;
            const float* src = A.dfeat_in + ((size_t)wi * T + t) * C;
#pragma unroll
            for (int c = 0; c < CP; ++c) df[c] = (c < C && wi < A.B) ? src[c] : 0.f;
        };
        // dgrad into the encoder output + encoder-specific backward up to the first-conv pre-activation
        if constexpr (ENC == ENC_NONE) {
            if (A.dx) {
                for (int r = tid; r < rows; r += NT) {
                    const int t = r / W, w = r - t * W, wi = win0 + w;
                    if (wi >= A.B) continue;
                    float dxr[CBP];
                    conv_row<3, S4, CBP, S, CB>(Zs, RB, halo, W, r, wbd, nullptr, dxr);
                    float* dst = A.dx + ((size_t)wi * T + t) * CIN;
#pragma unroll
                    for (int c = 0; c < CBP; ++c) if (c < CIN) dst[c] = dxr[c];
                }
            }
        } else if constexpr (CONVP) {
            for (int r = tid; r < rows; r += NT) {
                float df[CP];
                if (A.dfeat_in) load_dfeat(r, df);
                else conv_row<3, S4, CP, S, C>(Zs, RB, halo, W, r, wbd, nullptr, df);
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
                if constexpr (PROJ > 0) {
                    float dl[CBP];
                    conv_row<3, S4, CBP, S, CB>(Zs, RB, halo, W, r, wbd, nullptr, dl);
#pragma unroll
                    for (int j = 0; j < CBP; ++j) g_bp[j] += dl[j];
                    store_row<CBP>(DLs, RB, halo, r, dl);
#pragma unroll
                    for (int cc = 0; cc < CP; ++cc) df[cc] = 0.f;
#pragma unroll
                    for (int j = 0; j < PROJ; ++j)
#pragma unroll
                        for (int cc = 0; cc < C; ++cc) df[cc] = fmaf(dl[j], wpd[j * CP + cc], df[cc]);
                } else {
                    if (A.dfeat_in) load_dfeat(r, df);
                    else conv_row<3, S4, CP, S, C>(Zs, RB, halo, W, r, wbd, nullptr, df);
                }
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
                xchg(XHs, C4, RB);
                g_w2.accumulate(HAs, RB, XHs, RB, halo, W, rows, tid);
                for (int r = tid; r < rows; r += NT) {
                    float dh[O1], d1[O1];
                    conv_row<3, C4, O1, C, H>(XHs, RB, halo, W, r, w2d, nullptr, dh);
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
            if constexpr (PROJ > 0) g_wp.accumulate(Fs, RB, DLs, RB, halo, W, rows, tid);
        }
        __syncthreads();
    }

    if (CLn > 1) cg::this_cluster().sync();                // no CTA exits while a neighbour may still read its shared memory
    // ================= flush per-CTA partial sums (deterministic order)
    float* out = A.partial + (size_t)blockIdx.x * A.NGP;
    if (A.mode == MODE_FWD) return;
    __syncthreads();
    const GradOff& go = A.go;
    if constexpr (ENC != ENC_NONE) {
        const int OUT1 = (ENC == ENC_INSOLE) ? H : C;
        g_w1.flush(stage, out + go.w1, nullptr, CIN, OUT1, tid);
        flush_rowacc<O1>(g_b1, OUT1, stage, out + go.b1, nullptr, tid);
    }
    if constexpr (ENC == ENC_INSOLE) {
        g_w2.flush(stage, out + go.w2, (A.skip_identity ? nullptr : out + go.wsk), H, C, tid);
        flush_rowacc<CP>(g_b2, C, stage, out + go.b2, (A.skip_identity ? nullptr : out + go.bsk), tid);
    }
    if constexpr (!CONVP && ENC != ENC_NONE) {
        flush_rowacc<CP>(g_lng, C, stage, out + go.lng, nullptr, tid);
        flush_rowacc<CP>(g_lnb, C, stage, out + go.lnb, nullptr, tid);
    }
    if (go.wbb >= 0) {                                    // absent in an encoder stage
        g_wb.flush(stage, out + go.wbb, nullptr, CB, S, tid);
        flush_rowacc<S>(g_bb, S, stage, out + go.bbb, nullptr, tid);
    }
    if constexpr (PROJ > 0) {
        g_wp.flush(stage, out + go.wp, nullptr, C, PROJ, tid);
        flush_rowacc<CBP>(g_bp, PROJ, stage, out + go.bp, nullptr, tid);
    }
    if (go.hw >= 0) head.flush(A, stage, out, tid);       // absent in encoder / trunk stages
}

}  // namespace gaitk
