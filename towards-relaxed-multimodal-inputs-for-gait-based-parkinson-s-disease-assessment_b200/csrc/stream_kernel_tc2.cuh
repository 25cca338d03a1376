// stream_kernel_tc2.cuh -- tensor-core stream kernel, TWO threads per row (256-thread CTAs).
//
// Same tile, buffers, MMA descriptors and phase structure as stream_kernel_tc.cuh.  What changes is who does the
// thread-local work between the MMAs.  That work (GELU / LayerNorm / ReLU epilogues and their backward, the
// tf32 conversion of the operands, the mma.sync weight gradients) is latency-bound: with one thread per row a CTA has
// four warps and the shared-memory footprint of a tile (52 - 111 KB) allows only 2-3 CTAs per SM, i.e. 8-12 warps,
// which leaves the schedulers idle most cycles (ncu: IPC 1.3-1.5, profiles/).  Here warps w and w+4 share TMEM lane
// quarter w (tcgen05.ld may address lanes 32 (warp % 4) .. +31 from either) and split the CHANNELS of each row, so the
// same shared memory carries twice the warps and every thread holds half the state (fewer registers):
//   * row r = tid & 127, half h = tid >> 7; a thread owns NCH/2 channels of a NCH-channel row (HalfMap below);
//   * LayerNorm needs the whole row: the two halves exchange (mean, M2) resp. (sum dxh, sum dxh*xh) through a
//     2 KB shared buffer and a 64-thread named barrier per lane quarter (Chan's parallel-variance merge);
//   * the mma.sync weight-gradient tiles keep their owners (warp % 4) and split the K (row) range between the halves;
//     the two partial sums are added in a fixed order at the final flush;
//   * geometry is compile time (T = 64, W = 2, 128 rows, 8 pooling bins): shared-memory addressing is immediate.
#pragma once
#include "stream_kernel_tc.cuh"

namespace gaitk {

constexpr int NT2 = 256;

__host__ __device__ constexpr int round_rb_c(int rows, int halo) {
    int rb = rows + 2 * halo;
    while (rb % 8 != GAITK_RB_MOD) ++rb;
    return rb;
}

template <class Cfg> struct Tc2MinBlocks { static constexpr int value = 2; };

// Which channels of a NCH-channel row (NCH = 12, 16, 24) a thread of half h owns, as float4 groups plus (NCH = 12)
// one float2:  local index i < 4 F4 -> chunk (h ? NCH/4 - F4 : 0) + i / 4;  the remaining two -> the middle chunk's
// elements 2h, 2h + 1.  Both halves run the same code with h-dependent base offsets.
template <int NCH>
struct HalfMap {
    static_assert(NCH % 4 == 0 && (NCH % 8 == 0 || NCH % 8 == 4), "");
    static constexpr int NH = NCH / 2;           // channels per thread
    static constexpr int F4 = NCH / 8;           // float4 groups per thread
    static constexpr int F2 = (NCH % 8) / 4;     // 0 or 1 float2
    static constexpr int HI = NCH / 4 - F4;      // first chunk of half 1
    __device__ __forceinline__ static int ch(int i, int h) {
        return i < 4 * F4 ? ((h ? HI : 0) + i / 4) * 4 + (i & 3) : F4 * 4 + 2 * h + (i - 4 * F4);
    }
    // TMEM accumulator columns -> v
    __device__ __forceinline__ static void tmem(uint32_t trow, int h, float (&v)[NH]) {
#pragma unroll
        for (int g = 0; g < F4; ++g) umma::ld_x4(trow + (uint32_t)(((h ? HI : 0) + g) * 4), &v[4 * g]);
        if constexpr (F2 == 1) umma::ld_x2(trow + (uint32_t)(F4 * 4 + 2 * h), &v[4 * F4]);
        umma::ld_wait();
    }
    template <int RB, bool TF32>
    __device__ __forceinline__ static void store(float* buf, int row, int h, const float (&v)[NH]) {
        float4* p = reinterpret_cast<float4*>(buf) + ((h ? HI : 0) * RB + row);
#pragma unroll
        for (int g = 0; g < F4; ++g) {
            if constexpr (TF32) p[g * RB] = make_float4(umma::to_tf32(v[4 * g]), umma::to_tf32(v[4 * g + 1]), umma::to_tf32(v[4 * g + 2]), umma::to_tf32(v[4 * g + 3]));
            else p[g * RB] = make_float4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
        }
        if constexpr (F2 == 1) {
            float2* q = reinterpret_cast<float2*>(buf + ((size_t)(F4 * RB + row) * 4 + 2 * h));
            if constexpr (TF32) *q = make_float2(umma::to_tf32(v[4 * F4]), umma::to_tf32(v[4 * F4 + 1]));
            else *q = make_float2(v[4 * F4], v[4 * F4 + 1]);
        }
    }
    template <int RB>
    __device__ __forceinline__ static void load(const float* buf, int row, int h, float (&v)[NH]) {
        const float4* p = reinterpret_cast<const float4*>(buf) + ((h ? HI : 0) * RB + row);
#pragma unroll
        for (int g = 0; g < F4; ++g) { const float4 t = p[g * RB]; v[4 * g] = t.x; v[4 * g + 1] = t.y; v[4 * g + 2] = t.z; v[4 * g + 3] = t.w; }
        if constexpr (F2 == 1) {
            const float2 t = *reinterpret_cast<const float2*>(buf + ((size_t)(F4 * RB + row) * 4 + 2 * h));
            v[4 * F4] = t.x; v[4 * F4 + 1] = t.y;
        }
    }
    // per-channel parameter vector in shared memory -> this thread's channels
    __device__ __forceinline__ static void params(const float* src, int h, float (&v)[NH]) {
#pragma unroll
        for (int i = 0; i < NH; ++i) v[i] = src[ch(i, h)];
    }
};

__device__ __forceinline__ void pair_barrier(int wq) {          // warps wq and wq + 4
    asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
}

// per-row-thread accumulators split over the two halves: deterministic block sum, channel c -> dst[c]
template <int NCH>
__device__ __forceinline__ void flush_rowacc2(const float (&v)[NCH / 2], int nreal, float* stage, float* dst, float* dst2, int tid) {
    using M = HalfMap<NCH>;
    __syncthreads();
    const int lane = tid & 31, wq = (tid >> 5) & 3, h = tid >> 7;
#pragma unroll
    for (int i = 0; i < NCH / 2; ++i) {
        const float s = warp_sum(v[i]);
        if (lane == 0) stage[M::ch(i, h) * 4 + wq] = s;
    }
    __syncthreads();
    if (tid < nreal) {
        const float s = (stage[tid * 4] + stage[tid * 4 + 1]) + (stage[tid * 4 + 2] + stage[tid * 4 + 3]);
        dst[tid] = s;
        if (dst2) dst2[tid] = s;
    }
    __syncthreads();
}

// HeadState::flush for a 256-thread CTA: the head runs on warps 0..W-1 (< 4); only the first four warps hold state
template <int NFL, int SC>
__device__ __forceinline__ void head_flush2(HeadState<NFL, SC>& hs, const StreamArgs& A, float* stage, float* out, int tid) {
    const int lane = tid & 31, wrp = tid >> 5, K = A.K, NF = A.NF;
    const GradOff& go = A.go;
    const bool lo = tid < NT;
    __syncthreads();
    if (lo) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) if (k < K)
#pragma unroll
            for (int i = 0; i < NFL; ++i) stage[(wrp * KMAX + k) * NF + lane + 32 * i] = hs.g_hw[k][i];
    }
    __syncthreads();
    for (int e = tid; e < K * NF; e += NT2) {
        const int k = e / NF, j = e - k * NF;
        float s = 0.f;
        for (int w = 0; w < 4; ++w) s += stage[(w * KMAX + k) * NF + j];
        out[go.hw + e] = s;
    }
    __syncthreads();
    if (A.head_norm) {
        if (lo) {
#pragma unroll
            for (int i = 0; i < NFL; ++i) { stage[wrp * 2 * NF + lane + 32 * i] = hs.g_hng[i]; stage[wrp * 2 * NF + NF + lane + 32 * i] = hs.g_hnb[i]; }
        }
        __syncthreads();
        for (int e = tid; e < 2 * NF; e += NT2) {
            float s = 0.f;
            for (int w = 0; w < 4; ++w) s += stage[w * 2 * NF + e];
            if (e < NF) out[go.hng + e] = s; else out[go.hnb + e - NF] = s;
        }
        __syncthreads();
    }
    if (lo && lane == 0) {
#pragma unroll
        for (int k = 0; k < KMAX; ++k) stage[wrp * 8 + k] = hs.g_hb[k];
        stage[wrp * 8 + 4] = hs.acc_loss; stage[wrp * 8 + 5] = hs.acc_correct;
    }
    __syncthreads();
    if (tid < 6) {
        float s = 0.f;
        for (int w = 0; w < 4; ++w) s += stage[w * 8 + tid];
        if (tid < 4) { if (tid < K && go.hb >= 0) out[go.hb + tid] = s; }
        else out[go.total + (tid - 4)] = s;      // [NG] = loss, [NG+1] = correct
    }
}

template <class Cfg>
__global__ void __launch_bounds__(NT2, Tc2MinBlocks<Cfg>::value) stream_kernel_tc2(const StreamArgs A, const TcPlan SP) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar, ldbar;
    __shared__ uint32_t tmem_slot;
    __shared__ int ys_tile[2][WMAX];                      // labels of the current / next tile
    __shared__ uint64_t dtab[5][2 * 12];                  // MMA descriptor pairs: conv1, conv2, bb, bb dgrad, conv2 dgrad
    __shared__ __align__(16) float lnx[NT][4];                          // LayerNorm exchange between the two halves of a row
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5, wq = wrp & 3, h = tid >> 7;
    constexpr int ENC = Cfg::ENC, CIN = Cfg::CIN, CI4 = Cfg::CI4, KT1 = Cfg::KT1, H = Cfg::H, H4 = Cfg::H4;
    constexpr int C = Cfg::C, C4 = Cfg::C4, CP = Cfg::CP, S = Cfg::S, S4 = Cfg::S4, NFL = Cfg::NFL;
    static_assert(ENC == ENC_CONV_GELU_LN || ENC == ENC_INSOLE, "tensor-core kernel: WearGait encoders");
    static_assert(C == CP && C % 4 == 0 && S % 8 == 0 && (H == 0 || H % 8 == 0), "channel split");
    constexpr int KX = (CI4 + 1) / 2 * 2;                 // even chunk counts (K = 8 per MMA)
    constexpr int KH = (H4 + 1) / 2 * 2, KC = (C4 + 1) / 2 * 2, KS = (S4 + 1) / 2 * 2;
    constexpr int N1 = (ENC == ENC_INSOLE) ? ((H + 15) / 16) * 16 : ((C + 15) / 16) * 16;   // first conv outputs
    constexpr int NC = ((C + 15) / 16) * 16, NS = ((S + 15) / 16) * 16, NH = ((H + 15) / 16) * 16;
    constexpr int O1 = (ENC == ENC_INSOLE) ? H4 * 4 : CP;
    static_assert(N1 <= 32 && NC <= 32 && NS <= 32 && (H == 0 || NH <= 32), "accumulator fits 32 TMEM columns");
    // compile-time tile geometry (the host selects this kernel only for T = 64, W = 2, bdim = 8)
    constexpr int W = 2, T = 64, rows = 128, logW = 1, bdim = 8, binsz = T / bdim;
    constexpr int halo = (KT1 / 2 > 1 ? KT1 / 2 : 1) * W;
    constexpr int RB = round_rb_c(rows, halo);
    using MC = HalfMap<C>;                                // encoder output row (12 channels: 6 + 6)
    using MS = HalfMap<S>;                                // backbone row (16: 8 + 8)
    using MH = HalfMap<(ENC == ENC_INSOLE ? O1 : 8)>;     // insole hidden row (24: 12 + 12)
    const int K = A.K, NF = A.NF;
    const bool train = A.mode != MODE_FWD;

    float* Xs = sm + SP.X; float* HAs = sm + SP.HA; float* D1s = sm + SP.D1; float* XHs = sm + SP.XH;
    float* Ds = sm + SP.D; float* Fs = sm + SP.F; float* Zs = sm + SP.Z;
    float* w1b = sm + SP.W1B; float* b1s = sm + SP.B1; float* w2b = sm + SP.W2B; float* b2s = sm + SP.B2; float* w2d = sm + SP.W2D;
    float* lngs = sm + SP.LNG; float* lnbs = sm + SP.LNB; float* wbb = sm + SP.WBB; float* bbs = sm + SP.BB; float* wbd = sm + SP.WBD;
    float* hws = sm + SP.HW; float* hbs = sm + SP.HB; float* hngs = sm + SP.HNG; float* hnbs = sm + SP.HNB; float* inws = sm + SP.INW;
    float* DPs = sm + SP.DP; int* bins = reinterpret_cast<int*>(sm + SP.BINS); float* stage = sm + SP.STAGE;
    float* STGs = sm + SP.STG;                            // raw window bytes of the NEXT tile (TMA bulk prefetch)
    float* Ps = sm + SP.P;

    // ---- one-time setup
    for (int i = tid; i < SP.total; i += NT2) sm[i] = 0.f;
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_init(&ldbar, 1); umma::fence_mbar_init(); }
    if (wrp == 0) umma::tmem_alloc(&tmem_slot, 32);
    __syncthreads();
    {
        const int OUT1 = (ENC == ENC_INSOLE) ? H : C;
        for (int i = tid; i < OUT1 * CIN * KT1; i += NT2) {
            const int o = i / (CIN * KT1), ci = (i / KT1) % CIN, tap = i % KT1;
            w1b[((tap * KX + (ci >> 2)) * N1 + o) * 4 + (ci & 3)] = umma::to_tf32(A.w1[i]);
        }
        for (int i = tid; i < OUT1; i += NT2) b1s[i] = A.b1[i];
    }
    if constexpr (ENC == ENC_INSOLE) {
        for (int i = tid; i < C * H * 3; i += NT2) {
            const int o = i / (H * 3), ci = (i / 3) % H, tap = i % 3;
            float w = A.w2[i];
            if (tap == 1) w += A.skip_identity ? (o == ci ? 1.f : 0.f) : A.wsk[o * H + ci];
            w = umma::to_tf32(w);
            w2b[((tap * KH + (ci >> 2)) * NC + o) * 4 + (ci & 3)] = w;                    // fwd: n = o, k = ci
            w2d[(((2 - tap) * KC + (o >> 2)) * NH + ci) * 4 + (o & 3)] = w;               // dgrad: n = ci, k = o, flipped taps
        }
        for (int i = tid; i < C; i += NT2) b2s[i] = A.b2[i] + (A.skip_identity ? 0.f : A.bsk[i]);
    }
    for (int i = tid; i < C; i += NT2) { lngs[i] = A.lng[i]; lnbs[i] = A.lnb[i]; }
    for (int i = tid; i < S * C * 3; i += NT2) {
        const int o = i / (C * 3), ci = (i / 3) % C, tap = i % 3;
        const float w = umma::to_tf32(A.wbb[i]);
        wbb[((tap * KC + (ci >> 2)) * NS + o) * 4 + (ci & 3)] = w;
        wbd[(((2 - tap) * KS + (o >> 2)) * NC + ci) * 4 + (o & 3)] = w;
    }
    for (int i = tid; i < S; i += NT2) bbs[i] = A.bbb[i];
    for (int i = tid; i < K * NF; i += NT2) hws[i] = A.hw[i];
    if (A.hb) for (int i = tid; i < K; i += NT2) hbs[i] = A.hb[i];
    if (A.head_norm) for (int i = tid; i < NF; i += NT2) { hngs[i] = A.hng[i]; hnbs[i] = A.hnb[i]; }
    int* bin_s = bins; int* bin_e = bins + bdim;
    for (int b = tid; b < bdim; b += NT2) { bin_s[b] = b * binsz; bin_e[b] = (b + 1) * binsz; }
    if (A.head_cos && wrp < K) {
        float s = 0.f;
        for (int j = lane; j < NF; j += 32) s = fmaf(hws[wrp * NF + j], hws[wrp * NF + j], s);
        s = warp_sum(s);
        if (lane == 0) inws[wrp] = 1.0f / fmaxf(sqrtf(s), 1e-8f);
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16);           // this warp's 32 TMEM lanes
    uint32_t phase = 0;
    const uint32_t sX = umma::smem_u32(Xs), sHA = umma::smem_u32(HAs), sXH = umma::smem_u32(XHs), sF = umma::smem_u32(Fs),
                   sZ = umma::smem_u32(Zs), sW1 = umma::smem_u32(w1b), sW2 = umma::smem_u32(w2b), sW2D = umma::smem_u32(w2d),
                   sWB = umma::smem_u32(wbb), sWBD = umma::smem_u32(wbd);

    static_assert(KT1 * (KX / 2) <= 12 && 3 * (KH / 2) <= 12 && 3 * (KC / 2) <= 12 && 3 * (KS / 2) <= 12, "descriptor table size");
    if (tid == 0) {
        build_conv_descs<KT1, KX, N1>(dtab[0], sX, RB, halo, W, sW1);
        if constexpr (ENC == ENC_INSOLE) {
            build_conv_descs<3, KH, NC>(dtab[1], sHA, RB, halo, W, sW2);
            build_conv_descs<3, KC, NH>(dtab[4], sXH, RB, halo, W, sW2D);
        }
        build_conv_descs<3, KC, NS>(dtab[2], sF, RB, halo, W, sWB);
        build_conv_descs<3, KS, NC>(dtab[3], sZ, RB, halo, W, sWBD);
    }
    __syncthreads();

    // ---- persistent accumulators (each thread: its half of the channels / its half of the rows)
    WgradMma<KT1, CI4 * 4, (O1 + 7) / 8 * 8> g_w1;
    WgradMma<3, (ENC == ENC_INSOLE ? H4 * 4 : 4), (ENC == ENC_INSOLE ? NC : 8)> g_w2;
    WgradMma<3, CP, NS> g_wb;
    float g_b1[(ENC == ENC_INSOLE) ? MH::NH : MC::NH], g_b2[MC::NH], g_lng[MC::NH], g_lnb[MC::NH], g_bb[MS::NH];
    HeadState<NFL, S> head; head.zero();
    HeadCtx hc; hc.Zs = Zs; hc.RB = RB; hc.halo = halo; hc.W = W; hc.S = S; hc.bin_s = bin_s; hc.bin_e = bin_e;
    hc.hws = hws; hc.hbs = hbs; hc.hngs = hngs; hc.hnbs = hnbs; hc.inws = inws; hc.DPs = DPs;
    hc.Ps = Ps; hc.ys = nullptr; hc.inv_bin = 0.f;
    g_w1.zero(); g_w2.zero(); g_wb.zero();
#pragma unroll
    for (int i = 0; i < ((ENC == ENC_INSOLE) ? MH::NH : MC::NH); ++i) g_b1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < MC::NH; ++i) { g_b2[i] = 0.f; g_lng[i] = 0.f; g_lnb[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < MS::NH; ++i) g_bb[i] = 0.f;
    const float inv_denom = (A.mode == MODE_FUSED) ? 1.0f / A.denom[0] : 0.f;
    // this thread's per-channel parameters
    float bias_c[MC::NH], lng_c[MC::NH], lnb_c[MC::NH], bb_c[MS::NH];
    MC::params((ENC == ENC_INSOLE) ? b2s : b1s, h, bias_c); MC::params(lngs, h, lng_c); MC::params(lnbs, h, lnb_c);
    MS::params(bbs, h, bb_c);

#define GAITK_MMA_PHASE(ISSUE)                                              \
    do {                                                                    \
        umma::fence_smem_to_async(); umma::fence_before_sync();             \
        __syncthreads();                                                    \
        if (tid == 0) { umma::fence_after_sync(); ISSUE; umma::commit(&bar); } \
    } while (0)
#define GAITK_MMA_ISSUE_SYNCED(ISSUE)                                       \
    do { if (tid == 0) { umma::fence_after_sync(); ISSUE; umma::commit(&bar); } } while (0)
#define GAITK_MMA_WAIT()                                                    \
    do { umma::mbar_wait(&bar, phase); phase ^= 1u; umma::fence_after_sync(); } while (0)

    const int r = tid & (NT - 1);                         // row of the tile
    const int row = halo + r;                             // row inside a chunk plane
    const int k0 = h * (rows / 2), k1 = k0 + rows / 2;    // this half's K (row) range of the weight gradients
    const int ntiles = (A.B + W - 1) / W;
    constexpr int per_win = T * CIN;
    uint32_t ldphase = 0;
    int ylab_next = 0;
    constexpr int PF_WARP = 7;
    auto prefetch = [&](int tile) {
        if (wrp != PF_WARP) return;
        if (A.mode == MODE_FUSED && lane < W) { const int wi = tile * W + lane; ylab_next = wi < A.B ? (int)A.y[wi] : 0; }   // consumed later
        if (A.zero_input) return;
        uint32_t bytes = 0;
        for (int w = 0; w < W; ++w) {
            const int wi = tile * W + w;
            if (wi >= A.B) continue;
            const float* src = A.x + (A.win_start ? (size_t)A.win_start[wi] * CIN : (size_t)wi * per_win);
            if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) bytes += (uint32_t)per_win * 4u;
        }
        if (lane == 0) umma::mbar_expect_tx(&ldbar, bytes);
        __syncwarp();
        for (int w = 0; w < W; ++w) {
            const int wi = tile * W + w;
            if (wi >= A.B) continue;
            const float* src = A.x + (A.win_start ? (size_t)A.win_start[wi] * CIN : (size_t)wi * per_win);
            float* dst = STGs + (size_t)w * per_win;
            if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) {
                if (lane == 0) umma::bulk_g2s(dst, src, (uint32_t)per_win * 4u, &ldbar);
            } else {
                for (int e = lane; e < per_win; e += 32) dst[e] = __ldg(src + e);
            }
        }
    };
    int yslot = 0;
    if (blockIdx.x < ntiles) { prefetch(blockIdx.x); if (wrp == PF_WARP && lane < W) ys_tile[0][lane] = ylab_next; }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int win0 = tile * W;
        // ================= staged window bytes -> Xs [chunk][row][4] (tf32)
        if (!A.zero_input) {
            umma::mbar_wait(&ldbar, ldphase); ldphase ^= 1u;
            if constexpr (CIN % 4 == 0) {
                constexpr int pw4 = per_win / 4;
                for (int e = tid; e < W * pw4; e += NT2) {
                    const int w = e >= pw4, rem = e - w * pw4;
                    const int t = rem / (CIN / 4), c4 = rem - t * (CIN / 4);
                    float4 v = reinterpret_cast<const float4*>(STGs)[e];
                    if (win0 + w >= A.B) v = make_float4(0.f, 0.f, 0.f, 0.f);
                    reinterpret_cast<float4*>(Xs)[c4 * RB + halo + (t << logW) + w] =
                        make_float4(umma::to_tf32(v.x), umma::to_tf32(v.y), umma::to_tf32(v.z), umma::to_tf32(v.w));
                }
            } else {
                for (int e = tid; e < W * per_win; e += NT2) {
                    const int w = e >= per_win, rem = e - w * per_win;
                    const int t = rem / CIN, c = rem - t * CIN;
                    const float v = (win0 + w < A.B) ? STGs[e] : 0.f;
                    Xs[((c >> 2) * RB + halo + (t << logW) + w) * 4 + (c & 3)] = umma::to_tf32(v);
                }
            }
        }
        umma::fence_smem_to_async(); umma::fence_before_sync();
        __syncthreads();                                   // Xs complete, staging buffer free again
        hc.ys = (A.mode == MODE_FUSED) ? ys_tile[yslot] : nullptr;
        const bool has_next = tile + (int)gridDim.x < ntiles;
        // ================= encoder forward
        if constexpr (ENC == ENC_INSOLE) {
            GAITK_MMA_ISSUE_SYNCED((issue_conv_tab<KT1, KX, N1>(tmem, dtab[0])));
            GAITK_MMA_WAIT();
            {
                float a1[MH::NH], ha[MH::NH], d1[MH::NH], b1v[MH::NH];
                MH::tmem(trow, h, a1);
                MH::params(b1s, h, b1v);
#pragma unroll
                for (int c = 0; c < MH::NH; ++c) gelu_fwd_fast(a1[c] + b1v[c], ha[c], d1[c]);
                MH::template store<RB, true>(HAs, row, h, ha);
                if (train) MH::template store<RB, false>(D1s, row, h, d1);
            }
            GAITK_MMA_PHASE((issue_conv_tab<3, KH, NC>(tmem, dtab[1])));
        } else {
            GAITK_MMA_ISSUE_SYNCED((issue_conv_tab<KT1, KX, N1>(tmem, dtab[0])));
        }
        GAITK_MMA_WAIT();
        float rstd;
        {
            float a[MC::NH], g[MC::NH], d[MC::NH], xh[MC::NH], f[MC::NH];
            MC::tmem(trow, h, a);
#pragma unroll
            for (int c = 0; c < MC::NH; ++c) gelu_fwd_fast(a[c] + bias_c[c], g[c], d[c]);
            // LayerNorm over the 12 channels of the row: local (mean, M2) of 6, merged with the other half's
            float mu = 0.f;
#pragma unroll
            for (int c = 0; c < MC::NH; ++c) mu += g[c];
            mu *= (1.0f / MC::NH);
            float m2 = 0.f;
#pragma unroll
            for (int c = 0; c < MC::NH; ++c) { const float dd = g[c] - mu; m2 = fmaf(dd, dd, m2); }
            lnx[r][2 * h] = mu; lnx[r][2 * h + 1] = m2;
            pair_barrier(wq);
            const float4 ex = *reinterpret_cast<const float4*>(lnx[r]);
            const float delta = ex.z - ex.x;
            const float mean = 0.5f * (ex.x + ex.z);
            const float var = ((ex.y + ex.w) + delta * delta * (0.5f * MC::NH)) * (1.0f / C);
            rstd = rsqrtf(var + 1e-5f);
#pragma unroll
            for (int c = 0; c < MC::NH; ++c) { xh[c] = (g[c] - mean) * rstd; f[c] = fmaf(xh[c], lng_c[c], lnb_c[c]); }
            MC::template store<RB, true>(Fs, row, h, f);
            if (train) { MC::template store<RB, false>(Ds, row, h, d); MC::template store<RB, false>(XHs, row, h, xh); }
        }
        // ================= shared backbone forward
        GAITK_MMA_PHASE((issue_conv_tab<3, KC, NS>(tmem, dtab[2])));
        GAITK_MMA_WAIT();
        uint32_t zmask = 0;                                // ReLU mask of this thread's 8 channels (kept in a register)
        {
            float z[MS::NH];
            MS::tmem(trow, h, z);
#pragma unroll
            for (int s = 0; s < MS::NH; ++s) { z[s] = fmaxf(z[s] + bb_c[s], 0.f); zmask |= (z[s] > 0.f ? 1u : 0u) << s; }
            // rows of one (window, bin) are the lanes r, r+W, ..., r+(binsz-1)W of this warp: pool with shuffles
#pragma unroll
            for (int o = W; o < binsz * W; o <<= 1)
#pragma unroll
                for (int s = 0; s < MS::NH; ++s) z[s] += __shfl_xor_sync(0xffffffffu, z[s], o);
            const int tt = r >> logW, w = r & (W - 1);
            if ((tt & (binsz - 1)) == 0) {
                const float inv = 1.0f / (float)binsz;
                float4* dst = reinterpret_cast<float4*>(Ps + w * NF + (tt / binsz) * S + h * MS::NH);
#pragma unroll
                for (int s4 = 0; s4 < MS::NH / 4; ++s4)
                    dst[s4] = make_float4(z[s4 * 4] * inv, z[s4 * 4 + 1] * inv, z[s4 * 4 + 2] * inv, z[s4 * 4 + 3] * inv);
            }
        }
        umma::fence_before_sync();
        __syncthreads();
        // ================= head + loss (warp per window)
        if (wrp < W) head.run(A, hc, wrp, lane, win0, train, inv_denom);
        if (!train) { if (has_next) prefetch(tile + (int)gridDim.x); __syncthreads(); continue; }
        __syncthreads();
        // ================= dz through pool + ReLU (tf32: it is an MMA operand)
        {
            const int t = r >> logW, w = r & (W - 1);
            float dz[MS::NH];
            const float4* dp = reinterpret_cast<const float4*>(DPs + w * NF + (t / binsz) * S + h * MS::NH);
#pragma unroll
            for (int s4 = 0; s4 < MS::NH / 4; ++s4) {
                const float4 d = dp[s4];
                dz[s4 * 4] = d.x; dz[s4 * 4 + 1] = d.y; dz[s4 * 4 + 2] = d.z; dz[s4 * 4 + 3] = d.w;
            }
#pragma unroll
            for (int s = 0; s < MS::NH; ++s) { dz[s] = ((zmask >> s) & 1u) ? dz[s] : 0.f; g_bb[s] += dz[s]; }
            MS::template store<RB, true>(Zs, row, h, dz);
        }
        // backbone dgrad on the tensor core while all warps do the backbone weight gradient
        GAITK_MMA_PHASE((issue_conv_tab<3, KS, NC>(tmem, dtab[3])));
        if (has_next) prefetch(tile + (int)gridDim.x);          // TMA bulk copies + label loads for the next tile
        g_wb.accumulate_range(Fs, RB, Zs, RB, halo, W, k0, k1, wq, lane);
        GAITK_MMA_WAIT();
        {
            float df[MC::NH], xh[MC::NH], dxh[MC::NH], da[MC::NH], d[MC::NH];
            MC::tmem(trow, h, df);
            MC::template load<RB>(XHs, row, h, xh);
            MC::template load<RB>(Ds, row, h, d);
            float m1 = 0.f, m2 = 0.f;
#pragma unroll
            for (int c = 0; c < MC::NH; ++c) {
                g_lng[c] = fmaf(df[c], xh[c], g_lng[c]); g_lnb[c] += df[c];
                dxh[c] = df[c] * lng_c[c];
                m1 += dxh[c]; m2 = fmaf(dxh[c], xh[c], m2);
            }
            lnx[r][2 * h] = m1; lnx[r][2 * h + 1] = m2;
            pair_barrier(wq);
            const float4 ex = *reinterpret_cast<const float4*>(lnx[r]);
            m1 = (ex.x + ex.z) * (1.0f / C); m2 = (ex.y + ex.w) * (1.0f / C);
#pragma unroll
            for (int c = 0; c < MC::NH; ++c) da[c] = rstd * (dxh[c] - m1 - xh[c] * m2) * d[c];
            if constexpr (ENC == ENC_INSOLE) {
#pragma unroll
                for (int c = 0; c < MC::NH; ++c) g_b2[c] += da[c];
            } else {
#pragma unroll
                for (int c = 0; c < MC::NH; ++c) g_b1[c] += da[c];
            }
            MC::template store<RB, true>(XHs, row, h, da);     // dA over XH (row-private)
        }
        if constexpr (ENC == ENC_INSOLE) {
            // conv2 dgrad on the tensor core, conv2 weight gradient on the warps
            GAITK_MMA_PHASE((issue_conv_tab<3, KC, NH>(tmem, dtab[4])));
            g_w2.accumulate_range(HAs, RB, XHs, RB, halo, W, k0, k1, wq, lane);
            GAITK_MMA_WAIT();
            {
                float dh[MH::NH], d1[MH::NH], da1[MH::NH];
                MH::tmem(trow, h, dh);
                MH::template load<RB>(D1s, row, h, d1);
#pragma unroll
                for (int c = 0; c < MH::NH; ++c) { da1[c] = dh[c] * d1[c]; g_b1[c] += da1[c]; }
                MH::template store<RB, true>(D1s, row, h, da1);      // dA1 over D1 (row-private)
            }
            umma::fence_before_sync();
            __syncthreads();
            g_w1.accumulate_range(Xs, RB, D1s, RB, halo, W, k0, k1, wq, lane);
        } else {
            umma::fence_before_sync();
            __syncthreads();
            g_w1.accumulate_range(Xs, RB, XHs, RB, halo, W, k0, k1, wq, lane);
        }
        if (has_next && wrp == PF_WARP && lane < W) ys_tile[yslot ^ 1][lane] = ylab_next;
        yslot ^= 1;
        __syncthreads();
    }
#undef GAITK_MMA_PHASE
#undef GAITK_MMA_ISSUE_SYNCED
#undef GAITK_MMA_WAIT

    // ================= teardown + flush
    umma::fence_before_sync();
    __syncthreads();
    if (wrp == 0) umma::tmem_dealloc(tmem, 32);
    if (A.mode == MODE_FWD) return;
    float* out = A.partial + (size_t)blockIdx.x * A.NGP;
    const GradOff& go = A.go;
    // weight-gradient tiles: half 0 stores, half 1 adds (fixed order -> deterministic)
    for (int pass = 0; pass < 2; ++pass) {
        if (h == pass) {
            const int OUT1 = (ENC == ENC_INSOLE) ? H : C;
            g_w1.flush_acc(out + go.w1, nullptr, CIN, OUT1, wq, lane, pass == 1);
            if constexpr (ENC == ENC_INSOLE)
                g_w2.flush_acc(out + go.w2, (A.skip_identity ? nullptr : out + go.wsk), H, C, wq, lane, pass == 1);
            g_wb.flush_acc(out + go.wbb, nullptr, C, S, wq, lane, pass == 1);
        }
        __threadfence_block();
        __syncthreads();
    }
    {
        const int OUT1 = (ENC == ENC_INSOLE) ? H : C;
        if constexpr (ENC == ENC_INSOLE) {
            flush_rowacc2<O1>(g_b1, OUT1, stage, out + go.b1, nullptr, tid);
            flush_rowacc2<C>(g_b2, C, stage, out + go.b2, (A.skip_identity ? nullptr : out + go.bsk), tid);
        } else {
            flush_rowacc2<C>(g_b1, OUT1, stage, out + go.b1, nullptr, tid);
        }
    }
    flush_rowacc2<C>(g_lng, C, stage, out + go.lng, nullptr, tid);
    flush_rowacc2<C>(g_lnb, C, stage, out + go.lnb, nullptr, tid);
    flush_rowacc2<S>(g_bb, S, stage, out + go.bbb, nullptr, tid);
    head_flush2<NFL, S>(head, A, stage, out, tid);
}

}  // namespace gaitk
