// stream_kernel_tc.cuh -- tensor-core variant of the fused stream kernel ("GAITK_DTYPE_TF32"), WearGait encoders.
//
// Same tiling, buffers and phase structure as stream_kernel.cuh; the contractions move to
// the tensor cores:
//   * forward and data-gradient convolutions  ->  tcgen05.mma kind::tf32, M = 128 (one row tile),
//     N = padded output channels (16 / 32), K = 8 per instruction.  The A operand is the activation
//     buffer [chunk][row][4] itself, addressed K-major / no-swizzle (SBO = 128 B, LBO = chunk stride); tap
//     t of a k-tap convolution is the same buffer with the start address shifted by (t - k/2) * W rows
//     = 16 B per row, the zero halo supplying the "same" padding.  The B operand is the weight matrix
//     [tap][chunk][n][4] staged once per CTA.  Accumulators live in TMEM (32 columns) and come back
//     with tcgen05.ld 32x32b: thread i gets row i, so GELU / LayerNorm / ReLU epilogues are thread-local.
//   * weight gradients need the TRANSPOSED activations as operand.  tcgen05 accepts MN-major tf32 operands
//     only in the SWIZZLE_128B_BASE32B layout (CUTLASS sm100_common.inl:92; a no-swizzle MN-major tf32
//     descriptor is silently a no-op on hardware, tests/test_gpu_umma.py), which is incompatible with the
//     K-major use of the same buffer.  They therefore run on mma.sync m16n8k8 tf32, whose per-lane
//     fragment loads read the very same buffers with arbitrary row shifts; the dW tiles are owned by
//     warps and stay in registers for the whole kernel.
// All MMA operands are rounded to tf32 (cvt.rna) when they are stored, so both tensor paths see the same
// values.  Accumulation is fp32 throughout.
#pragma once
#include "stream_common.cuh"
#include "umma.cuh"
#include "stream_tc_plan.h"

namespace gaitk {

// resident CTAs per SM the register allocation aims for: the insole stream needs ~100 KB of shared memory
// (two CTAs at most), the single-conv encoders fit three
template <class Cfg> struct TcMinBlocks { static constexpr int value = Cfg::ENC == ENC_INSOLE ? 2 : 3; };


__device__ __forceinline__ void mma_sync_tf32(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// dW[j][n] += sum_r IN[r + (tap - KT/2) W][ci] * DOUT[r][n],   j = tap * CIP + ci.
// 16 x 8 output tiles; M-tile mt belongs to warp mt % 4 (all its N-tiles), accumulators persist in registers.
// The MMA's row / column indices are PERMUTED so that a lane's fragment elements are adjacent in shared memory and
// come in with one 64-bit load each (half the load instructions of the natural order):
//   MMA row m of M-tile mt      <->  j  = 16 mt + 2 (m & 7) + (m >> 3)     (a0, a1 = channels ci, ci + 1 of one row)
//   MMA col n of N-tile 2p + q  <->  co = 16 p + 2 n + q                    (b of tiles 2p, 2p+1 = channels co, co + 1)
//   a last unpaired N-tile i    <->  co = 8 i + n
// Half-warp wavefronts of these loads touch 32 distinct banks for RB = 4 (mod 8).
// Work split over the four warps of a row tile: KS = false -> M-tile mt belongs to warp mt % 4 (all rows);
// KS = true -> every warp holds ALL M-tiles for its quarter of the rows (balanced when MT is not a multiple of 4, and
// the B fragments are shared by all M-tiles), the four partial sums being added in warp order at the final flush.
template <int KT, int CIP, int NOUT, bool KS = false>
struct WgradMma {
    static_assert(CIP % 4 == 0 && NOUT % 8 == 0, "");
    static constexpr int J = KT * CIP;
    static constexpr int MT = (J + 15) / 16;
    static constexpr int NT8 = NOUT / 8;
    static constexpr int NP = NT8 / 2;                // paired N-tiles
    static constexpr int MTW = KS ? MT : (MT + 3) / 4;
    float acc[MTW][NT8][4];
    __device__ __forceinline__ static int tile_of(int i, int wrp) { return KS ? i : wrp + 4 * i; }

    __device__ __forceinline__ void zero() {
#pragma unroll
        for (int i = 0; i < MTW; ++i)
#pragma unroll
            for (int n = 0; n < NT8; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[i][n][e] = 0.f;
    }
    // rows [r_begin, r_end) of the tile (two thread groups may split the K range)
    __device__ __forceinline__ void accumulate_range(const float* __restrict__ in, int RBin, const float* __restrict__ dout, int RBout,
                                                     int halo, int W, int r_begin, int r_end, int wrp, int lane) {
        const int g = lane >> 2, t = lane & 3;
        int aoff[MTW];
#pragma unroll
        for (int i = 0; i < MTW; ++i) {
            int j = tile_of(i, wrp) * 16 + 2 * g;
            if (j >= J) j = 0;                                   // padding rows of the last tile: results ignored
            const int tap = j / CIP, ci = j - tap * CIP;
            aoff[i] = ((ci >> 2) * RBin + halo + (tap - KT / 2) * W + t) * 4 + (ci & 3);
        }
        int boff[NP + 1];
#pragma unroll
        for (int p = 0; p < NP; ++p) { const int c = 16 * p + 2 * g; boff[p] = ((c >> 2) * RBout + halo + t) * 4 + (c & 3); }
        { const int c = 8 * (NT8 - 1) + g; boff[NP] = ((c >> 2) * RBout + halo + t) * 4 + (c & 3); }
#pragma unroll 4
        for (int r0 = r_begin; r0 < r_end; r0 += 8) {
            uint32_t b[NT8][2];
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                const float2 lo = *reinterpret_cast<const float2*>(dout + boff[p] + r0 * 4);
                const float2 hi = *reinterpret_cast<const float2*>(dout + boff[p] + r0 * 4 + 16);
                b[2 * p][0] = __float_as_uint(lo.x); b[2 * p + 1][0] = __float_as_uint(lo.y);
                b[2 * p][1] = __float_as_uint(hi.x); b[2 * p + 1][1] = __float_as_uint(hi.y);
            }
            if constexpr (NT8 % 2 == 1) {
                b[NT8 - 1][0] = __float_as_uint(dout[boff[NP] + r0 * 4]);
                b[NT8 - 1][1] = __float_as_uint(dout[boff[NP] + r0 * 4 + 16]);
            }
#pragma unroll
            for (int i = 0; i < MTW; ++i) {
                if (tile_of(i, wrp) >= MT) continue;
                const float2 lo = *reinterpret_cast<const float2*>(in + aoff[i] + r0 * 4);
                const float2 hi = *reinterpret_cast<const float2*>(in + aoff[i] + r0 * 4 + 16);
                const uint32_t a[4] = {__float_as_uint(lo.x), __float_as_uint(lo.y), __float_as_uint(hi.x), __float_as_uint(hi.y)};
#pragma unroll
                for (int n = 0; n < NT8; ++n) mma_sync_tf32(acc[i][n], a, b[n]);
            }
        }
    }
    __device__ __forceinline__ void accumulate(const float* __restrict__ in, int RBin, const float* __restrict__ dout, int RBout,
                                               int halo, int W, int rows, int wrp, int lane) {
        if constexpr (KS) accumulate_range(in, RBin, dout, RBout, halo, W, wrp * (rows / 4), (wrp + 1) * (rows / 4), wrp, lane);
        else accumulate_range(in, RBin, dout, RBout, halo, W, 0, rows, wrp, lane);
    }
    // each tile has exactly one owner: write straight into the PyTorch weight layout (CO, CI, KT);
    // dst2 (optional) receives the centre tap as (CO, CI, 1) -- the folded 1x1 skip; `add` = read-modify-write
    // (second partial sum of a K-split tile)
    __device__ __forceinline__ void flush_acc(float* dst, float* dst2, int CIN, int COUT, int wrp, int lane, bool add) {
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int i = 0; i < MTW; ++i) {
            if (tile_of(i, wrp) >= MT) continue;
#pragma unroll
            for (int n = 0; n < NT8; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int j = tile_of(i, wrp) * 16 + 2 * g + (e >> 1), col = 2 * t + (e & 1);
                    const int co = n < 2 * NP ? 16 * (n >> 1) + 2 * col + (n & 1) : 8 * n + col;
                    if (j >= J) continue;
                    const int tap = j / CIP, ci = j - tap * CIP;
                    if (ci < CIN && co < COUT) {
                        float* q = dst + (co * CIN + ci) * KT + tap;
                        const float v = add ? *q + acc[i][n][e] : acc[i][n][e];
                        *q = v;
                        if (dst2 && tap == KT / 2) dst2[co * CIN + ci] = v;
                    }
                }
        }
    }
    // called by all threads of the CTA (KS: four ordered passes separated by CTA barriers)
    __device__ __forceinline__ void flush(float* dst, float* dst2, int CIN, int COUT, int wrp, int lane) {
        if constexpr (KS) {
            for (int pass = 0; pass < 4; ++pass) {
                if (wrp == pass) flush_acc(dst, dst2, CIN, COUT, wrp, lane, pass > 0);
                __syncthreads();
            }
        } else {
            flush_acc(dst, dst2, CIN, COUT, wrp, lane, false);
        }
    }
};

// one thread: D[tmem] = sum_tap A(buffer shifted by (tap - KT/2) W rows) * B_tap ;  KCH (even) chunks of 4 channels
template <int KT, int KCH, int N>
__device__ __forceinline__ void issue_conv(uint32_t tmem_d, uint32_t a_base, int RBx, int halo, int W, uint32_t b_base) {
    constexpr uint32_t idesc = umma::make_idesc_tf32(128, N, false, false);
#pragma unroll
    for (int tap = 0; tap < KT; ++tap) {
#pragma unroll
        for (int kp = 0; kp < KCH / 2; ++kp) {
            const uint32_t a = a_base + (uint32_t)(((2 * kp) * RBx + halo + (tap - KT / 2) * W) * 16);
            const uint32_t b = b_base + (uint32_t)(((tap * KCH + 2 * kp) * N) * 16);
            umma::mma_tf32(tmem_d, umma::make_desc(a, (uint32_t)RBx * 16u, 128u), umma::make_desc(b, (uint32_t)N * 16u, 128u), idesc,
                           (tap | kp) != 0 ? 1u : 0u);
        }
    }
}

// The descriptors of a convolution do not depend on the tile: build them once per kernel into shared memory
// (KT * KCH/2 pairs) so that issuing an MMA is two LDS.64 plus the instruction.
template <int KT, int KCH, int N>
__device__ __forceinline__ void build_conv_descs(uint64_t* tab, uint32_t a_base, int RBx, int halo, int W, uint32_t b_base) {
    for (int i = 0; i < KT * (KCH / 2); ++i) {
        const int tap = i / (KCH / 2), kp = i - tap * (KCH / 2);
        const uint32_t a = a_base + (uint32_t)(((2 * kp) * RBx + halo + (tap - KT / 2) * W) * 16);
        const uint32_t b = b_base + (uint32_t)(((tap * KCH + 2 * kp) * N) * 16);
        tab[2 * i] = umma::make_desc(a, (uint32_t)RBx * 16u, 128u);
        tab[2 * i + 1] = umma::make_desc(b, (uint32_t)N * 16u, 128u);
    }
}
template <int KT, int KCH, int N>
__device__ __forceinline__ void issue_conv_tab(uint32_t tmem_d, const uint64_t* tab) {
    constexpr uint32_t idesc = umma::make_idesc_tf32(128, N, false, false);
    constexpr int NM = KT * (KCH / 2);
    uint64_t d[2 * NM];
#pragma unroll
    for (int i = 0; i < 2 * NM; ++i) d[i] = tab[i];
#pragma unroll
    for (int i = 0; i < NM; ++i) umma::mma_tf32(tmem_d, d[2 * i], d[2 * i + 1], idesc, i != 0 ? 1u : 0u);
}

template <int N>
__device__ __forceinline__ void store_row_tf32(float* buf, int RBx, int halo, int r, const float (&v)[N]) {
    float4* p = reinterpret_cast<float4*>(buf) + (halo + r);
#pragma unroll
    for (int c4 = 0; c4 < N / 4; ++c4)
        p[c4 * RBx] = make_float4(umma::to_tf32(v[c4 * 4]), umma::to_tf32(v[c4 * 4 + 1]), umma::to_tf32(v[c4 * 4 + 2]), umma::to_tf32(v[c4 * 4 + 3]));
}
// this thread's accumulator row: N (multiple of 8) fp32 columns
template <int N>
__device__ __forceinline__ void tmem_row(uint32_t taddr, float (&v)[N]) {
#pragma unroll
    for (int c = 0; c < N; c += 8) umma::ld_x8(taddr + c, &v[c]);
    umma::ld_wait();
}

#ifdef GAITK_PHASE_TIMING
#define PH(i) do { if (lane == 0) { const long long t_ = clock64(); ph_acc[wrp][i] += t_ - ph_last; ph_last = t_; } } while (0)
#else
#define PH(i) do { } while (0)
#endif

// FX: the WearGait default geometry (T = 64 -> W = 2 windows per 128-row tile, 8 pooling bins) as compile-time
// constants, so that shared-memory addressing folds into immediates; FX = false keeps every 128-row geometry.
__host__ __device__ constexpr int tc_round_rb(int rows, int halo) {
    int rb = rows + 2 * halo;
    while (rb % 8 != GAITK_RB_MOD) ++rb;
    return rb;
}
template <class Cfg, bool FX>
__global__ void __launch_bounds__(NT, TcMinBlocks<Cfg>::value) stream_kernel_tc(const StreamArgs A, const TcPlan SP) {
    extern __shared__ __align__(1024) float sm[];
    __shared__ uint64_t bar, ldbar;
    __shared__ uint32_t tmem_slot;
    __shared__ int ys_tile[2][WMAX];                      // labels of the current / next tile
    __shared__ uint64_t dtab[5][2 * 12];                  // MMA descriptor pairs: conv1, conv2, bb, bb dgrad, conv2 dgrad
#ifdef GAITK_PHASE_TIMING
    __shared__ long long ph_acc[4][16];
    long long ph_last = 0;
    if (threadIdx.x < 64) ph_acc[threadIdx.x / 16][threadIdx.x % 16] = 0;
#endif
    const int tid = threadIdx.x, lane = tid & 31, wrp = tid >> 5;
    constexpr int ENC = Cfg::ENC, CIN = Cfg::CIN, CI4 = Cfg::CI4, KT1 = Cfg::KT1, H = Cfg::H, H4 = Cfg::H4;
    constexpr int C = Cfg::C, C4 = Cfg::C4, CP = Cfg::CP, S = Cfg::S, S4 = Cfg::S4, NFL = Cfg::NFL;
    static_assert(ENC == ENC_CONV_GELU_LN || ENC == ENC_INSOLE, "tensor-core kernel: WearGait encoders");
    constexpr int KX = (CI4 + 1) / 2 * 2;                 // even chunk counts (K = 8 per MMA)
    constexpr int KH = (H4 + 1) / 2 * 2, KC = (C4 + 1) / 2 * 2, KS = (S4 + 1) / 2 * 2;
    constexpr int N1 = (ENC == ENC_INSOLE) ? ((H + 15) / 16) * 16 : ((C + 15) / 16) * 16;   // first conv outputs
    constexpr int NC = ((C + 15) / 16) * 16, NS = ((S + 15) / 16) * 16, NH = ((H + 15) / 16) * 16;
    constexpr int O1 = (ENC == ENC_INSOLE) ? H4 * 4 : CP;
    static_assert(N1 <= 32 && NC <= 32 && NS <= 32 && (H == 0 || NH <= 32), "accumulator fits 32 TMEM columns");
    static_assert(S <= 32, "ReLU mask is one 32-bit word");
    constexpr int FX_HALO = (KT1 / 2 > 1 ? KT1 / 2 : 1) * 2;
    const int W = FX ? 2 : A.W, halo = FX ? FX_HALO : A.halo, RB = FX ? tc_round_rb(128, FX_HALO) : A.RB;
    const int rows = FX ? 128 : A.rows, T = FX ? 64 : A.T;
    const int K = A.K, NF = A.NF, bdim = FX ? 8 : A.bdim;
    const bool train = A.mode != MODE_FWD;

    float* Xs = sm + SP.X; float* HAs = sm + SP.HA; float* D1s = sm + SP.D1; float* XHs = sm + SP.XH;
    float* Ds = sm + SP.D; float* Fs = sm + SP.F; float* Zs = sm + SP.Z;
    float* w1b = sm + SP.W1B; float* b1s = sm + SP.B1; float* w2b = sm + SP.W2B; float* b2s = sm + SP.B2; float* w2d = sm + SP.W2D;
    float* lngs = sm + SP.LNG; float* lnbs = sm + SP.LNB; float* wbb = sm + SP.WBB; float* bbs = sm + SP.BB; float* wbd = sm + SP.WBD;
    float* hws = sm + SP.HW; float* hbs = sm + SP.HB; float* hngs = sm + SP.HNG; float* hnbs = sm + SP.HNB; float* inws = sm + SP.INW;
    float* DPs = sm + SP.DP; int* bins = reinterpret_cast<int*>(sm + SP.BINS); float* stage = sm + SP.STAGE;
    float* STGs = sm + SP.STG;                            // raw window bytes of the NEXT tile (TMA bulk prefetch)

    // ---- one-time setup
    for (int i = tid; i < SP.total; i += NT) sm[i] = 0.f;
    if (tid == 0) { umma::mbar_init(&bar, 1); umma::mbar_init(&ldbar, 1); umma::fence_mbar_init(); }
    if (wrp == 0) umma::tmem_alloc(&tmem_slot, 32);
    __syncthreads();
    // B operands [tap][chunk][n][4]: PyTorch (O, CIN, KT) -> value(n = o, k = ci)
    {
        const int OUT1 = (ENC == ENC_INSOLE) ? H : C;
        for (int i = tid; i < OUT1 * CIN * KT1; i += NT) {
            const int o = i / (CIN * KT1), ci = (i / KT1) % CIN, tap = i % KT1;
            w1b[((tap * KX + (ci >> 2)) * N1 + o) * 4 + (ci & 3)] = umma::to_tf32(A.w1[i]);
        }
        for (int i = tid; i < OUT1; i += NT) b1s[i] = A.b1[i];
    }
    if constexpr (ENC == ENC_INSOLE) {
        for (int i = tid; i < C * H * 3; i += NT) {
            const int o = i / (H * 3), ci = (i / 3) % H, tap = i % 3;
            float w = A.w2[i];
            if (tap == 1) w += A.skip_identity ? (o == ci ? 1.f : 0.f) : A.wsk[o * H + ci];
            w = umma::to_tf32(w);
            w2b[((tap * KH + (ci >> 2)) * NC + o) * 4 + (ci & 3)] = w;                    // fwd: n = o, k = ci
            w2d[(((2 - tap) * KC + (o >> 2)) * NH + ci) * 4 + (o & 3)] = w;               // dgrad: n = ci, k = o, flipped taps
        }
        for (int i = tid; i < C; i += NT) b2s[i] = A.b2[i] + (A.skip_identity ? 0.f : A.bsk[i]);
    }
    for (int i = tid; i < C; i += NT) { lngs[i] = A.lng[i]; lnbs[i] = A.lnb[i]; }
    for (int i = tid; i < S * C * 3; i += NT) {
        const int o = i / (C * 3), ci = (i / 3) % C, tap = i % 3;
        const float w = umma::to_tf32(A.wbb[i]);
        wbb[((tap * KC + (ci >> 2)) * NS + o) * 4 + (ci & 3)] = w;
        wbd[(((2 - tap) * KS + (o >> 2)) * NC + ci) * 4 + (o & 3)] = w;
    }
    for (int i = tid; i < S; i += NT) bbs[i] = A.bbb[i];
    for (int i = tid; i < K * NF; i += NT) hws[i] = A.hw[i];
    if (A.hb) for (int i = tid; i < K; i += NT) hbs[i] = A.hb[i];
    if (A.head_norm) for (int i = tid; i < NF; i += NT) { hngs[i] = A.hng[i]; hnbs[i] = A.hnb[i]; }
    int* bin_s = bins; int* bin_e = bins + bdim; int* t_lo = bins + 2 * bdim; int* t_hi = t_lo + T;
    for (int b = tid; b < bdim; b += NT) { bin_s[b] = (b * T) / bdim; bin_e[b] = ((b + 1) * T + bdim - 1) / bdim; }
    __syncthreads();
    for (int t = tid; t < T; t += NT) {
        int lo = bdim, hi = -1;
        for (int b = 0; b < bdim; ++b) if (t >= bin_s[b] && t < bin_e[b]) { lo = min(lo, b); hi = max(hi, b); }
        t_lo[t] = lo; t_hi[t] = hi;
    }
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
    const uint32_t trow = tmem + ((uint32_t)(wrp * 32) << 16);          // this warp's 32 TMEM lanes
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

    // ---- persistent accumulators
    // row-split (balanced) weight gradients where the register budget allows it
    constexpr bool KS_W1 = (ENC != ENC_INSOLE) && CIN <= 4, KS_WB = KS_W1;       // walkway only: the IMU stream is at its register cap
    WgradMma<KT1, CI4 * 4, (O1 + 7) / 8 * 8, KS_W1> g_w1;
    WgradMma<3, (ENC == ENC_INSOLE ? H4 * 4 : 4), (ENC == ENC_INSOLE ? NC : 8)> g_w2;
    WgradMma<3, CP, NS, KS_WB> g_wb;
    float g_b1[O1], g_b2[ENC == ENC_INSOLE ? CP : 4], g_lng[CP], g_lnb[CP], g_bb[S];
    HeadState<NFL, S> head; head.zero();
    HeadCtx hc; hc.Zs = Zs; hc.RB = RB; hc.halo = halo; hc.W = W; hc.S = S; hc.bin_s = bin_s; hc.bin_e = bin_e;
    hc.hws = hws; hc.hbs = hbs; hc.hngs = hngs; hc.hnbs = hnbs; hc.inws = inws; hc.DPs = DPs;
    // uniform power-of-two bins inside a warp: pool with shuffles in the backbone epilogue (all threads)
    const int binsz = (T % bdim == 0) ? T / bdim : 0;
    const bool pool_shfl = binsz > 0 && (binsz & (binsz - 1)) == 0 && binsz * W <= 32;
    float* Ps = sm + SP.P;
    hc.Ps = pool_shfl ? Ps : nullptr;
    hc.ys = nullptr;
    hc.inv_bin = (binsz > 0 && (binsz & (binsz - 1)) == 0) ? 1.0f / (float)binsz : 0.f;
    const int logW = FX ? 1 : 31 - __clz(W);             // W is a power of two (planner)
    g_w1.zero(); g_w2.zero(); g_wb.zero();
#pragma unroll
    for (int i = 0; i < O1; ++i) g_b1[i] = 0.f;
#pragma unroll
    for (int i = 0; i < (ENC == ENC_INSOLE ? CP : 4); ++i) g_b2[i] = 0.f;
#pragma unroll
    for (int i = 0; i < CP; ++i) { g_lng[i] = 0.f; g_lnb[i] = 0.f; }
#pragma unroll
    for (int i = 0; i < S; ++i) g_bb[i] = 0.f;
    const float inv_denom = (A.mode == MODE_FUSED) ? 1.0f / A.denom[0] : 0.f;

    // issue (one thread) + wait (all): the accumulator is ready in TMEM afterwards
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

    const int r = tid;                                    // rows == NT for the supported configurations
    const int ntiles = (A.B + W - 1) / W;
    const int per_win = T * CIN;
    uint32_t ldphase = 0;
    // warp 0 prefetches the W windows of a tile into the staging buffer: one cp.async.bulk per window
    // (16-byte aligned source), or a plain copy for a window whose frame-store offset is not aligned.
    int ylab_next = 0;
    constexpr int PF_WARP = 3;                           // has slack in the backbone weight-gradient phase (owns no tile there)
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
#ifdef GAITK_PHASE_TIMING
    __syncthreads(); ph_last = clock64();
#endif
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int win0 = tile * W;
        // ================= staged window bytes -> Xs [chunk][row][4] (tf32)
        if (!A.zero_input) {
            umma::mbar_wait(&ldbar, ldphase); ldphase ^= 1u;
            PH(0);
            // thread r converts row r = (t, w): CIN consecutive floats of window w's frame t -> one 16-byte word per chunk
            {
                const int t = r >> logW, w = r & (W - 1);
                const bool live = win0 + w < A.B;
                const float* src = STGs + w * per_win + t * CIN;
                float4* dst = reinterpret_cast<float4*>(Xs) + (halo + r);
                if constexpr (CIN % 4 == 0) {
#pragma unroll
                    for (int c4 = 0; c4 < CI4; ++c4) {
                        float4 v = reinterpret_cast<const float4*>(src)[c4];
                        if (!live) v = make_float4(0.f, 0.f, 0.f, 0.f);
                        dst[c4 * RB] = make_float4(umma::to_tf32(v.x), umma::to_tf32(v.y), umma::to_tf32(v.z), umma::to_tf32(v.w));
                    }
                } else {
                    float v[CI4 * 4];
#pragma unroll
                    for (int c = 0; c < CI4 * 4; ++c) v[c] = (c < CIN && live) ? umma::to_tf32(src[c]) : 0.f;
#pragma unroll
                    for (int c4 = 0; c4 < CI4; ++c4) dst[c4 * RB] = make_float4(v[4 * c4], v[4 * c4 + 1], v[4 * c4 + 2], v[4 * c4 + 3]);
                }
            }
        }
        umma::fence_smem_to_async(); umma::fence_before_sync();
        PH(1);
        __syncthreads();                                   // Xs complete, staging buffer free again
        PH(2);
        hc.ys = (A.mode == MODE_FUSED) ? ys_tile[yslot] : nullptr;
        const bool has_next = tile + (int)gridDim.x < ntiles;
        // ================= encoder forward
        if constexpr (ENC == ENC_INSOLE) {
            GAITK_MMA_ISSUE_SYNCED((issue_conv_tab<KT1, KX, N1>(tmem, dtab[0])));
            GAITK_MMA_WAIT();
            PH(3);
            {
                float a1[N1], ha[O1], d1[O1];
                tmem_row<N1>(trow, a1);
#pragma unroll
                for (int c = 0; c < O1; ++c) { if (c < H) gelu_fwd_fast(a1[c] + b1s[c], ha[c], d1[c]); else { ha[c] = 0.f; d1[c] = 0.f; } }
                store_row_tf32<O1>(HAs, RB, halo, r, ha);
                if (train) store_row<O1>(D1s, RB, halo, r, d1);
            }
            PH(4);
            GAITK_MMA_PHASE((issue_conv_tab<3, KH, NC>(tmem, dtab[1])));
        } else {
            GAITK_MMA_ISSUE_SYNCED((issue_conv_tab<KT1, KX, N1>(tmem, dtab[0])));
        }
        GAITK_MMA_WAIT();
        PH(5);
        float rstd_row = 0.f;                            // LayerNorm 1/sigma of this row, kept for the backward
        {
            float a[NC], g[CP], d[CP], xh[CP], f[CP]; float rstd;
            tmem_row<NC>(trow, a);
            const float* bias = (ENC == ENC_INSOLE) ? b2s : b1s;
#pragma unroll
            for (int c = 0; c < CP; ++c) { if (c < C) gelu_fwd_fast(a[c] + bias[c], g[c], d[c]); else { g[c] = 0.f; d[c] = 0.f; } }
            ln_fwd<CP, C>(g, xh, rstd);
#pragma unroll
            for (int c = 0; c < CP; ++c) f[c] = c < C ? fmaf(xh[c], lngs[c], lnbs[c]) : 0.f;
            store_row_tf32<CP>(Fs, RB, halo, r, f);
            if (train) { store_row<CP>(Ds, RB, halo, r, d); store_row<CP>(XHs, RB, halo, r, xh); rstd_row = rstd; }
        }
        PH(6);
        // ================= shared backbone forward
        GAITK_MMA_PHASE((issue_conv_tab<3, KC, NS>(tmem, dtab[2])));
        GAITK_MMA_WAIT();
        PH(7);
        uint32_t zmask = 0;                              // ReLU mask of this row's S backbone channels
        {
            float z[NS];
            tmem_row<NS>(trow, z);
            float zz[S];
#pragma unroll
            for (int s = 0; s < S; ++s) { zz[s] = fmaxf(z[s] + bbs[s], 0.f); zmask |= (zz[s] > 0.f ? 1u : 0u) << s; }
            // with shuffle pooling nothing reads Z itself again: the backward needs only the ReLU mask (a register)
            if (!pool_shfl) store_row<S>(Zs, RB, halo, r, zz);
            if constexpr (FX && S == 16) {
                // W = 2, 8 frames per bin: the 8 rows of one (window, bin) are the lanes with equal (lane & 1, lane >> 4).
                // Reduce-scatter butterfly over lane bits 3, 2, 1: at every stage a lane keeps half of its channels and
                // hands the other half to its partner (14 shuffles instead of the 48 of an all-reduce); afterwards
                // lane l holds the sums of channels 2 * ((l >> 1) & 7) and + 1 of its (window, bin).
                float a8[8], a4[4], a2[2];
                const bool h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float keep = h8 ? zz[8 + i] : zz[i], send = h8 ? zz[i] : zz[8 + i];
                    a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float keep = h4 ? a8[4 + i] : a8[i], send = h4 ? a8[i] : a8[4 + i];
                    a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const float keep = h2 ? a4[2 + i] : a4[i], send = h2 ? a4[i] : a4[2 + i];
                    a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
                }
                const int ch = (h8 ? 8 : 0) + (h4 ? 4 : 0) + (h2 ? 2 : 0);
                const int w = r & 1, bin = r >> 4;                        // r = 32 wrp + lane: 16 rows per bin
                *reinterpret_cast<float2*>(Ps + w * NF + bin * S + ch) = make_float2(a2[0] * 0.125f, a2[1] * 0.125f);
            } else if (pool_shfl) {
                // rows of one (window, bin) are the lanes r, r+W, ..., r+(binsz-1)W of this warp
                for (int o = W; o < binsz * W; o <<= 1)
#pragma unroll
                    for (int s = 0; s < S; ++s) zz[s] += __shfl_xor_sync(0xffffffffu, zz[s], o);
                const int tt = r >> logW, w = r & (W - 1);
                if ((tt & (binsz - 1)) == 0) {
                    const float inv = 1.0f / (float)binsz;
                    float4* dst = reinterpret_cast<float4*>(Ps + w * NF + (tt / binsz) * S);
#pragma unroll
                    for (int s4 = 0; s4 < S / 4; ++s4)
                        dst[s4] = make_float4(zz[s4 * 4] * inv, zz[s4 * 4 + 1] * inv, zz[s4 * 4 + 2] * inv, zz[s4 * 4 + 3] * inv);
                }
            }
        }
        PH(8);
        umma::fence_before_sync();
        __syncthreads();
        PH(9);
        // ================= pool + head + loss (warp per window)
        if (wrp < W) {
            if (K == 2) head.template run<true, 2>(A, hc, wrp, lane, win0, train, inv_denom);
            else if (K == 3) head.template run<true, 3>(A, hc, wrp, lane, win0, train, inv_denom);
            else head.template run<true, 0>(A, hc, wrp, lane, win0, train, inv_denom);
        }
        if (!train) { if (has_next) prefetch(tile + (int)gridDim.x); __syncthreads(); continue; }
        PH(10);
        __syncthreads();
        PH(11);
        // ================= dz through pool + ReLU, in place over Z (tf32: it is an MMA operand now)
        {
            const int t = r >> logW, w = r & (W - 1);
            float dz[S];
            const int lo = t_lo[t], hi = t_hi[t];
            if (lo == hi) {
                const float4* dp = reinterpret_cast<const float4*>(DPs + w * NF + lo * S);
#pragma unroll
                for (int s4 = 0; s4 < S / 4; ++s4) {
                    const float4 d = dp[s4];
                    dz[s4 * 4] = d.x; dz[s4 * 4 + 1] = d.y; dz[s4 * 4 + 2] = d.z; dz[s4 * 4 + 3] = d.w;
                }
            } else {
#pragma unroll
                for (int s = 0; s < S; ++s) {
                    float d = 0.f;
                    for (int b = lo; b <= hi; ++b) d += DPs[w * NF + b * S + s];
                    dz[s] = d;
                }
            }
#pragma unroll
            for (int s = 0; s < S; ++s) { dz[s] = ((zmask >> s) & 1u) ? dz[s] : 0.f; g_bb[s] += dz[s]; }
            store_row_tf32<S>(Zs, RB, halo, r, dz);
        }
        PH(12);
        // backbone dgrad on the tensor core while all warps do the backbone weight gradient
        GAITK_MMA_PHASE((issue_conv_tab<3, KS, NC>(tmem, dtab[3])));
        if (has_next) prefetch(tile + (int)gridDim.x);          // TMA bulk copies + label loads for the next tile (warp 3)
        g_wb.accumulate(Fs, RB, Zs, RB, halo, W, rows, wrp, lane);
        PH(13);
        GAITK_MMA_WAIT();
        PH(14);
        {
            float df[NC], xh[CP], dxh[CP], dg[CP], da[CP], d[CP];
            tmem_row<NC>(trow, df);
            load_row<CP>(XHs, RB, halo, r, xh);
            load_row<CP>(Ds, RB, halo, r, d);
            const float rstd = rstd_row;
#pragma unroll
            for (int c = 0; c < CP; ++c) {
                const float dfc = c < C ? df[c] : 0.f;
                g_lng[c] = fmaf(dfc, xh[c], g_lng[c]); g_lnb[c] += dfc;
                dxh[c] = c < C ? dfc * lngs[c] : 0.f;
            }
            ln_bwd<CP, C>(dxh, xh, rstd, dg);
#pragma unroll
            for (int c = 0; c < CP; ++c) da[c] = dg[c] * d[c];
            if constexpr (ENC == ENC_INSOLE) {
#pragma unroll
                for (int c = 0; c < CP; ++c) g_b2[c] += da[c];
            } else {
#pragma unroll
                for (int c = 0; c < CP; ++c) g_b1[c] += da[c];
            }
            store_row_tf32<CP>(XHs, RB, halo, r, da);          // dA over XH (row-private)
        }
        if constexpr (ENC == ENC_INSOLE) {
            // conv2 dgrad on the tensor core, conv2 weight gradient on the warps
            GAITK_MMA_PHASE((issue_conv_tab<3, KC, NH>(tmem, dtab[4])));
            g_w2.accumulate(HAs, RB, XHs, RB, halo, W, rows, wrp, lane);
            GAITK_MMA_WAIT();
            {
                float dh[NH], d1[O1], da1[O1];
                tmem_row<NH>(trow, dh);
                load_row<O1>(D1s, RB, halo, r, d1);
#pragma unroll
                for (int c = 0; c < O1; ++c) { da1[c] = c < H ? dh[c] * d1[c] : 0.f; g_b1[c] += da1[c]; }
                store_row_tf32<O1>(D1s, RB, halo, r, da1);      // dA1 over D1 (row-private)
            }
            umma::fence_before_sync();
            __syncthreads();
            g_w1.accumulate(Xs, RB, D1s, RB, halo, W, rows, wrp, lane);
        } else {
            umma::fence_before_sync();
            __syncthreads();
            g_w1.accumulate(Xs, RB, XHs, RB, halo, W, rows, wrp, lane);
        }
        if (has_next && wrp == PF_WARP && lane < W) ys_tile[yslot ^ 1][lane] = ylab_next;
        yslot ^= 1;
        PH(15);
        __syncthreads();
    }
#ifdef GAITK_PHASE_TIMING
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x < 4) {
        const int w = threadIdx.x;
        printf("ENC%d CIN%d warp%d:", ENC, CIN, w);
        for (int i = 0; i < 16; ++i) printf(" %lld", ph_acc[w][i] / max(1, (ntiles + (int)gridDim.x - 1) / (int)gridDim.x));
        printf("\n");
    }
#endif
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
    {
        const int OUT1 = (ENC == ENC_INSOLE) ? H : C;
        g_w1.flush(out + go.w1, nullptr, CIN, OUT1, wrp, lane);
        flush_rowacc<O1>(g_b1, OUT1, stage, out + go.b1, nullptr, tid);
    }
    if constexpr (ENC == ENC_INSOLE) {
        g_w2.flush(out + go.w2, (A.skip_identity ? nullptr : out + go.wsk), H, C, wrp, lane);
        flush_rowacc<CP>(g_b2, C, stage, out + go.b2, (A.skip_identity ? nullptr : out + go.bsk), tid);
    }
    flush_rowacc<CP>(g_lng, C, stage, out + go.lng, nullptr, tid);
    flush_rowacc<CP>(g_lnb, C, stage, out + go.lnb, nullptr, tid);
    g_wb.flush(out + go.wbb, nullptr, C, S, wrp, lane);
    flush_rowacc<S>(g_bb, S, stage, out + go.bbb, nullptr, tid);
    head.flush(A, stage, out, tid);
}

}  // namespace gaitk
