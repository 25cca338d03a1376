// stream_kernel_ws.cuh -- warp-specialised, all-tcgen05 fused stream kernel ("GAITK_DTYPE_BF16X3"), WearGait encoders.
//
// One persistent CTA per SM.  G independent 128-row tiles ("groups", 4 row warps each, thread = row) are in flight at
// once and share ONE copy of the weights, ONE TMEM allocation and a service warpgroup in which warp s belongs to group s:
// its lane 0 issues every tcgen05.mma of that group (forward / data-gradient convolutions first, the weight- and
// bias-gradient MMAs nobody waits for behind them), the warp issues the TMA bulk copies (cp.async.bulk) of the group's
// next tile + labels.
//
// Row warps and service warps talk through mbarriers only (per group: operands-ready rdy[k] with one arrival per row warp, done /
// free / wfree committed by tcgen05.commit, ld completed by the bulk copies' transaction bytes), so the tensor core
// works on one group's convolution while the other groups run their GELU / LayerNorm / pooling / head epilogues: the
// barrier -> issue -> spin cycle of stream_kernel_tc.cuh is gone, and so is every CTA-wide barrier in the tile loop.
//
// Arithmetic ("bf16x3"): every MMA operand x is stored as the bf16 PAIR (hi, lo), hi = bf16(x), lo = bf16(x - hi)
// -- the same 4 bytes per element as fp32/tf32 -- and every contraction is three kind::f16 MMAs hi*hi + lo*hi + hi*lo
// with fp32 accumulation in TMEM: relative operand error 2^-16 (tf32: 2^-11), i.e. fp32-grade gradients at tensor-core
// speed.  16-bit operands may be MN-major in the no-swizzle layout (tf32 may not, tests/test_gpu_umma.py), so the WEIGHT
// GRADIENTS run on tcgen05 too: dW[tap] = IN_shifted^T * DOUT with both operands read straight out of the activation
// planes (rows = K).  Their accumulators live in TMEM for the whole kernel (accumulate across tiles and groups, read back
// once), and the bias / LayerNorm-beta gradients are MMAs against an identity matrix into persistent TMEM columns: no
// mma.sync, no per-thread gradient registers.
//
// Layout: an activation buffer is a set of PLANES [row][8 x bf16] (16 bytes per row, RB = 128 + 2 * HALO rows with zero
// halos), ordered [part (hi, lo)][chunk of 8 channels].  K-major use (convolutions): rows = M, tap t = descriptor start
// shifted by (t - k/2) * 2 rows; one K step of 16 channels pairs plane 2k with plane 2k+1 through LBO, an odd plane count
// pairs its last plane with a shared zero plane.  MN-major use (weight gradients): rows = K, SBO = plane stride.
// Reference semantics: see stream_common.cuh.  Geometry: the WearGait default window (T = 64, 2 windows per tile, 8
// pooling bins); other geometries use stream_kernel_tc / stream_kernel.
#pragma once
#include <stdio.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <type_traits>

#include "stream_common.cuh"
#include "umma.cuh"

namespace gaitk {

constexpr int WS_NPH = 6;                    // operand-ready sync points per tile (simple encoders use 4)

template <class Cfg, int G_, int SL_ = 1>
struct WsLayout {
    static constexpr int G = G_;
    // SL tiles ("slots") per row warpgroup: a row thread alternates between the same phase of its SL tiles, so the MMA of one
    // tile runs under the epilogue of the other (ping-pong) instead of under other warps' epilogues; RG = G / SL row warpgroups
    static constexpr int SL = SL_, RG = G_ / SL_;
    static_assert(G_ % SL_ == 0, "groups = row warpgroups x slots");
    static constexpr int ENC = Cfg::ENC, CIN = Cfg::CIN, KT1 = Cfg::KT1, H = Cfg::H, C = Cfg::C, S = Cfg::S;
    static constexpr bool INS = ENC == ENC_INSOLE;
    static_assert(ENC == ENC_CONV_GELU_LN || ENC == ENC_INSOLE, "WearGait encoders");
    static_assert(C <= 16 && S == 16 && Cfg::NFL == 4, "C <= 16, S = 16, 128 head features");
    static constexpr int HALO = (KT1 / 2 > 1 ? KT1 / 2 : 1) * 2;
    static constexpr int RB = 128 + 2 * HALO;
    static constexpr int PL = RB * 16;                                   // plane bytes
    static constexpr int NX8 = (CIN + 7) / 8, NX8E = (NX8 + 1) / 2 * 2;
    static constexpr int NH8 = INS ? (H + 7) / 8 : 0, NH8E = (NH8 + 1) / 2 * 2;
    static constexpr int NC8 = 2, NS8 = 2;                               // C (<= 16) and S (= 16) as chunks of 8
    static constexpr int O1 = INS ? H : C;                               // real outputs of the first convolution
    static constexpr int N1 = ((O1 + 7) / 8 * 8) < 16 ? 16 : (O1 + 7) / 8 * 8;     // its MMA N
    static constexpr int NH = INS ? (H + 7) / 8 * 8 : 0;                 // conv2 data-gradient N
    static_assert(N1 <= 32 && NH <= 32, "accumulator fits 32 TMEM columns");
    // ---- group planes
    static constexpr int P_X = 0;
    static constexpr int P_HA = P_X + 2 * NX8;
    static constexpr int P_XH = P_HA + 2 * NH8;                          // xh fp32 (3 planes), later dA hi/lo (4 planes)
    static constexpr int P_D1 = P_XH + 4;
    static constexpr int ND1 = INS ? (H * 2 + 15) / 16 : 0;              // conv1 GELU derivative, fp16
    static constexpr int P_D = P_D1 + ND1;
    static constexpr int ND = (C * 2 + 15) / 16;                         // GELU derivative, fp16
    static constexpr int P_F = P_D + ND;                                 // F | Z adjacent: the TMA staging area aliases them
    static constexpr int P_Z = P_F + 2 * NC8;
    static constexpr int NPLANES = P_Z + 2 * NS8;
    static constexpr int O_P = NPLANES * PL;                             // pooled features [2][128] fp32
    static constexpr int O_DP = O_P + 1024;                              // their gradients
    static constexpr int O_YS = O_DP + 1024;                             // labels [2 slots][2]
    static constexpr int GRP = (O_YS + 32 + 127) / 128 * 128;
    // TMA staging inside F|Z: window w, frames [j FPC, (j+1) FPC) -> interior of plane w * CPW + j
    static constexpr int FPC = CIN == 13 ? 36 : (CIN == 24 ? 16 : 64);   // frames per bulk copy (<= 2048 B, multiple of 16 B)
    static constexpr int CPW = (64 + FPC - 1) / FPC;
    static_assert(FPC * CIN * 4 <= 2048 && (FPC * CIN * 4) % 16 == 0 && 2 * CPW <= 2 * NC8 + 2 * NS8, "staging geometry");
    // ---- shared region
    static constexpr int O_ZERO = G * GRP;
    // weights: K-major B operands [tap][chunk8][2 N rows: hi rows then lo rows][8 x bf16]
    static constexpr int W1B = KT1 * NX8E * 2 * N1 * 16;
    static constexpr int O_W1 = O_ZERO + PL;
    static constexpr int W2B = INS ? 3 * NH8E * 2 * 16 * 16 : 0;
    static constexpr int O_W2 = O_W1 + W1B;
    static constexpr int W2DB = INS ? 3 * NC8 * 2 * NH * 16 : 0;
    static constexpr int O_W2D = O_W2 + W2B;
    static constexpr int WBB = 3 * NC8 * 2 * 16 * 16;
    static constexpr int O_WB = O_W2D + W2DB;
    static constexpr int WBDB = 3 * NS8 * 2 * 16 * 16;
    static constexpr int O_WBD = O_WB + WBB;
    static constexpr int O_ID = O_WBD + WBDB;                            // identity [chunk8 (4)][K half (2)][32 rows][8] bf16
    static constexpr int O_F32 = O_ID + 4096;                            // b1[32] b2[16] lng[16] lnb[16] bb[16] hw[4*128] hb[4]
    static constexpr int F_B1 = 0, F_B2 = 32, F_LNG = 48, F_LNB = 64, F_BB = 80, F_HW = 96, F_HB = 96 + 512, F_END = 96 + 512 + 8;
    static constexpr int O_BAR = O_F32 + F_END * 4;                      // per group 16 mbarriers
    static constexpr int O_Q = O_BAR + G * 16 * 8 + 16;                  // tile queue: [next slot][tile id of group g, parity p]
    static constexpr int O_END = O_Q + 16 + G * 8;
    // M = 128 MN-major operands read 16 planes from their start: keep that inside the allocation
    static constexpr int SPAN = (G - 1) * GRP + (P_F + NC8 + 17) * PL;
    static constexpr int TOTAL = ((O_END > SPAN ? O_END : SPAN) + 127) / 128 * 128;
    // ---- TMEM columns
    // per group 32 working columns; then the persistent weight-gradient blocks (M = 64: [hi | lo] input planes on the
    // lanes, 32 columns = [hi | lo] output-gradient planes, one block per tap) and the per-row bias sums
    static constexpr int C_W1 = G * 32;
    static constexpr int C_W2 = C_W1 + 32 * KT1;
    static constexpr int C_WB = C_W2 + (INS ? 96 : 0);
    static constexpr int C_B1 = C_WB + 96;
    static constexpr int C_B2 = C_B1 + N1;
    static constexpr int C_BB = C_B2 + (INS ? 16 : 0);
    static constexpr int C_END = C_BB + 16;
    static_assert(C_END <= 512, "TMEM columns");
    static constexpr int NTH = (RG + 1) * 128;                           // RG row warpgroups + one service warpgroup
    static constexpr int REG_SERVICE = G == 3 ? 80 : 64;
    static constexpr int REG_LAUNCH = (65536 / NTH) / 8 * 8;
    static constexpr int REG_ROW_ = ((NTH * REG_LAUNCH - 128 * REG_SERVICE) / (RG * 128)) / 8 * 8;
    static constexpr int REG_ROW = REG_ROW_ > 232 ? 232 : REG_ROW_;
};

namespace ws {
__device__ __forceinline__ uint32_t cvt_bf16x2(float hi, float lo) {          // {bf16(hi) : bf16(lo)}
    uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r;
}
// 8 fp32 -> 8 bf16 hi + 8 bf16 lo (x = hi + lo up to 2^-16)
__device__ __forceinline__ void split8(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        const float x0 = v[2 * p], x1 = v[2 * p + 1];
        const uint32_t Hh = cvt_bf16x2(x1, x0);
        const float h0 = __uint_as_float(Hh << 16), h1 = __uint_as_float(Hh & 0xffff0000u);
        h[p] = Hh; l[p] = cvt_bf16x2(x1 - h1, x0 - h0);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]); lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// row `row` of an operand buffer [part][chunk8] planes: N (multiple of 8) values
template <int N, int PL>
__device__ __forceinline__ void store_split(uint8_t* buf, int row, const float (&v)[N]) {
    constexpr int N8 = N / 8;
#pragma unroll
    for (int c = 0; c < N8; ++c) {
        uint4 hi, lo; split8(&v[8 * c], hi, lo);
        *reinterpret_cast<uint4*>(buf + c * PL + row * 16) = hi;
        *reinterpret_cast<uint4*>(buf + (N8 + c) * PL + row * 16) = lo;
    }
}
// GELU derivatives (values in (-0.2, 1.2)) saved as fp16
template <int N, int PL>
__device__ __forceinline__ void store_half(uint8_t* buf, int row, const float (&v)[N]) {
    static_assert(N % 4 == 0, "");
    uint32_t w[N / 2];
#pragma unroll
    for (int i = 0; i < N / 2; ++i) { const __half2 h = __floats2half2_rn(v[2 * i], v[2 * i + 1]); w[i] = *reinterpret_cast<const uint32_t*>(&h); }
#pragma unroll
    for (int c = 0; c < N / 8; ++c) *reinterpret_cast<uint4*>(buf + c * PL + row * 16) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
    if constexpr (N % 8 != 0) *reinterpret_cast<uint2*>(buf + (N / 8) * PL + row * 16) = make_uint2(w[N / 2 - 2], w[N / 2 - 1]);
}
template <int N, int PL>
__device__ __forceinline__ void load_half(const uint8_t* buf, int row, float (&v)[N]) {
    uint32_t w[N / 2];
#pragma unroll
    for (int c = 0; c < N / 8; ++c) {
        const uint4 q = *reinterpret_cast<const uint4*>(buf + c * PL + row * 16);
        w[4 * c] = q.x; w[4 * c + 1] = q.y; w[4 * c + 2] = q.z; w[4 * c + 3] = q.w;
    }
    if constexpr (N % 8 != 0) { const uint2 q = *reinterpret_cast<const uint2*>(buf + (N / 8) * PL + row * 16); w[N / 2 - 2] = q.x; w[N / 2 - 1] = q.y; }
#pragma unroll
    for (int i = 0; i < N / 2; ++i) { const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i])); v[2 * i] = f.x; v[2 * i + 1] = f.y; }
}
// accumulator of a MERGE convolution: columns [0,16) + [16,32)
__device__ __forceinline__ void ld_merged16(uint32_t taddr, float (&v)[16]) {
    float t[16];
    umma::ld_x16(taddr, v); umma::ld_x16(taddr + 16, t); umma::ld_wait();
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] += t[i];
}
__device__ __forceinline__ void st_zero_x8(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
template <int N, class F> __device__ __forceinline__ void static_for(F&& f) {          // f(integral_constant<0>) ... f(integral_constant<N - 1>)
    if constexpr (N > 0) { static_for<N - 1>(f); f(std::integral_constant<int, N - 1>{}); }
}
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
template <int R> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R)); }
template <int R> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R)); }

// ---- MMA issue (one thread).  All descriptors: no swizzle; K-major SBO = 128 B, MN-major LBO = 128 B.
// A tcgen05.mma of this size costs ~45 SM clocks whatever M (64 / 128) and N (<= 32; 48 clocks at N = 64) are
// (scratch/umma_bench.py, profiles/r2_umma_bench.log): the design minimises the NUMBER of MMAs.
// convolution: D = sum_tap A(shifted by (tap - KT/2) * 2 rows) * W_tap over N8 chunks of 8 channels.
//   MERGE (2 N <= 32): two MMAs per (tap, K step): A_hi * [W_hi | W_lo] into columns [0, 2N) and A_lo * W_hi into [0, N); the
//   epilogue adds the two column blocks.  Otherwise three passes hi*hi + lo*hi + hi*lo into columns [0, N).
template <int KT, int N8, int N, bool MERGE, int HALO, int PL>
__device__ __forceinline__ void issue_conv(bool leader, uint32_t d, uint32_t a_hi, uint32_t zero, uint32_t w) {
    constexpr int N8E = (N8 + 1) / 2 * 2;
    constexpr uint32_t idesc1 = umma::make_idesc_bf16(128, N, false, false), idesc2 = umma::make_idesc_bf16(128, 2 * N, false, false);
    uint32_t acc = 0u;
#pragma unroll
    for (int pass = 0; pass < (MERGE ? 2 : 3); ++pass) {
        const uint32_t a = a_hi + (pass == 1 ? N8 * PL : 0);
#pragma unroll
        for (int tap = 0; tap < KT; ++tap)
#pragma unroll
            for (int ks = 0; ks < N8E / 2; ++ks) {
                const uint32_t a0 = a + (uint32_t)((2 * ks) * PL + (HALO + (tap - KT / 2) * 2) * 16);
                const uint32_t lbo = (2 * ks + 1 < N8) ? (uint32_t)PL : zero - (a + (uint32_t)((2 * ks) * PL));
                const uint32_t b0 = w + (uint32_t)(((tap * N8E + 2 * ks) * 2 * N + (pass == 2 ? N : 0)) * 16);
                if (leader)
                    umma::mma_bf16(d, umma::make_desc(a0, lbo, 128u), umma::make_desc(b0, (uint32_t)(2 * N) * 16u, 128u),
                                   (MERGE && pass == 0) ? idesc2 : idesc1, acc);
                acc = 1u;
            }
    }
}
// weight gradient, ONE MMA per (tap, 16 rows): D_tap[m][n] += sum_r A[r (+ shift)][m] * B[r (+ shift)][n] with both operands
// MN-major; the M = 64 rows are A's planes [hi chunks | lo chunks] (8 channels each), the N = 32 columns B's planes
// [hi | lo]: all four hi/lo products at once, added up when the accumulator is read back.
template <int KT, bool SHIFT_A, int HALO, int PL>
__device__ __forceinline__ void issue_wgrad(bool leader, uint32_t d0, uint32_t a_hi, uint32_t b_hi) {
    constexpr uint32_t idesc = umma::make_idesc_bf16(64, 32, true, true);
#pragma unroll
    for (int tap = 0; tap < KT; ++tap) {
        const int sh = (tap - KT / 2) * 2;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const uint32_t a0 = a_hi + (uint32_t)((HALO + 16 * ks + (SHIFT_A ? sh : 0)) * 16);
            const uint32_t b0 = b_hi + (uint32_t)((HALO + 16 * ks + (SHIFT_A ? 0 : sh)) * 16);
            if (leader)
                umma::mma_bf16(d0 + (uint32_t)(tap * 32), umma::make_desc(a0, 128u, (uint32_t)PL), umma::make_desc(b0, 128u, (uint32_t)PL), idesc, 1u);
        }
    }
}
// the same, one tap (8 MMAs): the unit in which the MMA thread interleaves gradient work with latency-critical jobs
template <int KT, bool SHIFT_A, int HALO, int PL>
__device__ __forceinline__ void issue_wgrad_tap(uint32_t d0, uint32_t a_hi, uint32_t b_hi, int tap) {
    constexpr uint32_t idesc = umma::make_idesc_bf16(64, 32, true, true);
    const int sh = (tap - KT / 2) * 2;
    const uint32_t a1 = a_hi + (uint32_t)((HALO + (SHIFT_A ? sh : 0)) * 16), b1 = b_hi + (uint32_t)((HALO + (SHIFT_A ? 0 : sh)) * 16);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)
        umma::mma_bf16(d0 + (uint32_t)(tap * 32), umma::make_desc(a1 + (uint32_t)(256 * ks), 128u, (uint32_t)PL),
                       umma::make_desc(b1 + (uint32_t)(256 * ks), 128u, (uint32_t)PL), idesc, 1u);
}
// per-row running sums: D[r][n] += sum_k (A_hi + A_lo)[r][k] * I[k][n]  (bias gradients; reduced over rows at the very end).
// One MMA per chunk of 8 channels: its K = 16 step pairs the chunk's hi plane with its lo plane (LBO = N8 planes) against
// the identity block of that chunk, stored twice (one copy per K half).
template <int N8, int N, int HALO, int PL>
__device__ __forceinline__ void issue_ident(bool leader, uint32_t d, uint32_t a_hi, uint32_t idm) {
    constexpr uint32_t idesc = umma::make_idesc_bf16(128, N, false, false);
#pragma unroll
    for (int c = 0; c < N8; ++c)
        if (leader) umma::mma_bf16(d, umma::make_desc(a_hi + (uint32_t)(c * PL + HALO * 16), (uint32_t)(N8 * PL), 128u),
                       umma::make_desc(idm + (uint32_t)(c * 1024), 512u, 128u), idesc, 1u);
}

// plain linear head + margin / scale cross entropy for ONE window by one warp (lane l owns features l, l+32, l+64, l+96)
template <int K>
struct HeadLite {
    float g_hw[K][4], g_hb[K], acc_loss, acc_correct;
    __device__ __forceinline__ void zero() {
        acc_loss = 0.f; acc_correct = 0.f;
#pragma unroll
        for (int k = 0; k < K; ++k) { g_hb[k] = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) g_hw[k][i] = 0.f; }
    }
    __device__ __forceinline__ void run(const StreamArgs& A, const float* Ps, float* DPs, const float* hws, const float* hbs, int lane,
                                        int wi, bool train, float inv_denom, int y) {
        float f[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) f[i] = Ps[lane + 32 * i];
        float logit[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            float d = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) d = fmaf(f[i], hws[k * 128 + lane + 32 * i], d);
            logit[k] = d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int k = 0; k < K; ++k) logit[k] += __shfl_xor_sync(0xffffffffu, logit[k], o);
#pragma unroll
        for (int k = 0; k < K; ++k) logit[k] += hbs[k];
        const bool livew = wi < A.B;
        if (livew && A.logits && lane < K) {
            float v = logit[0];
#pragma unroll
            for (int k = 1; k < K; ++k) v = lane == k ? logit[k] : v;
            A.logits[(size_t)wi * K + lane] = v;
        }
        if (!train) return;
        float dl[K];
#pragma unroll
        for (int k = 0; k < K; ++k) dl[k] = 0.f;
        if (livew) {
            if (A.mode == MODE_FUSED) {
                float zz[K]; float mx = -INFINITY; int am = 0; float best = -INFINITY;
#pragma unroll
                for (int k = 0; k < K; ++k) {
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
                for (int k = 0; k < K; ++k) se += __expf(zz[k] - mx);
                const float lse = mx + __logf(se);
                float zy = 0.f, wy = 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) if (k == y) { zy = zz[k]; wy = A.cls_w[k]; }
                acc_loss += wy * (lse - zy) * inv_denom;
                acc_correct += (am == y) ? 1.f : 0.f;
#pragma unroll
                for (int k = 0; k < K; ++k) dl[k] = A.scale * wy * inv_denom * (__expf(zz[k] - lse) - (k == y ? 1.f : 0.f));
            } else {
#pragma unroll
                for (int k = 0; k < K; ++k) dl[k] = A.dlogits_ext[(size_t)wi * K + k];
            }
        }
        float dxn[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int k = 0; k < K; ++k) {
            g_hb[k] += dl[k];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                dxn[i] = fmaf(dl[k], hws[k * 128 + lane + 32 * i], dxn[i]);
                g_hw[k][i] = fmaf(dl[k], f[i], g_hw[k][i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) DPs[lane + 32 * i] = dxn[i] * 0.125f;          // 8 frames per pooling bin
    }
};
}  // namespace ws

#ifdef GAITK_WS_TIMING          // phase clocks of group 0 in CTA 0 (row thread 0 / issuing lane), printed at the end
#define WT(i) do { if (wt_on) { const long long t_ = clock64(); wt_acc[i] += t_ - wt_last; wt_last = t_; } } while (0)
#else
#define WT(i) do { } while (0)
#endif

template <class Cfg, int G, int K, int SL = 1>
__global__ void __launch_bounds__(WsLayout<Cfg, G, SL>::NTH, 1) stream_kernel_ws(const StreamArgs A) {
#ifdef GAITK_WS_TIMING
    long long wt_acc[16], wt_last = 0; bool wt_on = false;
#pragma unroll
    for (int i = 0; i < 16; ++i) wt_acc[i] = 0;
#endif
    using L = WsLayout<Cfg, G, SL>;
    constexpr int RG = L::RG;
#ifdef GAITK_WS_TIMING
    const long long wt_k0 = clock64();
#endif
    constexpr int CIN = L::CIN, KT1 = L::KT1, H = L::H, C = L::C, HALO = L::HALO, PL = L::PL;
    constexpr bool INS = L::INS;
    constexpr int NX8 = L::NX8, NX8E = L::NX8E, NH8 = L::NH8, NH8E = L::NH8E, NC8 = L::NC8, NS8 = L::NS8, N1 = L::N1, O1 = L::O1, NH = L::NH;
    constexpr int NPH = INS ? 6 : 4;                       // operand-ready points per training tile
    extern __shared__ __align__(1024) uint8_t smw[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);        // warp-uniform for the compiler too (uniform datapath)
    const bool train = A.mode != MODE_FWD;
    float* f32 = reinterpret_cast<float*>(smw + L::O_F32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smw + L::O_BAR);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smw + L::O_BAR + G * 16 * 8);
    auto bar_rdy = [&](int g, int k) { return bars + g * 16 + k; };
    auto bar_done = [&](int g) { return bars + g * 16 + 8; };
    auto bar_free = [&](int g) { return bars + g * 16 + 9; };
    auto bar_wfree = [&](int g) { return bars + g * 16 + 10; };
    auto bar_ld = [&](int g) { return bars + g * 16 + 11; };
    // Work distribution inside the CTA: the CTA owns the tiles (it * gridDim + blockIdx) * G + q, q < G, it < nit, as a QUEUE of
    // slots j = it * G + q.  A group's service warp takes the next slot when it starts the load of the group's next tile and
    // publishes the tile id next to the load barrier; the row warps read it there.  (Static ownership -- slot q belongs to group q --
    // let the groups the warp arbiter favours run ~9 tiles ahead: the last quarter of every kernel ran with one or two groups.)
    int* q_next = reinterpret_cast<int*>(smw + L::O_Q);
    auto tile_slot = [&](int g, uint32_t par) { return reinterpret_cast<volatile int*>(smw + L::O_Q + 16) + g * 2 + par; };

    // ------------------------------------------------------------------------------------------ one-time setup
    for (int i = tid; i < L::TOTAL / 16; i += L::NTH) reinterpret_cast<uint4*>(smw)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (tid == 0) {
        for (int g = 0; g < G; ++g) {
            for (int k = 0; k < WS_NPH; ++k) umma::mbar_init(bar_rdy(g, k), 4);      // one arrival per row warp
            umma::mbar_init(bar_done(g), 1); umma::mbar_init(bar_free(g), 1); umma::mbar_init(bar_wfree(g), 1); umma::mbar_init(bar_ld(g), 1);
        }
        umma::fence_mbar_init();
    }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    {   // weights as bf16 (hi, lo) K-major B operands [part][tap][chunk8][n][8]
        // (block = tap * chunks + chunk8, n, k & 7) -> hi at row n, lo at row N + n of the block's 2 N rows
        auto put = [&](int off, int N, int block, int n, int k7, float w) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
            reinterpret_cast<__nv_bfloat16*>(smw + off)[(block * 2 * N + n) * 8 + k7] = hi;
            reinterpret_cast<__nv_bfloat16*>(smw + off)[(block * 2 * N + N + n) * 8 + k7] = lo;
        };
        for (int i = tid; i < O1 * CIN * KT1; i += L::NTH) {
            const int o = i / (CIN * KT1), ci = (i / KT1) % CIN, tap = i % KT1;
            put(L::O_W1, N1, tap * NX8E + (ci >> 3), o, ci & 7, A.w1[i]);
        }
        for (int i = tid; i < O1; i += L::NTH) f32[L::F_B1 + i] = A.b1[i];
        if constexpr (INS) {
            for (int i = tid; i < C * H * 3; i += L::NTH) {
                const int o = i / (H * 3), ci = (i / 3) % H, tap = i % 3;
                float w = A.w2[i];
                if (tap == 1) w += A.skip_identity ? (o == ci ? 1.f : 0.f) : A.wsk[o * H + ci];
                put(L::O_W2, 16, tap * NH8E + (ci >> 3), o, ci & 7, w);                 // fwd: n = o, k = ci
                put(L::O_W2D, NH, (2 - tap) * NC8 + (o >> 3), ci, o & 7, w);            // dgrad: n = ci, k = o, flipped taps
            }
            for (int i = tid; i < C; i += L::NTH) f32[L::F_B2 + i] = A.b2[i] + (A.skip_identity ? 0.f : A.bsk[i]);
        }
        for (int i = tid; i < C; i += L::NTH) { f32[L::F_LNG + i] = A.lng[i]; f32[L::F_LNB + i] = A.lnb[i]; }
        for (int i = tid; i < 16 * C * 3; i += L::NTH) {
            const int o = i / (C * 3), ci = (i / 3) % C, tap = i % 3;
            put(L::O_WB, 16, tap * NC8 + (ci >> 3), o, ci & 7, A.wbb[i]);
            put(L::O_WBD, 16, (2 - tap) * NS8 + (o >> 3), ci, o & 7, A.wbb[i]);
        }
        for (int i = tid; i < 16; i += L::NTH) f32[L::F_BB + i] = A.bbb[i];
        for (int i = tid; i < K * 128; i += L::NTH) f32[L::F_HW + i] = A.hw[i];
        if (A.hb) for (int i = tid; i < K; i += L::NTH) f32[L::F_HB + i] = A.hb[i];
        for (int i = tid; i < 64; i += L::NTH) {           // identity: element (n, k), n = k = i & 31, in both K halves of chunk k >> 3
            const int k = i & 31, h = i >> 5;
            reinterpret_cast<uint16_t*>(smw + L::O_ID)[(((k >> 3) * 2 + h) * 32 + k) * 8 + (k & 7)] = 0x3F80;
        }
    }
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);
    if (warp < 4) {                                        // persistent accumulator columns start at zero
        for (int c = L::C_W1; c < L::C_END; c += 8) ws::st_zero_x8(tmem + ((uint32_t)(warp * 32) << 16) + c);
        ws::st_wait();
    }
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();

    const uint32_t sbase = umma::smem_u32(smw);
    const int ntiles = (A.B + 1) / 2;
    const int per_it = (int)gridDim.x * G;
    const int nit = (ntiles + per_it - 1) / per_it;
    constexpr int ROW_WARPS = 4 * RG;

#ifdef GAITK_WS_TIMING
    const long long wt_k1 = clock64();
#endif
    if (warp >= ROW_WARPS) {                               // the LAST warpgroup: the hardware arbiter prefers high warp ids, so an
                                                           // MMA / TMA issue never queues behind the row warps' epilogue instructions
        // ====================================================================================== service warpgroup
        ws::reg_dec<L::REG_SERVICE>();
        const int sw = warp - ROW_WARPS;
        const uint32_t zero = sbase + L::O_ZERO, idm = sbase + L::O_ID;
        // Service warp s is the MMA + TMA warp of group s: lane 0 issues all of that group's MMAs in the group's natural
        // order with blocking (hardware-suspended) mbarrier waits, the whole warp stages the next tile's windows.  Issuing a
        // small tcgen05.mma costs the issuing thread ~45 clocks (scratch/umma_bench.py) although the tensor pipe executes it
        // faster: G issuers in parallel keep the pipe fed (one issuer for all groups was 3x slower, profiles/r2_ws_history.md).
        // The groups drift apart in phase, so one group's MMAs run under the other groups' epilogues.  (The order in which
        // different groups' weight-gradient MMAs accumulate into the shared TMEM blocks depends on timing: sums are
        // reproducible to fp32 rounding, not bit for bit.)
        if (sw < G) {
            const int g = sw;
            const uint32_t gb = sbase + g * L::GRP, acc = tmem + g * 32;
            uint8_t* gbp = smw + g * L::GRP;
            const int per_win = 64 * CIN;
            // window w's source address for the NEXT load, held by lane w (its win_start entry is read one tile ahead: no
            // global-load latency on the issuing path); tile_next = the tile those addresses belong to (-1: the queue is empty)
            const float* src_next = nullptr;
            int tile_next = -1;
            auto grab = [&]() {                            // whole warp: next slot of the CTA's queue -> tile id or -1
                int j = 0;
                if (lane == 0) j = atomicAdd(q_next, 1);
                j = __shfl_sync(0xffffffffu, j, 0);
                const int it_ = j / G, q_ = j - it_ * G;
                const int tile = (it_ * (int)gridDim.x + (int)blockIdx.x) * G + q_;
                tile_next = (j < nit * G && tile < ntiles) ? tile : -1;
                if (lane < 2) {
                    const int wi = tile_next * 2 + lane;
                    src_next = (tile_next >= 0 && wi < A.B && !A.zero_input)
                                   ? A.x + (A.win_start ? (size_t)A.win_start[wi] * CIN : (size_t)wi * per_win) : nullptr;
                }
            };
            // whole warp: windows of tile_next -> staging (inside F|Z) and its id -> tile_slot(g, par); then takes the following slot.
            // With an empty queue it publishes -1 (the row warps leave their loop).  Returns the tile it loaded.
            auto load_tile = [&](uint32_t par) {
                const int tile = tile_next;
                const float* src0 = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)src_next, 0);
                const float* src1 = (const float*)__shfl_sync(0xffffffffu, (unsigned long long)src_next, 1);
                if (tile >= 0) grab();
                uint32_t bytes = 0;
#pragma unroll
                for (int w = 0; w < 2; ++w) {
                    const float* src = w ? src1 : src0;
                    if (!src) continue;
                    if ((reinterpret_cast<uintptr_t>(src) & 15) == 0) { bytes += (uint32_t)per_win * 4u; continue; }
                    for (int e = lane; e < per_win; e += 32) {            // unaligned window: plain copy into the same staging layout
                        const int t = e / CIN, j = t / L::FPC;
                        reinterpret_cast<float*>(gbp + (L::P_F + w * L::CPW + j) * PL + HALO * 16)[e - j * L::FPC * CIN] = __ldg(src + e);
                    }
                }
                __syncwarp();
                WT(11);
                if (lane == 0) {
                    *tile_slot(g, par) = tile;             // ordered before the arrival below (release), read after the rows' wait (acquire)
                    if (bytes == 0) {
                        umma::mbar_arrive(bar_ld(g));
                    } else {
                        // (no proxy fence here: the row threads fenced their generic writes to F / Z before the arrivals this
                        // thread has observed, and the MMAs that read them have completed)
                        umma::mbar_expect_tx(bar_ld(g), bytes);
                        WT(12);
#pragma unroll
                        for (int w = 0; w < 2; ++w) {
                            const float* src = w ? src1 : src0;
                            if (!src || (reinterpret_cast<uintptr_t>(src) & 15) != 0) continue;
#pragma unroll
                            for (int j = 0; j < L::CPW; ++j) {
                                const int fr = (j + 1) * L::FPC <= 64 ? L::FPC : 64 - j * L::FPC;
                                umma::bulk_g2s(gbp + (L::P_F + w * L::CPW + j) * PL + HALO * 16, src + j * L::FPC * CIN, (uint32_t)(fr * CIN * 4), bar_ld(g));
                            }
                        }
                        WT(13);
                    }
                }
                __syncwarp();
                return tile;
            };
            const bool leader = umma::elect_one();
            auto wait_rdy = [&](int k, uint32_t par) { umma::mbar_wait(bar_rdy(g, k), par); umma::fence_after_sync(); };
            auto commit = [&](uint64_t* bar) { if (leader) umma::commit(bar); };
            grab();
            int cur = load_tile(0u);
#ifdef GAITK_WS_TIMING
            wt_on = blockIdx.x == 0 && g == 0 && lane == 0; wt_last = clock64();
#endif
            for (int it = 0; cur >= 0; ++it) {             // `it` counts THIS group's tiles (barrier parities are per group)
                const uint32_t par = it & 1;
                int k = 0;
                wait_rdy(k++, par);
                WT(0);
                ws::issue_conv<KT1, NX8, N1, (N1 <= 16), HALO, PL>(leader, acc, gb + L::P_X * PL, zero, sbase + L::O_W1);
                commit(bar_done(g));
                WT(1);
                if constexpr (INS) {
                    wait_rdy(k++, par);
                    ws::issue_conv<3, NH8, 16, true, HALO, PL>(leader, acc, gb + L::P_HA * PL, zero, sbase + L::O_W2);
                    commit(bar_done(g));
                }
                wait_rdy(k++, par);
                WT(2);
                ws::issue_conv<3, NC8, 16, true, HALO, PL>(leader, acc, gb + L::P_F * PL, zero, sbase + L::O_WB);
                commit(bar_done(g));
                WT(3);
                if (train) {
                    wait_rdy(k++, par);                    // dz is in the Z planes
                    WT(4);
                    ws::issue_conv<3, NS8, 16, true, HALO, PL>(leader, acc, gb + L::P_Z * PL, zero, sbase + L::O_WBD);
                    commit(bar_done(g));
                    WT(5);
                    ws::issue_ident<NS8, 16, HALO, PL>(leader, tmem + L::C_BB, gb + L::P_Z * PL, idm);
                    ws::issue_wgrad<3, true, HALO, PL>(leader, tmem + L::C_WB, gb + L::P_F * PL, gb + L::P_Z * PL);
                    commit(bar_free(g));
                    WT(6);
                } else {
                    commit(bar_free(g));
                }
                // this warp's backbone MMAs (data gradient included) have read F and Z: the staging area inside them is free
                umma::mbar_wait(bar_free(g), par);
                WT(7);
                cur = load_tile(par ^ 1u);                 // the group's next tile, or the end-of-queue mark
                WT(8);
                if (train) {
                    wait_rdy(INS ? 4 : 3, par);            // dA is in the XH planes
                    WT(9);
                    if constexpr (INS) {
                        ws::issue_conv<3, NC8, NH, false, HALO, PL>(leader, acc, gb + L::P_XH * PL, zero, sbase + L::O_W2D);
                        commit(bar_done(g));
                        ws::issue_wgrad<3, true, HALO, PL>(leader, tmem + L::C_W2, gb + L::P_HA * PL, gb + L::P_XH * PL);
                        ws::issue_ident<NC8, 16, HALO, PL>(leader, tmem + L::C_B2, gb + L::P_XH * PL, idm);
                        commit(bar_wfree(g));
                        wait_rdy(5, par);                  // dA1 is in the HA planes: lanes = conv1 output channel, columns = input channel
                        ws::issue_ident<NH8, N1, HALO, PL>(leader, tmem + L::C_B1, gb + L::P_HA * PL, idm);
                        ws::issue_wgrad<KT1, false, HALO, PL>(leader, tmem + L::C_W1, gb + L::P_HA * PL, gb + L::P_X * PL);
                    } else {
                        ws::issue_ident<NC8, 16, HALO, PL>(leader, tmem + L::C_B1, gb + L::P_XH * PL, idm);
                        ws::issue_wgrad<KT1, true, HALO, PL>(leader, tmem + L::C_W1, gb + L::P_X * PL, gb + L::P_XH * PL);
                    }
                    commit(bar_done(g));
                    WT(10);
                }
            }
#ifdef GAITK_WS_TIMING
            if (wt_on) printf("ISSUER CIN%d nit %d: rdy0 %lld conv1 %lld | rdy1 %lld bb %lld | rdyZ %lld dgrad %lld wgradbb %lld | free %lld load %lld rdydA %lld wgrad1(+conv2 bwd) %lld || load: pre %lld expect %lld bulk %lld\n", CIN, nit,
                              wt_acc[0] / nit, wt_acc[1] / nit, wt_acc[2] / nit, wt_acc[3] / nit, wt_acc[4] / nit, wt_acc[5] / nit, wt_acc[6] / nit, wt_acc[7] / nit, wt_acc[8] / nit, wt_acc[9] / nit, wt_acc[10] / nit,
                              wt_acc[11] / nit, wt_acc[12] / nit, wt_acc[13] / nit);
#endif
        }
        umma::fence_before_sync();
        asm volatile("bar.sync 0;" ::: "memory");
        asm volatile("bar.sync 0;" ::: "memory");
    } else {
        // ====================================================================================== row warps: thread = row of its group's tile
        ws::reg_inc<L::REG_ROW>();
        float g_lng[16], g_lnb[16];                        // per-row sums of dF * xh and dF (LayerNorm affine gradients)
        ws::HeadLite<K> head;
#pragma unroll
        for (int i = 0; i < 16; ++i) { g_lng[i] = 0.f; g_lnb[i] = 0.f; }
        head.zero();
        const int rw = warp;                               // row warp index
        const int rg = rw >> 2, wq = rw & 3, r = wq * 32 + lane;      // row warpgroup, lane quarter, row of the tile
        const float inv_denom = (A.mode == MODE_FUSED) ? 1.0f / A.denom[0] : 0.f;
        const float* b1s = f32 + L::F_B1; const float* b2s = f32 + L::F_B2; const float* lngs = f32 + L::F_LNG; const float* lnbs = f32 + L::F_LNB;
        const float* bbs = f32 + L::F_BB;
        const int t = r >> 1, w = r & 1;
        // per-slot state: slot s of this row warpgroup <-> group g = rg * SL + s (own planes, barriers, TMEM columns, service warp)
        float rstd_row[SL]; uint32_t zmask[SL], dph[SL]; int ylab[SL], cur_tile[SL];
#pragma unroll
        for (int q = 0; q < SL; ++q) { rstd_row[q] = 0.f; zmask[q] = 0u; dph[q] = 0u; ylab[q] = 0; cur_tile[q] = 0; }
        // One phase of one slot.  Every phase ends with the arrival that lets the slot's service warp issue the next MMAs, and the
        // thread moves on to the SAME phase of its next slot: that slot's accumulator has been computed meanwhile.
        constexpr int NHP = INS ? NH8 * 8 : 8;              // padded conv1 width (the insole-only phases are never run otherwise)
        auto phase = [&](auto P_, auto S_, int it) {
            constexpr int P = decltype(P_)::value, sl = decltype(S_)::value;
            const int g = rg * SL + sl;
            uint8_t* gb = smw + g * L::GRP;
            const uint32_t trow = tmem + ((uint32_t)(wq * 32) << 16) + g * 32;
            float* Ps = reinterpret_cast<float*>(gb + L::O_P); float* DPs = reinterpret_cast<float*>(gb + L::O_DP);
            const uint32_t par = it & 1;
            if (P > 0 && cur_tile[sl] < 0) return;         // this slot's group has drained the queue
            int win0 = cur_tile[sl] * 2;
            // every thread orders its operand stores before the async proxy, the warp converges, ONE lane arrives
            auto arrive = [&](int k) {
                umma::fence_smem_to_async(); umma::fence_before_sync();
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(bar_rdy(g, k));
            };
            auto wait_done = [&]() { umma::mbar_wait(bar_done(g), dph[sl]); dph[sl] ^= 1u; umma::fence_after_sync(); };
            if constexpr (P == 0) {
                if (cur_tile[sl] < 0) return;
                // ------------------------------------------------ staged window bytes -> X planes (hi, lo)
                umma::mbar_wait(bar_ld(g), par);
                cur_tile[sl] = *tile_slot(g, par);         // published by the service warp before the load barrier completed
                if (cur_tile[sl] < 0) return;
                win0 = cur_tile[sl] * 2;
                // the head warps fetch their window's label now; it is consumed a few thousand clocks later
                ylab[sl] = 0;
                if (wq < 2 && A.mode == MODE_FUSED && win0 + wq < A.B) ylab[sl] = (int)A.y[win0 + wq];
                if (!A.zero_input) {
                    const bool live = win0 + w < A.B;
                    const int j = t / L::FPC;
                    const float* src = reinterpret_cast<const float*>(gb + (L::P_F + w * L::CPW + j) * PL + HALO * 16) + (t - j * L::FPC) * CIN;
                    float v[NX8 * 8];
                    if constexpr (CIN % 4 == 0) {
#pragma unroll
                        for (int c4 = 0; c4 < CIN / 4; ++c4) {
                            const float4 q = reinterpret_cast<const float4*>(src)[c4];
                            v[4 * c4] = q.x; v[4 * c4 + 1] = q.y; v[4 * c4 + 2] = q.z; v[4 * c4 + 3] = q.w;
                        }
                    } else if constexpr (CIN == 2) {
                        const float2 q = *reinterpret_cast<const float2*>(src); v[0] = q.x; v[1] = q.y;
                    } else {
#pragma unroll
                        for (int c = 0; c < CIN; ++c) v[c] = src[c];
                    }
#pragma unroll
                    for (int c = 0; c < NX8 * 8; ++c) v[c] = (c < CIN && live) ? v[c] : 0.f;
                    ws::store_split<NX8 * 8, PL>(gb + L::P_X * PL, HALO + r, v);
                }
                arrive(0);
            }
            if constexpr (P == 1 && INS) {
                // ------------------------------------------------ conv1 epilogue: GELU -> HA planes (+ saved derivative)
                wait_done();
                float a1[32], ha[NHP], d1[NHP];
                umma::ld_x16(trow, a1); umma::ld_x8(trow + 16, a1 + 16); umma::ld_wait();
#pragma unroll
                for (int c = 0; c < NHP; ++c) { if (c < H) gelu_fwd_fast(a1[c] + b1s[c], ha[c], d1[c]); else { ha[c] = 0.f; d1[c] = 0.f; } }
                ws::store_split<NHP, PL>(gb + L::P_HA * PL, HALO + r, ha);
                if (train) ws::store_half<NHP, PL>(gb + L::P_D1 * PL, HALO + r, d1);
                arrive(1);
            }
            if constexpr (P == 2) {
                // ------------------------------------------------ encoder epilogue: GELU, LayerNorm -> F planes
                wait_done();
                float a[16], gl[16], d[16], xh[16], f[16]; float rstd;
                ws::ld_merged16(trow, a);
                const float* bias = INS ? b2s : b1s;
#pragma unroll
                for (int c = 0; c < 16; ++c) { if (c < C) gelu_fwd_fast(a[c] + bias[c], gl[c], d[c]); else { gl[c] = 0.f; d[c] = 0.f; } }
                ln_fwd<16, C>(gl, xh, rstd);
#pragma unroll
                for (int c = 0; c < 16; ++c) f[c] = c < C ? fmaf(xh[c], lngs[c], lnbs[c]) : 0.f;
                ws::store_split<16, PL>(gb + L::P_F * PL, HALO + r, f);
                if (train) {
                    float dd[(C + 3) / 4 * 4];
#pragma unroll
                    for (int c = 0; c < (C + 3) / 4 * 4; ++c) dd[c] = d[c];
                    ws::store_half<(C + 3) / 4 * 4, PL>(gb + L::P_D * PL, HALO + r, dd);
                    float4* xp = reinterpret_cast<float4*>(gb + L::P_XH * PL) + (HALO + r);
#pragma unroll
                    for (int c4 = 0; c4 < (C + 3) / 4; ++c4) xp[c4 * L::RB] = make_float4(xh[4 * c4], xh[4 * c4 + 1], xh[4 * c4 + 2], xh[4 * c4 + 3]);
                    rstd_row[sl] = rstd;
                }
                arrive(INS ? 2 : 1);
            }
            if constexpr (P == 3) {
                // ------------------------------------------------ shared backbone forward: ReLU + adaptive pooling (8 frames per bin)
                wait_done();
                uint32_t zm = 0;
                {
                    float z[16];
                    ws::ld_merged16(trow, z);
                    umma::fence_before_sync();
                    float zz[16];
#pragma unroll
                    for (int q = 0; q < 16; ++q) { zz[q] = fmaxf(z[q] + bbs[q], 0.f); zm |= (zz[q] > 0.f ? 1u : 0u) << q; }
                    // the 8 rows of one (window, bin) are the lanes with equal (lane & 1, lane >> 4): reduce-scatter butterfly over lane
                    // bits 3, 2, 1; afterwards lane l holds the sums of channels 2 * ((l >> 1) & 7) and + 1
                    float a8[8], a4[4], a2[2];
                    const bool h8 = lane & 8, h4 = lane & 4, h2 = lane & 2;
#pragma unroll
                    for (int i = 0; i < 8; ++i) { const float keep = h8 ? zz[8 + i] : zz[i], send = h8 ? zz[i] : zz[8 + i]; a8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8); }
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const float keep = h4 ? a8[4 + i] : a8[i], send = h4 ? a8[i] : a8[4 + i]; a4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4); }
#pragma unroll
                    for (int i = 0; i < 2; ++i) { const float keep = h2 ? a4[2 + i] : a4[i], send = h2 ? a4[i] : a4[2 + i]; a2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2); }
                    const int ch = (h8 ? 8 : 0) + (h4 ? 4 : 0) + (h2 ? 2 : 0);
                    *reinterpret_cast<float2*>(Ps + w * 128 + (r >> 4) * 16 + ch) = make_float2(a2[0] * 0.125f, a2[1] * 0.125f);
                }
                zmask[sl] = zm;
                ws::bar_sync(1 + rg, 128);
                // ------------------------------------------------ head + loss: warp q < 2 <-> window q of the tile
                if (wq < 2) head.run(A, Ps + wq * 128, DPs + wq * 128, f32 + L::F_HW, f32 + L::F_HB, lane, win0 + wq, train, inv_denom, ylab[sl]);
                ws::bar_sync(1 + rg, 128);
                if (train) {
                    // ------------------------------------------------ dz through pooling + ReLU -> Z planes
                    float dz[16];
                    const float4* dp = reinterpret_cast<const float4*>(DPs + w * 128 + (t >> 3) * 16);
#pragma unroll
                    for (int s4 = 0; s4 < 4; ++s4) { const float4 q = dp[s4]; dz[4 * s4] = q.x; dz[4 * s4 + 1] = q.y; dz[4 * s4 + 2] = q.z; dz[4 * s4 + 3] = q.w; }
#pragma unroll
                    for (int q = 0; q < 16; ++q) dz[q] = ((zm >> q) & 1u) ? dz[q] : 0.f;
                    ws::store_split<16, PL>(gb + L::P_Z * PL, HALO + r, dz);
                    arrive(INS ? 3 : 2);
                }
            }
            if constexpr (P == 4) {
                // ------------------------------------------------ LayerNorm / GELU backward -> dA (over XH)
                wait_done();
                float df[16], xh[16], dxh[16], dg[16], d[(C + 3) / 4 * 4], da[16];
                ws::ld_merged16(trow, df);
                const float4* xp = reinterpret_cast<const float4*>(gb + L::P_XH * PL) + (HALO + r);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c4 < (C + 3) / 4) q = xp[c4 * L::RB];
                    xh[4 * c4] = q.x; xh[4 * c4 + 1] = q.y; xh[4 * c4 + 2] = q.z; xh[4 * c4 + 3] = q.w;
                }
                ws::load_half<(C + 3) / 4 * 4, PL>(gb + L::P_D * PL, HALO + r, d);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float dfc = c < C ? df[c] : 0.f;
                    g_lng[c] = fmaf(dfc, xh[c], g_lng[c]); g_lnb[c] += dfc;
                    dxh[c] = c < C ? dfc * lngs[c] : 0.f;
                }
                ln_bwd<16, C>(dxh, xh, rstd_row[sl], dg);
#pragma unroll
                for (int c = 0; c < 16; ++c) da[c] = c < C ? dg[c] * d[c] : 0.f;
                ws::store_split<16, PL>(gb + L::P_XH * PL, HALO + r, da);
                arrive(INS ? 4 : 3);
            }
            if constexpr (P == 5 && INS) {
                // ------------------------------------------------ conv2 data gradient -> dA1 (over HA, once conv2's weight gradient has read it)
                wait_done();
                float dh[32], d1[NHP], da1[NHP];
                umma::ld_x16(trow, dh); umma::ld_x8(trow + 16, dh + 16); umma::ld_wait();
                ws::load_half<NHP, PL>(gb + L::P_D1 * PL, HALO + r, d1);
#pragma unroll
                for (int c = 0; c < NHP; ++c) da1[c] = c < H ? dh[c] * d1[c] : 0.f;
                umma::mbar_wait(bar_wfree(g), par);
                ws::store_split<NHP, PL>(gb + L::P_HA * PL, HALO + r, da1);
                arrive(5);
            }
            if constexpr (P == 6) wait_done();             // the first-layer weight gradient has read X and dA: the tile's buffers are free
        };
        auto all_slots = [&](auto P_, int it) { ws::static_for<SL>([&](auto S_) { phase(P_, S_, it); }); };
        for (int it = 0;; ++it) {                          // `it` counts this warpgroup's rounds (barrier parities are per group)
            all_slots(std::integral_constant<int, 0>{}, it);
            bool any = false;
#pragma unroll
            for (int q = 0; q < SL; ++q) any = any || cur_tile[q] >= 0;
            if (!any) break;
            if constexpr (INS) all_slots(std::integral_constant<int, 1>{}, it);
            all_slots(std::integral_constant<int, 2>{}, it);
            all_slots(std::integral_constant<int, 3>{}, it);
            if (!train) continue;
            all_slots(std::integral_constant<int, 4>{}, it);
            if constexpr (INS) all_slots(std::integral_constant<int, 5>{}, it);
            all_slots(std::integral_constant<int, 6>{}, it);
        }
        // -------------------------------------------------------------------------------------- teardown + flush
#ifdef GAITK_WS_TIMING
        const long long wt_k2 = clock64();
#endif
        umma::fence_before_sync();
        asm volatile("bar.sync 0;" ::: "memory");          // every group has seen its last commit: all MMAs are complete
        umma::fence_after_sync();
#ifdef GAITK_WS_TIMING
        const long long wt_k3 = clock64(); long long wt_k4 = 0, wt_k5 = 0, wt_k6 = 0;
#endif
        if (A.mode != MODE_FWD) {
        float* out = A.partial + (size_t)blockIdx.x * A.NGP;
        const GradOff& go = A.go;
        float* stage = reinterpret_cast<float*>(smw);      // the groups' planes are dead
        // per-row LayerNorm affine sums and the head warps' sums: every row warp stages its partial
#pragma unroll
        for (int c = 0; c < 16; ++c) {
            const float s1 = warp_sum(g_lng[c]), s2 = warp_sum(g_lnb[c]);
            if (lane == 0) { stage[rw * 32 + c] = s1; stage[rw * 32 + 16 + c] = s2; }
        }
        float* hstage = stage + ROW_WARPS * 32;            // [head warp][K * 128 + K + 2]
        constexpr int HS = K * 128 + K + 2;
        if (wq < 2) {
            float* hs = hstage + (rg * 2 + wq) * HS;
#pragma unroll
            for (int k = 0; k < K; ++k)
#pragma unroll
                for (int i = 0; i < 4; ++i) hs[k * 128 + lane + 32 * i] = head.g_hw[k][i];
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < K; ++k) hs[K * 128 + k] = head.g_hb[k];
                hs[K * 128 + K] = head.acc_loss; hs[K * 128 + K + 1] = head.acc_correct;
            }
        }
        ws::bar_sync(15, ROW_WARPS * 32);
#ifdef GAITK_WS_TIMING
        wt_k4 = clock64();
#endif
        const int rt = tid;                                // row threads 0 .. 128 G - 1
        if (rt < 2 * C) {
            const int c = rt % C, which = rt / C;
            float s = 0.f;
            for (int q = 0; q < ROW_WARPS; ++q) s += stage[q * 32 + which * 16 + c];
            out[(which ? go.lnb : go.lng) + c] = s;
        }
        for (int e = rt; e < HS; e += ROW_WARPS * 32) {
            float s = 0.f;
            for (int q = 0; q < 2 * RG; ++q) s += hstage[q * HS + e];
            if (e < K * 128) out[go.hw + e] = s;
            else if (e < K * 128 + K) { if (go.hb >= 0) out[go.hb + e - K * 128] = s; }
            else out[go.total + (e - K * 128 - K)] = s;    // [NG] = loss, [NG + 1] = correct
        }
        if (rg == 0) {
            const uint32_t tq = tmem + ((uint32_t)(wq * 32) << 16);
            float* bst = hstage + 2 * RG * HS;             // [4 warps][64] per-row sums
            float* wst = bst + 256;                        // [tap][64 rows][32 columns] one weight-gradient region at a time
            // per-row sums (bias gradients): reduce over the 128 lanes
            auto colsum = [&](int col0, int n, int slot) {
                for (int c0 = 0; c0 < n; c0 += 8) {
                    float v[8];
                    umma::ld_x8(tq + col0 + c0, v); umma::ld_wait();
#pragma unroll
                    for (int c = 0; c < 8; ++c) { const float s = warp_sum(v[c]); if (lane == 0) bst[wq * 64 + slot + c0 + c] = s; }
                }
            };
            colsum(L::C_B1, N1, 0);
            if constexpr (INS) colsum(L::C_B2, 16, 32);
            colsum(L::C_BB, 16, 48);
#ifdef GAITK_WS_TIMING
            wt_k5 = clock64();
#endif
            // weight-gradient regions: M = 64 accumulator rows sit on lanes 0..15 of every lane quarter (row = 16 * quarter + lane);
            // rows = A's planes [hi chunks | lo chunks] x 8 channels, columns = B's [hi 16 | lo 16]: dW = hh + hl + lh + ll
            auto dump = [&](int col0, int KT) {
                for (int tap = 0; tap < KT; ++tap) {
                    float v[16], v2[16];
                    umma::ld_x16(tq + col0 + tap * 32, v); umma::ld_x16(tq + col0 + tap * 32 + 16, v2); umma::ld_wait();
                    if (lane < 16) {
                        float* d = wst + (tap * 64 + wq * 16 + lane) * 32;
#pragma unroll
                        for (int c = 0; c < 16; ++c) { d[c] = v[c]; d[16 + c] = v2[c]; }
                    }
                }
                umma::fence_before_sync();
                ws::bar_sync(14, 128);
            };
            auto wval = [&](int tap, int NA8, int rc, int cc) {
                const float* lo_ = wst + (tap * 64 + (NA8 + (rc >> 3)) * 8 + (rc & 7)) * 32;
                const float* hi_ = wst + (tap * 64 + (rc >> 3) * 8 + (rc & 7)) * 32;
                return (hi_[cc] + hi_[16 + cc]) + (lo_[cc] + lo_[16 + cc]);
            };
            dump(L::C_W1, KT1);
            if constexpr (INS) {                           // rows = output channel (dA1 planes), columns = input channel
                for (int e = rt; e < KT1 * H * CIN; e += 128) {
                    const int tap = e / (H * CIN), co = (e / CIN) % H, ci = e % CIN;
                    out[go.w1 + (co * CIN + ci) * KT1 + tap] = wval(tap, NH8, co, ci);
                }
            } else {                                       // rows = input channel, columns = output channel
                for (int e = rt; e < KT1 * CIN * C; e += 128) {
                    const int tap = e / (CIN * C), ci = (e / C) % CIN, co = e % C;
                    out[go.w1 + (co * CIN + ci) * KT1 + tap] = wval(tap, NX8, ci, co);
                }
            }
            ws::bar_sync(14, 128);
            if constexpr (INS) {
                dump(L::C_W2, 3);
                for (int e = rt; e < 3 * H * C; e += 128) {
                    const int tap = e / (H * C), ci = (e / C) % H, co = e % C;
                    const float v = wval(tap, NH8, ci, co);
                    out[go.w2 + (co * H + ci) * 3 + tap] = v;
                    if (tap == 1 && !A.skip_identity) out[go.wsk + co * H + ci] = v;
                }
                ws::bar_sync(14, 128);
            }
            dump(L::C_WB, 3);
            for (int e = rt; e < 3 * C * 16; e += 128) {
                const int tap = e / (C * 16), ci = (e / 16) % C, co = e % 16;
                out[go.wbb + (co * C + ci) * 3 + tap] = wval(tap, NC8, ci, co);
            }
#ifdef GAITK_WS_TIMING
            wt_k6 = clock64();
#endif
            if (rt < 64) {
                const float s = (bst[rt] + bst[64 + rt]) + (bst[128 + rt] + bst[192 + rt]);
                if (rt < 32) { if (rt < O1) out[go.b1 + rt] = s; }
                else if (rt < 48) { if (INS && rt - 32 < C) { out[go.b2 + rt - 32] = s; if (!A.skip_identity) out[go.bsk + rt - 32] = s; } }
                else out[go.bbb + rt - 48] = s;
            }
        }
        }
        umma::fence_before_sync();
        asm volatile("bar.sync 0;" ::: "memory");
#ifdef GAITK_WS_TIMING
        if (blockIdx.x == 0 && tid == 0)
            printf("KERNEL CIN%d grid %d nit %d: setup %lld | tile loop %lld (%lld per round) | teardown %lld clocks = barrier %lld + stage %lld + sums %lld + dumps %lld + rest %lld\n", CIN, (int)gridDim.x, nit,
                   wt_k1 - wt_k0, wt_k2 - wt_k1, (wt_k2 - wt_k1) / (nit > 0 ? nit : 1), clock64() - wt_k2, wt_k3 - wt_k2, wt_k4 - wt_k3, wt_k5 - wt_k4, wt_k6 - wt_k5, clock64() - wt_k6);
#endif
    }
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

}  // namespace gaitk
