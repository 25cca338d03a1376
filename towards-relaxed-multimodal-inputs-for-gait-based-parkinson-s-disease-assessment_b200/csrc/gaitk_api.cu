// gaitk_api.cu -- libgaitk.so: plan, parameter layout, launch geometry and the C-ABI of include/gaitk.h.
// No torch headers, no allocation or synchronisation inside any entry (CUDA-graph capturable).
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/gaitk.h"
#include "stream_common.cuh"
#include "stream_tc_plan.h"
#include "stream_dispatch.h"
#include "update_kernels.cuh"
#include "umma_selftest.cuh"

using namespace gaitk;

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
extern "C" int gaitk_set_error(int code, const char* fmt, ...) {            // for the other translation units (fusion.cu)
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
    return code;
}
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail((int)e_, "%s: %s", #x, cudaGetErrorString(e_)); } while (0)
#define LAUNCH_CHECK() do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) return fail((int)e_, "kernel launch: %s", cudaGetErrorString(e_)); } while (0)

// ------------------------------------------------------------------------------------------ plan
struct ParamInfo {
    std::string name; long long off; int numel; int group; int dims[4]; int shared_off;
};

struct StreamPlan {
    int enc, CIN, T_in, T, W, rows_in, rows, halo, RBi, RB, pool_sensor;
    int pooled_taps;                                                   // sensor encoder in its pooled-taps form (ENC_POOL_LINEAR kernel)
    int cl;                                                            // thread-block cluster size (long windows split by time), else 1
    int H, C, S, NFL, KT1, skip_identity, PROJ;
    StreamKernelFn fn;
    SmemPlan sp; GradOff go; int NGP; size_t smem_bytes; int ctas_per_sm;
    StreamKernelTcFn fn_tc; TcPlan tp; size_t smem_tc; int ctas_tc;      // tensor-core variant (nullptr when unavailable)
    int tc_threads;                                                    // 128 (one thread per row) or 256 (two per row)
    WsKernel ws; int has_ws;                                           // warp-specialised split-bf16 kernel (GAITK_DTYPE_BF16X3)
    int p_w1, p_b1, p_w2, p_b2, p_wsk, p_bsk, p_lng, p_lnb, p_hng, p_hnb, p_hw, p_hb, p_wp, p_bp;   // param indices (-1 = none)
    int nseg; Seg seg[MAX_SEG];
};

struct gaitk_plan {
    gaitk_model_desc d; int device; int sm_count;
    std::vector<ParamInfo> params; long long NP; int P;
    int n_streams; StreamPlan st[GAITK_MAX_STREAMS];
    int p_wbb, p_bbb;
    int NF;
    // side streams + events: the stream kernels of a small batch (no kernel fills the GPU) run concurrently, forked from and
    // joined back into the caller's stream (capturable: the fork / join become graph edges); created on first use
    cudaStream_t side[GAITK_MAX_STREAMS - 1] = {nullptr, nullptr};
    cudaEvent_t ev_fork = nullptr, ev_join[GAITK_MAX_STREAMS - 1] = {nullptr, nullptr};
    ~gaitk_plan() {
        for (auto& e : ev_join) if (e) cudaEventDestroy(e);
        if (ev_fork) cudaEventDestroy(ev_fork);
        for (auto& q : side) if (q) cudaStreamDestroy(q);
    }
};
static int ensure_side_streams(gaitk_plan* pl) {
    if (pl->ev_fork) return 0;
    for (auto& q : pl->side) CUDA_TRY(cudaStreamCreateWithFlags(&q, cudaStreamNonBlocking));
    for (auto& e : pl->ev_join) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
    return 0;
}

static int add_param(gaitk_plan* pl, const std::string& name, int group, int d0, int d1 = 0, int d2 = 0) {
    ParamInfo p; p.name = name; p.off = pl->NP; p.group = group;
    p.dims[0] = d0; p.dims[1] = d1; p.dims[2] = d2; p.dims[3] = 0;
    p.numel = d0 * (d1 ? d1 : 1) * (d2 ? d2 : 1);
    p.shared_off = -1;
    if (group == 0) { p.shared_off = pl->P; pl->P += p.numel; }
    pl->NP += p.numel;
    pl->params.push_back(p);
    return (int)pl->params.size() - 1;
}

static int round_rb(int rows, int halo) {          // rows per chunk, == 1 (mod 8): conflict-free chunk planes
    int rb = rows + 2 * halo;
    while (rb % 8 != GAITK_RB_MOD) ++rb;
    return rb;
}

static int plan_stream(gaitk_plan* pl, int s, int enc, int CIN, int KT1, int H, int T_in, int T, int pool_sensor) {
    StreamPlan& sp = pl->st[s];
    const gaitk_model_desc& d = pl->d;
    sp.enc = enc; sp.CIN = CIN; sp.KT1 = KT1; sp.H = H; sp.C = d.enc_out_ch; sp.S = d.shared_out_ch;
    sp.T_in = T_in; sp.T = T; sp.pool_sensor = pool_sensor;
    sp.skip_identity = (enc == ENC_INSOLE && H == sp.C);
    sp.PROJ = (d.family == GAITK_FAMILY_WEARGAIT) ? d.reserved[0] : 0;      // proj_ch
    const int NF = d.backbone_dim * d.shared_out_ch;
    if (NF % 32 != 0 || NF > 512) return fail(GAITK_E_SHAPE, "backbone_dim*shared_out_ch = %d must be a multiple of 32 (<= 512)", NF);
    if (sp.S % 4 != 0) return fail(GAITK_E_SHAPE, "shared_out_ch must be a multiple of 4");
    sp.NFL = NF / 32;
    const int Tmax = std::max(T, T_in);
    sp.W = (pool_sensor || Tmax > NT / 2) ? 1 : std::min(WMAX, NT / Tmax);
    while (sp.W & (sp.W - 1)) --sp.W;                   // power of two (row <-> (t, w) by shifts)
    sp.rows = T * sp.W; sp.rows_in = T_in * sp.W; sp.halo = (KT1 / 2 > 1 ? KT1 / 2 : 1) * sp.W;
    // long windows (T a multiple of 128 above 128, e.g. the scaled sweep's T = 256): the window is split by time over a
    // thread-block cluster of T / 128 CTAs (stream_kernel.cuh); pooling bins must not straddle CTAs
    sp.cl = 1;
    if (!pool_sensor && T_in == T && T > NT && T % NT == 0 && enc != ENC_NONE) {
        const int cl = T / NT;
        if (cl > 8) return fail(GAITK_E_SHAPE, "window length %d needs a cluster of %d CTAs (> 8)", T, cl);
        if (T % d.backbone_dim != 0 || d.backbone_dim % cl != 0)
            return fail(GAITK_E_SHAPE, "window length %d with %d pooling bins: bins would straddle the %d CTAs of a cluster", T, d.backbone_dim, cl);
        sp.cl = cl; sp.rows = NT; sp.rows_in = NT;
    }
    // SensorEncoder that pools (no non-linearity between Conv1d and AdaptiveAvgPool1d): the pooled-taps kernel pools the raw clip
    // while loading it and runs Linear(3 CIN -> C) over the T pooled rows (stream_common.cuh, ENC_POOL_LINEAR); GAITK_POOLED_TAPS=0
    // keeps the literal conv-over-T_in-rows-then-pool kernel (the A/B switch of the parity tests)
    static const int pooled_mode = [] { const char* e = getenv("GAITK_POOLED_TAPS"); return e ? atoi(e) : 1; }();
    sp.pooled_taps = 0;
    KernelKey key = {enc, CIN, KT1, H, sp.C, sp.S, sp.NFL, sp.PROJ};
    sp.fn = nullptr;
    if (enc == ENC_CONV_POOL && pool_sensor && KT1 == 3 && pooled_mode) {
        KernelKey kp = {ENC_POOL_LINEAR, 3 * CIN, 1, H, sp.C, sp.S, sp.NFL, sp.PROJ};
        sp.fn = find_kernel(kp);
        if (sp.fn) { sp.pooled_taps = 1; sp.rows_in = sp.rows; key = kp; }
    }
    sp.RB = round_rb(sp.rows, sp.halo); sp.RBi = round_rb(sp.rows_in, sp.halo);
    if (!sp.fn) sp.fn = find_kernel(key);
    if (!sp.fn)
        return fail(GAITK_E_SHAPE, "no sm_100a kernel instantiated for stream %d (enc=%d Cin=%d k=%d H=%d C=%d S=%d NF=%d); "
                    "supported: WearGait C=12/S=16 and C=24/S=32 (bdim=8), FoG and FBG defaults", s, enc, CIN, KT1, H, sp.C, sp.S, NF);
    const int raw_cin = CIN;
    if (sp.pooled_taps) { CIN = 3 * CIN; KT1 = 1; }                    // what the kernel convolves (sp.CIN / sp.KT1 keep the model's)
    // ---- shared memory plan (floats)
    const int CI4 = (CIN + 3) / 4, C4 = (sp.C + 3) / 4, CP = C4 * 4, H4 = (H + 3) / 4, S4 = sp.S / 4;
    const int O1 = (enc == ENC_INSOLE) ? H4 * 4 : CP;
    const int CB = sp.PROJ ? sp.PROJ : sp.C, CB4 = (CB + 3) / 4, CBP = CB4 * 4;
    SmemPlan& m = sp.sp; memset(&m, 0, sizeof(m));
    int o = 0;
    auto take = [&](int n) { int r = o; o += (n + 3) / 4 * 4; return r; };
    m.X = take(CI4 * sp.RBi * 4);
    if (enc == ENC_INSOLE) { m.HA = take(H4 * sp.RB * 4); m.D1 = take(H4 * sp.RB * 4); }
    m.XH = take(C4 * sp.RB * 4);
    if (enc == ENC_CONV_GELU_LN || enc == ENC_INSOLE) m.D = take(C4 * sp.RB * 4);
    m.F = take(C4 * sp.RB * 4);
    m.RSTD = take(sp.RB);
    m.Z = take(S4 * sp.RB * 4);
    if (enc == ENC_CONV_POOL) m.A = take(C4 * sp.RBi * 4);
    m.W1F = take(KT1 * CI4 * 4 * O1); m.B1 = take(O1);
    if (enc == ENC_INSOLE) { m.W2F = take(3 * H4 * 4 * CP); m.B2 = take(CP); m.W2D = take(3 * CP * H4 * 4); }
    m.LNG = take(CP); m.LNB = take(CP);
    m.WBF = take(3 * CBP * sp.S); m.BB = take(sp.S); m.WBD = take(3 * sp.S * CBP);
    if (sp.PROJ) { m.L = take(CB4 * sp.RB * 4); m.DL = take(CB4 * sp.RB * 4); m.WPF = take(CP * CBP); m.BP = take(CBP); m.WPD = take(sp.PROJ * CP); }
    m.HW = take(KMAX * NF); m.HB = take(KMAX); m.HNG = take(NF); m.HNB = take(NF); m.INW = take(KMAX);
    m.P = take(WMAX * NF); m.DP = take(WMAX * NF); m.LOGIT = take(WMAX * KMAX);
    m.BINS = take(2 * d.backbone_dim + 4 * T + 2 * T_in + 8);
    int nblk_max = std::max(std::max(KT1 * CI4 * (O1 / 4), 3 * CB4 * S4), enc == ENC_INSOLE ? 3 * H4 * C4 : 0);
    const int stage_floats = 16 * std::max(std::max(NF, 256), std::max(nblk_max, NT));
    m.STAGE = take(stage_floats);
    // bulk-prefetch staging of the next tile's raw windows (stream_kernel.cuh): the FoG / FBG streams, whose kernels were
    // bound by the synchronous global -> shared scatter (the pooled-taps kernel has no other load path), and the WearGait
    // encoders of the fp32 path (GAITK_F32_PREFETCH = 2: FoG / FBG only, 0: off)
    static const int prefetch_mode = [] { const char* e = getenv("GAITK_F32_PREFETCH"); return e ? atoi(e) : 1; }();
    const bool pf_enc = enc == ENC_LINEAR_LN_RELU || (prefetch_mode >= 1 && prefetch_mode != 2 && (enc == ENC_CONV_GELU_LN || enc == ENC_INSOLE));
    if (sp.pooled_taps || (prefetch_mode && pf_enc && sp.cl == 1 && T_in == T && T_in * raw_cin >= 8)) {
        m.STGN = (T_in * raw_cin + 3) / 4 * 4 + 4;
        // the flush scratch is idle inside the tile loop (and no copy is in flight when the flush starts): the staging aliases it
        m.STG = sp.W * m.STGN <= stage_floats ? m.STAGE : take(sp.W * m.STGN);
        m.MBAR = take(4);
    }
    m.total = o;
    sp.smem_bytes = (size_t)o * sizeof(float);
    if (sp.smem_bytes > 227 * 1024)
        return fail(GAITK_E_SHAPE, "stream %d needs %zu B of shared memory (> 227 KB)", s, sp.smem_bytes);
    CUDA_TRY(cudaFuncSetAttribute((const void*)sp.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem_bytes));
    int occ = 0;
    CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, (const void*)sp.fn, NT, sp.smem_bytes));
    if (occ < 1) return fail(GAITK_E_SHAPE, "stream %d kernel does not fit on an SM", s);
    sp.ctas_per_sm = occ;
    // ---- tensor-core variant (tcgen05 + mma.sync, tf32): full 128-row tiles only
    const bool fixed_geo = T == 64 && sp.W == 2 && d.backbone_dim == 8;     // WearGait default window: compile-time geometry
    sp.fn_tc = (sp.rows == NT && T_in == T && sp.PROJ == 0 && sp.cl == 1) ? find_kernel_tc(key, fixed_geo) : nullptr;
    if (sp.fn_tc) {
        auto ev = [](int x) { return (x + 1) / 2 * 2; };
        const int KX = ev(CI4), KH = ev(H4), KC = ev(C4), KS = ev(S4);
        auto r16 = [](int x) { return (x + 15) / 16 * 16; };
        const int N1 = enc == ENC_INSOLE ? r16(H) : r16(sp.C), NC = r16(sp.C), NS = r16(sp.S), NH = r16(H);
        TcPlan& t = sp.tp; memset(&t, 0, sizeof(t));
        int q = 0;
        auto tk = [&](int n) { int rr = q; q += (n + 3) / 4 * 4; return rr; };
        t.X = tk(KX * sp.RB * 4);
        if (enc == ENC_INSOLE) { t.HA = tk(KH * sp.RB * 4); t.D1 = tk(H4 * sp.RB * 4); }
        t.XH = tk(KC * sp.RB * 4); t.D = tk(C4 * sp.RB * 4); t.F = tk(KC * sp.RB * 4); t.RSTD = tk(sp.RB); t.Z = tk(KS * sp.RB * 4);
        t.W1B = tk(KT1 * KX * N1 * 4); t.B1 = tk(O1);
        if (enc == ENC_INSOLE) { t.W2B = tk(3 * KH * NC * 4); t.B2 = tk(CP); t.W2D = tk(3 * KC * NH * 4); }
        t.LNG = tk(CP); t.LNB = tk(CP);
        t.WBB = tk(3 * KC * NS * 4); t.BB = tk(sp.S); t.WBD = tk(3 * KS * NC * 4);
        t.HW = tk(d.num_classes * NF); t.HB = tk(KMAX); t.HNG = tk(NF); t.HNB = tk(NF); t.INW = tk(KMAX);
        t.DP = tk(sp.W * NF); t.P = tk(sp.W * NF); t.BINS = tk(2 * d.backbone_dim + 2 * T + 8);
        t.STG = tk(sp.W * T * CIN);                      // TMA bulk-prefetch staging (raw window bytes)
        t.STAGE = 0;                                     // flush staging aliases the (dead by then) activations
        if (q < 16 * std::max(NF, 256)) q = 16 * std::max(NF, 256);
        t.total = q;
        sp.smem_tc = (size_t)q * sizeof(float);
        if (sp.smem_tc > 200 * 1024) sp.fn_tc = nullptr;
    }
    sp.tc_threads = NT;
#ifdef GAITK_WITH_TC2
    if (sp.fn_tc && fixed_geo) {
        const char* v = getenv("GAITK_TC_VARIANT");          // 2 = two threads per row (stream_kernel_tc2)
        StreamKernelTcFn f2 = (v && atoi(v) == 2) ? find_kernel_tc2(key) : nullptr;
        if (f2) { sp.fn_tc = f2; sp.tc_threads = NT2; }
    }
#endif
    // ---- warp-specialised split-bf16 variant (all contractions on tcgen05): WearGait default geometry, plain linear head
    sp.has_ws = 0;
    if (fixed_geo && !d.use_norm && !d.use_cosine && find_kernel_ws(key, d.num_classes, &sp.ws)) {
        CUDA_TRY(cudaFuncSetAttribute((const void*)sp.ws.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, sp.ws.smem));
        sp.has_ws = 1;
    }
    if (sp.fn_tc) {
        CUDA_TRY(cudaFuncSetAttribute((const void*)sp.fn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sp.smem_tc));
        // cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for kernels that allocate TMEM although
        // registers and shared memory admit more (ncu: occupancy_limit_registers / _shared_mem).  The kernel is
        // persistent with a grid-stride tile loop, so any grid is correct: size it from registers / shared
        // memory / TMEM columns and let the hardware co-schedule what it can.
        cudaFuncAttributes fa;
        CUDA_TRY(cudaFuncGetAttributes(&fa, (const void*)sp.fn_tc));
        const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * sp.tc_threads;
        const size_t smem_per_cta = sp.smem_tc + fa.sharedSizeBytes + 1024;
        int by_regs = 65536 / std::max(regs_per_cta, 1), by_smem = (int)((size_t)227 * 1024 / smem_per_cta);
        int occ2 = std::max(1, std::min(std::min(by_regs, by_smem), 16));      // 32 TMEM columns per CTA -> <= 16
        if (const char* e = getenv("GAITK_TC_CTAS")) occ2 = std::max(1, atoi(e));
        sp.ctas_tc = occ2;
    }
    return 0;
}

// stream-local gradient layout + reduce segments
static void layout_stream_grads(gaitk_plan* pl, int s) {
    StreamPlan& sp = pl->st[s];
    GradOff& go = sp.go; memset(&go, 0xff, sizeof(go));     // all -1
    int o = 0; sp.nseg = 0;
    auto seg = [&](int pidx, int& field) {
        if (pidx < 0) { field = -1; return; }
        const ParamInfo& p = pl->params[pidx];
        field = o;
        Seg sg; sg.src = o; sg.len = p.numel; sg.shared_off = p.shared_off; sg.param_off = (int)p.off;
        sp.seg[sp.nseg++] = sg;
        o += (p.numel + 3) / 4 * 4;
    };
    seg(sp.p_w1, go.w1); seg(sp.p_b1, go.b1); seg(sp.p_w2, go.w2); seg(sp.p_b2, go.b2);
    seg(sp.p_wsk, go.wsk); seg(sp.p_bsk, go.bsk); seg(sp.p_lng, go.lng); seg(sp.p_lnb, go.lnb);
    seg(sp.p_wp, go.wp); seg(sp.p_bp, go.bp);
    seg(pl->p_wbb, go.wbb); seg(pl->p_bbb, go.bbb);
    seg(sp.p_hng, go.hng); seg(sp.p_hnb, go.hnb); seg(sp.p_hw, go.hw); seg(sp.p_hb, go.hb);
    go.total = o;
    sp.NGP = o + 4;
}

extern "C" int gaitk_version(void) { return GAITK_VERSION; }
extern "C" const char* gaitk_last_error(void) { return g_err; }

extern "C" int gaitk_plan_create(const gaitk_model_desc* desc, int device, gaitk_plan** out) {
    if (!desc || !out) return fail(GAITK_E_BADARG, "null argument");
    *out = nullptr;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(GAITK_E_ARCH, "device %d is sm_%d%d; libgaitk is built for sm_100a only (no fallback path)", device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(device));
    const gaitk_model_desc& d = *desc;
    if (d.num_classes < 2 || d.num_classes > KMAX) return fail(GAITK_E_SHAPE, "num_classes must be in [2,%d]", KMAX);
    if (d.T < 1 || d.enc_out_ch < 1 || d.shared_out_ch < 1 || d.backbone_dim < 1) return fail(GAITK_E_SHAPE, "bad dims");
    if (d.backbone_dim > d.T) return fail(GAITK_E_SHAPE, "backbone_dim > T");
    gaitk_plan* pl = new gaitk_plan();
    pl->d = d; pl->device = device; pl->sm_count = prop.multiProcessorCount; pl->NP = 0; pl->P = 0;
    const int C = d.enc_out_ch, S = d.shared_out_ch, K = d.num_classes, NF = d.backbone_dim * S;
    pl->NF = NF;
    const bool norm = d.use_norm || d.use_cosine, cosn = d.use_cosine != 0, sync = d.synchronized != 0;
    int rc = 0;
    auto head = [&](const std::string& pre, int group, StreamPlan& sp) {
        sp.p_hng = sp.p_hnb = sp.p_hb = -1;
        if (norm) { sp.p_hng = add_param(pl, pre + "norm.weight", group, NF); sp.p_hnb = add_param(pl, pre + "norm.bias", group, NF); }
        sp.p_hw = add_param(pl, pre + "fc.weight", group, K, NF);
        if (!cosn) sp.p_hb = add_param(pl, pre + "fc.bias", group, K);
    };
    for (int s = 0; s < GAITK_MAX_STREAMS; ++s) {
        StreamPlan& sp = pl->st[s];
        sp.p_w1 = sp.p_b1 = sp.p_w2 = sp.p_b2 = sp.p_wsk = sp.p_bsk = sp.p_lng = sp.p_lnb = sp.p_hng = sp.p_hnb = sp.p_hw = sp.p_hb = sp.p_wp = sp.p_bp = -1;
    }
    if (d.family == GAITK_FAMILY_WEARGAIT) {
        // named_parameters() order of WearGaitThreeModal (weargait_encoders.py:121-141)
        pl->n_streams = 3;
        const int H = 2 * C;
        StreamPlan &w = pl->st[0], &i = pl->st[1], &m = pl->st[2];
        w.p_w1 = add_param(pl, "enc_w.conv.weight", 1, C, 2, 3); w.p_b1 = add_param(pl, "enc_w.conv.bias", 1, C);
        w.p_lng = add_param(pl, "enc_w.ln.weight", 1, C); w.p_lnb = add_param(pl, "enc_w.ln.bias", 1, C);
        i.p_w1 = add_param(pl, "enc_i.conv1.weight", 2, H, 13, 5); i.p_b1 = add_param(pl, "enc_i.conv1.bias", 2, H);
        add_param(pl, "enc_i.ln1.weight", -1, H); add_param(pl, "enc_i.ln1.bias", -1, H);     // constructed, never used (:81,:93-101)
        i.p_w2 = add_param(pl, "enc_i.conv2.weight", 2, C, H, 3); i.p_b2 = add_param(pl, "enc_i.conv2.bias", 2, C);
        i.p_lng = add_param(pl, "enc_i.ln2.weight", 2, C); i.p_lnb = add_param(pl, "enc_i.ln2.bias", 2, C);
        i.p_wsk = add_param(pl, "enc_i.skip.weight", 2, C, H, 1); i.p_bsk = add_param(pl, "enc_i.skip.bias", 2, C);
        m.p_w1 = add_param(pl, "enc_m.conv.weight", 3, C, 24, 3); m.p_b1 = add_param(pl, "enc_m.conv.bias", 3, C);
        m.p_lng = add_param(pl, "enc_m.ln.weight", 3, C); m.p_lnb = add_param(pl, "enc_m.ln.bias", 3, C);
        const int proj = d.reserved[0];                       // SharedLatent3 projection width (0 = WearGaitThreeModal)
        if (proj > 0) {
            w.p_wp = add_param(pl, "proj_w.weight", 1, proj, C); w.p_bp = add_param(pl, "proj_w.bias", 1, proj);
            i.p_wp = add_param(pl, "proj_i.weight", 2, proj, C); i.p_bp = add_param(pl, "proj_i.bias", 2, proj);
            m.p_wp = add_param(pl, "proj_m.weight", 3, proj, C); m.p_bp = add_param(pl, "proj_m.bias", 3, proj);
        }
        pl->p_wbb = add_param(pl, "backbone.conv.weight", 0, S, proj > 0 ? proj : C, 3); pl->p_bbb = add_param(pl, "backbone.conv.bias", 0, S);
        if (sync) {
            // _shared_or_three_heads assigns _shared_head first (weargait_encoders.py:306), WearGaitThreeModal head_w first (:135)
            head(proj > 0 ? "_shared_head." : "head_w.", 0, w);
            i.p_hng = m.p_hng = w.p_hng; i.p_hnb = m.p_hnb = w.p_hnb; i.p_hw = m.p_hw = w.p_hw; i.p_hb = m.p_hb = w.p_hb;
        } else {
            head("head_w.", 1, w); head("head_i.", 2, i); head("head_m.", 3, m);
        }
        if ((rc = plan_stream(pl, 0, ENC_CONV_GELU_LN, 2, 3, 0, d.T, d.T, 0))) { delete pl; return rc; }
        if ((rc = plan_stream(pl, 1, ENC_INSOLE, 13, 5, H, d.T, d.T, 0))) { delete pl; return rc; }
        if ((rc = plan_stream(pl, 2, ENC_CONV_GELU_LN, 24, 3, 0, d.T, d.T, 0))) { delete pl; return rc; }
    } else if (d.family == GAITK_FAMILY_FOG) {
        // named_parameters() order of MultiModalMultiTaskModel (feature_encoder.py:175-216)
        pl->n_streams = 2;
        StreamPlan &k = pl->st[0], &e = pl->st[1];
        k.p_w1 = add_param(pl, "skeleton_encoder.fc1.weight", 1, C, d.skel_in_dim); k.p_b1 = add_param(pl, "skeleton_encoder.fc1.bias", 1, C);
        k.p_lng = add_param(pl, "skeleton_encoder.ln1.weight", 1, C); k.p_lnb = add_param(pl, "skeleton_encoder.ln1.bias", 1, C);
        e.p_w1 = add_param(pl, "sensor_encoder.conv1d.weight", 2, C, d.sensor_in_ch, 3); e.p_b1 = add_param(pl, "sensor_encoder.conv1d.bias", 2, C);
        pl->p_wbb = add_param(pl, "backbone.conv1d.weight", 0, S, C, 3); pl->p_bbb = add_param(pl, "backbone.conv1d.bias", 0, S);
        if (sync) {
            head("task_head_shared.", 0, k);
            e.p_hng = k.p_hng; e.p_hnb = k.p_hnb; e.p_hw = k.p_hw; e.p_hb = k.p_hb;
        } else {
            head("task_head_skel.", 1, k); head("task_head_sensor.", 2, e);
        }
        if (d.sensor_out_len != d.T) { delete pl; return fail(GAITK_E_SHAPE, "sensor_out_len (%d) must equal pose length T (%d)", d.sensor_out_len, d.T); }
        if ((rc = plan_stream(pl, 0, ENC_LINEAR_LN_RELU, d.skel_in_dim, 1, 0, d.T, d.T, 0))) { delete pl; return rc; }
        // SensorEncoder pools only when the input length equals sensor_length (feature_encoder.py:55); the
        // backbone then sees sensor_out_len rows.  (A sensor clip of any other length is not produced by the
        // reference pipeline: pad_or_trim fixes it to sensor_length.)
        if ((rc = plan_stream(pl, 1, ENC_CONV_POOL, d.sensor_in_ch, 3, 0, d.sensor_len, d.sensor_out_len, 1))) { delete pl; return rc; }
    } else {
        delete pl; return fail(GAITK_E_BADARG, "unknown family %d", d.family);
    }
    if ((int)pl->params.size() > MAX_PARAMS) { delete pl; return fail(GAITK_E_SHAPE, "too many parameters"); }
    for (int s = 0; s < pl->n_streams; ++s) layout_stream_grads(pl, s);
    *out = pl;
    return 0;
}

extern "C" void gaitk_plan_destroy(gaitk_plan* plan) { delete plan; }
extern "C" int gaitk_param_count(const gaitk_plan* p) { return p ? (int)p->params.size() : 0; }
extern "C" int64_t gaitk_param_total(const gaitk_plan* p) { return p ? p->NP : 0; }
extern "C" int64_t gaitk_shared_total(const gaitk_plan* p) { return p ? p->P : 0; }
extern "C" int gaitk_num_streams(const gaitk_plan* p) { return p ? p->n_streams : 0; }
extern "C" int gaitk_stream_in_dim(const gaitk_plan* p, int s) { return (p && s >= 0 && s < p->n_streams) ? p->st[s].CIN : 0; }
extern "C" int gaitk_stream_in_len(const gaitk_plan* p, int s) { return (p && s >= 0 && s < p->n_streams) ? p->st[s].T_in : 0; }
extern "C" int gaitk_stream_geometry(const gaitk_plan* p, int s, int dtype, int* ctas_per_sm, int* windows_per_tile, size_t* smem_bytes) {
    if (!p || s < 0 || s >= p->n_streams) return fail(GAITK_E_BADARG, "bad stream");
    const StreamPlan& sp = p->st[s];
    const bool tc = dtype == GAITK_DTYPE_TF32;
    if (tc && !sp.fn_tc) return fail(GAITK_E_DTYPE, "no tensor-core kernel for this stream");
    if (dtype == GAITK_DTYPE_BF16X3) {
        if (!sp.has_ws) return fail(GAITK_E_DTYPE, "no split-bf16 kernel for this stream");
        if (ctas_per_sm) *ctas_per_sm = 1;
        if (windows_per_tile) *windows_per_tile = 2 * sp.ws.groups;      // tiles in flight per CTA x windows per tile
        if (smem_bytes) *smem_bytes = (size_t)sp.ws.smem;
        return 0;
    }
    if (ctas_per_sm) *ctas_per_sm = tc ? sp.ctas_tc : sp.ctas_per_sm;
    if (windows_per_tile) *windows_per_tile = sp.W;
    if (smem_bytes) *smem_bytes = tc ? sp.smem_tc : sp.smem_bytes;
    return 0;
}
extern "C" int64_t gaitk_gbuf_floats(const gaitk_plan* p) { return p ? (int64_t)MAXT * p->P + p->NP + 8 : 0; }

extern "C" int gaitk_param_info(const gaitk_plan* p, int index, char* name, size_t cap, int64_t* offset, int64_t* numel,
                                int32_t* group, int32_t* dims) {
    if (!p || index < 0 || index >= (int)p->params.size()) return fail(GAITK_E_BADARG, "bad parameter index");
    const ParamInfo& q = p->params[index];
    if (name && cap) { strncpy(name, q.name.c_str(), cap - 1); name[cap - 1] = 0; }
    if (offset) *offset = q.off;
    if (numel) *numel = q.numel;
    if (group) *group = q.group;
    if (dims) for (int i = 0; i < 4; ++i) dims[i] = q.dims[i];
    return 0;
}

constexpr int COUPLE_MAX_CTAS = 1024;          // CTAs (256 samples each) of the consistency-coupling kernel
static int stream_grid(const gaitk_plan* pl, const StreamPlan& sp, int B, int dtype = GAITK_DTYPE_F32) {
    const int ntiles = (B + sp.W - 1) / sp.W;
    if (dtype == GAITK_DTYPE_BF16X3)                     // one persistent CTA per SM, `groups` tiles in flight in each
        return std::max(1, std::min((ntiles + sp.ws.groups - 1) / sp.ws.groups, pl->sm_count));
    const int per_sm = dtype == GAITK_DTYPE_TF32 ? sp.ctas_tc : sp.ctas_per_sm;
    if (sp.cl > 1 && dtype == GAITK_DTYPE_F32)            // one cluster per window at a time: grid = clusters * cl
        return std::max(1, std::min(ntiles, pl->sm_count * per_sm / sp.cl)) * sp.cl;
    return std::max(1, std::min(ntiles, pl->sm_count * per_sm));
}
static int check_dtype(const gaitk_plan* pl, int dtype) {
    if (dtype == GAITK_DTYPE_F32) return 0;
    if (dtype == GAITK_DTYPE_BF16X3) {
        for (int s = 0; s < pl->n_streams; ++s)
            if (!pl->st[s].has_ws) return fail(GAITK_E_DTYPE, "split-bf16 (bf16x3) kernels exist for WearGait C=12/S=16/T=64 with a plain linear head; stream %d has none", s);
        return 0;
    }
    if (dtype != GAITK_DTYPE_TF32) return fail(GAITK_E_DTYPE, "unknown dtype %d", dtype);
    for (int s = 0; s < pl->n_streams; ++s)
        if (!pl->st[s].fn_tc) return fail(GAITK_E_DTYPE, "tensor-core (tf32) kernels exist for WearGait C=12/S=16 with 128-row tiles; stream %d has none", s);
    return 0;
}
static size_t stream_ws_floats(const gaitk_plan* pl, const StreamPlan& sp, int B) {
    return (size_t)std::max(stream_grid(pl, sp, B), sp.fn_tc ? stream_grid(pl, sp, B, GAITK_DTYPE_TF32) : 0) * sp.NGP;
}
extern "C" size_t gaitk_workspace_bytes(const gaitk_plan* pl, int B) {
    if (!pl) return 0;
    size_t sum = 0;                                      // every stream keeps its own partial rows until the single reduce launch
    for (int s = 0; s < pl->n_streams; ++s) sum += (stream_ws_floats(pl, pl->st[s], B) + 63) / 64 * 64;
    // consistency-coupled step (2 streams): logits[2] + dlogits[4] of (B, K) + the coupling kernel's per-CTA partials
    if (pl->n_streams == 2) sum += (size_t)6 * ((size_t)std::max(B, 0) * KMAX + 64) + (size_t)COUPLE_MAX_CTAS * 8;
    return sum * sizeof(float) + 256;
}

static const float* pp(const float* params, const gaitk_plan* pl, int idx) { return idx < 0 ? nullptr : params + pl->params[idx].off; }

static void fill_args(const gaitk_plan* pl, int s, const float* params, const float* x, const int64_t* ws, int B,
                      int mode, int zero_input, StreamArgs& a) {
    const StreamPlan& sp = pl->st[s];
    memset(&a, 0, sizeof(a));
    a.x = x; a.win_start = (const long long*)ws; a.B = B; a.T_in = sp.T_in; a.T = sp.T; a.W = sp.W;
    a.bdim = pl->d.backbone_dim; a.K = pl->d.num_classes; a.NF = pl->NF;
    a.rows_in = sp.rows_in; a.rows = sp.rows; a.halo = sp.halo; a.RBi = sp.RBi; a.RB = sp.RB;
    a.mode = mode; a.zero_input = zero_input; a.pool_sensor = sp.pooled_taps ? 0 : sp.pool_sensor; a.cl = sp.cl;
    a.w1 = pp(params, pl, sp.p_w1); a.b1 = pp(params, pl, sp.p_b1); a.w2 = pp(params, pl, sp.p_w2); a.b2 = pp(params, pl, sp.p_b2);
    a.wsk = pp(params, pl, sp.p_wsk); a.bsk = pp(params, pl, sp.p_bsk); a.lng = pp(params, pl, sp.p_lng); a.lnb = pp(params, pl, sp.p_lnb);
    a.wbb = pp(params, pl, pl->p_wbb); a.bbb = pp(params, pl, pl->p_bbb);
    a.hng = pp(params, pl, sp.p_hng); a.hnb = pp(params, pl, sp.p_hnb); a.hw = pp(params, pl, sp.p_hw); a.hb = pp(params, pl, sp.p_hb);
    a.wp = pp(params, pl, sp.p_wp); a.bp = pp(params, pl, sp.p_bp);
    a.head_norm = sp.p_hng >= 0; a.head_cos = pl->d.use_cosine != 0; a.skip_identity = sp.skip_identity;
    a.scale = 1.f; for (int k = 0; k < KMAX; ++k) { a.margin[k] = 0.f; a.cls_w[k] = 1.f; }
    a.go = sp.go; a.NGP = sp.NGP;
}

static int launch_stream(const gaitk_plan* pl, int s, const StreamArgs& a, int grid, cudaStream_t st, int dtype = GAITK_DTYPE_F32) {
    const StreamPlan& sp = pl->st[s];
    if (dtype == GAITK_DTYPE_BF16X3) sp.ws.fn<<<grid, sp.ws.threads, sp.ws.smem, st>>>(a);
    else if (dtype == GAITK_DTYPE_TF32) sp.fn_tc<<<grid, sp.tc_threads, sp.smem_tc, st>>>(a, sp.tp);
    else if (sp.cl > 1) {
        cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid); cfg.blockDim = dim3(NT); cfg.dynamicSmemBytes = sp.smem_bytes; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = sp.cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        CUDA_TRY(cudaLaunchKernelEx(&cfg, sp.fn, a, sp.sp));
    }
    else sp.fn<<<grid, NT, sp.smem_bytes, st>>>(a, sp.sp);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gaitk_forward(gaitk_plan* pl, const float* params, const float* const* x, const int64_t* const* win_start,
                             int B, uint32_t enabled_mask, float* const* logits, int dtype, void* stream) {
    if (!pl || !params || !x || !logits) return fail(GAITK_E_BADARG, "null argument");
    { int rc_ = check_dtype(pl, dtype); if (rc_) return rc_; }
    if (B <= 0) return 0;
    for (int s = 0; s < pl->n_streams; ++s) {
        if (!logits[s]) continue;
        StreamArgs a; fill_args(pl, s, params, x[s], win_start ? win_start[s] : nullptr, B, MODE_FWD, !(enabled_mask & (1u << s)), a);
        a.logits = logits[s];
        int rc = launch_stream(pl, s, a, stream_grid(pl, pl->st[s], B, dtype), (cudaStream_t)stream, dtype);
        if (rc) return rc;
    }
    return 0;
}

static void set_loss(StreamArgs& a, const gaitk_loss_desc& L, int K) {
    a.scale = L.scale; a.nan_degenerate = L.nan_if_degenerate;
    for (int k = 0; k < KMAX; ++k) { a.margin[k] = k < K ? L.margin[k] : 0.f; a.cls_w[k] = k < K ? L.cls_weight[k] : 0.f; }
}

static void fill_reduce(const gaitk_plan* pl, int s, const float* partial, int grid, float* gbuf, int task, float mult,
                        int stat_slot, ReduceArgs& R) {
    const StreamPlan& sp = pl->st[s];
    memset(&R, 0, sizeof(R));
    R.partial = partial; R.grid = grid; R.NGP = sp.NGP; R.NG = sp.go.total; R.nseg = sp.nseg;
    for (int i = 0; i < sp.nseg; ++i) R.seg[i] = sp.seg[i];
    R.gbuf = gbuf; R.P = pl->P; R.NP = pl->NP; R.task = task; R.private_mult = mult; R.stat_slot = stat_slot;
}
static int launch_reduce(const gaitk_plan* pl, int s, const float* partial, int grid, float* gbuf, int task, float mult,
                         int stat_slot, cudaStream_t st) {
    ReduceArgs R; fill_reduce(pl, s, partial, grid, gbuf, task, mult, stat_slot, R);
    const int n = R.NG + 2;
    reduce_partials_kernel<<<(n + 127) / 128, dim3(128, 8), 0, st>>>(R);
    LAUNCH_CHECK();
    return 0;
}

struct LossArgs { float scale; float margin[KMAX]; float cls_w[KMAX]; int nan_degenerate; };

// ------------------------------------------------------------------------------------------ consistency-coupled step
// fbg_fog_train.py:81-89,121-124 (synchronised FoG, wm = gcl): l_s = criterion_s(logits_s, y_s) + 0.5 lambda (KL(q || p) + KL(p || q)),
// p = softmax(logits_0), q = softmax(logits_1), both KL terms 'batchmean' over the GLOBAL batch and differentiated through BOTH
// arguments.  With u = log p - log q:  d KLsym / d logits_0[j] = p_j (u_j - sum_k p_k u_k) + p_j - q_j  (and symmetrically for
// logits_1), so every task's gradient reaches both streams.  One thread per sample writes, per stream s,
//   dl[2 s]     = d criterion_s / d logits_s + 0.5 lambda d KLsym / d logits_s     (task s through its own stream)
//   dl[2 s + 1] =                               0.5 lambda d KLsym / d logits_s     (the OTHER task through stream s)
// and the per-CTA partial sums of (loss_0, loss_1, correct_0, correct_1); the last CTA adds them in CTA order (deterministic).
struct CoupleArgs {
    const float* logits[2]; const long long* y[2]; const float* logit_off[2]; LossArgs L[2];
    const float* denom; float lambda; int B, K;
    float* dl[4]; float* partial; unsigned* ticket; float* stats;     // stats = gbuf tail: loss[4] | correct[4]
};
__global__ void __launch_bounds__(256) couple_kernel(const CoupleArgs A) {
    __shared__ float sh[4][8];
    __shared__ int last;
    const int b = blockIdx.x * blockDim.x + threadIdx.x, K = A.K;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    if (b < A.B) {
        float lp[2][KMAX], pr[2][KMAX];
        const float inv_B = 1.0f / A.denom[GAITK_DENOM_COUNT];
        for (int s = 0; s < 2; ++s) {
            const float* lg = A.logits[s] + (size_t)b * K;
            float mx = -INFINITY;
            for (int k = 0; k < K; ++k) mx = fmaxf(mx, lg[k]);
            float se = 0.f;
            for (int k = 0; k < K; ++k) se += expf(lg[k] - mx);
            const float lse = mx + logf(se);
            for (int k = 0; k < K; ++k) { lp[s][k] = lg[k] - lse; pr[s][k] = expf(lp[s][k]); }
        }
        float kl = 0.f, su0 = 0.f, su1 = 0.f;
        for (int k = 0; k < K; ++k) {
            const float u = lp[0][k] - lp[1][k];
            kl += (pr[0][k] - pr[1][k]) * u; su0 += pr[0][k] * u; su1 += pr[1][k] * u;
        }
        const float hl = 0.5f * A.lambda * inv_B;
        for (int s = 0; s < 2; ++s) {
            const LossArgs& L = A.L[s];
            const float* lg = A.logits[s] + (size_t)b * K;
            const int yy = (int)A.y[s][b];
            const float inv_denom = 1.0f / A.denom[s];
            float zz[KMAX]; float mx = -INFINITY, best = -INFINITY; int am = 0;
            for (int k = 0; k < K; ++k) {
                float z = lg[k];
                if (z > best) { best = z; am = k; }
                if (A.logit_off[s]) z -= A.logit_off[s][(size_t)b * K + k];
                if (k == yy) z -= L.margin[k];
                z *= L.scale;
                if (L.nan_degenerate) z = __int_as_float(0x7fc00000);
                zz[k] = z; mx = fmaxf(mx, z);
            }
            float se = 0.f;
            for (int k = 0; k < K; ++k) se += expf(zz[k] - mx);
            const float lse = mx + logf(se);
            const float wy = L.cls_w[yy];
            acc[s] = wy * (lse - zz[yy]) * inv_denom + hl * kl;
            acc[2 + s] = (am == yy) ? 1.f : 0.f;
            for (int k = 0; k < K; ++k) {
                const float u = lp[0][k] - lp[1][k];
                const float dk = s == 0 ? pr[0][k] * (u - su0) + pr[0][k] - pr[1][k] : pr[1][k] * (su1 - u) + pr[1][k] - pr[0][k];
                const float dce = L.scale * wy * inv_denom * (expf(zz[k] - lse) - (k == yy ? 1.f : 0.f));
                A.dl[2 * s][(size_t)b * K + k] = dce + hl * dk;
                A.dl[2 * s + 1][(size_t)b * K + k] = hl * dk;
            }
        }
    }
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    for (int i = 0; i < 4; ++i) {
        float v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) sh[i][wrp] = v;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        float v = 0.f;
        for (int w = 0; w < 8; ++w) v += sh[threadIdx.x][w];
        A.partial[(size_t)blockIdx.x * 4 + threadIdx.x] = v;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(A.ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (last && threadIdx.x < 4) {
        __threadfence();
        float v = 0.f;
        for (unsigned c = 0; c < gridDim.x; ++c) v += reinterpret_cast<volatile const float*>(A.partial)[(size_t)c * 4 + threadIdx.x];
        const int i = threadIdx.x;
        A.stats[(i >> 1) * 4 + (i & 1)] += v;          // loss[s] / correct[s]
        if (i == 0) *A.ticket = 0u;
    }
}

static int step_grads_coupled(gaitk_plan* pl, const float* params, const float* const* x, const int64_t* const* win_start,
                              const int64_t* const* y, int B, const gaitk_loss_desc* loss, const float* const* logit_off,
                              const float* denom, uint32_t enabled_mask, float private_mult, float lambda,
                              float* const* logits, float* gbuf, float* part, int dtype, cudaStream_t st) {
    const int K = pl->d.num_classes;
    if (pl->n_streams != 2) return fail(GAITK_E_BADARG, "the consistency term couples exactly two streams");
    if ((B + 255) / 256 > COUPLE_MAX_CTAS) return fail(GAITK_E_SHAPE, "coupled step: at most %d samples per call", COUPLE_MAX_CTAS * 256);
    float* my_part[2];
    for (int s = 0; s < 2; ++s) { my_part[s] = part; part += (stream_ws_floats(pl, pl->st[s], B) + 63) / 64 * 64; }
    const size_t per = ((size_t)B * KMAX + 63) / 64 * 64;
    float* lg[2]; float* dl[4];
    for (int s = 0; s < 2; ++s) { lg[s] = (logits && logits[s]) ? logits[s] : part; part += per; }
    for (int i = 0; i < 4; ++i) { dl[i] = part; part += per; }
    float* cpart = part; part += (size_t)COUPLE_MAX_CTAS * 4;
    unsigned* ticket = reinterpret_cast<unsigned*>(part);
    CUDA_TRY(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));
    // pass 1: logits of both streams
    for (int s = 0; s < 2; ++s) {
        StreamArgs a; fill_args(pl, s, params, x[s], win_start ? win_start[s] : nullptr, B, MODE_FWD, !(enabled_mask & (1u << s)), a);
        a.logits = lg[s];
        int rc = launch_stream(pl, s, a, stream_grid(pl, pl->st[s], B, dtype), st, dtype);
        if (rc) return rc;
    }
    // pass 2: losses + the four logit-gradient vectors
    CoupleArgs C; memset(&C, 0, sizeof(C));
    for (int s = 0; s < 2; ++s) {
        C.logits[s] = lg[s]; C.y[s] = (const long long*)y[s]; C.logit_off[s] = logit_off ? logit_off[s] : nullptr;
        C.L[s].scale = loss[s].scale; C.L[s].nan_degenerate = loss[s].nan_if_degenerate;
        for (int k = 0; k < KMAX; ++k) { C.L[s].margin[k] = k < K ? loss[s].margin[k] : 0.f; C.L[s].cls_w[k] = k < K ? loss[s].cls_weight[k] : 0.f; }
    }
    for (int i = 0; i < 4; ++i) C.dl[i] = dl[i];
    C.denom = denom; C.lambda = lambda; C.B = B; C.K = K; C.partial = cpart; C.ticket = ticket;
    C.stats = gbuf + (size_t)MAXT * pl->P + pl->NP;
    couple_kernel<<<(B + 255) / 256, 256, 0, st>>>(C);
    LAUNCH_CHECK();
    // pass 3: task t through stream s (recompute + backward with external logit gradients); shared parts -> column t of G,
    // private parts add up over the tasks (fbg_fog_train.py:146-152: every loss reaches both encoders)
    for (int t = 0; t < 2; ++t)
        for (int s = 0; s < 2; ++s) {
            StreamArgs a; fill_args(pl, s, params, x[s], win_start ? win_start[s] : nullptr, B, MODE_BWD_EXT, !(enabled_mask & (1u << s)), a);
            a.dlogits_ext = dl[2 * s + (t == s ? 0 : 1)]; a.partial = my_part[s];
            const int grid = stream_grid(pl, pl->st[s], B, dtype);
            int rc = launch_stream(pl, s, a, grid, st, dtype);
            if (rc) return rc;
            if ((rc = launch_reduce(pl, s, my_part[s], grid, gbuf, t, private_mult, -1, st))) return rc;
        }
    return 0;
}

__global__ void zero_floats_kernel(float* p, int n) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = 0.f;
}
extern "C" int gaitk_step_grads(gaitk_plan* pl, const float* params, const float* const* x, const int64_t* const* win_start,
                                const int64_t* const* y, int B, const gaitk_loss_desc* loss, const float* const* logit_off,
                                const float* denom, uint32_t enabled_mask, uint32_t task_mask, float private_mult,
                                float consistency_lambda, float* const* logits, float* gbuf, void* workspace,
                                size_t workspace_bytes, int dtype, void* stream) {
    if (!pl || !params || !x || !y || !loss || !denom || !gbuf || !workspace) return fail(GAITK_E_BADARG, "null argument");
    { int rc_ = check_dtype(pl, dtype); if (rc_) return rc_; }
    if (workspace_bytes < gaitk_workspace_bytes(pl, B)) return fail(GAITK_E_BADARG, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    // a kernel, not cudaMemsetAsync: gbuf may live in a peer-mapped symmetric allocation (gaitk_p2p_allreduce), where a
    // memset node inside a CUDA graph is not a plain device-local fill
    zero_floats_kernel<<<8, 256, 0, st>>>(gbuf, (int)gaitk_gbuf_floats(pl));
    LAUNCH_CHECK();
    if (B <= 0) return 0;
    if (consistency_lambda != 0.f) {
        if ((task_mask & 3u) != 3u) return fail(GAITK_E_BADARG, "the consistency term needs both tasks");
        return step_grads_coupled(pl, params, x, win_start, y, B, loss, logit_off, denom, enabled_mask, private_mult,
                                  consistency_lambda, logits, gbuf, (float*)workspace, dtype, st);
    }
    ReduceArgsMulti M; memset(&M, 0, sizeof(M));
    int n_active = 0, max_ng = 0;
    float* part = (float*)workspace;
    // the kernels of the streams run on forked streams (joined before the reduce)
    bool fork = false;
    {
        int live = 0, widest = 0;
        for (int s = 0; s < pl->n_streams; ++s)
            if (task_mask & (1u << s)) { ++live; widest = std::max(widest, stream_grid(pl, pl->st[s], B, dtype) * (dtype == GAITK_DTYPE_BF16X3 ? 1 : 1)); }
        // GAITK_FORK: 0 = never, 2 = small batches only, 1 = always (default): besides running small kernels side by side, the CTAs of
        // one persistent kernel that finish a round early are followed at once by the next kernel's CTAs (0.9915 -> 0.9863 ms at
        // B = 32768, repeatable)
        static const int fork_mode = [] { const char* e = getenv("GAITK_FORK"); return e ? atoi(e) : 1; }();
        // "always" applies to the warp-specialised kernels (one persistent CTA per SM each); the fp32 / tf32 kernels (several CTAs
        // per SM, full grids) slow each other down when they share the SMs (FoG step at B = 32768: 3.7 -> 5.1 ms): small batches only
        fork = live > 1 && ((fork_mode == 1 && dtype == GAITK_DTYPE_BF16X3) || (fork_mode != 0 && widest < pl->sm_count));
        if (fork) { int rc_ = ensure_side_streams(pl); if (rc_) return rc_; CUDA_TRY(cudaEventRecord(pl->ev_fork, st)); }
    }
    int n_forked = 0, n_launched = 0;
    float* parts[GAITK_MAX_STREAMS];
    for (int s = 0; s < pl->n_streams; ++s) { parts[s] = part; part += (stream_ws_floats(pl, pl->st[s], B) + 63) / 64 * 64; }
    // launch order: the most expensive stream first.  Forked kernels that each fill the SMs run nearly back to back and only the
    // LAST one's tail is exposed, so the long insole kernel should not be it (B = 32768: insole, IMU, walkway 0.891 ms per step;
    // walkway, insole, IMU 0.905; GAITK_ORDER = e.g. "012" overrides).  Results do not depend on the order (per-stream partials).
    int order[GAITK_MAX_STREAMS];
    for (int q = 0; q < pl->n_streams; ++q) order[q] = q;
    auto cost = [&](int s_) { const StreamPlan& sp = pl->st[s_]; return (sp.enc == ENC_INSOLE ? 1000 : 0) + sp.CIN * sp.KT1; };
    std::stable_sort(order, order + pl->n_streams, [&](int a_, int b_) { return cost(a_) > cost(b_); });
    static const char* order_env = getenv("GAITK_ORDER");
    if (order_env && (int)strlen(order_env) == pl->n_streams) {
        bool ok = true; int seen = 0;
        for (int q = 0; q < pl->n_streams; ++q) { const int c = order_env[q] - '0'; if (c < 0 || c >= pl->n_streams || (seen >> c & 1)) ok = false; else seen |= 1 << c; }
        if (ok) for (int q = 0; q < pl->n_streams; ++q) order[q] = order_env[q] - '0';
    }
    for (int q = 0; q < pl->n_streams; ++q) {
        const int s = order[q];
        float* my_part = parts[s];
        if (!(task_mask & (1u << s))) continue;
        StreamArgs a; fill_args(pl, s, params, x[s], win_start ? win_start[s] : nullptr, B, MODE_FUSED, !(enabled_mask & (1u << s)), a);
        set_loss(a, loss[s], pl->d.num_classes);
        a.y = (const long long*)y[s]; a.denom = denom + s;
        a.logit_off = logit_off ? logit_off[s] : nullptr;
        a.logits = logits ? logits[s] : nullptr;
        a.partial = my_part;
        const int grid = stream_grid(pl, pl->st[s], B, dtype);
        cudaStream_t ks = st;
        if (fork && n_launched > 0) {                     // first active stream stays on the caller's stream
            ks = pl->side[n_forked];
            CUDA_TRY(cudaStreamWaitEvent(ks, pl->ev_fork, 0));
        }
        int rc = launch_stream(pl, s, a, grid, ks, dtype);
        if (rc) return rc;
        ++n_launched;
        if (ks != st) { CUDA_TRY(cudaEventRecord(pl->ev_join[n_forked], ks)); CUDA_TRY(cudaStreamWaitEvent(st, pl->ev_join[n_forked], 0)); ++n_forked; }
        // (a per-stream reduce on the forked streams, right after each kernel, was tried: B = 4096 step 0.230 -> 0.224 ms, but the
        // data-parallel step with the peer-memory exchange went from 0.93 to 1.10 ms at N = 2 -- one reduce after the join it is)
        fill_reduce(pl, s, my_part, grid, gbuf, s, private_mult, s, M.r[n_active]);
        max_ng = std::max(max_ng, M.r[n_active].NG + 2);
        ++n_active;
    }
    if (n_active) {
        reduce_partials_multi_kernel<<<dim3((max_ng + 127) / 128, n_active), dim3(128, 8), 0, st>>>(M);
        LAUNCH_CHECK();
    }
    return 0;
}

static int accumulate_grads(const gaitk_plan* pl, const float* gbuf, float* grads, cudaStream_t st);

extern "C" int gaitk_backward(gaitk_plan* pl, const float* params, const float* const* x, const int64_t* const* win_start,
                              int B, uint32_t enabled_mask, const float* const* dlogits, float* grads, void* workspace,
                              size_t workspace_bytes, int dtype, void* stream) {
    if (!pl || !params || !x || !dlogits || !grads || !workspace) return fail(GAITK_E_BADARG, "null argument");
    { int rc_ = check_dtype(pl, dtype); if (rc_) return rc_; }
    // scratch: [gbuf | partials]
    const size_t gb = (size_t)gaitk_gbuf_floats(pl) * sizeof(float);
    if (workspace_bytes < gb + gaitk_workspace_bytes(pl, B)) return fail(GAITK_E_BADARG, "workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    float* gbuf = (float*)workspace; float* partial = (float*)((char*)workspace + ((gb + 255) / 256) * 256);
    CUDA_TRY(cudaMemsetAsync(gbuf, 0, gb, st));
    if (B <= 0) return 0;
    for (int s = 0; s < pl->n_streams; ++s) {
        if (!dlogits[s]) continue;
        StreamArgs a; fill_args(pl, s, params, x[s], win_start ? win_start[s] : nullptr, B, MODE_BWD_EXT, !(enabled_mask & (1u << s)), a);
        a.dlogits_ext = dlogits[s]; a.partial = partial;
        const int grid = stream_grid(pl, pl->st[s], B, dtype);
        int rc = launch_stream(pl, s, a, grid, st, dtype);
        if (rc) return rc;
        // all shared contributions land in task column 0 with unit coefficient
        if ((rc = launch_reduce(pl, s, partial, grid, gbuf, 0, 1.0f, -1, st))) return rc;
    }
    // grads += [shared from G column 0 | private]
    return accumulate_grads(pl, gbuf, grads, st);
}

__global__ void accumulate_grads_kernel(const float* gbuf, float* grads, const UpdateArgs U) {
    const float* G = gbuf; const float* PG = gbuf + (size_t)MAXT * U.P;
    for (int ip = 0; ip < U.nparams; ++ip) {
        const ParamSeg s = U.ps[ip];
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < s.numel; e += gridDim.x * blockDim.x)
            grads[s.off + e] += s.shared_off >= 0 ? G[s.shared_off + e] : PG[s.off + e];
    }
}
static void fill_param_segs(const gaitk_plan* pl, UpdateArgs& U, const uint8_t* has_grad) {
    U.nparams = (int)pl->params.size(); U.P = pl->P; U.NP = pl->NP;
    for (int i = 0; i < U.nparams; ++i) {
        const ParamInfo& p = pl->params[i];
        U.ps[i].off = p.off; U.ps[i].numel = p.numel; U.ps[i].shared_off = p.shared_off;
        U.ps[i].has_grad = has_grad ? (has_grad[i] != 0) : (p.group >= 0);
    }
}
static int accumulate_grads(const gaitk_plan* pl, const float* gbuf, float* grads, cudaStream_t st) {
    UpdateArgs U; memset(&U, 0, sizeof(U)); fill_param_segs(pl, U, nullptr);
    accumulate_grads_kernel<<<8, 256, 0, st>>>(gbuf, grads, U);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gaitk_step_update(gaitk_plan* pl, float* params, float* momentum, const float* gbuf, uint32_t task_mask,
                                 float cagrad_c, float max_norm, float lr, float mom, float weight_decay, float* grads_out,
                                 float* diag, int solver, void* stream) {
    if (!pl || !gbuf) return fail(GAITK_E_BADARG, "null argument");
    const bool do_sgd = params && momentum;
    if (!do_sgd && !grads_out) return fail(GAITK_E_BADARG, "nothing to do: pass params+momentum and/or grads_out");
    task_mask &= (1u << pl->n_streams) - 1u;
    if (!task_mask) return fail(GAITK_E_BADARG, "empty task mask");
    UpdateArgs U; memset(&U, 0, sizeof(U)); fill_param_segs(pl, U, nullptr);
    // a stream whose task is masked out contributes no private gradients either: those parameters are skipped
    for (int i = 0; i < U.nparams; ++i) {
        const int g = pl->params[i].group;
        if (g >= 1 && !(task_mask & (1u << (g - 1)))) U.ps[i].has_grad = 0;
    }
    U.params = params; U.momentum = momentum; U.gbuf = gbuf; U.grads_out = grads_out; U.diag = diag;
    U.task_mask = task_mask; U.n_tasks_max = pl->n_streams; U.alpha = cagrad_c; U.max_norm = max_norm;
    U.lr = lr; U.mom = mom; U.wd = weight_decay; U.do_sgd = do_sgd ? 1 : 0; U.solver = solver & 0xff;
    U.check_exchange = (solver & GAITK_SOLVER_FLAG_CHECK_EXCHANGE) ? 1 : 0;
    cagrad_update_kernel<<<1, UPD_THREADS, 0, (cudaStream_t)stream>>>(U);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gaitk_p2p_allreduce(gaitk_plan* pl, const float* const* peer_gbuf_dev, uint32_t* const* peer_flag_dev,
                                   uint32_t* counter, int rank, int world, float* gsum, float* diag, void* stream) {
    if (!pl || !peer_gbuf_dev || !peer_flag_dev || !counter || !gsum) return fail(GAITK_E_BADARG, "null argument");
    if (world < 1 || world > 64 || rank < 0 || rank >= world) return fail(GAITK_E_BADARG, "bad rank / world size");
    P2PArgs Q; Q.peer_gbuf = peer_gbuf_dev; Q.peer_flag = (unsigned* const*)peer_flag_dev; Q.counter = counter;
    Q.gsum = gsum; Q.n = (int)gaitk_gbuf_floats(pl); Q.rank = rank; Q.world = world; Q.diag = diag;
    Q.timeout_cycles = 8ll << 30;                              // ~4 s of SM clocks
    if (const char* e = getenv("GAITK_P2P_TIMEOUT_CYCLES")) { const long long v = atoll(e); if (v > 0) Q.timeout_cycles = v; }
    const int grid = std::max(1, std::min(16, (Q.n + P2P_THREADS * 4 - 1) / (P2P_THREADS * 4)));
    p2p_allreduce_kernel<<<grid, P2P_THREADS, 0, (cudaStream_t)stream>>>(Q);
    LAUNCH_CHECK();
    bump_counter_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(counter);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gaitk_cagrad(const float* G, int P, int n_tasks, float c, float max_norm, float* shared_grad, float* diag, int solver,
                            void* stream) {
    if (!G || !shared_grad || P <= 0 || n_tasks < 1 || n_tasks > MAXT) return fail(GAITK_E_BADARG, "bad argument");
    // G is (n_tasks x P) task-major (the kernel addresses task t at G + t*P); one pseudo-parameter covers all of G
    UpdateArgs U; memset(&U, 0, sizeof(U));
    U.gbuf = G; U.P = P; U.NP = 0; U.nparams = 1;
    U.ps[0].off = 0; U.ps[0].numel = P; U.ps[0].shared_off = 0; U.ps[0].has_grad = 1;
    U.grads_out = shared_grad; U.diag = diag; U.task_mask = (1u << n_tasks) - 1u; U.n_tasks_max = n_tasks;
    U.alpha = c; U.max_norm = max_norm; U.do_sgd = 0; U.solver = solver;
    cagrad_update_kernel<<<1, UPD_THREADS, 0, (cudaStream_t)stream>>>(U);
    LAUNCH_CHECK();
    return 0;
}

// host-side evaluation of the simplex solve (same code the device runs): unit tests and A/B debugging
extern "C" int gaitk_cagrad_solve_host(const float* gram3x3, int n_tasks, float alpha, int solver, double* w_out, int* iters_out) {
    if (!gram3x3 || !w_out || n_tasks < 1 || n_tasks > MAXT) return fail(GAITK_E_BADARG, "bad argument");
    double c = 0; int it = 0;
    const int mode = cagrad_weights(gram3x3, n_tasks, alpha, solver, w_out, &c, &it);
    if (iters_out) *iters_out = it;
    return mode;
}

__global__ void sgd_flat_kernel(float* params, const float* grads, float* momentum, const UpdateArgs U) {
    for (int ip = 0; ip < U.nparams; ++ip) {
        const ParamSeg s = U.ps[ip];
        if (!s.has_grad) continue;
        for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < s.numel; e += gridDim.x * blockDim.x) {
            const float p = params[s.off + e];
            const float g = fmaf(U.wd, p, grads[s.off + e]);
            const float b = fmaf(U.mom, momentum[s.off + e], g);
            momentum[s.off + e] = b;
            params[s.off + e] = p - U.lr * b;
        }
    }
}
extern "C" int gaitk_sgd(gaitk_plan* pl, float* params, const float* grads, float* momentum, const uint8_t* has_grad,
                         float lr, float mom, float weight_decay, void* stream) {
    if (!pl || !params || !grads || !momentum) return fail(GAITK_E_BADARG, "null argument");
    UpdateArgs U; memset(&U, 0, sizeof(U)); fill_param_segs(pl, U, has_grad);
    U.lr = lr; U.mom = mom; U.wd = weight_decay;
    sgd_flat_kernel<<<8, 256, 0, (cudaStream_t)stream>>>(params, grads, momentum, U);
    LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------ losses

// one CTA; deterministic tree reduction
__global__ void __launch_bounds__(256) loss_kernel(const float* logits, const long long* y, int B, int K, LossArgs L,
                                                   const float* logit_off, float* loss_out, int* correct_out, float* dlogits) {
    __shared__ double sh[3][256];
    double num = 0, den = 0, cor = 0;
    for (int b = threadIdx.x; b < B; b += blockDim.x) den += (double)L.cls_w[(int)y[b]];
    sh[0][threadIdx.x] = den; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) sh[0][threadIdx.x] += sh[0][threadIdx.x + o]; __syncthreads(); }
    const float inv_denom = 1.0f / (float)sh[0][0];
    __syncthreads();
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        const int yy = (int)y[b];
        float zz[KMAX]; float mx = -INFINITY, best = -INFINITY; int am = 0;
        for (int k = 0; k < K; ++k) {
            float z = logits[(size_t)b * K + k];
            if (z > best) { best = z; am = k; }
            if (logit_off) z -= logit_off[(size_t)b * K + k];
            if (k == yy) z -= L.margin[k];
            z *= L.scale;
            if (L.nan_degenerate) z = __int_as_float(0x7fc00000);
            zz[k] = z; mx = fmaxf(mx, z);
        }
        float se = 0.f;
        for (int k = 0; k < K; ++k) se += expf(zz[k] - mx);
        const float lse = mx + logf(se);
        const float wy = L.cls_w[yy];
        num += (double)(wy * (lse - zz[yy]));
        cor += (am == yy) ? 1.0 : 0.0;
        if (dlogits)
            for (int k = 0; k < K; ++k)
                dlogits[(size_t)b * K + k] = L.scale * wy * inv_denom * (expf(zz[k] - lse) - (k == yy ? 1.f : 0.f));
    }
    sh[1][threadIdx.x] = num; sh[2][threadIdx.x] = cor; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) { sh[1][threadIdx.x] += sh[1][threadIdx.x + o]; sh[2][threadIdx.x] += sh[2][threadIdx.x + o]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (loss_out) loss_out[0] = (float)(sh[1][0] / sh[0][0]);
        if (correct_out) correct_out[0] = (int)(sh[2][0] + 0.5);
    }
}

extern "C" int gaitk_loss(const float* logits, const int64_t* y, int B, int K, const gaitk_loss_desc* desc, const float* logit_off,
                          float* loss_out, int32_t* correct_out, float* dlogits, void* stream) {
    if (!logits || !y || !desc || B <= 0 || K < 2 || K > KMAX) return fail(GAITK_E_BADARG, "bad argument");
    LossArgs L; L.scale = desc->scale; L.nan_degenerate = desc->nan_if_degenerate;
    for (int k = 0; k < KMAX; ++k) { L.margin[k] = desc->margin[k]; L.cls_w[k] = desc->cls_weight[k]; }
    loss_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(logits, (const long long*)y, B, K, L, logit_off, loss_out, correct_out, dlogits);
    LAUNCH_CHECK();
    return 0;
}

struct DenomArgs { const long long* y[GAITK_MAX_STREAMS]; int count[GAITK_MAX_STREAMS]; float cls_w[GAITK_MAX_STREAMS][KMAX]; int same_as[GAITK_MAX_STREAMS]; int n; };
// Class histogram (exact integers) -> sum_c count_c * w_c in fp64, per distinct label vector.  DENOM_CTAS CTAs per vector
// add their partial histograms with integer atomics (order-independent, so the result is deterministic); the last CTA to
// finish (ticket) turns the histogram into the denominators of every task that shares the vector.  Scratch (histograms +
// tickets) lives behind the four result floats in `denom` and is zeroed by the call.  The data-parallel step scans the
// labels of the GLOBAL batch here, so this must not be a single-CTA latency chain (it was: 21 us at 32 K labels, 170 us
// at 8 x 32 K -- most of the 8-GPU scaling loss).
constexpr int DENOM_CTAS = 32;
__global__ void __launch_bounds__(256) denom_kernel(DenomArgs D, float* denom) {
    __shared__ int hist[8][KMAX];
    __shared__ int last_s;
    const int s = blockIdx.y;
    if (D.same_as[s] != s) return;
    int* ghist = reinterpret_cast<int*>(denom + 4) + s * KMAX;
    int* ticket = reinterpret_cast<int*>(denom + 4) + GAITK_MAX_STREAMS * KMAX + s;
    int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
    const long long* y = D.y[s];
    const int n = D.count[s];
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < n; b += gridDim.x * blockDim.x) {
        const int yy = (int)y[b];
        c0 += yy == 0; c1 += yy == 1; c2 += yy == 2; c3 += yy == 3;
    }
    for (int o = 16; o > 0; o >>= 1) {
        c0 += __shfl_xor_sync(0xffffffffu, c0, o); c1 += __shfl_xor_sync(0xffffffffu, c1, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o); c3 += __shfl_xor_sync(0xffffffffu, c3, o);
    }
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5;
    if (lane == 0) { hist[wrp][0] = c0; hist[wrp][1] = c1; hist[wrp][2] = c2; hist[wrp][3] = c3; }
    __syncthreads();
    if (threadIdx.x < KMAX) {
        int h = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) h += hist[w][threadIdx.x];
        if (h) atomicAdd(ghist + threadIdx.x, h);
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last_s = atomicAdd(ticket, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (last_s && threadIdx.x == 0) {
        __threadfence();
        long long h[KMAX];
        for (int k = 0; k < KMAX; ++k) h[k] = *reinterpret_cast<volatile int*>(ghist + k);
        for (int t = 0; t < D.n; ++t) {
            if (D.same_as[t] != s) continue;
            double acc = 0;
            for (int k = 0; k < KMAX; ++k) acc += (double)h[k] * (double)D.cls_w[t][k];
            denom[t] = (float)acc;
            denom[GAITK_DENOM_COUNT + t] = (float)(h[0] + h[1] + h[2] + h[3]);   // labels in the GLOBAL batch (KL batchmean)
        }
    }
}
extern "C" int gaitk_loss_denominators(const int64_t* const* y, const int* counts, int n_streams, const gaitk_loss_desc* loss,
                                       float* denom, void* stream) {
    if (!y || !counts || !loss || !denom || n_streams < 1 || n_streams > GAITK_MAX_STREAMS) return fail(GAITK_E_BADARG, "bad argument");
    DenomArgs D; memset(&D, 0, sizeof(D)); D.n = n_streams;
    for (int s = 0; s < n_streams; ++s) {
        D.y[s] = (const long long*)y[s]; D.count[s] = counts[s]; D.same_as[s] = s;
        for (int t = 0; t < s; ++t) if (y[t] == y[s] && counts[t] == counts[s]) { D.same_as[s] = D.same_as[t]; break; }
        for (int k = 0; k < KMAX; ++k) D.cls_w[s][k] = loss[s].cls_weight[k];
    }
    static_assert(4 + GAITK_MAX_STREAMS * KMAX + GAITK_MAX_STREAMS <= GAITK_DENOM_COUNT && GAITK_DENOM_COUNT + GAITK_MAX_STREAMS <= GAITK_DENOM_FLOATS, "denominator scratch");
    CUDA_TRY(cudaMemsetAsync(denom + 4, 0, (GAITK_DENOM_FLOATS - 4) * sizeof(float), (cudaStream_t)stream));
    denom_kernel<<<dim3(DENOM_CTAS, n_streams), 256, 0, (cudaStream_t)stream>>>(D, denom);
    LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------ stages (fusion baselines)
// EarlyFusion3 / CheapXAttn3 (weargait_encoders.py:209-245, 338-387) and the 2-stream twins (feature_encoder.py:346-596) couple
// the streams BETWEEN encoder and backbone, so they run as stages with the intermediate tensors in HBM:
//   encoder stage (one stream kernel, encoder only) -> fusion op (concat / gaitk_xattn_*) -> trunk stage (ENC_NONE stream
//   kernel: backbone conv + ReLU + adaptive pooling) -> gaitk_linear_* head -> gaitk_loss.
// A stage is a one-stream gaitk_plan whose parameters are all private: its flat layout (gaitk_param_info) is
//   conv encoders  w1 b1 lng lnb | insole  w1 b1 w2 b2 lng lnb wsk bsk | linear-LN-ReLU  w1 b1 lng lnb | conv-pool  w1 b1 | trunk  wbb bbb
extern "C" int gaitk_stage_create(const gaitk_stage_desc* sd, int device, gaitk_plan** out) {
    if (!sd || !out) return fail(GAITK_E_BADARG, "null argument");
    *out = nullptr;
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(GAITK_E_ARCH, "device %d is sm_%d%d; libgaitk is built for sm_100a only (no fallback path)", device, prop.major, prop.minor);
    CUDA_TRY(cudaSetDevice(device));
    if (sd->T < 1 || sd->C < 1 || sd->CIN < 1 || sd->S < 1 || sd->bdim < 1 || sd->bdim > sd->T) return fail(GAITK_E_SHAPE, "bad stage dims");
    gaitk_plan* pl = new gaitk_plan();
    memset(&pl->d, 0, sizeof(pl->d));
    pl->d.family = GAITK_FAMILY_STAGE; pl->d.T = sd->T; pl->d.enc_out_ch = sd->enc == ENC_NONE ? sd->CIN : sd->C;
    pl->d.shared_out_ch = sd->S; pl->d.backbone_dim = sd->bdim; pl->d.num_classes = 2;
    pl->device = device; pl->sm_count = prop.multiProcessorCount; pl->NP = 0; pl->P = 0; pl->NF = sd->bdim * sd->S;
    pl->n_streams = 1; pl->p_wbb = pl->p_bbb = -1;
    for (int s = 0; s < GAITK_MAX_STREAMS; ++s) {
        StreamPlan& sp = pl->st[s];
        sp.p_w1 = sp.p_b1 = sp.p_w2 = sp.p_b2 = sp.p_wsk = sp.p_bsk = sp.p_lng = sp.p_lnb = sp.p_hng = sp.p_hnb = sp.p_hw = sp.p_hb = sp.p_wp = sp.p_bp = -1;
    }
    StreamPlan& e = pl->st[0];
    const int C = sd->C, CIN = sd->CIN;
    int KT1 = 1, H = 0;
    switch (sd->enc) {
        case ENC_CONV_GELU_LN:
            KT1 = 3;
            e.p_w1 = add_param(pl, "w1", 1, C, CIN, 3); e.p_b1 = add_param(pl, "b1", 1, C);
            e.p_lng = add_param(pl, "lng", 1, C); e.p_lnb = add_param(pl, "lnb", 1, C);
            break;
        case ENC_INSOLE:
            KT1 = 5; H = sd->H;
            e.p_w1 = add_param(pl, "w1", 1, H, CIN, 5); e.p_b1 = add_param(pl, "b1", 1, H);
            e.p_w2 = add_param(pl, "w2", 1, C, H, 3); e.p_b2 = add_param(pl, "b2", 1, C);
            e.p_lng = add_param(pl, "lng", 1, C); e.p_lnb = add_param(pl, "lnb", 1, C);
            e.p_wsk = add_param(pl, "wsk", 1, C, H, 1); e.p_bsk = add_param(pl, "bsk", 1, C);
            break;
        case ENC_LINEAR_LN_RELU:
            e.p_w1 = add_param(pl, "w1", 1, C, CIN); e.p_b1 = add_param(pl, "b1", 1, C);
            e.p_lng = add_param(pl, "lng", 1, C); e.p_lnb = add_param(pl, "lnb", 1, C);
            break;
        case ENC_CONV_POOL:
            KT1 = 3;
            e.p_w1 = add_param(pl, "w1", 1, C, CIN, 3); e.p_b1 = add_param(pl, "b1", 1, C);
            break;
        case ENC_NONE:
            pl->p_wbb = add_param(pl, "wbb", 1, sd->S, CIN, 3); pl->p_bbb = add_param(pl, "bbb", 1, sd->S);
            break;
        default: delete pl; return fail(GAITK_E_BADARG, "unknown stage kind %d", sd->enc);
    }
    int rc = plan_stream(pl, 0, sd->enc, CIN, KT1, H, sd->T_in > 0 ? sd->T_in : sd->T, sd->T, sd->pool_sensor);
    if (rc) { delete pl; return rc; }
    layout_stream_grads(pl, 0);
    *out = pl;
    return 0;
}
// out: encoder stage (B, T, C); trunk stage (B, NF).  x: (B, T_in, CIN) dense, or a frame store with win_start.
extern "C" int gaitk_stage_forward(gaitk_plan* pl, const float* params, const float* x, const int64_t* win_start, int B, int zero_input,
                                   float* out, void* stream) {
    if (!pl || !params || !x || !out || pl->d.family != GAITK_FAMILY_STAGE) return fail(GAITK_E_BADARG, "bad argument (stage plan expected)");
    if (B <= 0) return 0;
    StreamArgs a; fill_args(pl, 0, params, x, win_start, B, MODE_FWD, zero_input, a);
    if (pl->st[0].enc == ENC_NONE) a.repr_out = out; else a.feat_out = out;
    return launch_stream(pl, 0, a, stream_grid(pl, pl->st[0], B), (cudaStream_t)stream);
}
// dout: gradient of what gaitk_stage_forward returned.  grads: flat, stage layout, gaitk_param_total + 8 floats, accumulated (+=).
// dx (trunk stage only, optional): gradient of the trunk input (B, T, CIN).
extern "C" int gaitk_stage_backward(gaitk_plan* pl, const float* params, const float* x, const int64_t* win_start, int B, int zero_input,
                                    const float* dout, float* dx, float* grads, void* workspace, size_t workspace_bytes, void* stream) {
    if (!pl || !params || !x || !dout || !grads || !workspace || pl->d.family != GAITK_FAMILY_STAGE) return fail(GAITK_E_BADARG, "bad argument (stage plan expected)");
    if (workspace_bytes < gaitk_workspace_bytes(pl, B)) return fail(GAITK_E_BADARG, "workspace too small");
    if (B <= 0) return 0;
    StreamArgs a; fill_args(pl, 0, params, x, win_start, B, MODE_BWD_EXT, zero_input, a);
    if (pl->st[0].enc == ENC_NONE) { a.drepr_in = dout; a.dx = dx; } else a.dfeat_in = dout;
    a.partial = (float*)workspace;
    const int grid = stream_grid(pl, pl->st[0], B);
    int rc = launch_stream(pl, 0, a, grid, (cudaStream_t)stream);
    if (rc) return rc;
    return launch_reduce(pl, 0, (const float*)workspace, grid, grads, 0, 1.0f, -1, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------ data path
extern "C" int64_t gaitk_window_indices(int64_t n_frames, int64_t win, int64_t hop, int64_t* out, int64_t cap) {
    if (n_frames <= 0 || win <= 0 || hop <= 0 || n_frames < win) return 0;
    int64_t n = 0;
    for (int64_t w = 0; w + win <= n_frames; w += hop, ++n)
        if (out && n < cap) { out[3 * n] = n; out[3 * n + 1] = w; out[3 * n + 2] = w + win; }
    return n;
}

// per-channel sum / sum of squares / count over finite values.  One CTA per channel group keeps the
// accumulation order fixed (deterministic), threads stride over frames (coalesced across channels).
__global__ void __launch_bounds__(256) stats_kernel(const double* __restrict__ f, long long N, int D, double* acc) {
    __shared__ double sh[3][256];
    const int d = blockIdx.x;
    double s = 0, ss = 0, n = 0;
    for (long long i = threadIdx.x; i < N; i += blockDim.x) {
        const double v = f[i * D + d];
        if (isfinite(v)) { s += v; ss = fma(v, v, ss); n += 1.0; }
    }
    sh[0][threadIdx.x] = s; sh[1][threadIdx.x] = ss; sh[2][threadIdx.x] = n; __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) for (int j = 0; j < 3; ++j) sh[j][threadIdx.x] += sh[j][threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) { acc[d] += sh[0][0]; acc[D + d] += sh[1][0]; acc[2 * D + d] += sh[2][0]; }
}
extern "C" int gaitk_stats_accumulate(const double* frames, int64_t N, int D, double* acc, void* stream) {
    if (!frames || !acc || D <= 0) return fail(GAITK_E_BADARG, "bad argument");
    if (N <= 0) return 0;
    stats_kernel<<<D, 256, 0, (cudaStream_t)stream>>>(frames, N, D, acc);
    LAUNCH_CHECK();
    return 0;
}
__global__ void stats_finalize_kernel(const double* acc, int D, double* mean, double* stdv) {
    const int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= D) return;
    const double n = acc[2 * D + d];
    if (n <= 0) { mean[d] = 0.0; stdv[d] = -1.0; return; }        // no statistics: channel passes through
    const double m = acc[d] / n;
    double var = acc[D + d] / n - m * m; if (var < 0) var = 0;
    double sd = sqrt(var); if (sd < 1e-6) sd = 1e-6;
    mean[d] = m; stdv[d] = sd;
}
extern "C" int gaitk_stats_finalize(const double* acc, int D, double* mean, double* stdv, void* stream) {
    if (!acc || !mean || !stdv || D <= 0) return fail(GAITK_E_BADARG, "bad argument");
    stats_finalize_kernel<<<(D + 63) / 64, 64, 0, (cudaStream_t)stream>>>(acc, D, mean, stdv);
    LAUNCH_CHECK();
    return 0;
}
__global__ void normalize_kernel(const double* __restrict__ f, long long total, int D, const double* __restrict__ mean,
                                 const double* __restrict__ stdv, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int d = (int)(i % D);
        double v = f[i];
        const double s = stdv[d];
        if (s < 0) { out[i] = (float)v; continue; }                 // channel without statistics
        const double m = mean[d];
        const double mm = isfinite(m) ? m : 0.0;
        if (!isfinite(v)) v = mm;
        const double se = (isfinite(s) && s > 1e-6) ? s : 1e-6;
        double z = __ddiv_rn(__dsub_rn(v, mm), se);
        if (!isfinite(z)) z = 0.0;
        out[i] = (float)z;
    }
}
extern "C" int gaitk_normalize_frames(const double* frames, int64_t N, int D, const double* mean, const double* stdv, float* out, void* stream) {
    if (!frames || !mean || !stdv || !out || D <= 0) return fail(GAITK_E_BADARG, "bad argument");
    if (N <= 0) return 0;
    const long long total = (long long)N * D;
    const int grid = (int)std::min<long long>((total + 255) / 256, 148 * 16);
    normalize_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(frames, total, D, mean, stdv, out);
    LAUNCH_CHECK();
    return 0;
}
// one window per (blockIdx.x + k*gridDim.x); the T*D floats of a window are contiguous in both source and destination
__global__ void __launch_bounds__(256) gather_kernel(const float* __restrict__ frames, int D, const long long* __restrict__ ws,
                                                     int B, int T, int enabled, float* __restrict__ out) {
    const int per = T * D;
    for (int b = blockIdx.x; b < B; b += gridDim.x) {
        const float* src = frames + (size_t)ws[b] * D;
        float* dst = out + (size_t)b * per;
        const bool vec = ((((uintptr_t)src | (uintptr_t)dst) & 15) == 0) && (per % 4 == 0);
        if (!enabled) {
            for (int i = threadIdx.x; i < per; i += blockDim.x) dst[i] = 0.f;
        } else if (vec) {
            const float4* s4 = reinterpret_cast<const float4*>(src); float4* d4 = reinterpret_cast<float4*>(dst);
            for (int i = threadIdx.x; i < per / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
        } else {
            for (int i = threadIdx.x; i < per; i += blockDim.x) dst[i] = __ldg(src + i);
        }
    }
}
extern "C" int gaitk_window_gather(const float* frames, int D, const int64_t* win_start, int B, int T, int enabled, float* out, void* stream) {
    if (!frames || !win_start || !out || D <= 0 || T <= 0) return fail(GAITK_E_BADARG, "bad argument");
    if (B <= 0) return 0;
    gather_kernel<<<std::min(B, 148 * 8), 256, 0, (cudaStream_t)stream>>>(frames, D, (const long long*)win_start, B, T, enabled, out);
    LAUNCH_CHECK();
    return 0;
}
// Relaxed-input evaluation: all seven presence masks of MASK_COMBOS (weargait_train.py:49-57) from ONE set of logits.
// eval_with_mask (:391-433) only ever reads the logits of the ENABLED streams, and a stream's logits depend on its own
// input alone, so the zero-filled forward passes of the reference add nothing: per window the three softmax rows are
// formed once, the seven ensembles (mean of the enabled rows, first-index argmax) are compared with the label, and the
// per-stream argmax hits (the async flavour and eval_one_epoch :322-350) are counted alongside.
// counts[0..6] = ensemble hits per mask in MASK_COMBOS order, counts[7..9] = per-stream hits (own labels), int32.
__global__ void __launch_bounds__(256) mask_eval_kernel(const float* __restrict__ lw, const float* __restrict__ li,
                                                        const float* __restrict__ lm, const long long* __restrict__ yw,
                                                        const long long* __restrict__ yi, const long long* __restrict__ ym,
                                                        int B, int K, int* __restrict__ counts) {
    int c[10];
#pragma unroll
    for (int i = 0; i < 10; ++i) c[i] = 0;
    for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < B; b += gridDim.x * blockDim.x) {
        const float* L[3] = {lw + (size_t)b * K, li + (size_t)b * K, lm + (size_t)b * K};
        const long long Y[3] = {yw[b], yi[b], ym[b]};
        float p[3][KMAX];
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            float z[KMAX], mx = -INFINITY; int am = 0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) { z[k] = k < K ? L[s][k] : -INFINITY; if (z[k] > mx) { mx = z[k]; am = k; } }
            c[7 + s] += (am == (int)Y[s]);
            float e[KMAX];
#pragma unroll
            for (int k = 0; k < KMAX; ++k) e[k] = k < K ? expf(z[k] - mx) : 0.f;
            // butterfly order of ATen's warp softmax (persistent_softmax.cuh): (e0 + e2) + (e1 + e3)
            const float sum = K <= 2 ? e[0] + e[1] : (e[0] + e[2]) + (e[1] + e[3]);
#pragma unroll
            for (int k = 0; k < KMAX; ++k) p[s][k] = e[k] / sum;
        }
#pragma unroll
        for (int m = 0; m < 7; ++m) {
            // MASK_COMBOS order: W, I, M, W+I, W+M, I+M, W+I+M
            const bool uw = (m == 0 || m == 3 || m == 4 || m == 6), ui = (m == 1 || m == 3 || m == 5 || m == 6),
                       um = (m == 2 || m == 4 || m == 5 || m == 6);
            const int n = (int)uw + (int)ui + (int)um;
            const float inv = 1.0f / (float)n;                 // tensor / python-int on CUDA multiplies by the reciprocal
            float best = -INFINITY; int am = 0;
#pragma unroll
            for (int k = 0; k < KMAX; ++k) {
                if (k >= K) break;
                float a = 0.f;                                  // sum(probs): 0 + p_first + ...
                if (uw) a = a + p[0][k];
                if (ui) a = a + p[1][k];
                if (um) a = a + p[2][k];
                a = a * inv;
                if (a > best) { best = a; am = k; }
            }
            c[m] += (am == (int)Y[0]);                          // sync: one label per window (yw)
        }
    }
#pragma unroll
    for (int i = 0; i < 10; ++i) {
        int v = c[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0 && v) atomicAdd(counts + i, v);
    }
}
extern "C" int gaitk_mask_eval(const float* const* logits, const int64_t* const* y, int B, int K, int32_t* counts, void* stream) {
    if (!logits || !y || !counts || !logits[0] || !logits[1] || !logits[2] || !y[0] || !y[1] || !y[2])
        return fail(GAITK_E_BADARG, "null argument");
    if (K < 2 || K > KMAX) return fail(GAITK_E_SHAPE, "num_classes must be in [2,%d]", KMAX);
    if (B <= 0) return 0;
    mask_eval_kernel<<<std::min((B + 255) / 256, 148 * 4), 256, 0, (cudaStream_t)stream>>>(
        logits[0], logits[1], logits[2], (const long long*)y[0], (const long long*)y[1], (const long long*)y[2], B, K, counts);
    LAUNCH_CHECK();
    return 0;
}
// FoG pose clip: subtract joint 0, per-coordinate min-max over (L, J) of the CENTRED clip, pad/trim to T_out.
// One CTA per clip.  fp64 arithmetic in the reference's operation order, then cast to fp32.
__global__ void __launch_bounds__(128) fog_pose_kernel(const double* __restrict__ poses, const long long* __restrict__ cs,
                                                       const long long* __restrict__ cl, int J, int T_out, float* __restrict__ out) {
    __shared__ double smin[3][128], smax[3][128];
    const int c = blockIdx.x;
    const double* p = poses + (size_t)cs[c] * J * 3;
    const long long L = cl[c];
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (long long i = threadIdx.x; i < L * J; i += blockDim.x) {
        const long long t = i / J;
        for (int k = 0; k < 3; ++k) {
            const double v = p[i * 3 + k] - p[t * J * 3 + k];
            mn[k] = fmin(mn[k], v); mx[k] = fmax(mx[k], v);
        }
    }
    for (int k = 0; k < 3; ++k) { smin[k][threadIdx.x] = mn[k]; smax[k][threadIdx.x] = mx[k]; }
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
        if (threadIdx.x < o) for (int k = 0; k < 3; ++k) {
            smin[k][threadIdx.x] = fmin(smin[k][threadIdx.x], smin[k][threadIdx.x + o]);
            smax[k][threadIdx.x] = fmax(smax[k][threadIdx.x], smax[k][threadIdx.x + o]);
        }
        __syncthreads();
    }
    float* o_ = out + (size_t)c * T_out * J * 3;
    for (long long i = threadIdx.x; i < (long long)T_out * J; i += blockDim.x) {
        const long long t = i / J;
        for (int k = 0; k < 3; ++k) {
            double v = 0.0;
            if (t < L) {
                const double cen = __dsub_rn(p[i * 3 + k], p[t * J * 3 + k]);
                v = __ddiv_rn(__dsub_rn(cen, smin[k][0]), __dadd_rn(__dsub_rn(smax[k][0], smin[k][0]), 1e-6));
            }
            o_[i * 3 + k] = (float)v;
        }
    }
}
extern "C" int gaitk_fog_prepare_pose(const double* poses, const int64_t* clip_start, const int64_t* clip_len, int n_clips, int J,
                                      int T_out, float* out, void* stream) {
    if (!poses || !clip_start || !clip_len || !out || J <= 0 || T_out <= 0) return fail(GAITK_E_BADARG, "bad argument");
    if (n_clips <= 0) return 0;
    fog_pose_kernel<<<n_clips, 128, 0, (cudaStream_t)stream>>>(poses, (const long long*)clip_start, (const long long*)clip_len, J, T_out, out);
    LAUNCH_CHECK();
    return 0;
}
__global__ void __launch_bounds__(128) fog_sensor_kernel(const double* __restrict__ sens, const long long* __restrict__ cs,
                                                         const long long* __restrict__ cl, int D, int T_out, float* __restrict__ out) {
    const int c = blockIdx.x;
    const double* p = sens + (size_t)cs[c] * D;
    const long long L = cl[c];
    float* o_ = out + (size_t)c * T_out * D;
    for (long long i = threadIdx.x; i < (long long)T_out * D; i += blockDim.x) o_[i] = (i / D < L) ? (float)p[i] : 0.f;
}
extern "C" int gaitk_fog_prepare_sensor(const double* sens, const int64_t* clip_start, const int64_t* clip_len, int n_clips, int D,
                                        int T_out, float* out, void* stream) {
    if (!sens || !clip_start || !clip_len || !out || D <= 0 || T_out <= 0) return fail(GAITK_E_BADARG, "bad argument");
    if (n_clips <= 0) return 0;
    fog_sensor_kernel<<<n_clips, 128, 0, (cudaStream_t)stream>>>(sens, (const long long*)clip_start, (const long long*)clip_len, D, T_out, out);
    LAUNCH_CHECK();
    return 0;
}

// ------------------------------------------------------------------------------------------ tcgen05 self-test
extern "C" int gaitk_umma_selftest(const float* A, int nA, const float* B, int nB, const uint32_t* ops, int nops, int ncols,
                                   float* D, void* stream) {
    if (!A || !B || !ops || !D || nops < 1 || (ncols != 32 && ncols != 64 && ncols != 128 && ncols != 256)) return fail(GAITK_E_BADARG, "bad argument");
    const size_t smem = ((size_t)((nA + 255) / 256) * 256 + nB + 64) * sizeof(float);
    if (smem > 200 * 1024) return fail(GAITK_E_SHAPE, "operands too large");
    CUDA_TRY(cudaFuncSetAttribute((const void*)umma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, nA, B, nB, (const UmmaOp*)ops, nops, ncols, D);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gaitk_umma_selftest_bf16(const uint16_t* A, int nA, const uint16_t* B, int nB, const uint32_t* ops, int nops, int ncols,
                                        float* D, void* stream) {
    if (!A || !B || !ops || !D || nops < 1 || (ncols != 32 && ncols != 64 && ncols != 128 && ncols != 256)) return fail(GAITK_E_BADARG, "bad argument");
    const size_t smem = ((size_t)((nA + 511) / 512) * 512 + nB + 128) * sizeof(uint16_t);
    if (smem > 200 * 1024) return fail(GAITK_E_SHAPE, "operands too large");
    CUDA_TRY(cudaFuncSetAttribute((const void*)umma_selftest_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    umma_selftest_bf16_kernel<<<1, 128, smem, (cudaStream_t)stream>>>(A, nA, B, nB, (const UmmaOp*)ops, nops, ncols, D);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gaitk_umma_bench(const uint32_t* ops, int nops, int reps, int ncols, int smem_bytes, int64_t* cycles, void* stream) {
    if (!ops || !cycles || nops < 1 || reps < 1 || smem_bytes < 1024 || smem_bytes > 200 * 1024) return fail(GAITK_E_BADARG, "bad argument");
    CUDA_TRY(cudaFuncSetAttribute((const void*)umma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    umma_bench_kernel<<<1, 128, smem_bytes, (cudaStream_t)stream>>>((const UmmaOp*)ops, nops, reps, ncols, smem_bytes, (long long*)cycles);
    LAUNCH_CHECK();
    return 0;
}

extern "C" int gaitk_umma_bench_multi(const uint32_t* ops, int nops, int reps, int ncols, int smem_bytes, int n_issuers, int64_t* cycles, void* stream) {
    if (!ops || !cycles || nops < 1 || reps < 1 || smem_bytes < 1024 || smem_bytes > 200 * 1024 || n_issuers < 1 || n_issuers > 4)
        return fail(GAITK_E_BADARG, "bad argument");
    CUDA_TRY(cudaFuncSetAttribute((const void*)umma_bench_multi_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    umma_bench_multi_kernel<<<1, 128, smem_bytes, (cudaStream_t)stream>>>((const UmmaOp*)ops, nops, reps, ncols, smem_bytes, n_issuers, (long long*)cycles);
    LAUNCH_CHECK();
    return 0;
}
