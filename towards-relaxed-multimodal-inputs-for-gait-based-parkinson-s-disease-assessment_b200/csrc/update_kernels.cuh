// update_kernels.cuh -- cross-CTA gradient reduction, on-device CAGrad solve + clip, SGD.
//
// Reference semantics (paths relative to the reference root):
//   CAGrad.cagrad / overwrite_grad / backward   train/learning/optimizers/multitask_weighting.py:694-776
//   clip_grad_norm_(shared, max_norm)           :775
//   private-gradient multiplicity               train/weargait_train.py:218-242, train/fbg_fog_train.py:146-152
//   SGD(momentum 0.9, weight_decay 1e-4)        train/weargait_train.py:560, train/fbg_fog_train.py:288
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cagrad_solver.cuh"

namespace gaitk {

constexpr int MAX_SEG = 16;
constexpr int MAX_PARAMS = 40;
constexpr int MAXT = 3;
// diag[16]: STICKY exchange status of the data-parallel step -- 0 = ok, -1 = a peer never arrived.  Written only by the
// exchange kernel; cagrad_update_kernel reads it and refuses to touch parameters / momentum after a failed exchange
// (replicas must not silently diverge); the host resets it (FusedTrainStep.reset_exchange_status).
constexpr int DIAG_EXCHANGE = 16;
constexpr int DIAG_FLOATS = 24;

// one stream-local gradient segment and where it goes
struct Seg { int src, len, shared_off, param_off; };     // shared_off >= 0 -> G column; else private
struct ReduceArgs {
    const float* partial; int grid, NGP, NG;
    int nseg; Seg seg[MAX_SEG];
    float* gbuf; int P; long long NP;                    // gbuf = [G (MAXT x P) | PG (NP) | loss[4] | correct[4]]
    int task; float private_mult; int stat_slot;         // stat_slot < 0: do not record loss/correct
};

// sum the per-CTA partial rows in fixed order (8 row slices per output, combined in slice order), then add
// into gbuf (passes run in launch order).  blockDim = (128, 8).
__device__ __forceinline__ void reduce_partials_body(const ReduceArgs& R) {
    __shared__ float sh[8][128];
    const int e = blockIdx.x * 128 + threadIdx.x, sl = threadIdx.y;
    float s0 = 0.f, s1 = 0.f;
    if (e < R.NG + 2) {
        const float* p = R.partial + e;
        int g = sl;
        for (; g + 8 < R.grid; g += 16) { s0 += p[(size_t)g * R.NGP]; s1 += p[(size_t)(g + 8) * R.NGP]; }
        if (g < R.grid) s0 += p[(size_t)g * R.NGP];
    }
    sh[sl][threadIdx.x] = s0 + s1;
    __syncthreads();
    if (sl != 0 || e >= R.NG + 2) return;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += sh[i][threadIdx.x];
    float* G = R.gbuf; float* PG = R.gbuf + (size_t)MAXT * R.P; float* ST = PG + R.NP;
    if (e >= R.NG) {
        if (R.stat_slot >= 0) ST[(e - R.NG) * 4 + R.stat_slot] += s;
        return;
    }
    for (int i = 0; i < R.nseg; ++i) {
        const Seg sg = R.seg[i];
        if (e >= sg.src && e < sg.src + sg.len) {
            const int o = e - sg.src;
            if (sg.shared_off >= 0) G[(size_t)R.task * R.P + sg.shared_off + o] += s;
            else PG[sg.param_off + o] += R.private_mult * s;
            return;
        }
    }
}
__global__ void __launch_bounds__(1024) reduce_partials_kernel(const ReduceArgs R) { reduce_partials_body(R); }
// all streams of a fused step in ONE launch (blockIdx.y = stream): every output element of gbuf receives exactly one
// contribution there (task column = stream, private segments and stat slots are disjoint), so the order is immaterial
struct ReduceArgsMulti { ReduceArgs r[MAXT]; };
__global__ void __launch_bounds__(1024) reduce_partials_multi_kernel(const ReduceArgsMulti M) {
    const ReduceArgs& R = M.r[blockIdx.y];
    if ((int)blockIdx.x * 128 >= R.NG + 2) return;
    reduce_partials_body(R);
}

struct ParamSeg { long long off; int numel; int shared_off; int has_grad; };
struct UpdateArgs {
    float* params; float* momentum; const float* gbuf; float* grads_out; float* diag;
    int P; long long NP; int nparams; ParamSeg ps[MAX_PARAMS];
    unsigned task_mask; int n_tasks_max;
    float alpha, max_norm, lr, mom, wd;
    int do_sgd, solver;
    int check_exchange;                // diag has DIAG_FLOATS entries and diag[DIAG_EXCHANGE] gates the update
};

// single CTA: Gram -> solve -> combine -> clip -> (optional) SGD over the flat parameter buffer.
// Every thread first loads "its" parameter / momentum / gradient elements (<= EPT each), so that global
// latency overlaps the latency-bound simplex solve that warp 0 runs in between.
constexpr int UPD_THREADS = 1024;
constexpr int UPD_EPT = 8;                                  // elements per thread: NP <= 8192 per pass
__global__ void __launch_bounds__(UPD_THREADS) cagrad_update_kernel(const UpdateArgs U) {
    __shared__ double shd[6][32];
    __shared__ double coef[MAXT];
    __shared__ int tl[MAXT];
    __shared__ int nt_s;
    const int tid = threadIdx.x, nth = blockDim.x, lane = tid & 31, wrp = tid >> 5;
#ifdef GAITK_UPDATE_TIMING          // phase clocks -> diag[17..19] (scratch/update_timing.py)
    const long long tk0 = clock64(); long long tk1 = 0, tk2 = 0;
#endif
    const float* G = U.gbuf; const float* PG = U.gbuf + (size_t)MAXT * U.P;
    // a failed data-parallel exchange (a peer never published its gradients) must not be applied: gbuf is incomplete
    if (U.check_exchange && U.diag && U.diag[DIAG_EXCHANGE] < 0.f) return;
    if (tid == 0) {
        int n = 0;
        for (int t = 0; t < U.n_tasks_max; ++t) if (U.task_mask & (1u << t)) tl[n++] = t;
        nt_s = n;
    }
    __syncthreads();
    const int n = nt_s;
    const int t0 = tl[0], t1 = n > 1 ? tl[1] : tl[0], t2 = n > 2 ? tl[2] : tl[0];
    // ---- Gram matrix: one pass, one combined block reduction (fp64)
    double gg[6] = {0, 0, 0, 0, 0, 0};     // 00 01 02 11 12 22
    for (int p = tid; p < U.P; p += nth) {
        const double v0 = (double)G[(size_t)t0 * U.P + p];
        const double v1 = n > 1 ? (double)G[(size_t)t1 * U.P + p] : 0.0;
        const double v2 = n > 2 ? (double)G[(size_t)t2 * U.P + p] : 0.0;
        gg[0] += v0 * v0; gg[1] += v0 * v1; gg[2] += v0 * v2; gg[3] += v1 * v1; gg[4] += v1 * v2; gg[5] += v2 * v2;
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        for (int o = 16; o > 0; o >>= 1) gg[i] += __shfl_xor_sync(0xffffffffu, gg[i], o);
        if (lane == 0) shd[i][wrp] = gg[i];
    }
    // ---- this thread's elements of the flat parameter buffer (loads in flight during the solve)
    const long long total = U.NP > 0 ? U.NP : (U.nparams ? U.ps[0].numel : 0);      // gaitk_cagrad: one pseudo-parameter
    const int passes = (int)((total + (long long)nth * UPD_EPT - 1) / ((long long)nth * UPD_EPT));
    __syncthreads();
#ifdef GAITK_UPDATE_TIMING
    tk1 = clock64();
#endif
    if (tid < 32) {
        // warp 0 runs the solve (all lanes redundantly; the QP enumeration is lane-parallel)
        double g6[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) { double s = 0; for (int w = 0; w < (nth >> 5); ++w) s += shd[i][w]; g6[i] = s; }
        // the reference forms GG in fp32 (torch mm) and hands the fp32 values to SciPy
        const float a[3][3] = {{(float)g6[0], (float)g6[1], (float)g6[2]}, {(float)g6[1], (float)g6[3], (float)g6[4]},
                               {(float)g6[2], (float)g6[4], (float)g6[5]}};
        Quad3 q;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) q.A[i][j] = (double)a[i][j];
        double w[3] = {0, 0, 0}; double c = 0; int iters = 0;
        const bool plain_mean = U.solver == 2;             // GAITK_SOLVER_MEAN: torch.stack(losses).mean().backward()
        if (plain_mean) { for (int i = 0; i < n; ++i) w[i] = 1.0 / n; }
        else cagrad_weights(&a[0][0], n, U.alpha, U.solver, w, &c, &iters);
        q.c = c;
        for (int i = 0; i < n; ++i) { double rs = 0; for (int j = 0; j < n; ++j) rs += q.A[i][j]; q.Ab[i] = rs / n; }
        // gw = G w with w cast to fp32 as torch.Tensor(w_cpu) does (:719)
        for (int i = 0; i < n; ++i) w[i] = (double)(float)w[i];
        double gw2 = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) gw2 += w[i] * q.A[i][j] * w[j];
        const double lam = c / (sqrt(gw2 > 0 ? gw2 : 0) + 1e-8);
        // g = n * (mean_i G_i + lam * sum_i w_i G_i) / (1 + alpha^2)  = sum_i k_i G_i
        double k[3], norm2 = 0;
        for (int i = 0; i < n; ++i)
            k[i] = plain_mean ? 1.0 / n : (double)n * (1.0 / n + lam * w[i]) / (1.0 + (double)U.alpha * U.alpha);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) norm2 += k[i] * q.A[i][j] * k[j];
        const double norm = sqrt(norm2 > 0 ? norm2 : 0);
        double clip = 1.0;
        if (U.max_norm > 0) { clip = (double)U.max_norm / (norm + 1e-6); if (clip > 1.0) clip = 1.0; }
        if (tid == 0) for (int i = 0; i < MAXT; ++i) coef[i] = i < n ? k[i] * clip : 0.0;
        if (U.diag && tid == 0) {
            for (int i = 0; i < 3; ++i) U.diag[i] = 0.f;
            for (int i = 0; i < n; ++i) U.diag[tl[i]] = (float)w[i];
            for (int i = 0; i < 9; ++i) U.diag[3 + i] = 0.f;
            for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) U.diag[3 + tl[i] * 3 + tl[j]] = a[i][j];
            U.diag[12] = (float)norm; U.diag[13] = (float)cg_obj(q, w, n); U.diag[14] = (float)iters; U.diag[15] = (float)clip;
        }
    }
    for (int pass = 0; pass < passes; ++pass) {
        // gather phase (before the barrier that publishes the coefficients): element -> segment, loads
        float gsh0[UPD_EPT], gsh1[UPD_EPT], gsh2[UPD_EPT], pv[UPD_EPT], mv[UPD_EPT];
        int kind[UPD_EPT];                                  // 0 none, 1 shared, 2 private ; bit 2: has_grad
#pragma unroll
        for (int j = 0; j < UPD_EPT; ++j) {
            const long long e = (long long)pass * nth * UPD_EPT + (long long)j * nth + tid;
            kind[j] = 0; gsh0[j] = gsh1[j] = gsh2[j] = 0.f; pv[j] = mv[j] = 0.f;
            if (e >= total) continue;
            int ip = 0;
            for (int i = 1; i < U.nparams; ++i) if (e >= U.ps[i].off) ip = i;       // segments are sorted by offset
            const ParamSeg s = U.ps[ip];
            const int o = (int)(e - s.off);
            if (s.shared_off >= 0) {
                const size_t p = (size_t)s.shared_off + o;
                gsh0[j] = G[(size_t)t0 * U.P + p];
                if (n > 1) gsh1[j] = G[(size_t)t1 * U.P + p];
                if (n > 2) gsh2[j] = G[(size_t)t2 * U.P + p];
                kind[j] = 1;
            } else {
                gsh0[j] = PG[e]; kind[j] = 2;
            }
            if (s.has_grad) kind[j] |= 4;
            if (U.do_sgd && s.has_grad) { pv[j] = U.params[e]; mv[j] = U.momentum[e]; }
        }
        if (pass == 0) __syncthreads();                    // coefficients from the solve
#ifdef GAITK_UPDATE_TIMING
        if (pass == 0) tk2 = clock64();
#endif
        const float k0 = (float)coef[0], k1 = (float)coef[1], k2 = (float)coef[2];
#pragma unroll
        for (int j = 0; j < UPD_EPT; ++j) {
            const long long e = (long long)pass * nth * UPD_EPT + (long long)j * nth + tid;
            if (e >= total || kind[j] == 0) continue;
            float g = gsh0[j];
            if ((kind[j] & 3) == 1) { g = k0 * gsh0[j]; if (n > 1) g = fmaf(k1, gsh1[j], g); if (n > 2) g = fmaf(k2, gsh2[j], g); }
            const bool hg = (kind[j] & 4) != 0;
            if (U.grads_out) U.grads_out[e] = hg ? g : 0.f;
            if (U.do_sgd && hg) {
                g = fmaf(U.wd, pv[j], g);
                const float b = fmaf(U.mom, mv[j], g);
                U.momentum[e] = b;
                U.params[e] = pv[j] - U.lr * b;
            }
        }
    }
#ifdef GAITK_UPDATE_TIMING
    if (tid == 0 && U.diag) { U.diag[17] = (float)(tk1 - tk0); U.diag[18] = (float)(tk2 - tk1); U.diag[19] = (float)(clock64() - tk2); }
#endif
}

// ------------------------------------------------------------------------------------------------------------
// Data-parallel exchange fused with the update: one-shot all-reduce of gbuf over NVLink peer memory.
// Every rank's gbuf lives in a symmetric allocation (the same buffer is mapped by all ranks of the node).  After its
// reduce kernels, a rank publishes `step` in its own flag word (release, system scope); every CTA of this kernel
// waits until all peers have published (acquire), then sums the N buffers IN RANK ORDER into the local `gsum` -- the
// same order on every rank, so replicas stay bit-identical -- and the ordinary single-CTA update runs on gsum.  No NCCL
// call, no host involvement; the whole step stays one CUDA graph.  gbuf is double-buffered by step parity: a rank can
// start writing buffer (k+2)&1 only after its step-(k+1) exchange, which completes only after every peer published
// step k+1, i.e. after every peer finished reading buffer k&1.
struct P2PArgs {
    const float* const* peer_gbuf;     // device array [world]: gbuf of this step's parity on every rank (peer-mapped)
    unsigned* const* peer_flag;        // device array [world]: flag word of every rank
    const unsigned* counter;           // local: number of completed exchanges (step = counter + 1)
    float* gsum; int n; int rank, world; float* diag;
    long long timeout_cycles;          // give up waiting for a peer after this many SM clocks
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v; asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory"); return v;
}
constexpr int P2P_THREADS = 256;
__global__ void __launch_bounds__(P2P_THREADS) p2p_allreduce_kernel(const P2PArgs Q) {
    __shared__ int ok_s;
    const unsigned step = *Q.counter + 1u;
    if (threadIdx.x == 0) ok_s = 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) { __threadfence_system(); st_release_sys(Q.peer_flag[Q.rank], step); }
    __syncthreads();
    if ((int)threadIdx.x < Q.world) {
        const unsigned* f = Q.peer_flag[threadIdx.x];
        const long long t0 = clock64();
        // step numbers only grow; the comparison is wrap-safe.  A peer that never arrives (crashed rank) must not hang
        // the GPU: give up after ~4 s of SM clocks and flag the step as failed (diag[15] = -1).
        while ((int)(ld_acquire_sys(f) - step) < 0) {
            if (clock64() - t0 > Q.timeout_cycles) { ok_s = 0; break; }
            __nanosleep(64);
        }
    }
    __syncthreads();
    if (!ok_s && Q.diag && threadIdx.x == 0) Q.diag[DIAG_EXCHANGE] = -1.0f;       // every CTA that saw it (same value)
    // all remote loads of an element are issued before the first one is consumed (rank order is kept in the sum)
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < Q.n; e += gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int r0 = 0; r0 < Q.world; r0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = r0 + i < Q.world ? ld_relaxed_sys(Q.peer_gbuf[r0 + i] + e) : 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s += v[i];
        }
        Q.gsum[e] = s;
    }
}
__global__ void bump_counter_kernel(unsigned* counter) { *counter += 1u; }

}  // namespace gaitk
