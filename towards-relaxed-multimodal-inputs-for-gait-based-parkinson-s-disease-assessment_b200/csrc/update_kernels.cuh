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

// one stream-local gradient segment and where it goes
struct Seg { int src, len, shared_off, param_off; };     // shared_off >= 0 -> G column; else private
struct ReduceArgs {
    const float* partial; int grid, NGP, NG;
    int nseg; Seg seg[MAX_SEG];
    float* gbuf; int P; long long NP;                    // gbuf = [G (MAXT x P) | PG (NP) | loss[4] | correct[4]]
    int task; float private_mult; int stat_slot;         // stat_slot < 0: do not record loss/correct
};

// sum the per-CTA partial rows in fixed order, then add into gbuf (passes run in launch order)
__global__ void reduce_partials_kernel(const ReduceArgs R) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= R.NG + 2) return;
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
    const float* p = R.partial + e;
    int g = 0;
    for (; g + 3 < R.grid; g += 4) {
        s0 += p[(size_t)g * R.NGP]; s1 += p[(size_t)(g + 1) * R.NGP];
        s2 += p[(size_t)(g + 2) * R.NGP]; s3 += p[(size_t)(g + 3) * R.NGP];
    }
    for (; g < R.grid; ++g) s0 += p[(size_t)g * R.NGP];
    const float s = (s0 + s1) + (s2 + s3);
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

struct ParamSeg { long long off; int numel; int shared_off; int has_grad; };
struct UpdateArgs {
    float* params; float* momentum; const float* gbuf; float* grads_out; float* diag;
    int P; long long NP; int nparams; ParamSeg ps[MAX_PARAMS];
    unsigned task_mask; int n_tasks_max;
    float alpha, max_norm, lr, mom, wd;
    int do_sgd, solver;
};

// block-wide sum of doubles (blockDim.x multiple of 32, <= 1024)
__device__ inline double block_sum_d(double v, double* sh) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, wrp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) sh[wrp] = v;
    __syncthreads();
    double s = 0;
    for (int i = 0; i < nw; ++i) s += sh[i];
    return s;
}

// single CTA: Gram -> solve -> combine -> clip -> (optional) SGD over the flat parameter buffer
__global__ void __launch_bounds__(1024) cagrad_update_kernel(const UpdateArgs U) {
    __shared__ double shd[32];
    __shared__ double coef[MAXT];
    __shared__ int tl[MAXT];
    __shared__ int nt_s;
    const int tid = threadIdx.x, nth = blockDim.x;
    const float* G = U.gbuf; const float* PG = U.gbuf + (size_t)MAXT * U.P;
    if (tid == 0) {
        int n = 0;
        for (int t = 0; t < U.n_tasks_max; ++t) if (U.task_mask & (1u << t)) tl[n++] = t;
        nt_s = n;
    }
    __syncthreads();
    const int n = nt_s;
    double gg[6] = {0, 0, 0, 0, 0, 0};     // 00 01 02 11 12 22
    for (int p = tid; p < U.P; p += nth) {
        double v[MAXT];
        for (int i = 0; i < MAXT; ++i) v[i] = i < n ? (double)G[(size_t)tl[i] * U.P + p] : 0.0;
        gg[0] += v[0] * v[0]; gg[1] += v[0] * v[1]; gg[2] += v[0] * v[2];
        gg[3] += v[1] * v[1]; gg[4] += v[1] * v[2]; gg[5] += v[2] * v[2];
    }
    for (int i = 0; i < 6; ++i) gg[i] = block_sum_d(gg[i], shd);
    if (tid == 0) {
        // the reference forms GG in fp32 (torch mm) and hands the fp32 values to SciPy
        const float a[3][3] = {{(float)gg[0], (float)gg[1], (float)gg[2]}, {(float)gg[1], (float)gg[3], (float)gg[4]},
                               {(float)gg[2], (float)gg[4], (float)gg[5]}};
        Quad3 q;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) q.A[i][j] = (double)a[i][j];
        double w[3] = {0, 0, 0}; double c = 0; int iters = 0;
        cagrad_weights(&a[0][0], n, U.alpha, U.solver, w, &c, &iters);
        q.c = c;
        for (int i = 0; i < n; ++i) { q.Ab[i] = 0; for (int j = 0; j < n; ++j) q.Ab[i] += q.A[i][j] / n; }
        // gw = G w with w cast to fp32 as torch.Tensor(w_cpu) does (:719)
        for (int i = 0; i < n; ++i) w[i] = (double)(float)w[i];
        double gw2 = 0;
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) gw2 += w[i] * q.A[i][j] * w[j];
        const double lam = c / (sqrt(gw2 > 0 ? gw2 : 0) + 1e-8);
        // g = n * (mean_i G_i + lam * sum_i w_i G_i) / (1 + alpha^2)  = sum_i k_i G_i
        double k[3], norm2 = 0;
        for (int i = 0; i < n; ++i) k[i] = (double)n * (1.0 / n + lam * w[i]) / (1.0 + (double)U.alpha * U.alpha);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) norm2 += k[i] * q.A[i][j] * k[j];
        const double norm = sqrt(norm2 > 0 ? norm2 : 0);
        double clip = 1.0;
        if (U.max_norm > 0) { clip = (double)U.max_norm / (norm + 1e-6); if (clip > 1.0) clip = 1.0; }
        for (int i = 0; i < MAXT; ++i) coef[i] = i < n ? k[i] * clip : 0.0;
        if (U.diag) {
            for (int i = 0; i < 3; ++i) U.diag[i] = 0.f;
            for (int i = 0; i < n; ++i) U.diag[tl[i]] = (float)w[i];
            for (int i = 0; i < 9; ++i) U.diag[3 + i] = 0.f;
            for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) U.diag[3 + tl[i] * 3 + tl[j]] = a[i][j];
            U.diag[12] = (float)norm; U.diag[13] = (float)cg_obj(q, w, n); U.diag[14] = (float)iters; U.diag[15] = (float)clip;
        }
    }
    __syncthreads();
    const float k0 = (float)coef[0], k1 = (float)coef[1], k2 = (float)coef[2];
    const int t0 = tl[0], t1 = n > 1 ? tl[1] : tl[0], t2 = n > 2 ? tl[2] : tl[0];
    for (int ip = 0; ip < U.nparams; ++ip) {
        const ParamSeg s = U.ps[ip];
        for (int e = tid; e < s.numel; e += nth) {
            float g;
            if (s.shared_off >= 0) {
                const size_t p = (size_t)s.shared_off + e;
                g = k0 * G[(size_t)t0 * U.P + p];
                if (n > 1) g = fmaf(k1, G[(size_t)t1 * U.P + p], g);
                if (n > 2) g = fmaf(k2, G[(size_t)t2 * U.P + p], g);
            } else {
                g = PG[s.off + e];
            }
            if (U.grads_out) U.grads_out[s.off + e] = s.has_grad ? g : 0.f;
            if (U.do_sgd && s.has_grad) {
                const float p = U.params[s.off + e];
                g = fmaf(U.wd, p, g);
                const float b = fmaf(U.mom, U.momentum[s.off + e], g);
                U.momentum[s.off + e] = b;
                U.params[s.off + e] = p - U.lr * b;
            }
        }
    }
}

}  // namespace gaitk
