// cagrad_solver.cuh -- the simplex problem inside CAGrad (multitask_weighting.py:702-718):
//
//     minimise  f(w) = w^T A b + c * sqrt(w^T A w + 1e-8),   b = 1/n,   sum w = 1,   0 <= w <= 1
//
// The reference calls scipy.optimize.minimize(objfn, x_start, bounds=, constraints=) which selects SLSQP
// (Kraft 1988, DFVLR-FB 88-28) with ftol = 1e-6 and maxiter = 100.  SLSQP stops early by design, so its
// answer is NOT the exact optimum (and is the uniform start whenever |g^T s| < ftol, e.g. for small
// gradients) -- parity with the reference therefore needs SLSQP's own iteration, not a better solver.
// `slsqp_simplex` restates that iteration for this problem class (n <= 3, one linear equality, box
// bounds), from the published algorithm:
//   * QP sub-problem  min 1/2 s^T B s + g^T s,  sum s = 1 - sum x,  -x <= s <= 1 - x  (SciPy solves it
//     with Lawson-Hanson LSEI/LDP/NNLS; it is strictly convex, so the unique solution is found here by
//     enumerating the <= 27 bound-activity patterns and checking the KKT conditions);
//   * convergence test |g^T s| < acc before the line search;
//   * Armijo test phi(alpha) - phi(0) <= alpha*phi'(0)/10 with quadratic-interpolation back-off
//     alpha <- max(h3 / (2 (h3 - h1)), 0.1), at most 10 reductions (the constraint is linear and every
//     iterate feasible, so the L1 merit function reduces to f);
//   * convergence test |f - f0| < acc or ||s|| < acc after the step;
//   * BFGS update with Powell damping (theta from s^T u < 0.2 s^T B s); reset to identity on a
//     non-descent direction, at most 5 times, then the relaxed (10*acc) test.
// Validated against SciPy 1.18.1 on random / collinear / conflicting / tiny / dominant Gram matrices with
// entries up to 1e6: weights agree to <= 1.2e-4 (p99 1.5e-7); tests/test_host_cpu.py repeats this.  Above
// ~2e6 SciPy's own LSQ sub-solver breaks down erratically (returns the start point or "mode 4"); that
// regime (shared-gradient norms > 1400) is outside the parity range.
//
// `exact_simplex` is the true optimum (closed-form line minima + bisection), kept as solver mode 1.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define GAITK_HD __host__ __device__
#else
#define GAITK_HD
#endif

namespace gaitk {

struct Quad3 { double A[3][3]; double Ab[3]; double c; };

GAITK_HD inline double cg_obj(const Quad3& q, const double* w, int n) {
    double lin = 0, quad = 0;
    for (int i = 0; i < n; ++i) { lin += w[i] * q.Ab[i]; for (int j = 0; j < n; ++j) quad += w[i] * q.A[i][j] * w[j]; }
    return lin + q.c * sqrt(quad + 1e-8);
}
GAITK_HD inline void cg_grad(const Quad3& q, const double* w, int n, double* g) {
    double Aw[3] = {0, 0, 0}, quad = 0;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) Aw[i] += q.A[i][j] * w[j];
    for (int i = 0; i < n; ++i) quad += w[i] * Aw[i];
    const double r = q.c / sqrt(quad + 1e-8);
    for (int i = 0; i < n; ++i) g[i] = q.Ab[i] + r * Aw[i];
}
GAITK_HD inline void cg_clip01(const double* x, double* y, int n) {
    for (int i = 0; i < n; ++i) y[i] = x[i] < 0 ? 0.0 : (x[i] > 1 ? 1.0 : x[i]);
}

// ---- everything below is templated on the number of tasks N (2 or 3) and written with STATIC array indices only: on the
// device one warp runs the solve while the whole update CTA waits for it, and the first version (runtime n, index lists,
// a run-time-sized elimination) kept its small arrays in local memory: ~31 K clocks per SLSQP iteration, 48 - 140 us per
// step (scratch/update_timing.py).  With static indices everything lives in registers.
template <int N> GAITK_HD inline double cg_obj_t(const Quad3& q, const double* w) {
    double lin = 0, quad = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        lin += w[i] * q.Ab[i];
#pragma unroll
        for (int j = 0; j < N; ++j) quad += w[i] * q.A[i][j] * w[j];
    }
    return lin + q.c * sqrt(quad + 1e-8);
}
template <int N> GAITK_HD inline void cg_grad_t(const Quad3& q, const double* w, double* g) {
    double Aw[N], quad = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        Aw[i] = 0;
#pragma unroll
        for (int j = 0; j < N; ++j) Aw[i] += q.A[i][j] * w[j];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) quad += w[i] * Aw[i];
    const double r = q.c / sqrt(quad + 1e-8);
#pragma unroll
    for (int i = 0; i < N; ++i) g[i] = q.Ab[i] + r * Aw[i];
}

// Gaussian elimination with partial pivoting on an M x M system (M <= 4), rows swapped by selects.  false when singular.
template <int M> GAITK_HD inline bool cg_solve_static(double (&K)[M][M], double (&r)[M]) {
    double inv[M];
#pragma unroll
    for (int col = 0; col < M; ++col) {
        int piv = col; double best = fabs(K[col][col]);
#pragma unroll
        for (int i = col + 1; i < M; ++i) { const double a = fabs(K[i][col]); if (a > best) { best = a; piv = i; } }
        if (!(best > 1e-300)) return false;
#pragma unroll
        for (int i = col + 1; i < M; ++i) {
            const bool sw = piv == i;
#pragma unroll
            for (int j = 0; j < M; ++j) { const double a = K[col][j], b = K[i][j]; K[col][j] = sw ? b : a; K[i][j] = sw ? a : b; }
            const double a = r[col], b = r[i]; r[col] = sw ? b : a; r[i] = sw ? a : b;
        }
        inv[col] = 1.0 / K[col][col];
#pragma unroll
        for (int i = col + 1; i < M; ++i) {
            const double f = K[i][col] * inv[col];
#pragma unroll
            for (int j = col; j < M; ++j) K[i][j] -= f * K[col][j];
            r[i] -= f * r[col];
        }
    }
#pragma unroll
    for (int i = M - 1; i >= 0; --i) {
        double acc = r[i];
#pragma unroll
        for (int j = i + 1; j < M; ++j) acc -= K[i][j] * r[j];
        r[i] = acc * inv[i];
    }
    return true;
}

// min 1/2 s^T B s + g^T s   s.t.  sum s = c0,  lo <= s <= hi     (B positive definite)
// One bound-activity pattern `comb` (base-3 digits: 0 free, 1 at lower, 2 at upper): solve the equality-constrained
// sub-problem and check the KKT conditions.  Returns true with (s, val) when it is THE solution.  The sub-problem is ONE
// (N + 1) x (N + 1) system for every pattern: a free variable contributes its stationarity row  B_i. s + lambda = -g_i,
// a fixed variable the identity row  s_i = bound,  the last row the constraint over the free variables.
template <int N>
GAITK_HD inline bool cg_qp_combo_t(const double (&B)[3][3], const double* g, const double* lo, const double* hi, double c0,
                                   int comb, double* s, double* val_out) {
    int st[N];
    { int t = comb;
#pragma unroll
      for (int i = N - 1; i >= 0; --i) { st[i] = t % 3; t /= 3; } }
    int nf = 0; double fixed_sum = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const bool fr = st[i] == 0;
        nf += fr ? 1 : 0;
        s[i] = fr ? 0.0 : (st[i] == 1 ? lo[i] : hi[i]);
        fixed_sum += s[i];
    }
    const double rhs_sum = c0 - fixed_sum;
    double lam = 0; const bool has_lam = nf > 0;
    if (!has_lam) {
        if (fabs(rhs_sum) > 1e-12) return false;
    } else {
        double K[N + 1][N + 1], r[N + 1];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const bool fr = st[i] == 0;
#pragma unroll
            for (int j = 0; j < N; ++j) K[i][j] = fr ? B[i][j] : (i == j ? 1.0 : 0.0);
            K[i][N] = fr ? 1.0 : 0.0;
            r[i] = fr ? -g[i] : s[i];
            K[N][i] = fr ? 1.0 : 0.0;
        }
        K[N][N] = 0.0; r[N] = rhs_sum;
        if (!cg_solve_static<N + 1>(K, r)) return false;
#pragma unroll
        for (int i = 0; i < N; ++i) s[i] = st[i] == 0 ? r[i] : s[i];
        lam = r[N];
    }
#pragma unroll
    for (int i = 0; i < N; ++i) if (st[i] == 0 && (s[i] < lo[i] - 1e-13 || s[i] > hi[i] + 1e-13)) return false;
    double grad[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        grad[i] = g[i];
#pragma unroll
        for (int j = 0; j < N; ++j) grad[i] += B[i][j] * s[j];
    }
    if (!has_lam) {
        double lo_l = -INFINITY, hi_l = INFINITY;
#pragma unroll
        for (int i = 0; i < N; ++i) { if (st[i] == 1) lo_l = fmax(lo_l, -grad[i]); else hi_l = fmin(hi_l, -grad[i]); }
        if (lo_l > hi_l + 1e-12) return false;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const double m = grad[i] + lam, t = 1e-12 * fmax(1.0, fabs(grad[i]));
            if (st[i] == 1 && m < -t) return false;
            if (st[i] == 2 && m > t) return false;
        }
    }
    double val = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        val += g[i] * s[i];
#pragma unroll
        for (int j = 0; j < N; ++j) val += 0.5 * s[i] * B[i][j] * s[j];
    }
    *val_out = val;
    return true;
}

template <int N>
GAITK_HD inline void cg_qp_t(const double (&B)[3][3], const double* g, const double* lo, const double* hi, double c0, double* s_out) {
    constexpr int ncomb = N == 2 ? 9 : 27;
#if defined(__CUDA_ARCH__)
    // device: the <= 27 patterns are evaluated by the lanes of the (fully active) calling warp
    const int lane = threadIdx.x & 31;
    double s[N], val = INFINITY;
#pragma unroll
    for (int i = 0; i < N; ++i) s[i] = 0.0;
    const bool ok = lane < ncomb && cg_qp_combo_t<N>(B, g, lo, hi, c0, lane, s, &val);
    if (!ok) val = INFINITY;
    double best = val; int who = lane;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, o); const int ow = __shfl_xor_sync(0xffffffffu, who, o);
        if (ov < best || (ov == best && ow < who)) { best = ov; who = ow; }
    }
#pragma unroll
    for (int i = 0; i < N; ++i) { const double v = __shfl_sync(0xffffffffu, s[i], who); s_out[i] = best < INFINITY ? v : 0.0; }
#else
    double best_val = 0; bool have = false;
    for (int i = 0; i < N; ++i) s_out[i] = 0.0;
    for (int comb = 0; comb < ncomb; ++comb) {
        double s[N], val;
        if (!cg_qp_combo_t<N>(B, g, lo, hi, c0, comb, s, &val)) continue;
        if (!have || val < best_val - 1e-15) { have = true; best_val = val; for (int i = 0; i < N; ++i) s_out[i] = s[i]; }
    }
#endif
}

// SLSQP iteration from x = 1/n.  Returns the exit mode (0 converged, 8 positive directional derivative,
// 9 iteration limit); *iters = major iterations.
template <int N>
GAITK_HD inline int slsqp_simplex_t(const Quad3& q, double* x, int* iters, double acc = 1e-6, int itermax = 100) {
#pragma unroll
    for (int i = 0; i < N; ++i) x[i] = 1.0 / N;
    double B[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    double xc[N], g[N], s[N], x0[N], gn[N], u[N], v[N], lo[N], hi[N];
    cg_clip01(x, xc, N);
    double fx = cg_obj_t<N>(q, xc); cg_grad_t<N>(q, xc, g);
    const double tol = 10.0 * acc;
    int ireset = 1, it = 0;
    for (;;) {
        ++it;
        if (it > itermax) { *iters = it; return 9; }
        double sx = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) { lo[i] = 0.0 - x[i]; hi[i] = 1.0 - x[i]; sx += x[i]; }
        cg_qp_t<N>(B, g, lo, hi, 1.0 - sx, s);
        const double f0 = fx;
        double gs = 0, snorm = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) { x0[i] = x[i]; gs += g[i] * s[i]; snorm += s[i] * s[i]; }
        snorm = sqrt(snorm);
        if (fabs(gs) < acc && fabs(1.0 - sx) < acc) { *iters = it; return 0; }
        double h3 = gs;
        if (h3 >= 0.0) {
            ++ireset;
            if (ireset > 5) { *iters = it; return (fabs(fx - f0) < tol || snorm < tol) ? 0 : 8; }
#pragma unroll
            for (int i = 0; i < 3; ++i)
#pragma unroll
                for (int j = 0; j < 3; ++j) B[i][j] = i == j ? 1.0 : 0.0;
            continue;
        }
        int line = 0; double alpha = 1.0;
        for (;;) {
            ++line;
            h3 *= alpha;
#pragma unroll
            for (int i = 0; i < N; ++i) { s[i] *= alpha; x[i] = x0[i] + s[i]; }
            cg_clip01(x, xc, N);
            fx = cg_obj_t<N>(q, xc);
            const double h1 = fx - f0;
            if (h1 <= h3 / 10.0 || line > 10) break;
            alpha = fmax(h3 / (2.0 * (h3 - h1)), 0.1);
        }
        snorm = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) snorm += s[i] * s[i];
        snorm = sqrt(snorm);
        if (fabs(fx - f0) < acc || snorm < acc) { *iters = it; return 0; }
        cg_grad_t<N>(q, xc, gn);
        double h1 = 0, h2 = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) {
            u[i] = gn[i] - g[i]; v[i] = 0;
#pragma unroll
            for (int j = 0; j < N; ++j) v[i] += B[i][j] * s[j];
        }
#pragma unroll
        for (int i = 0; i < N; ++i) { h1 += s[i] * u[i]; h2 += s[i] * v[i]; }
        const double h3b = 0.2 * h2;
        if (h1 < h3b) {
            const double h4 = (h2 - h3b) / (h2 - h1);
            h1 = h3b;
#pragma unroll
            for (int i = 0; i < N; ++i) u[i] = h4 * u[i] + (1.0 - h4) * v[i];
        }
        // two reciprocals instead of 2 n^2 divisions (an fp64 division is a long dependent chain for the one warp that solves)
        const double r1 = 1.0 / h1, r2 = 1.0 / h2;
#pragma unroll
        for (int i = 0; i < N; ++i)
#pragma unroll
            for (int j = 0; j < N; ++j) B[i][j] += u[i] * u[j] * r1 - v[i] * v[j] * r2;
#pragma unroll
        for (int i = 0; i < N; ++i) g[i] = gn[i];
    }
}
GAITK_HD inline int slsqp_simplex(const Quad3& q, int n, double* x, int* iters) {
    if (n == 1) { x[0] = 1.0; *iters = 0; return 0; }
    return n == 2 ? slsqp_simplex_t<2>(q, x, iters) : slsqp_simplex_t<3>(q, x, iters);
}

// ---- exact optimum (solver mode 1) ---------------------------------------------------------------------
// argmin over tau in [0,1] of f(p + tau * d): closed form
GAITK_HD inline double cg_line_min(const Quad3& q, const double* p, const double* d, int n) {
    double a1 = 0, q0 = 1e-8, q1 = 0, q2 = 0;
    for (int i = 0; i < n; ++i) {
        a1 += d[i] * q.Ab[i];
        for (int j = 0; j < n; ++j) { q0 += p[i] * q.A[i][j] * p[j]; q1 += p[i] * q.A[i][j] * d[j]; q2 += d[i] * q.A[i][j] * d[j]; }
    }
    if (q2 < 0) q2 = 0;
    const double c = q.c;
    const double lim = c * sqrt(q2);           // h'(+-inf) = a1 +- c sqrt(q2)
    if (!(fabs(a1) < lim)) return a1 > 0 ? 0.0 : (a1 < 0 ? 1.0 : 0.0);
    const double rho = -a1 / c;
    double disc = q0 * q2 - q1 * q1; if (disc < 0) disc = 0;
    const double m = (rho >= 0 ? 1.0 : -1.0) * fabs(rho) * sqrt(disc / (q2 - rho * rho));
    const double tau = (m - q1) / q2;
    return tau < 0 ? 0.0 : (tau > 1 ? 1.0 : tau);
}
GAITK_HD inline double cg_phi_prime(const Quad3& q, double uu, double* wo) {
    const double p[3] = {uu, 0, 1 - uu}, d[3] = {0, 1 - uu, -(1 - uu)};
    const double t = (1 - uu > 0) ? cg_line_min(q, p, d, 3) : 0.0;
    wo[0] = uu; wo[1] = (1 - uu) * t; wo[2] = (1 - uu) * (1 - t);
    double g[3]; cg_grad(q, wo, 3, g);
    return g[0] - t * g[1] - (1 - t) * g[2];
}
GAITK_HD inline void exact_simplex(const Quad3& q, int n, double* w, int* iters) {
    *iters = 0;
    if (n == 1) { w[0] = 1; return; }
    if (n == 2) {
        const double p[3] = {0, 1, 0}, d[3] = {1, -1, 0};
        const double t = cg_line_min(q, p, d, 2);
        w[0] = t; w[1] = 1 - t; return;
    }
    // w = (u, (1-u) tau, (1-u)(1-tau)); phi(u) = min_tau f is convex, phi'(u) = g0 - tau g1 - (1-tau) g2
    double wl[3];
    if (cg_phi_prime(q, 0.0, wl) >= 0) { w[0] = wl[0]; w[1] = wl[1]; w[2] = wl[2]; return; }
    if (cg_phi_prime(q, 1.0 - 1e-12, wl) <= 0) { w[0] = 1; w[1] = 0; w[2] = 0; return; }
    double lo = 0, hi = 1;
    for (int it = 0; it < 64; ++it) {
        const double mid = 0.5 * (lo + hi);
        if (cg_phi_prime(q, mid, wl) > 0) hi = mid; else lo = mid;
        ++*iters;
        if (hi - lo < 1e-15) break;
    }
    cg_phi_prime(q, 0.5 * (lo + hi), w);
}

// A (fp32 Gram values, n x n row-major with leading dimension 3) + alpha -> weights.  Follows
// multitask_weighting.py:699-718: g0 = sqrt(mean(GG) + 1e-8) in fp32, c = alpha * g0 + 1e-8.
GAITK_HD inline int cagrad_weights(const float* a, int n, float alpha, int solver, double* w, double* c_out, int* iters) {
    Quad3 q;
    double mean = 0;
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) { q.A[i][j] = (double)a[i * 3 + j]; mean += (double)a[i * 3 + j]; }
    mean /= (double)(n * n);
    const float g0 = sqrtf((float)mean + 1e-8f);
    // the reference evaluates (alpha * g0_norm + 1e-8) as a float32 tensor expression, then .item()
    q.c = (double)(alpha * g0 + 1e-8f);
    for (int i = 0; i < n; ++i) { double rs = 0; for (int j = 0; j < n; ++j) rs += q.A[i][j]; q.Ab[i] = rs / n; }      // A b, b = 1/n
    *c_out = q.c;
    int mode = 0;
    if (solver == 1) exact_simplex(q, n, w, iters); else mode = slsqp_simplex(q, n, w, iters);
    for (int i = 0; i < n; ++i) w[i] = w[i] < 0 ? 0.0 : (w[i] > 1 ? 1.0 : w[i]);
    return mode;
}

}  // namespace gaitk
