// dd_kalman.cuh -- 8-state constant-velocity Kalman filter in f64 (deep_sort/kalman_filter.py).
// One group (warp) per track; cross-lane values go through a 96-double scratch (shared memory on
// the device).  All arithmetic is unfused f64 so thresholded decisions downstream (gating, IoU,
// count-line) see the same roundings as numpy's element-wise ops wherever the reference's own
// operation order is deterministic (predict, project, initiate); the LAPACK/BLAS-backed steps
// (Cholesky, gain, covariance update) agree to ~1e-16 relative.
#pragma once
#include "dd_common.cuh"

#define DD_STD_POS (1.0 / 20)     // kalman_filter.py:52
#define DD_STD_VEL (1.0 / 160)    // kalman_filter.py:53
#define DD_CHI2INV95_4 9.4877     // kalman_filter.py:11-20
#define DD_CHI2INV95_2 5.9915

// kalman_filter.py:55-86.  z = (x, y, a, h).  Lanes cooperate over the 64 covariance entries.
template <class G>
DD_HD void dd_kf_initiate(const G& g, const double* z, double* mean, double* cov) {
    const double h = z[3];
    const double sp = dd_mul(2 * DD_STD_POS, h), sv = dd_mul(10 * DD_STD_VEL, h);
    for (int e = g.lane; e < 64; e += G::NL) {
        const int i = e >> 3, j = e & 7;
        double v = 0.0;
        if (i == j) {
            double s = (i == 2) ? 1e-2 : (i == 6) ? 1e-5 : (i < 4 ? sp : sv);
            v = dd_mul(s, s);
        }
        cov[e] = v;
    }
    for (int i = g.lane; i < 8; i += G::NL) mean[i] = i < 4 ? z[i] : 0.0;
}

// kalman_filter.py:88-123.  P' = F (P F^T) + Q with F = [[I, I], [0, I]] -- with 0/1 coefficients
// every entry is a sum of at most four terms in a fixed order, so this is bit-identical to numpy.
template <class G>
DD_HD void dd_kf_predict(const G& g, double* mean, double* cov) {
    constexpr int PER = 64 / G::NL;
    double outv[PER];
    const double h = mean[3];
    const double sp = dd_mul(DD_STD_POS, h), sv = dd_mul(DD_STD_VEL, h);
    int n = 0;
    for (int e = g.lane; e < 64; e += G::NL, ++n) {
        const int i = e >> 3, j = e & 7;
        double x = cov[i * 8 + j];                                // (P F^T)[i][j]
        if (j < 4) x = dd_add(x, cov[i * 8 + j + 4]);
        if (i < 4) {
            double y = cov[(i + 4) * 8 + j];                      // (P F^T)[i+4][j]
            if (j < 4) y = dd_add(y, cov[(i + 4) * 8 + j + 4]);
            x = dd_add(x, y);
        }
        if (i == j) {
            double s = (i == 2) ? 1e-2 : (i == 6) ? 1e-5 : (i < 4 ? sp : sv);
            x = dd_add(x, dd_mul(s, s));
        }
        outv[n] = x;
    }
    double m_new = 0.0;
    const int mi = g.lane;                                        // HostG: handled by the loop below
    if (G::NL > 1 && mi < 4) m_new = dd_add(mean[mi], mean[mi + 4]);
    g.sync();
    n = 0;
    for (int e = g.lane; e < 64; e += G::NL, ++n) cov[e] = outv[n];
    if (G::NL > 1) {
        if (mi < 4) mean[mi] = m_new;
    } else {
        for (int i = 0; i < 4; ++i) mean[i] = dd_add(mean[i], mean[i + 4]);
    }
    g.sync();
}

// kalman_filter.py:125-152 -- projected mean = mean[:4]; S = P[:4,:4] + diag(std^2) (exact).
DD_HD void dd_kf_project_cov(const double* mean, const double* cov, double* S) {
    const double sp = dd_mul(DD_STD_POS, mean[3]);
    const double r0 = dd_mul(sp, sp), r2 = dd_mul(1e-1, 1e-1);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            double v = cov[i * 8 + j];
            if (i == j) v = dd_add(v, i == 2 ? r2 : r0);
            S[i * 4 + j] = v;
        }
}

// Lower Cholesky factor of the leading N x N block of a 4x4 SPD matrix (row-major, stride 4), fully
// unrolled so L lives in registers; rinv[j] = 1 / L[j][j] is returned so the triangular solves multiply
// by a reciprocal instead of dividing (LAPACK/OpenBLAS trsm kernels do the same; results differ from a
// true division by <= 1 ulp, far inside the 1e-4 Kalman tolerance and the same for every code path here).
template <int N>
DD_HD void dd_chol(const double* S, double* L, double* rinv) {
#pragma unroll
    for (int j = 0; j < N; ++j) {
        double d = S[j * 4 + j];
#pragma unroll
        for (int k = 0; k < j; ++k) d = dd_sub(d, dd_mul(L[j * 4 + k], L[j * 4 + k]));
        const double ljj = dd_sqrt(d);
        L[j * 4 + j] = ljj;
        const double r = dd_div(1.0, ljj);
        rinv[j] = r;
#pragma unroll
        for (int i = j + 1; i < N; ++i) {
            double v = S[i * 4 + j];
#pragma unroll
            for (int k = 0; k < j; ++k) v = dd_sub(v, dd_mul(L[i * 4 + k], L[j * 4 + k]));
            L[i * 4 + j] = dd_mul(v, r);
        }
    }
}

// kalman_filter.py:223-228 -- squared Mahalanobis distance of one measurement: solve L z = d.
template <int N>
DD_HD double dd_maha_sq(const double* L, const double* rinv, const double* pmean, const double* meas) {
    double z[N];
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < N; ++i) {
        double v = dd_sub(meas[i], pmean[i]);
#pragma unroll
        for (int k = 0; k < i; ++k) v = dd_sub(v, dd_mul(L[i * 4 + k], z[k]));
        z[i] = dd_mul(v, rinv[i]);
        const double sq = dd_mul(z[i], z[i]);
        acc = (i == 0) ? sq : dd_add(acc, sq);
    }
    return acc;
}

// kalman_filter.py:154-186.  scratch: >= 64 doubles visible to the whole group.
template <class G>
DD_HD void dd_kf_update(const G& g, double* mean, double* cov, const double* z, double* scratch) {
    double* Ksh = scratch;        // gain K [8][4]
    double* Msh = scratch + 32;   // S K^T  [4][8]
    double S[16], L[16], rinv[4];
    dd_kf_project_cov(mean, cov, S);
    dd_chol<4>(S, L, rinv);
    double innov[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) innov[i] = dd_sub(z[i], mean[i]);
    // K[c][:] = S^-1 (P H^T)^T[:, c]  -> cho_solve with rhs column c = P[c][0..3]
    for (int c = g.lane; c < 8; c += G::NL) {
        double y[4], x[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            double v = cov[c * 8 + i];
#pragma unroll
            for (int k = 0; k < i; ++k) v = dd_sub(v, dd_mul(L[i * 4 + k], y[k]));
            y[i] = dd_mul(v, rinv[i]);
        }
#pragma unroll
        for (int i = 3; i >= 0; --i) {
            double v = y[i];
#pragma unroll
            for (int k = i + 1; k < 4; ++k) v = dd_sub(v, dd_mul(L[k * 4 + i], x[k]));
            x[i] = dd_mul(v, rinv[i]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) Ksh[c * 4 + i] = x[i];
    }
    g.sync();
    for (int e = g.lane; e < 32; e += G::NL) {       // M = S K^T
        const int i = e >> 3, j = e & 7;
        double v = dd_mul(S[i * 4 + 0], Ksh[j * 4 + 0]);
#pragma unroll
        for (int k = 1; k < 4; ++k) v = dd_add(v, dd_mul(S[i * 4 + k], Ksh[j * 4 + k]));
        Msh[e] = v;
    }
    g.sync();
    constexpr int PER = 64 / G::NL;
    double outv[PER];
    int n = 0;
    for (int e = g.lane; e < 64; e += G::NL, ++n) {  // P - K M
        const int a = e >> 3, b = e & 7;
        double v = dd_mul(Ksh[a * 4 + 0], Msh[0 * 8 + b]);
#pragma unroll
        for (int i = 1; i < 4; ++i) v = dd_add(v, dd_mul(Ksh[a * 4 + i], Msh[i * 8 + b]));
        outv[n] = dd_sub(cov[e], v);
    }
    double mnew[8 / (G::NL > 8 ? 8 : G::NL)];
    n = 0;
    for (int j = g.lane; j < 8; j += G::NL, ++n) {   // mean + innov . K^T
        double v = dd_mul(innov[0], Ksh[j * 4 + 0]);
#pragma unroll
        for (int i = 1; i < 4; ++i) v = dd_add(v, dd_mul(innov[i], Ksh[j * 4 + i]));
        mnew[n] = dd_add(mean[j], v);
    }
    g.sync();
    n = 0;
    for (int e = g.lane; e < 64; e += G::NL, ++n) cov[e] = outv[n];
    n = 0;
    for (int j = g.lane; j < 8; j += G::NL, ++n) mean[j] = mnew[n];
    g.sync();
}
