// nei.cuh -- cNEI / cNEHVI sample values: joint posterior of the refit model over [X_baseline; X].
//
// qConformalNoisyExpectedImprovement._nonconformal_fwd (conformalbo/acquisitions.py:115-134) samples f jointly at the
// baseline points and the candidates under every refit model and scores  softplus(f_x - smoothmax(f_b)).  The covariance of
// that joint posterior does not depend on the grid point, so ONE (nb+q)^2 Cholesky per joint batch serves all G grid points and
// MC samples (the reference factorises per (candidate, grid point)).  nb + q <= 64: botorch prunes the baseline to the few
// points that can still be the best (hydra_config/acq_fn/cnei.yaml `prune_baseline: true`; sampling.prune_inferior_points).
#pragma once
#include "common.cuh"
#include "rows.cuh"
#include "score_q.cuh"

namespace cbo {

constexpr int NEI_MAX_M = 64;                  // rows of the joint factor: nb + q <= 64 (each lane owns rows lane and lane + 32)
constexpr int NEI_MAX_NB = NEI_MAX_M - 1;

struct NeiArgs {
    int N, Np, d, nb, G, S, n_cand, tiles_n;
    double noise, kappa, temp;
    const double* Xq;        // rows x d
    const double* inv_ls;    // d
    const double* Xb;        // nb x d
    const double* KbX;       // nb x Np   k(X_b, X)
    const double* Sbb;       // nb x nb   posterior covariance of f(X_b) given the data (noise-free)
    const double* mub;       // nb        posterior mean of f(X_b)
    const double* W;         // rows x Np
    const double* rn_partial;
    const double* mu;        // rows
    const double* cscale;    // rows or nullptr
    const double* eps_grid;  // G x q or n_cand x G x q
    int eps_per_candidate;
    const double* eps_mc;    // S x G x (nb+q)
    double* values;          // n_cand x S x G x q
    int raw;                 // 1: write the joint-posterior samples f(x) themselves (cNEHVI), 0: the noisy-EI improvement
    const double* V;         // rows x Np (v = L^-1 u); rows = n_cand * q
};

__host__ __device__ inline int nei_ld(int m) { return m | 1; }                         // odd row stride: conflict-free rows and columns
__host__ __device__ inline size_t nei_warp_doubles(int nb, int q) { return size_t(nb + q) * nei_ld(nb + q) + 2 * size_t(nb) * q; }

// Per joint batch: Schur complement S, its inverse, the factor of the (row-scaled) grid covariance, c = Cov_D(f(X_b), y(X)),
// c S^-1, and the Cholesky factor L (row stride ld) of the joint refit covariance.  One warp; every lane ends up with the small
// q x q matrices, the shared arrays hold L, c (cv) and c S^-1 (cs).  Returns false when a pivot is not positive.
template <int Q>
__device__ __forceinline__ bool nei_setup(const NeiArgs& p, int c, int lane, double* L, double* cv, double* cs,
                                          double (&Sm)[Q][Q], double (&Si)[Q][Q], double (&Lc)[Q][Q]) {
    const int nb = p.nb, m = nb + Q, r0 = c * Q, ld = nei_ld(m);
    double Ls[Q][Q], Cov[Q][Q];
#pragma unroll
    for (int j = 0; j < Q; ++j)
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            if (k > j) continue;
            double acc = 0.0;
            for (int i = lane; i < p.N; i += 32) acc = fma(p.V[int64_t(r0 + j) * p.Np + i], p.V[int64_t(r0 + k) * p.Np + i], acc);
            acc = warp_sum(acc);
            double r2 = 0.0;
            for (int e = 0; e < p.d; ++e) {
                const double df = (p.Xq[int64_t(r0 + j) * p.d + e] - p.Xq[int64_t(r0 + k) * p.d + e]) * p.inv_ls[e];
                r2 = fma(df, df, r2);
            }
            const double r = sqrt(fmax(r2, 1e-30));
            const double kk = p.kappa * (1.0 + kSqrt5 * r + (5.0 / 3.0) * r * r) * exp(-kSqrt5 * r);
            const double v = kk - acc + ((j == k) ? p.noise : 0.0);
            Sm[j][k] = v; Sm[k][j] = v;
        }
    bool ok = chol_lower<Q>(Sm, Ls);
    inv_from_chol<Q>(Ls, Si);
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        const double rs = p.cscale ? p.cscale[r0 + j] : 1.0;
#pragma unroll
        for (int k = 0; k < Q; ++k) Cov[j][k] = rs * Sm[j][k];
    }
    ok = chol_lower<Q>(Cov, Lc) && ok;
    // c[r][j] = k(X_b,r, x_j) - K_bX[r,:] . w_j
    for (int r = 0; r < nb; ++r) {
        double acc[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) acc[j] = 0.0;
        for (int i = lane; i < p.N; i += 32) {
            const double kb = p.KbX[int64_t(r) * p.Np + i];
#pragma unroll
            for (int j = 0; j < Q; ++j) acc[j] = fma(kb, p.W[int64_t(r0 + j) * p.Np + i], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            acc[j] = warp_sum(acc[j]);
            if (lane == j) {
                double r2 = 0.0;
                for (int e = 0; e < p.d; ++e) {
                    const double df = (p.Xb[r * p.d + e] - p.Xq[int64_t(r0 + j) * p.d + e]) * p.inv_ls[e];
                    r2 = fma(df, df, r2);
                }
                const double rr = sqrt(fmax(r2, 1e-30));
                cv[r * Q + j] = p.kappa * (1.0 + kSqrt5 * rr + (5.0 / 3.0) * rr * rr) * exp(-kSqrt5 * rr) - acc[j];
            }
        }
    }
    __syncwarp();
    for (int r = lane; r < nb; r += 32) {
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k) a = fma(cv[r * Q + k], 0.5 * (Si[k][j] + Si[j][k]), a);
            cs[r * Q + j] = a;
        }
    }
    __syncwarp();
    // covariance (lower triangle), a lane owns rows lane, lane + 32
    for (int r = lane; r < m; r += 32) {
        for (int k = 0; k <= r; ++k) {
            double v;
            if (r < nb) {
                v = p.Sbb[r * nb + k];
#pragma unroll
                for (int j = 0; j < Q; ++j) v -= cs[r * Q + j] * cv[k * Q + j];
            } else {
                const int j = r - nb;
                if (k < nb) v = p.noise * cs[k * Q + j];
                else {
                    const int j2 = k - nb;
                    double sij = 0.0;
#pragma unroll
                    for (int a = 0; a < Q; ++a)
#pragma unroll
                        for (int b2 = 0; b2 < Q; ++b2)
                            if (a == j && b2 == j2) sij = 0.5 * (Si[a][b2] + Si[b2][a]);
                    v = p.noise * (((j == j2) ? 1.0 : 0.0) - p.noise * sij);
                }
            }
            L[r * ld + k] = v;
        }
        for (int k = r + 1; k < m; ++k) L[r * ld + k] = 0.0;
    }
    __syncwarp();
    // right-looking Cholesky
    for (int j = 0; j < m; ++j) {
        const double djj = L[j * ld + j];
        ok = ok && (djj > 0.0);
        const double ljj = sqrt(djj);
        __syncwarp();
        if (lane == 0) L[j * ld + j] = ljj;
        for (int r = lane; r < m; r += 32)
            if (r > j) L[r * ld + j] /= ljj;
        __syncwarp();
        for (int r = lane; r < m; r += 32)
            if (r > j) {
                const double lij = L[r * ld + j];
                for (int k = j + 1; k <= r; ++k) L[r * ld + k] -= lij * L[k * ld + j];
            }
        __syncwarp();
    }
    return ok;
}

// Joint batches (q = Q candidate points per batch, nb + Q <= 64): with c = Cov_D(f(X_b), y(X)) (nb x Q),
// S = K_qq + noise I - V^T V and B_g = S^-1 (y_g - mu) the refit joint posterior over [X_baseline; X] is
//   mean = [mu_b + c B_g ; mu + (y_g - mu) - noise B_g],
//   cov  = [[Sigma_bb - c S^-1 c^T, noise c S^-1], [., noise (I - noise S^-1)]]       -- independent of the grid point.
// One warp per joint batch for the factor; then one lane per (MC sample, grid point).  values: n_cand x S x G x Q, read by
// k_score (Q = 1) / k_score_q.
constexpr int NEI_WARPS = 4;

template <int Q>
__global__ void __launch_bounds__(NEI_WARPS * 32)
k_nei_values_q(const NeiArgs p) {
    extern __shared__ double nei_sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * NEI_WARPS + warp;
    if (c >= p.n_cand) return;
    const int nb = p.nb, m = nb + Q, r0 = c * Q, ld = nei_ld(m);
    double* L = nei_sm + size_t(warp) * nei_warp_doubles(nb, Q);
    double* cv = L + size_t(m) * ld;      // c[r][j]
    double* cs = cv + size_t(nb) * Q;     // (c S^-1)[r][j]
    double Sm[Q][Q], Si[Q][Q], Lc[Q][Q];
    const bool ok = nei_setup<Q>(p, c, lane, L, cv, cs, Sm, Si, Lc);
    double mu[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) mu[j] = p.mu[r0 + j];
    const double inv_temp = p.temp > 0.0 ? 1.0 / p.temp : 0.0;
    for (int sg = lane; sg < p.S * p.G; sg += 32) {
        const int smp = sg / p.G, g = sg - smp * p.G;
        const double* eg = p.eps_grid + (p.eps_per_candidate ? int64_t(c) * p.G + g : int64_t(g)) * Q;
        double dev[Q], Bg[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k <= j) a = fma(Lc[j][k], eg[k], a);
            dev[j] = a;
        }
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k) a = fma(Si[j][k], dev[k], a);
            Bg[j] = a;
        }
        const double* e = p.eps_mc + (int64_t(smp) * p.G + g) * m;
        // baseline rows: running (max, sum of exponentials) of f_b / temp; then the candidate rows
        double mx = -INFINITY, se = 0.0, fx[Q];
        for (int r = 0; r < m; ++r) {
            double acc;
            if (r < nb) {
                acc = p.mub[r];
#pragma unroll
                for (int j = 0; j < Q; ++j) acc = fma(cv[r * Q + j], Bg[j], acc);
            } else {
                acc = 0.0;
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    if (j == r - nb) acc = mu[j] + dev[j] - p.noise * Bg[j];
            }
            for (int k = 0; k <= r; ++k) acc = fma(L[r * ld + k], e[k], acc);
            if (r < nb) {
                if (p.temp > 0.0) {
                    if (acc > mx) { se = se * exp((mx - acc) * inv_temp) + 1.0; mx = acc; }
                    else se += exp((acc - mx) * inv_temp);
                } else mx = fmax(mx, acc);
            } else {
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    if (j == r - nb) fx[j] = acc;
            }
        }
        const double smax = (p.temp > 0.0) ? mx + p.temp * log(se) : mx;     // temp * logsumexp(f_b / temp)
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double val;
            if (p.temp > 0.0) {
                const double xt = (fx[j] - smax) * inv_temp;
                val = p.temp * (fmax(xt, 0.0) + log1p(exp(-fabs(xt))));
            } else {
                val = fmax(fx[j] - smax, 0.0);
            }
            if (p.raw) val = fx[j];
            p.values[((int64_t(c) * p.S + smp) * p.G + g) * Q + j] = ok ? val : nan("");
        }
    }
}

template <int Q>
inline cudaError_t launch_nei_values_t(const NeiArgs& a, cudaStream_t s) {
    const size_t smem = sizeof(double) * NEI_WARPS * nei_warp_doubles(a.nb, Q);
    static SmemOptIn optin;
    if (cudaError_t e = optin.ensure(k_nei_values_q<Q>, smem); e != cudaSuccess) return e;
    CBO_LAUNCHED(), k_nei_values_q<Q><<<(a.n_cand + NEI_WARPS - 1) / NEI_WARPS, NEI_WARPS * 32, smem, s>>>(a);
    return cudaGetLastError();
}

inline cudaError_t launch_nei_values_q(const NeiArgs& a, int q, cudaStream_t s) {
    switch (q) {
        case 1: return launch_nei_values_t<1>(a, s);
        case 2: return launch_nei_values_t<2>(a, s);
        case 3: return launch_nei_values_t<3>(a, s);
        case 4: return launch_nei_values_t<4>(a, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace cbo
