// nei.cuh -- cNEI sample values: joint posterior of the refit model over [X_baseline; x] (q = 1).
//
// qConformalNoisyExpectedImprovement._nonconformal_fwd (conformalbo/acquisitions.py:115-134) samples f jointly at the
// baseline points and the candidate under every refit model and scores  softplus(f_x - smoothmax(f_b)).  With
//   c_j = Cov_D(f(X_b,j), y(x)) = k(X_b,j, x) - K_bX[j,:] w,   P = Var_D(f(x)) = s - noise,   b_g = (y_g - mu)/s
// the refit joint posterior is  mean = [mu_b + c b_g ; mu + P b_g],
//   cov = [[Sigma_bb - c c^T / s,  c noise/s], [.,  P noise/s]]       -- independent of the grid point,
// so one (nb+1)^2 Cholesky per candidate serves all G grid points (the reference factorises per (candidate, grid point)).
// One warp per candidate; nb <= 31 (botorch prunes the baseline to the few points that can still be the best:
// hydra_config/acq_fn/cnei.yaml `prune_baseline: true`).
#pragma once
#include "common.cuh"
#include "rows.cuh"
#include "score_q.cuh"

namespace cbo {

constexpr int NEI_MAX_NB = 31;

struct NeiArgs {
    int N, Np, d, nb, G, S, n_cand, tiles_n;
    double noise, kappa, temp;
    const double* Xq;        // n_cand x d
    const double* inv_ls;    // d
    const double* Xb;        // nb x d
    const double* KbX;       // nb x Np   k(X_b, X)
    const double* Sbb;       // nb x nb   posterior covariance of f(X_b) given the data (noise-free)
    const double* mub;       // nb        posterior mean of f(X_b)
    const double* W;         // n_cand x Np
    const double* rn_partial;
    const double* mu;        // n_cand
    const double* cscale;    // n_cand or nullptr
    const double* eps_grid;  // G or n_cand x G
    int eps_per_candidate;
    const double* eps_mc;    // S x G x (nb+1)
    double* values;          // n_cand x S x G (x q)
    int raw;                 // 1: write the joint-posterior samples f(x) themselves (cNEHVI), 0: the noisy-EI improvement
    const double* V;         // joint batches only: rows x Np (v = L^-1 u); rows = n_cand * q
};

__global__ void __launch_bounds__(128)
k_nei_values(const NeiArgs p) {
    __shared__ double sL[4][(NEI_MAX_NB + 1) * (NEI_MAX_NB + 1)];
    __shared__ double sc[4][NEI_MAX_NB + 1];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 4 + warp;
    if (c >= p.n_cand) return;
    const int nb = p.nb, n1 = nb + 1;
    double* L = sL[warp];
    double* cv = sc[warp];
    double vnorm = 0.0;
    for (int t = 0; t < p.tiles_n; ++t) vnorm += p.rn_partial[int64_t(c) * p.tiles_n + t];
    const double s = p.kappa + p.noise - vnorm, P = s - p.noise, mu = p.mu[c];
    // c_j: lanes stride over the train points, one baseline point at a time
    const double* Wc = p.W + int64_t(c) * p.Np;
    for (int j = 0; j < nb; ++j) {
        double acc = 0.0;
        for (int i = lane; i < p.N; i += 32) acc = fma(p.KbX[int64_t(j) * p.Np + i], Wc[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            double r2 = 0.0;
            for (int e = 0; e < p.d; ++e) {
                double df = (p.Xb[j * p.d + e] - p.Xq[int64_t(c) * p.d + e]) * p.inv_ls[e];
                r2 = fma(df, df, r2);
            }
            double r = sqrt(fmax(r2, 1e-30));
            cv[j] = p.kappa * (1.0 + kSqrt5 * r + (5.0 / 3.0) * r * r) * exp(-kSqrt5 * r) - acc;
        }
    }
    __syncwarp();
    // covariance (lower triangle), lane = row
    if (lane < n1) {
        for (int k = 0; k <= lane; ++k) {
            double v;
            if (lane < nb) v = p.Sbb[lane * nb + k] - cv[lane] * cv[k] / s;
            else v = (k < nb) ? cv[k] * p.noise / s : P * p.noise / s;
            L[lane * n1 + k] = v;
        }
    }
    __syncwarp();
    // right-looking Cholesky, lane = row
    bool ok = true;
    for (int j = 0; j < n1; ++j) {
        double djj = L[j * n1 + j];
        ok = ok && (djj > 0.0);
        double ljj = sqrt(djj);
        __syncwarp();
        if (lane == j) L[j * n1 + j] = ljj;
        if (lane > j && lane < n1) L[lane * n1 + j] /= ljj;
        __syncwarp();
        if (lane > j && lane < n1) {
            double lij = L[lane * n1 + j];
            for (int k = j + 1; k <= lane; ++k) L[lane * n1 + k] -= lij * L[k * n1 + j];
        }
        __syncwarp();
    }
    const double cscale = p.cscale ? p.cscale[c] : 1.0;
    const double gsd = sqrt(cscale * s);
    const double inv_temp = p.temp > 0.0 ? 1.0 / p.temp : 0.0;
    for (int sg = lane; sg < p.S * p.G; sg += 32) {
        const int smp = sg / p.G, g = sg - smp * p.G;
        double eps = p.eps_grid[p.eps_per_candidate ? int64_t(c) * p.G + g : g];
        double bg = gsd * eps / s;
        const double* e = p.eps_mc + (int64_t(smp) * p.G + g) * n1;
        // f = mean + L eps ; smooth max over the baseline, softplus of the improvement
        double mx = -INFINITY, fx = 0.0;
        double fb[NEI_MAX_NB];
        for (int r = 0; r < n1; ++r) {
            double acc = (r < nb) ? p.mub[r] + cv[r] * bg : mu + P * bg;
            for (int k = 0; k <= r; ++k) acc = fma(L[r * n1 + k], e[k], acc);
            if (r < nb) { fb[r] = acc; mx = fmax(mx, acc); }
            else fx = acc;
        }
        double val;
        if (p.temp > 0.0) {
            double se = 0.0;
            for (int r = 0; r < nb; ++r) se += exp((fb[r] - mx) * inv_temp);
            double smax = mx + p.temp * log(se);                     // temp * logsumexp(f_b / temp)
            double xt = (fx - smax) * inv_temp;
            val = p.temp * (fmax(xt, 0.0) + log1p(exp(-fabs(xt))));
        } else {
            val = fmax(fx - mx, 0.0);
        }
        if (p.raw) val = fx;
        p.values[(int64_t(c) * p.S + smp) * p.G + g] = ok ? val : nan("");
    }
}

// Joint batches (q = Q candidate points per batch, nb + Q <= 32): with c = Cov_D(f(X_b), y(X)) (nb x Q),
// S = K_qq + noise I - V^T V and B_g = S^-1 (y_g - mu) the refit joint posterior over [X_baseline; X] is
//   mean = [mu_b + c B_g ; mu + (y_g - mu) - noise B_g],
//   cov  = [[Sigma_bb - c S^-1 c^T, noise c S^-1], [., noise (I - noise S^-1)]]       -- independent of the grid point.
// One warp per joint batch (lane = row of the (nb+Q)^2 factor); values: n_cand x S x G x Q, read by k_score_q.
template <int Q>
__global__ void __launch_bounds__(128)
k_nei_values_q(const NeiArgs p) {
    __shared__ double sL[4][32 * 32];
    __shared__ double scv[4][NEI_MAX_NB * Q];
    __shared__ double scs[4][NEI_MAX_NB * Q];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * 4 + warp;
    if (c >= p.n_cand) return;
    const int nb = p.nb, m = nb + Q, r0 = c * Q;
    double* L = sL[warp];
    double* cv = scv[warp];      // c[r][j]
    double* cs = scs[warp];      // (c S^-1)[r][j]
    // Schur complement of the joint batch, its inverse and the factor of the (row-scaled) grid covariance: every lane redundantly
    double Sm[Q][Q], Ls[Q][Q], Si[Q][Q], Cov[Q][Q], Lc[Q][Q];
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
    if (lane < nb) {
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double a = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k) a = fma(cv[lane * Q + k], 0.5 * (Si[k][j] + Si[j][k]), a);
            cs[lane * Q + j] = a;
        }
    }
    __syncwarp();
    // covariance (lower triangle), lane = row
    if (lane < m) {
        for (int k = 0; k <= lane; ++k) {
            double v;
            if (lane < nb) {
                v = p.Sbb[lane * nb + k];
#pragma unroll
                for (int j = 0; j < Q; ++j) v -= cs[lane * Q + j] * cv[k * Q + j];
            } else {
                const int j = lane - nb;
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
            L[lane * m + k] = v;
        }
    }
    __syncwarp();
    for (int j = 0; j < m; ++j) {
        const double djj = L[j * m + j];
        ok = ok && (djj > 0.0);
        const double ljj = sqrt(djj);
        __syncwarp();
        if (lane == j) L[j * m + j] = ljj;
        if (lane > j && lane < m) L[lane * m + j] /= ljj;
        __syncwarp();
        if (lane > j && lane < m) {
            const double lij = L[lane * m + j];
            for (int k = j + 1; k <= lane; ++k) L[lane * m + k] -= lij * L[k * m + j];
        }
        __syncwarp();
    }
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
        double mx = -INFINITY;
        double fb[NEI_MAX_NB], fx[Q];
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
            for (int k = 0; k <= r; ++k) acc = fma(L[r * m + k], e[k], acc);
            if (r < nb) { fb[r] = acc; mx = fmax(mx, acc); }
            else {
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    if (j == r - nb) fx[j] = acc;
            }
        }
        double smax = mx;
        if (p.temp > 0.0) {
            double se = 0.0;
            for (int r = 0; r < nb; ++r) se += exp((fb[r] - mx) * inv_temp);
            smax = mx + p.temp * log(se);                            // temp * logsumexp(f_b / temp)
        }
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

inline cudaError_t launch_nei_values_q(const NeiArgs& a, int q, cudaStream_t s) {
    const int grid = (a.n_cand + 3) / 4;
    switch (q) {
        case 2: CBO_LAUNCHED(), k_nei_values_q<2><<<grid, 128, 0, s>>>(a); break;
        case 3: CBO_LAUNCHED(), k_nei_values_q<3><<<grid, 128, 0, s>>>(a); break;
        case 4: CBO_LAUNCHED(), k_nei_values_q<4><<<grid, 128, 0, s>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace cbo
