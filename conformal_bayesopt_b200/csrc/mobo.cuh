// mobo.cuh -- cEHVI inner expectation and its conformal integration over a multi-output grid (q = 1).
//
// qConformalExpectedHypervolumeImprovement (conformalbo/mobo_acquisitions.py:24-45): the model is m independent
// GPs; every output k contributes its own conformal mask / grid / log-density (one cbo_conformal_acq_f64 call per
// output), the joint grid point g is (y_g^1, ..., y_g^m).  _conformal_integration (acquisitions.py:225-248) collapses
// the output axis by PRODUCT for the masks and SUM for the log-density, the value is botorch's _compute_qehvi
// (q = 1):  mean_s sum_cells prod_k clamp(min(f_k, u_cell,k) - l_cell,k, 0)  with
// f_k = refit mean_k(g) + refit sd_k * eps'_{s,g,k}.
// The reference's own EHVI integration collapses the wrong axes and raises unless b in {1, G} (App. C-5): this
// kernel implements the evident intent (product over outputs, integrate over the grid, one value per candidate).
#pragma once
#include "common.cuh"

namespace cbo {

constexpr int MAX_M = 4;

struct EhviArgs {
    int m, b, G, S, n_cells, optimistic;
    double alpha_lvl;
    const double* mask[MAX_M];     // b x G
    const double* logp[MAX_M];     // b x G
    const double* grid[MAX_M];     // b x G
    const double* mu[MAX_M];       // b
    const double* ood[MAX_M];      // b
    const double* mf[MAX_M];       // b x G   refit posterior mean at the candidate
    const double* sd[MAX_M];       // b       refit posterior sd at the candidate
    const double* eps_mc;          // S x G x m
    const double* cell_lower;      // n_cells x m
    const double* cell_upper;      // n_cells x m
    double* acq;                   // b
};

__global__ void __launch_bounds__(256)
k_ehvi(const EhviArgs p) {
    extern __shared__ double s_cells[];            // lower | upper
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int e = threadIdx.x; e < p.n_cells * p.m; e += blockDim.x) {
        s_cells[e] = p.cell_lower[e];
        s_cells[p.n_cells * p.m + e] = p.cell_upper[e];
    }
    __syncthreads();
    const double* lo = s_cells;
    const double* up = s_cells + p.n_cells * p.m;
    const int c = blockIdx.x * (blockDim.x >> 5) + warp;
    if (c >= p.b) return;
    double ood = 1.0, sdk[MAX_M], muk[MAX_M];
#pragma unroll
    for (int k = 0; k < MAX_M; ++k)
        if (k < p.m) { ood *= p.ood[k][c]; sdk[k] = p.sd[k][c]; muk[k] = p.mu[k][c]; }
    double scw = 0.0, snc = 0.0, A = 0.0, B = 0.0;
    for (int g = lane; g < p.G; g += 32) {
        double mk = 1.0, op = 1.0, lp = 0.0, mfk[MAX_M];
#pragma unroll
        for (int k = 0; k < MAX_M; ++k)
            if (k < p.m) {
                int64_t ix = int64_t(c) * p.G + g;
                mk *= p.mask[k][ix];
                lp += p.logp[k][ix];
                if (p.optimistic) op *= (p.grid[k][ix] >= muk[k]) ? 1.0 : 0.0;
                mfk[k] = p.mf[k][ix];
            }
        double val = 0.0;
        for (int s = 0; s < p.S; ++s) {
            double f[MAX_M];
#pragma unroll
            for (int k = 0; k < MAX_M; ++k)
                if (k < p.m) f[k] = fma(sdk[k], p.eps_mc[(int64_t(s) * p.G + g) * p.m + k], mfk[k]);
            double hv = 0.0;
            for (int cell = 0; cell < p.n_cells; ++cell) {
                double vol = 1.0;
#pragma unroll
                for (int k = 0; k < MAX_M; ++k)
                    if (k < p.m) vol *= fmax(fmin(f[k], up[cell * p.m + k]) - lo[cell * p.m + k], 0.0);
                hv += vol;
            }
            val += hv;
        }
        val /= double(p.S);
        double cw = mk * op / fmax(exp(lp), 1e-6), nc = (1.0 - mk) * op;
        scw += cw; snc += nc;
        A = fma(cw, val, A); B = fma(nc, val, B);
    }
    scw = warp_sum(scw); snc = warp_sum(snc); A = warp_sum(A); B = warp_sum(B);
    if (lane == 0)
        p.acq[c] = (1.0 - ood) * ((1.0 - p.alpha_lvl) * A / fmax(scw, 1e-6) + p.alpha_lvl * B / fmax(snc, 1e-6));
}

// ---- joint batches (q <= 4, incl. pending points) and per-sample cells (cNEHVI) -----------------------------------------
// k_ehvi_q: the same integration for q candidate points evaluated jointly.  Per output k the per-output call
// (k_score / k_score_q) supplies mask (b x G x q), the joint log-density (b x G), the grid, mu, ood and the q x q Schur
// complement S_k; the refit posterior at the candidates is  mean = mu + (I - noise S^-1)(y_g - mu),
// cov = noise (I - noise sym(S^-1))  (SURVEY.md App. A-6/A-8), sampled with the lower Cholesky factor and the base samples
// eps_mc (S x G x q x m).  Masks / optimistic masks / ood multiply over outputs AND over the q points, log-densities add
// (acquisitions.py:230-235 collapse every trailing axis), the value is botorch's _compute_qehvi: per cell, inclusion-
// exclusion over the non-empty subsets J of the q points of  prod_k clamp(min(min_{j in J} f_jk, u_k) - l_k, 0), summed
// over cells, averaged over the MC samples (App. B-9).
// cNEHVI (mobo_acquisitions.py:196-213): the samples f(X) come from the JOINT refit posterior over [X_baseline; X] of every
// output (k_nei_values*, raw-sample mode: fx[k], b x S x G x q) and every MC sample s has its own box decomposition
// (cells_per_sample: cell_lower / cell_upper are S x n_cells x m).
template <int Q>
__device__ __forceinline__ bool small_chol(const double* A, double* L) {      // row-major Q x Q, lower triangle of A read
    bool ok = true;
    for (int j = 0; j < Q; ++j) {
        double dsum = A[j * Q + j];
        for (int k = 0; k < j; ++k) dsum -= L[j * Q + k] * L[j * Q + k];
        ok = ok && (dsum > 0.0);
        const double ljj = sqrt(dsum);
        L[j * Q + j] = ljj;
        for (int i = 0; i < Q; ++i) {
            if (i < j) L[i * Q + j] = 0.0;
            if (i > j) {
                double s = A[i * Q + j];
                for (int k = 0; k < j; ++k) s -= L[i * Q + k] * L[j * Q + k];
                L[i * Q + j] = s / ljj;
            }
        }
    }
    return ok;
}

struct EhviQArgs {
    int m, b, G, S, q, n_cells, optimistic, cells_per_sample;
    double alpha_lvl;
    double noise[MAX_M];
    const double* mask[MAX_M];     // b x G x q
    const double* logp[MAX_M];     // b x G
    const double* grid[MAX_M];     // b x G x q
    const double* mu[MAX_M];       // b x q
    const double* ood[MAX_M];      // b x q
    const double* cov[MAX_M];      // b x q x q   S = K_qq + noise I - V^T V  (internal sampling; unused with fx)
    const double* fx[MAX_M];       // b x S x G x q  joint-posterior samples of f(X) (cNEHVI) or nullptr
    const double* eps_mc;          // S x G x q x m  (internal sampling)
    const double* cell_lower;      // [S x] n_cells x m
    const double* cell_upper;
    double* acq;                   // b
    double* values;                // b x G or nullptr: MC-averaged hyper-volume improvement per grid point
    // reverse pass (k_ehvi_q<Q, true>; values is then required): adjoints of d acq[c] for the per-output gradient calls
    // (cbo_ext_adjoints_f64)
    double* mask_bar[MAX_M];       // b x G x q
    double* ood_bar[MAX_M];        // b x q
    double* f_bar[MAX_M];          // b x S x G x q   adjoint of the samples (internal sampling or fx)
};

constexpr int EHVIQ_WARPS = 8;

template <int Q, bool BWD = false>
__global__ void __launch_bounds__(EHVIQ_WARPS * 32)
k_ehvi_q(const EhviQArgs p) {
    extern __shared__ double s_dyn[];              // cells: lower | upper ; then per warp: m x (shr[Q*Q], Lf[Q*Q])
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m = p.m, G = p.G, S = p.S;
    const int ncell_tot = p.n_cells * (p.cells_per_sample ? S : 1);
    for (int e = threadIdx.x; e < ncell_tot * m; e += blockDim.x) {
        s_dyn[e] = p.cell_lower[e];
        s_dyn[ncell_tot * m + e] = p.cell_upper[e];
    }
    const double* lo = s_dyn;
    const double* up = s_dyn + ncell_tot * m;
    double* wmat = s_dyn + 2 * ncell_tot * m + size_t(warp) * MAX_M * 2 * Q * Q;
    __syncthreads();
    const int c = blockIdx.x * EHVIQ_WARPS + warp;
    if (c >= p.b) return;
    int bad = 0;
    if (lane < m && !p.fx[lane]) {                 // one lane per output: shrinkage and refit factor of this candidate batch
        const int k = lane;
        double Sm[Q * Q], Ls[Q * Q], Li[Q * Q], Si[Q * Q], Cf[Q * Q];
        for (int e = 0; e < Q * Q; ++e) Sm[e] = p.cov[k][int64_t(c) * Q * Q + e];
        bad |= !small_chol<Q>(Sm, Ls);
        for (int j = 0; j < Q; ++j)                // L^-1 (lower)
            for (int i = 0; i < Q; ++i) {
                if (i < j) Li[i * Q + j] = 0.0;
                else if (i == j) Li[i * Q + j] = 1.0 / Ls[i * Q + i];
                else {
                    double s = 0.0;
                    for (int kk = j; kk < i; ++kk) s += Ls[i * Q + kk] * Li[kk * Q + j];
                    Li[i * Q + j] = -s / Ls[i * Q + i];
                }
            }
        for (int i = 0; i < Q; ++i)
            for (int j = 0; j < Q; ++j) {
                double s = 0.0;
                for (int kk = 0; kk < Q; ++kk) s += Li[kk * Q + i] * Li[kk * Q + j];
                Si[i * Q + j] = s;
            }
        const double nz = p.noise[k];
        double* shr = wmat + size_t(k) * 2 * Q * Q;
        double* Lf = shr + Q * Q;
        for (int i = 0; i < Q; ++i)
            for (int j = 0; j < Q; ++j) {
                shr[i * Q + j] = ((i == j) ? 1.0 : 0.0) - nz * Si[i * Q + j];
                Cf[i * Q + j] = nz * (((i == j) ? 1.0 : 0.0) - nz * 0.5 * (Si[i * Q + j] + Si[j * Q + i]));
            }
        if (Q == 1) { Lf[0] = sqrt(fmax(Cf[0], 0.0)); }
        else bad |= !small_chol<Q>(Cf, Lf);
    }
    bad = __any_sync(0xffffffffu, bad);
    __syncwarp();
    double ood = 1.0;
    for (int k = 0; k < m; ++k)
        for (int j = 0; j < Q; ++j) ood *= p.ood[k][int64_t(c) * Q + j];
    double scw = 0.0, snc = 0.0, A = 0.0, B = 0.0;
    for (int g = lane; g < G; g += 32) {
        double mk = 1.0, op = 1.0, lp = 0.0;
        double mf[MAX_M][Q];
        for (int k = 0; k < m; ++k) {
            const int64_t ix = (int64_t(c) * G + g) * Q;
            lp += p.logp[k][int64_t(c) * G + g];
            double dev[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                const double muj = p.mu[k][int64_t(c) * Q + j], yg = p.grid[k][ix + j];
                mk *= p.mask[k][ix + j];
                if (p.optimistic) op *= (yg >= muj) ? 1.0 : 0.0;
                dev[j] = yg - muj;
                mf[k][j] = muj;
            }
            if (!p.fx[k]) {
                const double* shr = wmat + size_t(k) * 2 * Q * Q;
#pragma unroll
                for (int j = 0; j < Q; ++j)
#pragma unroll
                    for (int l = 0; l < Q; ++l) mf[k][j] = fma(shr[j * Q + l], dev[l], mf[k][j]);
            }
        }
        double val = 0.0;
        for (int s = 0; s < S; ++s) {
            double f[MAX_M][Q];
            for (int k = 0; k < m; ++k) {
                if (p.fx[k]) {
#pragma unroll
                    for (int j = 0; j < Q; ++j) f[k][j] = p.fx[k][((int64_t(c) * S + s) * G + g) * Q + j];
                } else {
                    const double* Lf = wmat + size_t(k) * 2 * Q * Q + Q * Q;
#pragma unroll
                    for (int j = 0; j < Q; ++j) {
                        double a = mf[k][j];
#pragma unroll
                        for (int l = 0; l < Q; ++l)
                            if (l <= j) a = fma(Lf[j * Q + l], p.eps_mc[((int64_t(s) * G + g) * Q + l) * m + k], a);
                        f[k][j] = a;
                    }
                }
            }
            const double* los = lo + (p.cells_per_sample ? size_t(s) * p.n_cells * m : 0);
            const double* ups = up + (p.cells_per_sample ? size_t(s) * p.n_cells * m : 0);
            double hv = 0.0;
            for (int sub = 1; sub < (1 << Q); ++sub) {          // inclusion-exclusion over the non-empty subsets of the q points
                double vmin[MAX_M];
                for (int k = 0; k < m; ++k) {
                    double v = INFINITY;
#pragma unroll
                    for (int j = 0; j < Q; ++j)
                        if ((sub >> j) & 1) v = fmin(v, f[k][j]);
                    vmin[k] = v;
                }
                double tot = 0.0;
                for (int cell = 0; cell < p.n_cells; ++cell) {
                    double vol = 1.0;
                    for (int k = 0; k < m; ++k) vol *= fmax(fmin(vmin[k], ups[cell * m + k]) - los[cell * m + k], 0.0);
                    tot += vol;
                }
                hv += (__popc(sub) & 1) ? tot : -tot;
            }
            val += hv;
        }
        val /= double(S);
        if (p.values) p.values[int64_t(c) * G + g] = bad ? nan("") : val;
        const double cw = mk * op / fmax(exp(lp), 1e-6), nc = (1.0 - mk) * op;
        scw += cw; snc += nc;
        A = fma(cw, val, A); B = fma(nc, val, B);
    }
    scw = warp_sum(scw); snc = warp_sum(snc); A = warp_sum(A); B = warp_sum(B);
    if (lane == 0) {
        const double r = (1.0 - ood) * ((1.0 - p.alpha_lvl) * A / fmax(scw, 1e-6) + p.alpha_lvl * B / fmax(snc, 1e-6));
        p.acq[c] = bad ? nan("") : r;
    }
    if constexpr (BWD) {
        // ---- reverse: d acq[c] / d (mask_kgj, ood_kj, f_sgkj); the grid, its log-density and the optimistic mask are constants ----
        __syncwarp();
        const double dcw = fmax(scw, 1e-6), dnc = fmax(snc, 1e-6);
        const double inner = (1.0 - p.alpha_lvl) * A / dcw + p.alpha_lvl * B / dnc;
        const double nanv = nan("");
        if (lane == 0)
            for (int k = 0; k < m; ++k)
                for (int j = 0; j < Q; ++j) {
                    double others = 1.0;
                    for (int k2 = 0; k2 < m; ++k2)
                        for (int j2 = 0; j2 < Q; ++j2)
                            if (k2 != k || j2 != j) others *= p.ood[k2][int64_t(c) * Q + j2];
                    p.ood_bar[k][int64_t(c) * Q + j] = bad ? nanv : -inner * others;
                }
        const double c1 = (1.0 - ood) * (1.0 - p.alpha_lvl), c2 = (1.0 - ood) * p.alpha_lvl;
        for (int g = lane; g < G; g += 32) {
            const int64_t ix = (int64_t(c) * G + g) * Q;
            double mk = 1.0, op = 1.0, lp = 0.0;
            double mf[MAX_M][Q];
            for (int k = 0; k < m; ++k) {
                lp += p.logp[k][int64_t(c) * G + g];
                double dev[Q];
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    const double muj = p.mu[k][int64_t(c) * Q + j], yg = p.grid[k][ix + j];
                    mk *= p.mask[k][ix + j];
                    if (p.optimistic) op *= (yg >= muj) ? 1.0 : 0.0;
                    dev[j] = yg - muj;
                    mf[k][j] = muj;
                }
                if (!p.fx[k]) {
                    const double* shr = wmat + size_t(k) * 2 * Q * Q;
#pragma unroll
                    for (int j = 0; j < Q; ++j)
#pragma unroll
                        for (int l = 0; l < Q; ++l) mf[k][j] = fma(shr[j * Q + l], dev[l], mf[k][j]);
                }
            }
            const double kd = 1.0 / fmax(exp(lp), 1e-6);
            const double val = p.values[int64_t(c) * G + g];
            const double cw = mk * op * kd, nc = (1.0 - mk) * op;
            const double val_bar = (c1 * cw / dcw + c2 * nc / dnc) / double(S);
            const double cw_bar = c1 * (val / dcw - ((scw > 1e-6) ? A / (dcw * dcw) : 0.0));
            const double nc_bar = c2 * (val / dnc - ((snc > 1e-6) ? B / (dnc * dnc) : 0.0));
            const double mk_bar = (cw_bar * kd - nc_bar) * op;
            for (int k = 0; k < m; ++k)
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    double others = 1.0;
                    for (int k2 = 0; k2 < m; ++k2)
#pragma unroll
                        for (int j2 = 0; j2 < Q; ++j2)
                            if (k2 != k || j2 != j) others *= p.mask[k2][ix + j2];
                    p.mask_bar[k][ix + j] = bad ? nanv : mk_bar * others;
                }
            for (int s = 0; s < S; ++s) {
                double f[MAX_M][Q], fb[MAX_M][Q];
                for (int k = 0; k < m; ++k) {
#pragma unroll
                    for (int j = 0; j < Q; ++j) fb[k][j] = 0.0;
                    if (p.fx[k]) {
#pragma unroll
                        for (int j = 0; j < Q; ++j) f[k][j] = p.fx[k][((int64_t(c) * S + s) * G + g) * Q + j];
                    } else {
                        const double* Lf = wmat + size_t(k) * 2 * Q * Q + Q * Q;
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            double a = mf[k][j];
#pragma unroll
                            for (int l = 0; l < Q; ++l)
                                if (l <= j) a = fma(Lf[j * Q + l], p.eps_mc[((int64_t(s) * G + g) * Q + l) * m + k], a);
                            f[k][j] = a;
                        }
                    }
                }
                const double* los = lo + (p.cells_per_sample ? size_t(s) * p.n_cells * m : 0);
                const double* ups = up + (p.cells_per_sample ? size_t(s) * p.n_cells * m : 0);
                for (int sub = 1; sub < (1 << Q); ++sub) {
                    double vmin[MAX_M], dv[MAX_M];
                    int arg[MAX_M];
                    for (int k = 0; k < m; ++k) {
                        double v = INFINITY;
                        int a = 0;
#pragma unroll
                        for (int j = 0; j < Q; ++j)
                            if (((sub >> j) & 1) && f[k][j] < v) { v = f[k][j]; a = j; }
                        vmin[k] = v; arg[k] = a; dv[k] = 0.0;
                    }
                    for (int cell = 0; cell < p.n_cells; ++cell) {
                        double len[MAX_M];
                        for (int k = 0; k < m; ++k) len[k] = fmax(fmin(vmin[k], ups[cell * m + k]) - los[cell * m + k], 0.0);
                        for (int k = 0; k < m; ++k) {
                            if (!(vmin[k] < ups[cell * m + k]) || !(len[k] > 0.0)) continue;
                            double o = 1.0;
                            for (int k2 = 0; k2 < m; ++k2)
                                if (k2 != k) o *= len[k2];
                            dv[k] += o;
                        }
                    }
                    const double sgn = (__popc(sub) & 1) ? 1.0 : -1.0;
                    for (int k = 0; k < m; ++k)
#pragma unroll
                        for (int j = 0; j < Q; ++j)
                            if (j == arg[k]) fb[k][j] = fma(sgn, dv[k], fb[k][j]);
                }
                for (int k = 0; k < m; ++k)
                    if (p.f_bar[k])
#pragma unroll
                        for (int j = 0; j < Q; ++j)
                            p.f_bar[k][((int64_t(c) * S + s) * G + g) * Q + j] = bad ? nanv : val_bar * fb[k][j];
            }
        }
    }
}

inline size_t ehvi_q_smem_bytes(int m, int q, int n_cells_total) {
    return sizeof(double) * (size_t(2) * n_cells_total * m + size_t(EHVIQ_WARPS) * MAX_M * 2 * q * q);
}

}  // namespace cbo
