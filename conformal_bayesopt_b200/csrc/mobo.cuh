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

}  // namespace cbo
