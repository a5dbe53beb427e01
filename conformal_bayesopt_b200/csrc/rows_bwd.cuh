// rows_bwd.cuh -- last stage of d acq / d X (q = 1): from the adjoint of the kernel row back to the candidate.
//
//   u_bar_i = (A^-1 w_bar')_i + mu_bar alpha_i - 2 s_bar w_i
//             (w = A^-1 u, mu = m + u.alpha, s = kappa + noise - u.w; SURVEY.md App. A-1.  The probe term
//             sum_s ut_bar_s T_is = (A^-1 z ut_bar)_i is already folded into w_bar' = w_bar + z ut_bar by k_score_bwd)
//   x_bar_k = (1/l_k) sum_i u_bar_i k'(r_i) (x_k - X_ik)/l_k / r_i,   k'(r)/r = -(5 os/3)(1 + sqrt5 r) exp(-sqrt5 r)
//             (Matern-5/2: no singularity at r = 0)
//   (ratio_mma.cuh adds drc_bar * d ratio(x) / d x afterwards)
//
// Same organisation as k_kernel_rows: persistent 16-warp CTAs, XsT / alpha / T^T staged in shared memory, one warp owns
// ROWS_CPW candidates and sweeps i = lane, lane + 32, ...; the two row reads (A^-1 w_bar and w) are 256-byte coalesced
// streaming loads.  Input dimensions are processed in register chunks of 8 (one sweep for d <= 8).
#pragma once
#include "rows.cuh"
#include "score_bwd.cuh"

namespace cbo {

constexpr int RB_DK = 8;

template <bool STAGE>
__global__ void __launch_bounds__(ROWS_WARPS * 32, 1)
k_rows_bwd(const double* __restrict__ Xq, int n, const double* __restrict__ XsT_g, int N, int Np, int d,
           const double* __restrict__ inv_ls, double outputscale, const double* __restrict__ alpha_g,
           const double* __restrict__ Aw, const double* __restrict__ W,
           const double* __restrict__ adj, double* __restrict__ dX) {
    extern __shared__ double sx[];  // tab[16] | [WARPS][CPW][d] coords | (STAGE) alpha[Np] | XsT[d][Np]
    double* tab = sx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* myx = sx + 16 + warp * ROWS_CPW * d;
    double* s_alpha = sx + 16 + ROWS_WARPS * ROWS_CPW * d;
    double* s_xst = s_alpha + Np;
    if (threadIdx.x < 16) tab[threadIdx.x] = exp2(double(threadIdx.x) / 16.0);
    if (STAGE) {
        for (int e = threadIdx.x; e < Np; e += blockDim.x) s_alpha[e] = alpha_g[e];
        for (int e = threadIdx.x; e < d * Np; e += blockDim.x) s_xst[e] = XsT_g[e];
    }
    __syncthreads();
    const double* __restrict__ XsT = STAGE ? s_xst : XsT_g;
    const double* __restrict__ alpha = STAGE ? s_alpha : alpha_g;
    const double kfac = -(5.0 / 3.0) * outputscale;
    const int groups = (n + ROWS_CPW - 1) / ROWS_CPW;
    for (int grp = blockIdx.x * ROWS_WARPS + warp; grp < groups; grp += gridDim.x * ROWS_WARPS) {
        const int c0 = grp * ROWS_CPW;
        __syncwarp();
        for (int e = lane; e < ROWS_CPW * d; e += 32) {
            int cc = e / d, k = e - cc * d;
            int c = min(c0 + cc, n - 1);
            myx[e] = Xq[int64_t(c) * d + k] * inv_ls[k];
        }
        __syncwarp();
        double mu_bar[ROWS_CPW], s2_bar[ROWS_CPW];
#pragma unroll
        for (int cc = 0; cc < ROWS_CPW; ++cc) {
            const int c = min(c0 + cc, n - 1);
            const double* a = adj + int64_t(c) * ADJ_STRIDE;
            mu_bar[cc] = a[0];
            s2_bar[cc] = -2.0 * a[1];
        }
        for (int k0 = 0; k0 < d; k0 += RB_DK) {
            double acc[ROWS_CPW][RB_DK];
#pragma unroll
            for (int cc = 0; cc < ROWS_CPW; ++cc)
#pragma unroll
                for (int kk = 0; kk < RB_DK; ++kk) acc[cc][kk] = 0.0;
            for (int i = lane; i < N; i += 32) {
                double r2[ROWS_CPW];
#pragma unroll
                for (int cc = 0; cc < ROWS_CPW; ++cc) r2[cc] = 0.0;
#pragma unroll 4
                for (int k = 0; k < d; ++k) {
                    double xi = XsT[k * Np + i];
#pragma unroll
                    for (int cc = 0; cc < ROWS_CPW; ++cc) {
                        double df = myx[cc * d + k] - xi;
                        r2[cc] = fma(df, df, r2[cc]);
                    }
                }
                const double al = alpha[i];
                double coef[ROWS_CPW];
#pragma unroll
                for (int cc = 0; cc < ROWS_CPW; ++cc) {
                    const int c = min(c0 + cc, n - 1);
                    double ub = __ldcs(Aw + int64_t(c) * Np + i);
                    ub = fma(mu_bar[cc], al, ub);
                    ub = fma(s2_bar[cc], __ldcs(W + int64_t(c) * Np + i), ub);
                    coef[cc] = ub;
                }
#pragma unroll
                for (int cc = 0; cc < ROWS_CPW; ++cc) {
                    double t = kSqrt5 * sqrt_fast(fmax(r2[cc], 1e-30));
                    coef[cc] *= kfac * (1.0 + t) * exp_neg_fast(t, tab);          // u_bar_i k'(r_i)/r_i
                }
#pragma unroll
                for (int kk = 0; kk < RB_DK; ++kk) {
                    if (k0 + kk < d) {
                        double xi = XsT[(k0 + kk) * Np + i];
#pragma unroll
                        for (int cc = 0; cc < ROWS_CPW; ++cc)
                            acc[cc][kk] = fma(coef[cc], myx[cc * d + k0 + kk] - xi, acc[cc][kk]);
                    }
                }
            }
#pragma unroll
            for (int cc = 0; cc < ROWS_CPW; ++cc)
#pragma unroll
                for (int kk = 0; kk < RB_DK; ++kk) {
                    if (k0 + kk < d) {
                        double v = warp_sum(acc[cc][kk]);
                        if (lane == 0 && c0 + cc < n) dX[int64_t(c0 + cc) * d + k0 + kk] = v * inv_ls[k0 + kk];
                    }
                }
        }
    }
}

template <bool STAGE>
inline cudaError_t launch_rows_bwd_t(const double* Xq, int n, const double* XsT, int N, int Np, int d, const double* inv_ls,
                                     double outputscale, const double* alpha, const double* Aw,
                                     const double* W, const double* adj, double* dX, int num_sms, cudaStream_t st) {
    const size_t smem = rows_smem_bytes(Np, d, 0, STAGE);
    static SmemOptIn optin;
    if (cudaError_t e = optin.ensure(k_rows_bwd<STAGE>, smem); e != cudaSuccess) return e;
    const int groups = (n + ROWS_CPW - 1) / ROWS_CPW;
    int grid = (groups + ROWS_WARPS - 1) / ROWS_WARPS;
    if (grid > num_sms) grid = num_sms;
    CBO_LAUNCHED(), k_rows_bwd<STAGE><<<grid, ROWS_WARPS * 32, smem, st>>>(Xq, n, XsT, N, Np, d, inv_ls, outputscale, alpha, Aw, W, adj, dX);
    return cudaGetLastError();
}

inline cudaError_t launch_rows_bwd(const double* Xq, int n, const double* XsT, int N, int Np, int d, const double* inv_ls,
                                   double outputscale, const double* alpha, const double* Aw, const double* W,
                                   const double* adj, double* dX, int num_sms, cudaStream_t st) {
    if (rows_smem_bytes(Np, d, 0, true) <= 220 * 1024)
        return launch_rows_bwd_t<true>(Xq, n, XsT, N, Np, d, inv_ls, outputscale, alpha, Aw, W, adj, dX, num_sms, st);
    return launch_rows_bwd_t<false>(Xq, n, XsT, N, Np, d, inv_ls, outputscale, alpha, Aw, W, adj, dX, num_sms, st);
}

}  // namespace cbo
