// score_q.cuh -- joint batches q > 1: rank-q bordered fantasy update, per-column conformal rank, joint grid.
//
// Same path as score.cuh (helpers.py:94-333, acquisitions.py:53-76, 88-98, 151-159, 225-248) for q candidate points
// evaluated JOINTLY:
//   S = K_qq + noise I - V^T V,  y_g = mu + chol(c S) eps_g,  B_g = S^-1 (y_g - mu),
//   r_i = noise (alpha_i - sum_j W_ij B_gj),  r_cj = noise B_gj,  C = S^-1 (z_c - U^T T),
//   var_i = noise (h0_i + noise/(Sp+1e-6) sum_j w_ij (z_i . C_j)),
// column j of the rank tensor thresholds on candidate j's score against ALL N + q scores, but its importance weights
// hold the N train weights and only candidate j's own weight (helpers.py:288-297); the randomised step takes its
// D from the LAST row only (helpers.py:307).  Grid refinement scales the ROWS of the grid covariance
// (helpers.py:170,178) and the Cholesky reads its lower triangle.
// One CTA (256 threads) per candidate batch: joint batches are scored in small numbers (batch_limit 8 x restarts), the
// q = 1 sweep kernel in score.cuh is the throughput path.  The (train point, grid point, column) soft ranks use the same
// 12-instruction sigmoid as score.cuh (pattern table + cubic + MUFU.RCP64H + Newton).
#pragma once
#include "common.cuh"
#include "rows.cuh"
#include "score.cuh"
#include "score_bwd.cuh"

namespace cbo {

constexpr int SQ_THREADS = 256;     // value+reverse kernel, and the forward kernel for large batches
constexpr int SQ_WARPS = 8;
constexpr int SQ_THREADS_SMALL = 512;   // forward kernel when a handful of joint batches must fill their SMs (latency)
constexpr int MAX_Q = 4;

struct ScoreQArgs {
    ScoreArgs base;          // W / rn_partial / mu / ut / drc are indexed by ROW = candidate * q + j
    const double* V;         // rows x Np  (v = L^-1 u, still intact after pass 2)
    const double* Xq;        // n_cand x q x d
    const double* inv_ls;    // d
    int d;
    double* post_cov;        // n_cand x q x q or nullptr
};

__host__ __device__ inline size_t score_q_smem_bytes(int Np, int G, int Q) {
    // per-i tables: lv, ap, dr, wp[Q] | per-(g,j): mask, opt, dev, thr, cnt, sr, sw | per-g: logp | small scratch
    return sizeof(double) * (size_t(3 + Q) * Np + size_t(7 * Q + 1) * G + 512);
}

template <int Q>
__device__ inline bool chol_lower(const double (&A)[Q][Q], double (&L)[Q][Q]) {
    bool ok = true;
#pragma unroll
    for (int j = 0; j < Q; ++j) {
#pragma unroll
        for (int i = 0; i < Q; ++i) L[i][j] = 0.0;
    }
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        double dsum = A[j][j];
#pragma unroll
        for (int k = 0; k < Q; ++k)
            if (k < j) dsum -= L[j][k] * L[j][k];
        ok = ok && (dsum > 0.0);
        double ljj = sqrt(dsum);
        L[j][j] = ljj;
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            if (i > j) {
                double s = A[i][j];                       // lower triangle only
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k < j) s -= L[i][k] * L[j][k];
                L[i][j] = s / ljj;
            }
        }
    }
    return ok;
}

template <int Q>
__device__ inline void inv_from_chol(const double (&L)[Q][Q], double (&Ainv)[Q][Q]) {
    double Li[Q][Q];                                      // L^-1 (lower)
#pragma unroll
    for (int j = 0; j < Q; ++j) {
#pragma unroll
        for (int i = 0; i < Q; ++i) {
            if (i < j) Li[i][j] = 0.0;
            else if (i == j) Li[i][j] = 1.0 / L[i][i];
            else {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k >= j && k < i) s += L[i][k] * Li[k][j];
                Li[i][j] = -s / L[i][i];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < Q; ++i)
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k) s += Li[k][i] * Li[k][j];
            Ainv[i][j] = s;
        }
}

__device__ __forceinline__ double block_sum(double v, double* red, int tid, int nwarps = SQ_WARPS) {
    v = warp_sum(v);
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < nwarps; ++w) t += red[w];
    return t;
}

template <int Q, bool SOFT>
__global__ void __launch_bounds__(Q <= 2 ? SQ_THREADS_SMALL : SQ_THREADS)      // q > 2 needs more than 128 registers
k_score_q(const ScoreQArgs pq) {
    const ScoreArgs& p = pq.base;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nthr = blockDim.x, nwarps = blockDim.x >> 5;       // 256 or 512 threads
    const int N = p.N, Np = p.Np, G = p.G, Sp = p.Sp, S = p.S;
    double* s_lv = sm;                       // Np
    double* s_ap = s_lv + Np;                // Np
    double* s_dr = s_ap + Np;                // Np
    double* s_wp = s_dr + Np;                // Q x Np
    double* g_mask = s_wp + size_t(Q) * Np;  // G x Q
    double* g_opt = g_mask + size_t(Q) * G;  // G x Q
    double* g_dev = g_opt + size_t(Q) * G;   // G x Q
    double* g_thr = g_dev + size_t(Q) * G;   // G x Q   candidate scores; raw sums of phase B until phase C has run
    double* g_cnt = g_thr + size_t(Q) * G;
    double* g_sr = g_cnt + size_t(Q) * G;
    double* g_sw = g_sr + size_t(Q) * G;
    double* g_lp = g_sw + size_t(Q) * G;     // G
    double* sc = g_lp + G;                   // scratch: 512 doubles
    double* red = sc;                        // [4]
    double* s_S = sc + 16;                   // Q x Q
    double* s_Sinv = s_S + 16;
    double* s_Lc = s_Sinv + 16;
    double* s_Lf = s_Lc + 16;
    double* s_C = s_Lf + 16;                 // Q x MAX_SP
    double* s_misc = s_C + Q * MAX_SP;       // var_c[Q], logdet, bad
    const double kvar = p.noise / (double(Sp) + 1e-6);
    const double inv_temp = SOFT ? 1.0 / p.temp : 0.0;
    const double c1q = inv_temp * 23.083120654223414;                    // 16 / (temp ln 2): sigmoid argument in table steps
    const double sum_dr = p.sum_dr[0];
    __shared__ __align__(128) double tab[16];
    if (tid < 16) exp_tab_init(tab, tid);
    const unsigned tab_s = smem_u32(tab);

    for (int i = tid; i < N; i += nthr) s_dr[i] = p.dr[i];
    __syncthreads();

    for (int c = blockIdx.x; c < p.n_cand; c += gridDim.x) {
        const int r0 = c * Q;
        // ---- phase 0: Schur complement, its factor / inverse, probe correction -------------------------------
        double gsum[Q][Q];
#pragma unroll
        for (int j = 0; j < Q; ++j)
#pragma unroll
            for (int k = 0; k < Q; ++k) gsum[j][k] = 0.0;
        for (int i = tid; i < N; i += nthr) {
            double v[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) v[j] = pq.V[int64_t(r0 + j) * Np + i];
#pragma unroll
            for (int j = 0; j < Q; ++j)
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k <= j) gsum[j][k] = fma(v[j], v[k], gsum[j][k]);
        }
#pragma unroll
        for (int j = 0; j < Q; ++j)
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k <= j) gsum[j][k] = block_sum(gsum[j][k], red, tid, nwarps);
        if (tid == 0) {
            double Sm[Q][Q], Ls[Q][Q], Si[Q][Q];
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k <= j; ++k) {
                    double r2 = 0.0;
                    for (int e = 0; e < pq.d; ++e) {
                        double df = (pq.Xq[(int64_t(r0 + j)) * pq.d + e] - pq.Xq[(int64_t(r0 + k)) * pq.d + e]) * pq.inv_ls[e];
                        r2 = fma(df, df, r2);
                    }
                    double r = sqrt(fmax(r2, 1e-30));
                    double kk = p.kappa * (1.0 + kSqrt5 * r + (5.0 / 3.0) * r * r) * exp(-kSqrt5 * r);
                    double v = kk - gsum[j][k] + ((j == k) ? p.noise : 0.0);
                    Sm[j][k] = v; Sm[k][j] = v;
                }
            bool ok = chol_lower<Q>(Sm, Ls);
            inv_from_chol<Q>(Ls, Si);
            // grid covariance: rows scaled by the refinement factor, Cholesky of the lower triangle
            double Cov[Q][Q], Lc[Q][Q];
            for (int j = 0; j < Q; ++j) {
                double rs = p.cscale ? p.cscale[r0 + j] : 1.0;
                for (int k = 0; k < Q; ++k) Cov[j][k] = rs * Sm[j][k];
            }
            ok = chol_lower<Q>(Cov, Lc) && ok;
            double logdet = 0.0;
            for (int j = 0; j < Q; ++j) logdet += 2.0 * log(Lc[j][j]);
            // refit posterior covariance at the candidates: noise (I - noise S^-1), symmetrised
            double Cf[Q][Q], Lf[Q][Q];
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k)
                    Cf[j][k] = p.noise * (((j == k) ? 1.0 : 0.0) - p.noise * 0.5 * (Si[j][k] + Si[k][j]));
            ok = chol_lower<Q>(Cf, Lf) && ok;
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    s_S[j * Q + k] = Sm[j][k]; s_Sinv[j * Q + k] = Si[j][k];
                    s_Lc[j * Q + k] = Lc[j][k]; s_Lf[j * Q + k] = Lf[j][k];
                }
            for (int j = 0; j < Q; ++j) {
                double zc_c = 0.0;
                for (int s = 0; s < Sp; ++s) {
                    double cv = 0.0;
                    for (int k = 0; k < Q; ++k)
                        cv += Si[j][k] * (p.z_cand[k * Sp + s] - p.ut[int64_t(r0 + k) * Sp + s]);
                    s_C[j * MAX_SP + s] = cv;
                    zc_c = fma(p.z_cand[j * Sp + s], cv, zc_c);
                }
                double vc = p.noise * (1.0 + (double(Sp) - p.noise * zc_c) / (double(Sp) + 1e-6));
                s_misc[j] = vc;
                ok = ok && (vc > 0.0);
            }
            s_misc[Q] = logdet;
            s_misc[Q + 1] = ok ? 0.0 : 1.0;
        }
        __syncthreads();
        int bad = s_misc[Q + 1] != 0.0;

        // ---- phase A: per-train-point tables -------------------------------------------------------------------
        for (int i = tid; i < N; i += nthr) {
            uint32_t bits = p.zb[i];
            double w[Q], acc = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                w[j] = p.W[int64_t(r0 + j) * Np + i];
                double zdot = 0.0;
                for (int s = 0; s < Sp; ++s) zdot += ((bits >> s) & 1u) ? s_C[j * MAX_SP + s] : -s_C[j * MAX_SP + s];
                acc = fma(w[j], zdot, acc);
            }
            double var = p.noise * fma(kvar, acc, p.h0[i]);
            bad |= !(var > 0.0);
            double iv = 0.5 / var;
            double siv = sqrt(iv);
            s_lv[i] = -0.5 * log(var) - 0.5 * kLog2Pi;
            s_ap[i] = p.a[i] * siv;
#pragma unroll
            for (int j = 0; j < Q; ++j) s_wp[size_t(j) * Np + i] = p.noise * w[j] * siv;
        }
        bad = __syncthreads_or(bad);

        double mu[Q], drc[Q], norm[Q], omega[Q], ivc[Q], lvc[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            mu[j] = p.mu[r0 + j];
            drc[j] = p.drc ? p.drc[r0 + j] : 1.0;
            norm[j] = sum_dr + drc[j];
            omega[j] = drc[j] / norm[j];
            ivc[j] = 0.5 / s_misc[j];
            lvc[j] = -0.5 * log(s_misc[j]) - 0.5 * kLog2Pi;
        }
        const double logdet = s_misc[Q];

        // ---- phase B + C: warp per grid point --------------------------------------------------------------------
        for (int g = warp; g < G; g += nwarps) {
            double eps[Q], dev[Q], Bg[Q], thr[Q];
            double e2 = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                eps[j] = p.eps_grid[(p.eps_per_candidate ? int64_t(c) * G + g : g) * Q + j];
                e2 = fma(eps[j], eps[j], e2);
            }
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < Q; ++k)
                    if (k <= j) s = fma(s_Lc[j * Q + k], eps[k], s);
                dev[j] = s;
            }
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < Q; ++k) s = fma(s_Sinv[j * Q + k], dev[k], s);
                Bg[j] = s;
                double rc = p.noise * s;
                thr[j] = fma(-rc * rc, ivc[j], lvc[j]);
            }
            double sr[Q], sw[Q];
            int cnt[Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) { sr[j] = 0.0; sw[j] = 0.0; cnt[j] = 0; }
            if (!bad) {
                // two train points per lane and step (i, i + 32): 2 Q independent sigmoid chains instead of Q
                for (int i = lane; i < N; i += 64) {
                    const int i2 = i + 32;
                    const bool ok2 = i2 < N;
                    const int i2c = ok2 ? i2 : i;
                    double rp0 = s_ap[i], rp1 = s_ap[i2c];
#pragma unroll
                    for (int j = 0; j < Q; ++j) {
                        rp0 = fma(-s_wp[size_t(j) * Np + i], Bg[j], rp0);
                        rp1 = fma(-s_wp[size_t(j) * Np + i2c], Bg[j], rp1);
                    }
                    const double sci0 = fma(-rp0, rp0, s_lv[i]), sci1 = fma(-rp1, rp1, s_lv[i2c]);
                    const double dr0 = s_dr[i], dr1 = ok2 ? s_dr[i2c] : 0.0;
#pragma unroll
                    for (int j = 0; j < Q; ++j) {
                        const double dd0 = thr[j] - sci0, dd1 = thr[j] - sci1;
                        const bool ge0 = dd0 >= 0.0, ge1 = ok2 && (dd1 >= 0.0);
                        cnt[j] += (ge0 ? 1 : 0) + (ge1 ? 1 : 0);
                        if (SOFT) {
                            const double rho0 = soft_rho<true>(dd0 * c1q, tab_s, p.k_magic, p.k_c0, p.k_c1n, p.k_c2, p.k_c3n);
                            double rho1 = soft_rho<true>(dd1 * c1q, tab_s, p.k_magic, p.k_c0, p.k_c1n, p.k_c2, p.k_c3n);
                            rho1 = ok2 ? rho1 : 0.0;
                            sr[j] += rho0 + rho1;
                            sw[j] = fma(rho0, dr0, fma(rho1, dr1, sw[j]));
                        } else {
                            sw[j] += (ge0 ? dr0 : 0.0) + (ge1 ? dr1 : 0.0);
                        }
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                cnt[j] = warp_sum_int(cnt[j]);
                sw[j] = warp_sum(sw[j]);
                if (SOFT) sr[j] = warp_sum(sr[j]);
            }
            if (lane == 0) {           // raw sums only: the per-(grid point, column) epilogue runs thread-parallel below
                g_lp[g] = e2;
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    const int o = g * Q + j;
                    g_dev[o] = dev[j]; g_thr[o] = thr[j]; g_cnt[o] = double(cnt[j]); g_sr[o] = sr[j]; g_sw[o] = sw[j];
                }
            }
        }
        __syncthreads();
        // ---- phase C: thread per (grid point, column) ------------------------------------------------------------------------
        for (int o = tid; o < G * Q; o += nthr) {
            const int g = o / Q, j = o - g * Q;
            // rows of the q candidates themselves (weights zero except the own row, counted in n_g and sum rho)
            int n_g = int(g_cnt[o]);
            double srj = g_sr[o];
            const double thr_j = g_thr[o], sw_j = g_sw[o];
            for (int k = 0; k < Q; ++k) {
                double dd = thr_j - g_thr[g * Q + k];
                n_g += (dd >= 0.0) ? 1 : 0;
                if (SOFT) srj += sigmoid_acc(fmin(fmax(dd * inv_temp, -745.0), 745.0));
            }
            double drc_j = drc[0], norm_j = norm[0], omega_j = omega[0], mu_j = mu[0];
#pragma unroll
            for (int k = 1; k < Q; ++k)
                if (k == j) { drc_j = drc[k]; norm_j = norm[k]; omega_j = omega[k]; mu_j = mu[k]; }
            double M, rho_self;
            if (SOFT) {
                double scale = double(n_g) / srj;
                rho_self = 0.5 * scale;
                M = scale * (sw_j + 0.5 * drc_j) / norm_j;
            } else {
                rho_self = 1.0;
                M = (sw_j + drc_j) / norm_j;
            }
            if (p.randomized) {
                double D = (j == Q - 1) ? fmax(rho_self * omega_j, 1e-6) : 1e-6;     // helpers.py:307: last row only
                double pr = fmin(fmax((M - p.alpha_lvl) / D, 0.0), 1.0);
                double u = p.u_rand[(int64_t(c) * G + g) * Q + j];
                M -= (u < pr) ? 0.0 : D;
            }
            double mk = SOFT ? sigmoid_acc((M - p.alpha_lvl) * inv_temp) : ((M > p.alpha_lvl) ? 1.0 : 0.0);
            const double dv = g_dev[o];
            double yg = mu_j + dv;
            if (bad) mk = nan("");
            g_mask[o] = mk;
            g_opt[o] = p.optimistic ? ((yg >= mu_j) ? 1.0 : 0.0) : 1.0;
            if (p.grid) p.grid[(int64_t(c) * G + g) * Q + j] = yg;
            if (p.mask) p.mask[(int64_t(c) * G + g) * Q + j] = mk;
            if (p.qscore) p.qscore[(int64_t(c) * G + g) * Q + j] = bad ? nan("") : thr_j;
        }
        for (int g = tid; g < G; g += nthr) {       // g_lp is not read by the loop above
            double lp = -0.5 * (g_lp[g] + logdet + double(Q) * kLog2Pi);
            g_lp[g] = bad ? nan("") : lp;
            if (p.logp) p.logp[int64_t(c) * G + g] = bad ? nan("") : lp;
        }
        __syncthreads();

        // ---- phase D: integration per column, max over q, mean over the MC samples ---------------------------------
        double ood[Q], scw[Q], snc[Q], nacc[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            ood[j] = SOFT ? sigmoid_acc((omega[j] - p.alpha_lvl) * inv_temp) : ((omega[j] >= p.alpha_lvl) ? 1.0 : 0.0);
            double a1 = 0.0, a2 = 0.0, a3 = 0.0;
            for (int g = tid; g < G; g += nthr) {
                double mk = g_mask[g * Q + j], op = g_opt[g * Q + j];
                a1 += mk * op / fmax(exp(g_lp[g]), 1e-6);
                a2 += (1.0 - mk) * op;
                a3 += (mk >= 0.5) ? 1.0 : 0.0;
            }
            scw[j] = block_sum(a1, red, tid, nwarps);
            snc[j] = block_sum(a2, red, tid, nwarps);
            nacc[j] = block_sum(a3, red, tid, nwarps);
        }
        double tot = 0.0;
        for (int s = 0; s < S; ++s) {
            double best = -INFINITY;
            bool isnan_any = false;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                double A = 0.0, B = 0.0;
                for (int g = tid; g < G; g += nthr) {
                    // refit mean at candidate j: mu_j + sum_k (I - noise S^-1)_jk dev_k ; sample f = mean + (L_f eps')_j
                    double mf = mu[j], fs = 0.0, ms = 0.0;
#pragma unroll
                    for (int k = 0; k < Q; ++k) {
                        double shr = ((j == k) ? 1.0 : 0.0) - p.noise * s_Sinv[j * Q + k];
                        mf = fma(shr, g_dev[g * Q + k], mf);
                        if (k <= j) {
                            fs = fma(s_Lf[j * Q + k], p.eps_mc[(int64_t(s) * G + g) * Q + k], fs);
                            if (p.kind == 0 && S > 1) {
                                double em = 0.0;
                                for (int s2 = 0; s2 < S; ++s2) em += p.eps_mc[(int64_t(s2) * G + g) * Q + k];
                                ms = fma(s_Lf[j * Q + k], em / double(S), ms);
                            }
                        }
                    }
                    double f = mf + fs, val;
                    if (p.ext_values) {               // cNEI: joint-sample values from k_nei_values_q
                        val = p.ext_values[((int64_t(c) * S + s) * G + g) * Q + j];
                    } else if (p.kind == 0) {
                        if (S > 1) { double m = mf + ms; val = m + p.beta_prime * fabs(f - m); }
                        else val = f;
                    } else {
                        double xx = f - p.best_f;
                        if (p.temp > 0.0) { double xt = xx / p.temp; val = p.temp * (fmax(xt, 0.0) + log1p(exp(-fabs(xt)))); }
                        else val = fmax(xx, 0.0);
                    }
                    double mk = g_mask[g * Q + j], op = g_opt[g * Q + j];
                    A = fma(mk * op / fmax(exp(g_lp[g]), 1e-6), val, A);
                    B = fma((1.0 - mk) * op, val, B);
                }
                A = block_sum(A, red, tid, nwarps);
                B = block_sum(B, red, tid, nwarps);
                double r = (1.0 - ood[j]) * ((1.0 - p.alpha_lvl) * A / fmax(scw[j], 1e-6) + p.alpha_lvl * B / fmax(snc[j], 1e-6));
                isnan_any = isnan_any || (r != r);
                best = fmax(best, r);
            }
            tot += isnan_any ? nan("") : best;
        }
        if (tid == 0) {
            p.acq[c] = tot / double(S);
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                if (p.ood) p.ood[r0 + j] = ood[j];
                if (p.post_mean) p.post_mean[r0 + j] = mu[j];
                if (p.num_acc) p.num_acc[r0 + j] = nacc[j];
                if (p.post_var) p.post_var[r0 + j] = s_S[j * Q + j];
                if (pq.post_cov)
                    for (int k = 0; k < Q; ++k) pq.post_cov[(int64_t(c) * Q + j) * Q + k] = s_S[j * Q + k];
            }
        }
        __syncthreads();
    }
}

}  // namespace cbo
