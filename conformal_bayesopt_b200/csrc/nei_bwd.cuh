// nei_bwd.cuh -- reverse pass of k_nei_values_q: the joint posterior of the refit models over [X_baseline; X].
//
// cNEI / cNEHVI under autograd (conformalbo/optim/optimize.py:224-235 reaches every acquisition that carries a ratio
// estimator, conformalbo/utils.py:243-246): k_score_q_bwd hands back d acq / d values_sgj (val_bar); this kernel carries it
// through  values <- f = mean + L e,  L = chol(Sigma),  Sigma = [[Sigma_bb - c S^-1 c^T, noise c S^-1], [., noise (I - noise S^-1)]],
// mean = [mu_b + c B_g ; y_g - noise B_g],  B_g = S^-1 (y_g - mu)   (the grid y_g is a constant, helpers.py:118-119)
// down to the three things the rest of the reverse pass understands:
//   mu_bar_j          added to the adjoint record of row j   (adjq[0])
//   S_bar + S_bar^T   added to the record                    (adjq[2..])
//   c_bar (nb x Q)    c_rj = k(X_b,r, x_j) - K_bX[r,:] w_j:  w_bar'_j -= K_bX^T c_bar_j (right-hand side of the A^-1 solve),
//                     and the direct term sum_r c_bar_rj dk(X_b,r, x_j)/dx_j, added by k_rows_bwd_q from cv_bar.
// One warp per joint batch, lane = row of the (nb+Q)^2 factor; the Cholesky reverse  Sigma_bar = L^-T Phi(L^T L_bar) L^-1
// is two triangular solves with one column / one row per lane.  oracle: torch.autograd through oracle/closed_form.py.
#pragma once
#include "nei.cuh"
#include "score_q_bwd.cuh"

namespace cbo {

constexpr int NEI_LD = 33;                     // padded row stride of the 32 x 32 work matrices (conflict-free rows AND columns)

struct NeiBwdArgs {
    NeiArgs f;
    const double* val_bar;     // n_cand x S x G x Q   d acq / d values (raw = 1: d L / d f(x) samples)
    double* adjq;              // rows x ADJQ_STRIDE   (+=)
    double* wbar;              // rows x Np            (+=)
    double* cv_bar;            // n_cand x nb x Q      (out)
};

template <int Q>
__global__ void __launch_bounds__(32)
k_nei_values_q_bwd(const NeiBwdArgs qb) {
    const NeiArgs& p = qb.f;
    __shared__ double sL[32 * NEI_LD], sLb[32 * NEI_LD], sY[32 * NEI_LD];
    __shared__ double cv[32 * Q], cs[32 * Q], cvb_s[32 * Q];
    const int lane = threadIdx.x;
    const int c = blockIdx.x;
    if (c >= p.n_cand) return;
    const int nb = p.nb, m = nb + Q, r0 = c * Q;
    const double noise = p.noise;
    // ---- forward set-up (as k_nei_values_q) -------------------------------------------------------------------------
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
            const double v = kk - acc + ((j == k) ? noise : 0.0);
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
    double Sis[Q][Q];
#pragma unroll
    for (int j = 0; j < Q; ++j)
#pragma unroll
        for (int k = 0; k < Q; ++k) Sis[j][k] = 0.5 * (Si[j][k] + Si[k][j]);
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
            for (int k = 0; k < Q; ++k) a = fma(cv[lane * Q + k], Sis[k][j], a);
            cs[lane * Q + j] = a;
        }
    }
    __syncwarp();
    double* L = sL;
    double* Lb = sLb;
    double* Y = sY;
    for (int k = 0; k < 32; ++k) { L[lane * NEI_LD + k] = 0.0; Lb[lane * NEI_LD + k] = 0.0; Y[lane * NEI_LD + k] = 0.0; }
    if (lane < m) {
        for (int k = 0; k <= lane; ++k) {
            double v;
            if (lane < nb) {
                v = p.Sbb[lane * nb + k];
#pragma unroll
                for (int j = 0; j < Q; ++j) v -= cs[lane * Q + j] * cv[k * Q + j];
            } else {
                const int j = lane - nb;
                if (k < nb) v = noise * cs[k * Q + j];
                else {
                    const int j2 = k - nb;
                    double sij = 0.0;
#pragma unroll
                    for (int a = 0; a < Q; ++a)
#pragma unroll
                        for (int b2 = 0; b2 < Q; ++b2)
                            if (a == j && b2 == j2) sij = Sis[a][b2];
                    v = noise * (((j == j2) ? 1.0 : 0.0) - noise * sij);
                }
            }
            L[lane * NEI_LD + k] = v;
        }
    }
    __syncwarp();
    for (int j = 0; j < m; ++j) {
        const double djj = L[j * NEI_LD + j];
        ok = ok && (djj > 0.0);
        const double ljj = sqrt(djj);
        __syncwarp();
        if (lane == j) L[j * NEI_LD + j] = ljj;
        if (lane > j && lane < m) L[lane * NEI_LD + j] /= ljj;
        __syncwarp();
        if (lane > j && lane < m) {
            const double lij = L[lane * NEI_LD + j];
            for (int k = j + 1; k <= lane; ++k) L[lane * NEI_LD + k] -= lij * L[k * NEI_LD + j];
        }
        __syncwarp();
    }
    double mu[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) mu[j] = p.mu[r0 + j];
    const double inv_temp = p.temp > 0.0 ? 1.0 / p.temp : 0.0;

    // ---- reverse sweep over (MC sample, grid point): all lanes in lockstep, lane = row of the joint sample ----------------
    double cvb[Q], mub[Q], Sib[Q][Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        cvb[j] = 0.0; mub[j] = 0.0;
#pragma unroll
        for (int k = 0; k < Q; ++k) Sib[j][k] = 0.0;
    }
    double cvr[Q];                                   // this lane's row of c
#pragma unroll
    for (int j = 0; j < Q; ++j) cvr[j] = (lane < nb) ? cv[lane * Q + j] : 0.0;
    const double mub_r = (lane < nb) ? p.mub[lane] : 0.0;
    for (int sg = 0; sg < p.S * p.G; ++sg) {
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
        double f = 0.0;
        if (lane < m) {
            if (lane < nb) {
                f = mub_r;
#pragma unroll
                for (int j = 0; j < Q; ++j) f = fma(cvr[j], Bg[j], f);
            } else {
#pragma unroll
                for (int j = 0; j < Q; ++j)
                    if (j == lane - nb) f = mu[j] + dev[j] - noise * Bg[j];
            }
            for (int k = 0; k <= lane; ++k) f = fma(L[lane * NEI_LD + k], e[k], f);
        }
        double vb[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) vb[j] = qb.val_bar[((int64_t(c) * p.S + smp) * p.G + g) * Q + j];
        double fbar = 0.0, fxb[Q];                   // fxb[j]: adjoint of the sample at candidate point j (every lane)
        if (p.raw) {
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                fxb[j] = vb[j];
                if (lane == nb + j) fbar = vb[j];
            }
        } else {
            double fb = (lane < nb) ? f : -INFINITY;
            double mx = fb;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            double sm_r, smax;
            if (p.temp > 0.0) {
                const double ex = (lane < nb) ? exp((f - mx) * inv_temp) : 0.0;
                const double se = warp_sum(ex);
                smax = mx + p.temp * log(se);
                sm_r = ex / se;
            } else {                                  // hard maximum: its adjoint goes to the first arg-max
                const unsigned hit = __ballot_sync(0xffffffffu, lane < nb && f == mx);
                sm_r = (lane == __ffs(hit) - 1) ? 1.0 : 0.0;
                smax = mx;
            }
            double tsum = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                const double fx = __shfl_sync(0xffffffffu, f, nb + j);
                double dv;
                if (p.temp > 0.0) dv = sigmoid_acc((fx - smax) * inv_temp);
                else dv = (fx - smax > 0.0) ? 1.0 : 0.0;
                fxb[j] = vb[j] * dv;
                tsum += fxb[j];
                if (lane == nb + j) fbar = fxb[j];
            }
            if (lane < nb) fbar = -tsum * sm_r;
        }
        if (lane < m)
            for (int k = 0; k <= lane; ++k) Lb[lane * NEI_LD + k] = fma(fbar, e[k], Lb[lane * NEI_LD + k]);
        double Bgb[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            if (lane < nb) cvb[j] = fma(fbar, Bg[j], cvb[j]);
            Bgb[j] = warp_sum((lane < nb) ? fbar * cvr[j] : 0.0) - noise * fxb[j];
        }
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            double db = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                db = fma(Si[j][k], Bgb[j], db);
                Sib[j][k] = fma(Bgb[j], dev[k], Sib[j][k]);
            }
            mub[k] -= db;                            // dev = y_g - mu with the grid point held fixed
        }
    }
    __syncwarp();

    // ---- Cholesky reverse: Sigma_bar = L^-T Phi(L^T L_bar) L^-1 ---------------------------------------------------------------
    if (lane < m) {                                   // row `lane` of Phi(L^T L_bar) -> Y
        for (int k = 0; k <= lane; ++k) {
            double s = 0.0;
            for (int i = lane; i < m; ++i) s = fma(L[i * NEI_LD + lane], Lb[i * NEI_LD + k], s);
            Y[lane * NEI_LD + k] = (k == lane) ? 0.5 * s : s;
        }
    }
    __syncwarp();
    if (lane < m) {                                   // column `lane`: L^T y = phi  (back substitution, in place)
        for (int i = m - 1; i >= 0; --i) {
            double s = Y[i * NEI_LD + lane];
            for (int t = i + 1; t < m; ++t) s = fma(-L[t * NEI_LD + i], Y[t * NEI_LD + lane], s);
            Y[i * NEI_LD + lane] = s / L[i * NEI_LD + i];
        }
    }
    __syncwarp();
    if (lane < m) {                                   // row `lane`: x L = y  (in place)
        for (int k = m - 1; k >= 0; --k) {
            double s = Y[lane * NEI_LD + k];
            for (int t = k + 1; t < m; ++t) s = fma(-Y[lane * NEI_LD + t], L[t * NEI_LD + k], s);
            Y[lane * NEI_LD + k] = s / L[k * NEI_LD + k];
        }
    }
    __syncwarp();
    // symmetrised adjoint M = (Y + Y^T)/2 into Lb
    if (lane < m)
        for (int k = 0; k < m; ++k) Lb[lane * NEI_LD + k] = 0.5 * (Y[lane * NEI_LD + k] + Y[k * NEI_LD + lane]);
    __syncwarp();
    const double* M = Lb;

    // ---- Sigma_bar -> c_bar, S^-1_bar ------------------------------------------------------------------------------------
    double csb[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) csb[j] = 0.0;
    if (lane < nb) {
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double a = 0.0, b2 = 0.0;
            for (int k = 0; k < nb; ++k) {
                a = fma(M[lane * NEI_LD + k], cv[k * Q + j], a);
                b2 = fma(M[lane * NEI_LD + k], cs[k * Q + j], b2);
            }
            csb[j] = -a + 2.0 * noise * M[(nb + j) * NEI_LD + lane];
            cvb[j] -= b2;
        }
    }
    double Sisb[Q][Q];
#pragma unroll
    for (int k = 0; k < Q; ++k)
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            Sisb[k][j] = warp_sum((lane < nb) ? cvr[k] * csb[j] : 0.0) - noise * noise * M[(nb + k) * NEI_LD + nb + j];
        }
    if (lane < nb) {
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            double a = cvb[k];
#pragma unroll
            for (int j = 0; j < Q; ++j) a = fma(csb[j], Sis[k][j], a);
            cvb_s[lane * Q + k] = a;
            qb.cv_bar[(int64_t(c) * nb + lane) * Q + k] = ok ? a : nan("");
        }
    }
    __syncwarp();
    // S_bar = -S^-T (Sinv_bar) S^-T with Sinv_bar = direct part + symmetrised part
    double Sb[Q][Q], Tm[Q][Q];
#pragma unroll
    for (int j = 0; j < Q; ++j)
#pragma unroll
        for (int k = 0; k < Q; ++k) Sib[j][k] += 0.5 * (Sisb[j][k] + Sisb[k][j]);
#pragma unroll
    for (int j = 0; j < Q; ++j)
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < Q; ++t) s = fma(Sib[j][t], Si[k][t], s);
            Tm[j][k] = s;
        }
#pragma unroll
    for (int j = 0; j < Q; ++j)
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            double s = 0.0;
#pragma unroll
            for (int t = 0; t < Q; ++t) s = fma(Si[t][j], Tm[t][k], s);
            Sb[j][k] = -s;
        }
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double* adj = qb.adjq + int64_t(r0 + j) * ADJQ_STRIDE;
            adj[0] += ok ? mub[j] : nan("");
#pragma unroll
            for (int k = 0; k < Q; ++k) adj[2 + k] += Sb[j][k] + Sb[k][j];
        }
    }
    // right-hand sides of the A^-1 solve: c_rj = ... - K_bX[r,:] w_j
    for (int i = lane; i < p.N; i += 32) {
        double acc[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) acc[j] = 0.0;
        for (int r = 0; r < nb; ++r) {
            const double kb = p.KbX[int64_t(r) * p.Np + i];
#pragma unroll
            for (int j = 0; j < Q; ++j) acc[j] = fma(kb, cvb_s[r * Q + j], acc[j]);
        }
#pragma unroll
        for (int j = 0; j < Q; ++j) qb.wbar[int64_t(r0 + j) * p.Np + i] -= acc[j];
    }
}

inline cudaError_t launch_nei_values_q_bwd(const NeiBwdArgs& a, int q, cudaStream_t s) {
    switch (q) {
        case 1: CBO_LAUNCHED(), k_nei_values_q_bwd<1><<<a.f.n_cand, 32, 0, s>>>(a); break;
        case 2: CBO_LAUNCHED(), k_nei_values_q_bwd<2><<<a.f.n_cand, 32, 0, s>>>(a); break;
        case 3: CBO_LAUNCHED(), k_nei_values_q_bwd<3><<<a.f.n_cand, 32, 0, s>>>(a); break;
        case 4: CBO_LAUNCHED(), k_nei_values_q_bwd<4><<<a.f.n_cand, 32, 0, s>>>(a); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace cbo
