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
// One warp per joint batch, a lane owns rows lane and lane + 32 of the (nb+Q)^2 factor; the Cholesky reverse  Sigma_bar = L^-T Phi(L^T L_bar) L^-1
// is two triangular solves with one column / one row per lane.  oracle: torch.autograd through oracle/closed_form.py.
#pragma once
#include "nei.cuh"
#include "score_q_bwd.cuh"

namespace cbo {

struct NeiBwdArgs {
    NeiArgs f;
    const double* val_bar;     // n_cand x S x G x Q   d acq / d values (raw = 1: d L / d f(x) samples)
    double* adjq;              // rows x ADJQ_STRIDE   (+=)
    double* wbar;              // rows x Np            (+=)
    double* cv_bar;            // n_cand x nb x Q      (out)
    int grid_leak;             // cbo_params.grid_grad_leak: add this kernel's part of the adjoint of the prior block k(x_j, x_k)
                               // that flows through the target grid (score_q_bwd.cuh, grid_leak_from_moments) to adjq[ADJQ_LEAK..]
};

__host__ __device__ inline size_t nei_bwd_doubles(int nb, int q) { return 3 * size_t(nb + q) * nei_ld(nb + q) + 3 * size_t(nb + q) * q; }

template <int Q>
__global__ void __launch_bounds__(32)
k_nei_values_q_bwd(const NeiBwdArgs qb) {
    const NeiArgs& p = qb.f;
    extern __shared__ double nei_sm[];
    const int lane = threadIdx.x;
    const int c = blockIdx.x;
    if (c >= p.n_cand) return;
    const int nb = p.nb, m = nb + Q, r0 = c * Q, ld = nei_ld(m);
    const double noise = p.noise;
    double* L = nei_sm;
    double* Lb = L + size_t(m) * ld;
    double* Y = Lb + size_t(m) * ld;
    double* cv = Y + size_t(m) * ld;
    double* cs = cv + size_t(m) * Q;
    double* cvb_s = cs + size_t(m) * Q;
    // ---- forward set-up (as k_nei_values_q) -------------------------------------------------------------------------
    double Sm[Q][Q], Si[Q][Q], Lc[Q][Q], Sis[Q][Q];
    const bool ok = nei_setup<Q>(p, c, lane, L, cv, cs, Sm, Si, Lc);
#pragma unroll
    for (int j = 0; j < Q; ++j)
#pragma unroll
        for (int k = 0; k < Q; ++k) Sis[j][k] = 0.5 * (Si[j][k] + Si[k][j]);
    for (int r = lane; r < m; r += 32)
        for (int k = 0; k < m; ++k) { Lb[r * ld + k] = 0.0; Y[r * ld + k] = 0.0; }
    double mu[Q];
#pragma unroll
    for (int j = 0; j < Q; ++j) mu[j] = p.mu[r0 + j];
    const double inv_temp = p.temp > 0.0 ? 1.0 / p.temp : 0.0;
    __syncwarp();

    // ---- reverse sweep over (MC sample, grid point): all lanes in lockstep, a lane owns rows lane (slot 0) and lane + 32 (slot 1) ----
    double cvb[2][Q], cvr[2][Q], mubr[2], mub[Q], Sib[Q][Q], Fx[Q][Q];    // Fx = sum f_x_bar dev^T (grid leak: f_x = y_g - noise B_g)
#pragma unroll
    for (int j = 0; j < Q; ++j) {
        mub[j] = 0.0;
#pragma unroll
        for (int k = 0; k < Q; ++k) Sib[j][k] = Fx[j][k] = 0.0;
    }
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int r = lane + 32 * t;
        mubr[t] = (r < nb) ? p.mub[r] : 0.0;
#pragma unroll
        for (int j = 0; j < Q; ++j) { cvb[t][j] = 0.0; cvr[t][j] = (r < nb) ? cv[r * Q + j] : 0.0; }
    }
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
        double f[2];
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int r = lane + 32 * t;
            f[t] = 0.0;
            if (r < m) {
                if (r < nb) {
                    f[t] = mubr[t];
#pragma unroll
                    for (int j = 0; j < Q; ++j) f[t] = fma(cvr[t][j], Bg[j], f[t]);
                } else {
#pragma unroll
                    for (int j = 0; j < Q; ++j)
                        if (j == r - nb) f[t] = mu[j] + dev[j] - noise * Bg[j];
                }
                for (int k = 0; k <= r; ++k) f[t] = fma(L[r * ld + k], e[k], f[t]);
            }
        }
        double vb[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) vb[j] = qb.val_bar[((int64_t(c) * p.S + smp) * p.G + g) * Q + j];
        double fbar[2] = {0.0, 0.0}, fxb[Q];           // fxb[j]: adjoint of the sample at candidate point j (every lane)
        if (p.raw) {
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                fxb[j] = vb[j];
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (lane + 32 * t == nb + j) fbar[t] = vb[j];
            }
        } else {
            double mx = fmax((lane < nb) ? f[0] : -INFINITY, (lane + 32 < nb) ? f[1] : -INFINITY);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            double sm_r[2], smax;
            if (p.temp > 0.0) {
                const double ex0 = (lane < nb) ? exp((f[0] - mx) * inv_temp) : 0.0;
                const double ex1 = (lane + 32 < nb) ? exp((f[1] - mx) * inv_temp) : 0.0;
                const double se = warp_sum(ex0 + ex1);
                smax = mx + p.temp * log(se);
                sm_r[0] = ex0 / se; sm_r[1] = ex1 / se;
            } else {                                  // hard maximum: its adjoint goes to the first arg-max
                const unsigned hit0 = __ballot_sync(0xffffffffu, lane < nb && f[0] == mx);
                const unsigned hit1 = __ballot_sync(0xffffffffu, lane + 32 < nb && f[1] == mx);
                sm_r[0] = (hit0 && lane == __ffs(hit0) - 1) ? 1.0 : 0.0;
                sm_r[1] = (!hit0 && hit1 && lane == __ffs(hit1) - 1) ? 1.0 : 0.0;
                smax = mx;
            }
            double tsum = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                const int rx = nb + j;
                const double fx = __shfl_sync(0xffffffffu, (rx >= 32) ? f[1] : f[0], rx & 31);
                double dv;
                if (p.temp > 0.0) dv = sigmoid_acc((fx - smax) * inv_temp);
                else dv = (fx - smax > 0.0) ? 1.0 : 0.0;
                fxb[j] = vb[j] * dv;
                tsum += fxb[j];
#pragma unroll
                for (int t = 0; t < 2; ++t)
                    if (lane + 32 * t == rx) fbar[t] = fxb[j];
            }
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (lane + 32 * t < nb) fbar[t] = -tsum * sm_r[t];
        }
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int r = lane + 32 * t;
            if (r < m)
                for (int k = 0; k <= r; ++k) Lb[r * ld + k] = fma(fbar[t], e[k], Lb[r * ld + k]);
        }
        double Bgb[Q];
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            double part = 0.0;
#pragma unroll
            for (int t = 0; t < 2; ++t)
                if (lane + 32 * t < nb) { cvb[t][j] = fma(fbar[t], Bg[j], cvb[t][j]); part = fma(fbar[t], cvr[t][j], part); }
            Bgb[j] = warp_sum(part) - noise * fxb[j];
        }
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            double db = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                db = fma(Si[j][k], Bgb[j], db);
                Sib[j][k] = fma(Bgb[j], dev[k], Sib[j][k]);
                Fx[j][k] = fma(fxb[j], dev[k], Fx[j][k]);
            }
            mub[k] -= db;                            // dev = y_g - mu with the grid point held fixed
        }
    }
    __syncwarp();
    // grid leak: dev_bar_g = f_x_bar_g + Sinv^T B_bar_g, so sum_g dev_bar_g dev_g^T = Fx + Sinv^T (sum_g B_bar_g dev_g^T)
    double leak[Q][Q];
#pragma unroll
    for (int j = 0; j < Q; ++j)
#pragma unroll
        for (int k = 0; k < Q; ++k) leak[j][k] = 0.0;
    if (Q > 1 && qb.grid_leak && p.cscale == nullptr) {
        double Dm[Q][Q];
#pragma unroll
        for (int a = 0; a < Q; ++a)
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                double s2 = Fx[a][k];
#pragma unroll
                for (int j = 0; j < Q; ++j) s2 = fma(Si[j][a], Sib[j][k], s2);
                Dm[a][k] = s2;
            }
        grid_leak_from_moments<Q>(Dm, Lc, leak);
    }

    // ---- Cholesky reverse: Sigma_bar = L^-T Phi(L^T L_bar) L^-1 ---------------------------------------------------------------
    for (int r = lane; r < m; r += 32)                // row r of Phi(L^T L_bar) -> Y
        for (int k = 0; k <= r; ++k) {
            double s = 0.0;
            for (int i = r; i < m; ++i) s = fma(L[i * ld + r], Lb[i * ld + k], s);
            Y[r * ld + k] = (k == r) ? 0.5 * s : s;
        }
    __syncwarp();
    for (int col = lane; col < m; col += 32)          // column col: L^T y = phi  (back substitution, in place)
        for (int i = m - 1; i >= 0; --i) {
            double s = Y[i * ld + col];
            for (int t = i + 1; t < m; ++t) s = fma(-L[t * ld + i], Y[t * ld + col], s);
            Y[i * ld + col] = s / L[i * ld + i];
        }
    __syncwarp();
    for (int r = lane; r < m; r += 32)                // row r: x L = y  (in place)
        for (int k = m - 1; k >= 0; --k) {
            double s = Y[r * ld + k];
            for (int t = k + 1; t < m; ++t) s = fma(-Y[r * ld + t], L[t * ld + k], s);
            Y[r * ld + k] = s / L[k * ld + k];
        }
    __syncwarp();
    for (int r = lane; r < m; r += 32)                // symmetrised adjoint M = (Y + Y^T)/2 into Lb
        for (int k = 0; k < m; ++k) Lb[r * ld + k] = 0.5 * (Y[r * ld + k] + Y[k * ld + r]);
    __syncwarp();
    const double* M = Lb;

    // ---- Sigma_bar -> c_bar, S^-1_bar ------------------------------------------------------------------------------------
    double csb[2][Q];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int r = lane + 32 * t;
#pragma unroll
        for (int j = 0; j < Q; ++j) {
            csb[t][j] = 0.0;
            if (r < nb) {
                double a = 0.0, b2 = 0.0;
                for (int k = 0; k < nb; ++k) {
                    a = fma(M[r * ld + k], cv[k * Q + j], a);
                    b2 = fma(M[r * ld + k], cs[k * Q + j], b2);
                }
                csb[t][j] = -a + 2.0 * noise * M[(nb + j) * ld + r];
                cvb[t][j] -= b2;
            }
        }
    }
    double Sisb[Q][Q];
#pragma unroll
    for (int k = 0; k < Q; ++k)
#pragma unroll
        for (int j = 0; j < Q; ++j)
            Sisb[k][j] = warp_sum(cvr[0][k] * csb[0][j] + cvr[1][k] * csb[1][j]) - noise * noise * M[(nb + k) * ld + nb + j];
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const int r = lane + 32 * t;
        if (r < nb) {
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                double a = cvb[t][k];
#pragma unroll
                for (int j = 0; j < Q; ++j) a = fma(csb[t][j], Sis[k][j], a);
                cvb_s[r * Q + k] = a;
                qb.cv_bar[(int64_t(c) * nb + r) * Q + k] = ok ? a : nan("");
            }
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
#pragma unroll
            for (int k = 0; k < Q; ++k)
                if (k != j && ok) adj[ADJQ_LEAK + k] += leak[j][k] + leak[k][j];
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

template <int Q>
inline cudaError_t launch_nei_values_bwd_t(const NeiBwdArgs& a, cudaStream_t s) {
    const size_t smem = sizeof(double) * nei_bwd_doubles(a.f.nb, Q);
    static SmemOptIn optin;
    if (cudaError_t e = optin.ensure(k_nei_values_q_bwd<Q>, smem); e != cudaSuccess) return e;
    CBO_LAUNCHED(), k_nei_values_q_bwd<Q><<<a.f.n_cand, 32, smem, s>>>(a);
    return cudaGetLastError();
}

inline cudaError_t launch_nei_values_q_bwd(const NeiBwdArgs& a, int q, cudaStream_t s) {
    switch (q) {
        case 1: return launch_nei_values_bwd_t<1>(a, s);
        case 2: return launch_nei_values_bwd_t<2>(a, s);
        case 3: return launch_nei_values_bwd_t<3>(a, s);
        case 4: return launch_nei_values_bwd_t<4>(a, s);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace cbo
