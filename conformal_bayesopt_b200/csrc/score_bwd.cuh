// score_bwd.cuh -- value AND reverse pass of the q = 1 score / rank / mask / integration stage (cUCB, cEI).
//
// What the reference obtains with autograd through conformal_gp_regression + _conformal_integration
// (conformalbo/optim/optimize.py:224-235: `acq_function(X).sum().backward()` inside optimize_acqf_sgld; the grid, its
// log-density and the optimistic mask are no_grad constants, helpers.py:118-119, acquisitions.py:58-59) is evaluated
// here from the closed forms, per candidate, in one kernel:
//   phases A-D  as score.cuh (the same row sweep `score_rows`), keeping the per-grid sums
//   phase  E    (thread per grid point) d acq / d mask_g, d acq / d values_g -> P_g = d acq/d sum_i rho_ig dr_i and
//               Q_g = d acq/d sum_i rho_ig, so that d acq / d rho_ig = P_g dr_i + Q_g
//   pass 2      second sweep over (train point, grid point): Gamma_ig = (P_g dr_i + Q_g) rho_ig (1 - rho_ig) / temp,
//               E_k,i = sum_g Gamma_ig b_g^k, k = 0,1,2.  Because the score difference is the quadratic
//               Delta_ig = A_i b_g^2 + B_i b_g + C_i, the three moments carry everything: d acq/d(A_i,B_i,C_i) =
//               (E_2, E_1, E_0) and sum_g b_bar_g b_g^m = sum_i (2 A_i E_{m+1} + B_i E_m).  A thread owns 2 train points
//               (registers) and loops over the grid (per-grid record broadcast from shared memory): no atomics, no
//               cross-lane traffic in the loop; 15 fp64-pipe instructions per (train point, grid point).
//   epilogue    per train point: w_bar_i (written as one row of `wbar`, the right-hand side of the A^-1 solve done by
//               the same two triangular GEMMs as the forward pass); per candidate: mu_bar, s_bar, drc_bar, ut_bar.
// oracle/analytic_grad.py is the line-by-line CPU restatement of this algebra (checked against torch.autograd).
#pragma once
#include "score.cuh"

namespace cbo {

constexpr int SB_RI = 2;              // train points per thread in pass 2 ...
constexpr int SB_GI = 2;              // ... times grid points per step = independent dependency chains per thread
constexpr int SB_GARR = 12;           // per-grid arrays
constexpr int ADJ_STRIDE = 4 + MAX_SP;   // adj[c] = {mu_bar, s_bar, drc_bar, -, ut_bar[Sp]}

struct ScoreBwdArgs {
    ScoreArgs f;
    double* wbar;                     // n_cand x Np   d acq / d w  (zero beyond N)
    double* adj;                      // n_cand x ADJ_STRIDE
};

__host__ __device__ inline size_t score_bwd_group_doubles(int Np, int G) {
    return 3 * size_t(Np) + size_t(SB_GARR) * G + 64 + 16 * 32 * 2 + 32 + 96;
}
__host__ __device__ inline size_t score_bwd_smem_bytes(int Np, int G, int ngroups) {
    return score_smem_shared_bytes(Np) + size_t(ngroups) * score_bwd_group_doubles(Np, G) * sizeof(double);
}
inline int score_bwd_groups_for(int Np, int G, int wpg) {
    size_t avail = 227 * 1024, one = score_bwd_group_doubles(Np, G) * sizeof(double);
    if (score_smem_shared_bytes(Np) + one > avail) return 0;
    size_t w = (avail - score_smem_shared_bytes(Np)) / one;
    size_t cap = size_t(2 * SCORE_MAX_PAIRS / wpg);
    return int(w > cap ? cap : w);
}

// rho = sigmoid(x ln2 / 16): the instruction sequence of score_rows, with rho materialised (12 fp64-pipe instructions)
template <bool CLAMP>
__device__ __forceinline__ double soft_rho(double x, unsigned tab_s, double kMagic, double kC0, double kC1n,
                                           double kC2, double kC3n) {
    constexpr int kSatHi = 0x408d9000;
    double xc = x;
    if (CLAMP) {
        int hi = __double2hiint(x);
        int ch = min(hi & 0x7fffffff, kSatHi) | (hi & 0x80000000);
        xc = __hiloint2double(ch, __double2loint(x));
    }
    double t = kMagic - xc;
    double yn = xc + (t - kMagic);
    double pl = fma(yn, kC3n, kC2);
    pl = fma(pl, yn, kC1n);
    pl = fma(pl, yn, kC0);
    double dd = fma(exp_tab_scaled(tab_s, __double2loint(t)), pl, 1.0), y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(dd));
    double err1 = fma(-dd, y0, 2.0);
    return y0 * err1;
}

// Pass 2 for SB_RI train points of one thread: E_k = sum_g Gamma_g b_g^k.
// rec[g] = UNIT ? {b, pq, pq b, pq b^2} : {b, b^2, P, Q}   (P, Q, pq already divided by temp)
// rho by the instruction sequence of score_rows_fx (qcm = M - qc: the tables of the soft path hold it that way).
template <bool CLAMP, bool UNIT>
__device__ __forceinline__ void moment_sweep(const ScoreArgs& p, const double* __restrict__ rec, int G,
                                             unsigned tabfx_s, const double (&qa)[SB_RI],
                                             const double (&qb)[SB_RI], const double (&qcm)[SB_RI], const double (&dri)[SB_RI],
                                             double (&E0)[SB_RI], double (&E1)[SB_RI], double (&E2)[SB_RI]) {
    const double kB2 = p.fx_b2, kB1 = p.fx_b1, kB0 = p.fx_b0;
    // Stage-wise over SB_RI rows x SB_GI grid points (as score_rows): written out so that the dependency chains are
    // interleaved instead of being scheduled one after the other; 2 x 2 keeps the live coefficients / moments at 12 doubles.
    static_assert(SB_GI == 2, "moment_sweep is written for two grid points per step");
#pragma unroll 1
    for (int g = 0; g < G; g += SB_GI) {
        const int g1 = (g + 1 < G) ? g + 1 : g;
        const double w1 = (g + 1 < G) ? 1.0 : 0.0;               // odd G: the padded second slot contributes nothing
        double2 r01[SB_GI], r23[SB_GI];
        r01[0] = *reinterpret_cast<const double2*>(rec + 4 * g);
        r23[0] = *reinterpret_cast<const double2*>(rec + 4 * g + 2);
        r01[1] = *reinterpret_cast<const double2*>(rec + 4 * g1);
        r23[1] = *reinterpret_cast<const double2*>(rec + 4 * g1 + 2);
        double t[SB_RI][SB_GI], dd[SB_RI][SB_GI];
#pragma unroll
        for (int r = 0; r < SB_RI; ++r)
#pragma unroll
            for (int k = 0; k < SB_GI; ++k) {
                const double bg = r01[k].x;
                double tt = fma(-fma(qa[r], bg, qb[r]), bg, qcm[r]);         // M - x
                if (CLAMP) {
                    constexpr int kLo = 0x4137fc4e, kHi = 0x413803b2;        // high words of M - 946, M + 946
                    const int h = min(max(__double2hiint(tt), kLo), kHi);
                    tt = __hiloint2double(h, __double2loint(tt));
                }
                t[r][k] = tt;
            }
#pragma unroll
        for (int r = 0; r < SB_RI; ++r)
#pragma unroll
            for (int k = 0; k < SB_GI; ++k) {
                const unsigned z = unsigned(__double2loint(t[r][k]));
                const int hi = __double2hiint(t[r][k]);
                const double w = __hiloint2double(int(0x3ff00000u + (z >> 12)), int(z << 20));
                double pl = w + kB2;
                pl = fma(pl, w, kB1);
                pl = fma(pl, w, kB0);
                dd[r][k] = fma(exp_tab_scaled(tabfx_s, hi), pl, 1.0);
            }
#pragma unroll
        for (int r = 0; r < SB_RI; ++r)
#pragma unroll
            for (int k = 0; k < SB_GI; ++k) {
                double y0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(dd[r][k]));
                const double err1 = fma(-dd[r][k], y0, 2.0);
                const double rho = y0 * err1;
                double h = fma(-rho, rho, rho);
                if (k == 1) h *= w1;
                if (UNIT) {
                    E0[r] = fma(r01[k].y, h, E0[r]);
                    E1[r] = fma(r23[k].x, h, E1[r]);
                    E2[r] = fma(r23[k].y, h, E2[r]);
                } else {
                    const double gam = fma(r23[k].x, dri[r], r23[k].y) * h;
                    E0[r] += gam;
                    E1[r] = fma(gam, r01[k].x, E1[r]);
                    E2[r] = fma(gam, r01[k].y, E2[r]);
                }
            }
    }
}

template <bool SOFT>
__global__ void __launch_bounds__(SCORE_MAX_PAIRS * 64, 1)
k_score_bwd(const ScoreBwdArgs q) {
    const ScoreArgs& p = q.f;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wl = p.wpg_log2, GT = 32 << wl, WPG = 1 << wl;
    const int grp = warp >> wl, half = warp & (WPG - 1), ptid = tid & (GT - 1);
    const int ngroups = blockDim.x >> (5 + wl);
    const int N = p.N, Np = p.Np, G = p.G, Sp = p.Sp, S = p.S;
    const int N32 = (N + 31) & ~31;

    __shared__ __align__(128) double tab[16];
    double* sdr = sm + 16;
    double* pbase = sdr + Np + size_t(grp) * score_bwd_group_doubles(Np, G);
    double* sq = pbase;                                 // [i][3] scaled quadratic coefficients
    double* ga = sq + 3 * size_t(Np);                   // SB_GARR per-grid arrays
    double* g_cnt = ga;            double* g_sr = ga + G;        double* g_sw = ga + 2 * G;   double* g_mk = ga + 3 * G;
    double* g_kap = ga + 4 * G;    double* g_opt = ga + 5 * G;   double* g_vs = ga + 6 * G;   double* g_dmf = ga + 7 * G;
    double* g_dsd = ga + 8 * G;    double* g_scale = ga + 9 * G; double* g_bg = ga + 10 * G;  double* g_kl = ga + 11 * G;
    double* rec = ga + 4 * G;                           // pass-2 records (4 per grid point) overwrite kap..dmf AFTER phase E read them
    double* zt = ga + size_t(SB_GARR) * G;              // 64 nibble-table entries
    double* part1 = zt + 64;                            // [WPG][32]: phase C/D partial sums, then the pass-2 partials
    double* part2 = part1 + 16 * 32;                    // [WPG][32]: phase E partial sums
    double* s_utb = part2 + 16 * 32;                    // ut_bar[Sp] of the current candidate

    __shared__ __align__(128) double tabfx[16];         // score_rows_fx / moment_sweep (soft path)
    if (tid < 16) { exp_tab_init(tab, tid); exp_tabfx_init(tabfx, tid); }
    const unsigned tab_s = smem_u32(tab), tabfx_s = smem_u32(tabfx);
    for (int i = tid; i < Np; i += blockDim.x) sdr[i] = (i < N) ? p.dr[i] : 0.0;
    __syncthreads();

    const double sum_dr = p.sum_dr[0];
    const double noise = p.noise;
    const double kvar = noise / (double(Sp) + 1e-6);
    const double inv_temp = SOFT ? 1.0 / p.temp : 0.0;
    const double c1 = SOFT ? inv_temp * 23.083120654223414 : 1.0;
    const double alpha_lvl = p.alpha_lvl;

    const int total_groups = gridDim.x * ngroups;
    for (int c = blockIdx.x * ngroups + grp; c < p.n_cand; c += total_groups) {
        // ---- phase A: identical to k_score ------------------------------------------------------------
        {
            int cn = c + total_groups;
            if (cn < p.n_cand)
                for (int o = ptid * 16; o < Np; o += GT * 16)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.W + int64_t(cn) * Np + o));
        }
        double vnorm = 0.0;
        for (int t = 0; t < p.tiles_n; ++t) vnorm += p.rn_partial[int64_t(c) * p.tiles_n + t];
        const double s_var = p.kappa + noise - vnorm;
        const double mu = p.mu[c];
        const double inv_s = rcp_full(s_var);
        const double* Wc = p.W + int64_t(c) * Np;
        double zcc = 0.0;
        for (int s = 0; s < Sp; ++s) {
            double zcs = p.z_cand[s];
            zcc = fma(zcs, (zcs - p.ut[int64_t(c) * Sp + s]) * inv_s, zcc);
        }
        if (ptid < 64) {
            const int t4 = (ptid >> 4) * 4, e = ptid & 15;
            double acc = 0.0;
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) {
                int s = t4 + bb;
                if (s < Sp) {
                    double csv = (p.z_cand[s] - p.ut[int64_t(c) * Sp + s]) * inv_s;
                    acc += ((e >> bb) & 1) ? csv : -csv;
                }
            }
            zt[ptid] = acc;
        }
        const double var_c = fma(kvar, double(Sp) - noise * zcc, noise);
        int bad = !(var_c > 0.0) || !(s_var > 0.0);
        const double iv_c = 0.5 * rcp_full(var_c);
        const double lv_c = -0.5 * log_fast(var_c) - 0.5 * kLog2Pi;
        const double qa_c = -iv_c * noise * noise;
        const double cscale = p.cscale ? p.cscale[c] : 1.0;
        const double gvar = cscale * s_var;
        const double gsd = sqrt_fast(gvar);
        const double dens_c = 0.3989422804014327 * rcp_full(gsd);       // 1 / sqrt(2 pi gvar): grid density at its mean
        const double bmax = gsd * p.eps_absmax[p.eps_per_candidate ? c : 0] * inv_s;
        int big = 0;
        pair_barrier(grp, GT);
        for (int i0 = ptid; i0 < N32; i0 += 4 * GT) {
            double w[4], h0[4], ai[4];
            uint32_t bits[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int i = i0 + GT * j;
                bool ok = i < N;
                w[j] = ok ? Wc[i] : 0.0;
                h0[j] = ok ? p.h0[i] : 1.0;
                ai[j] = ok ? p.a[i] : 0.0;
                bits[j] = ok ? p.zb[i] : 0u;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                int i = i0 + GT * j;
                if (i < N32) {
                    double qa = 0.0, qb = 0.0, qc = -900.0;
                    if (i < N) {
                        double zdot = zt[bits[j] & 15] + zt[16 + ((bits[j] >> 4) & 15)];
                        if (Sp > 8) zdot += zt[32 + ((bits[j] >> 8) & 15)] + zt[48 + ((bits[j] >> 12) & 15)];
                        double var = noise * fma(kvar * w[j], zdot, h0[j]);
                        bool okv = var > 0.0;
                        bad |= !okv;
                        var = okv ? var : 1.0;
                        double iv = 0.5 * rcp_full(var);
                        double lv = -0.5 * log_fast(var) - 0.5 * kLog2Pi;
                        double wi = noise * w[j];
                        qa = c1 * fma(iv * wi, wi, qa_c);
                        qb = c1 * (-2.0 * iv * ai[j] * wi);
                        qc = c1 * (fma(iv * ai[j], ai[j], lv_c - lv));
                        big |= !(fma(fma(fabs(qa), bmax, fabs(qb)), bmax, fabs(qc)) < 900.0);
                    }
                    sq[3 * i] = qa; sq[3 * i + 1] = qb; sq[3 * i + 2] = SOFT ? kFxMagic - qc : qc;   // soft path: score_rows_fx wants M - qc
                }
            }
        }
        const double drc = p.drc ? p.drc[c] : 1.0;
        const double norm = sum_dr + drc;
        const double inv_norm = rcp_full(norm);
        const double omega_c = drc / norm;
        bad = pair_barrier_or(grp, GT, bad);
        big = SOFT ? pair_barrier_or(grp, GT, big) : 0;

        // ---- phase B: forward row sweep, 4 grid points per pass ----------------------------------------
        for (int g0 = half * SCORE_GPW; g0 < G; g0 += (SCORE_GPW << wl)) {
            double bg[SCORE_GPW], sr[SCORE_GPW], sw[SCORE_GPW];
            unsigned cneg[SCORE_GPW];
#pragma unroll
            for (int k = 0; k < SCORE_GPW; ++k) {
                int g = min(g0 + k, G - 1);
                double eps = p.eps_grid[p.eps_per_candidate ? int64_t(c) * G + g : g];
                bg[k] = gsd * eps * inv_s;
                sr[k] = 0.0; sw[k] = 0.0; cneg[k] = 0u;
            }
            if (!bad) {
                if (SOFT) {
                    if (p.unit_dr) {
                        if (big) score_rows_fx<true, true, SCORE_GPW>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                        else score_rows_fx<false, true, SCORE_GPW>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                    } else {
                        if (big) score_rows_fx<true, false, SCORE_GPW>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                        else score_rows_fx<false, false, SCORE_GPW>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                    }
                } else {
                    if (p.unit_dr) score_rows<false, false, true, SCORE_GPW>(p, sq, sdr, tab_s, lane, N32, bg, sr, sw, cneg);
                    else score_rows<false, false, false, SCORE_GPW>(p, sq, sdr, tab_s, lane, N32, bg, sr, sw, cneg);
                }
            }
            {   // transposed reductions: lanes 8k..8k+7 end up with the totals of grid point g0 + k
                static_assert(SCORE_GPW == 4, "warp_sum4");
                int ci[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) ci[k] = int(cneg[k]);
                const int k = lane >> 3;
                const int cnt = N32 - warp_sum4(ci, lane);
                const double tr = SOFT ? warp_sum4(sr, lane) : 0.0;
                const double tw = p.unit_dr ? (SOFT ? tr : double(cnt)) : warp_sum4(sw, lane);
                if ((lane & 7) == 0 && g0 + k < G) { g_cnt[g0 + k] = double(cnt); g_sr[g0 + k] = tr; g_sw[g0 + k] = tw; }
            }
        }
        pair_barrier(grp, GT);

        // ---- phase C: thread per grid point (forward), partial sums of phase D ---------------------------
        const double shrink = 1.0 - noise * inv_s;
        const double sd_f = sqrt(fmax(noise * shrink, 0.0));
        const double ood = SOFT ? sigmoid_acc((omega_c - alpha_lvl) * inv_temp) : ((omega_c >= alpha_lvl) ? 1.0 : 0.0);
        double p_scw = 0.0, p_snc = 0.0, p_A = 0.0, p_B = 0.0;
        for (int g = ptid; g < G; g += GT) {
            const double cnt = g_cnt[g], my_sr = g_sr[g], my_sw = g_sw[g];
            const double n_g = cnt + 1.0;
            double M, rho_self, scale = 0.0;
            if (SOFT) {
                scale = n_g * rcp_full(my_sr + 0.5);
                rho_self = 0.5 * scale;
                M = scale * (my_sw + 0.5 * drc) * inv_norm;
            } else {
                rho_self = 1.0;
                M = (my_sw + drc) / norm;
            }
            double keep_live = 0.0;
            if (p.randomized) {
                double Dr = rho_self * omega_c;
                double D = fmax(Dr, 1e-6);
                double pr = fmin(fmax((M - alpha_lvl) / D, 0.0), 1.0);
                double u = p.u_rand[int64_t(c) * G + g];
                bool keep = !(u < pr);
                M -= keep ? D : 0.0;
                keep_live = (keep && Dr > 1e-6) ? 1.0 : 0.0;
            }
            const double mk = SOFT ? sigmoid_pat((M - alpha_lvl) * inv_temp, tab_s) : ((M > alpha_lvl) ? 1.0 : 0.0);
            const double eps = p.eps_grid[p.eps_per_candidate ? int64_t(c) * G + g : g];
            const double dev = gsd * eps;
            const double dens = dens_c * exp_neg_pat(0.5 * eps * eps, tab_s);     // exp(log-density of y_g)
            const double opt = p.optimistic ? ((mu + dev >= mu) ? 1.0 : 0.0) : 1.0;
            const double kap = opt * rcp_full(fmax(dens, 1e-6));
            const double mf = mu + dev * shrink;
            double vs = 0.0, dmf = 0.0, dsd = 0.0;
            double ebar = 0.0;
            if (p.kind == 0 && S > 1) {
                for (int s2 = 0; s2 < S; ++s2) ebar += p.eps_mc[int64_t(s2) * G + g];
                ebar /= double(S);
            }
            for (int s = 0; s < S; ++s) {
                const double em = p.eps_mc[int64_t(s) * G + g];
                if (p.kind == 0) {
                    double coef = (S > 1) ? ebar + p.beta_prime * fabs(em - ebar) : em;
                    vs += fma(sd_f, coef, mf); dmf += 1.0; dsd += coef;
                } else {
                    double xx = fma(sd_f, em, mf) - p.best_f;
                    double val, dv;
                    if (p.temp > 0.0) {
                        double xt = xx / p.temp;
                        val = p.temp * (fmax(xt, 0.0) + log1p(exp(-fabs(xt))));
                        dv = sigmoid_acc(xt);
                    } else {
                        val = fmax(xx, 0.0);
                        dv = (xx > 0.0) ? 1.0 : 0.0;
                    }
                    vs += val; dmf += dv; dsd += dv * em;
                }
            }
            const double cw = mk * kap, nc = (1.0 - mk) * opt;
            p_scw += cw; p_snc += nc; p_A = fma(cw, vs, p_A); p_B = fma(nc, vs, p_B);
            g_mk[g] = mk; g_kap[g] = kap; g_opt[g] = opt; g_vs[g] = vs; g_dmf[g] = dmf; g_dsd[g] = dsd;
            g_scale[g] = scale; g_bg[g] = dev * inv_s; g_kl[g] = keep_live;
        }
        p_scw = warp_sum(p_scw); p_snc = warp_sum(p_snc); p_A = warp_sum(p_A); p_B = warp_sum(p_B);
        if (lane == 0) { double* pp = part1 + half * 32; pp[0] = p_scw; pp[1] = p_snc; pp[2] = p_A; pp[3] = p_B; }
        pair_barrier(grp, GT);

        // ---- phase D: value (every thread redundantly) ----------------------------------------------------
        double scw = 0.0, snc = 0.0, Asum = 0.0, Bsum = 0.0;
        for (int w = 0; w < WPG; ++w) {
            const double* pp = part1 + w * 32;
            scw += pp[0]; snc += pp[1]; Asum += pp[2]; Bsum += pp[3];
        }
        const double dcw = fmax(scw, 1e-6), dnc = fmax(snc, 1e-6);
        const double invS = 1.0 / double(S);
        const double inner = ((1.0 - alpha_lvl) * Asum / dcw + alpha_lvl * Bsum / dnc) * invS;
        const double acq = (1.0 - ood) * inner;
        const double ka = (1.0 - ood) * (1.0 - alpha_lvl) / dcw * invS;
        const double kb = (1.0 - ood) * alpha_lvl / dnc * invS;
        const double scw_bar = (scw > 1e-6) ? -ka * Asum / dcw : 0.0;
        const double snc_bar = (snc > 1e-6) ? -kb * Bsum / dnc : 0.0;

        // ---- phase E: thread per grid point (reverse): P_g, Q_g and the scalar adjoint partial sums ------------
        double e_smf = 0.0, e_smfb = 0.0, e_sd = 0.0, e_drc = 0.0, e_norm = 0.0, e_om = 0.0;
        {
            // records alias g_kap..g_dmf: read everything of this thread's grid points first, write after a barrier
            constexpr int kMaxOwn = 4;                       // G <= 4 * GT is checked by the launcher
            double r_b[kMaxOwn], r_P[kMaxOwn], r_Q[kMaxOwn];
#pragma unroll
            for (int own = 0; own < kMaxOwn; ++own) {
                const int g = ptid + own * GT;
                r_b[own] = r_P[own] = r_Q[own] = 0.0;
                if (g < G) {
                    const double mk = g_mk[g], kap = g_kap[g], opt = g_opt[g], vs = g_vs[g], bgv = g_bg[g];
                    const double cw = mk * kap, nc = (1.0 - mk) * opt;
                    const double vbar = fma(ka, cw, kb * nc);                   // d acq / d values_sg (same for every s)
                    const double mf_bar = vbar * g_dmf[g];
                    e_smf += mf_bar; e_smfb = fma(mf_bar, bgv, e_smfb); e_sd = fma(vbar, g_dsd[g], e_sd);
                    double P = 0.0, Q = 0.0;
                    if (SOFT) {
                        const double mk_bar = fma(ka, vs, scw_bar) * kap - fma(kb, vs, snc_bar) * opt;
                        const double M_bar = mk_bar * mk * (1.0 - mk) * inv_temp;
                        const double scale = g_scale[g], my_sr = g_sr[g], my_sw = g_sw[g], kl = g_kl[g];
                        const double Mraw = scale * (my_sw + 0.5 * drc) * inv_norm;
                        double scale_bar = M_bar * (my_sw + 0.5 * drc) * inv_norm;
                        const double D_bar = -kl * M_bar;                       // randomised step: M -= max(rho_self omega, 1e-6)
                        scale_bar = fma(D_bar, 0.5 * omega_c, scale_bar);
                        e_om = fma(D_bar, 0.5 * scale, e_om);
                        P = M_bar * scale * inv_norm;
                        Q = -scale_bar * scale * rcp_full(my_sr + 0.5);
                        e_drc = fma(0.5, P, e_drc);
                        e_norm -= M_bar * Mraw * inv_norm;
                    }
                    r_b[own] = bgv; r_P[own] = P * inv_temp; r_Q[own] = Q * inv_temp;
                }
            }
            pair_barrier(grp, GT);                                               // every per-grid array has been consumed
#pragma unroll
            for (int own = 0; own < kMaxOwn; ++own) {
                const int g = ptid + own * GT;
                if (g < G) {
                    const double b1 = r_b[own], P = r_P[own], Q = r_Q[own];
                    double* rr = rec + 4 * g;
                    if (p.unit_dr) { const double pq = P + Q; rr[0] = b1; rr[1] = pq; rr[2] = pq * b1; rr[3] = pq * b1 * b1; }
                    else { rr[0] = b1; rr[1] = b1 * b1; rr[2] = P; rr[3] = Q; }
                }
            }
        }
        e_smf = warp_sum(e_smf); e_smfb = warp_sum(e_smfb); e_sd = warp_sum(e_sd);
        e_drc = warp_sum(e_drc); e_norm = warp_sum(e_norm); e_om = warp_sum(e_om);
        if (lane == 0) {
            double* pp = part2 + half * 32;
            pp[0] = e_smf; pp[1] = e_smfb; pp[2] = e_sd; pp[3] = e_drc; pp[4] = e_norm; pp[5] = e_om;
        }
        pair_barrier(grp, GT);                                                   // records and part2 visible

        // ---- pass 2: moments E_k,i and the per-train-point epilogue ---------------------------------------
        double t_sb0 = 0.0, t_sb1 = 0.0, t_e0 = 0.0, t_e2 = 0.0;
        double* wrow = q.wbar + int64_t(c) * Np;
        if (SOFT && !bad) {
            for (int i0 = ptid; i0 < N32; i0 += GT * SB_RI) {
                double qa[SB_RI], qb[SB_RI], qc[SB_RI], dri[SB_RI], E0[SB_RI], E1[SB_RI], E2[SB_RI];
#pragma unroll
                for (int r = 0; r < SB_RI; ++r) {
                    int i = i0 + GT * r;
                    bool ok = i < N32;
                    qa[r] = ok ? sq[3 * i] : 0.0; qb[r] = ok ? sq[3 * i + 1] : 0.0; qc[r] = ok ? sq[3 * i + 2] : kFxMagic + 900.0;   // M - qc
                    dri[r] = (ok && !p.unit_dr) ? sdr[i] : 1.0;
                    E0[r] = E1[r] = E2[r] = 0.0;
                }
                if (p.unit_dr) {
                    if (big) moment_sweep<true, true>(p, rec, G, tabfx_s, qa, qb, qc, dri, E0, E1, E2);
                    else moment_sweep<false, true>(p, rec, G, tabfx_s, qa, qb, qc, dri, E0, E1, E2);
                } else {
                    if (big) moment_sweep<true, false>(p, rec, G, tabfx_s, qa, qb, qc, dri, E0, E1, E2);
                    else moment_sweep<false, false>(p, rec, G, tabfx_s, qa, qb, qc, dri, E0, E1, E2);
                }
#pragma unroll
                for (int r = 0; r < SB_RI; ++r) {
                    int i = i0 + GT * r;
                    if (i < N) {
                        const double w = Wc[i], h0 = p.h0[i], ai = p.a[i];
                        const uint32_t bits = p.zb[i];
                        double zdot = zt[bits & 15] + zt[16 + ((bits >> 4) & 15)];
                        if (Sp > 8) zdot += zt[32 + ((bits >> 8) & 15)] + zt[48 + ((bits >> 12) & 15)];
                        const double var = noise * fma(kvar * w, zdot, h0);
                        const double iv = 0.5 * rcp_full(var);
                        const double wi = noise * w;
                        const double Ai = fma(iv * wi, wi, qa_c), Bi = -2.0 * iv * ai * wi;
                        t_sb0 += fma(2.0 * Ai, E1[r], Bi * E0[r]);
                        t_sb1 += fma(2.0 * Ai, E2[r], Bi * E1[r]);
                        const double iv_bar = fma(E2[r] * wi, wi, fma(-2.0 * E1[r] * ai, wi, E0[r] * ai * ai));
                        double w_bar = 2.0 * iv * noise * fma(E2[r], wi, -E1[r] * ai);
                        const double var_bar = fma(-2.0 * iv * iv, iv_bar, iv * E0[r]);
                        const double vk = var_bar * noise * kvar;
                        w_bar = fma(vk, zdot, w_bar);
                        sq[3 * i] = vk * w;                  // d acq / d (z_i . c): this row's coefficients are dead now
                        t_e0 += E0[r]; t_e2 += E2[r];
                        wrow[i] = w_bar;
                    } else if (i < Np) {
                        wrow[i] = 0.0;
                    }
                }
            }
            for (int i = N32 + ptid; i < Np; i += GT) wrow[i] = 0.0;
        } else {
            for (int i = ptid; i < Np; i += GT) wrow[i] = 0.0;
        }
        t_sb0 = warp_sum(t_sb0); t_sb1 = warp_sum(t_sb1); t_e0 = warp_sum(t_e0); t_e2 = warp_sum(t_e2);
        // c_bar_s = sum_i (+-) d acq/d(z_i . c): second look at this thread's own rows (kept out of the sweep's registers)
        double cbar[MAX_SP];
#pragma unroll
        for (int s = 0; s < MAX_SP; ++s) cbar[s] = 0.0;
        if (SOFT && !bad) {
            for (int i = ptid; i < N; i += GT) {
                const double zd_bar = sq[3 * i];
                const uint32_t bits = p.zb[i];
#pragma unroll
                for (int s = 0; s < MAX_SP; ++s)
                    if (s < Sp) cbar[s] += ((bits >> s) & 1u) ? zd_bar : -zd_bar;
            }
        }
#pragma unroll
        for (int s = 0; s < MAX_SP; ++s)
            if (s < Sp) cbar[s] = warp_sum(cbar[s]);
        if (lane == 0) {
            double* pp = part1 + half * 32;
            pp[0] = t_sb0; pp[1] = t_sb1; pp[2] = t_e0; pp[3] = t_e2;
#pragma unroll
            for (int s = 0; s < MAX_SP; ++s)
                if (s < Sp) pp[4 + s] = cbar[s];
        }
        pair_barrier(grp, GT);

        // ---- per-candidate adjoints (one thread) ---------------------------------------------------------
        if (ptid == 0) {
            double Sb0 = 0.0, Sb1 = 0.0, sE0 = 0.0, sE2 = 0.0, smf = 0.0, smfb = 0.0, sd_bar = 0.0, drc_bar = 0.0,
                   norm_bar = 0.0, om_bar = 0.0;
            double cb[MAX_SP];
            for (int s = 0; s < Sp; ++s) cb[s] = 0.0;
            for (int w = 0; w < WPG; ++w) {
                const double* p1 = part1 + w * 32;
                const double* p2 = part2 + w * 32;
                Sb0 += p1[0]; Sb1 += p1[1]; sE0 += p1[2]; sE2 += p1[3];
                for (int s = 0; s < Sp; ++s) cb[s] += p1[4 + s];
                smf += p2[0]; smfb += p2[1]; sd_bar += p2[2]; drc_bar += p2[3]; norm_bar += p2[4]; om_bar += p2[5];
            }
            double s_bar = 0.0;
            double* adj = q.adj + int64_t(c) * ADJ_STRIDE;
            if (SOFT) {
                om_bar += -inner * ood * (1.0 - ood) * inv_temp;                 // d acq / d ood = -inner
                drc_bar += om_bar / norm;
                norm_bar -= om_bar * omega_c / norm;
                drc_bar += norm_bar;
                const double iv_c_bar = -noise * noise * sE2, lv_c_bar = sE0;
                const double var_c_bar = -2.0 * iv_c * iv_c * iv_c_bar - iv_c * lv_c_bar;
                const double zc_coef = -var_c_bar * kvar * noise;
                for (int s = 0; s < Sp; ++s) {
                    const double zcs = p.z_cand[s];
                    const double cs = (zcs - p.ut[int64_t(c) * Sp + s]) * inv_s;
                    const double cbs = fma(zc_coef, zcs, cb[s]);
                    adj[4 + s] = s_utb[s] = bad ? 0.0 : -cbs * inv_s;
                    s_bar -= cbs * cs * inv_s;
                }
            } else {
                drc_bar = 0.0;
                for (int s = 0; s < Sp; ++s) adj[4 + s] = s_utb[s] = 0.0;
            }
            const double Pf = s_var - noise;
            Sb0 = fma(Pf, smf, Sb0);
            Sb1 = fma(Pf, smfb, Sb1);
            double mu_bar = smf - Sb0 * inv_s;
            s_bar += smfb - Sb1 * inv_s;
            if (sd_f > 0.0) s_bar += sd_bar * noise * noise * inv_s * inv_s / (2.0 * sd_f);
            if (bad) { mu_bar = s_bar = drc_bar = nan(""); }
            adj[0] = mu_bar; adj[1] = s_bar; adj[2] = drc_bar; adj[3] = 0.0;
            p.acq[c] = bad ? nan("") : acq;
        }
        // ---- fold the probe term into the solve's right-hand side: w_bar' = w_bar + z ut_bar (T = A^-1 z) -------------
        if (SOFT) {
            pair_barrier(grp, GT);
            if (!bad)
                for (int i = ptid; i < N; i += GT) {
                    const uint32_t bits = p.zb[i];
                    double acc = wrow[i];
#pragma unroll
                    for (int s = 0; s < MAX_SP; ++s)
                        if (s < Sp) acc += ((bits >> s) & 1u) ? s_utb[s] : -s_utb[s];
                    wrow[i] = acc;
                }
        }
        // part1 / part2 / s_utb are next written after further group barriers of the following candidate
    }
}

}  // namespace cbo
