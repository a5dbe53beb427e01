// score_q_bwd.cuh -- value AND reverse pass for joint batches 2 <= q <= 4 (cUCB / cEI): what the reference's SGLD loop
// differentiates with its default q_batch_size = 2 (conformalbo/optim/optimize.py:224-235, hydra_config/black_box_opt.yaml:10).
// oracle/analytic_grad_q.py is the line-by-line CPU restatement (checked against torch.autograd to 1e-15).
//
// One CTA (256 threads) per candidate batch, as k_score_q (latency regime: batch_limit x restarts candidates per call):
//   phases 0, A, B+C, D   as k_score_q, keeping the per-(grid point, column) sums
//   phase E  (thread per grid point)  d acq/d values, d acq/d mask -> P_gj, Q_gj (d acq / d rho_ijg = P_gj dr_i + Q_gj),
//            refit-mean / refit-factor / shrinkage adjoints, candidate-vs-candidate rank terms
//   pass 2   (thread per train point, loop over the grid)  Gam_ijg = (P_gj dr_i + Q_gj) rho (1 - rho) / temp; per train
//            point the moments m0, mrr, mr, mrd_k of Gam_ig = sum_j Gam_ijg against rp = sqrt(iv_i) r_ig; per column the totals
//            of Gam_ijg {1, B_j^2, B_j, B_j dev_k}
//   final    (one thread) q x q algebra: Cholesky reverse of the refit factor, S^-1 adjoint, S_bar = -S^-T Sinv_bar S^-T;
//            per row: mu_bar, drc_bar, Ssym row, ut_bar;  w_bar rows (right-hand sides of the A^-1 solve) get z ut_bar folded in.
#pragma once
#include <cooperative_groups.h>
#include "score_q.cuh"

namespace cbo {

constexpr int ADJQ_STRIDE = 8 + MAX_SP + MAX_Q;   // per row: {mu_bar, drc_bar, Ssym[0..3], -, -, ut_bar[MAX_SP], Kleak[0..3]}
constexpr int ADJQ_LEAK = 8 + MAX_SP;         // Kleak: adjoint of the PRIOR block k(x_j, x_k) through the target grid (grid_leak)
constexpr int SQB_GJ = 12;                     // per-(grid point, column) arrays
constexpr int SQB_NT = 4 * MAX_Q + 4 * MAX_Q * MAX_Q + 4 * MAX_Q;   // totals: see enum below

struct ScoreQBwdArgs {
    ScoreQArgs f;
    double* wbar;          // rows x Np
    double* adjq;          // rows x ADJQ_STRIDE
    // f.base.ext_values (cNEI): the sample values are another kernel's -- constants of THIS reverse pass; the adjoints
    // d acq / d values_sgj go to val_bar (n_cand x S x G x Q) for that kernel's own reverse pass (k_nei_values_q_bwd).
    double* val_bar;
    // External adjoints (cEHVI / cNEHVI and any other outer integration over several outputs): the kernel is then the
    // vector-Jacobian product of (conf_pred_mask, ood_mask, samples of the refit posterior) with respect to X; its own
    // integration (phase D) is not differentiated and acq is not written.
    const double* ext_mask_bar;   // n_cand x G x Q   d L / d mask_gj        (all three null: the kernel's own integration)
    const double* ext_ood_bar;    // n_cand x Q       d L / d ood_j
    const double* ext_f_bar;      // n_cand x S x G x Q   d L / d f_sgj of the internally sampled refit posterior, or null
    // cbo_params.grid_grad_leak: the reference builds the target grid under torch.no_grad(), but the prior block k(x_j, x_k) of
    // its lazily evaluated covariance is only evaluated when the grid is sampled and so tracks d/dX: add
    // d acq / d y_g  d y_g / d k(x_j, x_k) for joint batches (cNEI / cNEHVI together with k_nei_values_q_bwd, which adds
    // the part that flows through its sample values; cEHVI / cNEHVI per output in vector-Jacobian mode), unscaled grid covariance.
    int grid_leak;
};

__host__ __device__ inline size_t score_q_bwd_smem_bytes(int Np, int G, int Q) {
    return sizeof(double) * (size_t(4 + Q) * Np + size_t(SQB_GJ * Q + 1) * G + 1024 + SQ_WARPS * SQB_NT + SQB_NT +
                             size_t(SQ_THREADS / 2) * (3 + Q));
}

// totals layout (doubles): per column / per pair accumulators reduced over the block
struct SqbIdx {
    static __host__ __device__ constexpr int drc(int j) { return j; }                               // sum_g 0.5 P
    static __host__ __device__ constexpr int nrm(int j) { return MAX_Q + j; }                       // norm_bar partial
    static __host__ __device__ constexpr int om(int j) { return 2 * MAX_Q + j; }                    // omega_bar partial
    static __host__ __device__ constexpr int mu(int j) { return 3 * MAX_Q + j; }                    // sum_g meanf_bar_gj
    static __host__ __device__ constexpr int lf(int j, int k) { return 4 * MAX_Q + j * MAX_Q + k; } // Lf_bar
    static __host__ __device__ constexpr int shr(int j, int k) { return 4 * MAX_Q + MAX_Q * MAX_Q + j * MAX_Q + k; }
    static __host__ __device__ constexpr int t3(int j, int k) { return 4 * MAX_Q + 2 * MAX_Q * MAX_Q + j * MAX_Q + k; }
    static __host__ __device__ constexpr int sinv(int j, int k) { return 4 * MAX_Q + 3 * MAX_Q * MAX_Q + j * MAX_Q + k; }
    static __host__ __device__ constexpr int t1(int j) { return 4 * MAX_Q + 4 * MAX_Q * MAX_Q + j; }
    static __host__ __device__ constexpr int t2(int j) { return 5 * MAX_Q + 4 * MAX_Q * MAX_Q + j; }
    static __host__ __device__ constexpr int t4(int j) { return 6 * MAX_Q + 4 * MAX_Q * MAX_Q + j; }
    static __host__ __device__ constexpr int bsum(int j) { return 7 * MAX_Q + 4 * MAX_Q * MAX_Q + j; }
};
static_assert(SqbIdx::bsum(MAX_Q - 1) < SQB_NT, "totals layout");

// Gradient that reaches the PRIOR block k(x_j, x_k) of the grid covariance through the target grid (cbo_params.grid_grad_leak).
// Everything depends on the grid through dev_g = y_g - mu = Lc eps_g; with Dm = sum_g dev_bar_g dev_g^T and eps_g = Lc^-1 dev_g:
// Lc_bar = tril(Dm Lc^-T), Cov_bar = Lc^-T Phi(Lc^T Lc_bar) Lc^-1 (returned unsymmetrised; the caller adds Cov_bar + Cov_bar^T).
template <int Q>
__device__ inline void grid_leak_from_moments(const double (&Dm)[Q][Q], const double (&Lc)[Q][Q], double (&leak)[Q][Q]) {
    double Lci[Q][Q], Lcb[Q][Q], Pc[Q][Q], Tc[Q][Q];
    for (int j = 0; j < Q; ++j)                                                   // Lci = Lc^-1 (lower)
        for (int i2 = 0; i2 < Q; ++i2) {
            if (i2 < j) Lci[i2][j] = 0.0;
            else if (i2 == j) Lci[i2][j] = 1.0 / Lc[i2][i2];
            else {
                double s = 0.0;
                for (int k = j; k < i2; ++k) s += Lc[i2][k] * Lci[k][j];
                Lci[i2][j] = -s / Lc[i2][i2];
            }
        }
    for (int a = 0; a < Q; ++a)
        for (int m = 0; m < Q; ++m) {
            double s = 0.0;
            if (m <= a)
                for (int k = 0; k < Q; ++k) s += Dm[a][k] * Lci[m][k];
            Lcb[a][m] = s;
        }
    for (int j = 0; j < Q; ++j)
        for (int k = 0; k < Q; ++k) {
            double s = 0.0;
            for (int m = 0; m < Q; ++m) s += Lc[m][j] * Lcb[m][k];
            Pc[j][k] = (k < j) ? s : (k == j ? 0.5 * s : 0.0);
        }
    for (int j = 0; j < Q; ++j)
        for (int k = 0; k < Q; ++k) {
            double s = 0.0;
            for (int m = 0; m < Q; ++m) s += Pc[j][m] * Lci[m][k];
            Tc[j][k] = s;
        }
    for (int j = 0; j < Q; ++j)
        for (int k = 0; k < Q; ++k) {
            double s = 0.0;
            for (int m = 0; m < Q; ++m) s += Lci[m][j] * Tc[m][k];
            leak[j][k] = s;
        }
}

// add this thread's value to totals slot `slot` (deterministic: warp tree, then fixed-order sum of the warp partials)
struct SqbReducer {
    double* wred;          // [SQ_WARPS][SQB_NT]
    double* tot;           // [SQB_NT]
    int tid;
    __device__ __forceinline__ void put(int slot, double v) const {
        v = warp_sum(v);
        if ((tid & 31) == 0) wred[(tid >> 5) * SQB_NT + slot] = v;
    }
    __device__ __forceinline__ void zero_partials() const {
        for (int e = tid; e < SQ_WARPS * SQB_NT; e += SQ_THREADS) wred[e] = 0.0;
    }
    __device__ __forceinline__ void commit() const {          // call between two __syncthreads()
        for (int e = tid; e < SQB_NT; e += SQ_THREADS) {
            double s = 0.0;
#pragma unroll
            for (int w = 0; w < SQ_WARPS; ++w) s += wred[w * SQB_NT + e];
            tot[e] += s;
        }
    }
};

// NV block sums behind ONE pair of barriers; every value is summed exactly as block_sum does it (same bits)
template <int NV>
__device__ __forceinline__ void block_sum_n(double (&v)[NV], double* red, int tid) {
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = warp_sum(v[k]);
    __syncthreads();
    if ((tid & 31) == 0) {
#pragma unroll
        for (int k = 0; k < NV; ++k) red[(tid >> 5) * NV + k] = v[k];
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < SQ_WARPS; ++w) t += red[w * NV + k];
        v[k] = t;
    }
}

// CL = true: a thread-block CLUSTER of SQB_CLUSTER CTAs works on ONE joint batch (a handful of joint batches per call would
// otherwise keep 8 SMs busy): every CTA owns a window of the train points for the O(N G q) sweeps (phase 0 Gram sums, phase A
// tables, phase B, pass 2), the partial sums are exchanged through distributed shared memory (cluster.map_shared_rank) at
// three cluster barriers, the O(G q) phases C-E run redundantly in every CTA, rank 0 does the final q x q algebra.
constexpr int SQB_CLUSTER = 8;

// Phase clocks of CTA 0 (experiment builds only: tools/sqb_phase_clocks.py compiles with -DCBO_SQB_CLOCKS; never in libcbo.so)
#ifdef CBO_SQB_CLOCKS
__device__ long long g_sqb_clk[32];
#define SQB_CLK(k) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_sqb_clk[k] = clock64(); } while (0)
#else
#define SQB_CLK(k) do { } while (0)
#endif

template <int Q, bool SOFT, bool CL>
__global__ void __launch_bounds__(SQ_THREADS, 1)
k_score_q_bwd(const ScoreQBwdArgs qb) {
    namespace cg = cooperative_groups;
    const ScoreQArgs& pq = qb.f;
    const ScoreArgs& p = pq.base;
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int N = p.N, Np = p.Np, G = p.G, Sp = p.Sp, S = p.S;
    double* s_lv = sm;                       // Np
    double* s_ap = s_lv + Np;                // Np
    double* s_dr = s_ap + Np;                // Np
    double* s_siv = s_dr + Np;               // Np   sqrt(iv_i)
    double* s_wp = s_siv + Np;               // Q x Np   (pass 2 overwrites row j, entry i with zdot_bar_ji)
    double* gj = s_wp + size_t(Q) * Np;      // SQB_GJ arrays of G x Q
    const size_t GQ = size_t(G) * Q;
    double* g_mask = gj;             double* g_opt = gj + GQ;        double* g_dev = gj + 2 * GQ;    double* g_Bg = gj + 3 * GQ;
    double* g_thr = gj + 4 * GQ;     double* g_sr = gj + 5 * GQ;     double* g_sw = gj + 6 * GQ;     double* g_scale = gj + 7 * GQ;
    double* g_P = gj + 8 * GQ;       double* g_Q = gj + 9 * GQ;      double* g_kl = gj + 10 * GQ;    double* g_Mraw = gj + 11 * GQ;
    double* g_kd = gj + size_t(SQB_GJ) * GQ; // G   1 / max(exp(logp), 1e-6)
    double* sc = g_kd + G;                   // scratch: 1024 doubles
    double* red = sc;                        // [4]
    double* s_S = sc + 8;                    // Q x Q
    double* s_Sinv = s_S + 16;
    double* s_Lc = s_Sinv + 16;
    double* s_Lf = s_Lc + 16;
    double* s_C = s_Lf + 16;                 // Q x MAX_SP
    double* s_R = s_C + MAX_Q * MAX_SP;      // Q x MAX_SP   z_c - U^T T
    double* s_misc = s_R + MAX_Q * MAX_SP;   // var_c[Q], logdet, bad
    double* s_A = s_misc + 16;               // [MAX_S][MAX_Q]  A_sj
    double* s_B = s_A + MAX_S * MAX_Q;       // [MAX_S][MAX_Q]
    double* s_kap = s_B + MAX_S * MAX_Q;     // [MAX_S][MAX_Q]  d acq / d res_sj
    double* s_Cbar = s_kap + MAX_S * MAX_Q;  // Q x MAX_SP
    double* s_utb = s_Cbar + MAX_Q * MAX_SP; // Q x MAX_SP
    double* s_xg = s_utb + MAX_Q * MAX_SP;   // [Q*Q + 1]: this CTA's Gram partial sums and its `bad` flag (cluster exchange)
    double* red_n = sc + 576;                // [SQ_WARPS][2 MAX_Q]  batched block sums (phase D)
    double* wred = sc + 1024;                // [SQ_WARPS][SQB_NT]
    double* tot = wred + SQ_WARPS * SQB_NT;  // [SQB_NT]
    double* s_x2 = tot + SQB_NT;             // [SQ_THREADS / 2][3 + Q]  pass 2 (cluster): moments of the second half of the grid
    const SqbReducer R{wred, tot, tid};
    const double noise = p.noise;
    const double kvar = noise / (double(Sp) + 1e-6);
    const double inv_temp = SOFT ? 1.0 / p.temp : 0.0;
    const double c1q = inv_temp * 23.083120654223414;                    // 16 / (temp ln 2)
    const double sum_dr = p.sum_dr[0];
    const double alpha_lvl = p.alpha_lvl;
    const bool ext_adj = qb.ext_mask_bar != nullptr;
    __shared__ __align__(128) double tab[16];
    if (tid < 16) exp_tab_init(tab, tid);
    const unsigned tab_s = smem_u32(tab);

    for (int i = tid; i < N; i += SQ_THREADS) s_dr[i] = p.dr[i];
    __syncthreads();

    // window of train points of this CTA: everything (one CTA per joint batch) or 1 / SQB_CLUSTER of it
    int crank = 0;
    if constexpr (CL) crank = int(cg::this_cluster().block_rank());
    const int NS = CL ? ((N + SQB_CLUSTER * 32 - 1) / (SQB_CLUSTER * 32)) * 32 : N;
    const int i_lo = CL ? min(N, crank * NS) : 0, i_hi = CL ? min(N, (crank + 1) * NS) : N;
    const int p_lo = CL ? min(Np, crank * NS) : 0;                       // pass 2 also zero-fills the padding rows
    const int p_hi = CL ? ((crank == SQB_CLUSTER - 1) ? Np : min(Np, (crank + 1) * NS)) : Np;
    const int cstep = CL ? gridDim.x / SQB_CLUSTER : gridDim.x;

    for (int c = CL ? blockIdx.x / SQB_CLUSTER : blockIdx.x; c < p.n_cand; c += cstep) {
        const int r0 = c * Q;
        for (int e = tid; e < SQB_NT; e += SQ_THREADS) tot[e] = 0.0;
        SQB_CLK(0);
        // ---- phase 0 (as k_score_q) -----------------------------------------------------------------------------------
        double gsum[Q][Q];
#pragma unroll
        for (int j = 0; j < Q; ++j)
#pragma unroll
            for (int k = 0; k < Q; ++k) gsum[j][k] = 0.0;
        for (int i = i_lo + tid; i < i_hi; i += SQ_THREADS) {
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
                if (k <= j) gsum[j][k] = block_sum(gsum[j][k], red, tid);
        SQB_CLK(1);
        if constexpr (CL) {                                              // Gram sums over all windows, fixed rank order
            if (tid == 0)
                for (int j = 0; j < Q; ++j)
                    for (int k = 0; k <= j; ++k) s_xg[j * Q + k] = gsum[j][k];
            cg::this_cluster().sync();
            if (tid == 0)
                for (int j = 0; j < Q; ++j)
                    for (int k = 0; k <= j; ++k) {
                        double t = 0.0;
                        for (int r = 0; r < SQB_CLUSTER; ++r) t += cg::this_cluster().map_shared_rank(s_xg, r)[j * Q + k];
                        gsum[j][k] = t;
                    }
        }
        SQB_CLK(2);
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
                    double v = kk - gsum[j][k] + ((j == k) ? noise : 0.0);
                    Sm[j][k] = v; Sm[k][j] = v;
                }
            bool ok = chol_lower<Q>(Sm, Ls);
            inv_from_chol<Q>(Ls, Si);
            double Cov[Q][Q], Lc[Q][Q];
            for (int j = 0; j < Q; ++j) {
                double rs = p.cscale ? p.cscale[r0 + j] : 1.0;
                for (int k = 0; k < Q; ++k) Cov[j][k] = rs * Sm[j][k];
            }
            ok = chol_lower<Q>(Cov, Lc) && ok;
            double logdet = 0.0;
            for (int j = 0; j < Q; ++j) logdet += 2.0 * log(Lc[j][j]);
            double Cf[Q][Q], Lf[Q][Q];
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k)
                    Cf[j][k] = noise * (((j == k) ? 1.0 : 0.0) - noise * 0.5 * (Si[j][k] + Si[k][j]));
            ok = chol_lower<Q>(Cf, Lf) && ok;
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    s_S[j * Q + k] = Sm[j][k]; s_Sinv[j * Q + k] = Si[j][k];
                    s_Lc[j * Q + k] = Lc[j][k]; s_Lf[j * Q + k] = Lf[j][k];
                }
            for (int j = 0; j < Q; ++j)
                for (int s = 0; s < Sp; ++s) s_R[j * MAX_SP + s] = p.z_cand[j * Sp + s] - p.ut[int64_t(r0 + j) * Sp + s];
            for (int j = 0; j < Q; ++j) {
                double zc_c = 0.0;
                for (int s = 0; s < Sp; ++s) {
                    double cv = 0.0;
                    for (int k = 0; k < Q; ++k) cv += Si[j][k] * s_R[k * MAX_SP + s];
                    s_C[j * MAX_SP + s] = cv;
                    zc_c = fma(p.z_cand[j * Sp + s], cv, zc_c);
                }
                double vc = noise * (1.0 + (double(Sp) - noise * zc_c) / (double(Sp) + 1e-6));
                s_misc[j] = vc;
                ok = ok && (vc > 0.0);
            }
            s_misc[Q] = logdet;
            s_misc[Q + 1] = ok ? 0.0 : 1.0;
        }
        __syncthreads();
        SQB_CLK(3);
        int bad = s_misc[Q + 1] != 0.0;

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

        // ---- phase A: per-train-point tables ---------------------------------------------------------------------------
        for (int i = i_lo + tid; i < i_hi; i += SQ_THREADS) {
            uint32_t bits = p.zb[i];
            double w[Q], acc = 0.0;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                w[j] = p.W[int64_t(r0 + j) * Np + i];
                double zdot = 0.0;
                for (int s = 0; s < Sp; ++s) zdot += ((bits >> s) & 1u) ? s_C[j * MAX_SP + s] : -s_C[j * MAX_SP + s];
                acc = fma(w[j], zdot, acc);
            }
            double var = noise * fma(kvar, acc, p.h0[i]);
            bad |= !(var > 0.0);
            double iv = 0.5 / var;
            double siv = sqrt(iv);
            s_lv[i] = -0.5 * log(var) - 0.5 * kLog2Pi;
            s_ap[i] = p.a[i] * siv;
            s_siv[i] = siv;
#pragma unroll
            for (int j = 0; j < Q; ++j) s_wp[size_t(j) * Np + i] = noise * w[j] * siv;
        }
        // per-grid-point quantities (dev = Lc eps, B = S^-1 dev, threshold score), thread per grid point, counted from the
        // top of the block so that they run beside phase A (which occupies the low threads of a cluster window)
        for (int g = SQ_THREADS - 1 - tid; g < G; g += SQ_THREADS) {
            double eps[Q], dev[Q];
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
            g_kd[g] = e2;
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                double s = 0.0;
#pragma unroll
                for (int k = 0; k < Q; ++k) s = fma(s_Sinv[j * Q + k], dev[k], s);
                const double rc = noise * s;
                g_dev[g * Q + j] = dev[j]; g_Bg[g * Q + j] = s; g_thr[g * Q + j] = fma(-rc * rc, ivc[j], lvc[j]);
            }
        }
        bad = __syncthreads_or(bad);                                     // CL: this window only; combined after phase B
        SQB_CLK(4);

        // ---- phase B + C: warp per grid point (forward), GU grid points in flight -------------------------------------------
        constexpr int GU = (Q <= 2) ? 2 : 1;
        for (int g0 = warp; g0 < G; g0 += GU * SQ_WARPS) {
            double Bg[GU][Q], thr[GU][Q], sr[GU][Q], sw[GU][Q];
            int cnt[GU][Q];
#pragma unroll
            for (int u = 0; u < GU; ++u) {
                const int g = min(g0 + u * SQ_WARPS, G - 1);
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    Bg[u][j] = g_Bg[g * Q + j]; thr[u][j] = g_thr[g * Q + j];
                    sr[u][j] = 0.0; sw[u][j] = 0.0; cnt[u][j] = 0;
                }
            }
            if (!bad) {
                // two train points per lane and step (i, i + 32): 2 GU Q independent sigmoid chains
                for (int i = i_lo + lane; i < i_hi; i += 64) {
                    const int i2 = i + 32;
                    const bool ok2 = i2 < i_hi;
                    const int i2c = ok2 ? i2 : i;
                    const double ap0 = s_ap[i], ap1 = s_ap[i2c], lv0 = s_lv[i], lv1 = s_lv[i2c];
                    const double dr0 = s_dr[i], dr1 = ok2 ? s_dr[i2c] : 0.0;
                    double w0[Q], w1[Q];
#pragma unroll
                    for (int j = 0; j < Q; ++j) { w0[j] = s_wp[size_t(j) * Np + i]; w1[j] = s_wp[size_t(j) * Np + i2c]; }
#pragma unroll
                    for (int u = 0; u < GU; ++u) {
                        double rp0 = ap0, rp1 = ap1;
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            rp0 = fma(-w0[j], Bg[u][j], rp0);
                            rp1 = fma(-w1[j], Bg[u][j], rp1);
                        }
                        const double sci0 = fma(-rp0, rp0, lv0), sci1 = fma(-rp1, rp1, lv1);
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            const double dd0 = thr[u][j] - sci0, dd1 = thr[u][j] - sci1;
                            const bool ge0 = dd0 >= 0.0, ge1 = ok2 && (dd1 >= 0.0);
                            cnt[u][j] += (ge0 ? 1 : 0) + (ge1 ? 1 : 0);
                            if (SOFT) {
                                const double rho0 = soft_rho<true>(dd0 * c1q, tab_s, p.k_magic, p.k_c0, p.k_c1n, p.k_c2, p.k_c3n);
                                double rho1 = soft_rho<true>(dd1 * c1q, tab_s, p.k_magic, p.k_c0, p.k_c1n, p.k_c2, p.k_c3n);
                                rho1 = ok2 ? rho1 : 0.0;
                                sr[u][j] += rho0 + rho1;
                                sw[u][j] = fma(rho0, dr0, fma(rho1, dr1, sw[u][j]));
                            } else {
                                sw[u][j] += (ge0 ? dr0 : 0.0) + (ge1 ? dr1 : 0.0);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < GU; ++u)
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    cnt[u][j] = warp_sum_int(cnt[u][j]);
                    sw[u][j] = warp_sum(sw[u][j]);
                    if (SOFT) sr[u][j] = warp_sum(sr[u][j]);
                }
            if (lane == 0) {           // raw sums only: the per-(grid point, column) epilogue runs thread-parallel below
#pragma unroll
                for (int u = 0; u < GU; ++u) {
                    const int g = g0 + u * SQ_WARPS;
                    if (g < G) {
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            const int o = g * Q + j;
                            if (CL) { g_P[o] = sr[u][j]; g_Q[o] = sw[u][j]; g_kl[o] = double(cnt[u][j]); }   // window partials
                            else { g_sr[o] = sr[u][j]; g_sw[o] = sw[u][j]; g_scale[o] = double(cnt[u][j]); }
                        }
                    }
                }
            }
        }
        SQB_CLK(5);
        if constexpr (CL) {        // sums over the windows of all CTAs of the cluster (fixed rank order) + the `bad` flags
            if (tid == 0) s_xg[Q * Q] = bad ? 1.0 : 0.0;
            cg::this_cluster().sync();
            SQB_CLK(6);
            auto cl = cg::this_cluster();
            for (int o = tid; o < G * Q; o += SQ_THREADS) {
                double a = 0.0, b2 = 0.0, c2 = 0.0;
                for (int r = 0; r < SQB_CLUSTER; ++r) {
                    a += cl.map_shared_rank(g_P, r)[o]; b2 += cl.map_shared_rank(g_Q, r)[o]; c2 += cl.map_shared_rank(g_kl, r)[o];
                }
                g_sr[o] = a; g_sw[o] = b2; g_scale[o] = c2;
            }
            int anybad = 0;
            for (int r = 0; r < SQB_CLUSTER; ++r) anybad |= cl.map_shared_rank(s_xg, r)[Q * Q] != 0.0;
            bad = anybad;
            cl.sync();             // every remote read of the partial arrays is done before phase C reuses them
        }
        __syncthreads();
        SQB_CLK(7);
        // ---- phase C: thread per (grid point, column) ------------------------------------------------------------------------
        for (int o = tid; o < G * Q; o += SQ_THREADS) {
            const int g = o / Q, j = o - g * Q;
            int n_g = int(g_scale[o]);
            double srj = g_sr[o];
            const double thr_j = g_thr[o], sw_j = g_sw[o];
            for (int k = 0; k < Q; ++k) {
                double dd = thr_j - g_thr[g * Q + k];
                n_g += (dd >= 0.0) ? 1 : 0;
                if (SOFT) srj += sigmoid_acc(fmin(fmax(dd * inv_temp, -745.0), 745.0));
            }
            // per-column scalars by index (the register arrays are indexed dynamically here)
            double drc_j = drc[0], norm_j = norm[0], omega_j = omega[0], mu_j = mu[0];
#pragma unroll
            for (int k = 1; k < Q; ++k)
                if (k == j) { drc_j = drc[k]; norm_j = norm[k]; omega_j = omega[k]; mu_j = mu[k]; }
            double M, rho_self, scale = 0.0;
            if (SOFT) {
                scale = double(n_g) / srj;
                rho_self = 0.5 * scale;
                M = scale * (sw_j + 0.5 * drc_j) / norm_j;
            } else {
                rho_self = 1.0;
                M = (sw_j + drc_j) / norm_j;
            }
            const double Mraw = M;
            double kl = 0.0;
            if (p.randomized) {
                double Dr = rho_self * omega_j;
                double D = (j == Q - 1) ? fmax(Dr, 1e-6) : 1e-6;
                double pr = fmin(fmax((M - alpha_lvl) / D, 0.0), 1.0);
                double u = p.u_rand[(int64_t(c) * G + g) * Q + j];
                bool keep = !(u < pr);
                M -= keep ? D : 0.0;
                kl = (keep && j == Q - 1 && Dr > 1e-6) ? 1.0 : 0.0;
            }
            double mk = SOFT ? sigmoid_acc((M - alpha_lvl) * inv_temp) : ((M > alpha_lvl) ? 1.0 : 0.0);
            double yg = mu_j + g_dev[o];
            g_mask[o] = mk;
            g_opt[o] = p.optimistic ? ((yg >= mu_j) ? 1.0 : 0.0) : 1.0;
            g_sr[o] = srj; g_scale[o] = scale; g_kl[o] = kl; g_Mraw[o] = Mraw; g_P[o] = 0.0; g_Q[o] = 0.0;
        }
        __syncthreads();
        for (int g = tid; g < G; g += SQ_THREADS) {
            double lp = -0.5 * (g_kd[g] + logdet + double(Q) * kLog2Pi);
            g_kd[g] = 1.0 / fmax(exp(lp), 1e-6);
        }
        __syncthreads();
        SQB_CLK(8);

        // sample value of (s, g, j): f = refit mean + (L_f eps)_j; returns val, d val / d f pieces
        auto sample_val = [&](int s, int g, int j, double mu_j, double& val, double& dv_f, double& sgn) {
            double mf = mu_j, fs = 0.0, ms = 0.0;
#pragma unroll
            for (int k = 0; k < Q; ++k) {
                double shr = ((j == k) ? 1.0 : 0.0) - noise * s_Sinv[j * Q + k];
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
            double f = mf + fs;
            sgn = 0.0; dv_f = 1.0;
            if (p.ext_values) {            // another kernel's values (cNEI): constants here, their adjoint is exported
                val = p.ext_values[((int64_t(c) * S + s) * G + g) * Q + j];
                dv_f = 0.0;
            } else if (p.kind == 0) {
                if (S > 1) { double m = mf + ms; double dm = f - m; sgn = (dm > 0.0) ? 1.0 : (dm < 0.0 ? -1.0 : 0.0); val = m + p.beta_prime * fabs(dm); }
                else val = f;
            } else {
                double xx = f - p.best_f;
                if (p.temp > 0.0) { double xt = xx / p.temp; val = p.temp * (fmax(xt, 0.0) + log1p(exp(-fabs(xt)))); dv_f = sigmoid_acc(xt); }
                else { val = fmax(xx, 0.0); dv_f = (xx > 0.0) ? 1.0 : 0.0; }
            }
        };

        // ---- phase D: forward integration per column; arg-max over q per MC sample ----------------------------------------
        double ood[Q], scw[Q], snc[Q];
        {
            double v[2 * Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                ood[j] = SOFT ? sigmoid_acc((omega[j] - alpha_lvl) * inv_temp) : ((omega[j] >= alpha_lvl) ? 1.0 : 0.0);
                double a1 = 0.0, a2 = 0.0;
                for (int g = tid; g < G; g += SQ_THREADS) {
                    double mk = g_mask[g * Q + j], op = g_opt[g * Q + j];
                    a1 += mk * op * g_kd[g];
                    a2 += (1.0 - mk) * op;
                }
                v[2 * j] = a1; v[2 * j + 1] = a2;
            }
            block_sum_n<2 * Q>(v, red_n, tid);
#pragma unroll
            for (int j = 0; j < Q; ++j) { scw[j] = v[2 * j]; snc[j] = v[2 * j + 1]; }
        }
        double tot_acq = 0.0;
        for (int s = 0; s < S; ++s) {
            double best = -INFINITY;
            int jbest = 0;
            bool isnan_any = false;
            double v[2 * Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                double A = 0.0, B = 0.0;
                for (int g = tid; g < G; g += SQ_THREADS) {
                    double val, dv, sg;
                    sample_val(s, g, j, mu[j], val, dv, sg);
                    double mk = g_mask[g * Q + j], op = g_opt[g * Q + j];
                    A = fma(mk * op * g_kd[g], val, A);
                    B = fma((1.0 - mk) * op, val, B);
                }
                v[2 * j] = A; v[2 * j + 1] = B;
            }
            block_sum_n<2 * Q>(v, red_n, tid);
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                const double A = v[2 * j], B = v[2 * j + 1];
                double r = (1.0 - ood[j]) * ((1.0 - alpha_lvl) * A / fmax(scw[j], 1e-6) + alpha_lvl * B / fmax(snc[j], 1e-6));
                isnan_any = isnan_any || (r != r);
                if (r > best) { best = r; jbest = j; }
                if (tid == 0) { s_A[s * MAX_Q + j] = A; s_B[s * MAX_Q + j] = B; }
            }
            tot_acq += isnan_any ? nan("") : best;
            if (tid == 0)
                for (int j = 0; j < Q; ++j) s_kap[s * MAX_Q + j] = (j == jbest) ? 1.0 / double(S) : 0.0;
        }
        __syncthreads();
        SQB_CLK(9);

        // ---- phase E: thread per (grid point, column), whole warps per column (reverse) ---------------------------------------
        R.zero_partials();
        __syncthreads();
        {
            constexpr int TPJ = ((SQ_THREADS / Q) / 32) * 32;
            const int jt = tid / TPJ, gt = tid - jt * TPJ;
            const bool jact = jt < Q;                                      // warp-uniform
            const int j = jact ? jt : 0;
            double ood_j = ood[0], scw_j = scw[0], snc_j = snc[0], drc_j = drc[0], norm_j = norm[0], omega_j = omega[0], mu_j = mu[0];
#pragma unroll
            for (int k = 1; k < Q; ++k)
                if (k == j) { ood_j = ood[k]; scw_j = scw[k]; snc_j = snc[k]; drc_j = drc[k]; norm_j = norm[k]; omega_j = omega[k]; mu_j = mu[k]; }
            double a_drc = 0.0, a_nrm = 0.0, a_om = 0.0, a_mu = 0.0, a_lf[Q], a_shr[Q];
#pragma unroll
            for (int k = 0; k < Q; ++k) a_lf[k] = a_shr[k] = 0.0;
            if (!bad && jact) {
                const double dcw = fmax(scw_j, 1e-6), dnc = fmax(snc_j, 1e-6);
                const double c1 = (1.0 - ood_j) * (1.0 - alpha_lvl), c2 = (1.0 - ood_j) * alpha_lvl;
                for (int g = gt; g < G; g += TPJ) {
                    const int o = g * Q + j;
                    const double kd = g_kd[g];
                    const double mk = g_mask[o], op = g_opt[o];
                    const double cw = mk * op * kd, nc = (1.0 - mk) * op;
                    const double wv = c1 * cw / dcw + c2 * nc / dnc;      // d res_sj / d val_sgj
                    double cw_bar = 0.0, nc_bar = 0.0, meanf_bar = 0.0;
                    double m_bar = 0.0;
                    if (ext_adj) {                 // adjoints of an outer integration: masks and samples only
                        if (qb.ext_f_bar)
                            for (int s = 0; s < S; ++s) {
                                const double f_bar = qb.ext_f_bar[((int64_t(c) * S + s) * G + g) * Q + j];
                                meanf_bar += f_bar;
#pragma unroll
                                for (int k = 0; k < Q; ++k)
                                    if (k <= j) a_lf[k] = fma(f_bar, p.eps_mc[(int64_t(s) * G + g) * Q + k], a_lf[k]);
                            }
                    } else {
                        // two sweeps over the MC samples for the S > 1 cUCB mean term
                        for (int s = 0; s < S; ++s) {
                            const double kap = s_kap[s * MAX_Q + j];
                            if (kap == 0.0) continue;
                            double val, dv, sg;
                            sample_val(s, g, j, mu_j, val, dv, sg);
                            cw_bar += kap * c1 * (val / dcw - ((scw_j > 1e-6) ? s_A[s * MAX_Q + j] / (dcw * dcw) : 0.0));
                            nc_bar += kap * c2 * (val / dnc - ((snc_j > 1e-6) ? s_B[s * MAX_Q + j] / (dnc * dnc) : 0.0));
                            if (p.kind == 0 && !p.ext_values) m_bar += kap * wv * (1.0 - p.beta_prime * sg);
                        }
                        for (int s = 0; s < S; ++s) {
                            const double kap = s_kap[s * MAX_Q + j];
                            double f_bar;
                            if (p.ext_values) {
                                if (qb.val_bar && crank == 0) qb.val_bar[((int64_t(c) * S + s) * G + g) * Q + j] = kap * wv;
                                continue;
                            }
                            if (p.kind == 0) {
                                double val, dv, sg = 0.0;
                                if (kap != 0.0 && S > 1) sample_val(s, g, j, mu_j, val, dv, sg);
                                f_bar = kap * wv * p.beta_prime * sg + m_bar / double(S);
                                if (S == 1) f_bar = kap * wv;
                            } else {
                                if (kap == 0.0) continue;
                                double val, dv, sg;
                                sample_val(s, g, j, mu_j, val, dv, sg);
                                f_bar = kap * wv * dv;
                            }
                            meanf_bar += f_bar;
#pragma unroll
                            for (int k = 0; k < Q; ++k)
                                if (k <= j) a_lf[k] = fma(f_bar, p.eps_mc[(int64_t(s) * G + g) * Q + k], a_lf[k]);
                        }
                    }
                    a_mu += meanf_bar;
#pragma unroll
                    for (int k = 0; k < Q; ++k) a_shr[k] = fma(meanf_bar, g_dev[g * Q + k], a_shr[k]);
                    if (SOFT) {
                        const double mk_bar = ext_adj ? qb.ext_mask_bar[(int64_t(c) * G + g) * Q + j] : (cw_bar * kd - nc_bar) * op;
                        const double M_bar = mk_bar * mk * (1.0 - mk) * inv_temp;
                        const double scale = g_scale[o], srj = g_sr[o], swj = g_sw[o], kl = g_kl[o], Mraw = g_Mraw[o];
                        double scale_bar = M_bar * (swj + 0.5 * drc_j) / norm_j;
                        const double D_bar = -kl * M_bar;
                        scale_bar = fma(D_bar, 0.5 * omega_j, scale_bar);
                        a_om = fma(D_bar, 0.5 * scale, a_om);
                        const double P = M_bar * scale / norm_j;
                        const double Qv = -scale_bar * scale / srj;
                        a_drc = fma(0.5, P, a_drc);
                        a_nrm -= M_bar * Mraw / norm_j;
                        g_P[o] = P * inv_temp; g_Q[o] = Qv * inv_temp;
                    }
                }
            }
            __syncthreads();
            // candidate-vs-candidate rank terms: d/d(thr_j - thr_k) = Q_gj rhoc (1 - rhoc) / temp, and rhoc (1 - rhoc) is the same
            // for the pair (k, j): thr_bar_gj = sum_k rhoc (1 - rhoc) (Q_gj - Q_gk) / temp
            double a_t1 = 0.0, a_t2 = 0.0, a_t4 = 0.0, a_t3[Q];
#pragma unroll
            for (int k = 0; k < Q; ++k) a_t3[k] = 0.0;
            if (SOFT && !bad && jact) {
                for (int g = gt; g < G; g += TPJ) {
                    const int o = g * Q + j;
                    const double thr_j = g_thr[o], gq_j = g_Q[o];
                    double tb = 0.0;
#pragma unroll
                    for (int k = 0; k < Q; ++k)
                        if (k != j) {
                            const double rc = sigmoid_acc(fmin(fmax((thr_j - g_thr[g * Q + k]) * inv_temp, -745.0), 745.0));
                            tb = fma(rc * (1.0 - rc), gq_j - g_Q[g * Q + k], tb);
                        }
                    const double Bj = g_Bg[o];
                    a_t1 += tb; a_t2 = fma(tb * Bj, Bj, a_t2); a_t4 = fma(tb, Bj, a_t4);
#pragma unroll
                    for (int k = 0; k < Q; ++k) a_t3[k] = fma(tb * Bj, g_dev[g * Q + k], a_t3[k]);
                }
            }
            SQB_CLK(19);
            if (jact) {
                R.put(SqbIdx::drc(j), a_drc); R.put(SqbIdx::nrm(j), a_nrm); R.put(SqbIdx::om(j), a_om);
                R.put(SqbIdx::mu(j), a_mu); R.put(SqbIdx::t1(j), a_t1); R.put(SqbIdx::t2(j), a_t2);
                R.put(SqbIdx::t4(j), a_t4);
#pragma unroll
                for (int k = 0; k < Q; ++k) {
                    R.put(SqbIdx::lf(j, k), a_lf[k]); R.put(SqbIdx::shr(j, k), a_shr[k]); R.put(SqbIdx::t3(j, k), a_t3[k]);
                }
            }
        }
        __syncthreads();
        SQB_CLK(10);
        R.commit();
        __syncthreads();
        if (CL && crank != 0) {    // phases C-E ran redundantly: their totals count once (rank 0) in the cluster sum
            for (int e = tid; e < SQB_NT; e += SQ_THREADS) tot[e] = 0.0;
            __syncthreads();
        }

        // ---- pass 2: thread per train point, loop over the grid ------------------------------------------------------------------
        R.zero_partials();
        __syncthreads();
        SQB_CLK(11);
        {
            double a_t1[Q], a_t2[Q], a_t4[Q], a_t3[Q][Q], a_bs[Q], a_si[Q][Q];
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                a_t1[j] = a_t2[j] = a_t4[j] = a_bs[j] = 0.0;
#pragma unroll
                for (int k = 0; k < Q; ++k) a_t3[j][k] = a_si[j][k] = 0.0;
            }
            if constexpr (CL) {
                // cluster window (N / 8 train points): the two halves of the block take the two halves of the grid for the same
                // train point; the second half hands its moments over through shared memory, the first half finishes the row
                constexpr int HT = SQ_THREADS / 2, XS = 3 + Q;
                const int half = tid / HT, th = tid - half * HT;
                const int Gh = (G + 1) / 2, g_lo = half ? Gh : 0, g_hi = half ? G : Gh;
                for (int base = p_lo; base < p_hi; base += HT) {
                    const int i = base + th;
                    const bool ok = (i < i_hi) && !bad;                  // SOFT on this path
                    const int ic = ok ? i : (i_lo < N ? i_lo : 0);
                    const double ap = s_ap[ic], lv = s_lv[ic], dr = s_dr[ic], siv = s_siv[ic];
                    double wp[Q], m0 = 0.0, mrr = 0.0, mr = 0.0, mrd[Q];
#pragma unroll
                    for (int j = 0; j < Q; ++j) { wp[j] = s_wp[size_t(j) * Np + ic]; mrd[j] = 0.0; }
                    if (ok) {
#pragma unroll 2
                        for (int g = g_lo; g < g_hi; ++g) {
                            double rp = ap;
#pragma unroll
                            for (int j = 0; j < Q; ++j) rp = fma(-wp[j], g_Bg[g * Q + j], rp);
                            const double sci = fma(-rp, rp, lv);
                            double gam = 0.0;
#pragma unroll
                            for (int j = 0; j < Q; ++j) {
                                const int o = g * Q + j;
                                const double rho = soft_rho<true>((g_thr[o] - sci) * c1q, tab_s, p.k_magic, p.k_c0, p.k_c1n, p.k_c2, p.k_c3n);
                                const double gj_ = fma(g_P[o], dr, g_Q[o]) * fma(-rho, rho, rho);
                                gam += gj_;
                                const double Bj = g_Bg[o], gB = gj_ * Bj;
                                a_t1[j] += gj_; a_t2[j] = fma(gB, Bj, a_t2[j]); a_t4[j] += gB;
#pragma unroll
                                for (int k = 0; k < Q; ++k) a_t3[j][k] = fma(gB, g_dev[g * Q + k], a_t3[j][k]);
                            }
                            const double grp = gam * rp;
                            m0 += gam; mrr = fma(grp, rp, mrr); mr += grp;
#pragma unroll
                            for (int k = 0; k < Q; ++k) mrd[k] = fma(grp, g_dev[g * Q + k], mrd[k]);
                        }
                    }
                    if (half == 1) {
                        double* x = s_x2 + th * XS;
                        x[0] = m0; x[1] = mrr; x[2] = mr;
#pragma unroll
                        for (int k = 0; k < Q; ++k) x[3 + k] = mrd[k];
                    }
                    __syncthreads();
                    if (half == 0 && i < p_hi) {
                        if (!ok) {
#pragma unroll
                            for (int j = 0; j < Q; ++j) {
                                qb.wbar[int64_t(r0 + j) * Np + i] = 0.0;
                                if (i < N) s_wp[size_t(j) * Np + i] = 0.0;
                            }
                        } else {
                            const double* x = s_x2 + th * XS;
                            m0 += x[0]; mrr += x[1]; mr += x[2];
#pragma unroll
                            for (int k = 0; k < Q; ++k) mrd[k] += x[3 + k];
                            const double iv = siv * siv;
                            const double var_bar = iv * (m0 - 2.0 * mrr);
                            const double vk = var_bar * noise * kvar;
                            const uint32_t bits = p.zb[i];
#pragma unroll
                            for (int j = 0; j < Q; ++j) {
                                double mrB = 0.0;
#pragma unroll
                                for (int k = 0; k < Q; ++k) mrB = fma(s_Sinv[j * Q + k], mrd[k], mrB);
                                double zdot = 0.0;
                                for (int s2 = 0; s2 < Sp; ++s2) zdot += ((bits >> s2) & 1u) ? s_C[j * MAX_SP + s2] : -s_C[j * MAX_SP + s2];
                                const double w_ji = wp[j] / (noise * siv);
                                qb.wbar[int64_t(r0 + j) * Np + i] = fma(vk, zdot, -2.0 * noise * siv * mrB);
                                s_wp[size_t(j) * Np + i] = vk * w_ji;                   // zdot_bar_ji (this row's table entry is dead)
                                a_bs[j] = fma(-2.0 * wp[j], mr, a_bs[j]);
#pragma unroll
                                for (int k = 0; k < Q; ++k) a_si[j][k] = fma(-2.0 * wp[j], mrd[k], a_si[j][k]);
                            }
                        }
                    }
                    __syncthreads();
                }
            } else {
            // RI train points per thread are swept together: RI Q independent sigmoid chains, and the per-column totals
            // (which depend on the grid point only) are accumulated once per grid point for all RI rows
            constexpr int RI = (Q <= 2) ? 4 : 2;
            for (int i0 = p_lo + tid; i0 < p_hi; i0 += SQ_THREADS * RI) {
                double ap[RI], lv[RI], dr[RI], siv[RI], vm[RI], wp[RI][Q], m0[RI], mrr[RI], mr[RI], mrd[RI][Q];
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    const int i = i0 + r * SQ_THREADS;
                    const bool ok = (i < i_hi) && SOFT && !bad;
                    const int ic = ok ? i : i_lo < N ? i_lo : 0;
                    vm[r] = ok ? 1.0 : 0.0;
                    ap[r] = s_ap[ic]; lv[r] = s_lv[ic]; dr[r] = s_dr[ic]; siv[r] = s_siv[ic];
                    m0[r] = mrr[r] = mr[r] = 0.0;
#pragma unroll
                    for (int j = 0; j < Q; ++j) { wp[r][j] = s_wp[size_t(j) * Np + ic]; mrd[r][j] = 0.0; }
                }
                if (SOFT && !bad && i0 < i_hi) {
                    for (int g = 0; g < G; ++g) {
                        double rp[RI], sci[RI], gam[RI];
#pragma unroll
                        for (int r = 0; r < RI; ++r) {
                            rp[r] = ap[r];
#pragma unroll
                            for (int j = 0; j < Q; ++j) rp[r] = fma(-wp[r][j], g_Bg[g * Q + j], rp[r]);
                            sci[r] = fma(-rp[r], rp[r], lv[r]);
                            gam[r] = 0.0;
                        }
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            const int o = g * Q + j;
                            const double thr_j = g_thr[o], Pj = g_P[o], Qj = g_Q[o];
                            double gsum = 0.0;
#pragma unroll
                            for (int r = 0; r < RI; ++r) {
                                const double rho = soft_rho<true>((thr_j - sci[r]) * c1q, tab_s, p.k_magic, p.k_c0, p.k_c1n, p.k_c2, p.k_c3n);
                                const double gj_ = vm[r] * fma(Pj, dr[r], Qj) * fma(-rho, rho, rho);
                                gam[r] += gj_;
                                gsum += gj_;
                            }
                            const double Bj = g_Bg[o], gB = gsum * Bj;
                            a_t1[j] += gsum; a_t2[j] = fma(gB, Bj, a_t2[j]); a_t4[j] += gB;
#pragma unroll
                            for (int k = 0; k < Q; ++k) a_t3[j][k] = fma(gB, g_dev[g * Q + k], a_t3[j][k]);
                        }
#pragma unroll
                        for (int r = 0; r < RI; ++r) {
                            const double grp = gam[r] * rp[r];
                            m0[r] += gam[r]; mrr[r] = fma(grp, rp[r], mrr[r]); mr[r] += grp;
#pragma unroll
                            for (int k = 0; k < Q; ++k) mrd[r][k] = fma(grp, g_dev[g * Q + k], mrd[r][k]);
                        }
                    }
                }
                // per-train-point epilogue
#pragma unroll
                for (int r = 0; r < RI; ++r) {
                    const int i = i0 + r * SQ_THREADS;
                    if (i >= p_hi) continue;
                    if (vm[r] == 0.0) {
#pragma unroll
                        for (int j = 0; j < Q; ++j) {
                            qb.wbar[int64_t(r0 + j) * Np + i] = 0.0;
                            if (i < N) s_wp[size_t(j) * Np + i] = 0.0;
                        }
                        continue;
                    }
                    const double iv = siv[r] * siv[r];
                    const double var_bar = iv * (m0[r] - 2.0 * mrr[r]);           // -2 iv^2 (mrr/iv) - iv (-m0)
                    const double vk = var_bar * noise * kvar;
                    const uint32_t bits = p.zb[i];
#pragma unroll
                    for (int j = 0; j < Q; ++j) {
                        double mrB = 0.0;
#pragma unroll
                        for (int k = 0; k < Q; ++k) mrB = fma(s_Sinv[j * Q + k], mrd[r][k], mrB);
                        double zdot = 0.0;
                        for (int s = 0; s < Sp; ++s) zdot += ((bits >> s) & 1u) ? s_C[j * MAX_SP + s] : -s_C[j * MAX_SP + s];
                        const double w_ji = wp[r][j] / (noise * siv[r]);
                        qb.wbar[int64_t(r0 + j) * Np + i] = fma(vk, zdot, -2.0 * noise * siv[r] * mrB);
                        s_wp[size_t(j) * Np + i] = vk * w_ji;                       // zdot_bar_ji (this row's table entry is dead)
                        a_bs[j] = fma(-2.0 * wp[r][j], mr[r], a_bs[j]);
#pragma unroll
                        for (int k = 0; k < Q; ++k) a_si[j][k] = fma(-2.0 * wp[r][j], mrd[r][k], a_si[j][k]);
                    }
                }
            }
            }
#pragma unroll
            for (int j = 0; j < Q; ++j) {
                R.put(SqbIdx::t1(j), a_t1[j]); R.put(SqbIdx::t2(j), a_t2[j]); R.put(SqbIdx::t4(j), a_t4[j]);
                R.put(SqbIdx::bsum(j), a_bs[j]);
#pragma unroll
                for (int k = 0; k < Q; ++k) { R.put(SqbIdx::t3(j, k), a_t3[j][k]); R.put(SqbIdx::sinv(j, k), a_si[j][k]); }
            }
        }
        __syncthreads();
        SQB_CLK(12);
        R.commit();
        // C_bar_js (train part) = sum_i (+-) zdot_bar_ji: thread = (pair (j, s), slice of the train points), then a fixed-order sum
        {
            const int npair = Q * Sp, nslice = SQ_THREADS / npair;               // npair <= 64
            double* cpart = wred;                                                 // [nslice][npair]; wred is free after commit
            __syncthreads();
            if (tid < nslice * npair) {
                const int pr = tid % npair, sl = tid / npair;
                const int j = pr / Sp, s2 = pr - j * Sp;
                double acc = 0.0;
                for (int i = i_lo + sl; i < i_hi; i += nslice) {
                    const double zb_ = s_wp[size_t(j) * Np + i];
                    acc += ((p.zb[i] >> s2) & 1u) ? zb_ : -zb_;
                }
                cpart[sl * npair + pr] = acc;
            }
            __syncthreads();
            if (tid < npair) {
                double acc = 0.0;
                for (int sl = 0; sl < nslice; ++sl) acc += cpart[sl * npair + tid];
                s_Cbar[(tid / Sp) * MAX_SP + (tid % Sp)] = acc;
            }
        }
        __syncthreads();
        SQB_CLK(13);
        if constexpr (CL) {        // totals and C_bar over the windows: rank 0 sums them (fixed rank order)
            auto cl = cg::this_cluster();
            cl.sync();
            if (crank == 0) {
                double tsum[(SQB_NT + SQ_THREADS - 1) / SQ_THREADS], csum = 0.0;
#pragma unroll
                for (int u = 0; u < (SQB_NT + SQ_THREADS - 1) / SQ_THREADS; ++u) {
                    const int e = tid + u * SQ_THREADS;
                    tsum[u] = 0.0;
                    if (e < SQB_NT)
                        for (int r = 0; r < SQB_CLUSTER; ++r) tsum[u] += cl.map_shared_rank(tot, r)[e];
                }
                const int cidx = (tid / Sp) * MAX_SP + (tid % Sp);
                if (tid < Q * Sp)
                    for (int r = 0; r < SQB_CLUSTER; ++r) csum += cl.map_shared_rank(s_Cbar, r)[cidx];
                __syncthreads();
#pragma unroll
                for (int u = 0; u < (SQB_NT + SQ_THREADS - 1) / SQ_THREADS; ++u) {
                    const int e = tid + u * SQ_THREADS;
                    if (e < SQB_NT) tot[e] = tsum[u];
                }
                if (tid < Q * Sp) s_Cbar[cidx] = csum;
                __syncthreads();
            }
        }

        SQB_CLK(14);
        // ---- final: q x q algebra (one thread; rank 0 of the cluster) ------------------------------------------------------------
        if (tid == 0 && crank == 0) {
            double Si[Q][Q], Lf[Q][Q], shr[Q][Q], shr_bar[Q][Q], Sinv_bar[Q][Q];
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    Si[j][k] = s_Sinv[j * Q + k]; Lf[j][k] = s_Lf[j * Q + k];
                    shr[j][k] = ((j == k) ? 1.0 : 0.0) - noise * Si[j][k];
                    shr_bar[j][k] = tot[SqbIdx::shr(j, k)];
                }
            // Cholesky reverse of Lf = chol(Cf): Cf_bar = Lf^-T Phi(Lf^T Lf_bar) Lf^-1, symmetrised
            double Lfb[Q][Q], Pm[Q][Q], Li[Q][Q], Tm[Q][Q], Cfb[Q][Q];
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) Lfb[j][k] = (k <= j) ? tot[SqbIdx::lf(j, k)] : 0.0;
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    double s = 0.0;
                    for (int m = 0; m < Q; ++m) s += Lf[m][j] * Lfb[m][k];                    // (Lf^T Lf_bar)_jk
                    Pm[j][k] = (k < j) ? s : (k == j ? 0.5 * s : 0.0);                        // Phi
                }
            for (int j = 0; j < Q; ++j)                                                       // Li = Lf^-1 (lower)
                for (int i2 = 0; i2 < Q; ++i2) {
                    if (i2 < j) Li[i2][j] = 0.0;
                    else if (i2 == j) Li[i2][j] = 1.0 / Lf[i2][i2];
                    else {
                        double s = 0.0;
                        for (int k = j; k < i2; ++k) s += Lf[i2][k] * Li[k][j];
                        Li[i2][j] = -s / Lf[i2][i2];
                    }
                }
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    double s = 0.0;
                    for (int m = 0; m < Q; ++m) s += Pm[j][m] * Li[m][k];
                    Tm[j][k] = s;
                }
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    double s = 0.0;
                    for (int m = 0; m < Q; ++m) s += Li[m][j] * Tm[m][k];
                    Cfb[j][k] = s;
                }
            // grid leak (see ScoreQBwdArgs::grid_leak).  Everything depends on the grid through dev_g = y_g - mu = Lc eps_g:
            // B_g = Sinv dev_g and the refit mean mu + shr dev_g, so  sum_g dev_bar_g dev_g^T = Sinv^T Mb + shr^T Mm  with the
            // moments Mb = sum_g B_bar_g dev_g^T, Mm = sum_g meanf_bar_g dev_g^T this reverse pass has anyway (they make
            // Sinv_bar).  With eps_g = Lc^-1 dev_g:  Lc_bar = tril((sum_g dev_bar_g dev_g^T) Lc^-T), then the Cholesky reverse.
            double leak[Q][Q];
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) leak[j][k] = 0.0;
            if (Q > 1 && qb.grid_leak && p.cscale == nullptr) {      // (also in vector-Jacobian mode: the moments are the handed-in adjoints')
                double Dm[Q][Q], Lc[Q][Q];
                for (int a = 0; a < Q; ++a)
                    for (int k = 0; k < Q; ++k) {
                        double s = 0.0;
                        for (int j = 0; j < Q; ++j) {
                            const double mb = tot[SqbIdx::sinv(j, k)] - 2.0 * ivc[j] * noise * noise * tot[SqbIdx::t3(j, k)];
                            s += Si[j][a] * mb + shr[j][a] * tot[SqbIdx::shr(j, k)];
                        }
                        Dm[a][k] = s;
                        Lc[a][k] = s_Lc[a * Q + k];
                    }
                grid_leak_from_moments<Q>(Dm, Lc, leak);
            }
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) shr_bar[j][k] += noise * 0.5 * (Cfb[j][k] + Cfb[k][j]);
            double var_c_bar[Q], Bsum[Q];
            for (int j = 0; j < Q; ++j) {
                const double lvb = tot[SqbIdx::t1(j)], ivb = -noise * noise * tot[SqbIdx::t2(j)];
                var_c_bar[j] = -2.0 * ivc[j] * ivc[j] * ivb - ivc[j] * lvb;
                Bsum[j] = tot[SqbIdx::bsum(j)] - 2.0 * ivc[j] * noise * noise * tot[SqbIdx::t4(j)];
                for (int s = 0; s < Sp; ++s) s_Cbar[j * MAX_SP + s] += -var_c_bar[j] * kvar * noise * p.z_cand[j * Sp + s];
            }
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    double cr = 0.0;
                    for (int s = 0; s < Sp; ++s) cr = fma(s_Cbar[j * MAX_SP + s], s_R[k * MAX_SP + s], cr);
                    Sinv_bar[j][k] = -noise * shr_bar[j][k] + tot[SqbIdx::sinv(j, k)]
                                     - 2.0 * ivc[j] * noise * noise * tot[SqbIdx::t3(j, k)] + cr;
                }
            // S_bar = -Sinv^T Sinv_bar Sinv^T
            double Sb[Q][Q];
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    double s = 0.0;
                    for (int m = 0; m < Q; ++m) s += Sinv_bar[j][m] * Si[k][m];               // (Sinv_bar Sinv^T)_jk
                    Tm[j][k] = s;
                }
            for (int j = 0; j < Q; ++j)
                for (int k = 0; k < Q; ++k) {
                    double s = 0.0;
                    for (int m = 0; m < Q; ++m) s += Si[m][j] * Tm[m][k];
                    Sb[j][k] = -s;
                }
            for (int j = 0; j < Q; ++j) {
                double* adj = qb.adjq + int64_t(r0 + j) * ADJQ_STRIDE;
                // mu_bar_k = mu_part_k - sum_j mu_part_j shr_jk - sum_j Sinv_jk Bsum_j   (index j of this loop plays k)
                double mb = tot[SqbIdx::mu(j)];
                for (int m = 0; m < Q; ++m) mb -= tot[SqbIdx::mu(m)] * shr[m][j] + Si[m][j] * Bsum[m];
                double drc_bar = 0.0;
                if (SOFT) {
                    double inner = 0.0;
                    for (int s = 0; s < S; ++s)
                        inner += s_kap[s * MAX_Q + j] * ((1.0 - alpha_lvl) * s_A[s * MAX_Q + j] / fmax(scw[j], 1e-6) +
                                                         alpha_lvl * s_B[s * MAX_Q + j] / fmax(snc[j], 1e-6));
                    if (ext_adj) inner = -qb.ext_ood_bar[int64_t(c) * Q + j];           // d L / d ood_j, handed in
                    const double om_bar = tot[SqbIdx::om(j)] - inner * ood[j] * (1.0 - ood[j]) * inv_temp;
                    const double nrm_bar = tot[SqbIdx::nrm(j)] - om_bar * omega[j] / norm[j];
                    drc_bar = tot[SqbIdx::drc(j)] + om_bar / norm[j] + nrm_bar;
                }
                adj[0] = bad ? nan("") : mb;
                adj[1] = bad ? nan("") : drc_bar;
                for (int k = 0; k < MAX_Q; ++k) adj[2 + k] = (k < Q) ? (bad ? nan("") : Sb[j][k] + Sb[k][j]) : 0.0;
                for (int k = 0; k < MAX_Q; ++k) adj[ADJQ_LEAK + k] = (k < Q && k != j && !bad) ? leak[j][k] + leak[k][j] : 0.0;
                for (int s = 0; s < Sp; ++s) {
                    double ub = 0.0;
                    for (int m = 0; m < Q; ++m) ub = fma(Si[m][j], s_Cbar[m * MAX_SP + s], ub);   // (Sinv^T C_bar)_js
                    s_utb[j * MAX_SP + s] = bad ? 0.0 : -ub;
                    adj[8 + s] = -ub;
                }
            }
            if (!ext_adj) p.acq[c] = tot_acq / double(S);
        }
        __syncthreads();
        SQB_CLK(15);
        if constexpr (CL) {        // ut_bar lives in rank 0's shared memory
            auto cl = cg::this_cluster();
            cl.sync();
            if (crank != 0 && tid < MAX_Q * MAX_SP) s_utb[tid] = cl.map_shared_rank(s_utb, 0)[tid];
            __syncthreads();
        }
        SQB_CLK(16);
        // fold the probe term into the right-hand sides: w_bar'_j = w_bar_j + z ut_bar_j   (T = A^-1 z)
        if (SOFT && !bad)
            for (int i = i_lo + tid; i < i_hi; i += SQ_THREADS) {
                const uint32_t bits = p.zb[i];
#pragma unroll
                for (int j = 0; j < Q; ++j) {
                    double acc = qb.wbar[int64_t(r0 + j) * Np + i];
                    for (int s = 0; s < Sp; ++s) acc += ((bits >> s) & 1u) ? s_utb[j * MAX_SP + s] : -s_utb[j * MAX_SP + s];
                    qb.wbar[int64_t(r0 + j) * Np + i] = acc;
                }
            }
        __syncthreads();
        SQB_CLK(17);
        if constexpr (CL) cg::this_cluster().sync();      // remote shared memory stays valid until every CTA is done with it
        SQB_CLK(18);
    }
}

// Last stage for joint batches: one CTA (256 threads) per candidate POINT (row r = c q + j) -- a handful of rows per call, so
// the train points of ONE row are spread over 8 warps (latency, not throughput).
//   u_bar_i = (A^-1 w_bar')_i + mu_bar alpha_i - sum_k Ssym_jk w_ki
//   x_bar   = sum_i u_bar_i k'(r_i)/r_i (x - X_i)/l^2 + sum_{k != j} Ssym_jk k'(r_jk)/r_jk (x_j - x_k)/l^2
// RBQ_THREADS = 256 is the throughput shape; calls with a handful of rows (the SGLD step, decided per CALL like split-K) use
// 1024 threads: one train point per thread at N = 1024, i.e. one round of loads instead of four in a row.
template <int RBQ_THREADS>
__global__ void __launch_bounds__(RBQ_THREADS)
k_rows_bwd_q(const double* __restrict__ Xq, int rows, int q, const double* __restrict__ XsT, int N, int Np, int d,
             const double* __restrict__ inv_ls, double outputscale, const double* __restrict__ alpha,
             const double* __restrict__ Aw, const double* __restrict__ W, const double* __restrict__ adjq,
             int adj_stride, int q1_layout, const double* __restrict__ Xb, int nb, const double* __restrict__ cv_bar,
             double* __restrict__ dX, int accumulate) {
    // Xb / cv_bar (cNEI): c_rj = k(X_b,r, x_j) - ... adds sum_r cv_bar[c, r, j] dk(X_b,r, x_j)/dx_j
    // q1_layout: the adjoint record of k_score_bwd ({mu_bar, s_bar, ...}; Ssym = 2 s_bar) instead of k_score_q_bwd's
    __shared__ double s_x[MAX_D], s_red[RBQ_THREADS / 32][8], s_adj[2 + MAX_Q];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r = blockIdx.x;
    if (r >= rows) return;
    const int c = r / q, j = r - c * q;
    if (tid < d) s_x[tid] = Xq[int64_t(r) * d + tid] * inv_ls[tid];
    if (tid < 2 + MAX_Q) {
        const double* a = adjq + int64_t(r) * adj_stride;
        s_adj[tid] = q1_layout ? (tid == 0 ? a[0] : (tid == 2 ? 2.0 * a[1] : 0.0)) : a[tid];
    }
    __syncthreads();
    const double mu_bar = s_adj[0];
    const double kfac = -(5.0 / 3.0) * outputscale;
    const double* leak = q1_layout ? nullptr : adjq + int64_t(r) * adj_stride + ADJQ_LEAK;   // prior-block adjoint through the grid
    for (int k0 = 0; k0 < d; k0 += 8) {
        double acc[8];
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) acc[kk] = 0.0;
        for (int i = tid; i < N; i += RBQ_THREADS) {
            double r2 = 0.0;
            for (int k = 0; k < d; ++k) {
                double df = s_x[k] - XsT[k * Np + i];
                r2 = fma(df, df, r2);
            }
            double ub = fma(mu_bar, alpha[i], Aw[int64_t(r) * Np + i]);
            for (int k = 0; k < q; ++k) ub = fma(-s_adj[2 + k], W[int64_t(c * q + k) * Np + i], ub);
            const double t = kSqrt5 * sqrt(fmax(r2, 1e-30));
            const double coef = ub * kfac * (1.0 + t) * exp(-t);
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
                if (k0 + kk < d) acc[kk] = fma(coef, s_x[k0 + kk] - XsT[(k0 + kk) * Np + i], acc[kk]);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            double v = warp_sum(acc[kk]);
            if (lane == 0) s_red[warp][kk] = v;
        }
        __syncthreads();
        if (tid < 8 && k0 + tid < d) {
            const int kd = k0 + tid;
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < RBQ_THREADS / 32; ++w) v += s_red[w][tid];
            // candidate-vs-candidate kernel entries (the diagonal of K_qq is a constant)
            for (int k = 0; k < q; ++k) {
                if (k == j) continue;
                double r2 = 0.0;
                for (int e = 0; e < d; ++e) {
                    double df = s_x[e] - Xq[int64_t(c * q + k) * d + e] * inv_ls[e];
                    r2 = fma(df, df, r2);
                }
                const double t = kSqrt5 * sqrt(fmax(r2, 1e-30));
                const double dfk = s_x[kd] - Xq[int64_t(c * q + k) * d + kd] * inv_ls[kd];
                v = fma((s_adj[2 + k] + (leak ? leak[k] : 0.0)) * kfac * (1.0 + t) * exp(-t), dfk, v);
            }
            for (int rb = 0; rb < nb; ++rb) {
                double r2 = 0.0;
                for (int e = 0; e < d; ++e) {
                    double df = s_x[e] - Xb[rb * d + e] * inv_ls[e];
                    r2 = fma(df, df, r2);
                }
                const double t = kSqrt5 * sqrt(fmax(r2, 1e-30));
                const double dfk = s_x[kd] - Xb[rb * d + kd] * inv_ls[kd];
                v = fma(cv_bar[(int64_t(c) * nb + rb) * q + j] * kfac * (1.0 + t) * exp(-t), dfk, v);
            }
            const double out = v * inv_ls[kd];
            dX[int64_t(r) * d + kd] = accumulate ? dX[int64_t(r) * d + kd] + out : out;   // accumulate: the ratio term is there already
        }
        __syncthreads();
    }
}

}  // namespace cbo
