// score_ws.cuh -- warp-specialised throughput variant of k_score (q = 1, many candidates).
//
// Same per-candidate arithmetic as k_score (score.cuh; conformalbo/helpers.py:266-325, conformalbo/acquisitions.py:57-76,
// 151-159, 88-98, 225-248), different division of labour inside the CTA:
//   * 16 COMPUTE warps = 4 candidate groups x 4 warps (one warp of every group on each SM sub-partition) that never
//     leave phase B -- the (train point x grid point) soft-rank sweep, the only phase that can fill the fp64 pipe;
//   * 4 HELPER warps (one per group, one per sub-partition) that build the NEXT candidate's coefficient tables and grid
//     offsets (phase A: global loads, log, reciprocals -- latency bound) and run the per-grid epilogue and the integration
//     of the PREVIOUS candidate (phases C/D) meanwhile.
// Tables, grid offsets and raw sums are double-buffered per group; producer/consumer hand-over through shared-memory
// mbarriers (full[buf]: helper -> compute, count 1; done[buf]: compute -> helper, count 4).  In k_score every warp pair
// walks A -> B -> C -> D itself, so at any time a few of the 4 warps of a sub-partition sit in the latency-bound phases
// and the fp64 pipe idles (65 % busy over the kernel against 81 % inside the loop); here a sub-partition
// always has 4 warps in the loop and a fifth that only waits on memory.
#pragma once
#include "score.cuh"

namespace cbo {

constexpr int SWS_GROUPS = 4;                                  // candidates in flight per CTA
constexpr int SWS_CW = 4;                                      // compute warps per candidate (3 groups x 8 warps at 72 registers
                                                               // was measured: 4.73 ms against 3.84 ms per 65536 candidates)
constexpr int SWS_COMPUTE_WARPS = SWS_GROUPS * SWS_CW;
constexpr int SWS_THREADS = (SWS_COMPUTE_WARPS + SWS_GROUPS) * 32;
constexpr int SWS_NSCAL = 16;                                  // per-candidate scalars handed from phase A to phases C/D

// per group: sq[2][3 Np] | raw[2][4][Gp] | bgs[2][Gp] | zt[64] | scal[2][16] | flags[2][2] (int) | full[2], done[2] (mbarrier)
// (the coefficient prefetch of score_rows reads 96 doubles past a table: the next table or the arrays behind it)
__host__ __device__ inline size_t score_ws_gp(int G) { return (size_t(G) + 3) & ~size_t(3); }
__host__ __device__ inline size_t score_ws_group_doubles(int Np, int G) {
    return 6 * size_t(Np) + 10 * score_ws_gp(G) + 64 + 2 * SWS_NSCAL + 2 + 4;
}
__host__ __device__ inline size_t score_ws_smem_bytes(int Np, int G) {
    return sizeof(double) * (size_t(Np) + SWS_GROUPS * score_ws_group_doubles(Np, G));
}
inline bool score_ws_fits(int Np, int G) { return score_ws_smem_bytes(Np, G) + 256 <= 227 * 1024 && 10 * score_ws_gp(G) + 72 >= 96; }

#ifdef CBO_WS_PROFILE
__device__ unsigned long long g_ws_prof[8];   // [0] compute: clk waiting on full, [1] compute: clk total, [2] helper: clk waiting on done, [3] helper total, [4] helper phase A, [5] helper C/D
#define WS_CLK() clock64()
#define WS_ADD(i, v) do { if (lane == 0) atomicAdd(&g_ws_prof[i], (unsigned long long)(v)); } while (0)
#else
#define WS_CLK() 0ll
#define WS_ADD(i, v) do { } while (0)
#endif

template <bool SOFT>
__global__ void __launch_bounds__(SWS_THREADS, 1)
k_score_ws(const ScoreArgs p) {
    extern __shared__ __align__(16) double sm[];
    __shared__ __align__(128) double tab[16];           // exp_neg_pat / sigmoid_pat of the helper warps (phases C / D)
    __shared__ __align__(128) double tabfx[16];         // score_rows_fx of the compute warps
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool helper = warp >= SWS_COMPUTE_WARPS;
    const int grp = helper ? warp - SWS_COMPUTE_WARPS : warp >> 2;
    const int cw = warp & 3;
    const int N = p.N, Np = p.Np, G = p.G, Sp = p.Sp;
    const int Gp = int(score_ws_gp(G));
    const int N32 = (N + 31) & ~31;

    double* sdr = sm;
    double* gbase = sm + Np + size_t(grp) * score_ws_group_doubles(Np, G);
    double* sq0 = gbase;
    double* raw0 = gbase + 6 * size_t(Np);
    double* bgs0 = raw0 + 8 * Gp;
    double* zt = bgs0 + 2 * Gp;
    double* scal0 = zt + 64;
    int* flags = reinterpret_cast<int*>(scal0 + 2 * SWS_NSCAL);
    uint64_t* full = reinterpret_cast<uint64_t*>(scal0 + 2 * SWS_NSCAL + 2);
    uint64_t* done = full + 2;

    if (tid < 16) { exp_tab_init(tab, tid); exp_tabfx_init(tabfx, tid); }
    const unsigned tab_s = smem_u32(tab), tabfx_s = smem_u32(tabfx);
    for (int i = tid; i < Np; i += blockDim.x) sdr[i] = (i < N) ? p.dr[i] : 0.0;
    if (helper && lane == 0) {
        mbar_init(&full[0], 1); mbar_init(&full[1], 1);
        mbar_init(&done[0], SWS_CW); mbar_init(&done[1], SWS_CW);
    }
    __syncthreads();                                    // the only block-wide barrier

    const int total_groups = gridDim.x * SWS_GROUPS;
    const int first = blockIdx.x * SWS_GROUPS + grp;
    const int n_mine = first < p.n_cand ? (p.n_cand - first + total_groups - 1) / total_groups : 0;

    if (!helper) {
        // =================================== compute warps: phase B only ========================================
        const long long t_start = WS_CLK();
        long long t_wait = 0;
        for (int k = 0; k < n_mine; ++k) {
            const int buf = k & 1;
            const long long t0 = WS_CLK();
            mbar_wait(&full[buf], (k >> 1) & 1);
            t_wait += WS_CLK() - t0;
            const double* sq = sq0 + size_t(buf) * 3 * Np;
            double* raw = raw0 + buf * 4 * Gp;
            const double* bgs = bgs0 + buf * Gp;
            const int bad = flags[2 * buf], big = flags[2 * buf + 1];
            for (int g0 = cw * 4; g0 < G; g0 += 4 * SWS_CW) {
                double bg[4], sr[4], sw[4];
                unsigned cneg[4];
                {
                    const double2 b01 = *reinterpret_cast<const double2*>(bgs + g0);
                    const double2 b23 = *reinterpret_cast<const double2*>(bgs + g0 + 2);
                    bg[0] = b01.x; bg[1] = b01.y; bg[2] = b23.x; bg[3] = b23.y;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) { sr[j] = 0.0; sw[j] = 0.0; cneg[j] = 0u; }
                if (!bad) {
                    if (SOFT) {          // tables hold (qa, qb, M - qc): score_rows_fx
                        if (p.unit_dr) {
                            if (big) score_rows_fx<true, true, 4>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                            else score_rows_fx<false, true, 4>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                        } else {
                            if (big) score_rows_fx<true, false, 4>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                            else score_rows_fx<false, false, 4>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                        }
                    } else {             // hard path: the sign of x only
                        if (p.unit_dr) score_rows<false, false, true, 4>(p, sq, sdr, tab_s, lane, N32, bg, sr, sw, cneg);
                        else score_rows<false, false, false, 4>(p, sq, sdr, tab_s, lane, N32, bg, sr, sw, cneg);
                    }
                }
                int ci[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) ci[j] = int(cneg[j]);
                const int j = lane >> 3;
                const int cnt = N32 - warp_sum4(ci, lane);
                const double tr = SOFT ? warp_sum4(sr, lane) : 0.0;
                const double tw = p.unit_dr ? (SOFT ? tr : double(cnt)) : warp_sum4(sw, lane);
                if ((lane & 7) == 0 && g0 + j < G) { raw[g0 + j] = double(cnt); raw[Gp + g0 + j] = tr; raw[2 * Gp + g0 + j] = tw; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&done[buf]);
        }
        WS_ADD(0, t_wait); WS_ADD(1, WS_CLK() - t_start);
        return;
    }

    // ======================================= helper warp: phases A, C, D ============================================
    const double sum_dr = p.sum_dr[0];
    const double kvar = p.noise / (double(Sp) + 1e-6);
    const double inv_temp = SOFT ? 1.0 / p.temp : 0.0;
    const double c1 = SOFT ? inv_temp * 23.083120654223414 : 1.0;        // 16 / (temp ln 2)

    const long long h_start = WS_CLK();
    long long h_wait = 0, h_a = 0, h_cd = 0;
    for (int k = 0; k <= n_mine; ++k) {
        const long long ta = WS_CLK();
        if (k < n_mine) {
            // ---- phase A of candidate k into buffer k & 1 (last read by candidate k - 2, whose done[] was awaited below) ----
            const int buf = k & 1;
            const int c = first + k * total_groups;
            double* sq = sq0 + size_t(buf) * 3 * Np;
            {
                const int cn = c + total_groups;
                if (cn < p.n_cand)
                    for (int o = lane * 16; o < Np; o += 32 * 16)
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(p.W + int64_t(cn) * Np + o));
            }
            double vnorm = 0.0;
            for (int t = 0; t < p.tiles_n; ++t) vnorm += p.rn_partial[int64_t(c) * p.tiles_n + t];
            const double s_var = p.kappa + p.noise - vnorm;
            const double mu = p.mu[c];
            const double inv_s = rcp_full(s_var);
            const double* Wc = p.W + int64_t(c) * Np;
            double zcc = 0.0;
            for (int s = 0; s < Sp; ++s) {
                double zcs = p.z_cand[s];
                zcc = fma(zcs, (zcs - p.ut[int64_t(c) * Sp + s]) * inv_s, zcc);
            }
            for (int e0 = lane; e0 < 64; e0 += 32) {
                const int t4 = (e0 >> 4) * 4, e = e0 & 15;
                double acc = 0.0;
#pragma unroll
                for (int bb = 0; bb < 4; ++bb) {
                    int s = t4 + bb;
                    if (s < Sp) {
                        double csv = (p.z_cand[s] - p.ut[int64_t(c) * Sp + s]) * inv_s;
                        acc += ((e >> bb) & 1) ? csv : -csv;
                    }
                }
                zt[e0] = acc;
            }
            const double var_c = fma(kvar, double(Sp) - p.noise * zcc, p.noise);
            int bad = !(var_c > 0.0) || !(s_var > 0.0);
            const double iv_c = 0.5 * rcp_full(var_c);
            const double lv_c = -0.5 * log_fast(var_c) - 0.5 * kLog2Pi;
            const double qa_c = -iv_c * p.noise * p.noise;
            const double cscale = p.cscale ? p.cscale[c] : 1.0;
            const double gvar = cscale * s_var;
            const double gsd = sqrt_fast(gvar);
            const double lp_c = -0.5 * (log_fast(gvar) + kLog2Pi);
            const double dens_c = 0.3989422804014327 * rcp_full(gsd);
            const double bmax = gsd * p.eps_absmax[p.eps_per_candidate ? c : 0] * inv_s;
            int big = 0;
            __syncwarp();                                                    // nibble tables visible
            for (int i0 = lane; i0 < N32; i0 += 4 * 32) {
                double w[4], h0[4], ai[4];
                uint32_t bits[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int i = i0 + 32 * j;
                    bool ok = i < N;
                    w[j] = ok ? Wc[i] : 0.0;
                    h0[j] = ok ? p.h0[i] : 1.0;
                    ai[j] = ok ? p.a[i] : 0.0;
                    bits[j] = ok ? p.zb[i] : 0u;
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    int i = i0 + 32 * j;
                    if (i < N32) {
                        double qa = 0.0, qb = 0.0, qc = -900.0;
                        if (i < N) {
                            double zdot = zt[bits[j] & 15] + zt[16 + ((bits[j] >> 4) & 15)];
                            if (Sp > 8) zdot += zt[32 + ((bits[j] >> 8) & 15)] + zt[48 + ((bits[j] >> 12) & 15)];
                            double var = p.noise * fma(kvar * w[j], zdot, h0[j]);
                            bool okv = var > 0.0;
                            bad |= !okv;
                            var = okv ? var : 1.0;
                            double iv = 0.5 * rcp_full(var);
                            double lv = -0.5 * log_fast(var) - 0.5 * kLog2Pi;
                            double wi = p.noise * w[j];
                            qa = c1 * fma(iv * wi, wi, qa_c);
                            qb = c1 * (-2.0 * iv * ai[j] * wi);
                            qc = c1 * (fma(iv * ai[j], ai[j], lv_c - lv));
                            big |= !(fma(fma(fabs(qa), bmax, fabs(qb)), bmax, fabs(qc)) < 900.0);
                        }
                        sq[3 * i] = qa; sq[3 * i + 1] = qb; sq[3 * i + 2] = SOFT ? kFxMagic - qc : qc;   // score_rows_fx wants M - qc
                    }
                }
            }
            {   // grid offsets b_g = (y_g - mu) / s of this candidate (padding repeats the last grid point)
                double* bgs = bgs0 + buf * Gp;
                for (int g = lane; g < Gp; g += 32) {
                    const int gg = min(g, G - 1);
                    const double eps = p.eps_grid[p.eps_per_candidate ? int64_t(c) * G + gg : gg];
                    bgs[g] = gsd * eps * inv_s;
                }
            }
            bad = __any_sync(0xffffffffu, bad);
            big = SOFT ? __any_sync(0xffffffffu, big) : 0;
            if (lane == 0) {
                const double drc = p.drc ? p.drc[c] : 1.0;
                double* sc = scal0 + buf * SWS_NSCAL;
                sc[0] = mu; sc[1] = gsd; sc[2] = inv_s; sc[3] = lp_c; sc[4] = dens_c; sc[5] = iv_c; sc[6] = lv_c;
                sc[7] = drc; sc[8] = s_var;
                flags[2 * buf] = bad; flags[2 * buf + 1] = big;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[buf]);
        }
        h_a += WS_CLK() - ta;
        if (k == 0) continue;

        // ---- phases C / D of candidate k - 1 ------------------------------------------------------------------------
        const int kp = k - 1, buf = kp & 1;
        const int c = first + kp * total_groups;
        {
            const long long t0 = WS_CLK();
            mbar_wait(&done[buf], (kp >> 1) & 1);
            h_wait += WS_CLK() - t0;
        }
        const long long tc = WS_CLK();
        double* raw = raw0 + buf * 4 * Gp;
        double* g_cw = raw, *g_nc = raw + Gp, *g_mf = raw + 2 * Gp, *g_mk = raw + 3 * Gp;
        const double* sc = scal0 + buf * SWS_NSCAL;
        const double mu = sc[0], gsd = sc[1], inv_s = sc[2], lp_c = sc[3], dens_c = sc[4], iv_c = sc[5], lv_c = sc[6];
        const double drc = sc[7], s_var = sc[8];
        const int bad = flags[2 * buf];
        const double norm = sum_dr + drc;
        const double inv_norm = rcp_full(norm);
        const double omega_c = drc / norm;
        for (int g = lane; g < G; g += 32) {
            const double my_cnt = g_cw[g], my_sr = g_nc[g], my_sw = g_mf[g];
            double n_g = my_cnt + 1.0;
            double M, rho_self;
            if (SOFT) {
                double scale = n_g * rcp_full(my_sr + 0.5);
                rho_self = 0.5 * scale;
                M = scale * (my_sw + 0.5 * drc) * inv_norm;
            } else {
                rho_self = 1.0;
                M = (my_sw + drc) / norm;
            }
            if (p.randomized) {
                double D = fmax(rho_self * omega_c, 1e-6);
                double pr = fmin(fmax((M - p.alpha_lvl) / D, 0.0), 1.0);
                double u = p.u_rand[int64_t(c) * G + g];
                M -= (u < pr) ? 0.0 : D;
            }
            double mk = SOFT ? sigmoid_pat((M - p.alpha_lvl) * inv_temp, tab_s) : ((M > p.alpha_lvl) ? 1.0 : 0.0);
            double eps = p.eps_grid[p.eps_per_candidate ? int64_t(c) * G + g : g];
            double yg = mu + gsd * eps;
            double lp = fma(-0.5 * eps, eps, lp_c);
            double dens = dens_c * exp_neg_pat(0.5 * eps * eps, tab_s);
            double opt = p.optimistic ? ((yg >= mu) ? 1.0 : 0.0) : 1.0;
            double shrink = 1.0 - p.noise * inv_s;
            double rc = p.noise * (gsd * eps * inv_s);
            double scc = fma(-rc * rc, iv_c, lv_c);
            double cwt = mk * opt * rcp_full(fmax(dens, 1e-6)), nc = (1.0 - mk) * opt, mf = mu + (yg - mu) * shrink;
            if (bad) { cwt = nc = mf = mk = lp = scc = nan(""); }
            g_mk[g] = mk; g_cw[g] = cwt; g_nc[g] = nc; g_mf[g] = mf;
            if (p.grid) p.grid[int64_t(c) * G + g] = yg;
            if (p.logp) p.logp[int64_t(c) * G + g] = lp;
            if (p.mask) p.mask[int64_t(c) * G + g] = mk;
            if (p.qscore) p.qscore[int64_t(c) * G + g] = scc;
            if (p.refit_mean) p.refit_mean[int64_t(c) * G + g] = mf;
        }
        __syncwarp();
        {
            const double sd_f = sqrt(fmax(p.noise * (1.0 - p.noise * inv_s), 0.0));
            const int S = p.S;
            double scw = 0.0, snc = 0.0, nacc = 0.0;
            for (int g = lane; g < G; g += 32) {
                scw += g_cw[g]; snc += g_nc[g]; nacc += (g_mk[g] >= 0.5) ? 1.0 : 0.0;
            }
            scw = warp_sum(scw); snc = warp_sum(snc); nacc = warp_sum(nacc);
            const double ood = SOFT ? sigmoid_acc((omega_c - p.alpha_lvl) * inv_temp)
                                    : ((omega_c >= p.alpha_lvl) ? 1.0 : 0.0);
            double tot = 0.0;
            for (int s = 0; s < S; ++s) {
                double A = 0.0, B = 0.0;
                for (int g = lane; g < G; g += 32) {
                    double cwt = g_cw[g], nc = g_nc[g], mf = g_mf[g];
                    double f = fma(sd_f, p.eps_mc[int64_t(s) * G + g], mf);
                    double val;
                    if (p.ext_values) {
                        val = p.ext_values[(int64_t(c) * S + s) * G + g];
                    } else if (p.kind == 0) {
                        if (S > 1) {
                            double em = 0.0;
                            for (int s2 = 0; s2 < S; ++s2) em += p.eps_mc[int64_t(s2) * G + g];
                            double m = fma(sd_f, em / double(S), mf);
                            val = m + p.beta_prime * fabs(f - m);
                        } else {
                            val = f;
                        }
                    } else {
                        double xx = f - p.best_f;
                        if (p.temp > 0.0) { double xt = xx / p.temp; val = p.temp * (fmax(xt, 0.0) + log1p(exp(-fabs(xt)))); }
                        else val = fmax(xx, 0.0);
                    }
                    A = fma(cwt, val, A);
                    B = fma(nc, val, B);
                }
                A = warp_sum(A); B = warp_sum(B);
                tot += (1.0 - ood) * ((1.0 - p.alpha_lvl) * A / fmax(scw, 1e-6) + p.alpha_lvl * B / fmax(snc, 1e-6));
            }
            if (lane == 0) {
                p.acq[c] = tot / double(S);
                if (p.ood) p.ood[c] = ood;
                if (p.post_mean) p.post_mean[c] = mu;
                if (p.post_var) p.post_var[c] = s_var;
                if (p.num_acc) p.num_acc[c] = nacc;
                if (p.refit_sd) p.refit_sd[c] = sd_f;
            }
        }
        __syncwarp();                                   // the fin arrays are free again before the next phase A hands the buffer over
        h_cd += WS_CLK() - tc;
    }
    WS_ADD(2, h_wait); WS_ADD(3, WS_CLK() - h_start); WS_ADD(4, h_a); WS_ADD(5, h_cd);
}

}  // namespace cbo
