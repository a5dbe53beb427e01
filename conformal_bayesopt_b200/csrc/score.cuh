// score.cuh -- nonconformity scores, density-ratio-weighted conformal rank, mask and the
// importance-weighted acquisition value.  One warp GROUP (2..16 warps) per candidate, group-local named barriers.
//
// Fuses, per candidate (q = 1, one output), what the reference evaluates as dense batched tensors
// of shape (b, G, N+1, ...) (conformalbo/helpers.py:266-325, conformalbo/acquisitions.py:57-76,
// 151-159, 88-98, 225-248), using the closed forms of SURVEY.md App. A-2..A-7:
//   phase A  per-train-point Hutchinson variance for THIS candidate (A-4):
//              var_i = noise (h0_i + noise/(Sp+1e-6) * w_i * (z_i . c)),  c = (z_c - u^T T)/s
//            and the coefficients of the score difference, which is a quadratic in b_g = (y_g - mu)/s:
//              x_i(g) = c1 (score_c(g) - score_i(g)) = (qa_i b_g + qb_i) b_g + qc_i,  c1 = 16/(temp ln 2)
//            -> pair-private shared-memory tables qa, qb, qc
//   phase B  4 grid points per pass, every train point: hard count, soft-rank sigmoid and the
//            weighted sums (A-3, A-5) -- the hot loop: 11 fp64-pipe instructions (13 with a ratio estimator) per
//            (train point, grid point), 4 independent dependency chains per lane, 16 warps per SM
//   phase C  (thread per grid point) weighted mass, randomised step, conformal mask, log-density,
//            optimistic mask, refit mean (A-5, A-6)
//   phase D  (warp reduce over the grid) importance-weighted integration (A-7)
// Groups never wait for each other (bar.sync on a group-private barrier id): while one group is in the
// latency-bound phases A/C/D the other warps of the SM sub-partition keep the fp64 pipe busy with phase B.
// Large batches use pairs (8 candidates in flight per SM); small batches widen the group (up to 16 warps on one
// candidate) so that a handful of candidates still uses whole SMs -- latency instead of throughput.
#pragma once
#include "common.cuh"
#include "rows.cuh"

namespace cbo {

#ifndef CBO_FWD_PAIRS
#define CBO_FWD_PAIRS 8
#endif
#ifndef CBO_FWD_GPW
#define CBO_FWD_GPW 4
#endif
constexpr int SCORE_MAX_PAIRS = 8;     // warp pairs per CTA of the value+reverse kernel
constexpr int SCORE_GPW = 4;           // grid points per pass (value+reverse kernel)
constexpr int SCORE_FWD_PAIRS = CBO_FWD_PAIRS;   // forward kernel: warp pairs per CTA ...
constexpr int SCORE_FWD_GPW = CBO_FWD_GPW;       // ... and grid points per pass (independent chains per lane)
constexpr int MAX_S = 16;

struct ScoreArgs {
    // geometry
    int N, Np, G, S, Sp, n_cand, tiles_n;
    // model scalars
    double noise, kappa, alpha_lvl, temp, beta_prime, best_f;
    int kind, optimistic, randomized, eps_per_candidate;
    int wpg_log2;           // log2(warps per candidate group): 1 (pairs) .. 4
    int unit_dr;            // no ratio estimator: all train-point importance weights are 1
    // sigmoid constants, passed as parameters so that they are constant-bank operands instead of per-iteration moves
    double k_magic, k_c0, k_c1n, k_c2, k_c3n;
    // score_rows_fx: the monic cubic in w = 1 + z 2^-32 (kFxB2, kFxB1, kFxB0)
    double fx_b2, fx_b1, fx_b0;
    // per-call tables (train side)
    const double* a;        // Np  noise*alpha
    const double* h0;       // Np
    const uint32_t* zb;     // Np
    const double* dr;       // Np
    const double* sum_dr;   // 1
    const double* z_cand;   // Sp  probe row of the candidate
    // per-candidate inputs
    const double* W;        // n_cand x Np
    const double* rn_partial;  // n_cand x tiles_n   partial ||v||^2
    const double* mu;       // n_cand
    const double* ut;       // n_cand x Sp   u^T T_s
    const double* drc;      // n_cand (already clamp_min 1e-6) or nullptr -> 1
    const double* cscale;   // n_cand grid covariance scale or nullptr -> 1
    const double* eps_grid; // G or n_cand x G
    const double* eps_absmax; // 1 or n_cand: max_g |eps_g| (bounds the sigmoid argument; see CLAMP)
    const double* eps_mc;   // S x G
    const double* u_rand;   // n_cand x G or nullptr
    const double* ext_values; // n_cand x S x G sample values computed by another kernel (cNEI), or nullptr
    // outputs (nullable except acq)
    double* acq; double* grid; double* logp; double* mask; double* ood; double* qscore;
    double* post_mean; double* post_var; double* num_acc; double* refit_mean; double* refit_sd;
};

// shared: tab[16] | dr[Np]        per pair: qa, qb, qc [Np] | 4 x [G] (raw cnt/sr/sw, then cw/nc/mf/mk) | zt[4][16] | pad[96]
__host__ __device__ inline size_t score_smem_shared_bytes(int Np) { return sizeof(double) * (16 + size_t(Np)); }
__host__ __device__ inline size_t score_smem_pair_bytes(int Np, int G) {
    return sizeof(double) * (3 * size_t(Np) + 4 * size_t(G) + 64 + 96);   // +96: the coefficient prefetch reads one row past N32
}
__host__ __device__ inline size_t score_smem_bytes(int Np, int G, int npairs) {
    return score_smem_shared_bytes(Np) + size_t(npairs) * score_smem_pair_bytes(Np, G);
}
// warp groups (of `wpg` warps) per CTA that fit 227 KB of shared memory and 16 warps (0: unsupported shape)
inline int score_groups_for(int Np, int G, int wpg) {
    size_t avail = 227 * 1024;
    if (score_smem_shared_bytes(Np) + score_smem_pair_bytes(Np, G) > avail) return 0;
    size_t w = (avail - score_smem_shared_bytes(Np)) / score_smem_pair_bytes(Np, G);
    size_t cap = size_t(2 * SCORE_FWD_PAIRS / wpg);
    return int(w > cap ? cap : w);
}
inline int score_pairs_for(int Np, int G) { return score_groups_for(Np, G, 2); }

// barrier over the `nthr` threads of one warp group (barrier ids 1..8; id 0 is __syncthreads)
__device__ __forceinline__ void pair_barrier(int group, int nthr) {
    asm volatile("bar.sync %0, %1;" ::"r"(group + 1), "r"(nthr) : "memory");
}
// ... that also ORs a predicate across them
__device__ __forceinline__ int pair_barrier_or(int group, int nthr, int pred) {
    int r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\tsetp.ne.s32 q, %3, 0;\n\tbar.red.or.pred p, %1, %2, q;\n\tselp.s32 %0, 1, 0, p;\n\t}"
        : "=r"(r)
        : "r"(group + 1), "r"(nthr), "r"(pred)
        : "memory");
    return r;
}

// Phase B row sweep for SCORE_GPW grid points: hard count, soft-rank sigmoid, weighted sums.
// Soft rank: rho = sigmoid(dd / temp) = 1 / (1 + e),  e = exp(-dd/temp) = 2^(-x/16),  x = c1 dd.  With
// n = -round(x) = 16 kk + j:  e = 2^kk * tab[j] * P(x - round(x)),  P(y) ~ exp(-y ln2/16) on [-1/2, 1/2] (16-entry
// table + near-minimax cubic, rel. err 1.2e-9); 1/(1+e) from a MUFU.RCP64H seed and one Newton step folded into the
// running sum: 11 fp64-pipe instructions per (train point, grid point) without a ratio estimator, 13 with one.  CLAMP = true clamps |x| at 946 = 41 * 16/ln2 in the integer domain
// (rho is 0 or 1 to 2e-18 beyond that); phase A proves |x| < 900 for most candidates and then the clamp (4 integer
// instructions per element of an issue-bound loop) is compiled out.
// UNIT = true: no ratio estimator, every train point has importance weight 1 (the reference's `ratio_estimator=None`
// branch, helpers.py:286-290): sum rho * dr == sum rho, one fp64 instruction and one shared-memory load less.
// 2^(n/16) for n = 16 kk + j from the 16-entry table of bit patterns  bits(2^(j/16)) - (j << 48):  adding n << 16 to the
// high word then yields exactly high(2^(j/16)) + (kk << 20).  Three non-fp64 instructions and one LDS: shift, LOP3
// (mask and OR with the 128-byte-aligned table address), load, LEA.
__device__ __forceinline__ double exp_tab_scaled(unsigned tab_s, int n) {
    unsigned addr = ((unsigned(n) << 3) & 0x78u) | tab_s;
    int lo, hi;
    asm("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(addr));
    return __hiloint2double(hi + (n << 16), lo);
}
__device__ __forceinline__ void exp_tab_init(double* tab, int j) {
    double v = exp2(double(j) / 16.0);
    tab[j] = __hiloint2double(__double2hiint(v) - (j << 16), __double2loint(v));
}

// exp(-t), t in [0, 700], relative error < 1e-15, from the same pattern table (rows.cuh exp_neg_fast with the table
// entry already carrying the exponent): 12 fp64 instructions instead of the ~35 of libdevice exp.
__device__ __forceinline__ double exp_neg_pat(double t, unsigned tab_s) {
    constexpr double kMagic = 6755399441055744.0;            // 1.5 * 2^52
    constexpr double kL2e16 = 23.083120654223414;            // 16 / ln 2
    constexpr double kLn2o16Hi = 0.04332169877307024;        // ln2/16, low 21 bits zeroed
    constexpr double kLn2o16Lo = 1.1926343307941173e-11;     // ln2/16 - hi
    t = fmin(t, 700.0);
    double tt = fma(-t, kL2e16, kMagic);
    int n = __double2loint(tt);                              // round(-t * 16 / ln2) <= 0
    double nf = tt - kMagic;
    double y = fma(nf, -kLn2o16Hi, -t);
    y = fma(nf, -kLn2o16Lo, y);
    double pp = fma(y, 1.0 / 5040.0, 1.0 / 720.0);
    pp = fma(pp, y, 1.0 / 120.0);
    pp = fma(pp, y, 1.0 / 24.0);
    pp = fma(pp, y, 1.0 / 6.0);
    pp = fma(pp, y, 0.5);
    pp = fma(pp, y, 1.0);
    pp = pp * y;
    double tj = exp_tab_scaled(tab_s, n);
    return fma(tj, pp, tj);
}
// sigmoid(z) to ~2 ulp (per-grid-point epilogue: conformal mask)
__device__ __forceinline__ double sigmoid_pat(double z, unsigned tab_s) {
    double e = exp_neg_pat(fabs(z), tab_s);
    double r = rcp_full(1.0 + e);
    return (z >= 0.0) ? r : e * r;
}

template <bool SOFT, bool CLAMP, bool UNIT, int GPW>
__device__ __forceinline__ void score_rows(const ScoreArgs& p, const double* __restrict__ sq, const double* __restrict__ sdr,
                                           unsigned tab_s, int lane, int N32,
                                           const double (&bg)[GPW], double (&sr)[GPW], double (&sw)[GPW],
                                           unsigned (&cneg)[GPW]) {
    const double kMagic = p.k_magic;                          // 1.5 * 2^52
    const double kC0 = p.k_c0, kC1n = p.k_c1n, kC2 = p.k_c2, kC3n = p.k_c3n;
    constexpr int kSatHi = 0x408d9000;                        // hiword(946.0)
    const double* tq = sq + 3 * lane;
    const double* td = sdr + lane;
    double qa = tq[0], qb = tq[1], qc = tq[2], dr = UNIT ? 1.0 : td[0];
    for (int i = lane; i < N32; i += 32) {
        tq += 96; td += 32;                                   // next row (reads past N32 stay inside shared memory)
        const double qa_n = tq[0], qb_n = tq[1], qc_n = tq[2], dr_n = UNIT ? 1.0 : td[0];
        double x[GPW];
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
            x[k] = fma(fma(qa, bg[k], qb), bg[k], qc);
            cneg[k] += unsigned(__double2hiint(x[k])) >> 31;  // counts score_c < score_i
        }
        if (SOFT) {
            double t[GPW], yn[GPW], e[GPW];
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                double xc = x[k];
                if (CLAMP) {
                    int hi = __double2hiint(x[k]);
                    int ch = min(hi & 0x7fffffff, kSatHi) | (hi & 0x80000000);
                    xc = __hiloint2double(ch, __double2loint(x[k]));         // clamp(x, -946, 946)
                }
                t[k] = kMagic - xc;                           // low word = -round(x)
                yn[k] = xc + (t[k] - kMagic);                 // x - round(x) in [-1/2, 1/2]
            }
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                double pl = fma(yn[k], kC3n, kC2);
                pl = fma(pl, yn[k], kC1n);
                pl = fma(pl, yn[k], kC0);                     // exp(-yn ln2/16)
                double ts = exp_tab_scaled(tab_s, __double2loint(t[k]));   // 2^(j/16) 2^kk
                e[k] = fma(ts, pl, 1.0);                      // 1 + e: the product and the sum are ONE fma
            }
#pragma unroll
            for (int k = 0; k < GPW; ++k) {
                double d = e[k];
                double y0;
                asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
                double err1 = fma(-d, y0, 2.0);               // Newton: rho = y0 (2 - d y0), folded into the running sum
                if (UNIT) {
                    sr[k] = fma(y0, err1, sr[k]);
                } else {
                    double rho = y0 * err1;
                    sr[k] += rho;
                    sw[k] = fma(rho, dr, sw[k]);
                }
            }
        } else if (!UNIT) {
#pragma unroll
            for (int k = 0; k < GPW; ++k) sw[k] += (__double2hiint(x[k]) >= 0) ? dr : 0.0;
        }
        qa = qa_n; qb = qb_n; qc = qc_n; dr = dr_n;
    }
}

// ---- score_rows_fx: the soft-rank sweep with the exponent split folded into the quadratic -------------------------------
// e = exp(-dd/temp) = 2^(X/16), X = -x.  The coefficient table holds (qa, qb, M - qc) with M = 1.5 * 2^20, so that the second
// fma of the quadratic already delivers t = M + X rounded to 32 fraction bits: the low word of t IS z = X - floor(X) as a
// 32-bit fraction, the high word carries floor(X) = 16 kk + j in its low bits (what exp_tab_scaled wants) and [x <= 0] in
// bit 19 (the hard count).  Two integer instructions turn z into the double w = 1 + z 2^-32 in [1, 2), and
// 2^(z/16) = c3 (((w + b2) w + b1) w + b0) (monic minimax cubic on [1, 2], rel. err 1.2e-9 like the cubic of score_rows; monic, so that every
// step has ONE constant operand -- a DFMA that reads three fresh register pairs issues every 3 clocks instead of 2,
// tools/dfma_operands*.cu; c3 folded into
// the table) -- the three DADDs that split x into round(x) and x - round(x) are gone: 8 fp64-pipe instructions per element
// instead of 11 (10 instead of 13 with a ratio estimator).  Measured (tools/score_loop_bench.cu VAR 50 against VAR 0): 112.9
// instead of 121.0 clocks per (32 train points x 4 grid points).  The same cubic on the integer pipes (three IMAD.HI in 2^-36
// fixed point, 5 fp64 instructions per element) was measured too and is SLOWER (132 clocks: the loop is bound by issue slots
// and register-file reads, not by the fp64 pipe; profiles/r02_score_loop_fx.log).
// Rounding X to 2^-32 moves a score difference by < 2e-13 * temp: far inside the tolerance of a rank tie; an exact tie
// x == 0 counts as "score_c < score_i" here and as ">=" in score_rows (measure zero for real data).
// Valid while |x| < 946 (the caller's `big` flag; the CLAMP variant clamps t first, two integer instructions).
constexpr double kFxMagic = 1572864.0;                                   // 1.5 * 2^20
constexpr double kFxC3 = 1.384801912537204e-05;
constexpr double kFxB2 = 64.74963271891839, kFxB1 = 2995.871356125411, kFxB0 = 69150.87307076537;
__device__ __forceinline__ void exp_tabfx_init(double* tab, int j) {
    double v = kFxC3 * exp2(double(j) / 16.0);
    tab[j] = __hiloint2double(__double2hiint(v) - (j << 16), __double2loint(v));
}

template <bool CLAMP, bool UNIT, int GPW>
__device__ __forceinline__ void score_rows_fx(const ScoreArgs& p, const double* __restrict__ sq, const double* __restrict__ sdr,
                                              unsigned tab_s, int lane, int N32,
                                              const double (&bg)[GPW], double (&sr)[GPW], double (&sw)[GPW],
                                              unsigned (&cneg)[GPW]) {
    const double kB2 = p.fx_b2, kB1 = p.fx_b1, kB0 = p.fx_b0;  // kernel parameters: constant-bank operands
    const double* tq = sq + 3 * lane;
    const double* td = sdr + lane;
    double qa = tq[0], qb = tq[1], qc = tq[2], dr = UNIT ? 1.0 : td[0];
    const int nsteps = N32 >> 5;
    for (int it = 0; it < nsteps; ++it) {
        tq += 96; td += 32;                                   // next row (reads past N32 stay inside shared memory)
        const double qa_n = tq[0], qb_n = tq[1], qc_n = tq[2], dr_n = UNIT ? 1.0 : td[0];
        double t[GPW], e[GPW];
#pragma unroll
        for (int k = 0; k < GPW; ++k) t[k] = fma(qa, bg[k], qb);
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
            t[k] = fma(-t[k], bg[k], qc);                     // M - x
            if (CLAMP) {   // clamp(t, M - 946, M + 947) on the high word: positive doubles order like their bit patterns, and a
                           // negative t (x > M) has a negative high word; the low word adds < 1 to |x|
                constexpr int kLo = 0x4137fc4e, kHi = 0x413803b2;   // high words of M - 946, M + 946
                const int h = min(max(__double2hiint(t[k]), kLo), kHi);
                t[k] = __hiloint2double(h, __double2loint(t[k]));
            }
        }
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
            const unsigned z = unsigned(__double2loint(t[k]));
            const int hi = __double2hiint(t[k]);
            cneg[k] += unsigned(hi) >> 19;                    // 0x826 + [x <= 0]; the constant is removed after the loop
            const double w = __hiloint2double(int(0x3ff00000u + (z >> 12)), int(z << 20));   // 1 + z 2^-32
            double pl = w + kB2;
            pl = fma(pl, w, kB1);
            pl = fma(pl, w, kB0);                             // 2^(z/16) / c3
            const double ts = exp_tab_scaled(tab_s, hi);      // c3 2^(j/16) 2^kk
            e[k] = fma(ts, pl, 1.0);                          // 1 + e
        }
#pragma unroll
        for (int k = 0; k < GPW; ++k) {
            double d = e[k];
            double y0;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(d));
            double err1 = fma(-d, y0, 2.0);                   // Newton: rho = y0 (2 - d y0), folded into the running sum
            if (UNIT) {
                sr[k] = fma(y0, err1, sr[k]);
            } else {
                double rho = y0 * err1;
                sr[k] += rho;
                sw[k] = fma(rho, dr, sw[k]);
            }
        }
        qa = qa_n; qb = qb_n; qc = qc_n; dr = dr_n;
    }
#pragma unroll
    for (int k = 0; k < GPW; ++k) cneg[k] -= unsigned(nsteps) * 0x826u;   // high word of t >> 19 = 0x826 + [x <= 0]
}

template <bool SOFT>
__global__ void __launch_bounds__(SCORE_FWD_PAIRS * 64, 1)
k_score(const ScoreArgs p) {
    extern __shared__ __align__(16) double sm[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wl = p.wpg_log2, GT = 32 << wl;                            // threads per candidate group
    const int pair = warp >> wl, half = warp & ((1 << wl) - 1), ptid = tid & (GT - 1);   // group, warp in group, thread in group
    const int npairs = blockDim.x >> (5 + wl);
    const int N = p.N, Np = p.Np, G = p.G, Sp = p.Sp;
    const int N32 = (N + 31) & ~31;                     // Np is a multiple of 128 >= N32

    __shared__ __align__(128) double tab[16];           // 2^(j/16): one 128-byte row, conflict-free for any index pattern;
                                                        // static, so that its address is an immediate of the table load
    double* sdr = sm + 16;                             // density ratio of train point i (0 beyond N)
    double* pbase = sdr + Np + size_t(pair) * (3 * size_t(Np) + 4 * size_t(G) + 64 + 96);
    double* sq = pbase;                                 // [i][3] = (qa_i, qb_i, qc_i): 24-byte stride, conflict-free LDS.64
    double* g_cw = sq + 3 * size_t(Np);                            // phase B stores the hard count here, phase C the conf. weight
    double* g_nc = g_cw + G;                            // sum rho        -> non-conformal weight
    double* g_mf = g_nc + G;                            // sum rho * dr   -> refit mean
    double* g_mk = g_mf + G;                            // mask
    double* zt = g_mk + G;                              // zt[t][e] = sum_{b<4} (+-) c_{4t+b}, sign from bit b of e

    // ---- once per CTA: candidate-independent tables ------------------------------------------------
    __shared__ __align__(128) double tabfx[16];         // c3 2^(j/16): score_rows_fx (soft path)
    if (tid < 16) { exp_tab_init(tab, tid); exp_tabfx_init(tabfx, tid); }
    const unsigned tab_s = smem_u32(tab), tabfx_s = smem_u32(tabfx);
    for (int i = tid; i < Np; i += blockDim.x) sdr[i] = (i < N) ? p.dr[i] : 0.0;
    __syncthreads();                                    // the only block-wide barrier

    const double sum_dr = p.sum_dr[0];
    const double kvar = p.noise / (double(Sp) + 1e-6);
    const double inv_temp = SOFT ? 1.0 / p.temp : 0.0;
    const double c1 = SOFT ? inv_temp * 23.083120654223414 : 1.0;        // 16 / (temp ln 2)

    const int total_pairs = gridDim.x * npairs;
    for (int c = blockIdx.x * npairs + pair; c < p.n_cand; c += total_pairs) {
        // ---- phase A (64 threads) ------------------------------------------------------------------
        {   // pull the next candidate's W row towards L2 while this one is processed
            int cn = c + total_pairs;
            if (cn < p.n_cand)
                for (int o = ptid * 16; o < Np; o += GT * 16)
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p.W + int64_t(cn) * Np + o));
        }
        double vnorm = 0.0;
        for (int t = 0; t < p.tiles_n; ++t) vnorm += p.rn_partial[int64_t(c) * p.tiles_n + t];
        const double s_var = p.kappa + p.noise - vnorm;                 // Schur complement (noisy predictive variance)
        const double mu = p.mu[c];
        const double inv_s = rcp_full(s_var);
        const double* Wc = p.W + int64_t(c) * Np;
        double zcc = 0.0;                                               // sum_s z_cs c_s
        for (int s = 0; s < Sp; ++s) {
            double zcs = p.z_cand[s];
            zcc = fma(zcs, (zcs - p.ut[int64_t(c) * Sp + s]) * inv_s, zcc);
        }
        if (ptid < 64) {   // nibble tables of z_i . c (z_i in {+-1}^Sp is stored as a bit mask): one entry per thread
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
        const double var_c = fma(kvar, double(Sp) - p.noise * zcc, p.noise);   // noise (1 + (Sp - noise zcc)/(Sp + 1e-6))
        int bad = !(var_c > 0.0) || !(s_var > 0.0);
        const double iv_c = 0.5 * rcp_full(var_c);
        const double lv_c = -0.5 * log_fast(var_c) - 0.5 * kLog2Pi;
        // score_c(g) = lv_c - iv_c (noise b_g)^2 ;  score_i(g) = lv_i - iv_i (a_i - noise w_i b_g)^2
        const double qa_c = -iv_c * p.noise * p.noise;
        // bmax >= |b_g| for every grid point of this candidate: the sigmoid argument is bounded by (|qa| bmax + |qb|) bmax + |qc|
        const double cscale = p.cscale ? p.cscale[c] : 1.0;
        const double gvar = cscale * s_var;
        const double gsd = sqrt_fast(gvar);                             // sd of the grid distribution
        const double lp_c = -0.5 * (log_fast(gvar) + kLog2Pi);          // log-density of the grid distribution at its mean
        const double dens_c = 0.3989422804014327 * rcp_full(gsd);       // exp(lp_c) = 1 / sqrt(2 pi gvar)
        const double bmax = gsd * p.eps_absmax[p.eps_per_candidate ? c : 0] * inv_s;
        int big = 0;
        pair_barrier(pair, GT);                                          // nibble tables visible
        for (int i0 = ptid; i0 < N32; i0 += 4 * GT) {                    // 4 independent loads in flight per thread
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
                    double qa = 0.0, qb = 0.0, qc = -900.0;              // padding: x = -900: never counted, rho ~ 1e-17, dr = 0
                    if (i < N) {
                        double zdot = zt[bits[j] & 15] + zt[16 + ((bits[j] >> 4) & 15)];
                        if (Sp > 8) zdot += zt[32 + ((bits[j] >> 8) & 15)] + zt[48 + ((bits[j] >> 12) & 15)];
                        double var = p.noise * fma(kvar * w[j], zdot, h0[j]);
                        bool okv = var > 0.0;
                        bad |= !okv;
                        var = okv ? var : 1.0;                           // flagged: the candidate's outputs become NaN
                        double iv = 0.5 * rcp_full(var);
                        double lv = -0.5 * log_fast(var) - 0.5 * kLog2Pi;
                        double wi = p.noise * w[j];
                        qa = c1 * fma(iv * wi, wi, qa_c);                // b_g^2 coefficient
                        qb = c1 * (-2.0 * iv * ai[j] * wi);              // b_g coefficient
                        qc = c1 * (fma(iv * ai[j], ai[j], lv_c - lv));   // constant
                        big |= !(fma(fma(fabs(qa), bmax, fabs(qb)), bmax, fabs(qc)) < 900.0);
                    }
                    sq[3 * i] = qa; sq[3 * i + 1] = qb; sq[3 * i + 2] = SOFT ? kFxMagic - qc : qc;   // soft path: score_rows_fx wants M - qc
                }
            }
        }
        const double drc = p.drc ? p.drc[c] : 1.0;
        const double norm = sum_dr + drc;
        const double inv_norm = rcp_full(norm);
        const double omega_c = drc / norm;                               // true division: the hard path compares it with alpha
        bad = pair_barrier_or(pair, GT, bad);                             // tables complete; any non-positive variance?
        big = SOFT ? pair_barrier_or(pair, GT, big) : 0;                  // can any sigmoid argument leave the unclamped range?

        // ---- phase B: 4 grid points per pass, passes alternate between the two warps ----------------
        for (int g0 = half * SCORE_FWD_GPW; g0 < G; g0 += (SCORE_FWD_GPW << wl)) {
            double bg[SCORE_FWD_GPW], sr[SCORE_FWD_GPW], sw[SCORE_FWD_GPW];
            unsigned cneg[SCORE_FWD_GPW];
#pragma unroll
            for (int k = 0; k < SCORE_FWD_GPW; ++k) {
                int g = min(g0 + k, G - 1);
                double eps = p.eps_grid[p.eps_per_candidate ? int64_t(c) * G + g : g];
                bg[k] = gsd * eps * inv_s;                              // (y_g - mu) / s
                sr[k] = 0.0; sw[k] = 0.0; cneg[k] = 0u;
            }
            if (!bad) {   // (no ratio estimator: padded rows carry rho ~ 1e-17 each into sum rho * dr, dr is not read: negligible)
                if (SOFT) {
                    if (p.unit_dr) {
                        if (big) score_rows_fx<true, true, SCORE_FWD_GPW>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                        else score_rows_fx<false, true, SCORE_FWD_GPW>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                    } else {
                        if (big) score_rows_fx<true, false, SCORE_FWD_GPW>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                        else score_rows_fx<false, false, SCORE_FWD_GPW>(p, sq, sdr, tabfx_s, lane, N32, bg, sr, sw, cneg);
                    }
                } else {
                    if (p.unit_dr) score_rows<false, false, true, SCORE_FWD_GPW>(p, sq, sdr, tab_s, lane, N32, bg, sr, sw, cneg);
                    else score_rows<false, false, false, SCORE_FWD_GPW>(p, sq, sdr, tab_s, lane, N32, bg, sr, sw, cneg);
                }
            }
            if (SCORE_FWD_GPW == 4) {   // transposed reductions: lanes 8k..8k+7 end up with the totals of grid point g0 + k
                int ci[4];
                double srr[4], swr[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) { ci[k] = int(cneg[k]); srr[k] = sr[k]; swr[k] = sw[k]; }
                const int k = lane >> 3;
                const int cnt = N32 - warp_sum4(ci, lane);                    // #{i <= N : score_c >= score_i}
                const double tr = SOFT ? warp_sum4(srr, lane) : 0.0;
                const double tw = p.unit_dr ? (SOFT ? tr : double(cnt)) : warp_sum4(swr, lane);
                if ((lane & 7) == 0 && g0 + k < G) { g_cw[g0 + k] = double(cnt); g_nc[g0 + k] = tr; g_mf[g0 + k] = tw; }
            } else {
#pragma unroll
                for (int k = 0; k < SCORE_FWD_GPW; ++k) {
                    int cnt = N32 - warp_sum_int(int(cneg[k]));
                    double tr = SOFT ? warp_sum(sr[k]) : 0.0;
                    double tw = p.unit_dr ? (SOFT ? tr : double(cnt)) : warp_sum(sw[k]);
                    if (lane == 0 && g0 + k < G) { g_cw[g0 + k] = double(cnt); g_nc[g0 + k] = tr; g_mf[g0 + k] = tw; }
                }
            }
        }
        pair_barrier(pair, GT);                                           // raw sums of every grid point are in shared memory

        // ---- phase C: thread per grid point -----------------------------------------------------------
        for (int g = ptid; g < G; g += GT) {
            const double my_cnt = g_cw[g], my_sr = g_nc[g], my_sw = g_mf[g];
            double n_g = my_cnt + 1.0;                                  // includes the candidate itself
            double M, rho_self;
            if (SOFT) {
                double scale = n_g * rcp_full(my_sr + 0.5);             // self: sigmoid(0) = 1/2
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
            double dens = dens_c * exp_neg_pat(0.5 * eps * eps, tab_s);  // exp(lp)
            double opt = p.optimistic ? ((yg >= mu) ? 1.0 : 0.0) : 1.0;
            double shrink = 1.0 - p.noise * inv_s;
            double rc = p.noise * (gsd * eps * inv_s);
            double scc = fma(-rc * rc, iv_c, lv_c);                     // candidate's own conformal score
            double cw = mk * opt * rcp_full(fmax(dens, 1e-6)), nc = (1.0 - mk) * opt, mf = mu + (yg - mu) * shrink;
            if (bad) { cw = nc = mf = mk = lp = scc = nan(""); }
            g_mk[g] = mk; g_cw[g] = cw; g_nc[g] = nc; g_mf[g] = mf;
            if (p.grid) p.grid[int64_t(c) * G + g] = yg;
            if (p.logp) p.logp[int64_t(c) * G + g] = lp;
            if (p.mask) p.mask[int64_t(c) * G + g] = mk;
            if (p.qscore) p.qscore[int64_t(c) * G + g] = scc;
            if (p.refit_mean) p.refit_mean[int64_t(c) * G + g] = mf;
        }
        pair_barrier(pair, GT);

        // ---- phase D (first warp of the pair): integrate over the grid ---------------------------------
        if (half == 0) {
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
                    double cw = g_cw[g], nc = g_nc[g], mf = g_mf[g];
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
                    A = fma(cw, val, A);
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
        // the next candidate's phase A touches only the tables; its barriers order the g-arrays
    }
}

}  // namespace cbo
