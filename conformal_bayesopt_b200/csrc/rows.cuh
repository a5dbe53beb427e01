// rows.cuh -- kernel rows of the candidates against the training set, plus the per-call tables.
//
//   k_kernel_rows    U[c, i] = os * (1 + sqrt5 r + 5/3 r^2) exp(-sqrt5 r),  r = ||(x_c - X_i) / l||
//                    mu[c]   = mean + sum_i U[c,i] alpha_i                   (helpers.py:119, acquisitions.py:59)
//                    ut[c,s] = sum_i U[c,i] T[i,s]                           (probe correction, App. A-4)
//   k_probe_matvec   out[i, s] = sum_j Mat[i, j] in[j, s]   (model builder: alpha = L^-T (L^-1 (y - m)))
//   k_train_tables   per-train-point tables for the score kernel             (helpers.py:396-405)
#pragma once
#include "common.cuh"

namespace cbo {

constexpr int ROWS_CPW = 4;          // candidates per warp
constexpr int ROWS_WARPS = 16;
constexpr int MAX_D = 64;
constexpr int MAX_SP = 16;
constexpr int MAX_H = 64;

// sqrt(v) for normal v > 0 to ~1 ulp: MUFU.RSQ64H seed + Goldschmidt step + residual correction (7 fp64 ops).
__device__ __forceinline__ double sqrt_fast(double v) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(v));
    double g = v * y, h = 0.5 * y;
    double r = fma(-g, h, 0.5);
    g = fma(g, r, g);
    h = fma(h, r, h);
    double dd = fma(-g, g, v);
    return fma(dd, h, g);
}

// exp(-t) for t in [0, 700], relative error < 1e-15: 16-entry 2^(j/16) table (conflict-free) + degree-7 Taylor on
// |y| <= ln2/32, two-constant Cody-Waite reduction.  12 fp64 ops instead of the ~35 of libdevice exp.
__device__ __forceinline__ double exp_neg_fast(double t, const double* __restrict__ tab) {
    constexpr double kMagic = 6755399441055744.0;            // 1.5 * 2^52
    constexpr double kL2e16 = 23.083120654223414;            // 16 / ln 2
    constexpr double kLn2o16Hi = 0.04332169877307024;        // ln2/16, low 21 bits zeroed
    constexpr double kLn2o16Lo = 1.1926343307941173e-11;     // ln2/16 - hi
    double tt = fma(-t, kL2e16, kMagic);
    int n = __double2loint(tt);                              // round(-t * 16 / ln2) <= 0
    double nf = tt - kMagic;
    double y = fma(nf, -kLn2o16Hi, -t);
    y = fma(nf, -kLn2o16Lo, y);
    double p = fma(y, 1.0 / 5040.0, 1.0 / 720.0);
    p = fma(p, y, 1.0 / 120.0);
    p = fma(p, y, 1.0 / 24.0);
    p = fma(p, y, 1.0 / 6.0);
    p = fma(p, y, 0.5);
    p = fma(p, y, 1.0);
    p = p * y;
    double tj = tab[n & 15];
    double e = fma(tj, p, tj);
    int hi = __double2hiint(e) + ((n >> 4) << 20);
    e = __hiloint2double(hi, __double2loint(e));
    return (n < -16000) ? 0.0 : e;                           // below 2^-1000: flush
}

// scaled, transposed training inputs: XsT[k][i] = X[i][k] / lengthscale[k]   (d x Np, zeros beyond N)
__global__ void k_scale_transpose(const double* __restrict__ X, int N, int Np, int d, const double* __restrict__ inv_ls,
                                  double* __restrict__ XsT) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= d * Np) return;
    int k = e / Np, i = e - k * Np;
    XsT[e] = (i < N) ? X[int64_t(i) * d + k] * inv_ls[k] : 0.0;
}

// Persistent CTAs of 16 warps; the scaled, transposed training inputs XsT (d x Np), alpha and the probe solves T^T
// (Sp x Np) are staged in shared memory once per CTA (STAGE) so that the sweep never waits on L2.  One warp owns
// ROWS_CPW candidates at a time and sweeps i = lane, lane+32, ...: X_i is read once per sweep step (conflict-free,
// lane-contiguous) and reused for the ROWS_CPW candidates; U stores are 256-byte coalesced streaming stores.
// ~41 + NSP fp64 ops per kernel value (distance 2d, sqrt 7, exp 12, Matern polynomial 4, mu 1, u^T T NSP).
template <bool STAGE, int NSP>
__global__ void __launch_bounds__(ROWS_WARPS * 32, 1)
k_kernel_rows(const double* __restrict__ Xq, int n, const double* __restrict__ XsT_g, int N, int Np, int d,
              const double* __restrict__ inv_ls, double outputscale, double mean,
              const double* __restrict__ alpha_g, const double* __restrict__ T, int Sp,
              double* __restrict__ U, double* __restrict__ mu, double* __restrict__ ut) {
    extern __shared__ double sx[];  // tab[16] | [WARPS][CPW][d] coords | (STAGE) alpha[Np] | XsT[d][Np] | TT[Sp][Np]
    double* tab = sx;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* myx = sx + 16 + warp * ROWS_CPW * d;
    double* s_alpha = sx + 16 + ROWS_WARPS * ROWS_CPW * d;
    double* s_xst = s_alpha + Np;
    double* s_tt = s_xst + size_t(d) * Np;
    const bool want_ut = (T != nullptr) && (ut != nullptr);
    if (threadIdx.x < 16) tab[threadIdx.x] = exp2(double(threadIdx.x) / 16.0);
    if (STAGE) {
        for (int e = threadIdx.x; e < Np; e += blockDim.x) s_alpha[e] = alpha_g[e];
        for (int e = threadIdx.x; e < d * Np; e += blockDim.x) s_xst[e] = XsT_g[e];
        if (want_ut)
            for (int e = threadIdx.x; e < Sp * Np; e += blockDim.x) s_tt[e] = T[e];      // T is stored transposed: (Sp x Np)
    }
    __syncthreads();
    const double* __restrict__ XsT = STAGE ? s_xst : XsT_g;
    const double* __restrict__ alpha = STAGE ? s_alpha : alpha_g;
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
        double acc_mu[ROWS_CPW], acc_ut[ROWS_CPW][NSP];
#pragma unroll
        for (int cc = 0; cc < ROWS_CPW; ++cc) {
            acc_mu[cc] = 0.0;
#pragma unroll
            for (int s = 0; s < NSP; ++s) acc_ut[cc][s] = 0.0;
        }
        for (int i = lane; i < Np; i += 32) {
            double kv[ROWS_CPW];
            if (i < N) {
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
                double al = alpha[i];
#pragma unroll
                for (int cc = 0; cc < ROWS_CPW; ++cc) {
                    double t = kSqrt5 * sqrt_fast(fmax(r2[cc], 1e-30));      // sqrt5 * r
                    double pm = fma(fma(t, 1.0 / 3.0, 1.0), t, 1.0);          // 1 + sqrt5 r + 5/3 r^2
                    kv[cc] = (outputscale * exp_neg_fast(t, tab)) * pm;
                    acc_mu[cc] = fma(kv[cc], al, acc_mu[cc]);
                }
                if (want_ut) {
#pragma unroll
                    for (int s = 0; s < NSP; ++s) {
                        if (s < Sp) {
                            double ts = STAGE ? s_tt[s * Np + i] : T[int64_t(s) * Np + i];
#pragma unroll
                            for (int cc = 0; cc < ROWS_CPW; ++cc) acc_ut[cc][s] = fma(kv[cc], ts, acc_ut[cc][s]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int cc = 0; cc < ROWS_CPW; ++cc) kv[cc] = 0.0;
            }
#pragma unroll
            for (int cc = 0; cc < ROWS_CPW; ++cc)
                if (c0 + cc < n) __stcs(U + int64_t(c0 + cc) * Np + i, kv[cc]);
        }
#pragma unroll
        for (int cc = 0; cc < ROWS_CPW; ++cc) {
            double m = warp_sum(acc_mu[cc]);
            if (lane == 0 && c0 + cc < n && mu != nullptr) mu[c0 + cc] = mean + m;
            if (want_ut) {
#pragma unroll
                for (int s = 0; s < NSP; ++s) {
                    if (s < Sp) {
                        double v = warp_sum(acc_ut[cc][s]);
                        if (lane == 0 && c0 + cc < n) ut[int64_t(c0 + cc) * Sp + s] = v;
                    }
                }
            }
        }
    }
}

// A handful of candidate points per call (the SGLD regime): one CTA (256 threads) per candidate point, the train points
// spread over its 8 warps (latency instead of throughput; no staging).  Same outputs as k_kernel_rows.
constexpr int ROWS_SMALL_THREADS = 256;
__global__ void __launch_bounds__(ROWS_SMALL_THREADS)
k_kernel_rows_small(const double* __restrict__ Xq, int n, const double* __restrict__ XsT, int N, int Np, int d,
                    const double* __restrict__ inv_ls, double outputscale, double mean, const double* __restrict__ alpha,
                    const double* __restrict__ T, int Sp, double* __restrict__ U, double* __restrict__ mu,
                    double* __restrict__ ut,
                    double* __restrict__ U2 = nullptr) {
    __shared__ double s_x[MAX_D], s_tab[16], s_red[ROWS_SMALL_THREADS / 32][1 + MAX_SP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x;
    if (tid < 16) s_tab[tid] = exp2(double(tid) / 16.0);
    if (tid < d) s_x[tid] = Xq[int64_t(c) * d + tid] * inv_ls[tid];
    __syncthreads();
    const bool want_ut = (T != nullptr) && (ut != nullptr);
    double acc_mu = 0.0, acc_ut[MAX_SP];
#pragma unroll
    for (int s = 0; s < MAX_SP; ++s) acc_ut[s] = 0.0;
    for (int i = tid; i < Np; i += ROWS_SMALL_THREADS) {
        double kv = 0.0;
        if (i < N) {
            double r2 = 0.0;
            for (int k = 0; k < d; ++k) {
                double df = s_x[k] - XsT[k * Np + i];
                r2 = fma(df, df, r2);
            }
            double t = kSqrt5 * sqrt_fast(fmax(r2, 1e-30));
            double pm = fma(fma(t, 1.0 / 3.0, 1.0), t, 1.0);
            kv = (outputscale * exp_neg_fast(t, s_tab)) * pm;
            acc_mu = fma(kv, alpha[i], acc_mu);
            if (want_ut) {
#pragma unroll
                for (int s = 0; s < MAX_SP; ++s)
                    if (s < Sp) acc_ut[s] = fma(kv, T[int64_t(s) * Np + i], acc_ut[s]);
            }
        }
        U[int64_t(c) * Np + i] = kv;
        if (U2) U2[int64_t(c) * Np + i] = kv;          // second copy for k_ut_small (pass 2 overwrites U while it runs)
    }
    acc_mu = warp_sum(acc_mu);
    if (lane == 0) s_red[warp][0] = acc_mu;
    if (want_ut) {
#pragma unroll
        for (int s = 0; s < MAX_SP; ++s)
            if (s < Sp) {
                double v = warp_sum(acc_ut[s]);
                if (lane == 0) s_red[warp][1 + s] = v;
            }
    }
    __syncthreads();
    if (tid <= (want_ut ? Sp : 0)) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < ROWS_SMALL_THREADS / 32; ++w) v += s_red[w][tid];
        if (tid == 0) { if (mu != nullptr) mu[c] = mean + v; }
        else ut[int64_t(c) * Sp + tid - 1] = v;
    }
}

// ut[c, s] = sum_i U[c, i] T[s, i] from the STORED kernel rows: the probe correction of k_kernel_rows_small as a kernel of
// its own (latency path: the rows are computed while T = A^-1 z is still being solved for on another stream).  Same
// per-thread order, warp reduction and warp order as the fused accumulation, so both give the same bits.
__global__ void __launch_bounds__(ROWS_SMALL_THREADS)
k_ut_small(const double* __restrict__ U, int N, int Np, const double* __restrict__ T, int Sp, double* __restrict__ ut) {
    __shared__ double s_red[ROWS_SMALL_THREADS / 32][MAX_SP];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int c = blockIdx.x;
    double acc_ut[MAX_SP];
#pragma unroll
    for (int s = 0; s < MAX_SP; ++s) acc_ut[s] = 0.0;
    // four train points per thread with all their loads in flight at once, accumulated in the order of the plain loop
    constexpr int UNR = 4;
    for (int i0 = tid; i0 < Np; i0 += ROWS_SMALL_THREADS * UNR) {
        double kv[UNR], tv[UNR][MAX_SP];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int i = i0 + u * ROWS_SMALL_THREADS;
            const bool ok = i < N;
            kv[u] = ok ? U[int64_t(c) * Np + i] : 0.0;
#pragma unroll
            for (int s = 0; s < MAX_SP; ++s) tv[u][s] = (ok && s < Sp) ? T[int64_t(s) * Np + i] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u)
            if (i0 + u * ROWS_SMALL_THREADS < N) {
#pragma unroll
                for (int s = 0; s < MAX_SP; ++s)
                    if (s < Sp) acc_ut[s] = fma(kv[u], tv[u][s], acc_ut[s]);
            }
    }
#pragma unroll
    for (int s = 0; s < MAX_SP; ++s) acc_ut[s] = warp_sum(acc_ut[s]);      // (independent chains; slots >= Sp are zeros)
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < MAX_SP; ++s)
            if (s < Sp) s_red[warp][s] = acc_ut[s];
    }
    __syncthreads();
    if (tid < Sp) {
        double v = 0.0;
#pragma unroll
        for (int w = 0; w < ROWS_SMALL_THREADS / 32; ++w) v += s_red[w][tid];
        ut[int64_t(c) * Sp + tid] = v;
    }
}

inline size_t rows_smem_bytes(int Np, int d, int Sp, bool stage) {
    return sizeof(double) * (16 + size_t(ROWS_WARPS) * ROWS_CPW * d + (stage ? size_t(Np) * (d + 1 + Sp) : 0));
}

template <bool STAGE, int NSP>
inline cudaError_t launch_kernel_rows_t(const double* Xq, int n, const double* XsT, int N, int Np, int d,
                                        const double* inv_ls, double outputscale, double mean, const double* alpha,
                                        const double* T, int Sp, double* U, double* mu, double* ut, int num_sms,
                                        cudaStream_t st) {
    const size_t smem = rows_smem_bytes(Np, d, (T && ut) ? Sp : 0, STAGE);
    static SmemOptIn optin;
    if (cudaError_t e = optin.ensure(k_kernel_rows<STAGE, NSP>, smem); e != cudaSuccess) return e;
    const int groups = (n + ROWS_CPW - 1) / ROWS_CPW;
    int grid = (groups + ROWS_WARPS - 1) / ROWS_WARPS;
    if (grid > num_sms) grid = num_sms;                              // persistent: one 16-warp CTA per SM
    CBO_LAUNCHED(), k_kernel_rows<STAGE, NSP><<<grid, ROWS_WARPS * 32, smem, st>>>(Xq, n, XsT, N, Np, d, inv_ls, outputscale, mean, alpha,
                                                                   T, Sp, U, mu, ut);
    return cudaGetLastError();
}

inline cudaError_t launch_kernel_rows(const double* Xq, int n, const double* XsT, int N, int Np, int d,
                                      const double* inv_ls, double outputscale, double mean, const double* alpha,
                                      const double* T, int Sp, double* U, double* mu, double* ut, int num_sms,
                                      cudaStream_t st, bool small_call = false, double* U2 = nullptr) {
    if (small_call) {   // decided per CALL by the API (<= 32 candidate points), never per chunk
        CBO_LAUNCHED(), k_kernel_rows_small<<<n, ROWS_SMALL_THREADS, 0, st>>>(Xq, n, XsT, N, Np, d, inv_ls, outputscale, mean, alpha, T, Sp, U,
                                                              mu, ut, U2);
        return cudaGetLastError();
    }
    const bool stage = rows_smem_bytes(Np, d, (T && ut) ? Sp : 0, true) <= 220 * 1024;
#define CBO_ROWS_DISPATCH(ST, NS) \
    return launch_kernel_rows_t<ST, NS>(Xq, n, XsT, N, Np, d, inv_ls, outputscale, mean, alpha, T, Sp, U, mu, ut, num_sms, st)
    if (stage) { if (Sp <= 8) CBO_ROWS_DISPATCH(true, 8); else CBO_ROWS_DISPATCH(true, 16); }
    else       { if (Sp <= 8) CBO_ROWS_DISPATCH(false, 8); else CBO_ROWS_DISPATCH(false, 16); }
#undef CBO_ROWS_DISPATCH
}

__global__ void k_fill(double* __restrict__ out, int n, double v) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = v;
}

// out[i, s] = sum_j Mat[i, j] * in[j, s]  for s < Sp.  One warp per row i; Mat is Np x Np.
__global__ void __launch_bounds__(256)
k_probe_matvec(const double* __restrict__ Mat, int Np, const double* __restrict__ in, int Sp,
               double* __restrict__ out, int j_begin_is_row, int j_end_is_row) {
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= Np) return;
    // triangular structure: lower matrix -> j <= i; upper -> j >= i
    int jb = j_begin_is_row ? (i & ~31) : 0;
    int je = j_end_is_row ? min(Np, (i | 31) + 1) : Np;
    double acc[MAX_SP];
#pragma unroll
    for (int s = 0; s < MAX_SP; ++s) acc[s] = 0.0;
    for (int j = jb + lane; j < je; j += 32) {
        double m = Mat[int64_t(i) * Np + j];
#pragma unroll
        for (int s = 0; s < MAX_SP; ++s)
            if (s < Sp) acc[s] = fma(m, in[int64_t(j) * Sp + s], acc[s]);
    }
#pragma unroll
    for (int s = 0; s < MAX_SP; ++s) {
        if (s < Sp) {
            double v = warp_sum(acc[s]);
            if (lane == 0) out[int64_t(i) * Sp + s] = v;
        }
    }
}

// out[c] = max_g |eps[c, g]|  (one warp per row; rows = 1 for a shared grid)
__global__ void k_eps_absmax(const double* __restrict__ eps, int rows, int G, double* __restrict__ out) {
    int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    double m = 0.0;
    for (int g = lane; g < G; g += 32) m = fmax(m, fabs(eps[int64_t(r) * G + g]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) out[r] = m;
}

// z (N+q) x Sp of +-1 -> padded TRANSPOSED copy zt (Sp x Np, zeros beyond N): the Sp probe vectors become Sp rows that go
// through the same two triangular GEMM passes as the candidates' kernel rows (T^T = Z^T A^-1)
__global__ void k_pad_probes(const double* __restrict__ z, int N, int Np, int Sp, double* __restrict__ zt) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= Np * Sp) return;
    int s = e / Np, i = e - s * Np;
    zt[e] = (i < N) ? z[int64_t(i) * Sp + s] : 0.0;
}

// Per-train-point tables of the score kernel:
//   a[i]   = noise * alpha_i
//   h0[i]  = 1 + (Sp - noise * sum_s z_is T_is) / (Sp + 1e-6)       (g- and candidate-independent part of var/noise)
//   zb[i]  = bit s set <=> z_is > 0
// and sum_dr = sum_i dr_i (single block, deterministic tree).
__global__ void __launch_bounds__(256)
k_train_tables(const double* __restrict__ alpha, const double* __restrict__ z, const double* __restrict__ T, int N,
               int Np, int Sp, double noise, const double* __restrict__ dr, double* __restrict__ a,
               double* __restrict__ h0, uint32_t* __restrict__ zb, double* __restrict__ sum_dr) {
    __shared__ double red[256];
    double part = 0.0;
    for (int i = threadIdx.x; i < Np; i += blockDim.x) {
        if (i < N) {
            double acc = 0.0;
            uint32_t bits = 0;
            for (int s = 0; s < Sp; ++s) {
                double zs = z[int64_t(i) * Sp + s];
                acc = fma(zs, T[int64_t(s) * Np + i], acc);
                bits |= (zs > 0.0) ? (1u << s) : 0u;
            }
            a[i] = noise * alpha[i];
            h0[i] = 1.0 + (double(Sp) - noise * acc) / (double(Sp) + 1e-6);
            zb[i] = bits;
            part += dr[i];
        } else {
            a[i] = 0.0;
            h0[i] = 1.0;
            zb[i] = 0;
        }
    }
    red[threadIdx.x] = part;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) sum_dr[0] = red[0];
}

}  // namespace cbo
