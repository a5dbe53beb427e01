// Rank-1 extension of a built model state when an observation is appended (second half): O(N^2) per point
// instead of the O(N^3) rebuild.  With k = k(X, x), v = L^-1 k, l^2 = kappa + noise - |v|^2, w = L^-T v:
//   new row of L^-1:  [-w^T / l, 1 / l]          alpha <- [alpha - w e / l^2 ; e / l^2],  e = (y - m) - k . alpha
// Three launches per point, every reduction in a fixed order (bit-reproducible).
#pragma once
#include "common.cuh"

namespace cbo {

constexpr int MA_THREADS = 256;

// kvec[i] = k(X_i, x), part[blk] = sum over the block's rows of kvec[i] alpha[i]
__global__ void __launch_bounds__(MA_THREADS)
k_append_kvec(const double* __restrict__ X, int n, int d, const double* __restrict__ inv_ls, double outputscale,
              const double* __restrict__ x_new, const double* __restrict__ alpha, double* __restrict__ kvec,
              double* __restrict__ part) {
    __shared__ double red[MA_THREADS / 32];
    const int i = blockIdx.x * MA_THREADS + threadIdx.x;
    double ka = 0.0;
    if (i < n) {
        double r2 = 0.0;
        for (int e = 0; e < d; ++e) {
            const double df = (X[int64_t(i) * d + e] - x_new[e]) * inv_ls[e];
            r2 = fma(df, df, r2);
        }
        const double t = kSqrt5 * sqrt(fmax(r2, 1e-30));
        const double kv = outputscale * (1.0 + t + t * t * (1.0 / 3.0)) * exp(-t);
        kvec[i] = kv;
        ka = kv * alpha[i];
    }
    ka = warp_sum(ka);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ka;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < MA_THREADS / 32; ++w) s += red[w];
        part[blockIdx.x] = s;
    }
}

// v[i] = sum_{j <= i} linv[i][j] kvec[j]   (warp per row)
__global__ void __launch_bounds__(MA_THREADS)
k_append_v(const double* __restrict__ linv, int Np, int n, const double* __restrict__ kvec, double* __restrict__ v) {
    const int i = blockIdx.x * (MA_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= n) return;
    const double* row = linv + int64_t(i) * Np;
    double acc = 0.0;
    for (int j = lane; j <= i; j += 32) acc = fma(row[j], kvec[j], acc);
    acc = warp_sum(acc);
    if (lane == 0) v[i] = acc;
}

// w[j] = sum_{i >= j} linv_t[j][i] v[i] (warp per column); writes row n of L^-1 / column n of L^-T, updates alpha, stores x.
__global__ void __launch_bounds__(MA_THREADS)
k_append_row(double* __restrict__ linv, double* __restrict__ linv_t, int Np, int n, int d, const double* __restrict__ v,
             const double* __restrict__ part, int nparts, double kappa_noise, const double* __restrict__ y_new, double mean,
             const double* __restrict__ x_new, double* __restrict__ X, double* __restrict__ alpha, int32_t* __restrict__ info) {
    __shared__ double red[MA_THREADS / 32];
    __shared__ double s_l2, s_e;
    // |v|^2 in a fixed order, identically in every CTA
    double vv = 0.0;
    for (int i = threadIdx.x; i < n; i += MA_THREADS) vv = fma(v[i], v[i], vv);
    vv = warp_sum(vv);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = vv;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < MA_THREADS / 32; ++w) s += red[w];
        double ka = 0.0;
        for (int b = 0; b < nparts; ++b) ka += part[b];
        s_l2 = kappa_noise - s;
        s_e = (*y_new - mean) - ka;
    }
    __syncthreads();
    const double l2 = s_l2, e = s_e;
    const bool ok = l2 > 0.0;
    const double inv_l_exact = ok ? 1.0 / sqrt(l2) : __longlong_as_double(0x7ff8000000000000ll);
    const int j = blockIdx.x * (MA_THREADS / 32) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (j < n) {
        const double* col = linv_t + int64_t(j) * Np;
        double acc = 0.0;
        for (int i = j + lane; i < n; i += 32) acc = fma(col[i], v[i], acc);
        acc = warp_sum(acc);
        if (lane == 0) {
            const double r = -acc * inv_l_exact;
            linv[int64_t(n) * Np + j] = r;
            linv_t[int64_t(j) * Np + n] = r;
            alpha[j] -= acc * (e / l2);
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        linv[int64_t(n) * Np + n] = inv_l_exact;
        linv_t[int64_t(n) * Np + n] = inv_l_exact;
        alpha[n] = e / l2;
        if (!ok && info && *info == 0) *info = n + 1;
    }
    if (blockIdx.x == 0 && threadIdx.x < d) X[int64_t(n) * d + threadIdx.x] = x_new[threadIdx.x];
}

inline size_t model_append_ws_doubles(int64_t Np) { return size_t(2 * Np + (Np + MA_THREADS - 1) / MA_THREADS + 8); }

inline cudaError_t model_append(double* X, int64_t N, int64_t d, int64_t Np, const double* inv_ls, double mean,
                                double outputscale, double noise, double* alpha, double* linv, double* linv_t,
                                const double* x_new, const double* y_new, int64_t q, int32_t* info, double* ws,
                                cudaStream_t s) {
    double* kvec = ws;
    double* v = ws + Np;
    double* part = v + Np;
    for (int64_t p = 0; p < q; ++p) {
        const int n = int(N + p);
        const double* xn = x_new + p * d;
        const int nblk = (n + MA_THREADS - 1) / MA_THREADS, nwb = (n + MA_THREADS / 32 - 1) / (MA_THREADS / 32);
        if (n > 0) {
            CBO_LAUNCHED(), k_append_kvec<<<nblk, MA_THREADS, 0, s>>>(X, n, (int)d, inv_ls, outputscale, xn, alpha, kvec, part);
            CBO_LAUNCHED(), k_append_v<<<nwb, MA_THREADS, 0, s>>>(linv, (int)Np, n, kvec, v);
        }
        CBO_LAUNCHED(), k_append_row<<<nwb > 0 ? nwb : 1, MA_THREADS, 0, s>>>(linv, linv_t, (int)Np, n, (int)d, v, part, n > 0 ? nblk : 0,
                                                            outputscale + noise, y_new + p, mean, xn, X, alpha, info);
    }
    return cudaGetLastError();
}

}  // namespace cbo
