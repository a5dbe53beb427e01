// model_build.cuh -- device-side builder of the model state the hot path consumes: from raw training
// data and hyper-parameters to  L^-1 (explicit inverse Cholesky factor of A = K + noise I), its transpose and
// alpha = A^-1 (y - m).  The reference gets the same quantities from GPyTorch after every refit
// (conformalbo/utils.py:382-390 builds SingleTaskGP; prediction_strategy.mean_cache / lik_train_train_covar.cholesky(),
// root_inv_decomposition for the fantasy update); here nothing outside libcbo is needed.
//
// Blocked (128 = the DMMA GEMM's tile) left-looking Cholesky followed by a blocked triangular inverse; all O(N^3) work runs
// in k_gemm_tn_f64 (fp64 tensor path), the 128 x 128 diagonal blocks are factorised and inverted by one CTA in shared memory:
//   step j:  S    = A_jj - L_j,0:j L_j,0:j^T                      (GEMM, K = 128 j)
//            L_jj = chol(S), T_j = L_jj^-1                         (k_potrf_trtri_128)
//            L_i,j = (A_ij - L_i,0:j L_j,0:j^T) T_j^T   for i > j  (GEMM, K = 128 j; GEMM, K = 128)
//   inverse: Linv_ii = T_i,  Linv_i,0:i = -T_i (L_i,0:i Linv_0:i,0:i), built transposed (linv_t) so that every product is the
//            kernel's "TN" form  C[m,n] = sum_k A[m,k] B[n,k].
// Padding rows/columns (N .. Np-1) carry an identity block during the factorisation and are zeroed at the end.
#pragma once
#include "common.cuh"
#include "gemm_tri.cuh"
#include "rows.cuh"

namespace cbo {

constexpr int MB_NB = 128;
constexpr int MB_LDS = MB_NB + 1;            // padded shared-memory row stride (conflict-free column walks)

// A[i][i] = outputscale + noise (i < N), 1 (padding); rows >= N are zero elsewhere (buffer pre-zeroed)
__global__ void k_fix_diag(double* __restrict__ A, int N, int Np, double diag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Np) A[int64_t(i) * Np + i] = (i < N) ? diag : 1.0;
}

// D[m][n] = A[m][n] - C[m][n]  (C == nullptr: copy) for an M x 128 block
__global__ void k_block_sub(const double* __restrict__ A, int64_t lda, const double* __restrict__ C, int64_t ldc,
                            double* __restrict__ D, int64_t ldd, int M) {
    int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * MB_NB) return;
    int m = e / MB_NB, n = e - m * MB_NB;
    double v = A[int64_t(m) * lda + n];
    if (C) v -= C[int64_t(m) * ldc + n];
    D[int64_t(m) * ldd + n] = v;
}

// One CTA (256 threads): S (128 x 128, symmetric, lower part read) -> L = chol(S) written to Lout (lower, upper zeroed),
// T = L^-1 and -T written as full 128 x 128 row-major blocks (upper parts zero).  info[0] receives 1 + the global row of the
// first non-positive pivot (0 stays untouched otherwise).
__global__ void __launch_bounds__(256, 1)
k_potrf_trtri_128(const double* __restrict__ S, int64_t lds_, double* __restrict__ Lout, int64_t ldl, double* __restrict__ T,
                  double* __restrict__ negT, int row0, int* __restrict__ info) {
    extern __shared__ double sa[];                       // [128][129] + dinv[128]
    double* dinv = sa + MB_NB * MB_LDS;
    const int tid = threadIdx.x;
    for (int e = tid; e < MB_NB * MB_NB; e += blockDim.x) {
        int r = e >> 7, c = e & 127;
        sa[r * MB_LDS + c] = S[int64_t(r) * lds_ + c];
    }
    const int r = tid & 127, h = tid >> 7;
    for (int k = 0; k < MB_NB; ++k) {
        __syncthreads();
        const double akk = sa[k * MB_LDS + k];
        if (!(akk > 0.0) && tid == 0 && info) atomicCAS(info, 0, row0 + k + 1);
        const double d = sqrt(akk), di = 1.0 / d;
        if (h == 0 && r > k) sa[r * MB_LDS + k] *= di;
        __syncthreads();                                  // every thread has read the pivot: it may be overwritten now
        if (tid == 0) { sa[k * MB_LDS + k] = d; dinv[k] = di; }
        if (r > k) {
            const double lrk = sa[r * MB_LDS + k];
            for (int c = k + 1 + h; c <= r; c += 2) sa[r * MB_LDS + c] = fma(-lrk, sa[c * MB_LDS + k], sa[r * MB_LDS + c]);
        }
    }
    __syncthreads();
    for (int e = tid; e < MB_NB * MB_NB; e += blockDim.x) {
        int rr = e >> 7, c = e & 127;
        Lout[int64_t(rr) * ldl + c] = (c <= rr) ? sa[rr * MB_LDS + c] : 0.0;
    }
    __syncthreads();
    // T = L^-1, column c by thread c: T[rr][c] = -dinv[rr] sum_{k=c}^{rr-1} L[rr][k] T[k][c]; T[k][c] is kept at sa[c][k] (upper part)
    if (tid < MB_NB) {
        const int c = tid;
        for (int rr = c + 1; rr < MB_NB; ++rr) {
            double acc = sa[rr * MB_LDS + c] * dinv[c];
            for (int k = c + 1; k < rr; ++k) acc = fma(sa[rr * MB_LDS + k], sa[c * MB_LDS + k], acc);
            sa[c * MB_LDS + rr] = -dinv[rr] * acc;
        }
    }
    __syncthreads();
    for (int e = tid; e < MB_NB * MB_NB; e += blockDim.x) {
        int rr = e >> 7, c = e & 127;
        double v = (c < rr) ? sa[c * MB_LDS + rr] : (c == rr ? dinv[rr] : 0.0);
        T[e] = v;
        negT[e] = -v;
    }
}

// out[c][r] = in[r][c] for a 128 x 128 block (in: contiguous 128 x 128; out: leading dimension ldo)
__global__ void k_block_transpose_out(const double* __restrict__ in, double* __restrict__ out, int64_t ldo) {
    __shared__ double tile[32][33];
    int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) tile[j][threadIdx.x] = in[(by + j) * MB_NB + bx + threadIdx.x];
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) out[int64_t(bx + j) * ldo + by + threadIdx.x] = tile[threadIdx.x][j];
}

// linv = linv_t^T (Np x Np), with the padding (rows / columns >= N) zeroed in both
__global__ void k_transpose_zero_pad(double* __restrict__ linv_t, double* __restrict__ linv, int N, int Np) {
    __shared__ double tile[32][33];
    int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        int rr = by + j, c = bx + threadIdx.x;
        double v = linv_t[int64_t(rr) * Np + c];
        if (rr >= N || c >= N) { v = 0.0; linv_t[int64_t(rr) * Np + c] = 0.0; }
        tile[j][threadIdx.x] = v;
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) linv[int64_t(bx + j) * Np + by + threadIdx.x] = tile[threadIdx.x][j];
}

__global__ void k_center_targets(const double* __restrict__ y, int N, int Np, double mean, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Np) out[i] = (i < N) ? y[i] - mean : 0.0;
}

struct ModelBuildWs {
    double *A, *XsT, *Tb, *negTb, *P1, *P2, *S, *yc, *t1;
    size_t bytes;
};

inline ModelBuildWs model_build_carve(void* ws, int64_t N, int64_t d) {
    const int64_t Np = (N + 127) / 128 * 128, nb = Np / MB_NB;
    char* base = static_cast<char*>(ws);
    size_t off = 0;
    auto take = [&](size_t n) { double* p = reinterpret_cast<double*>(base ? base + off : nullptr); off += (n * sizeof(double) + 255) / 256 * 256; return p; };
    ModelBuildWs w;
    w.A = take(size_t(Np) * Np);            // kernel matrix, overwritten by L
    w.XsT = take(size_t(Np) * d);
    w.Tb = take(size_t(nb) * MB_NB * MB_NB);
    w.negTb = take(size_t(nb) * MB_NB * MB_NB);
    w.P1 = take(size_t(Np) * MB_NB);
    w.P2 = take(size_t(Np) * MB_NB);
    w.S = take(size_t(MB_NB) * MB_NB);
    w.yc = take(Np);
    w.t1 = take(Np);
    w.bytes = off;
    return w;
}

inline cudaError_t model_build(const double* X, const double* y, int N, int d, const double* inv_ls, double mean,
                               double outputscale, double noise, double* alpha, double* linv, double* linv_t, int* info,
                               const ModelBuildWs& w, int num_sms, cudaStream_t s) {
    const int Np = (N + 127) / 128 * 128, nb = Np / MB_NB;
    cudaError_t e;
#define MB_CHECK(x) do { e = (x); if (e != cudaSuccess) return e; } while (0)
    if (info) MB_CHECK(cudaMemsetAsync(info, 0, sizeof(int), s));
    MB_CHECK(cudaMemsetAsync(w.A, 0, sizeof(double) * size_t(Np) * Np, s));
    MB_CHECK(cudaMemsetAsync(linv_t, 0, sizeof(double) * size_t(Np) * Np, s));
    // ---- A = K(X, X) + noise I (identity on the padding) ---------------------------------------------------------
    CBO_LAUNCHED(), k_scale_transpose<<<(Np * d + 255) / 256, 256, 0, s>>>(X, N, Np, d, inv_ls, w.XsT);
    MB_CHECK(launch_kernel_rows(X, N, w.XsT, N, Np, d, inv_ls, outputscale, 0.0, w.yc /* unused alpha: any Np buffer */,
                                nullptr, 0, w.A, nullptr, nullptr, num_sms, s));
    CBO_LAUNCHED(), k_fix_diag<<<(Np + 255) / 256, 256, 0, s>>>(w.A, N, Np, outputscale + noise);
    const size_t potrf_smem = sizeof(double) * (MB_NB * MB_LDS + MB_NB);
    static SmemOptIn optin;
    MB_CHECK(optin.ensure(k_potrf_trtri_128, potrf_smem));
    // ---- left-looking blocked Cholesky, in place in A ---------------------------------------------------------------
    for (int j = 0; j < nb; ++j) {
        double* Ajj = w.A + (int64_t(j) * MB_NB) * Np + int64_t(j) * MB_NB;
        const double* Lj0 = w.A + (int64_t(j) * MB_NB) * Np;             // L_j,0:j   (128 x 128 j)
        const int K = j * MB_NB;
        if (j > 0) {
            MB_CHECK(launch_gemm_tn(Lj0, Np, Lj0, Np, w.P1, MB_NB, MB_NB, MB_NB, K, TRI_FULL, nullptr, s, num_sms));
            CBO_LAUNCHED(), k_block_sub<<<(MB_NB * MB_NB + 255) / 256, 256, 0, s>>>(Ajj, Np, w.P1, MB_NB, w.S, MB_NB, MB_NB);
        } else {
            CBO_LAUNCHED(), k_block_sub<<<(MB_NB * MB_NB + 255) / 256, 256, 0, s>>>(Ajj, Np, nullptr, 0, w.S, MB_NB, MB_NB);
        }
        double* Tj = w.Tb + size_t(j) * MB_NB * MB_NB;
        double* nTj = w.negTb + size_t(j) * MB_NB * MB_NB;
        CBO_LAUNCHED(), k_potrf_trtri_128<<<1, 256, potrf_smem, s>>>(w.S, MB_NB, Ajj, Np, Tj, nTj, j * MB_NB, info);
        const int M = Np - (j + 1) * MB_NB;
        if (M > 0) {
            double* Aij = w.A + (int64_t(j + 1) * MB_NB) * Np + int64_t(j) * MB_NB;     // panel below the diagonal block
            const double* Li0 = w.A + (int64_t(j + 1) * MB_NB) * Np;
            if (j > 0) {
                MB_CHECK(launch_gemm_tn(Li0, Np, Lj0, Np, w.P1, MB_NB, M, MB_NB, K, TRI_FULL, nullptr, s, num_sms));
                CBO_LAUNCHED(), k_block_sub<<<(M * MB_NB + 255) / 256, 256, 0, s>>>(Aij, Np, w.P1, MB_NB, w.P2, MB_NB, M);
            } else {
                CBO_LAUNCHED(), k_block_sub<<<(M * MB_NB + 255) / 256, 256, 0, s>>>(Aij, Np, nullptr, 0, w.P2, MB_NB, M);
            }
            MB_CHECK(launch_gemm_tn(w.P2, MB_NB, Tj, MB_NB, Aij, Np, M, MB_NB, MB_NB, TRI_FULL, nullptr, s, num_sms));
        }
    }
    // ---- blocked triangular inverse, built transposed ------------------------------------------------------------------
    for (int i = 0; i < nb; ++i) {
        const double* Ti = w.Tb + size_t(i) * MB_NB * MB_NB;
        const double* nTi = w.negTb + size_t(i) * MB_NB * MB_NB;
        CBO_LAUNCHED(), k_block_transpose_out<<<dim3(4, 4), dim3(32, 8), 0, s>>>(Ti, linv_t + (int64_t(i) * MB_NB) * Np + int64_t(i) * MB_NB, Np);
        if (i > 0) {
            const int M = i * MB_NB;                                        // rows 0 .. 128 i of linv_t
            const double* Li0 = w.A + (int64_t(i) * MB_NB) * Np;
            // P^T[m][n] = sum_k linv_t[m][k] L_i[n][k]
            MB_CHECK(launch_gemm_tn(linv_t, Np, Li0, Np, w.P1, MB_NB, M, MB_NB, M, TRI_FULL, nullptr, s, num_sms));
            // linv_t[0:M, block i][m][n] = -sum_k P^T[m][k] T_i[n][k]
            MB_CHECK(launch_gemm_tn(w.P1, MB_NB, nTi, MB_NB, linv_t + int64_t(i) * MB_NB, Np, M, MB_NB, MB_NB, TRI_FULL,
                                    nullptr, s, num_sms));
        }
    }
    CBO_LAUNCHED(), k_transpose_zero_pad<<<dim3(Np / 32, Np / 32), dim3(32, 8), 0, s>>>(linv_t, linv, N, Np);
    // ---- alpha = L^-T (L^-1 (y - m)) ---------------------------------------------------------------------------------------
    CBO_LAUNCHED(), k_center_targets<<<(Np + 255) / 256, 256, 0, s>>>(y, N, Np, mean, w.yc);
    CBO_LAUNCHED(), k_probe_matvec<<<(Np + 7) / 8, 256, 0, s>>>(linv, Np, w.yc, 1, w.t1, 0, 1);
    CBO_LAUNCHED(), k_probe_matvec<<<(Np + 7) / 8, 256, 0, s>>>(linv_t, Np, w.t1, 1, alpha, 1, 0);
#undef MB_CHECK
    return cudaGetLastError();
}

}  // namespace cbo
