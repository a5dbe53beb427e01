// ratio_mma.cuh -- RatioEstimator.forward (conformalbo/ratio_estimation.py:97-99) and its reverse pass on the fp64
// tensor path (mma.sync m8n8k4 = DMMA.8x8x4): the d -> H -> ReLU -> H -> ReLU -> 1 target network evaluated for every
// candidate (helpers.py:292-293 evaluates it per (candidate, grid point); it is g-independent).
//
// One warp owns 8 m (points) x all H hidden units.  The DMMA accumulator layout (lane l holds C[l/4][2(l%4)+{0,1}]) is fed
// straight back as the next layer's A operand by PERMUTING the contraction index: k-slot q of step (tile, e) is hidden
// unit 8 tile + 2q + e, which lane (row, q) already holds -- no shuffles, no shared-memory round trip between layers.  The
// weights are staged once per CTA in "fragment order" (the 32 lanes of one B fragment are 32 consecutive doubles):
// conflict-free, one LDS.64 per DMMA (per two DMMAs with 2 m-tiles per warp).
//   forward:  H1 = relu(X W1^T + b1), A2 = H1 W2^T + b2, logit = relu(A2) . W3 + b3, w = min(p,1-1e-6)/max(1-p,1e-6)
//   reverse:  g2 = dlogit W3 [A2 > 0], G1 = (g2 W2) [H1 > 0], dX += G1 W1
#pragma once
#include "common.cuh"
#include "rows.cuh"
#include "score_bwd.cuh"

namespace cbo {

constexpr int RM_WARPS = 8;
constexpr int RM_NT = MAX_H / 8;          // hidden tiles of 8
constexpr int RM_KS1 = MAX_D / 4;         // input k-steps of 4

// doubles of shared memory: W1 fragments [KS1][NT][32], W2 fragments [NT][NT][2][32], b1, b2, W3 [MAX_H each];
// reverse adds W2 "transposed" fragments [NT][NT][2][32] and W1 reverse fragments [d tiles <= 8][NT][2][32]
inline size_t ratio_mma_smem_bytes(int d, bool reverse) {
    const size_t ks1 = (d + 3) / 4, dt = (d + 7) / 8;
    size_t n = ks1 * RM_NT * 32 + size_t(RM_NT) * RM_NT * 64 + 3 * MAX_H;
    if (reverse) n += size_t(RM_NT) * RM_NT * 64 + dt * RM_NT * 64;
    return n * sizeof(double);
}

struct RatioMmaSmem {
    double *w1f, *w2f, *b1, *b2, *w3, *w2r, *w1r;
};

__device__ __forceinline__ RatioMmaSmem ratio_mma_stage(double* sm, int d, int H, const double* __restrict__ W1,
                                                        const double* __restrict__ b1, const double* __restrict__ W2,
                                                        const double* __restrict__ b2, const double* __restrict__ W3,
                                                        bool reverse) {
    const int ks1 = (d + 3) / 4, dt = (d + 7) / 8;
    RatioMmaSmem s;
    s.w1f = sm;
    s.w2f = s.w1f + size_t(ks1) * RM_NT * 32;
    s.b1 = s.w2f + size_t(RM_NT) * RM_NT * 64;
    s.b2 = s.b1 + MAX_H;
    s.w3 = s.b2 + MAX_H;
    s.w2r = s.w3 + MAX_H;
    s.w1r = s.w2r + size_t(RM_NT) * RM_NT * 64;
    const int tid = threadIdx.x, nthr = blockDim.x;
    for (int e = tid; e < ks1 * RM_NT * 32; e += nthr) {          // B[k][n] = W1[n][k]: lane l <- W1[8 nt + l/4][4 ks + l%4]
        int l = e & 31, nt = (e >> 5) % RM_NT, ks = (e >> 5) / RM_NT;
        int n = 8 * nt + (l >> 2), k = 4 * ks + (l & 3);
        s.w1f[e] = (n < H && k < d) ? W1[n * d + k] : 0.0;
    }
    for (int e = tid; e < RM_NT * RM_NT * 64; e += nthr) {         // [nto][nti][e2][lane] <- W2[8 nto + l/4][8 nti + 2(l%4) + e2]
        int l = e & 31, e2 = (e >> 5) & 1, nti = (e >> 6) % RM_NT, nto = (e >> 6) / RM_NT;
        int n = 8 * nto + (l >> 2), k = 8 * nti + 2 * (l & 3) + e2;
        s.w2f[e] = (n < H && k < H) ? W2[n * H + k] : 0.0;
    }
    for (int e = tid; e < MAX_H; e += nthr) {
        s.b1[e] = e < H ? b1[e] : 0.0;
        s.b2[e] = e < H ? b2[e] : 0.0;
        s.w3[e] = e < H ? W3[e] : 0.0;
    }
    if (reverse) {
        for (int e = tid; e < RM_NT * RM_NT * 64; e += nthr) {     // G1 = g2 W2: B[k = j][n = m] = W2[j][m], j = 8 nti + 2(l%4) + e2
            int l = e & 31, e2 = (e >> 5) & 1, nti = (e >> 6) % RM_NT, nto = (e >> 6) / RM_NT;
            int m = 8 * nto + (l >> 2), j = 8 * nti + 2 * (l & 3) + e2;
            s.w2r[e] = (m < H && j < H) ? W2[j * H + m] : 0.0;
        }
        for (int e = tid; e < dt * RM_NT * 64; e += nthr) {        // dX = G1 W1: B[k = m][n = kk] = W1[m][kk], m = 8 nti + 2(l%4) + e2
            int l = e & 31, e2 = (e >> 5) & 1, nti = (e >> 6) % RM_NT, dto = (e >> 6) / RM_NT;
            int kk = 8 * dto + (l >> 2), m = 8 * nti + 2 * (l & 3) + e2;
            s.w1r[e] = (m < H && kk < d) ? W1[m * d + kk] : 0.0;
        }
    }
    __syncthreads();
    return s;
}

// H1 tile of 8 points starting at p0: h1[nt][e] = relu(...) for hidden unit 8 nt + 2(l%4) + e, row l/4
__device__ __forceinline__ void ratio_mma_layer1(const RatioMmaSmem& s, const double* __restrict__ x, int n, int d, int p0,
                                                 int lane, double (&h1)[RM_NT][2]) {
    const int r = lane >> 2, q = lane & 3, ks1 = (d + 3) / 4;
    const int p = min(p0 + r, n - 1);
#pragma unroll
    for (int nt = 0; nt < RM_NT; ++nt) { h1[nt][0] = s.b1[8 * nt + 2 * q]; h1[nt][1] = s.b1[8 * nt + 2 * q + 1]; }
    for (int ks = 0; ks < ks1; ++ks) {
        const int k = 4 * ks + q;
        const double a = (k < d) ? x[int64_t(p) * d + k] : 0.0;
#pragma unroll
        for (int nt = 0; nt < RM_NT; ++nt) dmma884(h1[nt][0], h1[nt][1], a, s.w1f[(ks * RM_NT + nt) * 32 + lane]);
    }
#pragma unroll
    for (int nt = 0; nt < RM_NT; ++nt) { h1[nt][0] = fmax(h1[nt][0], 0.0); h1[nt][1] = fmax(h1[nt][1], 0.0); }
}

__device__ __forceinline__ double ratio_from_logit(double logit, double clamp_min, double* dw_dlogit) {
    const double pr = sigmoid_acc(logit);
    const double num = fmin(pr, 1.0 - 1e-6), den = fmax(1.0 - pr, 1e-6);
    const double w = num / den;
    if (dw_dlogit) {
        double dw_dp = ((pr <= 1.0 - 1e-6) ? 1.0 / den : 0.0) + ((1.0 - pr >= 1e-6) ? num / (den * den) : 0.0);
        if (!(w >= clamp_min)) dw_dp = 0.0;
        *dw_dlogit = dw_dp * pr * (1.0 - pr);
    }
    return fmax(w, clamp_min);
}

// forward: 2 m-tiles (16 points) per warp pass
__global__ void __launch_bounds__(RM_WARPS * 32)
k_ratio_mlp_mma(const double* __restrict__ x, int n, int d, int H, const double* __restrict__ W1,
                const double* __restrict__ b1, const double* __restrict__ W2, const double* __restrict__ b2,
                const double* __restrict__ W3, const double* __restrict__ b3, double clamp_min, double* __restrict__ out) {
    extern __shared__ __align__(16) double rsm[];
    const RatioMmaSmem s = ratio_mma_stage(rsm, d, H, W1, b1, W2, b2, W3, false);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const double bias3 = b3[0];
    const int tiles = (n + 15) / 16;
    for (int t = blockIdx.x * RM_WARPS + warp; t < tiles; t += gridDim.x * RM_WARPS) {
        const int p0 = t * 16;
        double h1a[RM_NT][2], h1b[RM_NT][2];
        ratio_mma_layer1(s, x, n, d, p0, lane, h1a);
        ratio_mma_layer1(s, x, n, d, p0 + 8, lane, h1b);
        double pa = 0.0, pb = 0.0;
#pragma unroll 1
        for (int nto = 0; nto < RM_NT; ++nto) {
            const double bb0 = s.b2[8 * nto + 2 * q], bb1 = s.b2[8 * nto + 2 * q + 1];
            double a0 = bb0, a1 = bb1, c0 = bb0, c1 = bb1;
            const double* wf = s.w2f + size_t(nto) * RM_NT * 64 + lane;
#pragma unroll
            for (int nti = 0; nti < RM_NT; ++nti) {
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const double bfrag = wf[(nti * 2 + e) * 32];
                    dmma884(a0, a1, h1a[nti][e], bfrag);
                    dmma884(c0, c1, h1b[nti][e], bfrag);
                }
            }
            const double w30 = s.w3[8 * nto + 2 * q], w31 = s.w3[8 * nto + 2 * q + 1];
            pa = fma(w30, fmax(a0, 0.0), fma(w31, fmax(a1, 0.0), pa));
            pb = fma(w30, fmax(c0, 0.0), fma(w31, fmax(c1, 0.0), pb));
        }
        pa += __shfl_xor_sync(0xffffffffu, pa, 1); pa += __shfl_xor_sync(0xffffffffu, pa, 2);
        pb += __shfl_xor_sync(0xffffffffu, pb, 1); pb += __shfl_xor_sync(0xffffffffu, pb, 2);
        if (q == 0) {
            if (p0 + r < n) out[p0 + r] = ratio_from_logit(pa + bias3, clamp_min, nullptr);
            if (p0 + 8 + r < n) out[p0 + 8 + r] = ratio_from_logit(pb + bias3, clamp_min, nullptr);
        }
    }
}

// reverse: dX[p, :] (+)= adj[p][2] * d ratio(x_p) / d x   (1 m-tile per warp pass; clamps cut the gradient like torch.clamp)
__global__ void __launch_bounds__(RM_WARPS * 32)
k_ratio_mlp_mma_bwd(const double* __restrict__ x, int n, int d, int H, const double* __restrict__ W1,
                    const double* __restrict__ b1, const double* __restrict__ W2, const double* __restrict__ b2,
                    const double* __restrict__ W3, const double* __restrict__ b3, double clamp_min,
                    const double* __restrict__ adj, int adj_stride, int adj_off, double* __restrict__ dX, int accumulate) {
    extern __shared__ __align__(16) double rsm[];
    const RatioMmaSmem s = ratio_mma_stage(rsm, d, H, W1, b1, W2, b2, W3, true);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, r = lane >> 2, q = lane & 3;
    const double bias3 = b3[0];
    const int tiles = (n + 7) / 8, dt = (d + 7) / 8;
    for (int t = blockIdx.x * RM_WARPS + warp; t < tiles; t += gridDim.x * RM_WARPS) {
        const int p0 = t * 8;
        double h1[RM_NT][2], g2[RM_NT][2];
        ratio_mma_layer1(s, x, n, d, p0, lane, h1);
        double part = 0.0;
#pragma unroll
        for (int nto = 0; nto < RM_NT; ++nto) {
            double a0 = s.b2[8 * nto + 2 * q], a1 = s.b2[8 * nto + 2 * q + 1];
            const double* wf = s.w2f + size_t(nto) * RM_NT * 64 + lane;
#pragma unroll
            for (int nti = 0; nti < RM_NT; ++nti) {
                dmma884(a0, a1, h1[nti][0], wf[(nti * 2) * 32]);
                dmma884(a0, a1, h1[nti][1], wf[(nti * 2 + 1) * 32]);
            }
            const double w30 = s.w3[8 * nto + 2 * q], w31 = s.w3[8 * nto + 2 * q + 1];
            part = fma(w30, fmax(a0, 0.0), fma(w31, fmax(a1, 0.0), part));
            g2[nto][0] = (a0 > 0.0) ? w30 : 0.0;              // d logit / d A2, scaled by dlogit below
            g2[nto][1] = (a1 > 0.0) ? w31 : 0.0;
        }
        part += __shfl_xor_sync(0xffffffffu, part, 1);
        part += __shfl_xor_sync(0xffffffffu, part, 2);
        double dwl;
        ratio_from_logit(part + bias3, clamp_min, &dwl);
        const int p = min(p0 + r, n - 1);
        const double dlogit = adj[int64_t(p) * adj_stride + adj_off] * dwl;
#pragma unroll
        for (int nt = 0; nt < RM_NT; ++nt) { g2[nt][0] *= dlogit; g2[nt][1] *= dlogit; }
        // G1 = (g2 W2) [H1 > 0], overwriting h1
        double g1[RM_NT][2];
#pragma unroll
        for (int nto = 0; nto < RM_NT; ++nto) {
            double a0 = 0.0, a1 = 0.0;
            const double* wf = s.w2r + size_t(nto) * RM_NT * 64 + lane;
#pragma unroll
            for (int nti = 0; nti < RM_NT; ++nti) {
                dmma884(a0, a1, g2[nti][0], wf[(nti * 2) * 32]);
                dmma884(a0, a1, g2[nti][1], wf[(nti * 2 + 1) * 32]);
            }
            g1[nto][0] = (h1[nto][0] > 0.0) ? a0 : 0.0;
            g1[nto][1] = (h1[nto][1] > 0.0) ? a1 : 0.0;
        }
        // dX tile = G1 W1
        for (int dto = 0; dto < dt; ++dto) {
            double a0 = 0.0, a1 = 0.0;
            const double* wf = s.w1r + size_t(dto) * RM_NT * 64 + lane;
#pragma unroll
            for (int nti = 0; nti < RM_NT; ++nti) {
                dmma884(a0, a1, g1[nti][0], wf[(nti * 2) * 32]);
                dmma884(a0, a1, g1[nti][1], wf[(nti * 2 + 1) * 32]);
            }
            const int k0 = 8 * dto + 2 * q;
            if (p0 + r < n) {
                double* o = dX + int64_t(p0 + r) * d + k0;
                if (k0 < d) o[0] = accumulate ? o[0] + a0 : a0;          // accumulate = 0: this kernel runs FIRST (side stream) and
                if (k0 + 1 < d) o[1] = accumulate ? o[1] + a1 : a1;      // the kernel-row term is added to it afterwards
            }
        }
    }
}

inline cudaError_t launch_ratio_mlp(const double* x, int n, int d, int H, const double* W1, const double* b1,
                                    const double* W2, const double* b2, const double* W3, const double* b3,
                                    double clamp_min, double* out, int num_sms, cudaStream_t st) {
    const size_t smem = ratio_mma_smem_bytes(d, false);
    static SmemOptIn optin;
    if (cudaError_t e = optin.ensure(k_ratio_mlp_mma, smem); e != cudaSuccess) return e;
    int grid = ((n + 15) / 16 + RM_WARPS - 1) / RM_WARPS;
    if (grid > 2 * num_sms) grid = 2 * num_sms;
    CBO_LAUNCHED(), k_ratio_mlp_mma<<<grid, RM_WARPS * 32, smem, st>>>(x, n, d, H, W1, b1, W2, b2, W3, b3, clamp_min, out);
    return cudaGetLastError();
}

inline cudaError_t launch_ratio_mlp_bwd(const double* x, int n, int d, int H, const double* W1, const double* b1,
                                        const double* W2, const double* b2, const double* W3, const double* b3,
                                        double clamp_min, const double* adj, int adj_stride, int adj_off, double* dX,
                                        int num_sms, cudaStream_t st, bool accumulate = true) {
    const size_t smem = ratio_mma_smem_bytes(d, true);
    static SmemOptIn optin;
    if (cudaError_t e = optin.ensure(k_ratio_mlp_mma_bwd, smem); e != cudaSuccess) return e;
    int grid = ((n + 7) / 8 + RM_WARPS - 1) / RM_WARPS;
    if (grid > 2 * num_sms) grid = 2 * num_sms;
    CBO_LAUNCHED(), k_ratio_mlp_mma_bwd<<<grid, RM_WARPS * 32, smem, st>>>(x, n, d, H, W1, b1, W2, b2, W3, b3, clamp_min, adj, adj_stride, adj_off, dX, accumulate ? 1 : 0);
    return cudaGetLastError();
}

}  // namespace cbo
