// Training-side steps next to the hot path, kept on the device so that an SGLD step has no host round trip:
//   k_ratio_train_grad + k_ratio_adam_ema : RatioEstimator.optimize_callback from the minibatch on
//       (conformalbo/ratio_estimation.py:133-144: BCEWithLogitsLoss over [neg; pos], Adam step, update_target_network)
//   k_sgld_update : the SGHMC update of the candidates with the box penalty of optimize_acqf_sgld
//       (conformalbo/optim/sgld.py:9-47, conformalbo/optim/optimize.py:224-235)
// Parameters live in ONE flat fp64 vector in torch's parameter order (W1 H x d, b1, W2 H x H, b2, W3, b3), H = 64.
// The classifier is tiny (4.8 k parameters, 2 x N rows per step): the step is latency bound, so the design goal is two
// launches with fixed-order reductions (bit-reproducible), not tensor-core throughput.
#pragma once
#include "common.cuh"

namespace cbo {

constexpr int RT_H = 64;          // hidden width of the reference's classifier (ratio_estimation.py:73-79)
constexpr int RT_ROWS = 64;       // minibatch rows per CTA
constexpr int RT_THREADS = 256;
constexpr int RT_LD = RT_H + 1;   // padded leading dimension of the 64 x 64 shared tiles (conflict-free column reads)
constexpr int RT_MAX_D = 64;

__host__ __device__ inline int64_t ratio_num_params(int d) { return int64_t(RT_H) * d + RT_H + RT_H * RT_H + RT_H + RT_H + 1; }

__host__ __device__ inline size_t ratio_train_smem_bytes(int d) {
    const int ldx = d | 1;
    // W1 (H x ldx), X (ROWS x ldx), W2, H1, H2 / dZ1, dZ2 (64 x 65 each), b1, b2, W3, dl (64 each), 8 warp partials
    return sizeof(double) * (size_t(RT_H) * ldx + size_t(RT_ROWS) * ldx + 4 * size_t(RT_H) * RT_LD + 4 * RT_H + 16);
}

// One CTA: rows [tile*64, tile*64+64) of the minibatch [x_neg (label 0); x_pos[pos_idx] (label 1)]; writes this tile's
// contribution to the gradient of the MEAN loss into partial[tile][0..P) and its loss sum into partial[tile][P].
__global__ void __launch_bounds__(RT_THREADS)
k_ratio_train_grad(const double* __restrict__ theta, int d, const double* __restrict__ x_neg, int n_neg,
                   const double* __restrict__ x_pos, const int64_t* __restrict__ pos_idx, int n_pos,
                   double* __restrict__ partial) {
    extern __shared__ double sm[];
    const int ldx = d | 1;
    double* sW1 = sm;                          // H x ldx
    double* sX = sW1 + RT_H * ldx;             // ROWS x ldx
    double* sW2 = sX + RT_ROWS * ldx;          // H x LD   W2[k][m]
    double* sH1 = sW2 + RT_H * RT_LD;          // ROWS x LD
    double* sH2 = sH1 + RT_ROWS * RT_LD;       // ROWS x LD  (later dZ1)
    double* sD2 = sH2 + RT_ROWS * RT_LD;       // ROWS x LD  dZ2
    double* sb1 = sD2 + RT_ROWS * RT_LD;
    double* sb2 = sb1 + RT_H;
    double* sW3 = sb2 + RT_H;
    double* sdl = sW3 + RT_H;                  // ROWS
    double* sred = sdl + RT_ROWS;              // 8 warp partial loss sums (+ spare)

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tile = blockIdx.x;
    const int n_tot = n_neg + n_pos;
    const int64_t P = ratio_num_params(d);
    const int64_t oW1 = 0, ob1 = int64_t(RT_H) * d, oW2 = ob1 + RT_H, ob2 = oW2 + RT_H * RT_H, oW3 = ob2 + RT_H, ob3 = oW3 + RT_H;
    double* out = partial + int64_t(tile) * (P + 1);

    for (int o = tid; o < RT_H * d; o += RT_THREADS) sW1[(o / d) * ldx + (o % d)] = theta[oW1 + o];
    for (int o = tid; o < RT_H * RT_H; o += RT_THREADS) sW2[(o >> 6) * RT_LD + (o & 63)] = theta[oW2 + o];
    if (tid < RT_H) { sb1[tid] = theta[ob1 + tid]; sb2[tid] = theta[ob2 + tid]; sW3[tid] = theta[oW3 + tid]; }
    for (int o = tid; o < RT_ROWS * d; o += RT_THREADS) {
        const int r = o / d, j = o % d, rg = tile * RT_ROWS + r;
        double v = 0.0;
        if (rg < n_neg) v = x_neg[int64_t(rg) * d + j];
        else if (rg < n_tot) {
            const int64_t src = pos_idx ? pos_idx[rg - n_neg] : int64_t(rg - n_neg);
            v = x_pos[src * d + j];
        }
        sX[r * ldx + j] = v;
    }
    __syncthreads();

    // thread tile of the 64 x 64 products: rows 4 tr + a, columns tc + 16 b
    const int tr = tid >> 4, tc = tid & 15;

    {   // H1 = relu(X W1^T + b1)
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = sb1[tc + 16 * b];
        for (int j = 0; j < d; ++j) {
            double xa[4], wb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) xa[a] = sX[(4 * tr + a) * ldx + j];
#pragma unroll
            for (int b = 0; b < 4; ++b) wb[b] = sW1[(tc + 16 * b) * ldx + j];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(xa[a], wb[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) sH1[(4 * tr + a) * RT_LD + tc + 16 * b] = fmax(acc[a][b], 0.0);
    }
    __syncthreads();
    {   // H2 = relu(H1 W2^T + b2)
        double acc[4][4];
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc[a][b] = sb2[tc + 16 * b];
        for (int m = 0; m < RT_H; ++m) {
            double ha[4], wb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) ha[a] = sH1[(4 * tr + a) * RT_LD + m];
#pragma unroll
            for (int b = 0; b < 4; ++b) wb[b] = sW2[(tc + 16 * b) * RT_LD + m];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(ha[a], wb[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) sH2[(4 * tr + a) * RT_LD + tc + 16 * b] = fmax(acc[a][b], 0.0);
    }
    __syncthreads();
    {   // logits, loss and d loss / d logit: 4 threads per row
        const int r = tid >> 2, part = tid & 3, rg = tile * RT_ROWS + r;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 16; ++k) s = fma(sH2[r * RT_LD + part * 16 + k], sW3[part * 16 + k], s);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        double loss = 0.0;
        if (part == 0) {
            double dl = 0.0;
            if (rg < n_tot) {
                const double logit = s + theta[ob3];
                const double y = rg < n_neg ? 0.0 : 1.0;
                const double e = exp(-fabs(logit));
                loss = fmax(logit, 0.0) - logit * y + log1p(e);
                const double sig = logit >= 0.0 ? 1.0 / (1.0 + e) : e / (1.0 + e);
                dl = (sig - y) / double(n_tot);
            }
            sdl[r] = dl;
        }
        loss = warp_sum(loss);
        if (lane == 0) sred[warp] = loss;
    }
    __syncthreads();
    // dZ2 = dl W3^T . [H2 > 0]; gW3 = H2^T dl; gb3 = sum dl; tile loss
    for (int o = tid; o < RT_ROWS * RT_H; o += RT_THREADS) {
        const int r = o >> 6, k = o & 63;
        sD2[r * RT_LD + k] = sH2[r * RT_LD + k] > 0.0 ? sdl[r] * sW3[k] : 0.0;
    }
    if (tid < RT_H) {
        double g = 0.0;
        for (int r = 0; r < RT_ROWS; ++r) g = fma(sdl[r], sH2[r * RT_LD + tid], g);
        out[oW3 + tid] = g;
    } else if (tid == RT_H) {
        double g = 0.0;
        for (int r = 0; r < RT_ROWS; ++r) g += sdl[r];
        out[ob3] = g;
    } else if (tid == RT_H + 1) {
        double l = 0.0;
        for (int w = 0; w < RT_THREADS / 32; ++w) l += sred[w];
        out[P] = l;
    }
    __syncthreads();
    {   // gW2[k][m] = sum_r dZ2[r][k] H1[r][m]   (k = 4 tr + a, m = tc + 16 b)
        double acc[4][4] = {};
        for (int r = 0; r < RT_ROWS; ++r) {
            double da[4], hb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) da[a] = sD2[r * RT_LD + 4 * tr + a];
#pragma unroll
            for (int b = 0; b < 4; ++b) hb[b] = sH1[r * RT_LD + tc + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(da[a], hb[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) out[oW2 + (4 * tr + a) * RT_H + tc + 16 * b] = acc[a][b];
    }
    {   // dZ1 = (dZ2 W2) . [H1 > 0]  -> overwrites H2 (everything that read H2 finished before the last barrier)
        double acc[4][4] = {};
        for (int k = 0; k < RT_H; ++k) {
            double da[4], wb[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) da[a] = sD2[(4 * tr + a) * RT_LD + k];
#pragma unroll
            for (int b = 0; b < 4; ++b) wb[b] = sW2[k * RT_LD + tc + 16 * b];
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int b = 0; b < 4; ++b) acc[a][b] = fma(da[a], wb[b], acc[a][b]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int r = 4 * tr + a, m = tc + 16 * b;
                sH2[r * RT_LD + m] = sH1[r * RT_LD + m] > 0.0 ? acc[a][b] : 0.0;
            }
    }
    if (tid < RT_H) {   // gb2
        double g = 0.0;
        for (int r = 0; r < RT_ROWS; ++r) g += sD2[r * RT_LD + tid];
        out[ob2 + tid] = g;
    }
    __syncthreads();
    // gW1[k][j] = sum_r dZ1[r][k] X[r][j];  gb1[k] = sum_r dZ1[r][k]
    for (int o = tid; o < RT_H * d; o += RT_THREADS) {
        const int k = o / d, j = o % d;
        double g = 0.0;
        for (int r = 0; r < RT_ROWS; ++r) g = fma(sH2[r * RT_LD + k], sX[r * ldx + j], g);
        out[oW1 + o] = g;
    }
    if (tid < RT_H) {
        double g = 0.0;
        for (int r = 0; r < RT_ROWS; ++r) g += sH2[r * RT_LD + tid];
        out[ob1 + tid] = g;
    }
}

// Thread per parameter: gradient = sum of the tile partials in tile order (+ L2 weight decay), torch.optim.Adam update
// (bias-corrected, eps outside the square root), then the EMA of the target network.  loss (may be NULL) <- mean loss.
__global__ void k_ratio_adam_ema(double* __restrict__ theta, double* __restrict__ adam_m, double* __restrict__ adam_v,
                                 double* __restrict__ target, const double* __restrict__ partial, int ntiles, int64_t P,
                                 double n_rows, double step_size, double inv_sqrt_bc2, double beta1, double beta2,
                                 double eps, double weight_decay, double ema_weight, double* __restrict__ loss) {
    const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (p > P) return;
    double g = 0.0;
    for (int t = 0; t < ntiles; ++t) g += partial[int64_t(t) * (P + 1) + p];
    if (p == P) {
        if (loss) *loss = g / n_rows;
        return;
    }
    const double th = theta[p];
    g = fma(weight_decay, th, g);
    const double m = adam_m[p] + (1.0 - beta1) * (g - adam_m[p]);
    const double v = beta2 * adam_v[p] + (1.0 - beta2) * g * g;
    adam_m[p] = m;
    adam_v[p] = v;
    const double denom = sqrt(v) * inv_sqrt_bc2 + eps;
    const double tn = th - step_size * (m / denom);
    theta[p] = tn;
    target[p] = target[p] * (1.0 - ema_weight) + ema_weight * tn;
}

// X, buf: n x d (n = restarts * q rows); dacq: d acq / d X; noise: N(0,1) draws.  Gradient of the reference's loss
// -sum(acq) + 1e2 (sigmoid((lb - X)/1e-2).mean() + sigmoid((X - ub)/1e-2).mean()), then the SGHMC update.
__global__ void k_sgld_update(double* __restrict__ X, double* __restrict__ buf, const double* __restrict__ dacq,
                              const double* __restrict__ noise, const double* __restrict__ lb,
                              const double* __restrict__ ub, int64_t n_elem, int d, double lr, double momentum,
                              double noise_scale, double pen_scale, double pen_width) {
    const int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= n_elem) return;
    const int j = int(e % d);
    const double x = X[e];
    const double sl = 1.0 / (1.0 + exp(-(lb[j] - x) / pen_width));
    const double su = 1.0 / (1.0 + exp(-(x - ub[j]) / pen_width));
    const double g = -dacq[e] + pen_scale * (su * (1.0 - su) - sl * (1.0 - sl)) / (pen_width * double(n_elem));
    if (momentum == 0.0) {           // plain SGLD (sgld.py:41-46)
        X[e] = x - lr * g + noise_scale * noise[e];
        return;
    }
    const double b = buf[e] * (1.0 - momentum) - lr * g + noise_scale * noise[e];
    buf[e] = b;
    X[e] = x + b;
}

// k distinct pseudo-random indices out of [0, n): out[t] = P(t), P a keyed bijection of [0, n) -- a 6-round Feistel network on
// the next even-bit-width power of two (<= 4 n points) with cycle walking back into range (expected < 4 rounds per index).
// The first k values of a pseudo-random permutation: what RatioEstimator.optimize_callback draws with
// np.random.permutation(num_positive)[:num_negative] (conformalbo/ratio_estimation.py:129-131) without building the permutation.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
__global__ void k_sample_subset(int64_t n, int64_t k, uint64_t seed, int half_bits, int64_t* __restrict__ out) {
    const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= k) return;
    const uint32_t mask = (1u << half_bits) - 1u;
    uint32_t key[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) key[r] = mix32(uint32_t(seed >> ((r & 1) ? 32 : 0)) + 0x9e3779b9u * uint32_t(r + 1));
    uint64_t x = uint64_t(t);
    do {
        uint32_t L = uint32_t(x >> half_bits) & mask, R = uint32_t(x) & mask;
#pragma unroll
        for (int r = 0; r < 6; ++r) {
            const uint32_t F = mix32(R ^ key[r]) & mask;
            const uint32_t nl = R;
            R = L ^ F;
            L = nl;
        }
        x = (uint64_t(L) << half_bits) | R;
    } while (x >= uint64_t(n));
    out[t] = int64_t(x);
}

}  // namespace cbo
