// bounds.cuh -- conformal band endpoints from (grid, mask): conf_mask_to_bounds (conformalbo/helpers.py:15-91).
// The reference sorts on the device and then walks every candidate in a Python loop; here one CTA
// per (candidate, output) does a shared-memory bitonic sort of (y_g, mask_g), finds the first / last
// accepted position by a block min/max and interpolates to the 0.5 crossing (helpers.py:58-71).
#pragma once
#include <math_constants.h>
#include "common.cuh"

namespace cbo {

constexpr int BOUNDS_THREADS = 256;
constexpr int BOUNDS_MAX_G = 2048;

// grid, mask: b x G x t (row-major); lb, ub: b x t
__global__ void __launch_bounds__(BOUNDS_THREADS)
k_mask_to_bounds(const double* __restrict__ grid, const double* __restrict__ mask, int b, int G, int t, int Gp,
                 double* __restrict__ lb, double* __restrict__ ub) {
    extern __shared__ __align__(16) double sm[];
    double* key = sm;            // Gp
    double* val = sm + Gp;       // Gp
    __shared__ int s_first, s_last;
    const int pair = blockIdx.x;             // (candidate, output)
    const int c = pair / t, o = pair - c * t;
    const int tid = threadIdx.x;
    for (int g = tid; g < Gp; g += BOUNDS_THREADS) {
        if (g < G) {
            key[g] = grid[(int64_t(c) * G + g) * t + o];
            val[g] = mask[(int64_t(c) * G + g) * t + o];
        } else {
            key[g] = CUDART_INF;             // padding sorts to the end and is never accepted
            val[g] = 0.0;
        }
    }
    if (tid == 0) { s_first = Gp; s_last = -1; }
    __syncthreads();
    for (int k = 2; k <= Gp; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < Gp; i += BOUNDS_THREADS) {
                int ixj = i ^ j;
                if (ixj > i) {
                    bool up = (i & k) == 0;
                    double a = key[i], bb = key[ixj];
                    if ((a > bb) == up) {
                        key[i] = bb; key[ixj] = a;
                        double tv = val[i]; val[i] = val[ixj]; val[ixj] = tv;
                    }
                }
            }
            __syncthreads();
        }
    }
    int first = Gp, last = -1;
    for (int g = tid; g < G; g += BOUNDS_THREADS) {
        if (val[g] >= 0.5) { first = min(first, g); last = max(last, g); }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
        last = max(last, __shfl_xor_sync(0xffffffffu, last, off));
    }
    if ((tid & 31) == 0) { atomicMin(&s_first, first); atomicMax(&s_last, last); }
    __syncthreads();
    if (tid == 0) {
        first = s_first; last = s_last;
        double lo, hi;
        if (last < 0) {
            lo = hi = CUDART_INF;                                      // helpers.py:48-51
        } else {
            if (first != 0)
                lo = key[first - 1] + (0.5 - val[first - 1]) * (key[first] - key[first - 1]) / (val[first] - val[first - 1]);
            else
                lo = key[first];
            if (last != G - 1)
                hi = key[last] + (val[last] - 0.5) * (key[last + 1] - key[last]) / (val[last] - val[last + 1]);
            else
                hi = key[last];
        }
        lb[int64_t(c) * t + o] = lo;
        ub[int64_t(c) * t + o] = hi;
    }
}

}  // namespace cbo
