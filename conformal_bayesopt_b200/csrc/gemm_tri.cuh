// gemm_tri.cuh -- fp64 tensor-core (DMMA) "TN" GEMM with triangular K-range skipping.
//
//   C[m, n] = sum_{k in range(n-tile)} A[m, k] * B[n, k]
//
// This is the fantasy-update solve of the conformal path: with
// B = L^-1 (lower, explicit inverse of chol(K + noise I)) and A = U (kernel rows, one row per
// candidate) pass 1 gives V = U L^-T, i.e. v = L^-1 u for every candidate (tri = lower: only
// k < end of the n-tile contributes); with B = L^-T and A = V pass 2 gives w = L^-T v = A^-1 u
// (tri = upper: only k >= begin of the n-tile).  It replaces what the reference does through
// gp.condition_on_observations -> get_fantasy_strategy (cholesky_solve of the new kernel column,
// conformalbo/helpers.py:243) -- there once per (candidate, grid point), here once per candidate.
//
// Design (B200, sm_100a): fp64 has no tcgen05 kind; the fp64 tensor path on Blackwell is
// mma.sync m8n8k4 (SASS DMMA.8x8x4), measured at 37.0 TFLOP/s = the whole fp64 pipe
// (profiles/r01_fp64_peak.log).  CTA tile 128 x 128, BK = 16, 4-stage cp.async + mbarrier ring into
// padded shared memory (row stride 20 doubles: conflict-free 64-bit fragment loads), 8 warps as
// 2 (m) x 4 (n), warp tile 64 x 32 = 8 x 4 DMMA fragments, 64 fp64 accumulators per thread.
// Tiles are issued longest-K first so the triangular imbalance fills the tail of the grid.
#pragma once
#include "common.cuh"

namespace cbo {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 128;
constexpr int GEMM_BK = 16;
constexpr int GEMM_LDS = GEMM_BK + 4;       // padded row stride (doubles)
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_PREFETCH = 2;         // tiles in flight; STAGES - PREFETCH = tiles of slack between warps
constexpr int GEMM_THREADS = 256;
constexpr int GEMM_MGROUP = 64;          // m-tiles per L2 group (see the tile order in the kernel)

enum { TRI_FULL = 0, TRI_LOWER = 1, TRI_UPPER = 2 };

template <int BM>
__device__ __forceinline__ void gemm_load_stage(double* sA, double* sB, const double* __restrict__ A, int64_t lda,
                                                const double* __restrict__ B, int64_t ldb, int m0, int n0, int k0,
                                                int M, int tid) {
    // A tile: BM rows x 16 doubles = BM * 8 16-byte chunks; B tile: 128 rows = 1024 chunks (4 per thread).
    constexpr int A_CHUNKS = BM * GEMM_BK / 2;
#pragma unroll
    for (int it = 0; it < (A_CHUNKS + GEMM_THREADS - 1) / GEMM_THREADS; ++it) {
        int chunk = tid + it * GEMM_THREADS;
        if (A_CHUNKS % GEMM_THREADS == 0 || chunk < A_CHUNKS) {
            int r = chunk >> 3, cc = chunk & 7;
            int gr = m0 + r;
            gr = gr < M ? gr : M - 1;             // clamp: rows beyond M are computed but never stored
            cp_async16(sA + r * GEMM_LDS + 2 * cc, A + int64_t(gr) * lda + k0 + 2 * cc);
        }
    }
#pragma unroll
    for (int it = 0; it < (GEMM_BN * GEMM_BK / 2) / GEMM_THREADS; ++it) {
        int chunk = tid + it * GEMM_THREADS;
        int r = chunk >> 3, cc = chunk & 7;
        cp_async16(sB + r * GEMM_LDS + 2 * cc, B + int64_t(n0 + r) * ldb + k0 + 2 * cc);
    }
}

// MF = 8-row DMMA fragments per warp in m: CTA tile (16 MF) x 128.  MF = 8 is the throughput configuration; smaller MF
// keep all SMs busy (and cut the wasted rows) when there are only a few hundred candidate rows.
// STAGES / PER_SM: ring depth and co-resident CTAs per SM.  <8, 4, 1> is the 128 x 128 tile alone on its SM (160 KB ring, 202
// registers) -- the configuration in use.  <4, 3, 2>, TWO 64 x 128 tiles per SM (92 KB rings, 128 registers) so that one CTA's
// MMAs cover the other's launch, ring fill and epilogue stores (the 4.3 us of fixed cost per tile), was measured and is 5 %
// SLOWER (2.280 against 2.170 ms per triangular pass of 65536 rows: the smaller tile's main loop loses more than the overlap
// wins; profiles/r02_gemm_shape_time.log).
template <int MF, int STAGES = GEMM_STAGES, int PER_SM = 1>
__global__ void __launch_bounds__(GEMM_THREADS, PER_SM)
k_gemm_tn_f64(const double* __restrict__ A, int64_t lda, const double* __restrict__ B, int64_t ldb,
              double* __restrict__ C, int64_t ldc, int M, int Nn, int K, int tri,
              double* __restrict__ rownorm_partial, int splits, int64_t part_stride) {
    extern __shared__ __align__(16) double smem[];
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    // 2 x 4 warps.  Warps w and w+4 share an SM sub-partition (w % 4): give them n-slices wn and 3-wn so that the
    // triangular skipping on the diagonal block (below) loads every sub-partition's fp64 pipe equally.
    const int wm = warp >> 2, wn = (warp < 4) ? warp : 7 - warp;
    constexpr int BM = 16 * MF;
    constexpr int SD = (BM + GEMM_BN) * GEMM_LDS;          // doubles per ring stage (A tile + B tile, padded rows)
    constexpr int PF = (STAGES > GEMM_PREFETCH) ? GEMM_PREFETCH : STAGES - 1;
    const int tiles_m = (M + BM - 1) / BM;
    const int tiles_n = Nn / GEMM_BN;
    // longest-K tiles first: lower -> last n-tile first; upper / full -> first n-tile first
    // ... within groups of GEMM_MGROUP m-tiles: a group's A rows (64 x 128 rows = 64 MB at K = 1024) stay in L2 while its
    // n-tiles are swept, so A is read from HBM once instead of once per n-tile
    const int t = blockIdx.x;
    const int grp = t / (GEMM_MGROUP * tiles_n);
    const int gm0 = grp * GEMM_MGROUP;
    const int gsz = min(GEMM_MGROUP, tiles_m - gm0);                  // m-tiles in this group
    const int tg = t - grp * GEMM_MGROUP * tiles_n;
    const int n_ord = tg / gsz;
    const int n_tile = (tri == TRI_LOWER) ? (tiles_n - 1 - n_ord) : n_ord;
    const int m_tile = gm0 + tg - n_ord * gsz;
    const int m0 = m_tile * BM, n0 = n_tile * GEMM_BN;
    int k_begin = 0, k_end = K;
    if (tri == TRI_LOWER) k_end = min(K, n0 + GEMM_BN);
    if (tri == TRI_UPPER) k_begin = n0;
    if (splits > 1) {          // split-K: this CTA owns the k-tiles [kt_lo, kt_hi) of the tile's range and its own partial C
        const int nk_all = (k_end - k_begin) / GEMM_BK, per = (nk_all + splits - 1) / splits;
        const int kt_lo = min(nk_all, int(blockIdx.y) * per), kt_hi = min(nk_all, kt_lo + per);
        if (tri == TRI_UPPER) { const int e0 = k_end; k_end = e0 - kt_lo * GEMM_BK; k_begin = e0 - kt_hi * GEMM_BK; }
        else { const int b0 = k_begin; k_begin = b0 + kt_lo * GEMM_BK; k_end = b0 + kt_hi * GEMM_BK; }
        C += int64_t(blockIdx.y) * part_stride;
    }
    const int nk = (k_end - k_begin) / GEMM_BK;
    // K is walked so that the diagonal block (where half of the warps skip and the loads are not covered by MMAs) comes
    // LAST, when the prefetch ring is full: ascending for lower, descending for upper.
    auto kpos = [&](int kt) { return (tri == TRI_UPPER) ? k_end - (kt + 1) * GEMM_BK : k_begin + kt * GEMM_BK; };

    double acc[MF][4][2];
#pragma unroll
    for (int i = 0; i < MF; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // Warp-decoupled pipeline: no CTA-wide barrier in the main loop.  full[s] completes when all 256 threads' cp.async
    // copies of the tile in stage s have landed (cp.async.mbarrier.arrive); empty[s] completes when all 8 warps have
    // consumed it.  Tiles are prefetched PREFETCH ahead into a ring of STAGES, so warps may drift apart by
    // STAGES - PREFETCH tiles (absorbs the triangular imbalance on the diagonal block and scheduling jitter).
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + size_t(STAGES) * SD);
    uint64_t* empty = full + STAGES;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], GEMM_THREADS); mbar_init(&empty[s], GEMM_THREADS / 32); }
    }
    __syncthreads();
#pragma unroll
    for (int s = 0; s < PF; ++s) {
        if (s < nk) {
            double* sA = smem + s * SD;
            gemm_load_stage<BM>(sA, sA + BM * GEMM_LDS, A, lda, B, ldb, m0, n0, kpos(s), M, tid);
            cp_async_mbar_arrive(&full[s]);
        }
    }

    const int fr = lane >> 2, fk = lane & 3;                 // fragment row / k within the 8x4 fragment
    for (int kt = 0; kt < nk; ++kt) {
        {   // producer side: tile kt + PREFETCH goes into the stage last used by tile kt + PREFETCH - STAGES
            const int nt = kt + PF;
            if (nt < nk) {
                const int st = nt % STAGES;
                mbar_wait(&empty[st], ((nt / STAGES) & 1) ^ 1);
                double* sA = smem + st * SD;
                gemm_load_stage<BM>(sA, sA + BM * GEMM_LDS, A, lda, B, ldb, m0, n0, kpos(nt), M, tid);
                cp_async_mbar_arrive(&full[st]);
            }
        }
        const int st = kt % STAGES;
        // On the diagonal 128 x 128 block of the triangular factor this warp's 32 columns see only zeros for part of
        // the K range: lower  -> k < n_warp_end,  upper -> k + BK > n_warp_begin.  Skip those MMAs (warp-uniform).
        const int k0 = kpos(kt);
        const bool skip = (tri == TRI_LOWER && k0 >= n0 + 32 * (wn + 1)) || (tri == TRI_UPPER && k0 + GEMM_BK <= n0 + 32 * wn);
        if (!skip) {
            mbar_wait(&full[st], (kt / STAGES) & 1);
            const double* sA = smem + st * SD + (wm * (8 * MF) + fr) * GEMM_LDS + fk;
            const double* sB = smem + st * SD + BM * GEMM_LDS + (wn * 32 + fr) * GEMM_LDS + fk;
            // rel = position of this k-tile inside the warp's own 32 x 32 diagonal sub-block (0 or 16), else outside [0, 32)
            const int rel = (tri == TRI_FULL) ? -1 : k0 - (n0 + 32 * wn);
            if (rel >= 0 && rel < 32) {
                // the two k-tiles that cross the diagonal: 8-column fragment j sees only zeros at k-step ks when
                // lower: k0 + 4 ks >= n_frag + 8,  upper: k0 + 4 ks + 4 <= n_frag  -- skip those MMAs too (12 of 32 here)
#pragma unroll
                for (int ks = 0; ks < GEMM_BK / 4; ++ks) {
                    double a[MF], b[4];
#pragma unroll
                    for (int i = 0; i < MF; ++i) a[i] = sA[i * 8 * GEMM_LDS + ks * 4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) b[j] = sB[j * 8 * GEMM_LDS + ks * 4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool live = (tri == TRI_LOWER) ? (rel + 4 * ks < 8 * j + 8) : (rel + 4 * ks + 4 > 8 * j);
                        if (live) {
#pragma unroll
                            for (int i = 0; i < MF; ++i) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                        }
                    }
                }
            } else {
#pragma unroll
                for (int ks = 0; ks < GEMM_BK / 4; ++ks) {
                    double a[MF], b[4];
#pragma unroll
                    for (int i = 0; i < MF; ++i) a[i] = sA[i * 8 * GEMM_LDS + ks * 4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) b[j] = sB[j * 8 * GEMM_LDS + ks * 4];
#pragma unroll
                    for (int i = 0; i < MF; ++i)
#pragma unroll
                        for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], a[i], b[j]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&empty[st]);              // this warp is done with the stage (also when it skipped)
    }
    cp_async_wait<0>();
    __syncthreads();

    // epilogue: C tile; each thread owns rows (wm*8*MF + i*8 + fr), cols (wn*32 + j*8 + 2*fk + {0,1})
    double rn[MF];
#pragma unroll
    for (int i = 0; i < MF; ++i) {
        int row = m0 + wm * (8 * MF) + i * 8 + fr;
        double s2 = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s2 = fma(acc[i][j][0], acc[i][j][0], s2);
            s2 = fma(acc[i][j][1], acc[i][j][1], s2);
            if (row < M) {
                double2 v = make_double2(acc[i][j][0], acc[i][j][1]);
                *reinterpret_cast<double2*>(C + int64_t(row) * ldc + n0 + wn * 32 + j * 8 + 2 * fk) = v;
            }
        }
        rn[i] = s2;
    }
    if (rownorm_partial != nullptr) {
        __syncthreads();                                      // smem reuse
        double* red = smem;                                   // [BM rows][4 n-warps]
#pragma unroll
        for (int i = 0; i < MF; ++i) {
            double s2 = rn[i];
            s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
            s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
            if (fk == 0) red[(wm * (8 * MF) + i * 8 + fr) * 4 + wn] = s2;
        }
        __syncthreads();
        if (tid < BM) {
            int row = m0 + tid;
            if (row < M) {
                double s2 = (red[tid * 4 + 0] + red[tid * 4 + 1]) + (red[tid * 4 + 2] + red[tid * 4 + 3]);
                rownorm_partial[int64_t(row) * tiles_n + n_tile] = s2;
            }
        }
    }
}

__host__ __device__ constexpr size_t gemm_smem_bytes(int MF, int stages) {
    return size_t(stages) * (16 * MF + GEMM_BN) * GEMM_LDS * sizeof(double) + 2 * size_t(stages) * sizeof(uint64_t);
}

template <int MF, int STAGES = GEMM_STAGES, int PER_SM = 1>
inline cudaError_t launch_gemm_tn_t(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                                    int M, int Nn, int K, int tri, double* rownorm_partial, cudaStream_t stream,
                                    int splits = 1, int64_t part_stride = 0) {
    constexpr size_t smem = gemm_smem_bytes(MF, STAGES);
    static SmemOptIn optin;                      // per (function, device)
    if (cudaError_t e = optin.ensure(k_gemm_tn_f64<MF, STAGES, PER_SM>, smem); e != cudaSuccess) return e;
    int tiles_m = (M + 16 * MF - 1) / (16 * MF), tiles_n = Nn / GEMM_BN;
    CBO_LAUNCHED(), k_gemm_tn_f64<MF, STAGES, PER_SM><<<dim3(tiles_m * tiles_n, splits), GEMM_THREADS, smem, stream>>>(
        A, lda, B, ldb, C, ldc, M, Nn, K, tri, rownorm_partial, splits, part_stride);
    return cudaGetLastError();
}

// ---- split-K for a handful of rows (M <= 32: the SGLD regime) ---------------------------------------------------------------
// The tiled kernel alone would run 8 CTAs through up to 64 sequential k-tiles (~35 us of latency).  With blockIdx.y = split
// index every (n-tile, split) CTA walks 1/splits of the tile's k-tiles and writes its partial C tile; k_splitk_reduce sums
// the partials in a fixed order (deterministic) and k_rownorm_tiles produces the per-tile row norms.
// rownorm_partial[row][tile] = sum over the tile's 128 columns of C[row][n]^2 (one warp per (row, tile), fixed order)
__global__ void k_rownorm_tiles(const double* __restrict__ C, int64_t ldc, int M, int tiles_n, double* __restrict__ rn) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= M * tiles_n) return;
    const int row = w / tiles_n, tile = w - row * tiles_n;
    const double* c = C + int64_t(row) * ldc + tile * GEMM_BN;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < GEMM_BN / 32; ++j) { double v = c[lane + 32 * j]; s = fma(v, v, s); }
    s = warp_sum(s);
    if (lane == 0) rn[int64_t(row) * tiles_n + tile] = s;
}

// C[m][n] = sum_s part[s][m][n]  (fixed order)
__global__ void k_splitk_reduce(const double* __restrict__ part, int splits, int64_t part_stride, int64_t ldp,
                                double* __restrict__ C, int64_t ldc, int M, int Nn) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= M * Nn) return;
    const int m = e / Nn, n = e - m * Nn;
    double s = 0.0;
    for (int sp = 0; sp < splits; ++sp) s += part[sp * part_stride + int64_t(m) * ldp + n];
    C[int64_t(m) * ldc + n] = s;
}

constexpr int GEMM_SPLITK = 8;
constexpr int GEMM_SPLITK_MAX_M = 32;
inline size_t gemm_splitk_scratch_doubles(int Nn) { return size_t(GEMM_SPLITK) * GEMM_SPLITK_MAX_M * Nn; }

inline cudaError_t launch_gemm_tn(const double* A, int64_t lda, const double* B, int64_t ldb, double* C, int64_t ldc,
                                  int M, int Nn, int K, int tri, double* rownorm_partial, cudaStream_t stream,
                                  int num_sms = 148, double* splitk_scratch = nullptr) {
    // split-K is decided per CALL by the API (all rows of the call <= 32, scratch given), never per chunk: chunking stays
    // bit-invariant
    if (splitk_scratch != nullptr && M <= GEMM_SPLITK_MAX_M) {
        const int64_t ps = int64_t(GEMM_SPLITK_MAX_M) * Nn;
        cudaError_t e = (M > 16)
            ? launch_gemm_tn_t<2>(A, lda, B, ldb, splitk_scratch, Nn, M, Nn, K, tri, nullptr, stream, GEMM_SPLITK, ps)
            : launch_gemm_tn_t<1>(A, lda, B, ldb, splitk_scratch, Nn, M, Nn, K, tri, nullptr, stream, GEMM_SPLITK, ps);
        if (e != cudaSuccess) return e;
        CBO_LAUNCHED(), k_splitk_reduce<<<(M * Nn + 255) / 256, 256, 0, stream>>>(splitk_scratch, GEMM_SPLITK, ps, Nn, C, ldc, M, Nn);
        if (rownorm_partial != nullptr)
            CBO_LAUNCHED(), k_rownorm_tiles<<<(M * (Nn / GEMM_BN) + 7) / 8, 256, 0, stream>>>(C, ldc, M, Nn / GEMM_BN, rownorm_partial);
        return cudaGetLastError();
    }
    // pick the largest m-tile that still yields ~2 CTAs per SM; few rows -> small tiles (latency, no wasted rows)
    const int tiles_n = Nn / GEMM_BN;
    auto ctas = [&](int bm) { return ((M + bm - 1) / bm) * tiles_n; };
    if (ctas(128) >= 2 * num_sms || M > 2048)   // (launch_gemm_tn_t<4, 3, 2>, two 64 x 128 tiles per SM, was measured here: 5 % slower)
        return launch_gemm_tn_t<8>(A, lda, B, ldb, C, ldc, M, Nn, K, tri, rownorm_partial, stream);
    if (ctas(64) >= 2 * num_sms || M > 64)
        return launch_gemm_tn_t<4>(A, lda, B, ldb, C, ldc, M, Nn, K, tri, rownorm_partial, stream);
    if (M > 16)
        return launch_gemm_tn_t<2>(A, lda, B, ldb, C, ldc, M, Nn, K, tri, rownorm_partial, stream);
    return launch_gemm_tn_t<1>(A, lda, B, ldb, C, ldc, M, Nn, K, tri, rownorm_partial, stream);
}

}  // namespace cbo
