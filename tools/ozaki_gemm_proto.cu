// Prototype for the next round, NOT linked into libcbo: C (M x N, fp64) = A (M x K) * B (N x K)^T,
// optionally restricted to k < (n-tile end) like the triangular pass of the hot path, computed from 7-bit integer slices on the
// int8 tensor path (tcgen05.mma kind::i8, int32 accumulators in TMEM) -- the Ozaki scheme with fixed-point slices:
//   x = 2^e * sum_i d_i 2^(-7 i),  d_i in [-64, 64]   (A: one global exponent, B: one exponent per row; all steps exact in fp64)
//   C[m][n] = 2^(eA + eB[n]) * sum_t 2^(-7 t) D_t[m][n],   D_t = sum_{i+j=t} A_i B_j^T   (exact int32),  t = 2 .. S+1
// One CTA per 128 x 128 output tile; the S anti-diagonals are accumulated four at a time (4 x 128 TMEM columns), i.e. K is
// walked ceil(S/4) times.  Operands are staged by plain loads into the K-major no-swizzle canonical layout (see
// i8_mma_probe.cu); v0 has no load/MMA overlap -- it exists to pin the arithmetic and to give a first rate.
// Build: nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/ozaki_gemm_proto tools/ozaki_gemm_proto.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ unsigned long long g_t[4] = {~0ull, 0ull, ~0ull, 0ull};     // [spin start, spin end, gemm start, gemm end] (globaltimer ns)
__device__ __forceinline__ unsigned long long gtime() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }

constexpr int S = 8;            // slices per operand (8 x 7 bits = 56 bits)
constexpr int TM = 128, TN = 128, KB = 32, THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | ((uint64_t)1 << 46);
}
// K-major operand in a TMA-swizzled layout: rows of `row_bytes` (32 or 64) bytes, 8-row groups `8 * row_bytes` apart
__device__ __forceinline__ uint64_t make_desc_sw(uint32_t saddr, uint32_t row_bytes) {
    const uint64_t layout = row_bytes == 64 ? 4 : 6;                 // SWIZZLE_64B = 4, SWIZZLE_32B = 6
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)1 << 16) | ((uint64_t)(((8 * row_bytes) >> 4) & 0x3fff) << 32) |
           ((uint64_t)1 << 46) | (layout << 61);
}
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    for (long it = 0; it < 4000000L; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// ---- slicing: rows x K fp64 (ld) -> S planes of rows x K int8; per-row exponent (row_exp) or one global exponent -------------
__global__ void k_row_exponent(const double* __restrict__ X, int rows, int K, int ld, int* __restrict__ row_exp) {
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= rows) return;
    double m = 0.0;
    for (int k = lane; k < K; k += 32) m = fmax(m, fabs(X[(size_t)r * ld + k]));
    for (int o = 16; o; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) {
        int e = 0;
        if (m > 0.0) { frexp(m, &e); e += 1; }          // |x| < 2^(e-1): x 2^-e in (-0.5, 0.5)
        row_exp[r] = e;
    }
}

__global__ void k_slice(const double* __restrict__ X, int rows, int K, int ld, const int* __restrict__ row_exp, int global_exp,
                        int8_t* __restrict__ out) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (size_t)rows * K) return;
    const int r = (int)(idx / K), k = (int)(idx % K);
    const int e = row_exp ? row_exp[r] : global_exp;
    double x = ldexp(X[(size_t)r * ld + k], -e);        // exact
#pragma unroll
    for (int s = 0; s < S; ++s) {
        x *= 128.0;
        const double d = rint(x);
        x -= d;                                         // exact: |x| <= 0.5 again
        out[(size_t)s * rows * K + idx] = (int8_t)(int)d;
    }
}

// ---- the slice-product GEMM ---------------------------------------------------------------------------------------------------
// As: S x M x K int8, Bs: S x N x K int8.  tri != 0: only k < (n-tile end) contributes (B is lower triangular: B[n][k] = 0 for k > n).
__global__ void __launch_bounds__(THREADS)
k_ozaki_gemm(const int8_t* __restrict__ As, const int8_t* __restrict__ Bs, int M, int N, int K, int tri, int expA,
             const int* __restrict__ expB, double* __restrict__ C, int ldc, int* __restrict__ fail) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_n = N / TN;
    const int tile_m = blockIdx.x / tiles_n, tile_n = blockIdx.x % tiles_n;
    const int m0 = tile_m * TM, n0 = tile_n * TN;
    const int k_end = tri ? min(K, n0 + TN) : K;
    uint8_t* sA = smem;                                  // S planes x (KB/16 chunks x 128 rows x 16 B)
    uint8_t* sB = smem + (size_t)S * TM * KB;
    constexpr int PLANE = TM * KB;                       // bytes of one slice plane of a stage (TM == TN)
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = make_idesc_i8(TM, TN);
    uint32_t parity = 0;
    bool ok = true;
    long long t_load = 0, t_mma = 0, t_epi = 0, tc0;

    for (int pass = 0; pass < (S + 3) / 4; ++pass) {
        const int t_lo = 2 + 4 * pass, t_hi = min(S + 1, t_lo + 3);      // anti-diagonals of this pass
        const int s_max = t_hi - 1;                                      // slices 1..s_max are needed (1-based)
        for (int k0 = 0; k0 < k_end; k0 += KB) {
            // stage slices 1..s_max of A (rows m0..) and B (rows n0..), columns k0..k0+KB, into the canonical layout
            tc0 = clock64();
            const int chunks = KB / 16;
            for (int o = tid; o < 2 * s_max * TM * chunks; o += THREADS) {
                const int c = o % chunks, r = (o / chunks) % TM, pl = (o / (chunks * TM)) % s_max, which = o / (chunks * TM * s_max);
                const int8_t* src = which ? Bs + ((size_t)pl * N + n0 + r) * K + k0 + c * 16 : As + ((size_t)pl * M + m0 + r) * K + k0 + c * 16;
                uint8_t* dst = (which ? sB : sA) + (size_t)pl * PLANE + (size_t)c * TM * 16 + r * 16;
                *reinterpret_cast<int4*>(dst) = *reinterpret_cast<const int4*>(src);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncthreads();
            t_load += clock64() - tc0; tc0 = clock64();
            if (tid == 0) {
                asm volatile("tcgen05.fence::after_thread_sync;");
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
                for (int t = t_lo; t <= t_hi; ++t)
                    for (int i = 1; i < t; ++i) {
                        const int j = t - i;
                        if (i > S || j > S) continue;
                        const uint64_t ad = make_desc(a0 + (uint32_t)(i - 1) * PLANE, TM * 16, 128);
                        const uint64_t bd = make_desc(b0 + (uint32_t)(j - 1) * PLANE, TN * 16, 128);
                        mma_i8(tmem + (uint32_t)(t - t_lo) * TN, ad, bd, idesc, (k0 > 0) || (i > 1));
                    }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            ok = mbar_wait_bounded(smem_u32(&bar), parity) && ok;        // v0: no overlap of loads and MMAs
            parity ^= 1;
            t_mma += clock64() - tc0;
            if (!ok) break;
        }
        if (!ok) break;
        // epilogue of this pass: C (+)= 2^(eA + eB[n]) sum_t 2^(-7 t) D_t ; warps 0..3 own the four 32-lane TMEM quadrants
        asm volatile("tcgen05.fence::after_thread_sync;");
        tc0 = clock64();
        if (warp < 4) {
            const int row = m0 + warp * 32 + lane;
            for (int c0 = 0; c0 < TN; c0 += 32) {
                double acc[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0.0;
                for (int t = t_hi; t >= t_lo; --t) {                     // small terms first
                    uint32_t v[32];
                    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t - t_lo) * TN + c0;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(addr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const double w = ldexp(1.0, -7 * t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = fma((double)(int32_t)v[j], w, acc[j]);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = n0 + c0 + j;
                    const double val = ldexp(acc[j], expA + expB[col]);
                    double* dst = C + (size_t)row * ldc + col;
                    *dst = pass == 0 ? val : *dst + val;                 // later passes add the lower-order diagonals
                }
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();                                                 // TMEM is reused by the next pass
        asm volatile("tcgen05.fence::after_thread_sync;");
        t_epi += clock64() - tc0;
    }
    if (blockIdx.x == 0 && tid == 0) printf("  [v0 CTA 0] cycles: staging %lld, MMA issue+wait %lld, epilogue %lld\n", t_load, t_mma, t_epi);
    if (!ok && tid == 0) atomicExch(fail, 1);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- v1: warp-specialised pipeline (loaders -> MMA issuer -> epilogue) ----------------------------------------------------------
// warps 0..3 epilogue (TMEM lane quadrants), warp 4 issues the MMAs, warps 5..12 stage the slices (plain loads) into a ring of
// NSTAGE shared-memory stages; full[] / empty[] mbarriers per stage, acc_full / acc_empty around the TMEM accumulators.
constexpr int NSTAGE = 3, LOADERS = 256, V1_THREADS = 160 + LOADERS;

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__global__ void __launch_bounds__(V1_THREADS, 1)
k_ozaki_gemm_v1(const int8_t* __restrict__ As, const int8_t* __restrict__ Bs, int M, int N, int K, int tri, int expA,
                const int* __restrict__ expB, double* __restrict__ C, int ldc, int* __restrict__ fail) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[NSTAGE], empty[NSTAGE], acc_full, acc_empty;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_n = N / TN;
    // longest-K tiles first in the triangular case
    // ... within groups of 16 m-tiles, so that a group's A slices stay in L2 while its n-tiles are swept
    const int grp = blockIdx.x / (16 * tiles_n), rem = blockIdx.x % (16 * tiles_n);
    const int tile_m = grp * 16 + rem % 16, n_ord = rem / 16;
    const int tile_n = tri ? tiles_n - 1 - n_ord : n_ord;
    const int m0 = tile_m * TM, n0 = tile_n * TN;
    const int k_end = tri ? min(K, n0 + TN) : K;
    const int nk = k_end / KB;
    constexpr int PLANE = TM * KB, STAGE = 2 * S * PLANE;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&full[s])), "r"(LOADERS));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&acc_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(&acc_empty)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    constexpr int NPASS = (S + 3) / 4;
    bool ok = true;

    if (warp >= 5) {
        // ===== loaders =====
        const int lt = tid - 160;
        int it = 0;
        for (int pass = 0; pass < NPASS && ok; ++pass) {
            const int s_max = min(S + 1, 2 + 4 * pass + 3) - 1;
            for (int kt = 0; kt < nk && ok; ++kt, ++it) {
                const int st = it % NSTAGE;
                if (it >= NSTAGE) ok = mbar_wait_bounded(smem_u32(&empty[st]), ((it / NSTAGE) & 1) ^ 1);
                uint8_t* sA = smem + (size_t)st * STAGE;
                uint8_t* sB = sA + S * PLANE;
                const int k0 = kt * KB;
                constexpr int chunks = KB / 16;
                const int total = 2 * s_max * TM * chunks;
                for (int o0 = lt; o0 < total; o0 += LOADERS * 8) {
                    int4 v[8];
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int o = o0 + u * LOADERS;
                        if (o < total) {
                            const int c = o % chunks, r = (o / chunks) % TM, pl = (o / (chunks * TM)) % s_max, which = o / (chunks * TM * s_max);
                            const int8_t* src = which ? Bs + ((size_t)pl * N + n0 + r) * K + k0 + c * 16 : As + ((size_t)pl * M + m0 + r) * K + k0 + c * 16;
                            v[u] = *reinterpret_cast<const int4*>(src);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
                        const int o = o0 + u * LOADERS;
                        if (o < total) {
                            const int c = o % chunks, r = (o / chunks) % TM, pl = (o / (chunks * TM)) % s_max, which = o / (chunks * TM * s_max);
                            *reinterpret_cast<int4*>((which ? sB : sA) + (size_t)pl * PLANE + (size_t)c * TM * 16 + r * 16) = v[u];
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_arrive(smem_u32(&full[st]));
            }
        }
    } else if (warp == 4) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc_i8(TM, TN);
        int it = 0;
        for (int pass = 0; pass < NPASS && ok; ++pass) {
            const int t_lo = 2 + 4 * pass, t_hi = min(S + 1, t_lo + 3);
            if (pass > 0) ok = mbar_wait_bounded(smem_u32(&acc_empty), (pass - 1) & 1);     // epilogue has drained TMEM
            asm volatile("tcgen05.fence::after_thread_sync;");
            for (int kt = 0; kt < nk && ok; ++kt, ++it) {
                const int st = it % NSTAGE;
                ok = mbar_wait_bounded(smem_u32(&full[st]), (it / NSTAGE) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (lane == 0 && ok) {
                    const uint32_t a0 = smem_u32(smem + (size_t)st * STAGE), b0 = a0 + S * PLANE;
                    for (int t = t_lo; t <= t_hi; ++t)
                        for (int i = 1; i < t; ++i) {
                            const int j = t - i;
                            if (i > S || j > S) continue;
                            mma_i8(tmem + (uint32_t)(t - t_lo) * TN, make_desc(a0 + (uint32_t)(i - 1) * PLANE, TM * 16, 128),
                                   make_desc(b0 + (uint32_t)(j - 1) * PLANE, TN * 16, 128), idesc, (kt > 0) || (i > 1));
                        }
                    tc_commit(smem_u32(&empty[st]));
                    if (kt == nk - 1) tc_commit(smem_u32(&acc_full));
                }
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue warps 0..3 =====
        for (int pass = 0; pass < NPASS && ok; ++pass) {
            const int t_lo = 2 + 4 * pass, t_hi = min(S + 1, t_lo + 3);
            ok = mbar_wait_bounded(smem_u32(&acc_full), pass & 1);
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (!ok) break;
            const int row = m0 + warp * 32 + lane;
            for (int c0 = 0; c0 < TN; c0 += 32) {
                double acc[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0.0;
                for (int t = t_hi; t >= t_lo; --t) {
                    uint32_t v[32];
                    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t - t_lo) * TN + c0;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(addr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const double w = ldexp(1.0, -7 * t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = fma((double)(int32_t)v[j], w, acc[j]);
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int col = n0 + c0 + j;
                    const double val = ldexp(acc[j], expA + expB[col]);
                    double* dst = C + (size_t)row * ldc + col;
                    *dst = pass == 0 ? val : *dst + val;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            mbar_arrive(smem_u32(&acc_empty));
        }
    }
    if (!ok && lane == 0) atomicExch(fail, 1);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

// ---- v2: TMA-fed pipeline -----------------------------------------------------------------------------------------------------
// The slice planes are described to TMA as a 4-D byte tensor (16 B within a K-chunk, row, K-chunk, plane) so that ONE bulk copy
// with box (16, 128, KB/16, 1) lands a plane of a stage directly in the canonical no-swizzle layout [chunk][row][16 B].
// warp 4: TMA producer, warp 5: MMA issuer, warps 0..3: epilogue (coalesced through a padded shared-memory transpose).
constexpr int V2_THREADS = 192, V2_STAGES = 2, V3_THREADS = 320;

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar) : "memory");
}

__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(c2), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(V2_THREADS, 2)
k_ozaki_gemm_v2(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapA8, const __grid_constant__ CUtensorMap mapB8, int M, int N, int K, int tri,
                const double* __restrict__ scale, double* __restrict__ C, int ldc, int* __restrict__ fail) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t full[V2_STAGES], empty[V2_STAGES], acc_full, acc_empty;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_n = N / TN;
    const int grp = blockIdx.x / (16 * tiles_n), rem = blockIdx.x % (16 * tiles_n);
    const int tile_m = grp * 16 + rem % 16, n_ord = rem / 16;
    const int tile_n = tri ? tiles_n - 1 - n_ord : n_ord;
    const int m0 = tile_m * TM, n0 = tile_n * TN;
    const int k_end = tri ? min(K, n0 + TN) : K;
    const int nk = k_end / KB;
    constexpr int PLANE = TM * KB, STAGE = 2 * S * PLANE;
    double* s_tr = reinterpret_cast<double*>(smem + (size_t)V2_STAGES * STAGE);      // 4 warps x 32 x 33 doubles
    if (tid == 0) atomicMin(&g_t[2], gtime());
    if (tid == 0) {
        for (int s = 0; s < V2_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&acc_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 128;" ::"r"(smem_u32(&acc_empty)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    constexpr int NPASS = (S + 3) / 4;
    bool ok = true;

    if (warp == 4) {
        if (lane == 0) {            // ===== TMA producer =====
            int it = 0;
            for (int pass = 0; pass < NPASS && ok; ++pass) {
                const int s_max = min(S + 1, 2 + 4 * pass + 3) - 1;
                // the first pass needs 4 planes per operand: it takes 64 K-columns per stage, the second (8 planes) 32 -- both
                // fill the 64 KB stage, and a first-pass stage then carries 20 MMAs, enough to hide the copy latency
                const int kb = (s_max == 4) ? 2 * KB : KB, plane = TM * kb;
                for (int kt = 0; kt < k_end / kb && ok; ++kt, ++it) {
                    const int st = it % V2_STAGES;
                    if (it >= V2_STAGES) ok = mbar_wait_bounded(smem_u32(&empty[st]), ((it / V2_STAGES) & 1) ^ 1);
                    const uint32_t bar = smem_u32(&full[st]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(2 * s_max * plane) : "memory");
                    const uint32_t a0 = smem_u32(smem + (size_t)st * STAGE), b0 = a0 + s_max * plane;
                    // one bulk copy per operand: the box spans the 4 (first pass) or 8 (second pass) slice planes of the stage
                    tma_load_3d(a0, s_max == 4 ? &mapA : &mapA8, kt * kb, m0, 0, bar);
                    tma_load_3d(b0, s_max == 4 ? &mapB : &mapB8, kt * kb, n0, 0, bar);
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc_i8(TM, TN);
        int it = 0;
        long long w_full = 0, w_acc = 0, t_begin = clock64();
        for (int pass = 0; pass < NPASS && ok; ++pass) {
            const int t_lo = 2 + 4 * pass, t_hi = min(S + 1, t_lo + 3);
            long long c0_ = clock64();
            if (pass > 0) ok = mbar_wait_bounded(smem_u32(&acc_empty), (pass - 1) & 1);
            w_acc += clock64() - c0_;
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int s_max = t_hi - 1, kb = (s_max == 4) ? 2 * KB : KB, plane = TM * kb, nkp = k_end / kb;
            for (int kt = 0; kt < nkp && ok; ++kt, ++it) {
                const int st = it % V2_STAGES;
                c0_ = clock64();
                ok = mbar_wait_bounded(smem_u32(&full[st]), (it / V2_STAGES) & 1);
                w_full += clock64() - c0_;
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (lane == 0 && ok) {
                    const uint32_t a0 = smem_u32(smem + (size_t)st * STAGE), b0 = a0 + s_max * plane;
                    for (int sub = 0; sub < kb / 32; ++sub)                          // 32 K-columns = 2 chunks per MMA
                        for (int t = t_lo; t <= t_hi; ++t)
                            for (int i = 1; i < t; ++i) {
                                const int j = t - i;
                                if (i > S || j > S) continue;
                                mma_i8(tmem + (uint32_t)(t - t_lo) * TN,
                                       make_desc_sw(a0 + (uint32_t)(i - 1) * plane + sub * 32, kb),
                                       make_desc_sw(b0 + (uint32_t)(j - 1) * plane + sub * 32, kb), idesc,
                                       (kt > 0) || (sub > 0) || (i > 1));
                            }
                    tc_commit(smem_u32(&empty[st]));
                    if (kt == nkp - 1) tc_commit(smem_u32(&acc_full));
                }
                __syncwarp();
            }
        }
        if (blockIdx.x == 0 && lane == 0)
            printf("  [v2 CTA 0, MMA warp] issue loop %lld clk, of which waiting for operands %lld, waiting for the epilogue %lld\n",
                   clock64() - t_begin, w_full, w_acc);
    } else {
        // ===== epilogue warps 0..3 =====
        double* tr = s_tr + warp * 32 * 33;
        long long t_epi = 0, w_accf = 0;
        for (int pass = 0; pass < NPASS && ok; ++pass) {
            const int t_lo = 2 + 4 * pass, t_hi = min(S + 1, t_lo + 3);
            long long c1_ = clock64();
            ok = mbar_wait_bounded(smem_u32(&acc_full), pass & 1);
            w_accf += clock64() - c1_; c1_ = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (!ok) break;
            for (int c0 = 0; c0 < TN; c0 += 32) {
                // the anti-diagonals of a pass are recombined in 64-bit integers (D_t < 2^25, shifts <= 21 bits) and converted
                // once: the epilogue then leaves the fp64 pipe almost alone (co-residency with fp64-bound kernels)
                long long acc[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0;
                for (int t = t_lo; t <= t_hi; ++t) {
                    uint32_t v[32];
                    const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(t - t_lo) * TN + c0;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(addr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = acc[j] * 128 + (long long)(int32_t)v[j];
                }
                const double wlast = ldexp(1.0, -7 * t_hi);
                // transpose through shared memory: lane = row -> lane = column, 256-byte row segments to global memory
#pragma unroll
                for (int j = 0; j < 32; ++j) tr[lane * 33 + j] = (double)acc[j] * wlast;
                __syncwarp();
                const double sc = scale[n0 + c0 + lane];
                double* dst0 = C + (size_t)(m0 + warp * 32) * ldc + n0 + c0 + lane;
                if (pass == 0) {
#pragma unroll 8
                    for (int r = 0; r < 32; ++r) dst0[(size_t)r * ldc] = tr[r * 33 + lane] * sc;
                } else {                                                     // all 32 loads of the earlier partial sums in flight
                    double old[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r) old[r] = dst0[(size_t)r * ldc];
#pragma unroll
                    for (int r = 0; r < 32; ++r) dst0[(size_t)r * ldc] = old[r] + tr[r * 33 + lane] * sc;
                }
                __syncwarp();
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            mbar_arrive(smem_u32(&acc_empty));
            t_epi += clock64() - c1_;
        }
        if (blockIdx.x == 0 && tid == 0)
            printf("  [v2 CTA 0, epilogue warp 0] waiting for accumulators %lld clk, epilogue work %lld clk\n", w_accf, t_epi);
    }
    if (!ok && lane == 0) atomicExch(fail, 1);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    if (tid == 0) atomicMax(&g_t[3], gtime());
}

// ---- v3: the same pipeline with the tile width TN_ and the number of anti-diagonals per pass ACC as parameters ---------------------
// (TN_ = 256, ACC = 2: four passes over K, m128n256k32 MMAs; every bulk copy brings 2 slice planes)
template <int TN_, int ACC>
__global__ void __launch_bounds__(V3_THREADS, 1)
k_ozaki_gemm_v3(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                const __grid_constant__ CUtensorMap mapA8, const __grid_constant__ CUtensorMap mapB8,   /* mapX: 64-byte rows, mapX8: 32-byte rows; 2 planes per box */ int M, int N, int K, int tri,
                const double* __restrict__ scale, double* __restrict__ C, int ldc, int* __restrict__ fail) {
    extern __shared__ __align__(1024) uint8_t smem[];
    constexpr int STAGE = S * (TM + TN_) * KB, V3_STAGES = (TN_ > 128) ? 2 : 3;
    __shared__ __align__(8) uint64_t full[V3_STAGES], empty[V3_STAGES], acc_full, acc_empty;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tiles_n = N / TN_;
    const int grp = blockIdx.x / (16 * tiles_n), rem = blockIdx.x % (16 * tiles_n);
    const int tile_m = grp * 16 + rem % 16, n_ord = rem / 16;
    const int tile_n = tri ? tiles_n - 1 - n_ord : n_ord;
    const int m0 = tile_m * TM, n0 = tile_n * TN_;
    const int k_end = tri ? min(K, n0 + TN_) : K;
    const int nk = k_end / KB;
    double* s_tr = reinterpret_cast<double*>(smem + (size_t)V3_STAGES * STAGE);      // 4 warps x 32 x 33 doubles
    if (tid == 0) atomicMin(&g_t[2], gtime());
    if (tid == 0) {
        for (int s = 0; s < V3_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&empty[s])));
        }
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&acc_full)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 256;" ::"r"(smem_u32(&acc_empty)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    constexpr int NPASS = (S + ACC - 1) / ACC;
    bool ok = true;

    if (warp == 4) {
        if (lane == 0) {            // ===== TMA producer =====
            int it = 0;
            for (int pass = 0; pass < NPASS && ok; ++pass) {
                const int s_max = min(S + 1, 2 + ACC * pass + ACC - 1) - 1;
                // the first pass needs 4 planes per operand: it takes 64 K-columns per stage, the second (8 planes) 32 -- both
                // fill the 64 KB stage, and a first-pass stage then carries 20 MMAs, enough to hide the copy latency
                const int kb = (s_max <= 4) ? 2 * KB : KB, plane = TM * kb, planeB = TN_ * kb;
                for (int kt = 0; kt < k_end / kb && ok; ++kt, ++it) {
                    const int st = it % V3_STAGES;
                    if (it >= V3_STAGES) ok = mbar_wait_bounded(smem_u32(&empty[st]), ((it / V3_STAGES) & 1) ^ 1);
                    const uint32_t bar = smem_u32(&full[st]);
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(s_max * (plane + planeB)) : "memory");
                    const uint32_t a0 = smem_u32(smem + (size_t)st * STAGE), b0 = a0 + s_max * plane;
                    // one bulk copy per operand: the box spans the 4 (first pass) or 8 (second pass) slice planes of the stage
                    for (int pl = 0; pl < s_max; pl += 2) {
                        tma_load_3d(a0 + pl * plane, kb == 64 ? &mapA : &mapA8, kt * kb, m0, pl, bar);
                        tma_load_3d(b0 + pl * planeB, kb == 64 ? &mapB : &mapB8, kt * kb, n0, pl, bar);
                    }
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        const uint32_t idesc = make_idesc_i8(TM, TN_);
        int it = 0;
        long long w_full = 0, w_acc = 0, t_begin = clock64();
        for (int pass = 0; pass < NPASS && ok; ++pass) {
            const int t_lo = 2 + ACC * pass, t_hi = min(S + 1, t_lo + ACC - 1);
            long long c0_ = clock64();
            if (pass > 0) ok = mbar_wait_bounded(smem_u32(&acc_empty), (pass - 1) & 1);
            w_acc += clock64() - c0_;
            asm volatile("tcgen05.fence::after_thread_sync;");
            const int s_max = t_hi - 1, kb = (s_max <= 4) ? 2 * KB : KB, plane = TM * kb, planeB = TN_ * kb, nkp = k_end / kb;
            for (int kt = 0; kt < nkp && ok; ++kt, ++it) {
                const int st = it % V3_STAGES;
                c0_ = clock64();
                ok = mbar_wait_bounded(smem_u32(&full[st]), (it / V3_STAGES) & 1);
                w_full += clock64() - c0_;
                asm volatile("tcgen05.fence::after_thread_sync;");
                if (lane == 0 && ok) {
                    const uint32_t a0 = smem_u32(smem + (size_t)st * STAGE), b0 = a0 + s_max * plane;
                    for (int sub = 0; sub < kb / 32; ++sub)                          // 32 K-columns = 2 chunks per MMA
                        for (int t = t_lo; t <= t_hi; ++t)
                            for (int i = 1; i < t; ++i) {
                                const int j = t - i;
                                if (i > S || j > S) continue;
                                mma_i8(tmem + (uint32_t)(t - t_lo) * TN_,
                                       make_desc_sw(a0 + (uint32_t)(i - 1) * plane + sub * 32, kb),
                                       make_desc_sw(b0 + (uint32_t)(j - 1) * planeB + sub * 32, kb), idesc,
                                       (kt > 0) || (sub > 0) || (i > 1));
                            }
                    tc_commit(smem_u32(&empty[st]));
                    if (kt == nkp - 1) tc_commit(smem_u32(&acc_full));
                }
                __syncwarp();
            }
        }
        if (blockIdx.x == 0 && lane == 0)
            printf("  [v3 CTA 0, MMA warp] issue loop %lld clk, of which waiting for operands %lld, waiting for the epilogue %lld\n",
                   clock64() - t_begin, w_full, w_acc);
    } else {
        // ===== epilogue: warps 0..3 and 6..9; warps w and w' with w % 4 == w' % 4 share a TMEM lane quadrant and split the columns
        const int quad = warp & 3, half = warp < 4 ? 0 : 1;
        double* tr = s_tr + (half * 4 + quad) * 32 * 17;
        long long t_epi = 0, w_accf = 0;
        for (int pass = 0; pass < NPASS && ok; ++pass) {
            const int t_lo = 2 + ACC * pass, t_hi = min(S + 1, t_lo + ACC - 1);
            long long c1_ = clock64();
            ok = mbar_wait_bounded(smem_u32(&acc_full), pass & 1);
            w_accf += clock64() - c1_; c1_ = clock64();
            asm volatile("tcgen05.fence::after_thread_sync;");
            if (!ok) break;
            for (int c0 = half * 32; c0 < TN_; c0 += 64) {
                double acc[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = 0.0;
                for (int t = t_hi; t >= t_lo; --t) {
                    uint32_t v[32];
                    const uint32_t addr = tmem + ((uint32_t)(quad * 32) << 16) + (uint32_t)(t - t_lo) * TN_ + c0;
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                                 : "r"(addr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const double w = ldexp(1.0, -7 * t);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[j] = fma((double)(int32_t)v[j], w, acc[j]);
                }
                // transpose through shared memory in two 16-column halves: lane = row -> (lane & 15) = column, (lane >> 4) = row parity;
                // every store instruction then writes two 128-byte row segments
#pragma unroll
                for (int h = 0; h < 2; ++h) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) tr[lane * 17 + j] = acc[h * 16 + j];
                    __syncwarp();
                    const int col = lane & 15, rsub = lane >> 4;
                    const double sc = scale[n0 + c0 + h * 16 + col];
                    double* dst0 = C + (size_t)(m0 + quad * 32 + rsub) * ldc + n0 + c0 + h * 16 + col;
                    if (pass == 0) {
#pragma unroll
                        for (int r2 = 0; r2 < 16; ++r2) dst0[(size_t)(2 * r2) * ldc] = tr[(2 * r2 + rsub) * 17 + col] * sc;
                    } else {
                        double old[16];
#pragma unroll
                        for (int r2 = 0; r2 < 16; ++r2) old[r2] = dst0[(size_t)(2 * r2) * ldc];
#pragma unroll
                        for (int r2 = 0; r2 < 16; ++r2) dst0[(size_t)(2 * r2) * ldc] = old[r2] + tr[(2 * r2 + rsub) * 17 + col] * sc;
                    }
                    __syncwarp();
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;");
            mbar_arrive(smem_u32(&acc_empty));
            t_epi += clock64() - c1_;
        }
        if (blockIdx.x == 0 && tid == 0)
            printf("  [v3 CTA 0, epilogue warp 0] waiting for accumulators %lld clk, epilogue work %lld clk\n", w_accf, t_epi);
    }
    if (!ok && lane == 0) atomicExch(fail, 1);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
    if (tid == 0) atomicMax(&g_t[3], gtime());
}

__global__ void k_col_scale(const int* __restrict__ expB, int expA, int N, double* __restrict__ scale) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n < N) scale[n] = ldexp(1.0, expA + expB[n]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// 3-D byte tensor (K, rows, planes) with a swizzled box of `kb` (32 / 64) K-bytes x 128 rows x planes: rows of kb bytes in
// shared memory, 16-byte chunks XOR-swizzled the way the UMMA SWIZZLE_32B / _64B layouts expect
static CUtensorMap make_slice_map_sw(const int8_t* base, int rows, int K, int planes_in_box, int kb, int box_rows = 128) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); exit(1); }
    typedef CUresult (*Fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                           const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap m;
    const cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)S};
    const cuuint64_t gstr[2] = {(cuuint64_t)K, (cuuint64_t)rows * K};
    const cuuint32_t box[3] = {(cuuint32_t)kb, (cuuint32_t)box_rows, (cuuint32_t)planes_in_box};
    const cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = ((Fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          kb == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled (swizzled) failed: %d\n", (int)r); exit(1); }
    return m;
}

static CUtensorMap make_slice_map(const int8_t* base, int rows, int K, int planes_in_box = 1, int chunks_in_box = KB / 16) {
    static EncodeTiledFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess) { printf("cuTensorMapEncodeTiled not available\n"); exit(1); }
        encode = (EncodeTiledFn)fn;
    }
    CUtensorMap m;
    const cuuint64_t gdim[4] = {16, (cuuint64_t)rows, (cuuint64_t)(K / 16), (cuuint64_t)S};
    const cuuint64_t gstr[3] = {(cuuint64_t)K, 16, (cuuint64_t)rows * K};           // bytes, dims 1..3
    const cuuint32_t box[4] = {16, 128, (cuuint32_t)chunks_in_box, (cuuint32_t)planes_in_box};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 4, (void*)base, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d\n", (int)r); exit(1); }
    return m;
}

__global__ void k_sleep(int slot, unsigned ns) {
    if (threadIdx.x == 0) atomicMin(&g_t[slot], gtime());
    const unsigned long long t0 = gtime();
    while (gtime() - t0 < ns) __nanosleep(1000);
    if (threadIdx.x == 0) atomicMax(&g_t[slot + 1], gtime());
}

// fp64-pipe load for the co-residency experiment: 8 independent DFMA chains per thread, no shared memory
__global__ void __launch_bounds__(256) k_dfma_spin(int iters, double* out) {
    if (threadIdx.x == 0) atomicMin(&g_t[0], gtime());
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = threadIdx.x * 1e-3 + j;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = fma(a[j], 0.999999, 1e-9);
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 12345.678) out[0] = s;
    if (threadIdx.x == 0) atomicMax(&g_t[1], gtime());
}

static double urand() { return rand() / (double)RAND_MAX; }

int main(int argc, char** argv) {
    const int K = 1024, N = 1024;
    const int Mcheck = 256, Mbig = argc > 1 ? atoi(argv[1]) : 16384;
    srand(7);
    // A like kernel rows (values in [0, 1], wide dynamic range), B like L^-1: lower triangular, entries of mixed sign / magnitude
    std::vector<double> hA((size_t)Mbig * K), hB((size_t)N * K, 0.0);
    for (auto& x : hA) x = exp(-8.0 * urand());
    for (int n = 0; n < N; ++n)
        for (int k = 0; k <= n; ++k) hB[(size_t)n * K + k] = (k == n ? 5.0 + 5.0 * urand() : (urand() - 0.5) * exp(-6.0 * urand()));
    double *dA, *dB, *dC; int8_t *dAs, *dBs; int *dexpB, *dfail;
    CK(cudaMalloc(&dA, hA.size() * 8)); CK(cudaMalloc(&dB, hB.size() * 8)); CK(cudaMalloc(&dC, (size_t)Mbig * N * 8));
    CK(cudaMalloc(&dAs, (size_t)S * Mbig * K)); CK(cudaMalloc(&dBs, (size_t)S * N * K)); CK(cudaMalloc(&dexpB, N * 4)); CK(cudaMalloc(&dfail, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 8, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 8, cudaMemcpyHostToDevice));
    CK(cudaMemset(dfail, 0, 4));
    const int expA = 1;                                   // |a| <= 1 < 2^(expA - 1) ... a 2^-1 in [0, 0.5]
    k_row_exponent<<<(N + 7) / 8, 256>>>(dB, N, K, K, dexpB);
    k_slice<<<(unsigned)(((size_t)N * K + 255) / 256), 256>>>(dB, N, K, K, dexpB, 0, dBs);
    cudaEvent_t e0, e1, e2; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    const size_t smem = (size_t)2 * S * TM * KB;
    CK(cudaFuncSetAttribute(k_ozaki_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(k_ozaki_gemm_v1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(NSTAGE * smem)));
    double* dscale; CK(cudaMalloc(&dscale, N * 8));
    k_col_scale<<<(N + 255) / 256, 256>>>(dexpB, expA, N, dscale);
    static_assert(S == 8, "the v2 tensor maps assume 4 + 8 slice planes per stage");
    const CUtensorMap mapA = make_slice_map_sw(dAs, Mbig, K, 4, 2 * KB), mapB = make_slice_map_sw(dBs, N, K, 4, 2 * KB);
    const CUtensorMap mapA8 = make_slice_map_sw(dAs, Mbig, K, 8, KB), mapB8 = make_slice_map_sw(dBs, N, K, 8, KB);
    const size_t smem_v2 = (size_t)V2_STAGES * smem + 4 * 32 * 33 * 8;
    CK(cudaFuncSetAttribute(k_ozaki_gemm_v2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v2));
    const CUtensorMap m3A64 = make_slice_map_sw(dAs, Mbig, K, 2, 64), m3A32 = make_slice_map_sw(dAs, Mbig, K, 2, 32);
    const CUtensorMap m3B64 = make_slice_map_sw(dBs, N, K, 2, 64, 256), m3B32 = make_slice_map_sw(dBs, N, K, 2, 32, 256);
    const size_t smem_v3 = (size_t)2 * S * (TM + 256) * KB + 8 * 32 * 17 * 8;
    CK(cudaFuncSetAttribute(k_ozaki_gemm_v3<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_v3));
    for (int variant = 0; variant <= 3; ++variant)
    for (int tri = 0; tri <= 1; ++tri) {
        CK(cudaMemset(dC, 0, (size_t)Mbig * N * 8));
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0));
            k_slice<<<(unsigned)(((size_t)Mbig * K + 255) / 256), 256>>>(dA, Mbig, K, K, nullptr, expA, dAs);
            CK(cudaEventRecord(e1));
            if (variant == 0) k_ozaki_gemm<<<(Mbig / TM) * (N / TN), THREADS, smem>>>(dAs, dBs, Mbig, N, K, tri, expA, dexpB, dC, N, dfail);
            else if (variant == 1) k_ozaki_gemm_v1<<<(Mbig / TM) * (N / TN), V1_THREADS, NSTAGE * smem>>>(dAs, dBs, Mbig, N, K, tri, expA, dexpB, dC, N, dfail);
            else if (variant == 2) k_ozaki_gemm_v2<<<(Mbig / TM) * (N / TN), V2_THREADS, smem_v2>>>(mapA, mapB, mapA8, mapB8, Mbig, N, K, tri, dscale, dC, N, dfail);
            else k_ozaki_gemm_v3<256, 2><<<(Mbig / TM) * (N / 256), V3_THREADS, smem_v3>>>(m3A64, m3B64, m3A32, m3B32, Mbig, N, K, tri, dscale, dC, N, dfail);
            CK(cudaEventRecord(e2));
            CK(cudaDeviceSynchronize());
        }
        int hfail = 0; CK(cudaMemcpy(&hfail, dfail, 4, cudaMemcpyDeviceToHost));
        if (hfail) { printf("MMA wait timed out\n"); return 1; }
        float ms_s, ms_g; CK(cudaEventElapsedTime(&ms_s, e0, e1)); CK(cudaEventElapsedTime(&ms_g, e1, e2));
        const double flop = 2.0 * Mbig * (double)N * K * (tri ? 0.5 + 0.5 * TN / N : 1.0);
        printf("v%d tri=%d M=%d: slicing A %.3f ms, slice-product GEMM %.3f ms = %.1f fp64-equivalent TFLOP/s (%d slices, %d int8 MMAs per k-step)\n",
               variant, tri, Mbig, ms_s, ms_g, flop / (ms_g * 1e-3) / 1e12, S, S * (S + 1) / 2);
        // accuracy against a long-double CPU product on the first Mcheck rows
        std::vector<double> hC((size_t)Mcheck * N);
        CK(cudaMemcpy(hC.data(), dC, hC.size() * 8, cudaMemcpyDeviceToHost));
        double max_err = 0.0, max_ref = 0.0, max_rel_scale = 0.0;
        for (int m = 0; m < Mcheck; ++m)
            for (int n = 0; n < N; ++n) {
                long double s = 0.0L, sabs = 0.0L;
                for (int k = 0; k < K; ++k) {
                    const long double p = (long double)hA[(size_t)m * K + k] * hB[(size_t)n * K + k];
                    s += p; sabs += fabsl(p);
                }
                const double err = fabs((double)(hC[(size_t)m * N + n] - s));
                max_err = fmax(max_err, err); max_ref = fmax(max_ref, fabs((double)s));
                if (sabs > 0) max_rel_scale = fmax(max_rel_scale, err / (double)sabs);
            }
        printf("  accuracy on %d rows: max abs err %.3e (max |C| %.3e), max err / sum|a b| %.3e (fp64 dot product bound ~ %.1e)\n",
               Mcheck, max_err, max_ref, max_rel_scale, 1024 * 1.1e-16);
    }
    // ---- co-residency experiment: does the int8 tensor path run beside a saturated fp64 pipe? ----------------------------------
    {
        cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
        int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
        const int grid = (Mbig / TM) * (N / TN);
        const int iters = 30000;
        // both kernels must ask for the same shared-memory carve-out, or the SM is drained between them
        CK(cudaFuncSetAttribute(k_dfma_spin, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        CK(cudaFuncSetAttribute(k_ozaki_gemm_v2, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        auto gemm = [&](cudaStream_t st) { k_ozaki_gemm_v2<<<grid, V2_THREADS, smem_v2, st>>>(mapA, mapB, mapA8, mapB8, Mbig, N, K, 0, dscale, dC, N, dfail); };
        auto spin = [&](cudaStream_t st) { k_dfma_spin<<<sms, 256, 0, st>>>(iters, dC); };
        auto timed = [&](int which) {
            CK(cudaDeviceSynchronize());
            cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
            CK(cudaEventRecord(a, 0));
            CK(cudaStreamWaitEvent(s1, a, 0)); CK(cudaStreamWaitEvent(s2, a, 0));
            if (which & 1) spin(s1);
            if (which & 2) gemm(s2);
            cudaEvent_t e1_, e2_; CK(cudaEventCreate(&e1_)); CK(cudaEventCreate(&e2_));
            CK(cudaEventRecord(e1_, s1)); CK(cudaEventRecord(e2_, s2));
            CK(cudaStreamWaitEvent(0, e1_, 0)); CK(cudaStreamWaitEvent(0, e2_, 0));
            CK(cudaEventRecord(b, 0));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, a, b));
            return ms;
        };
        {   // harness check: two independent one-warp-per-SM sleep kernels on the two streams must overlap
            unsigned long long init0[4] = {~0ull, 0ull, ~0ull, 0ull}, got0[4];
            CK(cudaMemcpyToSymbol(g_t, init0, sizeof(init0)));
            CK(cudaDeviceSynchronize());
            k_sleep<<<sms, 32, 0, s1>>>(0, 300000);
            k_sleep<<<sms, 32, 0, s2>>>(2, 300000);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpyFromSymbol(got0, g_t, sizeof(got0)));
            const unsigned long long z0 = got0[0] < got0[2] ? got0[0] : got0[2];
            printf("harness check (two sleep kernels): [%.1f, %.1f] and [%.1f, %.1f] us\n", (got0[0] - z0) * 1e-3, (got0[1] - z0) * 1e-3,
                   (got0[2] - z0) * 1e-3, (got0[3] - z0) * 1e-3);
        }
        timed(3);
        const float t_spin = timed(1), t_gemm = timed(2);
        unsigned long long init[4] = {~0ull, 0ull, ~0ull, 0ull}, got[4];
        CK(cudaMemcpyToSymbol(g_t, init, sizeof(init)));
        const float t_both = timed(3);
        CK(cudaMemcpyFromSymbol(got, g_t, sizeof(got)));
        printf("device timeline of the concurrent run (us from the first start): DFMA kernel [%.1f, %.1f], GEMM [%.1f, %.1f]\n",
               (got[0] - (got[0] < got[2] ? got[0] : got[2])) * 1e-3, (got[1] - (got[0] < got[2] ? got[0] : got[2])) * 1e-3,
               (got[2] - (got[0] < got[2] ? got[0] : got[2])) * 1e-3, (got[3] - (got[0] < got[2] ? got[0] : got[2])) * 1e-3);
        printf("co-residency: DFMA kernel alone %.3f ms, slice-product GEMM (v2, full K) alone %.3f ms, both on two streams %.3f ms "
               "(serial would be %.3f ms)\n", t_spin, t_gemm, t_both, t_spin + t_gemm);
    }
    return 0;
}
