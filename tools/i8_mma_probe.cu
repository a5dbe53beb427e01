// Exploration for the next round: is the int8 tensor path of sm_100a (tcgen05.mma kind::i8, int32
// accumulators in TMEM) usable as the engine of an fp64-by-integer-slices GEMM?  This stand-alone probe
//   (1) checks one 128 x N x K int8 GEMM tile issued with hand-built shared-memory / instruction descriptors against the CPU,
//   (2) measures the sustained MMA rate of one CTA per SM re-issuing m128 n256 k32 on resident operands.
// Operands are staged with plain stores in the K-major no-swizzle ("interleave") canonical layout: 16-byte K-chunk c of row r
// at  c * ROWS * 16 + r * 16  (core matrix = 8 rows x 16 B contiguous; SBO = 128 B between 8-row groups, LBO = ROWS * 16 B
// between K-chunks).  Every wait is bounded (trap after ~1 s) so that a wrong descriptor cannot hang the box.
// Build: nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/i8_mma_probe tools/i8_mma_probe.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;            // descriptor version (Blackwell)
    return d;                          // base_offset 0, lbo_mode 0, layout_type 0 = no swizzle
}

__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
    // c_format [4,6) = 2 (S32), a_format [7,10) = 1 (signed 8 bit), b_format [10,13) = 1, K-major both, n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ void mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
    for (long it = 0; it < 20000000L; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return true;
    }
    return false;
}

// A: 128 x K (row-major int8), B: N x K (row-major int8), C: 128 x N int32 = A B^T.  reps > 1: re-issue the whole K loop
// (accumulating) to measure the MMA rate; C is then reps x the product.
template <int N>
__global__ void __launch_bounds__(128) k_probe(const int8_t* __restrict__ A, const int8_t* __restrict__ B, int K, int reps,
                                               int32_t* __restrict__ C, long long* __restrict__ cycles, int* __restrict__ fail) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* sA = smem;                         // (K/16) chunks x 128 rows x 16 B
    uint8_t* sB = smem + (size_t)128 * K;       // (K/16) chunks x N rows x 16 B
    const int chunks = K / 16;
    for (int o = tid; o < 128 * chunks; o += 128) {
        const int r = o % 128, c = o / 128;
        *reinterpret_cast<int4*>(sA + (size_t)c * 128 * 16 + r * 16) = *reinterpret_cast<const int4*>(A + (size_t)r * K + c * 16);
    }
    for (int o = tid; o < N * chunks; o += 128) {
        const int r = o % N, c = o / N;
        *reinterpret_cast<int4*>(sB + (size_t)c * N * 16 + r * 16) = *reinterpret_cast<const int4*>(B + (size_t)r * K + c * 16);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(N));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy stores -> visible to the tensor core
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base;
    const uint32_t idesc = make_idesc_i8(128, N);
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int rep = 0; rep < reps; ++rep)
            for (int k = 0; k < K / 32; ++k) {
                const uint64_t ad = make_desc(a0 + (uint32_t)k * 2 * 128 * 16, 128 * 16, 128);
                const uint64_t bd = make_desc(b0 + (uint32_t)k * 2 * N * 16, N * 16, 128);
                mma_i8(tmem, ad, bd, idesc, (rep | k) != 0);
            }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    const bool ok = mbar_wait_bounded(smem_u32(&bar), 0);
    long long t1 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;");
    if (!ok) { if (tid == 0) atomicExch(fail, 1); }
    else if (C) {
        int32_t* Cb = C + (size_t)blockIdx.x * 128 * N;
        for (int c0 = 0; c0 < N; c0 += 32) {
            uint32_t v[32];
            const uint32_t addr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                         "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                           "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                           "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                         : "r"(addr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 32; ++j) Cb[(size_t)(warp * 32 + lane) * N + c0 + j] = (int32_t)v[j];
        }
    }
    if (tid == 0 && cycles) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(N));
}

template <int N>
static int run(int K, int reps, int grid, bool check) {
    std::vector<int8_t> hA((size_t)128 * K), hB((size_t)N * K);
    srand(1234 + N + K);
    for (auto& x : hA) x = (int8_t)(rand() % 255 - 127);
    for (auto& x : hB) x = (int8_t)(rand() % 255 - 127);
    int8_t *dA, *dB; int32_t* dC; long long* dcyc; int* dfail;
    CK(cudaMalloc(&dA, hA.size())); CK(cudaMalloc(&dB, hB.size()));
    CK(cudaMalloc(&dC, (size_t)grid * 128 * N * 4)); CK(cudaMalloc(&dcyc, grid * 8)); CK(cudaMalloc(&dfail, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size(), cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size(), cudaMemcpyHostToDevice));
    CK(cudaMemset(dfail, 0, 4));
    const size_t smem = (size_t)(128 + N) * K;
    CK(cudaFuncSetAttribute(k_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_probe<N><<<grid, 128, smem>>>(dA, dB, K, reps, check ? dC : nullptr, dcyc, dfail);     // warm-up
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    k_probe<N><<<grid, 128, smem>>>(dA, dB, K, reps, check ? dC : nullptr, dcyc, dfail);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    int hfail = 0; CK(cudaMemcpy(&hfail, dfail, 4, cudaMemcpyDeviceToHost));
    if (hfail) { printf("N=%d K=%d: mbarrier wait timed out (MMA never completed)\n", N, K); return 1; }
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> cyc(grid); CK(cudaMemcpy(cyc.data(), dcyc, grid * 8, cudaMemcpyDeviceToHost));
    const double ops = 2.0 * 128 * N * K * (double)reps * grid;
    printf("N=%d K=%d reps=%d grid=%d: %.3f ms, %.1f TOP/s (int8 dense), %lld cycles in the issue+wait section of CTA 0 (%.1f clk per m128n%dk32)\n",
           N, K, reps, grid, ms, ops / (ms * 1e-3) / 1e12, cyc[0], (double)cyc[0] / ((double)reps * (K / 32)), N);
    int bad = 0;
    if (check) {
        std::vector<int32_t> hC((size_t)128 * N);
        CK(cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost));
        for (int m = 0; m < 128; ++m)
            for (int n = 0; n < N; ++n) {
                long long s = 0;
                for (int k = 0; k < K; ++k) s += (int)hA[(size_t)m * K + k] * (int)hB[(size_t)n * K + k];
                s *= reps;
                if ((int32_t)s != hC[(size_t)m * N + n]) { if (bad < 5) printf("  mismatch at (%d,%d): got %d want %lld\n", m, n, hC[(size_t)m * N + n], s); ++bad; }
            }
        printf("  check vs CPU: %s (%d mismatches of %d)\n", bad ? "FAIL" : "exact", bad, 128 * N);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC); cudaFree(dcyc); cudaFree(dfail);
    return bad != 0;
}

// The MMA sequence of the slice-product GEMM (tools/ozaki_gemm_proto.cu) on RESIDENT operands: 8 A planes and 8 B planes of one
// k-step (128 x 32 B each), 4 accumulators of 128 columns, the 10 + 26 (i, j) pairs of the two passes, no TMA traffic.
__global__ void __launch_bounds__(128) k_probe_pattern(int reps, long long* __restrict__ cycles, int* __restrict__ fail) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int o = tid; o < 16 * 4096 / 16; o += 128) reinterpret_cast<int4*>(smem)[o] = make_int4(o * 0x01010101, o, ~o, o * 7);
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    if (warp == 0) {
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = tmem_base, idesc = make_idesc_i8(128, 128);
    long long t0 = clock64();
    if (tid == 0) {
        const uint32_t a0 = smem_u32(smem), b0 = a0 + 8 * 4096;
        for (int rep = 0; rep < reps; ++rep)
            for (int pass = 0; pass < 2; ++pass)
                for (int t = 2 + 4 * pass; t <= 5 + 4 * pass; ++t)
                    for (int i = 1; i < t; ++i) {
                        const int j = t - i;
                        if (i > 8 || j > 8) continue;
                        mma_i8(tmem + (uint32_t)(t - 2 - 4 * pass) * 128, make_desc(a0 + (i - 1) * 4096, 128 * 16, 128),
                               make_desc(b0 + (j - 1) * 4096, 128 * 16, 128), idesc, 1);
                    }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    const bool ok = mbar_wait_bounded(smem_u32(&bar), 0);
    long long t1 = clock64();
    if (!ok && tid == 0) atomicExch(fail, 1);
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

static void run_pattern(int sms) {
    long long* dcyc; int* dfail;
    CK(cudaMalloc(&dcyc, sms * 8)); CK(cudaMalloc(&dfail, 4)); CK(cudaMemset(dfail, 0, 4));
    CK(cudaFuncSetAttribute(k_probe_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
    const int reps = 512;
    k_probe_pattern<<<sms, 128, 65536>>>(reps, dcyc, dfail);
    CK(cudaDeviceSynchronize());
    k_probe_pattern<<<sms, 128, 65536>>>(reps, dcyc, dfail);
    CK(cudaDeviceSynchronize());
    int hfail = 0; CK(cudaMemcpy(&hfail, dfail, 4, cudaMemcpyDeviceToHost));
    long long c0 = 0; CK(cudaMemcpy(&c0, dcyc, 8, cudaMemcpyDeviceToHost));
    printf("slice-product MMA pattern on resident operands (16 planes, 4 accumulators, 36 MMAs per k-step): %s, %.1f clk per m128n128k32\n",
           hfail ? "TIMED OUT" : "ok", (double)c0 / (reps * 36.0));
}

int main() {
    int sms = 0; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    int rc = 0;
    rc |= run<128>(64, 1, 1, true);
    rc |= run<128>(256, 1, 1, true);
    rc |= run<256>(256, 3, 2, true);
    if (rc) { printf("correctness failed: no throughput run\n"); return 1; }
    run<256>(256, 2048, sms, false);
    run<128>(256, 2048, sms, false);
    run<64>(256, 2048, sms, false);
    run_pattern(sms);
    return 0;
}
