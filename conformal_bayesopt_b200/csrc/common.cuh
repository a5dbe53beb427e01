// common.cuh -- device helpers shared by the libcbo kernels (sm_100a, fp64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace cbo {

// Kernels launched by the library since it was loaded (cbo_launch_count): every launch site is prefixed with CBO_LAUNCHED().
inline std::atomic<long long>& launch_counter() {
    static std::atomic<long long> c{0};
    return c;
}
#define CBO_LAUNCHED() ((void)cbo::launch_counter().fetch_add(1, std::memory_order_relaxed))

// ---- host: per-DEVICE state (one process may drive several GPUs from several threads; cbo.h promises re-entrancy) ------
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return -1;
    return dev;
}
// Opt-in to more than 48 KB of dynamic shared memory: cudaFuncAttributeMaxDynamicSharedMemorySize is per (function, device),
// so the "already configured" size is cached per device.  Concurrent callers may both set the attribute: harmless.
struct SmemOptIn {
    std::atomic<size_t> done[kMaxDevices];
    std::atomic<int> aux[kMaxDevices];          // optional per-device value computed together with the opt-in (occupancy)
    template <typename F>
    cudaError_t ensure(F* fn, size_t smem, bool* fresh = nullptr) {
        const int dev = current_device_slot();
        if (fresh) *fresh = false;
        if (dev >= 0 && smem <= done[dev].load(std::memory_order_acquire)) return cudaSuccess;
        if (dev >= 0 && smem <= 48 * 1024 && done[dev].load(std::memory_order_acquire) == 0 && !fresh) {
            done[dev].store(48 * 1024, std::memory_order_release);     // the default limit needs no opt-in
            return cudaSuccess;
        }
        cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        if (fresh) *fresh = true;
        if (dev >= 0) done[dev].store(smem, std::memory_order_release);
        return cudaSuccess;
    }
};

constexpr int kWarp = 32;
constexpr double kLog2Pi = 1.8378770664093453;   // log(2*pi)
constexpr double kSqrt5 = 2.23606797749979;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ int warp_sum_int(int v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Warp sums of FOUR per-lane values at once by transposition: 6 shuffles instead of 20.  Every lane returns the total of
// value k = (lane >> 3) & 3 (lanes 8k .. 8k+7 hold the sum of v[k] over the warp).
template <typename T>
__device__ __forceinline__ T warp_sum4(const T (&v)[4], int lane) {
    const bool up = lane & 16, mid = lane & 8;
    T a0 = up ? v[2] : v[0], a1 = up ? v[3] : v[1];
    const T s0 = up ? v[0] : v[2], s1 = up ? v[1] : v[3];
    a0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    a1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    T b = mid ? a1 : a0;
    const T s = mid ? a0 : a1;
    b += __shfl_xor_sync(0xffffffffu, s, 8);
    b += __shfl_xor_sync(0xffffffffu, b, 4);
    b += __shfl_xor_sync(0xffffffffu, b, 2);
    b += __shfl_xor_sync(0xffffffffu, b, 1);
    return b;
}

// ---- cp.async (LDGSTS) 16-byte copies -------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

// ---- mbarrier (shared::cta) -------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return static_cast<unsigned>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
    unsigned ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}
// the mbarrier receives one arrival from this thread when all of its prior cp.async copies have landed
__device__ __forceinline__ void cp_async_mbar_arrive(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- fp64 tensor-core MMA (DMMA.8x8x4 in SASS) ----------------------------------------------
// D(8x8) += A(8x4, row) * B(4x8, col).  Lane l holds A[l/4][l%4], B[l%4][l/4], C[l/4][2*(l%4)+{0,1}].
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// ---- reduced-cost fp64 transcendental pieces for the soft-rank inner loop -------------------
// exp(x) for x in [-745, 0], relative error <= 4e-11 (64-entry 2^(j/64) table + cubic).  The soft
// rank needs rho to ~1e-9; this costs 7 fp64-pipe instructions instead of the
// ~26 of libdevice exp (profiles/r01_fp64_peak.log).  `tab` is a 64-entry table of 2^(j/64).
__device__ __forceinline__ double fast_exp_neg(double x, const double* __restrict__ tab) {
    constexpr double kMagic = 6755399441055744.0;            // 1.5 * 2^52: integer lands in the low word
    constexpr double kL2e64 = 92.332482616893657;            // 64 / ln 2
    constexpr double kLn2o64 = 0.010830424696249145;         // ln 2 / 64
    double t = fma(x, kL2e64, kMagic);
    int n = __double2loint(t);                               // round(x * 64 / ln2), n <= 0
    double nf = t - kMagic;
    double y = fma(nf, -kLn2o64, x);                         // |y| <= ln2/128
    double p = fma(y, 1.0 / 6.0, 0.5);
    p = fma(p, y, 1.0);
    p = p * y;                                               // y + y^2/2 + y^3/6
    double tj = tab[n & 63];
    double e = fma(tj, p, tj);
    int k = n >> 6;                                          // floor(n / 64)
    int hi = __double2hiint(e) + (k << 20);                  // scale by 2^k (result stays normal for x > -700)
    e = __hiloint2double(hi, __double2loint(e));
    return (k < -1000) ? 0.0 : e;
}

// 1 / d for d in [1, 2]: MUFU.RCP64H seed (rel. err 2^-23) + one Newton step -> ~2^-45.
__device__ __forceinline__ double fast_rcp(double d) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(d));
    double e = fma(-d, y, 1.0);
    return fma(y, e, y);
}

// sigmoid(z) with absolute error ~1e-11; saturates exactly to 0/1 for |z| > 40.
__device__ __forceinline__ double fast_sigmoid(double z, const double* __restrict__ tab) {
    double az = fabs(z);
    double e = fast_exp_neg(-fmin(az, 64.0), tab);
    double r = fast_rcp(1.0 + e);
    return (z >= 0.0) ? r : e * r;
}

// 1 / v to ~1 ulp for normal v > 0: MUFU.RCP64H seed + two Newton steps (6 instructions instead of the ~40 of a
// full IEEE division; used once per (candidate, train point)).
__device__ __forceinline__ double rcp_full(double v) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;\n" : "=d"(y) : "d"(v));
    double e = fma(-v, y, 1.0);
    y = fma(y, e, y);
    e = fma(-v, y, 1.0);
    return fma(y, e, y);
}

// log(v) for normal v > 0, absolute error < 1e-13: v = 2^k m, m in [0.75, 1.5), log m = 2 atanh((m-1)/(m+1)) by an
// 8-term odd series (|s| <= 0.2).  ~30 instructions instead of the ~80 of libdevice log; the scores need log(var) to
// 1e-11 (the error is amplified by 1/temp in the soft rank).
__device__ __forceinline__ double log_fast(double v) {
    int hi = __double2hiint(v), lo = __double2loint(v);
    int k = (hi >> 20) - 1023;
    int mh = (hi & 0x000fffff) | 0x3ff00000;                 // m in [1, 2)
    if (mh >= 0x3ff80000) { mh -= 0x00100000; k += 1; }      // m >= 1.5 -> m / 2
    double m = __hiloint2double(mh, lo);
    double s = (m - 1.0) * rcp_full(m + 1.0);
    double s2 = s * s;
    double pl = fma(s2, 1.0 / 17.0, 1.0 / 15.0);
    pl = fma(pl, s2, 1.0 / 13.0);
    pl = fma(pl, s2, 1.0 / 11.0);
    pl = fma(pl, s2, 1.0 / 9.0);
    pl = fma(pl, s2, 1.0 / 7.0);
    pl = fma(pl, s2, 1.0 / 5.0);
    pl = fma(pl, s2, 1.0 / 3.0);
    pl = fma(pl, s2, 1.0);
    return fma(double(k), 0.6931471805599453, 2.0 * s * pl);
}

// accurate sigmoid for the per-(candidate, grid point) epilogue (not in the inner loop)
__device__ __forceinline__ double sigmoid_acc(double z) {
    if (z >= 0.0) return 1.0 / (1.0 + exp(-z));
    double e = exp(z);
    return e / (1.0 + e);
}

}  // namespace cbo
