// fp64 peak micro-benchmarks for B200 (sm_100a): DFMA, DMMA m8n8k4, DMMA m16n8k16,
// and an interleaved DFMA+DMMA kernel (do the two share one pipe?).
// Build: nvcc -cudart shared -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp64_peak tools/fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void __launch_bounds__(256) k_dfma(double* out, int iters, double a, double b) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma16816(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}
__device__ __forceinline__ void dmma1684(double* c, const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                 : "+d"(c[0]), "+d"(c[1]), "+d"(c[2]), "+d"(c[3])
                 : "d"(a[0]), "d"(a[1]), "d"(b[0]));
}

__global__ void __launch_bounds__(256) k_dmma884(double* out, int iters, double a, double b) {
    double c[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) dmma884(c[2 * i], c[2 * i + 1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma16816(double* out, int iters, double a0, double b0) {
    double c[16], a[8], b[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + i * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 + i * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma16816(c + 4 * i, a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

__global__ void __launch_bounds__(256) k_dmma1684(double* out, int iters, double a0, double b0) {
    double c[16], a[2], b[1];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
    a[0] = a0; a[1] = a0 * 1.0000001; b[0] = b0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) dmma1684(c + 4 * i, a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
    if (s == 12345.678) out[0] = s;
}

// interleaved: per iteration 8 DMMA884 (8*512 flop/warp) + 8*? DFMA per thread
__global__ void __launch_bounds__(256) k_mixed(double* out, int iters, double a, double b) {
    double c[16], acc[8];
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            dmma884(c[2 * i], c[2 * i + 1], a, b);
            acc[i] = fma(acc[i], a, b);
            acc[i] = fma(acc[i], a, b);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += c[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 12345.678) out[0] = s;
}

// fp64 exp / division throughput (libdevice), for sizing the score kernel
__global__ void __launch_bounds__(256) k_exp(double* out, int iters, double a) {
    double acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = -1.0 - threadIdx.x * 1e-3 - i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = exp(acc[i]) - a;
    }
    double s = acc[0] + acc[1] + acc[2] + acc[3];
    if (s == 12345.678) out[0] = s;
}
__global__ void __launch_bounds__(256) k_div(double* out, int iters, double a) {
    double acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = a / (1.0 + acc[i]);
    }
    double s = acc[0] + acc[1] + acc[2] + acc[3];
    if (s == 12345.678) out[0] = s;
}

template <typename F>
float time_ms(F f, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(e0);
        f();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    double* out; CK(cudaMalloc(&out, 8));
    const int iters = 20000;
    for (int bps : {1, 2, 4, 8}) {
        int grid = sms * bps;
        double nthreads = (double)grid * 256;
        float ms;
        ms = time_ms([&] { k_dfma<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d dfma       %8.3f ms  %7.2f TFLOP/s\n", bps, ms, nthreads * iters * 8 * 2 / ms / 1e9);
        ms = time_ms([&] { k_dmma884<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d dmma884    %8.3f ms  %7.2f TFLOP/s\n", bps, ms, nthreads / 32 * iters * 8 * 512.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma1684<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d dmma1684   %8.3f ms  %7.2f TFLOP/s\n", bps, ms, nthreads / 32 * iters * 4 * 1024.0 / ms / 1e9);
        ms = time_ms([&] { k_dmma16816<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d dmma16816  %8.3f ms  %7.2f TFLOP/s\n", bps, ms, nthreads / 32 * iters * 4 * 4096.0 / ms / 1e9);
        ms = time_ms([&] { k_mixed<<<grid, 256>>>(out, iters, 1.0000001, 1e-9); }, 5);
        printf("bps %d mixed      %8.3f ms  dmma %7.2f + dfma %7.2f TFLOP/s\n", bps, ms,
               nthreads / 32 * iters * 8 * 512.0 / ms / 1e9, nthreads * iters * 16 * 2 / ms / 1e9);
        ms = time_ms([&] { k_exp<<<grid, 256>>>(out, iters / 10, 0.5); }, 3);
        printf("bps %d exp        %8.3f ms  %7.2f Gexp/s\n", bps, ms, nthreads * (iters / 10) * 4 / ms / 1e6);
        ms = time_ms([&] { k_div<<<grid, 256>>>(out, iters / 10, 0.5); }, 3);
        printf("bps %d div        %8.3f ms  %7.2f Gdiv/s\n", bps, ms, nthreads * (iters / 10) * 4 / ms / 1e6);
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
